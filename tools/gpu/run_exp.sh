mkdir -p gpurun_out
rm -f gpurun_out/exp.jsonl
timeout 900 python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py -m gpu -x -q > gpurun_out/exp_tests.log 2>&1; tail -3 gpurun_out/exp_tests.log
for rep in 1 2; do
  timeout 120 python tools/ensemble_bench.py --bins 148 --steps 2000 >> gpurun_out/exp.jsonl 2>&1
  timeout 120 python tools/ensemble_bench.py --bins 200 --steps 2000 >> gpurun_out/exp.jsonl 2>&1
done
timeout 120 python tools/ensemble_bench.py --bins 200 --steps 20 >> gpurun_out/exp.jsonl 2>&1
timeout 120 python tools/ensemble_bench.py --bins 100 --steps 2000 --walkers 256 --n-l 5 --n-pts 30 >> gpurun_out/exp.jsonl 2>&1
timeout 120 python tools/ensemble_bench.py --bins 100 --steps 2000 --n-l 8 --n-pts 40 >> gpurun_out/exp.jsonl 2>&1
timeout 120 python tools/ensemble_bench.py --bins 100 --steps 2000 --n-l 7 >> gpurun_out/exp.jsonl 2>&1
