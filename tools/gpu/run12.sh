mkdir -p gpurun_out
rm -f gpurun_out/r02q.jsonl
timeout 600 python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; tail -3 gpurun_out/r02q_tests.log
for rep in 1 2; do
  timeout 120 python tools/ensemble_bench.py --bins 148 --steps 2000 --chain 1 >> gpurun_out/r02q.jsonl 2>&1
  timeout 120 python tools/ensemble_bench.py --bins 200 --steps 2000 --chain 1 >> gpurun_out/r02q.jsonl 2>&1
  timeout 120 python tools/ensemble_bench.py --bins 200 --steps 20 --chain 1 >> gpurun_out/r02q.jsonl 2>&1
done
timeout 120 python tools/ensemble_bench.py --bins 100 --steps 2000 --chain 1 --walkers 256 --n-l 5 --n-pts 30 >> gpurun_out/r02q.jsonl 2>&1
timeout 120 python tools/ensemble_bench.py --bins 100 --steps 2000 --chain 1 --n-l 8 >> gpurun_out/r02q.jsonl 2>&1
python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
