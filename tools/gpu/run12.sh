mkdir -p gpurun_out
rm -f gpurun_out/r02s.jsonl
timeout 600 python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py -m gpu -x -q > gpurun_out/r02s_tests.log 2>&1; tail -3 gpurun_out/r02s_tests.log
for bins in 148 200; do
for st in 2 10 20 20 40 100 2000; do
  timeout 120 python tools/ensemble_bench.py --bins $bins --steps $st --chain 1 >> gpurun_out/r02s.jsonl 2>&1
done
done
