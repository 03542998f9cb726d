mkdir -p gpurun_out
rm -f gpurun_out/r02g.jsonl
python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02g_tests.log; cat gpurun_out/r02g_tests.log
for b in 148 200 25; do for s in 20 2000; do python tools/ensemble_bench.py --bins $b --steps $s >> gpurun_out/r02g.jsonl 2>&1; done; done
python tools/ensemble_bench.py --bins 100 --walkers 256 --n-l 5 --n-pts 30 --steps 2000 >> gpurun_out/r02g.jsonl 2>&1
