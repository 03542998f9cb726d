mkdir -p gpurun_out
rm -f gpurun_out/r02g.jsonl
python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02g_tests.log; cat gpurun_out/r02g_tests.log
for l in 8 4 6 9; do python tools/ensemble_bench.py --bins 100 --walkers 512 --n-l $l --n-pts 40 --steps 2000 >> gpurun_out/r02g.jsonl 2>&1; done
python tools/ensemble_bench.py --bins 200 --steps 2000 >> gpurun_out/r02g.jsonl 2>&1
