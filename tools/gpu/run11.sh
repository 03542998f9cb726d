mkdir -p gpurun_out
python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py -m gpu -x -q > gpurun_out/r02i_tests.log 2>&1; tail -3 gpurun_out/r02i_tests.log
python tools/ensemble_bench.py --bins 148 --steps 2000 > gpurun_out/r02i.jsonl 2>&1
python tools/ensemble_bench.py --bins 200 --steps 2000 >> gpurun_out/r02i.jsonl 2>&1
python tools/ensemble_bench.py --bins 200 --steps 20 >> gpurun_out/r02i.jsonl 2>&1
python tools/ensemble_bench.py --bins 200 --steps 20 >> gpurun_out/r02i.jsonl 2>&1
python tools/ensemble_bench.py --bins 200 --steps 20 --n-l 8 >> gpurun_out/r02i.jsonl 2>&1
python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
