mkdir -p gpurun_out
python -m pytest tests/test_fit_gpu.py tests/test_pipeline_gpu.py tests/test_autocorr.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02i_tests.log; tail -25 gpurun_out/r02i_tests.log
python tools/fit_bench.py > gpurun_out/r02_fit_bench.jsonl 2>&1; tail -3 gpurun_out/r02_fit_bench.jsonl
