mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_quad_gpu.py tests/test_sampler_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02d_tests.log
tail -3 gpurun_out/r02d_tests.log
for impl in split dfma; do MDA_BATCH_IMPL=$impl python tools/sweep.py --shapes c4 > gpurun_out/r02d_sweep_c4_$impl.jsonl 2>&1; MDA_BATCH_IMPL=$impl python tools/sweep.py --shapes cross > gpurun_out/r02d_sweep_cross_$impl.jsonl 2>&1; done
python tools/sweep.py --shapes c4 > gpurun_out/r02d_sweep_c4_default.jsonl 2>&1
