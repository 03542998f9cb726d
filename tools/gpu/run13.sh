mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_sampler_gpu.py tests/test_fit_gpu.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02j_tests.log; cat gpurun_out/r02j_tests.log
for p in 0 1; do MDA_BATCH_PERSISTENT=$p python tools/sweep.py --shapes big > gpurun_out/r02j_sweep_big_p$p.jsonl 2>&1; done
python tools/sweep.py --shapes big > gpurun_out/r02j_sweep_big_default.jsonl 2>&1
for t in 64x4 128x2 128x4 256x2 64x2; do echo "tile $t" >> gpurun_out/r02j_tiles.txt; MDA_BATCH_PERSISTENT=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs --tile $t 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['batched_half_step']['roofline']['launch_us'])" >> gpurun_out/r02j_tiles.txt; done
