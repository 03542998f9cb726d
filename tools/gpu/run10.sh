mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_n1_20steps.json 2> gpurun_out/r02_bench_c4_n1_20steps.err; tail -2 gpurun_out/r02_bench_c4_n1_20steps.err
python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -2 gpurun_out/r02_bench_c4_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c4_n1_reference_arm.json 2> gpurun_out/r02_bench_ref.err
python tools/gpu/hbm_write_peak.py > gpurun_out/r02_hbm_write_peak.json 2>&1
