mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_n1_20steps.json 2> gpurun_out/r02_bench_c4_n1_20steps.err; tail -1 gpurun_out/r02_bench_c4_n1_20steps.err
python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -1 gpurun_out/r02_bench_c4_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c4_n1_reference_arm.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_c4_n1_20steps_b.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_c4_n1_20steps_c.json 2>/dev/null
