mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gputests.log; cat gpurun_out/gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_n1_20steps.json 2> gpurun_out/r02_bench_c4_n1_20steps.err; tail -1 gpurun_out/r02_bench_c4_n1_20steps.err
python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -1 gpurun_out/r02_bench_c4_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c4_n1_reference_arm.json 2> gpurun_out/r02_bench_ref.err
