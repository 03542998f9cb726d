mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gputests.log; cat gpurun_out/gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_n1_20steps.json 2> gpurun_out/r02_bench_c4_n1_20steps.err; tail -2 gpurun_out/r02_bench_c4_n1_20steps.err
