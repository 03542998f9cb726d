mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gputests.log; cat gpurun_out/gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_n1_20steps.json 2> gpurun_out/r02_bench_c4_n1_20steps.err; tail -1 gpurun_out/r02_bench_c4_n1_20steps.err
python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -1 gpurun_out/r02_bench_c4_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c4_n1_reference_arm.json 2> gpurun_out/r02_bench_ref.err
bash tools/gpu/ncu_any.sh r02_ncu_c4_qws_2000 ensemble_qws 2 -- python tools/ensemble_bench.py --bins 200 --steps 2000
bash tools/gpu/ncu_any.sh r02_ncu_c4_qws_20 ensemble_qws 2 -- python tools/ensemble_bench.py --bins 200 --steps 20
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_c4.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launchlist_ncu.log 2>&1
