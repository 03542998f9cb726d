mkdir -p gpurun_out
# C4 qws: ncu --set full of the 2000-step launch (cooperative launch) and of the driver's 20-step launch
bash tools/gpu/ncu_any.sh r02_ncu_c4_qws_2000 ensemble_qws 2 -- python tools/ensemble_bench.py --bins 200 --steps 2000
bash tools/gpu/ncu_any.sh r02_ncu_c4_qws_20 ensemble_qws 2 -- python tools/ensemble_bench.py --bins 200 --steps 20
# launch list of the bench command
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_for_launchlist.json 2> gpurun_out/r02_bench_for_launchlist.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_c4.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launchlist_ncu.log 2>&1
tail -2 gpurun_out/r02_launchlist_ncu.log
