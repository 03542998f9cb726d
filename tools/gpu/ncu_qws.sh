mkdir -p gpurun_out
BINS=${1:-148}; STEPS=${2:-500}; TAG=${3:-r02_qws}
python tools/ensemble_bench.py --bins $BINS --steps $STEPS > gpurun_out/${TAG}_plain.json 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:ensemble_q -s 2 -c 1 -f -o gpurun_out/${TAG}_${BINS} python tools/ensemble_bench.py --bins $BINS --steps $STEPS > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
