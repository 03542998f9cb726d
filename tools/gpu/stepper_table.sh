# The rows of DESIGN.md 4.2.1's table: ensemble steps of the device-resident sampler at the shapes quoted there.
mkdir -p gpurun_out
out=gpurun_out/r02_stepper_table.jsonl
rm -f $out
for bins in 25 148 200 296 592; do
  timeout 300 python tools/ensemble_bench.py --bins $bins --steps 2000 >> $out 2>&1
done
timeout 300 python tools/ensemble_bench.py --bins 200 --steps 20 >> $out 2>&1
timeout 300 python tools/ensemble_bench.py --bins 100 --steps 2000 --n-l 8 --n-pts 40 >> $out 2>&1
timeout 300 python tools/ensemble_bench.py --bins 100 --steps 2000 --walkers 256 --n-l 5 --n-pts 30 >> $out 2>&1
timeout 300 python tools/ensemble_bench.py --bins 100 --steps 2000 >> $out 2>&1
timeout 300 python tools/ensemble_bench.py --bins 148 --steps 2000 --chain 0 >> $out 2>&1
MDA_ENS_IMPL=direct timeout 300 python tools/ensemble_bench.py --bins 200 --steps 500 >> $out 2>&1
