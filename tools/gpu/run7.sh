mkdir -p gpurun_out
python tools/gpu/hbm_write_peak.py > gpurun_out/r02_hbm_write_peak.json 2>&1
rm -f gpurun_out/r02f.jsonl
for impl in qws quad; do for b in 296 592; do for c in 1 0; do echo "impl $impl bins $b chain $c" >> gpurun_out/r02f.jsonl; MDA_ENS_IMPL=$impl python tools/ensemble_bench.py --bins $b --steps 1000 --chain $c >> gpurun_out/r02f.jsonl 2>&1; done; done; done
MDA_E2E_TIMING=1 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_bench20b.json 2> gpurun_out/r02_bench20b.err
