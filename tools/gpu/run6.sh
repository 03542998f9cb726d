mkdir -p gpurun_out
rm -f gpurun_out/r02e.jsonl
for treg in 1 0; do for b in 74 148 296; do echo "treg $treg bins $b" >> gpurun_out/r02e.jsonl; MDA_ENS_TREG=$treg python tools/ensemble_bench.py --bins $b --steps 1000 >> gpurun_out/r02e.jsonl 2>&1; done; done
for b in 148 296 444 592; do echo "quad bins $b" >> gpurun_out/r02e.jsonl; MDA_ENS_IMPL=quad python tools/ensemble_bench.py --bins $b --steps 1000 >> gpurun_out/r02e.jsonl 2>&1; done
