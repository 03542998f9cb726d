mkdir -p gpurun_out
rm -f gpurun_out/r02n.jsonl
for s in 1 2 4 8 16 32 64; do python tools/ensemble_bench.py --bins 148 --steps $s >> gpurun_out/r02n.jsonl 2>&1; done
for s in 1 2 4 8 16 32 64; do python tools/ensemble_bench.py --bins 148 --steps $s --chain 0 >> gpurun_out/r02n.jsonl 2>&1; done
