mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err
for b in 25 148 200; do for s in 20 2000; do python tools/ensemble_bench.py --bins $b --steps $s --profile 1 >> gpurun_out/r02_qws_prof.jsonl 2>&1; python tools/ensemble_bench.py --bins $b --steps $s >> gpurun_out/r02_qws_noprof.jsonl 2>&1; done; done
