mkdir -p gpurun_out
python tools/ensemble_bench.py --bins 200 --steps 20 --profile 1 > gpurun_out/r02h.jsonl 2>&1
python tools/ensemble_bench.py --bins 148 --steps 20 --profile 1 >> gpurun_out/r02h.jsonl 2>&1
