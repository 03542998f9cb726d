mkdir -p gpurun_out
python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02b_gputests.log
rm -f gpurun_out/r02b_qws_prof.jsonl gpurun_out/r02b_qws_noprof.jsonl
for b in 25 148 200 296; do for s in 20 2000; do python tools/ensemble_bench.py --bins $b --steps $s >> gpurun_out/r02b_qws_noprof.jsonl 2>&1; done; done
for b in 148 200; do python tools/ensemble_bench.py --bins $b --steps 2000 --profile 1 >> gpurun_out/r02b_qws_prof.jsonl 2>&1; done
python tools/ensemble_bench.py --bins 100 --walkers 256 --n-l 5 --n-pts 30 --steps 2000 >> gpurun_out/r02b_qws_noprof.jsonl 2>&1
