mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests.log
tail -3 gpurun_out/gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err; tail -3 gpurun_out/r02_bench20.err
python bench.py > gpurun_out/r02_bench2000.json 2> gpurun_out/r02_bench2000.err; tail -3 gpurun_out/r02_bench2000.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
