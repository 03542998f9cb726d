mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/gputests.log
tail -5 gpurun_out/gputests.log
