mkdir -p gpurun_out
bash tools/gpu/run_n.sh 8 20 r02_bench_c4_n8_20steps
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --no-configs > gpurun_out/r02_bench_c4_n8.json 2> gpurun_out/r02_bench_c4_n8.err
