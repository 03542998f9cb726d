mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi "0x10de" $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node)" >> gpurun_out/r02_topo.txt; fi; done
lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/r02_topo.txt
bash tools/gpu/run_n.sh 8 20 r02_bench_c4_n8_20steps
MDA_BENCH_NUMA=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 --no-configs --no-strong > gpurun_out/r02_bench_c4_n8_20steps_nonuma.json 2> gpurun_out/r02_bench_c4_n8_nonuma.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --no-configs > gpurun_out/r02_bench_c4_n8.json 2> gpurun_out/r02_bench_c4_n8.err
