# usage: run_n.sh N STEPS TAG
N=$1; STEPS=$2; TAG=$3
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --steps $STEPS --warmup 3 > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2>> gpurun_out/${TAG}.err
fi
tail -2 gpurun_out/${TAG}.err
