mkdir -p gpurun_out
rm -f gpurun_out/r02c.jsonl
for dbg in 0 16 32 48; do echo "dbg $dbg" >> gpurun_out/r02c.jsonl; MDA_QWS_DEBUG=$dbg python tools/ensemble_bench.py --bins 148 --steps 2000 --profile 1 >> gpurun_out/r02c.jsonl 2>&1;  MDA_QWS_DEBUG=$dbg python tools/ensemble_bench.py --bins 148 --steps 2000 >> gpurun_out/r02c.jsonl 2>&1; done
echo "chain 0" >> gpurun_out/r02c.jsonl;  python tools/ensemble_bench.py --bins 148 --steps 2000 --profile 1 --chain 0 >> gpurun_out/r02c.jsonl 2>&1
echo "chain 0 dbg 1" >> gpurun_out/r02c.jsonl;  MDA_QWS_DEBUG=1 python tools/ensemble_bench.py --bins 148 --steps 2000 --profile 1 --chain 0 >> gpurun_out/r02c.jsonl 2>&1
