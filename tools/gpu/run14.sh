mkdir -p gpurun_out
for p in 0 1; do MDA_BATCH_IMPL=dfma MDA_BATCH_PERSISTENT=$p python tools/sweep.py --shapes big --tiles --launches 40 > gpurun_out/r02k_tiles_p$p.jsonl 2>&1; done
