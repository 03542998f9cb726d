mkdir -p gpurun_out
rm -f gpurun_out/r02m.jsonl
for K in 10 9; do
  sed -i "s/constexpr int kSplit = TREG ? (L + 1 < [0-9]* ? L + 1 : [0-9]*) : 0;/constexpr int kSplit = TREG ? (L + 1 < $K ? L + 1 : $K) : 0;/" mimic_mda_b200/csrc/ensemble_qws_kernel.cu
  python -c "from mimic_mda_b200 import build; build.build()" > /dev/null 2>gpurun_out/r02m_build_$K.err || { echo "build failed K=$K" >> gpurun_out/r02m.jsonl; continue; }
  echo "kSplit $K" >> gpurun_out/r02m.jsonl
  for b in 148 200 25; do for s in 20 2000; do python tools/ensemble_bench.py --bins $b --steps $s >> gpurun_out/r02m.jsonl 2>&1; done; done
  python tools/ensemble_bench.py --bins 100 --walkers 256 --n-l 5 --n-pts 30 --steps 2000 >> gpurun_out/r02m.jsonl 2>&1
  python tools/ensemble_bench.py --bins 100 --walkers 512 --n-l 8 --n-pts 40 --steps 2000 >> gpurun_out/r02m.jsonl 2>&1
done
python -m pytest tests/test_ensemble_gpu.py tests/test_quad_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02m_tests.log; cat gpurun_out/r02m_tests.log
