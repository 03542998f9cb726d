# fp64 peak micro-benchmark (csrc/peaks.cu via mdaMeasureFp64Peak) with the clocks sampled while it runs
mkdir -p gpurun_out
( for i in 1 2 3 4 5 6; do nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader; sleep 0.3; done ) > gpurun_out/r02_fp64_peak_clocks.txt &
python - > gpurun_out/r02_fp64_peak.json <<'PY'
import ctypes as ct, json, sys
sys.path.insert(0, ".")
from mimic_mda_b200 import chisq
lib = chisq.prepare_shared_object()
vals = []
for _ in range(3):
    p = ct.c_double()
    chisq.check(lib, lib.mdaMeasureFp64Peak(5, ct.byref(p)))
    vals.append(p.value)
print(json.dumps({"fp64_peak_tflops": vals, "how": "mdaMeasureFp64Peak(5): DFMA chains with one re-used operand, all SMs (csrc/peaks.cu); 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz = 37.2"}))
PY
wait
cat gpurun_out/r02_fp64_peak.json gpurun_out/r02_fp64_peak_clocks.txt
