mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02l_tests.log; cat gpurun_out/r02l_tests.log
python - <<'PY' > gpurun_out/r02l_scalar.txt 2>&1
import sys, time, ctypes as ct
sys.path.insert(0, '.')
import numpy as np
import bench
from mimic_mda_b200 import chisq
lib = chisq.prepare_shared_object()
for _ in range(3):
    print(bench.scalar_call_latency(lib, chisq, calls=2000))
PY
cat gpurun_out/r02l_scalar.txt
