"""Pure-write, pure-read and copy HBM rates of this B200 (torch fill_/sum/copy_ over 4 GiB, CUDA events, best of 10)."""
import json
import torch

def best(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = 1e30
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        t = min(t, a.elapsed_time(b) * 1e-3)
    return t

n = 1 << 29          # doubles: 4 GiB
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
out = {"bytes": n * 8}
out["write_gbs_fill"] = n * 8 / best(lambda: x.fill_(1.0)) / 1e9
out["write_gbs_memset"] = n * 8 / best(lambda: x.zero_()) / 1e9
out["read_gbs_sum"] = n * 8 / best(lambda: x.sum()) / 1e9
out["copy_gbs_read_plus_write"] = 2 * n * 8 / best(lambda: y.copy_(x)) / 1e9
print(json.dumps(out))
