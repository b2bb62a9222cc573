"""What a write-dominated stream can reach on this GPU: cudaMemset (pure write), a torch fill kernel, a device copy
(read + write, the MEASURED_PEAKS number) and a read-only reduction, each over 4 GiB."""
import json, torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
ms = t(lambda: a.zero_()); out["memset_GBs"] = 4 * n / ms / 1e6
ms = t(lambda: a.fill_(1.5)); out["fill_GBs"] = 4 * n / ms / 1e6
ms = t(lambda: b.copy_(a)); out["copy_GBs_read_plus_write"] = 8 * n / ms / 1e6
ms = t(lambda: a.sum()); out["read_sum_GBs"] = 4 * n / ms / 1e6
print(json.dumps(out))
