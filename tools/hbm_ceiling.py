"""Practical HBM ceilings for this kernel's read/write mix (26 % reads, 74 % writes)."""
import torch
n = 1 << 30  # 8 GiB of fp64
a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: b.copy_(a)); print(f"copy   : {2 * n * 8 / ms / 1e6:.0f} GB/s")
ms = t(lambda: a.zero_()); print(f"memset : {n * 8 / ms / 1e6:.0f} GB/s")
ms = t(lambda: a.fill_(1.5)); print(f"fill   : {n * 8 / ms / 1e6:.0f} GB/s")
ms = t(lambda: a.sum()); print(f"read   : {n * 8 / ms / 1e6:.0f} GB/s")
# 1 read stream + 3 write streams (25/75 mix like the fused kernel)
c = torch.empty(n // 4, dtype=torch.float64, device="cuda")
o = [torch.empty(n // 4, dtype=torch.float64, device="cuda") for _ in range(3)]
def mix():
    torch.add(c, 1.0, out=o[0]); torch.add(c, 2.0, out=o[1]); torch.add(c, 3.0, out=o[2])
ms = t(mix); print(f"3x (read 1, write 1) add: {6 * (n // 4) * 8 / ms / 1e6:.0f} GB/s")
