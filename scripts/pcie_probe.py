"""PCIe copy rates on the GPU box: pinned H2D of one c2 slice, D2H of U (strided and contiguous), both at once."""
import torch, time
T, S, k = 744, 1038240, 100
host = torch.empty((T, S), dtype=torch.float32, pin_memory=True); host.fill_(1.0)
dev = torch.empty((T, S), dtype=torch.float32, device="cuda")
Ubuf = torch.zeros((S, 112), device="cuda"); U = Ubuf[:, :k]
Uc = torch.zeros((S, k), device="cuda")
hU = torch.empty((S, k), dtype=torch.float32, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

ms = t(lambda: dev.copy_(host, non_blocking=True)); print(f"H2D 3.09 GB pinned: {ms:.1f} ms = {3.0898/ms*1e3:.1f} GB/s")
ms = t(lambda: hU.copy_(U, non_blocking=True)); print(f"D2H U strided (ld 112 -> 100): {ms:.1f} ms = {0.4153/ms*1e3:.1f} GB/s")
ms = t(lambda: hU.copy_(Uc, non_blocking=True)); print(f"D2H U contiguous: {ms:.1f} ms = {0.4153/ms*1e3:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): hU.copy_(Uc, non_blocking=True)
ms = t(both); print(f"H2D + D2H(contiguous) concurrently: {ms:.1f} ms")
def both2():
    with torch.cuda.stream(s1): dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): hU.copy_(U, non_blocking=True)
ms = t(both2); print(f"H2D + D2H(strided) concurrently: {ms:.1f} ms")
# H2D while a bandwidth-heavy kernel runs
a = torch.empty(1 << 28, device="cuda"); b = torch.empty(1 << 28, device="cuda")
def with_compute():
    with torch.cuda.stream(s1): dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        for _ in range(40): b.copy_(a)
ms = t(with_compute); print(f"H2D concurrently with 40 x 1 GiB device copies: {ms:.1f} ms")
