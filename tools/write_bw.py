"""Pure write / read / copy bandwidth of the box with torch (context for the pixel pass's store pattern)."""
import torch
n = 8 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
print("fill  %.0f GB/s" % (n / t(lambda: a.zero_()) / 1e6))
print("copy  %.0f GB/s (r+w)" % (2 * n / t(lambda: b.copy_(a)) / 1e6))
av = a.view(torch.int32)
print("read  %.0f GB/s" % (n / t(lambda: av.max()) / 1e6))
