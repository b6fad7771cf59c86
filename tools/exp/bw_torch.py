"""Practical HBM rates on this box for the traffic mixes of the step's main pass (read 161 MB, write 354 MB)."""
import torch
dev = torch.device("cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
N = 512 << 20
src = torch.empty(N, dtype=torch.uint8, device=dev); dst = torch.empty(N, dtype=torch.uint8, device=dev)
t = timed(lambda: dst.copy_(src)); print("copy 512MB->512MB: %.1f us, %.2f TB/s (r+w)" % (t * 1e3, 2 * N / t / 1e9))
t = timed(lambda: dst.zero_()); print("fill 512MB: %.1f us, %.2f TB/s (w)" % (t * 1e3, N / t / 1e9))
a32 = src[: 160 << 20].view(torch.int32); out = dst[: 480 << 20].view(torch.int32).view(3, -1)
t = timed(lambda: torch.add(a32.unsqueeze(0), 1, out=out[:1]) if False else out.copy_(a32.unsqueeze(0).expand(3, -1)))
print("read 160MB write 480MB: %.1f us, %.2f TB/s" % (t * 1e3, (640 << 20) / t / 1e9))
s = src.view(torch.int64); t = timed(lambda: s.sum()); print("read 512MB (sum): %.1f us, %.2f TB/s (r)" % (t * 1e3, N / t / 1e9))
