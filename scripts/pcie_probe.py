"""Development tool (GPU): pinned-memory PCIe bandwidth for the e2e step's transfer sizes (H2D 43.6 MB, D2H 24.9 MB), alone and concurrently."""
import json, torch
h_in = torch.empty(43_569_152 // 4, dtype=torch.float32).pin_memory(); h_out = torch.empty(24_883_216 // 4, dtype=torch.float32).pin_memory()
d_in = torch.empty_like(h_in, device="cuda"); d_out = torch.empty_like(h_out, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def h2d(): d_in.copy_(h_in, non_blocking=True)
def d2h(): h_out.copy_(d_out, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
a, b, c = t(h2d), t(d2h), t(both)
print(json.dumps({"h2d_ms": a, "h2d_GBps": h_in.numel() * 4 / a / 1e6, "d2h_ms": b, "d2h_GBps": h_out.numel() * 4 / b / 1e6, "both_ms": c}))
