import torch, time
n = 205357056 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
print(f"H2D 205MB {a:.3f} ms = {0.2054/a*1e3:.1f} GB/s; D2H {b:.3f} ms = {0.2054/b*1e3:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): both()
torch.cuda.synchronize()
print(f"both directions concurrently: {(time.perf_counter()-t0)/10*1e3:.3f} ms per pair")
