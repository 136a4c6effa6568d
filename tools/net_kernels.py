import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bench import NETS
from pointwise_b200 import nets, _lib
from pointwise_b200.synth import make_points
name = sys.argv[1]
per_gpu, N, cin, ncls, dist = NETS[name]
dev = torch.device("cuda")
pts = torch.from_numpy(make_points(per_gpu, N, dist, seed=0)).to(dev)
feats = pts.clone() if cin == 3 else torch.rand(per_gpu, N, cin, device=dev) * 2 - 1
net = (nets.PointConvNetSeg(ncls, cin) if name == "seg_net" else nets.PointConvNetCls(ncls, N, cin)).to(dev)
labels = torch.randint(0, ncls, (per_gpu, N) if name == "seg_net" else (per_gpu,), device=dev)
params = list(net.parameters())
def step():
    for p in params: p.grad = None
    loss = net.loss(net.model(pts, feats, True), labels); loss.backward()
for _ in range(5): step()
torch.cuda.synchronize()
L = _lib.lib(); L.conv3p_profile_enable(1)
from torch.profiler import profile, ProfilerActivity
for _ in range(10): step()
torch.cuda.synchronize()
buf = C.create_string_buffer(16384); L.conv3p_profile_read(buf, 16384); L.conv3p_profile_enable(0)
tot = 0
for ln in buf.value.decode().splitlines():
    n, c, t = ln.split(); print(f"{n:32s} {int(c)//10:3d} launches/step {float(t)/10:.4f} ms/step"); tot += float(t)/10
print("sum of library kernels", tot)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:25]
for e in rows: print(f"{e.key[:70]:70s} {e.count/5:6.1f} {e.device_time_total/5/1000:.4f} ms/step")
