"""Speed-of-light probes for the obs write stream (not part of the product)."""
import sys, time, torch
sys.path.insert(0, ".")
import jaxovercooked_b200 as ocb
from jaxovercooked_b200 import random as ocr
dev = torch.device("cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
lay = ocb.overcooked_layouts["cramped_room"]
env = ocb.make("overcooked", layout=lay)
obs, st = env.reset(ocr.split(ocr.PRNGKey(0), N))
S = 8
bufs = [torch.empty((2, N, 4, 5, 26), dtype=torch.uint8, device=dev) for _ in range(S)]
def timeit(fn, iters=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us
us = timeit(lambda i: bufs[i % S].fill_(1))
print(f"fill {bufs[0].numel()/1e6:.1f} MB: {us:.1f} us  -> {bufs[0].numel()/us/1e3:.0f} GB/s write-only; {N/us/1e3:.2f} G env/s equivalent")
src = torch.empty_like(bufs[0])
us = timeit(lambda i: bufs[i % S].copy_(bufs[(i + 3) % S]))
print(f"copy: {us:.1f} us -> {2*bufs[0].numel()/us/1e3:.0f} GB/s r+w")
# get_obs through the C ABI with preallocated outputs (monkeypatch _obs_buffers)
k = [0]
def ob(batch):
    b = bufs[k[0] % S]; k[0] += 1
    return b[0], b[1]
env._obs_buffers = ob
g = torch.cuda.CUDAGraph()
env.get_obs(st); torch.cuda.synchronize()
with torch.cuda.graph(g):
    for i in range(S): env.get_obs(st)
us = timeit(lambda i: g.replay(), 100) / S
print(f"get_obs: {us:.1f} us/launch -> {N/us/1e3:.2f} G env/s; obs bytes {2*N*520/us/1e3:.0f} GB/s")
