import sys; sys.path.insert(0, ".")
import torch, jaxovercooked_b200 as ocb
from jaxovercooked_b200 import random as ocr
lay = ocb.overcooked_layouts[sys.argv[1] if len(sys.argv) > 1 else "efficiency_test"]
env = ocb.make("overcooked", layout=lay, max_steps=17)
print("epw", env.info.envs_per_warp, "smem", env.info.smem_bytes_per_cta)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 203
obs, st = env.reset(ocr.split(ocr.PRNGKey(0), N))
torch.cuda.synchronize(); print("reset ok")
a = torch.randint(0, 6, (N,), dtype=torch.int32, device="cuda")
for t in range(3):
    obs, st, *_ = env.step(ocr.split(ocr.PRNGKey(t), N), st, {"agent_0": a, "agent_1": a}, donate=True)
torch.cuda.synchronize(); print("steps ok")
