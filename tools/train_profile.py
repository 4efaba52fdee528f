#!/usr/bin/env python
"""Per-kernel device time of ONE native training step (BASELINE configs[1]: bs 8, 256x256) from torch.profiler (CUPTI kernel
records; ncu serialises the ~2500 launches of a step and needs half an hour).  Prints a markdown table for profiles/.

    python tools/train_profile.py > gpurun_out/train_step_kernels.md
"""
import collections
import copy
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

from anoddpm_b200 import GaussianDiffusionModel, UNetModel, get_beta_schedule   # noqa: E402
from anoddpm_b200.optim import FusedAdamWEMA   # noqa: E402
from anoddpm_b200.synth import fill_state_dict   # noqa: E402

dev = torch.device("cuda")
m = UNetModel(256, 128, n_heads=2, attention_resolutions="16,8")
m.load_state_dict(fill_state_dict(m.state_dict(), seed=1))
m = m.to(dev).train()
ema = copy.deepcopy(m)
opt = FusedAdamWEMA(m, ema, lr=1e-4, weight_decay=0.0, max_norm=1.0)
D = GaussianDiffusionModel([256, 256], get_beta_schedule(1000, "linear"), noise="simplex")
targs = {"train_start": True, "sample_distance": 800, "Batch_Size": 8}
np.random.seed(0)
x = (torch.rand(8, 1, 256, 256) * 2 - 1).to(dev)


def step():
    opt.zero_grad()
    loss, _ = D.p_loss(m, x, targs)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time_total > 0:
        n = re.sub(r"\(.*", "", ev.name.replace("void ", "").replace("(anonymous namespace)::", ""))
        n = re.sub(r"^.*::", "", n) if ("at::" in n or "cuda::" in n) else n
        agg[n[:70]][0] += 1
        agg[n[:70]][1] += ev.device_time_total / 1e3
        tot += ev.device_time_total / 1e3
print(f"# native training step (bs 8, 256x256, BASELINE configs[1]): {ms_step:.2f} ms/step by CUDA events; kernel time by CUPTI "
      f"records of one step: {tot:.2f} ms in {sum(v[0] for v in agg.values())} launches\n")
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot < 0.002:
        continue
    print(f"| {k} | {n} | {t:.3f} | {100 * t / tot:.1f}% |")
