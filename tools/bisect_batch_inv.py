"""Which component makes sample b of a batched forward differ from its bs-1 forward?  (GPU tool)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anoddpm_b200.UNet import UNetModel
from oracle.unet_oracle import fill_state_dict

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())

kw = dict(img_size=256, base_channels=128, n_heads=2, attention_resolutions="16,8")
m = UNetModel(**kw).eval()
m.load_state_dict(fill_state_dict(m.state_dict(), seed=1))
m = m.cuda()
g = torch.Generator().manual_seed(8)
x = torch.randn(8, 1, 256, 256, generator=g).cuda()
t = torch.tensor([999, 0, 431, 7, 250, 800, 1, 500], device="cuda")
for impl, h1, fuse in (("exact", True, True), ("tc_tf32", True, True), ("tc2", True, True), ("tc", False, False), ("tc", True, False), ("tc", True, True)):
    m.conv_impl, m.h1_fp16, m.fuse_gn = impl, h1, fuse
    with torch.no_grad():
        yb = m(x, t)
        e = [rel(yb[b:b + 1], m(x[b:b + 1], t[b:b + 1])) for b in (0, 3, 7)]
        # same timestep for all samples
        t2 = torch.full((8,), 431, device="cuda")
        yb2 = m(x, t2)
        e2 = [rel(yb2[b:b + 1], m(x[b:b + 1], t2[b:b + 1])) for b in (0, 3, 7)]
    eng = m.engine()
    print(f"impl={impl} h1_fp16={h1} fuse_gn={fuse}: mixed t {['%.2e' % v for v in e]}  same t {['%.2e' % v for v in e2]}  "
          f"gn_tables={eng.gn_tables} launches={eng.launches_per_forward}", flush=True)
