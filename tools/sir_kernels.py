"""BASELINE config 3 (FSD SIR, 150k points, 256 groups) twice through the module, for `ncu --metrics gpu__time_duration.sum`.
Usage: ncu ... python tools/sir_kernels.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sst_b200.sir_modules import SIR  # noqa: E402

dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
N, G = 150000, 256
g = torch.Generator().manual_seed(3)
sp = torch.cat([torch.randn(N, 3, generator=g) * 10, torch.rand(N, 2, generator=g)], 1).to(dev)
sf = torch.randn(N, 79, generator=g).to(dev)
gid = torch.randint(0, G, (N,), generator=g)
sc = torch.stack([gid % 3, torch.zeros_like(gid), gid], 1).to(dev)
fcl = (torch.randn(N, 3, generator=g) * 2).to(dev)
torch.manual_seed(0)
sir = SIR(num_blocks=3, in_channels=[84, 133, 133], feat_channels=[[128, 128]] * 3, rel_mlp_hidden_dims=[[16, 32]] * 3,
          norm_cfg=dict(type="LN", eps=1e-3), mode="max", xyz_normalizer=[20, 20, 4], act="gelu", unique_once=True).eval().to(dev)
sir.precision = prec
with torch.no_grad():
    for it in range(2):
        if it == 1:
            torch.cuda.profiler.start()   # `ncu --profile-from-start off`: the warm pass only
        sir(sp, sf, sc, fcl)
        torch.cuda.synchronize()
print("done")
