"""recover_bev on the config-2 canvas (B=1, C=128, 468x468, ~30k voxels) twice, for an ncu launch list with DRAM bytes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sst_b200.sst_modules import SSTv2  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
M, C, ny, nx = 30000, 128, 468, 468
cells = torch.randperm(ny * nx, generator=g)[:M]
coors = torch.stack([torch.zeros_like(cells), torch.zeros_like(cells), cells // nx, cells % nx], 1).to(dev)
feat = torch.randn(M, C, generator=g).to(dev)
m = SSTv2(d_model=[C], nhead=[1], num_blocks=0, dim_feedforward=[C], output_shape=[ny, nx], num_attached_conv=0, to_bev=True)
for _ in range(2):
    m.recover_bev(feat, coors, 1)
    torch.cuda.synchronize()
print("done")
