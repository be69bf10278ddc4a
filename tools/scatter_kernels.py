"""DynamicScatter / scatter_v2 on the config-2 shape (150k points, C=128), twice each, for an ncu launch list with DRAM bytes:
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... python tools/scatter_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sst_b200 import flagship as fl, ops  # noqa: E402

dev = torch.device("cuda:0")
P, C = 150000, 128
pts = fl.synth_frame(1000, P).to(dev)
coors = ops.Voxelization(fl.VOXEL_SIZE, fl.PC_RANGE, -1)(pts)
feats = torch.randn(P, C, device=dev)
c64 = torch.cat([torch.zeros(P, 1, dtype=torch.int64, device=dev), coors.long()], 1)
for avg in (False, True):
    ds = ops.DynamicScatter(fl.VOXEL_SIZE, fl.PC_RANGE, avg)
    for it in range(2):
        if it == 1:
            torch.cuda.profiler.start()   # `ncu --profile-from-start off`: warm passes only
        ds(feats, coors)
        torch.cuda.synchronize()
        torch.arange(7, device=dev)
        torch.cuda.profiler.stop()
for it in range(2):
    if it == 1:
        torch.cuda.profiler.start()
    ops.scatter_v2(feats, c64, "max")
    torch.cuda.synchronize()
    torch.arange(7, device=dev)
print("done")
