"""SURVEY 8f next-1 / next-3 kernels once each inside a profiler range, for `ncu --set full --profile-from-start off`: the input-resolution
SubM convolution of the FSD U-Net in the three precisions (fp16 operands, split-fp16 = fp32 tolerance, FFMA), its weight gradient, the
neighbour tables, and one instance grouping.  No oracle import.      python tools/next_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sst_b200 import flagship as fl, fsd_modules as FM, spconv_modules as SP  # noqa: E402

dev = torch.device("cuda:0")
feats, coors = fl.fsd_sweep_voxels()
feats, coors = feats.to(dev), coors.to(dev)
shape = fl.FSD_UNET["sparse_shape"]
w = torch.randn((27, 64, 64), device=dev) * 0.02
g = torch.Generator().manual_seed(0)
centres = (torch.rand((20000, 3), generator=g) * 140 - 70).to(dev)


def once():
    nbr, inv = SP.conv_table(coors, coors, 1, shape, shape, [3, 3, 3], [1, 1, 1], [1, 1, 1], want_inv=True)
    oshape = SP.get_conv_output_size(shape, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1])
    SP.conv_out_coors(coors, 1, shape, oshape, [3, 3, 3], [2, 2, 2], [1, 1, 1])
    for prec in ("bf16", "fp32_tc", "fp32"):
        out = SP.indice_conv(feats, nbr, w, precision=prec)
    SP.indice_conv_backward_weight(feats, nbr, out, 27, 64, 64)
    FM.connected_components(centres, None, 0.6, batch_size=1, xy_bounds=([-80.0, -80.0], [80.0, 80.0]))


with torch.no_grad():
    once()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    once()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok")
