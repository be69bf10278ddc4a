"""Kernel-time table of one training step (torch.profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from sst_b200 import flagship as fl
from sst_b200.train import TrainStep
dev = torch.device("cuda:0")
cfg = fl.sst_cfg(); cfg["backbone"]["precision"] = "bf16"; cfg["middle_encoder"]["shuffle_voxels"] = True
vfe, il, bb = fl.build_sst(cfg)
ts = TrainStep(vfe.to(dev), il, bb.to(dev), fl.VOXEL_SIZE, fl.PC_RANGE)
B = int(os.environ.get("TRAIN_FRAMES", "4"))
frames = [fl.synth_frame(1000 + i, 150000).to(dev) for i in range(B)]
for _ in range(2):
    ts.step(frames)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(frames)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=60))
