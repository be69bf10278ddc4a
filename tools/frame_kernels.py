"""Run a few frames through the engine WITHOUT a CUDA graph (plain stream launches) so that ncu's per-kernel durations are
taken in the real data flow.  Usage: ncu --metrics gpu__time_duration.sum ... python tools/frame_kernels.py [precision]"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sst_b200 import flagship as fl
from sst_b200.engine import SSTEngine
dev = torch.device('cuda:0')
P = 150000
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
vfe, il, bb = fl.build_sst(fl.sst_cfg())
eng = SSTEngine(fl.VOXEL_SIZE, fl.PC_RANGE, vfe.to(dev), il, bb.to(dev), max_points=P, batch_size=1, precision=prec, device=dev, use_graph=False)
offs = torch.tensor([0, P], dtype=torch.int32, device=dev)
for i in range(3):
    if i == 2:
        torch.cuda.profiler.start()   # `ncu --profile-from-start off` captures exactly the third (warm) frame
    eng.load_frames_device(fl.synth_frame(1000 + i, P).to(dev), offs)
    eng.run()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("frames done; kernels per frame (python-side count n/a in stream mode)")
