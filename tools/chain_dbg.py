"""Timeline dump of the chain kernel (SSTB200_CHAIN_DBG=N prints every N-th launch): clock64 stamps of the three roles."""
import os, sys, ctypes as C
os.environ.setdefault("SSTB200_CHAIN_DBG", "7")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sst_b200 import flagship as fl, _lib as L
from sst_b200.engine import SSTEngine
dev = torch.device('cuda:0')
P = 150000
vfe, il, bb = fl.build_sst(fl.sst_cfg(num_blocks=2))
eng = SSTEngine(fl.VOXEL_SIZE, fl.PC_RANGE, vfe.to(dev), il, bb.to(dev), max_points=P, batch_size=1, precision='bf16', device=dev, use_graph=False)
eng.load_frames_device(fl.synth_frame(1000, P).to(dev), torch.tensor([0, P], dtype=torch.int32, device=dev))
eng.run(); torch.cuda.synchronize()
eng.run(); torch.cuda.synchronize()
