"""Stream-launch timing of one SRA layer call (optionally with SSTB200_DEBUG_SKIP) - back-to-back calls, no graph."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from sst_b200 import flagship as fl, _lib as L
from sst_b200.engine import SSTEngine
dev = torch.device('cuda:0')
P = 150000
eng = SSTEngine(fl.VOXEL_SIZE, fl.PC_RANGE, *[m.to(dev) if hasattr(m, 'to') else m for m in fl.build_sst(fl.sst_cfg(num_blocks=1))],
                max_points=P, batch_size=1, precision='bf16', device=dev)
eng.load_frames_device(fl.synth_frame(1000, P).to(dev), torch.tensor([0, P], dtype=torch.int32, device=dev))
eng.run(); torch.cuda.synchronize()
lib = L.lib(); st = eng.stream
ls, shift = eng._layers[0]
N = 50
with torch.cuda.stream(st):
    c = L.ctx(dev)
    def call():
        L.check(c, lib.sstb200_sra_layer_forward(c, C.byref(ls), C.byref(eng._plan_structs[shift]), eng.vf.data_ptr(), eng.x[0].data_ptr(), eng.cap, eng.num.data_ptr(), 1))
    for _ in range(5): call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(N): call()
    b.record(st)
torch.cuda.synchronize()
print("us per layer call (stream launches, back to back):", a.elapsed_time(b) * 1e3 / N)
