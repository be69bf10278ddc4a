"""In-graph latency of the frame engine as a function of the number of SRA blocks (slope = true per-layer cost)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sst_b200 import flagship as fl
from sst_b200.engine import SSTEngine
dev = torch.device('cuda:0')
P = 150000
pts = fl.synth_frame(1000, P).to(dev)
offs = torch.tensor([0, P], dtype=torch.int32, device=dev)
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
for nb in (0, 1, 2, 4, 6):
    cfg = fl.sst_cfg(num_blocks=max(nb, 1))
    vfe, il, bb = fl.build_sst(cfg)
    if nb == 0:
        bb.block_list = torch.nn.ModuleList([])
    eng = SSTEngine(fl.VOXEL_SIZE, fl.PC_RANGE, vfe.to(dev), il, bb.to(dev), max_points=P, batch_size=1, precision=prec, device=dev)
    st = eng.stream
    ts = []
    for i in range(12):
        with torch.cuda.stream(st):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.load_frames_device(pts, offs)
            a.record(st)
        eng.run()
        with torch.cuda.stream(st):
            b.record(st)
        ts.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ts[2:])
    print(f"blocks={nb} layers={2*nb} kernels={eng.launches_per_frame} other={eng.other_nodes_per_frame} latency_us={t[len(t)//2]*1e3:.1f} min={t[0]*1e3:.1f}")
