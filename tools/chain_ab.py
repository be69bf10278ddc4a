"""A/B of the SRA stack under SSTB200_CHAIN=0 (unfused GEMM launches) and =1 (warp-specialised TMA chain): each variant runs in
its own process (the switch is read once), dumps the 12-layer output, and the parent compares them and prints the timings.
    python tools/chain_ab.py            (parent)      python tools/chain_ab.py child <variant> <out.pt>"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(variant, out):
    import ctypes as C
    import torch
    from sst_b200 import flagship as fl, _lib as L
    from sst_b200.engine import SSTEngine
    dev = torch.device('cuda:0')
    P = int(os.environ.get("AB_POINTS", "150000"))
    nb = int(os.environ.get("AB_BLOCKS", "6"))
    vfe, il, bb = fl.build_sst(fl.sst_cfg(num_blocks=nb))
    eng = SSTEngine(fl.VOXEL_SIZE, fl.PC_RANGE, vfe.to(dev), il, bb.to(dev), max_points=P, batch_size=1, precision='bf16', device=dev)
    eng.load_frames_device(fl.synth_frame(1000, P).to(dev), torch.tensor([0, P], dtype=torch.int32, device=dev))
    feats, coors, num = eng.run()
    torch.cuda.synchronize()
    M = int(num.item())
    res = {"feats": feats[:M].cpu().clone(), "M": M}
    st = eng.stream
    # whole-frame graph replay timing + one stack call timing (stream launches)
    with torch.cuda.stream(st):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            eng.run()
        a.record(st)
        for _ in range(30):
            eng.run()
        b.record(st)
    torch.cuda.synchronize()
    res["frame_us"] = a.elapsed_time(b) * 1e3 / 30
    lib = L.lib()
    with torch.cuda.stream(st):
        c = L.ctx(dev)

        def call():
            L.check(c, lib.sstb200_sra_stack_forward(c, eng._layer_array, len(eng._layers), C.byref(eng._plan_structs[0]),
                                                     C.byref(eng._plan_structs[1]), eng.vf.data_ptr(), eng.x[0].data_ptr(),
                                                     eng.x[1].data_ptr(), eng.cap, eng.num.data_ptr(), 1))
        for _ in range(3):
            call()
        a.record(st)
        for _ in range(20):
            call()
        b.record(st)
    torch.cuda.synchronize()
    res["stack_us"] = a.elapsed_time(b) * 1e3 / 20
    torch.save(res, out)
    print(f"[{variant}] M={M} frame {res['frame_us']:.1f} us, stack ({2 * nb} layers) {res['stack_us']:.1f} us "
          f"= {res['stack_us'] / (2 * nb):.1f} us/layer", flush=True)


def main():
    import torch
    outs = {}
    for v in ("0", "1"):
        env = dict(os.environ, SSTB200_CHAIN=v)
        out = f"/tmp/chain_ab_{v}.pt"
        r = subprocess.run([sys.executable, __file__, "child", v, out], env=env, timeout=600)
        if r.returncode != 0:
            print(f"variant {v} failed rc={r.returncode}")
            continue
        outs[v] = torch.load(out)
    if len(outs) == 2:
        a, b = outs["0"]["feats"], outs["1"]["feats"]
        d = (a - b).abs().max().item()
        print(f"max |unfused - chain| = {d:.3e} (max |unfused| = {a.abs().max().item():.3e}); bit-equal: {torch.equal(a, b)}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2], sys.argv[3])
    else:
        main()
