"""Not a test (not collected): prints the bf16 path's error against the oracle after 1 / 3 / 6 blocks at the full config-2 size.
    python tests/probe_bf16_error.py   (GPU box)"""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import sst_oracle as O
import test_gpu_sst as T
from sst_b200.sst_modules import SSTInputLayerV2
cuda=torch.device('cuda:0')
feats, coors = T._voxels((1000,), 150000, C=128)
for blocks in (1,3,6):
    m = T._sst_pair(128, 8, 256, blocks)
    il = SSTInputLayerV2((T.DROP_TRAIN, T.DROP_TEST), (12, 12, 1), (468, 468, 1), shuffle_voxels=False, mute=True).eval()
    w = {k: v.clone() for k, v in m.state_dict().items()}
    info_o = O.input_layer_v2(feats, coors, T.DROP_TEST, (12, 12, 1), (468, 468, 1))
    ref = O.sstv2_forward(info_o, w, [8] * blocks, blocks)
    m = m.to(cuda); m.precision='bf16'
    with torch.no_grad():
        info_g = il(feats.to(cuda), coors.to(cuda), 1)
        got = m(info_g)[0]["voxel_feats"].cpu()
    d = (got-ref)
    print(blocks, 'max-rel', d.abs().max().item()/ref.abs().max().item(), 'rms-rel', (d.square().mean().sqrt()/ref.square().mean().sqrt()).item(), 'max|ref|', ref.abs().max().item())
