"""GPU: gradients of the training path (csrc/sra_train.cu through sst_b200.train) against autograd through the CPU oracle -
which tests/test_oracle_vs_reference.py::test_oracle_gradients_match_reference pins to the reference's own autograd."""
import pytest
import torch

from oracle import sst_oracle as O

pytestmark = pytest.mark.gpu

VS = (0.32, 0.32, 6)
RNG = [-74.88, -74.88, -2, 74.88, 74.88, 4]
DROP_TEST = {0: {'max_tokens': 30, 'drop_range': (0, 30)}, 1: {'max_tokens': 60, 'drop_range': (30, 60)},
             2: {'max_tokens': 100, 'drop_range': (60, 100)}, 3: {'max_tokens': 144, 'drop_range': (100, 100000)}}


def _setup(cuda, P, num_blocks, seed=0):
    from sst_b200.sst_modules import SSTInputLayerV2, SSTv2
    p = O.synth_frame(900 + seed, P)
    c = torch.unique(O.dynamic_voxelize(p, VS, RNG), dim=0)
    coors = torch.nn.functional.pad(c, (1, 0), value=0).int()
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(coors.shape[0], 128, generator=g)
    torch.manual_seed(seed)
    bb = SSTv2(d_model=[128] * num_blocks, nhead=[8] * num_blocks, num_blocks=num_blocks, dim_feedforward=[256] * num_blocks,
               output_shape=[468, 468], num_attached_conv=0, to_bev=False, precision="bf16")
    with torch.no_grad():
        for n_, q in bb.named_parameters():
            if q.dim() == 1:
                q.copy_(torch.randn(q.shape, generator=g) * 0.1 + (1.0 if "norm" in n_ and "weight" in n_ else 0.0))
    il = SSTInputLayerV2(DROP_TEST, (12, 12, 1), (468, 468, 1), shuffle_voxels=False, mute=True)
    return feats, coors, il, bb


@pytest.mark.parametrize("P,num_blocks", [(20000, 1), (60000, 2)])
def test_sra_stack_gradients_match_oracle_autograd(cuda, P, num_blocks):
    feats, coors, il, bb = _setup(cuda, P, num_blocks)
    # oracle (fp32, CPU autograd)
    w = {k: v.detach().clone().requires_grad_(True) for k, v in bb.state_dict().items()}
    xo = feats.clone().requires_grad_(True)
    info_o = O.input_layer_v2(xo, coors, DROP_TEST, (12, 12, 1), (468, 468, 1))
    yo = O.sstv2_forward(info_o, w, [8] * num_blocks, num_blocks)
    gy = torch.randn(yo.shape, generator=torch.Generator().manual_seed(5)) / yo.shape[0]
    (yo * gy).sum().backward()
    # CUDA training path
    bb = bb.to(cuda).train()
    xg = feats.to(cuda).requires_grad_(True)
    info = il.train()(xg, coors.to(cuda), 1)
    yg = bb(info)[0]["voxel_feats"]
    fwd_err = (yg.detach().cpu() - yo.detach()).abs().max().item() / yo.detach().abs().max().item()
    assert fwd_err < 1e-2, f"training forward rel err {fwd_err}"
    (yg * gy.to(cuda)).sum().backward()

    def rel(a, b):
        return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)

    assert rel(xg.grad.cpu(), xo.grad) < 2e-2, f"d input: {rel(xg.grad.cpu(), xo.grad)}"
    worst = {}
    for name, p in bb.named_parameters():
        assert p.grad is not None, name
        worst[name] = rel(p.grad.cpu(), w[name].grad)
    bad = {k: v for k, v in worst.items() if v > 2e-2}
    assert not bad, f"parameter gradients off (bf16 tolerance 2e-2 of max-abs): {bad}"


def test_flat_adamw_matches_torch(cuda):
    from sst_b200.train import FlatAdamW
    g = torch.Generator().manual_seed(0)
    ps = [torch.randn(s, generator=g) for s in ((33, 7), (128,), (5, 5, 5))]
    a = [p.clone().to(cuda).requires_grad_(True) for p in ps]
    b = [p.clone().to(cuda).requires_grad_(True) for p in ps]
    oa = FlatAdamW(a, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    ob = torch.optim.AdamW(b, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    for it in range(3):
        for x, y in zip(a, b):
            gr = torch.randn(x.shape, generator=g).to(cuda)
            x.grad.copy_(gr)
            y.grad = gr.clone()
        oa.step()
        ob.step()
    for x, y in zip(a, b):
        torch.testing.assert_close(x.detach(), y.detach(), rtol=1e-5, atol=1e-6)


def test_full_model_train_step_gradients(cuda):
    """DynamicVFE (training-mode batch-stat BN) -> SSTInputLayerV2 (training drop_info, shuffle off) -> SSTv2 (2 blocks): loss
    and every parameter gradient against autograd through the oracle (training=True restatement, pinned to the reference by
    tests/test_oracle_vs_reference.py)."""
    from sst_b200 import flagship as fl
    cfg = fl.sst_cfg(num_blocks=2)
    cfg["backbone"]["precision"] = "bf16"
    vfe, il, bb = fl.build_sst(cfg, seed=3)
    pts = torch.cat([O.synth_frame(41, 12000), O.synth_frame(42, 9000)])
    c3 = O.dynamic_voxelize(pts, VS, RNG)
    coors = torch.cat([torch.nn.functional.pad(c3[:12000], (1, 0), value=0), torch.nn.functional.pad(c3[12000:], (1, 0), value=1)])
    # oracle
    wv = {k: v.detach().clone().requires_grad_(k in dict(vfe.named_parameters())) for k, v in vfe.state_dict().items()}
    wb = {k: v.detach().clone().requires_grad_(True) for k, v in bb.state_dict().items()}
    vf_o, vc_o = O.dynamic_vfe_forward(pts, coors, wv, VS, RNG, 2, training=True)
    info_o = O.input_layer_v2(vf_o, vc_o, fl.DROP_TRAIN, (12, 12, 1), (468, 468, 1))
    out_o = O.sstv2_forward(info_o, wb, [8, 8], 2)
    loss_o = out_o.square().mean()
    loss_o.backward()
    # CUDA
    vfe, bb = vfe.to(cuda).train(), bb.to(cuda).train()
    il.train()
    il.shuffle_voxels = False
    vf, vc = vfe(pts.to(cuda), coors.to(cuda))
    assert torch.equal(vc.cpu().long(), vc_o.long())
    info = il(vf, vc, 2)
    assert torch.equal(info["voxel_keep_inds"].cpu(), info_o["voxel_keep_inds"])
    out = bb(info)[0]["voxel_feats"]
    loss = out.float().square().mean()
    loss.backward()
    assert abs(loss.item() - loss_o.item()) < 2e-2 * abs(loss_o.item())

    def rel(a, b):
        return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)

    bad = {}
    for name, p in list(vfe.named_parameters()) + list(bb.named_parameters()):
        ref = (wv if name in wv and wv[name].grad is not None and p.shape == wv[name].shape and name.startswith("vfe_layers") else wb)[name].grad
        r = rel(p.grad.cpu(), ref)
        if r > 3e-2:
            bad[name] = r
    assert not bad, bad
