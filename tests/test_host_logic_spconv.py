"""CPU: host-side logic of sst_b200/spconv_modules.py (container / layer glue, indice_key reuse, BN folding, fused vs unfused
epilogues, state-dict keys) with the three C-ABI calls replaced by the oracle's restatements - no compute call reaches the library.
The real kernels are checked on the GPU in tests/test_gpu_spconv.py."""
import os

import numpy as np
import pytest
import torch

from oracle import spconv_oracle as SO

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture()
def SP(monkeypatch):
    from sst_b200 import spconv_modules as SP

    def conv_out_coors(indices, batch_size, in_shape, out_shape, ksize, stride, padding):
        return SO.out_coors(SP._coors4(indices), batch_size, in_shape, ksize, stride, padding)

    def conv_table(in_indices, out_indices, batch_size, in_shape, out_shape, ksize, stride, padding, want_nbr=True, want_inv=False, check=True):
        ci, co = SP._coors4(in_indices), SP._coors4(out_indices)
        nbr = SO.neighbour_table(ci, co, batch_size, in_shape, ksize, stride, padding)
        inv = None
        if want_inv:
            inv = torch.full((ci.shape[0], nbr.shape[1]), -1, dtype=torch.int32)
            o, k = torch.nonzero(nbr >= 0, as_tuple=True)
            inv[nbr[o, k].long(), k] = o.int()
        return (nbr if want_nbr else None), inv

    def indice_conv(features, nbr, weight, weight_h16=None, scale=None, shift=None, residual=None, relu=False, precision="fp32",
                    transposed_table=None):
        if torch.is_grad_enabled() and (features.requires_grad or weight.requires_grad):
            assert scale is None and shift is None and residual is None and not relu and transposed_table is not None
            nt = transposed_table()   # the table the dX launch would run on: same pairs, transposed
            assert nt.shape == (features.shape[0], nbr.shape[1])
            o, k = torch.nonzero(nbr >= 0, as_tuple=True)
            assert torch.equal(nt[nbr[o, k].long(), k].long(), o) and int((nt >= 0).sum()) == o.numel()
            return SO.indice_conv(features, nbr, weight)   # torch ops: differentiable
        return SO.indice_conv(features, nbr, weight.detach(), scale, shift, residual, relu)

    monkeypatch.setattr(SP, "conv_out_coors", conv_out_coors)
    monkeypatch.setattr(SP, "conv_table", conv_table)
    monkeypatch.setattr(SP, "indice_conv", indice_conv)
    return SP


def _sorted(f, c):
    c = c.long()
    order = torch.argsort(((c[:, 0] * 4096 + c[:, 1]) * 4096 + c[:, 2]) * 4096 + c[:, 3])
    return f[order], c[order].int()


def test_unet_and_mixer_glue_against_reference_golden(SP):
    from sst_b200 import registry
    z = np.load(os.path.join(G, "spconv_unet.npz"))
    net = registry.MODELS.build(dict(type="SimpleSparseUNet", **SO.SP_UNET, return_multiscale_features=True))
    sd = {k[len("unet_sd."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("unet_sd.")}
    assert set(sd) == set(net.state_dict()), "state-dict keys differ from the reference's SimpleSparseUNet"
    net.load_state_dict(sd)
    net.eval()
    feats, coors = torch.from_numpy(z["unet_feats"]), torch.from_numpy(z["unet_coors"])
    with torch.no_grad():
        out = net(dict(voxel_feats=feats, voxel_coors=coors))[0]
        SP.FUSE_EPILOGUE = False
        try:
            out2 = net(dict(voxel_feats=feats, voxel_coors=coors))[0]
        finally:
            SP.FUSE_EPILOGUE = True
    for o in (out, out2):
        assert torch.equal(o["voxel_coors"], coors)
        torch.testing.assert_close(o["voxel_feats"], torch.from_numpy(z["unet_out"]), rtol=1e-3, atol=1e-4)
        for i, d in enumerate(o["decoder_features"]):
            f, c = _sorted(d.features, d.indices)
            assert torch.equal(c, torch.from_numpy(z[f"unet_ms{i}_c"]).int())
            torch.testing.assert_close(f, torch.from_numpy(z[f"unet_ms{i}_f"]), rtol=1e-3, atol=1e-4)
    mix = registry.MODELS.build(dict(type="VirtualVoxelMixer", **SO.SP_MIXER))
    sd = {k[len("mixer_sd."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("mixer_sd.")}
    assert set(sd) == set(mix.state_dict())
    mix.load_state_dict(sd)
    mix.eval()
    with torch.no_grad():
        f, c, shape = mix(torch.from_numpy(z["mixer_feats"]), torch.from_numpy(z["mixer_coors"]), 3)
    assert list(shape) == SO.SP_MIXER["sparse_shape"]
    torch.testing.assert_close(f, torch.from_numpy(z["mixer_out"]), rtol=1e-3, atol=1e-4)


def test_layer_glue_against_reference_golden(SP):
    z = np.load(os.path.join(G, "spconv_layers.npz"))
    feats, coors, shape = torch.from_numpy(z["l_feats"]), torch.from_numpy(z["l_coors"]), z["l_shape"].tolist()
    with torch.no_grad():
        for name in "abc":
            cfg = z[f"conv_{name}_cfg"].tolist()
            conv = SP.SparseConv3d(8, 12, cfg[0:3], stride=cfg[3:6], padding=cfg[6:9], bias=False, indice_key="k")
            inv = SP.SparseInverseConv3d(12, 8, cfg[0:3], indice_key="k", bias=False)
            conv.weight.copy_(torch.from_numpy(z[f"conv_{name}_w"]))
            inv.weight.copy_(torch.from_numpy(z[f"inv_{name}_w"]))
            y = conv(SP.SparseConvTensor(feats, coors, shape, 2))
            assert y.spatial_shape == z[f"conv_{name}_shape"].tolist()
            assert torch.equal(y.indices, torch.from_numpy(z[f"conv_{name}_coors"]).int())
            torch.testing.assert_close(y.features, torch.from_numpy(z[f"conv_{name}_out"]), rtol=1e-4, atol=1e-5)
            back = inv(y)
            assert torch.equal(back.indices, coors) and back.spatial_shape == shape
            torch.testing.assert_close(back.features, torch.from_numpy(z[f"inv_{name}_out"]), rtol=1e-4, atol=1e-5)
        sub = SP.SubMConv3d(8, 12, 3, padding=0, bias=True, indice_key="s")
        sub.weight.copy_(torch.from_numpy(z["subm_w"]))
        sub.bias.copy_(torch.from_numpy(z["subm_b"]))
        torch.testing.assert_close(sub(SP.SparseConvTensor(feats, coors, shape, 2)).features, torch.from_numpy(z["subm_out"]), rtol=1e-4, atol=1e-5)


def test_2d_layers_and_dense(SP):
    """ndim = 2 layers ride on the 3-D tables with a unit z axis; dense() matches the reference layout [B, C, *spatial]"""
    g = torch.Generator().manual_seed(0)
    yx = torch.unique(torch.randint(0, 12, (60, 2), generator=g), dim=0)
    coors = torch.cat([torch.zeros((yx.shape[0], 1), dtype=torch.long), yx], 1).int()
    feats = torch.randn((coors.shape[0], 4), generator=g)
    conv = SP.SparseConv2d(4, 8, 3, stride=2, padding=1, bias=True)
    with torch.no_grad():
        y = conv(SP.SparseConvTensor(feats, coors, [12, 12], 1))
        dense_in = SP.SparseConvTensor(feats, coors, [12, 12], 1).dense()
        ref = torch.nn.functional.conv2d(dense_in, conv.weight.permute(3, 2, 0, 1), conv.bias, stride=2, padding=1)
    assert y.indices.shape[1] == 3 and y.spatial_shape == [6, 6]
    d = y.dense()
    assert d.shape == ref.shape
    mask = (d != 0).any(1, keepdim=True)
    torch.testing.assert_close(d, ref * mask, rtol=1e-4, atol=1e-5)


def test_gradient_composition_matches_dense_autograd(SP):
    """with gradients enabled the containers run epilogue-free convolutions + torch BatchNorm / ReLU / residual; the gradient of that
    composition (table form) equals autograd through the oracle's dense restatement, for the input and every parameter"""
    from sst_b200 import registry
    cfg = dict(SO.SP_UNET)
    torch.manual_seed(2)
    net = registry.MODELS.build(dict(type="SimpleSparseUNet", **cfg)).eval()   # frozen-BN fine-tuning: eval statistics, grads on
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
        for p in net.parameters():
            if p.dim() == 5:
                p.mul_(2.0)
    feats, coors = SO.synth_sparse(6, 2, cfg["sparse_shape"], 300, 8)
    x = feats.clone().requires_grad_(True)
    out = net(dict(voxel_feats=x, voxel_coors=coors))[0]["voxel_feats"]
    probe = torch.randn(out.shape, generator=g)
    (out * probe).sum().backward()
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in net.state_dict().items()}
    x2 = feats.clone().requires_grad_(True)
    ref = SO.sparse_unet_forward(sd, x2, coors, 2, cfg["sparse_shape"], cfg["encoder_channels"], cfg["encoder_paddings"],
                                 cfg["decoder_channels"], cfg["decoder_paddings"])[0]
    torch.testing.assert_close(out.detach(), ref.detach(), rtol=1e-3, atol=1e-5)
    (ref * probe).sum().backward()

    def close(a, b):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-4 * float(b.abs().max()) + 1e-7)
    close(x.grad, x2.grad)
    n = 0
    for name, p in net.named_parameters():
        assert sd[name].grad is not None, name
        close(p.grad, sd[name].grad)
        n += 1
    assert n > 60


def test_dynamic_scatter_vfe_training_composition(monkeypatch):
    """host logic of DynamicScatterVFE's training composition with the two library calls replaced by the oracle's restatements:
    same outputs and gradients as the oracle's training=True forward (pinned to the reference class in train mode)"""
    from oracle import sst_oracle as O
    from sst_b200 import ops, registry
    monkeypatch.setattr(ops, "unique_rows", lambda coors, return_counts=False, bounds=None: torch.unique(coors, return_inverse=True, dim=0))
    monkeypatch.setattr(ops, "segment_reduce", lambda src, index, mode, num_segments=None, want_argmax=True: (O.segment_reduce(src, index, mode, num_segments), None))
    vs, rng = (0.25, 0.25, 0.2), [-80, -80, -2, 80, 80, 4]
    torch.manual_seed(0)
    m = registry.MODELS.build(dict(type='DynamicScatterVFE', in_channels=5, feat_channels=[32, 32], with_cluster_center=True, with_voxel_center=True,
                                   voxel_size=vs, point_cloud_range=rng, norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01),
                                   unique_once=True, rel_dist_scaler=10.0)).train()
    monkeypatch.setattr(type(m), "_check", lambda self, f, c: None)   # the CUDA-tensor guard: nothing reaches the library in this test
    pts = torch.cat([torch.cat([O.synth_frame(3 + b, 2000), torch.rand(2000, 2)], 1) for b in range(2)])
    co = torch.cat([torch.nn.functional.pad(O.dynamic_voxelize(pts[b * 2000:(b + 1) * 2000], vs, rng), (1, 0), value=b) for b in range(2)]).long()
    w = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in m.state_dict().items()}
    x_ref = pts.clone().requires_grad_(True)
    of, oc, oinv = O.dynamic_scatter_vfe_forward(x_ref, co, w, vs, rng, 2, rel_dist_scaler=10.0, training=True)
    (of.square()).sum().backward()
    x = pts.clone().requires_grad_(True)
    vf, vc, inv = m(x, co, return_inv=True)
    assert torch.equal(vc, oc) and torch.equal(inv, oinv)
    torch.testing.assert_close(vf, of, rtol=1e-5, atol=1e-6)
    (vf.square()).sum().backward()
    torch.testing.assert_close(x.grad, x_ref.grad, rtol=1e-4, atol=1e-6)
    for name, p in m.named_parameters():
        torch.testing.assert_close(p.grad, w[name].grad, rtol=1e-4, atol=1e-5 * float(w[name].grad.abs().max()) + 1e-8)


def test_sir_training_composition(monkeypatch):
    """host logic of SIR / SIRLayer when a gradient is needed (composition over ops.segment_reduce + torch Linear / LayerNorm / GELU) with
    the library calls replaced by the oracle's restatements: outputs and gradients equal autograd through the oracle's sir_forward, which
    tests/test_oracle_vs_reference.py pins to the reference's own autograd"""
    from oracle import sst_oracle as O
    from sst_b200 import ops
    from sst_b200.sir_modules import SIR
    monkeypatch.setattr(ops, "unique_rows", lambda coors, return_counts=False, bounds=None: torch.unique(coors, return_inverse=True, dim=0))
    monkeypatch.setattr(ops, "segment_reduce", lambda src, index, mode, num_segments=None, want_argmax=True: (O.segment_reduce(src, index, mode, num_segments), None))
    torch.manual_seed(0)
    m = SIR(num_blocks=3, in_channels=[20, 37, 37], feat_channels=[[32, 32]] * 3, rel_mlp_hidden_dims=[[16, 32]] * 3, norm_cfg=dict(type='LN', eps=1e-3),
            mode='max', xyz_normalizer=[20, 20, 4], act='gelu', unique_once=True).train()
    g = torch.Generator().manual_seed(1)
    N, G = 1500, 40
    points = torch.cat([torch.randn(N, 3, generator=g) * 10, torch.rand(N, 2, generator=g)], 1)
    feats = torch.randn(N, 15, generator=g)
    gid = torch.randint(0, G, (N,), generator=g)
    coors = torch.stack([gid % 3, torch.zeros_like(gid), gid], 1)
    fcl = torch.randn(N, 3, generator=g) * 2
    w = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    f_ref = feats.clone().requires_grad_(True)
    ref = O.sir_forward(points, f_ref, coors, fcl, w, 3, 3, 2, [20, 20, 4])
    (ref[0].square().sum() + ref[1].square().sum()).backward()
    f = feats.clone().requires_grad_(True)
    a, b, c = m(points, f, coors, fcl)
    assert torch.equal(c, ref[2])
    torch.testing.assert_close(a, ref[0], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(b, ref[1], rtol=1e-5, atol=1e-5)
    (a.square().sum() + b.square().sum()).backward()
    torch.testing.assert_close(f.grad, f_ref.grad, rtol=1e-4, atol=1e-5 * float(f_ref.grad.abs().max()))
    n = 0
    for name, p in m.named_parameters():
        torch.testing.assert_close(p.grad, w[name].grad, rtol=1e-4, atol=1e-5 * float(w[name].grad.abs().max()) + 1e-8)
        n += 1
    assert n >= 30
