"""CPU, build container only: pins oracle/spconv_oracle.py (dense restatement of the sparse convolutions) against the
reference's own vendored spconv v1 - its Python layers imported unmodified through oracle/ref_shim.load_spconv(), its C++ sources
compiled unmodified into oracle/_ref by oracle/build_ref.py - and against the reference's SimpleSparseUNet / VirtualVoxelMixer classes.
Skipped where /root/reference does not exist (the GPU box): there tests/golden/spconv_*.npz carry the reference's outputs."""
import pytest
import torch

from oracle import build_ref, ref_shim, spconv_oracle as SO

pytestmark = pytest.mark.skipif(not ref_shim.available() or not build_ref.built("sparse_conv_ext_ref"),
                                reason="reference tree / oracle/_ref spconv build not present")


@pytest.fixture(scope="module")
def S():
    return ref_shim.load_spconv()


def _by_coor(feats, coors):
    """rows sorted lexicographically by coordinate (the reference's SparseConv output order is hash-insertion order)"""
    c = coors.long()
    key = ((c[:, 0] * 4096 + c[:, 1]) * 4096 + c[:, 2]) * 4096 + c[:, 3]
    order = torch.argsort(key)
    return feats[order], coors[order]


@pytest.mark.parametrize("ks,stride,padding", [(3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0), (2, 2, 0)])
def test_sparse_conv_and_inverse_match_reference(S, ks, stride, padding):
    feats, coors = SO.synth_sparse(1, 2, (9, 20, 24), 300, 8)
    torch.manual_seed(0)
    conv = S.SparseConv3d(8, 12, ks, stride=stride, padding=padding, bias=False, indice_key="k")
    inv = S.SparseInverseConv3d(12, 6, ks, indice_key="k", bias=False)
    x = S.SparseConvTensor(feats, coors, [9, 20, 24], 2)
    with torch.no_grad():
        y = conv(x)
        z = inv(y)
    of, oc, oshape = SO.sparse_conv(feats, coors, 2, [9, 20, 24], conv.weight.detach(), conv.stride, conv.padding)
    assert list(y.spatial_shape) == oshape
    yf, yc = _by_coor(y.features, y.indices)
    assert torch.equal(yc.int(), oc)          # oracle order is lexicographic
    torch.testing.assert_close(of, yf, rtol=1e-4, atol=1e-5)
    # the neighbour table regrouped from the reference's index pairs
    nbr = SO.neighbour_table(coors, oc, 2, [9, 20, 24], conv.kernel_size, conv.stride, conv.padding)
    torch.testing.assert_close(SO.indice_conv(feats, nbr, conv.weight.detach()), of, rtol=1e-4, atol=1e-5)
    zi = SO.inverse_conv(of, oc, 2, oshape, coors, [9, 20, 24], inv.weight.detach(), conv.stride, conv.padding)
    assert torch.equal(z.indices, coors)      # the inverse conv restores the couple conv's input rows, in order
    torch.testing.assert_close(zi, z.features, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("ks,padding", [(3, 1), ((3, 1, 1), (1, 0, 0)), (3, 0), (3, (0, 1, 1))])
def test_subm_conv_matches_reference(S, ks, padding):
    """incl. layers built with padding != k//2: the index generation centres every SubM conv (spconv_ops.h:74-78)"""
    feats, coors = SO.synth_sparse(2, 2, (7, 18, 18), 250, 8)
    torch.manual_seed(1)
    conv = S.SubMConv3d(8, 10, ks, padding=padding, bias=True, indice_key="s")
    with torch.no_grad():
        y = conv(S.SparseConvTensor(feats, coors, [7, 18, 18], 2))
    assert torch.equal(y.indices, coors)
    o = SO.subm_conv(feats, coors, 2, [7, 18, 18], conv.weight.detach(), conv.bias.detach())
    torch.testing.assert_close(o, y.features, rtol=1e-4, atol=1e-5)
    nbr = SO.neighbour_table(coors, coors, 2, [7, 18, 18], conv.kernel_size, [1, 1, 1], [k // 2 for k in conv.kernel_size])
    torch.testing.assert_close(SO.indice_conv(feats, nbr, conv.weight.detach(), shift=conv.bias.detach()), y.features, rtol=1e-4, atol=1e-5)


UNET = dict(in_channels=8, sparse_shape=[9, 32, 32], order=('conv', 'norm', 'act'), norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01),
            base_channels=8, output_channels=16, encoder_channels=((8,), (8, 8, 8), (16, 16, 16), (16, 16, 16)),
            encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
            decoder_channels=((16, 16, 16), (16, 16, 8), (8, 8, 8), (8, 8, 8)), decoder_paddings=((1, 1), (1, 0), (0, 0), (0, 1)))


def _randomise_bn(m, seed):
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.2)


def test_simple_sparse_unet_matches_reference(S):
    torch.manual_seed(3)
    net = S.SimpleSparseUNet(**UNET, return_multiscale_features=True).eval()
    _randomise_bn(net, 5)
    feats, coors = SO.synth_sparse(4, 2, (9, 32, 32), 500, 8)
    with torch.no_grad():
        out = net(dict(voxel_feats=feats, voxel_coors=coors))[0]
    f, c, ms = SO.sparse_unet_forward(net.state_dict(), feats, coors, 2, UNET["sparse_shape"], UNET["encoder_channels"], UNET["encoder_paddings"],
                                      UNET["decoder_channels"], UNET["decoder_paddings"], return_multiscale=True)
    assert torch.equal(out["voxel_coors"], coors) and torch.equal(c, coors)
    torch.testing.assert_close(f, out["voxel_feats"], rtol=1e-3, atol=1e-4)
    assert len(ms) == len(out["decoder_features"]) == 4
    for (mf, mc), ref in zip(ms, out["decoder_features"]):
        rf, rc = _by_coor(ref.features, ref.indices)
        of, oc = _by_coor(mf, mc)
        assert torch.equal(oc.int(), rc.int())
        torch.testing.assert_close(of, rf, rtol=1e-3, atol=1e-4)


def test_virtual_voxel_mixer_matches_reference(S):
    """the mixer's conv_out is SubM kernel 3 / padding 0 on a fresh indice key - still a centred conv"""
    cfg = dict(in_channels=8, sparse_shape=[8, 24, 24], norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01), base_channels=8,
               output_channels=12, encoder_channels=((8,), (8, 8), (8, 8)), encoder_paddings=((1,), (1, 1), (1, 1)),
               decoder_channels=((8, 8, 8), (8, 8, 8), (8, 8, 8)), decoder_paddings=((1, 1), (1, 1), (1, 1)))
    torch.manual_seed(7)
    net = S.VirtualVoxelMixer(**cfg).eval()
    _randomise_bn(net, 9)
    feats, coors = SO.synth_sparse(8, 3, (8, 24, 24), 300, 8)
    with torch.no_grad():
        rf, rc, rshape = net(feats, coors, 3)
    f, c = SO.sparse_unet_forward(net.state_dict(), feats, coors, 3, cfg["sparse_shape"], cfg["encoder_channels"], cfg["encoder_paddings"],
                                  cfg["decoder_channels"], cfg["decoder_paddings"], mixer_out=True)
    assert torch.equal(rc, coors) and list(rshape) == cfg["sparse_shape"]
    torch.testing.assert_close(f, rf, rtol=1e-3, atol=1e-4)
