"""CPU: pins oracle/sst_oracle.py against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container) and against the reference's own DynamicScatter test."""
import os

import numpy as np
import pytest
import torch

from oracle import sst_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
VS = (0.32, 0.32, 6)
RNG = [-74.88, -74.88, -2, 74.88, 74.88, 4]
DROP_TRAIN = {0: {'max_tokens': 30, 'drop_range': (0, 30)}, 1: {'max_tokens': 60, 'drop_range': (30, 60)},
              2: {'max_tokens': 100, 'drop_range': (60, 100000)}}
DROP_TEST = {0: {'max_tokens': 30, 'drop_range': (0, 30)}, 1: {'max_tokens': 60, 'drop_range': (30, 60)},
             2: {'max_tokens': 100, 'drop_range': (60, 100)}, 3: {'max_tokens': 144, 'drop_range': (100, 100000)}}


def _load(name):
    z = np.load(os.path.join(G, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _w(z, prefix):
    return {k[len(prefix):]: v for k, v in z.items() if k.startswith(prefix)}


def test_dynamic_scatter_reference_test_restated():
    """tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:56-84 (seeded): sorted coors with negatives
    removed equal exactly; mean / max feats allclose(atol=1e-2, rtol=1e-5) to the brute-force loop."""
    z = _load("dynamic_scatter.npz")
    for red, ref in (("mean", z["ref_mean"]), ("max", z["ref_max"])):
        f, c, m, n = O.dynamic_point_to_voxel_forward(z["feats"], z["coors"], red)
        assert torch.equal(c, z["ref_coors"])
        assert torch.allclose(f, ref, atol=1e-2, rtol=1e-5)
        assert int(n.sum()) == int((z["coors"].min(dim=-1).values >= 0).sum())
    # empty input keeps shapes (:22-36)
    f, c, m, n = O.dynamic_point_to_voxel_forward(torch.rand(0, 3), torch.zeros((0, 3), dtype=torch.int32), "mean")
    assert f.shape == (0, 3) and c.shape == (0, 3)


def test_dynamic_scatter_backward_matches_autograd():
    g = torch.Generator().manual_seed(1)
    feats = torch.rand(500, 4, generator=g, dtype=torch.float64)
    coors = torch.randint(-1, 6, (500, 3), generator=g, dtype=torch.int32)
    for red in ("sum", "mean", "max"):
        f, c, m, n = O.dynamic_point_to_voxel_forward(feats, coors, red)
        gr = torch.rand(f.shape, generator=g, dtype=torch.float64)
        got = O.dynamic_point_to_voxel_backward(gr, feats, f, m, n, red)
        x = feats.clone().requires_grad_(True)
        valid = m >= 0
        idx = m[valid].long()[:, None].expand(-1, 4)
        if red == "max":
            y = torch.full(f.shape, -1e30, dtype=torch.float64).scatter_reduce(0, idx, x[valid], "amax")
        else:
            y = torch.zeros(f.shape, dtype=torch.float64).scatter_reduce(0, idx, x[valid], red, include_self=False)
        (y * gr).sum().backward()
        assert torch.allclose(got, x.grad, atol=1e-12)  # random doubles: no ties, so amax grad == lowest-index rule


def test_vfe_and_input_layer_golden():
    z = _load("sst_small.npz")
    vf, vc = O.dynamic_vfe_forward(z["points"], z["coors"], _w(z, "vfe."), VS, RNG, 2)
    assert torch.equal(vc, z["vfe_coors"])
    torch.testing.assert_close(vf, z["vfe_feats"], rtol=1e-5, atol=1e-5)
    dropped = False
    for tag, drop in (("eval", DROP_TEST), ("train", DROP_TRAIN)):
        info = O.input_layer_v2(z["vfe_feats"], z["vfe_coors"], drop, (12, 12, 1), (468, 468, 1))
        assert torch.equal(info["voxel_keep_inds"], z[f"{tag}.keep"])
        dropped |= info["voxel_keep_inds"].numel() < z["vfe_coors"].shape[0]
        for i in range(2):
            assert torch.equal(info[f"batch_win_inds_shift{i}"], z[f"{tag}.batch_win_inds_shift{i}"])
            assert torch.equal(info[f"coors_in_win_shift{i}"], z[f"{tag}.coors_in_win_shift{i}"])
            assert torch.equal(info[f"voxel_drop_level_shift{i}"], z[f"{tag}.drop_level_shift{i}"])
            levels = [k for k in info[f"flat2win_inds_shift{i}"] if not isinstance(k, str)]
            assert sorted(levels) == sorted(int(k.split(".")[2]) for k in z if k.startswith(f"{tag}.f2w{i}.") and k.endswith(".inds"))
            for dl in levels:
                inds, pos = info[f"flat2win_inds_shift{i}"][dl]
                assert torch.equal(inds, z[f"{tag}.f2w{i}.{dl}.inds"]) and torch.equal(pos[0], z[f"{tag}.f2w{i}.{dl}.pos"])
                assert torch.equal(info[f"pos_dict_shift{i}"][dl], z[f"{tag}.pos{i}.{dl}"])
                assert torch.equal(info[f"key_mask_shift{i}"][dl], z[f"{tag}.mask{i}.{dl}"])
    assert dropped, "the training fixture is meant to exercise voxel dropping"


@pytest.mark.parametrize("name,lc", [("plain", {}), ("cosine", dict(cosine=True, tau_min=0.01)),
                                     ("prebn", dict(post_norm=False, use_bn=True))])
def test_sstv2_golden(name, lc):
    z = _load("sst_small.npz")
    info = O.input_layer_v2(z["vfe_feats"], z["vfe_coors"], DROP_TEST, (12, 12, 1), (468, 468, 1))
    out = O.sstv2_forward(info, _w(z, f"sst.{name}.w."), [4, 4], 2, "gelu", lc)
    torch.testing.assert_close(out, z[f"sst.{name}.out"], rtol=1e-4, atol=1e-5)


def test_sir_golden():
    z = _load("sir_small.npz")
    a, b, c = O.sir_forward(z["points"], z["feats"], z["coors"], z["f_cluster"], _w(z, "w."), 3, 3, 2, [20, 20, 4])
    assert torch.equal(c, z["out_coors"])
    torch.testing.assert_close(a, z["out_point"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(b, z["out_group"], rtol=1e-4, atol=1e-5)


def test_dynamic_scatter_vfe_golden():
    z = _load("dsvfe_small.npz")
    f, c, inv = O.dynamic_scatter_vfe_forward(z["points"], z["coors"], _w(z, "w."), (0.25, 0.25, 0.2), [-80, -80, -2, 80, 80, 4], 2,
                                              rel_dist_scaler=10.0)
    assert torch.equal(c, z["vcoors"]) and torch.equal(inv, z["inv"])
    torch.testing.assert_close(f, z["feats"], rtol=1e-5, atol=1e-5)


def test_voxelize_float32_grid():
    """grid = ceil((max-min)/vs) must be evaluated in float32 like the C++ (voxelization_cpu.cpp:151-158)."""
    assert O.grid_size(VS, RNG) == [468, 468, 1]
    p = torch.tensor([[74.87, -74.88, 3.99], [80.0, -90.0, 10.0], [-74.88, 0.0, -2.0]])
    c = O.dynamic_voxelize(p, VS, RNG)
    assert c.tolist() == [[0, 0, 467], [0, 0, 467], [0, 234, 0]]


def test_ingroup_and_window_helpers():
    g = torch.tensor([3, 1, 3, 3, 0, 1])
    assert O.ingroup_indices(g).tolist() == [0, 0, 1, 2, 0, 1]
    assert O.make_continuous_inds(torch.tensor([7, 3, 7, 100])).tolist() == [1, 0, 1, 2]
    co = torch.tensor([[0, 0, 5, 13], [1, 0, 467, 467]])
    w, ciw = O.get_window_coors(co, (468, 468, 1), (12, 12, 1), False)
    # no shift: shift = window size (sst_ops.py:283-286) -> x=13+12=25 -> win 2 rem 1; y=5+12=17 -> win 1 rem 5
    assert w.tolist() == [2 * 80 + 1 * 2 + 0, 3200 + 39 * 80 + 39 * 2] and ciw.tolist() == [[0, 5, 1], [0, 11, 11]]


@pytest.mark.parametrize("name,xyz,norm", [("xyz", True, False), ("xyznorm", True, True), ("noxyz", False, False)])
def test_voxel2point_neck_golden(name, xyz, norm):
    """Outputs of the unmodified reference Voxel2PointScatterNeck (tests/golden/neck_small.npz, oracle/make_golden.py)."""
    z = _load("neck_small.npz")
    out, mask = O.voxel2point_neck(z["points"], z["coors"], z["voxel_feats"], z["inds"], VS, RNG, xyz, norm)
    assert torch.equal(mask, z[f"mask_{name}"]) and torch.equal(out, z[f"out_{name}"])


@pytest.mark.parametrize("name", ["roomy", "capped"])
def test_hard_voxelize_golden(name):
    """Outputs of the reference's own C++ hard_voxelize (voxelization_cpu.cpp:43-142 compiled unmodified into oracle/_ref)."""
    z = _load("hard_voxelize.npz")
    mp, mv = [int(v) for v in z[f"cfg_{name}"]]
    v, c, n = O.hard_voxelize(z["points"], VS, RNG, mp, mv)
    assert torch.equal(c, z[f"coors_{name}"]) and torch.equal(n, z[f"npts_{name}"]) and torch.equal(v, z[f"voxels_{name}"])


def test_sst_v1_golden():
    """configs/sst names: the v1 SSTInputLayer + SSTv1 of the reference (tests/golden/sst_v1_small.npz) are the same maths as v2 -
    the oracle's v2 restatement reproduces the v1 output rows, the window ids up to v1's own numbering and the drop levels."""
    z = _load("sst_v1_small.npz")
    info = O.input_layer_v2(z["voxel_feats"], z["voxel_coors"], DROP_TEST, (12, 12, 1), (468, 468, 1))
    assert torch.equal(info["voxel_keep_inds"], z["keep"]) and torch.equal(info["voxel_coors"], z["coors"])
    for i in range(2):
        assert torch.equal(info[f"voxel_drop_level_shift{i}"], z[f"lvl{i}"])
        assert torch.equal(info[f"coors_in_win_shift{i}"][:, [2, 1]], z[f"ciw{i}"])          # v1 stacks (x, y)
        # v1 numbers its windows differently (no +1 window offset, no z): the PARTITION is the same
        a, b = info[f"batch_win_inds_shift{i}"], z[f"bwi{i}"]
        assert torch.equal(torch.unique(a, return_inverse=True)[1], torch.unique(b, return_inverse=True)[1])
    out = O.sstv2_forward(info, _w(z, "w."), [4, 4], 2, "gelu", {})
    torch.testing.assert_close(out, z["bev_rows"], rtol=1e-4, atol=1e-5)
    assert abs(float(out.abs().double().sum()) - float(z["bev_abs_sum"])) < 1e-3 * float(z["bev_abs_sum"])


def test_vfe_small_golden():
    z = _load("vfe_small.npz")
    vf, vc = O.dynamic_vfe_forward(z["points"], z["coors"], _w(z, "w."), VS, RNG, 2)
    assert torch.equal(vc, z["vcoors"])
    torch.testing.assert_close(vf, z["feats"], rtol=1e-5, atol=1e-5)


FSDV2 = dict(vs=(0.5, 0.5, 0.5), rng=[-40, -40, -2, 40, 40, 4], target=[12, 160, 160],
             vfe=dict(num_layers=2, with_cluster_center=True, with_voxel_center=True, rel_dist_scaler=10.0, eps=1e-3))


def _fsdv2_inputs(z, with_ms):
    sampled = {k[len("sampled."):]: v for k, v in z.items() if k.startswith("sampled.")}
    origin = {k[len("origin."):]: v for k, v in z.items() if k.startswith("origin.")}
    ms = None
    if with_ms:
        ms = dict(levels=[(z[f"ms{i}.features"], z[f"ms{i}.indices"], z[f"ms{i}.shape"].tolist()) for i in range(3)],
                  target_sparse_shape=FSDV2["target"], fusion_mode="avg")
    return sampled, origin, ms


@pytest.mark.parametrize("tag", ["plain", "ms"])
def test_fsdv2_front_golden(tag):
    """FSDv2 virtual-voxel front (config 5): the oracle against what SingleStageFSDV2.extract_feat itself hands to its backbone
    (reference source run through oracle/ref_shim.reference_methods, fixture tests/golden/fsdv2_front_*.npz)."""
    z = _load(f"fsdv2_front_{tag}.npz")
    sampled, origin, ms = _fsdv2_inputs(z, tag == "ms")
    out = O.fsdv2_front(sampled, origin, _w(z, "w."), FSDV2["vs"], FSDV2["rng"], FSDV2["vfe"], ms=ms)
    assert torch.equal(out["coors"], z["coors"])
    assert torch.equal(out["voxel_coors"], z["backbone_coors"])
    torch.testing.assert_close(out["voxel_feats"], z["backbone_feats"], rtol=1e-5, atol=1e-5)
    # the part after the (identity) backbone: virtual voxels only
    vf, vc = out["voxel_feats"], out["voxel_coors"]
    if out["singlescale_mask"] is not None:
        vf, vc = vf[out["singlescale_mask"]], vc[out["singlescale_mask"]]
    assert torch.equal(vc[out["virtual_mask"]], z["virtual_coors"])
    torch.testing.assert_close(vf[out["virtual_mask"]], z["virtual_feats"], rtol=1e-5, atol=1e-5)


# ---- sparse convolution fixtures produced by the reference's vendored spconv v1 (oracle/make_golden.py spconv) ------
def _sp_sorted(f, c):
    c = c.long()
    order = torch.argsort(((c[:, 0] * 4096 + c[:, 1]) * 4096 + c[:, 2]) * 4096 + c[:, 3])
    return f[order], c[order].int()


def test_spconv_layers_golden():
    from oracle import spconv_oracle as SO
    z = np.load(os.path.join(G, "spconv_layers.npz"))
    feats, coors, shape = torch.from_numpy(z["l_feats"]), torch.from_numpy(z["l_coors"]), z["l_shape"].tolist()
    for name in "abc":
        cfg = z[f"conv_{name}_cfg"].tolist()
        ks, st, pd = cfg[0:3], cfg[3:6], cfg[6:9]
        w = torch.from_numpy(z[f"conv_{name}_w"])
        of, oc, oshape = SO.sparse_conv(feats, coors, 2, shape, w, st, pd)
        assert oshape == z[f"conv_{name}_shape"].tolist()
        assert torch.equal(oc, torch.from_numpy(z[f"conv_{name}_coors"]).int())
        torch.testing.assert_close(of, torch.from_numpy(z[f"conv_{name}_out"]), rtol=1e-4, atol=1e-5)
        nbr = SO.neighbour_table(coors, oc, 2, shape, ks, st, pd)
        torch.testing.assert_close(SO.indice_conv(feats, nbr, w), of, rtol=1e-4, atol=1e-5)
        zi = SO.inverse_conv(of, oc, 2, oshape, coors, shape, torch.from_numpy(z[f"inv_{name}_w"]), st, pd)
        torch.testing.assert_close(zi, torch.from_numpy(z[f"inv_{name}_out"]), rtol=1e-4, atol=1e-5)
    o = SO.subm_conv(feats, coors, 2, shape, torch.from_numpy(z["subm_w"]), torch.from_numpy(z["subm_b"]))
    torch.testing.assert_close(o, torch.from_numpy(z["subm_out"]), rtol=1e-4, atol=1e-5)


def test_spconv_unet_golden():
    from oracle import spconv_oracle as SO
    z = np.load(os.path.join(G, "spconv_unet.npz"))
    sd = {k[len("unet_sd."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("unet_sd.")}
    feats, coors = torch.from_numpy(z["unet_feats"]), torch.from_numpy(z["unet_coors"])
    U = SO.SP_UNET
    f, c, ms = SO.sparse_unet_forward(sd, feats, coors, 2, U["sparse_shape"], U["encoder_channels"], U["encoder_paddings"], U["decoder_channels"],
                                      U["decoder_paddings"], return_multiscale=True)
    torch.testing.assert_close(f, torch.from_numpy(z["unet_out"]), rtol=1e-3, atol=1e-4)
    for i, (mf, mc) in enumerate(ms):
        mf, mc = _sp_sorted(mf, mc)
        assert torch.equal(mc, torch.from_numpy(z[f"unet_ms{i}_c"]).int())
        torch.testing.assert_close(mf, torch.from_numpy(z[f"unet_ms{i}_f"]), rtol=1e-3, atol=1e-4)
    sd = {k[len("mixer_sd."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("mixer_sd.")}
    M = SO.SP_MIXER
    f, c = SO.sparse_unet_forward(sd, torch.from_numpy(z["mixer_feats"]), torch.from_numpy(z["mixer_coors"]), 3, M["sparse_shape"],
                                  M["encoder_channels"], M["encoder_paddings"], M["decoder_channels"], M["decoder_paddings"], mixer_out=True)
    torch.testing.assert_close(f, torch.from_numpy(z["mixer_out"]), rtol=1e-3, atol=1e-4)


def test_fsd_cluster_golden():
    """the grouping oracle against outputs of the reference's own source (oracle/make_golden.py fsd)"""
    from oracle import fsd_oracle as FO
    z = np.load(os.path.join(G, "fsd_cluster.npz"))
    for tag in ("", "_mixed"):
        pts, bidx = torch.from_numpy(z["cc_points" + tag]), torch.from_numpy(z["cc_batch" + tag])
        for d in (0.1, 0.6, 2.0):
            ref = torch.from_numpy(z[f"cc_labels{tag}_{d}"])
            assert torch.equal(FO.find_connected_components(pts, bidx, d), ref)
            assert torch.equal(FO.connected_components_large(pts, bidx, d), ref)
    vs = dict(Car=(0.3, 0.3, 6), Cyclist=(0.2, 0.2, 6), Pedestrian=(0.05, 0.05, 6))
    dist = dict(Car=0.6, Cyclist=0.4, Pedestrian=0.1)
    for i, name in enumerate(('Car', 'Cyclist', 'Pedestrian')):
        inds, mask = FO.cluster_assigner_single_class(torch.from_numpy(z[f"ca_points{i}"]), torch.from_numpy(z[f"ca_batch{i}"]), vs[name], 2,
                                                      [-80, -80, -2, 80, 80, 4], dist[name])
        assert torch.equal(mask, torch.from_numpy(z[f"ca_mask{i}"]))
        assert torch.equal(inds.long(), torch.from_numpy(z[f"ca_inds{i}"])[:, 1:].long())


def test_fsdv2_extract_feat_with_mixer_golden():
    """BASELINE config 5 end to end: front (oracle) -> VirtualVoxelMixer (spconv oracle) -> virtual voxels, against the reference's own
    extract_feat running its own VirtualVoxelMixer over its vendored spconv (fixture fsdv2_front_mixer.npz)."""
    from oracle import spconv_oracle as SO
    z = _load("fsdv2_front_mixer.npz")
    sampled, origin, ms = _fsdv2_inputs(z, True)
    out = O.fsdv2_front(sampled, origin, _w(z, "w."), FSDV2["vs"], FSDV2["rng"], FSDV2["vfe"], ms=ms)
    torch.testing.assert_close(out["voxel_feats"], z["backbone_feats"], rtol=1e-5, atol=1e-5)
    M = SO.FSDV2_MIXER
    f, c = SO.sparse_unet_forward(_w(z, "mix."), out["voxel_feats"], out["voxel_coors"].int(), int(z["batch_size"]), M["sparse_shape"],
                                  M["encoder_channels"], M["encoder_paddings"], M["decoder_channels"], M["decoder_paddings"], mixer_out=True)
    f, c = f[out["singlescale_mask"]], c[out["singlescale_mask"]]
    assert torch.equal(c[out["virtual_mask"]].long(), z["virtual_coors"].long())
    torch.testing.assert_close(f[out["virtual_mask"]], z["virtual_feats"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("seed", range(6))
def test_spconv_oracle_table_form_equals_dense_form(seed):
    """the two independent forms of the sparse-convolution checker agree on random geometries (kernel 1-3, stride 1-2, asymmetric padding,
    ragged batches): the neighbour-table form (what the CUDA path is compared with launch by launch) and the dense conv3d /
    conv_transpose3d form (what is pinned to the reference's spconv)"""
    from oracle import spconv_oracle as SO
    g = torch.Generator().manual_seed(100 + seed)
    shape = [int(v) for v in torch.randint(5, 14, (3,), generator=g)]
    ks = [int(v) for v in torch.randint(1, 4, (3,), generator=g)]
    st = [int(v) for v in torch.randint(1, 3, (3,), generator=g)]
    pd = [int(torch.randint(0, k, (1,), generator=g)) for k in ks]
    B = 1 + seed % 3
    feats, coors = SO.synth_sparse(seed, B, shape, 40 + 25 * seed, 4, clustered=bool(seed % 2))
    if seed == 5:   # one empty sample in the middle of the batch
        keep = coors[:, 0] != 1
        feats, coors = feats[keep], coors[keep]
    cin, cout = 4, 6
    w = torch.randn((*ks, cin, cout), generator=g)
    of, oc, oshape = SO.sparse_conv(feats, coors, B, shape, w, st, pd)
    nbr = SO.neighbour_table(coors, oc, B, shape, ks, st, pd)
    torch.testing.assert_close(SO.indice_conv(feats, nbr, w), of, rtol=1e-4, atol=1e-5)
    assert bool((nbr >= 0).any(1).all()), "every active output cell has at least one contributing input"
    # inverse conv through the transposed table
    kv = nbr.shape[1]
    inv = torch.full((coors.shape[0], kv), -1, dtype=torch.int32)
    o, k = torch.nonzero(nbr >= 0, as_tuple=True)
    inv[nbr[o, k].long(), k] = o.int()
    wi = torch.randn((*ks, cout, cin), generator=g)
    torch.testing.assert_close(SO.indice_conv(of, inv, wi), SO.inverse_conv(of, oc, B, oshape, coors, shape, wi, st, pd), rtol=1e-4, atol=1e-5)
    # SubM (odd kernels only: spconv centres them)
    ko = [k | 1 for k in ks]
    ws = torch.randn((*ko, cin, cout), generator=g)
    nb = SO.neighbour_table(coors, coors, B, shape, ko, [1, 1, 1], [k // 2 for k in ko])
    torch.testing.assert_close(SO.indice_conv(feats, nb, ws), SO.subm_conv(feats, coors, B, shape, ws), rtol=1e-4, atol=1e-5)
