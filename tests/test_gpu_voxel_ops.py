"""GPU parity: voxel ops (V1, V2, V3, V5, B2) through the C ABI vs the CPU oracle.  Integer outputs bit-exact."""
import pytest
import torch

from oracle import sst_oracle as O

pytestmark = pytest.mark.gpu

VS = (0.32, 0.32, 6)
RNG = [-74.88, -74.88, -2, 74.88, 74.88, 4]


@pytest.mark.parametrize("P,F", [(1, 3), (1000, 4), (20000, 3), (150000, 5)])
def test_dynamic_voxelize_bitexact(cuda, P, F):
    from sst_b200 import ops
    pts = O.synth_frame(1000 + P, P, extra_dims=F - 3)
    # push some points outside the range to exercise the clamp
    pts[::7, 0] += 200.0
    pts[::11, 1] -= 300.0
    ref = O.dynamic_voxelize(pts, VS, RNG)
    got = ops.Voxelization(VS, RNG, -1, (-1, -1))(pts.to(cuda))
    assert got.dtype == torch.int32 and torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("voxel_size,rng", [((0.25, 0.25, 0.2), [-80, -80, -2, 80, 80, 4]),
                                             ((0.1, 0.1, 0.15), [-51.2, -51.2, -5, 51.2, 51.2, 3])])
def test_dynamic_voxelize_other_grids(cuda, voxel_size, rng):
    from sst_b200 import ops
    pts = O.synth_frame(7, 30000)
    ref = O.dynamic_voxelize(pts, voxel_size, rng)
    coors = torch.zeros((pts.shape[0], 3), dtype=torch.int32, device=cuda)
    ops.dynamic_voxelize(pts.to(cuda), coors, voxel_size, rng, 3)
    assert torch.equal(coors.cpu(), ref)


def _dyn_scatter_case(seed, P, C, lo=-1, hi=20):
    g = torch.Generator().manual_seed(seed)
    feats = torch.rand(P, C, generator=g) * 100 - 50
    coors = torch.randint(lo, hi, (P, 3), generator=g, dtype=torch.int32)
    return feats, coors


@pytest.mark.parametrize("reduce", ["mean", "max", "sum"])
@pytest.mark.parametrize("P,C,lo", [(200000, 3, -1), (5000, 64, -1), (3000, 128, 0), (17, 5, 0), (1, 4, 0)])
def test_dynamic_scatter_forward(cuda, reduce, P, C, lo):
    """Restates tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:56-84 (seeded) and adds the
    no-invalid-row case (lo=0) that exposes the reference's unconditional first-row removal."""
    from sst_b200 import ops
    feats, coors = _dyn_scatter_case(P + C, P, C, lo)
    r_f, r_c, r_m, r_n = O.dynamic_point_to_voxel_forward(feats, coors, reduce)
    g_f, g_c, g_m, g_n = ops.dynamic_point_to_voxel_forward(feats.to(cuda), coors.to(cuda), reduce)
    assert torch.equal(g_c.cpu(), r_c)
    assert torch.equal(g_m.cpu(), r_m)
    assert torch.equal(g_n.cpu(), r_n)
    if reduce == "max":
        assert torch.equal(g_f.cpu(), r_f)
    else:
        torch.testing.assert_close(g_f.cpu(), r_f, rtol=1e-5, atol=1e-3 if reduce == "sum" else 1e-4)


def test_dynamic_scatter_empty_and_all_invalid(cuda):
    """test_dynamic_scatter.py:22-53: empty input keeps shapes; all-invalid coors give zero grads."""
    from sst_b200 import ops
    ds = ops.DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], True)
    f = torch.rand(0, 3, device=cuda, requires_grad=True)
    c = torch.randint(0, 5, (0, 3), dtype=torch.int32, device=cuda)
    vf, vc = ds(f, c)
    assert vf.shape == (0, 3) and vc.shape == (0, 3)
    f = torch.rand(200, 3, device=cuda, requires_grad=True)
    c = -torch.ones((200, 3), dtype=torch.int32, device=cuda)
    vf, vc = ds(f, c)
    assert vf.shape == (0, 3) and vc.shape == (0, 3)
    vf.sum().backward()
    assert (f.grad == 0).all()


@pytest.mark.parametrize("reduce", ["mean", "max", "sum"])
def test_dynamic_scatter_backward(cuda, reduce):
    from sst_b200 import ops
    P, C = 4000, 7
    feats, coors = _dyn_scatter_case(3, P, C)
    # force ties for max so the lowest-index rule is exercised
    feats = (feats / 10).round()
    r_f, r_c, r_m, r_n = O.dynamic_point_to_voxel_forward(feats, coors, reduce)
    gr = torch.randn(r_f.shape, generator=torch.Generator().manual_seed(1))
    ref = O.dynamic_point_to_voxel_backward(gr, feats, r_f, r_m, r_n, reduce)
    f = feats.to(cuda).requires_grad_(True)
    vf, vc = ops.dynamic_scatter(f, coors.to(cuda), reduce)
    vf.backward(gr.to(cuda))
    torch.testing.assert_close(f.grad.cpu(), ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("batch", [1, 3])
def test_dynamic_scatter_module_batched(cuda, batch):
    from sst_b200 import ops
    pts, cs = [], []
    for b in range(batch):
        p = O.synth_frame(50 + b, 20000)
        c = O.dynamic_voxelize(p, VS, RNG)
        pts.append(p)
        cs.append(torch.nn.functional.pad(c, (1, 0), value=b))
    pts, cs = torch.cat(pts), torch.cat(cs)
    for avg in (True, False):
        rf, rc = O.dynamic_scatter_module(pts, cs, avg)
        gf, gc = ops.DynamicScatter(VS, RNG, avg)(pts.to(cuda), cs.to(cuda))
        assert torch.equal(gc.cpu(), rc)
        torch.testing.assert_close(gf.cpu(), rf, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("k", [1, 3, 4])
def test_unique_rows_and_scatter_v2(cuda, k):
    from sst_b200 import ops
    g = torch.Generator().manual_seed(k)
    P = 50000
    coors = torch.stack([torch.randint(-3, 40 + 10 * d, (P,), generator=g) for d in range(k)], 1)
    feat = torch.randn(P, 16, generator=g)
    ru, ri, rc = torch.unique(coors, dim=0, return_inverse=True, return_counts=True)
    gu, gi, gc = ops.unique_rows(coors.to(cuda), return_counts=True)
    assert torch.equal(gu.cpu(), ru) and torch.equal(gi.cpu(), ri) and torch.equal(gc.cpu(), rc)
    for mode in ("avg", "max", "sum"):
        rf, rn, rinv = O.scatter_v2(feat, coors, mode)
        gf, gn, ginv = ops.scatter_v2(feat.to(cuda), coors.to(cuda), mode)
        assert torch.equal(gn.cpu(), rn) and torch.equal(ginv.cpu(), rinv)
        torch.testing.assert_close(gf.cpu(), rf, rtol=1e-5, atol=1e-4)
    rf, rn, rinv = O.scatter_v2(feat, coors, "max", min_points=3)
    gf, gn, ginv = ops.scatter_v2(feat.to(cuda), coors.to(cuda), "max", min_points=3)
    assert torch.equal(gn.cpu(), rn) and torch.equal(ginv.cpu(), rinv) and torch.equal(gf.cpu(), rf)


def test_scatter_v2_grad(cuda):
    from sst_b200 import ops
    g = torch.Generator().manual_seed(0)
    P = 3000
    coors = torch.randint(0, 12, (P, 3), generator=g)
    feat = torch.randn(P, 8, generator=g)
    new_coors, inv = torch.unique(coors, dim=0, return_inverse=True)
    for mode in ("mean", "max", "sum"):
        fr = feat.clone().requires_grad_(True)
        O.segment_reduce(fr, inv, mode, new_coors.shape[0]).square().sum().backward()
        fg = feat.to(cuda).requires_grad_(True)
        out, _, _ = ops.scatter_v2(fg, coors.to(cuda), mode, unq_inv=inv.to(cuda), new_coors=new_coors.to(cuda))
        out.square().sum().backward()
        torch.testing.assert_close(fg.grad.cpu(), fr.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("n,ngroups", [(1, 1), (1000, 7), (30000, 1300), (30000, 3), (5, 100000)])
def test_ingroup_indices(cuda, n, ngroups):
    from sst_b200 import ops
    g = torch.Generator().manual_seed(n + ngroups)
    grp = torch.randint(0, ngroups, (n,), generator=g)
    ref = O.ingroup_indices(grp)
    got = ops.get_inner_win_inds(grp.to(cuda))
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("P,F,mp,mv", [(20000, 4, 32, 20000), (20000, 4, 5, 300), (3000, 3, 1, 50), (1, 4, 3, 2)])
def test_hard_voxelize_bitexact(cuda, P, F, mp, mv):
    """`Voxelization(max_num_points > 0)` (hard voxelisation): voxel order = first appearance, first
    max_points points per voxel in input order, max_voxels cap - bit-exact against the oracle (itself pinned to the reference
    C++ in tests/test_ref_voxel_layer.py)."""
    from sst_b200 import ops
    pts = O.synth_frame(9 + P, P, extra_dims=F - 3)
    pts[::13, 0] += 500.0
    ov, oc, on = O.hard_voxelize(pts, VS, RNG, mp, mv)
    gv, gc, gn = ops.Voxelization(VS, RNG, mp, mv).eval()(pts.to(cuda))
    assert torch.equal(gc.cpu(), oc) and torch.equal(gn.cpu(), on) and torch.equal(gv.cpu(), ov)


def test_hard_voxelize_empty(cuda):
    from sst_b200 import ops
    v, c, n = ops.Voxelization(VS, RNG, 8, 100).eval()(torch.zeros(0, 4, device=cuda))
    assert v.shape == (0, 8, 4) and c.shape == (0, 3) and n.shape == (0,)


@pytest.mark.parametrize("name", ["roomy", "capped"])
def test_hard_voxelize_reference_golden(cuda, name):
    """Fixture written by the reference's own C++ hard_voxelize (tests/golden/hard_voxelize.npz, oracle/make_golden.py)."""
    import os
    import numpy as np
    from sst_b200 import ops
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hard_voxelize.npz"))
    mp, mv = [int(v) for v in z[f"cfg_{name}"]]
    v, c, n = ops.Voxelization(VS, RNG, mp, mv).eval()(torch.from_numpy(z["points"]).to(cuda))
    assert torch.equal(c.cpu(), torch.from_numpy(z[f"coors_{name}"]))
    assert torch.equal(n.cpu(), torch.from_numpy(z[f"npts_{name}"]))
    assert torch.equal(v.cpu(), torch.from_numpy(z[f"voxels_{name}"]))
