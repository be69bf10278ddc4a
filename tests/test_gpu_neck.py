"""GPU parity: FSD's voxel -> point neck through the registered module / C ABI vs the oracle."""
import pytest
import torch

from oracle import sst_oracle as O

pytestmark = pytest.mark.gpu

VS = (0.32, 0.32, 6)
RNG = [-74.88, -74.88, -2, 74.88, 74.88, 4]


@pytest.mark.parametrize("N,M,C,with_xyz,norm", [(150000, 30000, 128, True, False), (5000, 700, 16, True, True), (5000, 700, 37, False, False),
                                                (33, 5, 3, True, False)])
def test_voxel2point_neck_exact(cuda, N, M, C, with_xyz, norm):
    """Gather + compaction + centre offset: mask, row count and order bit-exact; the features are copies, the offsets the same
    fp32 expression as the reference ((c + 0.5) * vs + min with two roundings)."""
    from sst_b200.neck_modules import Voxel2PointScatterNeck
    g = torch.Generator().manual_seed(N + C)
    pts = O.synth_frame(3, N, extra_dims=1)
    coors = torch.nn.functional.pad(O.dynamic_voxelize(pts, VS, RNG), (1, 0), value=0).long()
    vf = torch.randn(M, C, generator=g)
    vf[::7] = -1.0
    inds = torch.randint(0, M, (N,), generator=g)
    o_out, o_mask = O.voxel2point_neck(pts, coors, vf, inds, VS, RNG, with_xyz, norm)
    neck = Voxel2PointScatterNeck(point_cloud_range=RNG, voxel_size=VS, with_xyz=with_xyz, normalize_local_xyz=norm).eval()
    with torch.no_grad():
        out, mask = neck(pts.to(cuda), coors.to(cuda), vf.to(cuda), inds.to(cuda))
    assert mask.dtype == torch.bool and torch.equal(mask.cpu(), o_mask)
    assert torch.equal(out.cpu(), o_out)


def test_voxel2point_neck_bad_index_fails(cuda):
    from sst_b200._lib import SSTB200Error
    from sst_b200.neck_modules import Voxel2PointScatterNeck
    neck = Voxel2PointScatterNeck(point_cloud_range=RNG, voxel_size=VS)
    with pytest.raises(SSTB200Error), torch.no_grad():
        neck(torch.zeros(4, 3, device=cuda), torch.zeros(4, 4, dtype=torch.long, device=cuda), torch.zeros(2, 8, device=cuda),
             torch.tensor([0, 1, 2, 0], device=cuda))


def test_reorder(cuda):
    """VoteSegmentor.reorder (single_stage_fsd.py:253-266)."""
    from sst_b200.neck_modules import reorder
    g = torch.Generator().manual_seed(0)
    n = 50
    shuffle = torch.randperm(n, generator=g)
    keep = torch.sort(torch.randperm(n, generator=g)[:35]).values
    data = torch.randn(35, 6, generator=g)
    temp = -torch.ones(n, 6)
    ref = -torch.ones(n, 6)
    temp[keep] = data
    ref[shuffle] = temp
    assert torch.equal(reorder(data.to(cuda), shuffle.to(cuda), keep.to(cuda)).cpu(), ref)


@pytest.mark.parametrize("name,xyz,norm", [("xyz", True, False), ("xyznorm", True, True), ("noxyz", False, False)])
def test_voxel2point_neck_reference_golden(cuda, name, xyz, norm):
    """Fixture written by the unmodified reference class (tests/golden/neck_small.npz)."""
    import os
    import numpy as np
    from sst_b200.neck_modules import Voxel2PointScatterNeck
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "neck_small.npz"))
    t = lambda k: torch.from_numpy(z[k]).to(cuda)
    neck = Voxel2PointScatterNeck(point_cloud_range=RNG, voxel_size=VS, with_xyz=xyz, normalize_local_xyz=norm).eval()
    with torch.no_grad():
        out, mask = neck(t("points"), t("coors"), t("voxel_feats"), t("inds"))
    assert torch.equal(mask.cpu(), torch.from_numpy(z[f"mask_{name}"]))
    assert torch.equal(out.cpu(), torch.from_numpy(z[f"out_{name}"]))
