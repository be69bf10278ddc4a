"""GPU: FSD instance grouping through the C ABI: connected-component labels are BIT-EXACT against
  * tests/golden/fsd_cluster.npz = outputs of the reference's own find_connected_componets / ClusterAssigner source,
  * the oracle (scipy restatement pinned to that source) at sizes the dense n x n matrix cannot reach, via its k-d tree form."""
import os

import numpy as np
import pytest
import torch

from oracle import fsd_oracle as FO

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_connected_components_reference_golden(cuda):
    from sst_b200 import fsd_modules as FM
    z = np.load(os.path.join(G, "fsd_cluster.npz"))
    for tag in ("", "_mixed"):
        pts, bidx = torch.from_numpy(z["cc_points" + tag]).to(cuda), torch.from_numpy(z["cc_batch" + tag]).to(cuda)
        for d in (0.1, 0.6, 2.0):
            ref = torch.from_numpy(z[f"cc_labels{tag}_{d}"])
            labels, num = FM.connected_components(pts, bidx, d)
            assert num == int(ref.max()) + 1
            assert torch.equal(labels.cpu(), ref.int())
            # explicit bounds (the ClusterAssigner path) and a grid that clips most centres into its border cells: same labels
            for bounds in (([-80.0, -80.0], [80.0, 80.0]), ([-5.0, -5.0], [5.0, 5.0])):
                l2, n2 = FM.connected_components(pts, bidx, d, batch_size=3, xy_bounds=bounds)
                assert n2 == num and torch.equal(l2, labels)
    # one sample, batch_idx ignored (find_connected_componets_single_batch)
    pts = torch.from_numpy(z["cc_points"])[:700]
    ref = FO.find_connected_components(pts, torch.zeros(700, dtype=torch.int32), 0.6)
    assert torch.equal(FM.find_connected_componets_single_batch(pts.to(cuda), None, 0.6).cpu(), ref)


def test_cluster_assigner_reference_golden(cuda):
    from sst_b200 import fsd_modules as FM
    z = np.load(os.path.join(G, "fsd_cluster.npz"))
    ca = FM.ClusterAssigner(cluster_voxel_size=dict(Car=(0.3, 0.3, 6), Cyclist=(0.2, 0.2, 6), Pedestrian=(0.05, 0.05, 6)), min_points=2,
                            point_cloud_range=[-80, -80, -2, 80, 80, 4], connected_dist=dict(Car=0.6, Cyclist=0.4, Pedestrian=0.1),
                            class_names=['Car', 'Cyclist', 'Pedestrian']).train()
    pts = [torch.from_numpy(z[f"ca_points{i}"]).to(cuda) for i in range(3)]
    bidx = [torch.from_numpy(z[f"ca_batch{i}"]).to(cuda) for i in range(3)]
    inds, masks = ca(pts, bidx, origin_points=[None] * 3)
    for i in range(3):
        assert torch.equal(masks[i].cpu(), torch.from_numpy(z[f"ca_mask{i}"]))
        assert torch.equal(inds[i].cpu().long(), torch.from_numpy(z[f"ca_inds{i}"]).long())


def test_connected_components_large_and_edge_cases(cuda):
    from sst_b200 import _lib as L, fsd_modules as FM
    pts, bidx = FO.synth_centres(21, 4, 15000, spread=70.0, blob=0.4, blobs=900)   # 60k centres: the n x n form would need 3.6e9 cells
    for d in (0.15, 0.6):
        ref = FO.connected_components_large(pts, bidx, d)
        labels, num = FM.connected_components(pts.to(cuda), bidx.to(cuda), d, xy_bounds=([-80.0, -80.0], [80.0, 80.0]))
        assert num == int(ref.max()) + 1
        assert torch.equal(labels.cpu(), ref.int())
    # every centre on one spot: a single component however many share a cell
    same = torch.zeros((5000, 3), device=cuda)
    labels, num = FM.connected_components(same, None, 0.1)
    assert num == 1 and int(labels.abs().sum()) == 0
    # nothing connected: labels = arange
    far = torch.arange(1000, device=cuda, dtype=torch.float32)[:, None].repeat(1, 3) * 5.0
    labels, num = FM.connected_components(far, None, 0.6)
    assert num == 1000 and torch.equal(labels.cpu(), torch.arange(1000, dtype=torch.int32))
    empty, n0 = FM.connected_components(torch.zeros((0, 3), device=cuda), None, 0.6)
    assert n0 == 0 and empty.shape == (0,)
    with pytest.raises(L.SSTB200Error):   # batch index outside [0, batch_size)
        FM.connected_components(far, torch.full((1000,), 3, dtype=torch.int32, device=cuda), 0.6, batch_size=2)
    with pytest.raises(L.SSTB200Error):   # no CPU fallback
        FM.connected_components(far.cpu(), None, 0.6)


def test_grouping_speed_vs_reference_cpu_path(cuda):
    """the reference labels the voted centres on the CPU (dense n x n distance matrix + scipy, single_stage_fsd.py:47-68) - timed here through
    the oracle's restatement of exactly that, beside sstb200_connected_components on the same centres (FSD scale: 3 classes x ~8k voxel
    centres of one sweep); labels identical, timings printed for profiles/"""
    import time
    from sst_b200 import fsd_modules as FM
    pts, bidx = FO.synth_centres(5, 1, 8000, spread=70.0, blob=0.5, blobs=400)
    t0 = time.perf_counter()
    ref = FO.find_connected_components(pts, bidx, 0.6)
    t_cpu = time.perf_counter() - t0
    p, b = pts.to(cuda), bidx.to(cuda)
    bounds = ([-80.0, -80.0], [80.0, 80.0])
    FM.connected_components(p, b, 0.6, batch_size=1, xy_bounds=bounds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        labels, num = FM.connected_components(p, b, 0.6, batch_size=1, xy_bounds=bounds)
    e1.record()
    torch.cuda.synchronize()
    assert torch.equal(labels.cpu(), ref.int())
    print(f"\n[FSD grouping, {pts.shape[0]} centres, dist 0.6 -> {num} components] reference CPU path (n x n matrix + scipy): {t_cpu * 1e3:.1f} ms | "
          f"sst_b200: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us per call incl. the host round trip for the component count")
