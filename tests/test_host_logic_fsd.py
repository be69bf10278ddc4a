"""CPU: host-side logic of sst_b200/fsd_modules.py (ClusterAssigner glue: per-class tables, almost-empty filter, eval / train branch,
class column) with the library calls replaced by the oracle's restatements - no compute call reaches the library."""
import os

import numpy as np
import torch

from oracle import fsd_oracle as FO, sst_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def test_cluster_assigner_glue_against_reference_golden(monkeypatch):
    from sst_b200 import fsd_modules as FM

    def unique_rows(coors, return_counts=False, bounds=None):
        u, inv, cnt = torch.unique(coors, return_inverse=True, return_counts=True, dim=0)
        return (u, inv, cnt) if return_counts else (u, inv)

    def connected_components(points, batch_idx, dist, batch_size=None, xy_bounds=None):
        b = torch.zeros(points.shape[0], dtype=torch.int32) if batch_idx is None else batch_idx.int()
        lab = FO.find_connected_components(points, b, dist)
        return lab, int(lab.max()) + 1

    monkeypatch.setattr(FM, "unique_rows", unique_rows)
    monkeypatch.setattr(FM, "scatter_v2", lambda feat, coors, mode, return_inv=True: O.scatter_v2(feat, coors, mode, return_inv))
    monkeypatch.setattr(FM, "connected_components", connected_components)
    z = np.load(os.path.join(G, "fsd_cluster.npz"))
    ca = FM.ClusterAssigner(cluster_voxel_size=dict(Car=(0.3, 0.3, 6), Cyclist=(0.2, 0.2, 6), Pedestrian=(0.05, 0.05, 6)), min_points=2,
                            point_cloud_range=[-80, -80, -2, 80, 80, 4], connected_dist=dict(Car=0.6, Cyclist=0.4, Pedestrian=0.1),
                            class_names=['Car', 'Cyclist', 'Pedestrian']).train()
    pts = [torch.from_numpy(z[f"ca_points{i}"]) for i in range(3)]
    bidx = [torch.from_numpy(z[f"ca_batch{i}"]) for i in range(3)]
    inds, masks = ca(pts, bidx, origin_points=[None] * 3)
    for i in range(3):
        assert torch.equal(masks[i], torch.from_numpy(z[f"ca_mask{i}"]))
        assert torch.equal(inds[i].long(), torch.from_numpy(z[f"ca_inds{i}"]).long())
    # list-valued tables and the eval (single-sample) branch
    ca2 = FM.ClusterAssigner(cluster_voxel_size=[(0.3, 0.3, 6)], min_points=2, point_cloud_range=[-80, -80, -2, 80, 80, 4],
                             connected_dist=[0.6], class_names=['Car']).eval()
    m0 = bidx[0] == 0
    i1, _ = ca2([pts[0][m0]], [bidx[0][m0]])
    ref, _ = FO.cluster_assigner_single_class(pts[0][m0], bidx[0][m0], (0.3, 0.3, 6), 2, [-80, -80, -2, 80, 80, 4], 0.6, single_batch=True)
    assert torch.equal(i1[0][:, 1:].int(), ref) and int(i1[0][:, 0].abs().sum()) == 0
