"""TEST INFRASTRUCTURE ONLY - never imported by the product package.

CPU restatement of FSD's instance grouping, following
  mmdet3d/models/detectors/single_stage_fsd.py:28-32   filter_almost_empty
  mmdet3d/models/detectors/single_stage_fsd.py:47-81   find_connected_componets / _single_batch (dense xy distance matrix < dist,
                                                       scipy.sparse.csgraph.connected_components, running base over the samples)
  mmdet3d/models/detectors/single_stage_fsd.py:144-151 modify_cluster_by_class
  mmdet3d/models/detectors/single_stage_fsd.py:922-999 ClusterAssigner.forward / forward_single_class
Pinned: tests/test_oracle_fsd_vs_reference.py runs these against the reference's own functions / class compiled straight from that
source file (oracle/ref_shim.reference_functions), and tests/golden/fsd_cluster.npz holds the reference's outputs for the GPU box.
`connected_components_large` is the same labelling for sizes where the n x n matrix does not fit: candidate pairs from a k-d tree,
the SAME fp32 distance test on them, scipy on the sparse graph; checked against the dense form in the CPU tests."""
import numpy as np
import torch
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

from oracle import sst_oracle as O


def filter_almost_empty(coors, min_points):
    _, inv, cnt = torch.unique(coors, return_inverse=True, return_counts=True, dim=0)
    return cnt[inv] >= min_points


def find_connected_components(points, batch_idx, dist):
    """single_stage_fsd.py:47-68"""
    bsz = int(batch_idx.max()) + 1
    base = 0
    out = torch.zeros_like(batch_idx) - 1
    for i in range(bsz):
        m = batch_idx == i
        if m.any():
            p = points[m]
            d = p[:, None, :2] - p[None, :, :2]
            adj = ((d ** 2).sum(2) ** 0.5 < dist).numpy()
            c = torch.from_numpy(connected_components(adj, directed=False)[1]).to(out.dtype) + base
            base = int(c.max()) + 1
            out[m] = c
    return out


def connected_components_large(points, batch_idx, dist):
    """same labels without the n x n matrix (k-d tree candidates + the reference's fp32 distance test)"""
    from scipy.spatial import cKDTree
    bsz = int(batch_idx.max()) + 1
    base = 0
    out = torch.zeros_like(batch_idx) - 1
    for i in range(bsz):
        m = batch_idx == i
        if m.any():
            p = points[m][:, :2].float()
            pairs = cKDTree(p.double().numpy()).query_pairs(float(dist) * 1.001 + 1e-6, output_type="ndarray")
            a, b = torch.from_numpy(pairs[:, 0]).long(), torch.from_numpy(pairs[:, 1]).long()
            d = p[a] - p[b]
            ok = ((d ** 2).sum(1) ** 0.5 < dist).numpy()
            n = p.shape[0]
            g = coo_matrix((np.ones(int(ok.sum()), dtype=bool), (pairs[ok, 0], pairs[ok, 1])), shape=(n, n))
            c = torch.from_numpy(connected_components(g, directed=False)[1]).to(out.dtype) + base
            base = int(c.max()) + 1
            out[m] = c
    return out


def cluster_assigner_single_class(points, batch_idx, cluster_voxel_size, min_points, point_cloud_range, dist, single_batch=False):
    """ClusterAssigner.forward_single_class (single_stage_fsd.py:954-996); single_batch = the eval branch (:989)"""
    batch_idx = batch_idx.int()
    vs = torch.tensor(cluster_voxel_size)
    lo = torch.tensor(point_cloud_range)[:3]
    coors = torch.div(points - lo[None], vs[None], rounding_mode='floor').int()
    coors = torch.cat([batch_idx[:, None], coors], 1)
    valid = filter_almost_empty(coors, min_points)
    if not valid.any():
        valid = ~valid
    points, batch_idx, coors = points[valid], batch_idx[valid], coors[valid]
    centres, vcoors, inv = O.scatter_v2(points, coors, 'avg')
    labels = find_connected_components(centres, torch.zeros_like(vcoors[:, 0]) if single_batch else vcoors[:, 0], dist)
    return torch.stack([batch_idx, labels[inv].int()], 1), valid


def synth_centres(seed, batch_size, n_per_sample, spread=40.0, blob=0.35, blobs=60):
    """voted centres: tight blobs (instances) + a few strays, samples interleaved on request by the caller"""
    g = torch.Generator().manual_seed(seed)
    pts, bidx = [], []
    for b in range(batch_size):
        c = (torch.rand((blobs, 2), generator=g) - 0.5) * 2 * spread
        pick = torch.randint(0, blobs, (n_per_sample,), generator=g)
        xy = c[pick] + torch.randn((n_per_sample, 2), generator=g) * blob
        z = torch.rand((n_per_sample, 1), generator=g) * 4 - 2
        pts.append(torch.cat([xy, z], 1))
        bidx.append(torch.full((n_per_sample,), b, dtype=torch.int32))
    return torch.cat(pts), torch.cat(bidx)
