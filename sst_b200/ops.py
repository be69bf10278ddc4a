"""Python face of the op surface the reference exports from `mmdet3d.ops`
(mmdet3d/ops/__init__.py:23-30), backed by libsstb200 (CUDA, sm_100a).  Same names, argument meaning
and error behaviour; no CPU path - CPU tensors raise.

    Voxelization / voxelization       ops/voxel/voxelize.py:11-130          (dynamic branch only, V1)
    DynamicScatter / dynamic_scatter  ops/voxel/scatter_points.py:9-110      (V2, V3)
    scatter_v2                        ops/sst/sst_ops.py:151-182             (V5)
    get_inner_win_inds                ops/sst/sst_ops.py:244-264             (B2)
    get_window_coors, make_continuous_inds, get_flat2win_inds(_v2), flat2window(_v2), window2flat(_v2)
                                      ops/sst/sst_ops.py:27-149,266-331      (B1, B4, B6)
    build_mlp, get_activation(_layer) ops/sst/sst_ops.py:334-392             (S3)
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L

_REDUCE = {"sum": 0, "mean": 1, "max": 2}


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.SSTB200Error("sst_b200 ops need CUDA tensors (B200); there is no CPU implementation")


# ----------------------------------------------------------------------------------------------
# V1 voxelization
# ----------------------------------------------------------------------------------------------
def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
    """voxel_layer.dynamic_voxelize (ops/voxel/src/voxelization.h:72-88): fills caller-allocated `coors`."""
    assert NDim == 3
    _need_cuda(points, coors)
    assert points.dtype == torch.float32 and coors.dtype == torch.int32
    assert points.is_contiguous() and coors.is_contiguous(), "points/coors must be contiguous"
    c = L.ctx(points.device)
    L.check(c, L.lib().sstb200_dynamic_voxelize(
        c, points.data_ptr(), points.shape[0], points.shape[1], L.arr(C.c_float, [float(v) for v in voxel_size]),
        L.arr(C.c_float, [float(v) for v in coors_range]), coors.data_ptr()))


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points, max_voxels, NDim=3):
    """voxel_layer.hard_voxelize (ops/voxel/src/voxelization.h:51-70): fills the caller's zero-initialised buffers, returns the
    number of voxels (int, one stream sync like the reference)."""
    assert NDim == 3
    _need_cuda(points, voxels, coors, num_points_per_voxel)
    assert points.dtype == torch.float32 and voxels.dtype == torch.float32
    assert coors.dtype == torch.int32 and num_points_per_voxel.dtype == torch.int32
    for t in (points, voxels, coors, num_points_per_voxel):
        assert t.is_contiguous()
    assert voxels.shape == (max_voxels, max_points, points.shape[1]) and coors.shape == (max_voxels, 3)
    num_dev = torch.empty((1,), dtype=torch.int32, device=points.device)
    num = C.c_int32(0)
    c = L.ctx(points.device)
    L.check(c, L.lib().sstb200_hard_voxelize(
        c, points.data_ptr(), points.shape[0], points.shape[1], L.arr(C.c_float, [float(v) for v in voxel_size]),
        L.arr(C.c_float, [float(v) for v in coors_range]), int(max_points), int(max_voxels), voxels.data_ptr(), coors.data_ptr(),
        num_points_per_voxel.data_ptr(), num_dev.data_ptr(), C.byref(num)))
    return num.value


class _Voxelization(Function):
    """ops/voxel/voxelize.py:11-72."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000):
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
            dynamic_voxelize(points.contiguous(), coors, voxel_size, coors_range, 3)
            return coors
        points = points.contiguous()
        voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
        coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
        num_points_per_voxel = points.new_zeros(size=(max_voxels,), dtype=torch.int)
        voxel_num = hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points, max_voxels, 3)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    """ops/voxel/voxelize.py:77-130."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        grid = torch.round((pcr[3:] - pcr[:3]) / torch.tensor(voxel_size, dtype=torch.float32)).long()
        self.grid_size = grid
        self.pcd_shape = [*grid[:2], 1][::-1]

    def forward(self, input):
        mv = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points, mv)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range="
                f"{self.point_cloud_range}, max_num_points={self.max_num_points}, max_voxels={self.max_voxels})")


# ----------------------------------------------------------------------------------------------
# V2/V3 DynamicScatter
# ----------------------------------------------------------------------------------------------
def dynamic_point_to_voxel_forward(feats, coors, reduce_type, coor_bounds=None):
    """voxel_layer.dynamic_point_to_voxel_forward -> [reduced_feats, out_coors, coors_map, reduce_count].
    `coor_bounds` = (lo[3], hi[3]) of the valid coordinates if the caller knows the grid (saves one reduction
    + sync); otherwise they are measured from `coors`."""
    _need_cuda(feats, coors)
    assert feats.dtype == torch.float32, "fp32 feats only"
    assert coors.dtype == torch.int32 and coors.dim() == 2 and coors.shape[1] == 3
    assert feats.is_contiguous() and coors.is_contiguous(), "feats/coors must be contiguous"
    P, Cc = feats.shape
    dev = feats.device
    if P == 0:
        return [feats.clone().detach(), coors.clone().detach(), coors.new_empty((0,), dtype=torch.int32),
                coors.new_empty((0,), dtype=torch.int32)]
    if coor_bounds is None:
        hi = coors.amax(0).clamp(min=0).tolist()
        lo = [0, 0, 0]
    else:
        lo, hi = coor_bounds
    reduced = torch.empty((P, Cc), dtype=torch.float32, device=dev)
    out_coors = torch.empty((P, 3), dtype=torch.int32, device=dev)
    cmap = torch.empty((P,), dtype=torch.int32, device=dev)
    count = torch.empty((P,), dtype=torch.int32, device=dev)
    num_dev = torch.empty((1,), dtype=torch.int32, device=dev)
    num_host = C.c_int32(0)
    c = L.ctx(dev)
    L.check(c, L.lib().sstb200_dynamic_point_to_voxel_forward(
        c, feats.data_ptr(), coors.data_ptr(), P, Cc, _REDUCE[reduce_type], L.arr(C.c_int32, [int(v) for v in lo]),
        L.arr(C.c_int32, [int(v) for v in hi]), reduced.data_ptr(), out_coors.data_ptr(), cmap.data_ptr(),
        count.data_ptr(), num_dev.data_ptr(), C.byref(num_host)))
    M = num_host.value
    return [reduced[:M], out_coors[:M], cmap, count[:M]]


def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_map, reduce_count,
                                    reduce_type):
    """voxel_layer.dynamic_point_to_voxel_backward: fills caller-allocated grad_feats."""
    _need_cuda(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_map, reduce_count)
    for t in (grad_feats, grad_reduced_feats, feats, reduced_feats, coors_map, reduce_count):
        assert t.is_contiguous()
    c = L.ctx(feats.device)
    L.check(c, L.lib().sstb200_dynamic_point_to_voxel_backward(
        c, grad_feats.data_ptr(), grad_reduced_feats.data_ptr(), feats.data_ptr(), reduced_feats.data_ptr(),
        coors_map.data_ptr(), reduce_count.data_ptr(), feats.shape[0], reduced_feats.shape[0], feats.shape[1],
        _REDUCE[reduce_type]))


class _dynamic_scatter(Function):
    """ops/voxel/scatter_points.py:9-49."""

    @staticmethod
    def forward(ctx, feats, coors, reduce_type="max", coor_bounds=None):
        voxel_feats, voxel_coors, point2voxel_map, voxel_points_count = dynamic_point_to_voxel_forward(
            feats, coors, reduce_type, coor_bounds)
        ctx.reduce_type = reduce_type
        ctx.save_for_backward(feats, voxel_feats, point2voxel_map, voxel_points_count)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        feats, voxel_feats, point2voxel_map, voxel_points_count = ctx.saved_tensors
        grad_feats = torch.empty_like(feats)
        dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats.contiguous(), feats, voxel_feats.contiguous(),
                                        point2voxel_map, voxel_points_count.contiguous(), ctx.reduce_type)
        return grad_feats, None, None, None


dynamic_scatter = _dynamic_scatter.apply


class DynamicScatter(nn.Module):
    """ops/voxel/scatter_points.py:52-110."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points
        # (z, y, x) grid of the voxelisation this layer is paired with (voxelize.py:95-100): the bitmap-rank index is
        # sized from it, so no reduction + host sync over the coordinates is needed.  Coordinates outside this grid
        # (impossible when they come from `Voxelization` with the same range) make the call fail loudly.
        self._bounds = None
        if voxel_size is not None and point_cloud_range is not None:
            r, v = point_cloud_range, voxel_size
            g = [int(math.ceil((r[3 + i] - r[i]) / v[i])) + 1 for i in range(3)]  # ceil-grid of voxelization_cpu.cpp:151-158, +1 slack
            self._bounds = ([0, 0, 0], [g[2], g[1], g[0]])

    def forward_single(self, points, coors):
        reduce = "mean" if self.average_points else "max"
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce, self._bounds)

    def forward(self, points, coors):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        batch_size = int(coors[-1, 0]) + 1
        voxels, voxel_coors = [], []
        for i in range(batch_size):
            inds = torch.where(coors[:, 0] == i)
            voxel, voxel_coor = self.forward_single(points[inds], coors[inds][:, 1:])
            voxel_coors.append(nn.functional.pad(voxel_coor, (1, 0), mode="constant", value=i))
            voxels.append(voxel)
        return torch.cat(voxels, dim=0), torch.cat(voxel_coors, dim=0)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range="
                f"{self.point_cloud_range}, average_points={self.average_points})")


# ----------------------------------------------------------------------------------------------
# V5 scatter_v2
# ----------------------------------------------------------------------------------------------
def unique_rows(coors, return_counts=False, bounds=None):
    """torch.unique(coors, dim=0, return_inverse=True[, return_counts=True]) for int64 [P,k<=4] rows."""
    _need_cuda(coors)
    assert coors.dtype == torch.int64 and coors.dim() == 2 and coors.shape[1] <= 4
    coors = coors.contiguous()
    P, k = coors.shape
    dev = coors.device
    if P == 0:
        e = coors.new_empty((0,))
        return (coors.clone(), e, e.int()) if return_counts else (coors.clone(), e)
    if bounds is None:
        lo, hi = coors.amin(0).tolist(), coors.amax(0).tolist()
    else:
        lo, hi = bounds
    new_coors = torch.empty((P, k), dtype=torch.int64, device=dev)
    inv = torch.empty((P,), dtype=torch.int64, device=dev)
    cnt = torch.empty((P,), dtype=torch.int32, device=dev) if return_counts else None
    num_dev = torch.empty((1,), dtype=torch.int32, device=dev)
    num_host = C.c_int32(0)
    c = L.ctx(dev)
    L.check(c, L.lib().sstb200_unique_rows_i64(
        c, coors.data_ptr(), P, k, L.arr(C.c_int64, [int(v) for v in lo]), L.arr(C.c_int64, [int(v) for v in hi]),
        new_coors.data_ptr(), inv.data_ptr(), L.ptr(cnt), num_dev.data_ptr(), C.byref(num_host)))
    M = num_host.value
    if return_counts:
        return new_coors[:M], inv, cnt[:M].long()
    return new_coors[:M], inv


class _SegmentReduce(Function):
    """torch_scatter.scatter / scatter_max, dim=0 (forward + backward)."""

    @staticmethod
    def forward(ctx, src, index, num_segments, mode, want_argmax=True):
        _need_cuda(src, index)
        assert src.dtype == torch.float32 and index.dtype == torch.int64
        src = src.contiguous()
        index = index.contiguous()
        P, Cc = src.shape
        out = torch.empty((num_segments, Cc), dtype=torch.float32, device=src.device)
        arg = torch.empty((num_segments, Cc), dtype=torch.int64, device=src.device) if mode == "max" and want_argmax else None
        c = L.ctx(src.device)
        L.check(c, L.lib().sstb200_segment_reduce(c, src.data_ptr(), index.data_ptr(), P, Cc, num_segments,
                                                  _REDUCE[mode], out.data_ptr(), L.ptr(arg)))
        ctx.mode = mode
        ctx.P = P
        ctx.save_for_backward(index, arg if arg is not None else index)
        if arg is not None:
            ctx.mark_non_differentiable(arg)
            return out, arg
        return out, index.new_empty(0)

    @staticmethod
    def backward(ctx, g, _=None):
        index, arg = ctx.saved_tensors
        g = g.contiguous()
        if ctx.mode == "max":
            gs = g.new_zeros((ctx.P + 1, g.shape[1]))
            gs.scatter_(0, arg, g)
            return gs[:ctx.P], None, None, None, None
        gv = g[index]
        if ctx.mode == "mean":
            cnt = torch.bincount(index, minlength=g.shape[0]).clamp(min=1)
            gv = gv / cnt[index][:, None].to(g.dtype)
        return gv, None, None, None, None


def segment_reduce(src, index, mode, num_segments=None, want_argmax=True):
    """torch_scatter.scatter(src, index, dim=0, reduce=mode) / scatter_max -> (out, argmax or None).  `want_argmax=False`
    skips the argmax output of 'max' (only valid when no gradient is needed): the kernel then runs one FMNMX per element."""
    if num_segments is None:
        num_segments = int(index.max()) + 1 if index.numel() else 0
    if mode == "max" and not want_argmax and src.requires_grad and torch.is_grad_enabled():
        want_argmax = True
    out, arg = _SegmentReduce.apply(src, index, num_segments, mode, want_argmax)
    return out, (arg if mode == "max" and want_argmax else None)


def scatter_v2(feat, coors, mode, return_inv=True, min_points=0, unq_inv=None, new_coors=None):
    """ops/sst/sst_ops.py:151-182."""
    assert feat.size(0) == coors.size(0)
    if mode == "avg":
        mode = "mean"
    if mode not in ("max", "mean", "sum"):
        raise NotImplementedError
    if unq_inv is None:
        new_coors, unq_inv, unq_cnt = unique_rows(coors.long(), return_counts=True)
    else:
        assert new_coors is not None, "please pass new_coors for interface consistency"
    if min_points > 0:
        cnt_per_point = unq_cnt[unq_inv]
        valid_mask = cnt_per_point >= min_points
        feat = feat[valid_mask]
        coors = coors[valid_mask]
        new_coors, unq_inv, unq_cnt = unique_rows(coors.long(), return_counts=True)
    new_feat, _ = segment_reduce(feat, unq_inv, mode, new_coors.shape[0], want_argmax=False)
    if not return_inv:
        return new_feat, new_coors
    return new_feat, new_coors, unq_inv


# ----------------------------------------------------------------------------------------------
# B2 get_inner_win_inds
# ----------------------------------------------------------------------------------------------
class IngroupIndicesFunction(Function):
    """ops/sst/sst_ops.py:244-262."""

    @staticmethod
    def forward(ctx, group_inds):
        _need_cuda(group_inds)
        assert group_inds.dtype == torch.int64 and group_inds.dim() == 1
        group_inds = group_inds.contiguous()
        out_inds = torch.zeros_like(group_inds) - 1
        n = group_inds.numel()
        if n:
            assert int(group_inds.min()) >= 0, "group ids must be non-negative"
            c = L.ctx(group_inds.device)
            L.check(c, L.lib().sstb200_ingroup_indices(c, group_inds.data_ptr(), n, int(group_inds.max()),
                                                       out_inds.data_ptr()))
        ctx.mark_non_differentiable(out_inds)
        return out_inds

    @staticmethod
    def backward(ctx, g):
        return None


get_inner_win_inds = IngroupIndicesFunction.apply


# ----------------------------------------------------------------------------------------------
# B1/B3/B4 fused window plan (one shift)
# ----------------------------------------------------------------------------------------------
class _WindowCfg(C.Structure):
    _fields_ = [("sparse_shape", C.c_int32 * 3), ("window_shape", C.c_int32 * 3), ("batch_size", C.c_int32),
                ("num_levels", C.c_int32), ("level_id", C.c_int32 * 8), ("level_lo", C.c_int32 * 8),
                ("level_hi", C.c_int32 * 8), ("level_max_tokens", C.c_int32 * 8)]


class _WindowShift(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("batch_win_inds", "coors_in_win", "drop_level", "flat2win_inds", "pos_code",
                                          "tok_win", "tok_inner", "win_offsets", "tok_perm", "win_level", "win_rank",
                                          "counters", "tok_slot", "win_batch")]


L.SIGNATURES["sstb200_window_plan"] = (C.c_int, [L.vp, L.vp, C.c_int, L.vp, C.POINTER(_WindowCfg), C.c_int, L.vp,
                                                 C.POINTER(_WindowShift), L.P_i32])
L.SIGNATURES["sstb200_window_plan_i32"] = (C.c_int, [L.vp, L.vp, C.c_int, L.vp, C.POINTER(_WindowCfg), C.c_int,
                                                     C.POINTER(_WindowShift)])


def _window_shape3(window_shape, sparse_shape):
    if len(window_shape) == 2:
        return (int(window_shape[0]), int(window_shape[1]), int(sparse_shape[-1]))
    return tuple(int(v) for v in window_shape)


def _window_cfg(sparse_shape, window_shape, drop_info, batch_size):
    cfg = _WindowCfg()
    w3 = _window_shape3(window_shape, sparse_shape)
    for i in range(3):
        cfg.sparse_shape[i] = int(sparse_shape[i])
        cfg.window_shape[i] = w3[i]
    cfg.batch_size = int(batch_size)
    keys = list(drop_info.keys())
    assert 1 <= len(keys) <= 8, "at most 8 drop levels supported"
    cfg.num_levels = len(keys)
    for s, k in enumerate(keys):
        lo, hi = drop_info[k]["drop_range"]
        cfg.level_id[s] = int(k)
        cfg.level_lo[s] = int(min(lo, 2 ** 31 - 1))
        cfg.level_hi[s] = int(min(hi, 2 ** 31 - 1))
        cfg.level_max_tokens[s] = int(drop_info[k]["max_tokens"])
    return cfg, keys


class WindowPlan:
    """Device-resident result of sstb200_window_plan for one shift."""
    __slots__ = ("n", "batch_win_inds", "coors_in_win", "drop_level", "flat2win_inds", "pos_code", "tok_win",
                 "tok_inner", "win_offsets", "tok_perm", "win_level", "win_rank", "counters", "tok_slot", "win_batch", "level_keys",
                 "num_windows", "level_windows", "level_tokens", "status")


def window_plan(coors, sparse_shape, window_shape, drop_info, do_shift, batch_size, token_level=None, sync=True):
    """coors [n,4] int64 (b,z,y,x) on CUDA.  Returns a WindowPlan (reference index tensors + window CSR)."""
    _need_cuda(coors)
    assert coors.dtype == torch.int64 and coors.dim() == 2 and coors.shape[1] == 4
    coors = coors.contiguous()
    n, dev = coors.shape[0], coors.device
    cfg, keys = _window_cfg(sparse_shape, window_shape, drop_info, batch_size)
    p = WindowPlan()
    p.n, p.level_keys = n, keys
    i64 = dict(dtype=torch.int64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    p.batch_win_inds = torch.empty((n,), **i64)
    p.coors_in_win = torch.empty((n, 3), **i64)
    p.drop_level = torch.empty((n,), **i64)
    p.flat2win_inds = torch.empty((n,), **i64)
    p.pos_code = torch.empty((n,), **i32)
    p.tok_win = torch.empty((n,), **i32)
    p.tok_inner = torch.empty((n,), **i32)
    p.win_offsets = torch.empty((n + 1,), **i32)
    p.tok_perm = torch.empty((n,), **i32)
    p.win_level = torch.empty((n,), **i32)
    p.win_rank = torch.empty((n,), **i32)
    p.counters = torch.zeros((20,), **i32)
    p.tok_slot = torch.empty((n,), **i32)
    p.win_batch = torch.empty((n + 16,), **i32)   # int4 records, n/112 + 2 of them
    out = _WindowShift(*[getattr(p, k).data_ptr() for k, _ in _WindowShift._fields_])
    status = (C.c_int32 * 18)()
    if token_level is not None:
        assert token_level.dtype == torch.int64 and token_level.is_cuda and token_level.is_contiguous()
    c = L.ctx(dev)
    L.check(c, L.lib().sstb200_window_plan(c, coors.data_ptr(), n, None, C.byref(cfg), int(bool(do_shift)),
                                           L.ptr(token_level), C.byref(out), status if sync else None))
    if sync:
        p.status = status[0]
        if status[0] & 1:
            raise L.SSTB200Error("window_plan: voxel coordinate outside batch_size/sparse_shape window grid")
        if status[0] & 2:
            raise AssertionError("a window's token count is not covered by any drop_range (reference asserts too)")
        p.num_windows = status[1]
        p.level_windows = list(status[2:2 + len(keys)])
        p.level_tokens = list(status[10:10 + len(keys)])
    return p


@torch.no_grad()
def get_window_coors(coors, sparse_shape, window_shape, do_shift):
    """ops/sst/sst_ops.py:266-314."""
    assert sparse_shape[2] < sparse_shape[0], "Usually holds... in case of wrong order"
    n = coors.shape[0]
    if n == 0:
        return coors.new_empty((0,)), coors.new_empty((0, 3))
    B = int(coors[:, 0].max()) + 1
    p = window_plan(coors.long(), sparse_shape, window_shape, {0: dict(max_tokens=1, drop_range=(0, 2 ** 31 - 1))},
                    do_shift, B)
    return p.batch_win_inds, p.coors_in_win


@torch.no_grad()
def make_continuous_inds(inds):
    """ops/sst/sst_ops.py:316-331."""
    if inds.numel() == 0:
        return inds.clone()
    return unique_rows(inds.long().view(-1, 1))[1].to(inds.dtype)


@torch.no_grad()
def get_flat2win_inds(batch_win_inds, voxel_drop_lvl, drop_info, debug=True):
    """ops/sst/sst_ops.py:27-64 (compat path; the fused path is window_plan)."""
    d = {}
    for dl in drop_info:
        dl_mask = voxel_drop_lvl == dl
        if not dl_mask.any():
            continue
        conti = make_continuous_inds(batch_win_inds[dl_mask])
        inner = get_inner_win_inds(conti)
        d[dl] = (conti * drop_info[dl]["max_tokens"] + inner, torch.where(dl_mask))
    return d


def get_flat2win_inds_v2(batch_win_inds, voxel_drop_lvl, drop_info, debug=True):
    d = get_flat2win_inds(batch_win_inds, voxel_drop_lvl, drop_info, debug)
    d["voxel_drop_level"] = voxel_drop_lvl
    d["batching_info"] = drop_info
    return d


def flat2window(feat, voxel_drop_lvl, flat2win_inds_dict, drop_info, padding=0):
    """ops/sst/sst_ops.py:67-104.  Reference-layout (padded [R,T,C]) materialisation: API parity only - the
    attention kernels of this build consume the ragged CSR and never call this."""
    out = {}
    for dl in drop_info:
        dl_mask = voxel_drop_lvl == dl
        if not dl_mask.any():
            continue
        this_inds = flat2win_inds_dict[dl][0]
        T = drop_info[dl]["max_tokens"]
        R = int((this_inds // T).max()) + 1
        buf = torch.full((R * T, feat.shape[-1]), padding, dtype=feat.dtype, device=feat.device)
        buf[this_inds] = feat[dl_mask]
        out[dl] = buf.reshape(R, T, feat.shape[-1])
    return out


def window2flat(feat_3d_dict, inds_dict):
    """ops/sst/sst_ops.py:106-132."""
    n = sum(inds_dict[dl][0].shape[0] for dl in inds_dict)
    first = feat_3d_dict[next(iter(feat_3d_dict))]
    out = torch.zeros((n, first.shape[-1]), device=first.device, dtype=first.dtype)
    for dl in feat_3d_dict:
        inds, flat_pos = inds_dict[dl]
        out[flat_pos] = feat_3d_dict[dl].reshape(-1, first.shape[-1])[inds]
    return out


def window2flat_v2(feat_3d_dict, inds_dict):
    return window2flat(feat_3d_dict, {k: inds_dict[k] for k in inds_dict if not isinstance(k, str)})


def flat2window_v2(feat, inds_dict, padding=0):
    assert "voxel_drop_level" in inds_dict, "voxel_drop_level should be in inds_dict in v2 function"
    inds_v1 = {k: inds_dict[k] for k in inds_dict if not isinstance(k, str)}
    return flat2window(feat, inds_dict["voxel_drop_level"], inds_v1, inds_dict["batching_info"], padding=padding)


# ----------------------------------------------------------------------------------------------
# S3 small builders (host logic only)
# ----------------------------------------------------------------------------------------------
def get_activation(activation):
    """ops/sst/sst_ops.py:363-371."""
    if activation == "relu":
        return torch.nn.functional.relu
    if activation == "gelu":
        return torch.nn.functional.gelu
    if activation == "glu":
        return torch.nn.functional.glu
    raise RuntimeError(f"activation should be relu/gelu, not {activation}.")


def get_activation_layer(act, dim=None):
    """ops/sst/sst_ops.py:373-392."""
    act = act.lower()
    table = {"relu": lambda: nn.ReLU(inplace=True), "gelu": nn.GELU, "leakyrelu": lambda: nn.LeakyReLU(inplace=True),
             "prelu": lambda: nn.PReLU(num_parameters=dim), "swish": lambda: nn.SiLU(inplace=True),
             "silu": lambda: nn.SiLU(inplace=True), "glu": nn.GLU, "elu": lambda: nn.ELU(inplace=True)}
    if act not in table:
        raise NotImplementedError
    return table[act]()


def build_norm_layer(cfg, num_features):
    """mmcv.cnn.build_norm_layer for the norm types the path uses; returns (name, layer)."""
    from .norm import NaiveSyncBatchNorm1d, NaiveSyncBatchNorm2d
    cfg = dict(cfg)
    t = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if t == "LN":
        return "ln", nn.LayerNorm(num_features, **cfg)
    if t in ("BN1d", "BN"):
        return "bn", nn.BatchNorm1d(num_features, **cfg)
    if t == "naiveSyncBN1d":
        return "bn", NaiveSyncBatchNorm1d(num_features, **cfg)
    if t == "BN2d":
        return "bn", nn.BatchNorm2d(num_features, **cfg)
    if t == "naiveSyncBN2d":
        return "bn", NaiveSyncBatchNorm2d(num_features, **cfg)
    raise NotImplementedError(f"norm type {t}")


def build_mlp(in_channel, hidden_dims, norm_cfg, is_head=False, act="relu", bias=False, dropout=0):
    """ops/sst/sst_ops.py:334-361 (parameter container: `mlp.{i}.0.weight`, `mlp.{i}.1.{weight,bias}`)."""
    layers = []
    last = in_channel
    if isinstance(hidden_dims, int):
        hidden_dims = [hidden_dims]
    for i, ch in enumerate(hidden_dims):
        if i == len(hidden_dims) - 1 and is_head:
            layers.append(nn.Linear(last, ch, bias=True))
        else:
            sq = [nn.Linear(last, ch, bias=bias), build_norm_layer(norm_cfg, ch)[1], get_activation_layer(act, ch)]
            if dropout > 0:
                sq.append(nn.Dropout(dropout))
            layers.append(nn.Sequential(*sq))
        last = ch
    return nn.Sequential(*layers)


L.SIGNATURES["sstb200_linear"] = (C.c_int, [L.vp, L.vp, L.vp, L.vp, L.vp, C.c_int, C.c_int, C.c_int, C.c_int])


def linear(x, weight, bias=None, act=None):
    """nn.Linear forward (fp32) through libsstb200; act in (None, 'relu', 'gelu')."""
    _need_cuda(x, weight)
    x = x.float().contiguous()
    w = weight.detach().float().contiguous()
    b = None if bias is None else bias.detach().float().contiguous()
    out = torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
    c = L.ctx(x.device)
    L.check(c, L.lib().sstb200_linear(c, x.data_ptr(), w.data_ptr(), L.ptr(b), out.data_ptr(), x.shape[0], w.shape[0],
                                      w.shape[1], {None: 0, "relu": 1, "gelu": 2}[act]))
    return out
