"""TEST INFRASTRUCTURE ONLY - never imported by the product package.

Loads the reference's own hot-path Python files *unmodified* from
/root/reference (read-only, only present in the build container, never on the
GPU box) so that

  * oracle/sst_oracle.py (the portable CPU restatement) can be pinned against
    the real reference code, and
  * oracle/make_golden.py can emit the fixtures under tests/golden/.

Nothing is copied: the reference modules are imported in place.  The reference
depends on mmcv / mmdet / torch_scatter / TorchEx(ingroup_indices) / ipdb which
are absent from this image, so `sys.modules` is pre-seeded with minimal stubs
(SURVEY.md Appendix A).  The two third-party *arithmetic* dependencies are
restated here because their sources are not under /root/reference:

  torch_scatter 2.0.9 (docs/overall_instructions.md:36) -> `scatter`,
      `scatter_max`: call sites mmdet3d/ops/sst/sst_ops.py:173-175.
  TorchEx `ingroup_indices.forward` (no version pinned; docs link only) ->
      rank of every element inside its group; specification recovered from the
      in-tree equivalents sst_input_layer.py:200-208 (`_slow`, stable order)
      and sst_ops.py:194-242.  Parity for both is "unpinned" by the
      reference's own tests.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("SST_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mmdet3d"))


# --------------------------------------------------------------------------
# third-party restatements (CPU, torch)
# --------------------------------------------------------------------------
def _ts_scatter(src, index, dim=0, reduce="sum", dim_size=None):
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("sum", "add"):
        return out.scatter_reduce(0, idx, src, "sum", include_self=False)
    if reduce == "mean":
        return out.scatter_reduce(0, idx, src, "mean", include_self=False)
    if reduce == "max":
        return out.scatter_reduce(0, idx, src, "amax", include_self=False)
    raise NotImplementedError(reduce)


def _ts_scatter_max(src, index, dim=0, dim_size=None):
    out = _ts_scatter(src, index, dim, "max", dim_size)
    # argmax: first (lowest) point index attaining the max, like torch_scatter
    n = out.shape[0]
    P = src.shape[0]
    hit = src == out[index]
    cand = torch.where(hit, torch.arange(P).view(-1, *([1] * (src.dim() - 1))).expand_as(src),
                       torch.full_like(src, P, dtype=torch.long))
    arg = torch.full(out.shape, P, dtype=torch.long)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    arg = arg.scatter_reduce(0, idx, cand, "amin", include_self=True)
    return out, arg


def _ingroup_forward(group_inds, out_inds):
    """Stable rank-in-group == get_inner_win_inds_slow (sst_input_layer.py:200-208)."""
    order = torch.sort(group_inds, stable=True).indices
    g = group_inds[order]
    n = g.numel()
    if n == 0:
        return
    start = torch.ones(n, dtype=torch.bool)
    start[1:] = g[1:] != g[:-1]
    seg_start = torch.cummax(torch.where(start, torch.arange(n), torch.zeros(n, dtype=torch.long)), 0).values
    out_inds[order] = torch.arange(n) - seg_start


class _Registry:
    def __init__(self):
        self.d = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.d[name or cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg, *a, **k):
        cfg = dict(cfg)
        return self.d[cfg.pop("type")](**cfg)


def _build_norm_layer(cfg, num_features, postfix=""):
    cfg = dict(cfg)
    t = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if t == "LN":
        return "ln", nn.LayerNorm(num_features, **cfg)
    if t in ("BN1d", "naiveSyncBN1d", "BN"):
        return f"bn{postfix}", nn.BatchNorm1d(num_features, **cfg)
    if t in ("BN2d", "naiveSyncBN2d"):
        return "bn", nn.BatchNorm2d(num_features, **cfg)
    raise NotImplementedError(t)


def _build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type="Conv2d"))
    t = cfg.pop("type")
    assert t in ("Conv2d", "Conv")
    return nn.Conv2d(*args, **kwargs, **cfg)


class _BaseModule(nn.Module):
    """mmcv.runner.BaseModule: nn.Module whose constructor takes init_cfg"""

    def __init__(self, init_cfg=None):
        super().__init__()


def _passthrough_deco(*dargs, **dkwargs):
    def deco(fn):
        return fn
    return deco


_LOADED = None


def load():
    """Import the reference hot-path modules; returns a namespace of classes."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    assert available(), f"reference tree not found at {REF_ROOT}"
    reg = _Registry()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("ipdb", set_trace=lambda *a, **k: None)
    mod("mmcv")
    mod("mmcv.runner", auto_fp16=_passthrough_deco, force_fp32=_passthrough_deco, BaseModule=_BaseModule)
    mod("mmcv.cnn", build_norm_layer=_build_norm_layer, build_conv_layer=_build_conv_layer,
        NORM_LAYERS=_Registry(), ConvModule=None)
    mod("mmdet")
    mod("mmdet.models", BACKBONES=reg, NECKS=reg)
    mod("torch_scatter", scatter=_ts_scatter, scatter_max=_ts_scatter_max)
    mod("ingroup_indices", forward=_ingroup_forward)

    def pkg(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_ROOT, rel)]
        sys.modules[name] = m
        return m

    pkg("mmdet3d", "mmdet3d")
    ops = pkg("mmdet3d.ops", "mmdet3d/ops")
    pkg("mmdet3d.ops.sst", "mmdet3d/ops/sst")
    sp = pkg("mmdet3d.ops.spconv", "mmdet3d/ops/spconv")
    sp.IS_SPCONV2_AVAILABLE = False
    ops.spconv = sp
    pkg("mmdet3d.models", "mmdet3d/models")
    bld = mod("mmdet3d.models.builder", MIDDLE_ENCODERS=reg, VOXEL_ENCODERS=reg, BACKBONES=reg,
              build_voxel_encoder=reg.build, build_fusion_layer=None)
    sys.modules["mmdet3d.models"].builder = bld
    ops.voxel = types.ModuleType("mmdet3d.ops.voxel")  # `from mmdet3d.ops import voxel` in necks/voxel2point_neck.py (unused there)
    for sub in ("middle_encoders", "backbones", "sst", "voxel_encoders", "necks"):
        pkg(f"mmdet3d.models.{sub}", f"mmdet3d/models/{sub}")

    sst_ops = importlib.import_module("mmdet3d.ops.sst.sst_ops")
    for k in dir(sst_ops):
        if not k.startswith("_"):
            setattr(ops, k, getattr(sst_ops, k))
    ops.make_sparse_convmodule = None

    # DynamicScatter on CPU: the reference has no CPU dynamic_point_to_voxel
    # (voxelization.h:96-108) -> use the oracle's restatement of
    # scatter_points_cuda.cu:183-234 wrapped in the reference's own per-sample
    # loop semantics (scatter_points.py:78-99).
    from oracle import sst_oracle as O

    class DynamicScatter(nn.Module):
        def __init__(self, voxel_size, point_cloud_range, average_points):
            super().__init__()
            self.average_points = average_points

        def forward(self, points, coors):
            return O.dynamic_scatter_module(points, coors, self.average_points)

    ops.DynamicScatter = DynamicScatter

    ve_utils = importlib.import_module("mmdet3d.models.voxel_encoders.utils")
    ve = importlib.import_module("mmdet3d.models.voxel_encoders.voxel_encoder")
    il2 = importlib.import_module("mmdet3d.models.middle_encoders.sst_input_layer_v2")
    blk = importlib.import_module("mmdet3d.models.sst.sst_basic_block_v2")
    sstv2 = importlib.import_module("mmdet3d.models.backbones.sst_v2")
    sir = importlib.import_module("mmdet3d.models.backbones.sir")
    cos = importlib.import_module("mmdet3d.models.sst.cosine_msa")
    v2p = importlib.import_module("mmdet3d.models.necks.voxel2point_neck")
    il1 = importlib.import_module("mmdet3d.models.middle_encoders.sst_input_layer")
    blk1 = importlib.import_module("mmdet3d.models.sst.sst_basic_block")
    sstv1 = importlib.import_module("mmdet3d.models.backbones.sst_v1")

    ns = types.SimpleNamespace(
        sst_ops=sst_ops, voxel_encoder=ve, ve_utils=ve_utils, input_layer_v2=il2, block_v2=blk,
        sst_v2=sstv2, sir=sir, cosine_msa=cos, registry=reg,
        DynamicVFE=ve.DynamicVFE, DynamicScatterVFE=ve.DynamicScatterVFE, SIRLayer=ve.SIRLayer,
        SSTInputLayerV2=il2.SSTInputLayerV2, SSTv2=sstv2.SSTv2, SIR=sir.SIR,
        EncoderLayer=blk.EncoderLayer, WindowAttention=blk.WindowAttention,
        Voxel2PointScatterNeck=v2p.Voxel2PointScatterNeck,
        input_layer_v1=il1, block_v1=blk1, sst_v1=sstv1, SSTInputLayer=il1.SSTInputLayer, SSTv1=sstv1.SSTv1,
    )
    _LOADED = ns
    return ns


class _RefBasicBlock(nn.Module):
    """mmdet.models.backbones.resnet.BasicBlock (mmdet 2.x, absent from this image) restated as far as
    mmdet3d/ops/sparse_block.py:81-143 uses it: attribute / state-dict names conv1, bn1, conv2, bn2, relu, downsample."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style="pytorch", with_cp=False, conv_cfg=None,
                 norm_cfg=dict(type="BN"), dcn=None, plugins=None, init_cfg=None):
        super().__init__()
        build_norm, build_conv = sys.modules["mmcv.cnn"].build_norm_layer, sys.modules["mmcv.cnn"].build_conv_layer
        self.norm1_name, norm1 = build_norm(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm(norm_cfg, planes, postfix=2)
        self.conv1 = build_conv(conv_cfg, inplanes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = build_conv(conv_cfg, planes, planes, 3, padding=1, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride, self.dilation, self.with_cp = stride, dilation, with_cp

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    @property
    def norm2(self):
        return getattr(self, self.norm2_name)


_SPCONV = None


def load_spconv():
    """The reference's vendored spconv v1 (mmdet3d/ops/spconv/*.py, imported unmodified, over oracle/_ref/sparse_conv_ext_ref*.so =
    its own C++/CUDA sources compiled by oracle/build_ref.py), its sparse_block.py and its sparse U-Nets."""
    global _SPCONV
    if _SPCONV is not None:
        return _SPCONV
    R = load()
    from oracle import build_ref
    ext = build_ref.load_module("sparse_conv_ext_ref")
    assert ext is not None, "oracle/_ref/sparse_conv_ext_ref*.so not built (python -m oracle.build_ref)"
    cnn = sys.modules["mmcv.cnn"]
    conv_reg = _Registry()
    cnn.CONV_LAYERS = conv_reg
    dense_build = cnn.build_conv_layer

    def build_conv_layer(cfg, *args, **kwargs):
        c = dict(cfg or dict(type="Conv2d"))
        if c.get("type") in conv_reg.d:
            return conv_reg.d[c.pop("type")](*args, **kwargs, **c)
        return dense_build(cfg, *args, **kwargs)

    cnn.build_conv_layer = build_conv_layer
    sp = sys.modules["mmdet3d.ops.spconv"]
    sp.sparse_conv_ext = ext
    sys.modules["mmdet3d.ops.spconv.sparse_conv_ext"] = ext
    structure = importlib.import_module("mmdet3d.ops.spconv.structure")
    modules = importlib.import_module("mmdet3d.ops.spconv.modules")
    conv = importlib.import_module("mmdet3d.ops.spconv.conv")
    if not hasattr(structure.SparseConvTensor, "replace_feature"):
        # sparse_unet.py:181-186 calls x.replace_feature (spconv 2.x API); on the v1 tensor it is a plain re-wrap
        def replace_feature(self, feature):
            t = structure.SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid)
            t.indice_dict = self.indice_dict
            return t
        structure.SparseConvTensor.replace_feature = replace_feature
    m = types.ModuleType("mmcv.ops")
    m.SparseConvTensor, m.SparseSequential, m.SparseModule = structure.SparseConvTensor, modules.SparseSequential, modules.SparseModule
    sys.modules["mmcv.ops"] = m
    for name in ("mmdet.models.backbones", "mmdet.models.backbones.resnet"):
        mm = types.ModuleType(name)
        sys.modules[name] = mm
    sys.modules["mmdet.models.backbones.resnet"].BasicBlock = _RefBasicBlock
    sys.modules["mmdet.models.backbones.resnet"].Bottleneck = _RefBasicBlock
    sb = importlib.import_module("mmdet3d.ops.sparse_block")
    ops = sys.modules["mmdet3d.ops"]
    ops.SparseBasicBlock, ops.make_sparse_convmodule = sb.SparseBasicBlock, sb.make_sparse_convmodule
    ops.SparseBottleneck = sb.SparseBottleneck
    unet = importlib.import_module("mmdet3d.models.middle_encoders.sparse_unet")
    _SPCONV = types.SimpleNamespace(ext=ext, structure=structure, modules=modules, conv=conv, sparse_block=sb, sparse_unet=unet,
                                    SparseConvTensor=structure.SparseConvTensor, SubMConv3d=conv.SubMConv3d,
                                    SparseConv3d=conv.SparseConv3d, SparseInverseConv3d=conv.SparseInverseConv3d,
                                    SimpleSparseUNet=unet.SimpleSparseUNet, VirtualVoxelMixer=unet.VirtualVoxelMixer,
                                    SparseBasicBlock=sb.SparseBasicBlock, make_sparse_convmodule=sb.make_sparse_convmodule)
    return _SPCONV


def reference_methods(relpath, class_name, names, extra_globals=None):
    """Compile selected METHODS of a reference class straight from its source file, unmodified, without importing the module
    (detector files import half of mmdet / mmseg / spconv at module level).  Returns {name: function}; decorators are dropped
    (`force_fp32` / `torch.no_grad` do not change CPU fp32 results).  The functions see the shim's `scatter_v2` / `build_mlp`."""
    import ast
    R = load()
    src = open(os.path.join(REF_ROOT, relpath)).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    g = {"torch": torch, "scatter_v2": R.sst_ops.scatter_v2, "build_mlp": R.sst_ops.build_mlp, "F": torch.nn.functional}
    g.update(extra_globals or {})
    out = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, os.path.join(REF_ROOT, relpath), "exec"), g)
            out[node.name] = g[node.name]
    missing = set(names) - set(out)
    assert not missing, f"{class_name} has no method(s) {missing}"
    return out


def reference_functions(relpath, names, extra_globals=None):
    """Compile selected MODULE-LEVEL functions / classes of a reference file straight from its source, unmodified, without importing the
    module (detector files import half of mmdet / mmseg / spconv at module level).  Returns a dict name -> object; the objects see
    torch, scipy's connected_components, the shim's scatter_v2 and a restated mmdet.core.multi_apply."""
    import ast
    from functools import partial
    R = load()
    src = open(os.path.join(REF_ROOT, relpath)).read()
    tree = ast.parse(src)

    def multi_apply(func, *args, **kwargs):   # mmdet/core/utils/misc.py
        pfunc = partial(func, **kwargs) if kwargs else func
        return tuple(map(list, zip(*map(pfunc, *args))))

    from scipy.sparse.csgraph import connected_components
    g = {"torch": torch, "nn": nn, "scatter_v2": R.sst_ops.scatter_v2, "connected_components": connected_components,
         "multi_apply": multi_apply, "F": torch.nn.functional, "cc_gpu": None}
    g.update(extra_globals or {})
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in body}
    assert not missing, f"{relpath} has no top-level {missing}"
    for node in body:
        if isinstance(node, ast.ClassDef):
            node.decorator_list = []
    mod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, os.path.join(REF_ROOT, relpath), "exec"), g)
    return {n: g[n] for n in names}
