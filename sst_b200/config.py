"""Tiny stand-in for mmcv.Config.fromfile (python configs with `_base_` inheritance and `_delete_`,
mmcv/utils/config.py) so that the reference's configs/sst*, configs/fsd*, configs/fsdv2 load with zero edits when
mmcv is absent.  Only what those configs use is implemented."""
import copy
import os
import runpy
import types


def _merge(base, new):
    out = copy.deepcopy(base)
    for k, v in new.items():
        if isinstance(v, dict) and k in out and isinstance(out[k], dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            v = copy.deepcopy(v)
            if isinstance(v, dict):
                v.pop("_delete_", None)
            out[k] = v
    return out


def _load(path):
    path = os.path.abspath(path)
    ns = runpy.run_path(path)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
    bases = cfg.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


class Config(dict):
    @classmethod
    def fromfile(cls, filename):
        return cls(_load(filename))

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return Config(v) if isinstance(v, dict) and not isinstance(v, Config) else v


# every type name the reference registers from a hot-path file - voxel_encoders/voxel_encoder.py,
# middle_encoders/sst_input_layer{,_v2}.py, middle_encoders/sparse_unet.py, backbones/{sst,sst_v1,sst_v2,sir}.py, necks/voxel2point_neck.py, ops/norm.py.
# A config that names one of these must find it in this package: "configs load unchanged" cannot pass by skipping.
REFERENCE_HOT_PATH_TYPES = frozenset((
    "DynamicVFE", "DynamicScatterVFE", "SIRLayer", "SSTInputLayer", "SSTInputLayerV2", "PseudoMiddleEncoderForSpconvFSD",
    "SST", "SSTv1", "SSTv2", "SIR", "Voxel2PointScatterNeck", "naiveSyncBN1d", "naiveSyncBN2d", "naiveSyncBN3d",
    # middle_encoders/sparse_unet.py
    "SparseUNet", "SimpleSparseUNet", "VirtualVoxelMixer"))


def find_hot_path_modules(cfg, registry, strict=True):
    """All sub-dicts of `cfg` whose `type` is registered here (voxel encoders, middle encoders, backbones, necks).
    strict: a `type` that the reference registers from a hot-path file but this package does not is an error."""
    found = []

    def walk(node, path):
        if isinstance(node, dict):
            t = node.get("type")
            if isinstance(t, str) and t in registry:
                found.append((path, node))
            elif strict and isinstance(t, str) and t in REFERENCE_HOT_PATH_TYPES and not t.startswith("naiveSyncBN"):
                raise KeyError(f"{path}: hot-path type {t!r} of the reference is not registered in sst_b200")
            for k, v in node.items():
                walk(v, f"{path}.{k}" if path else str(k))
        elif isinstance(node, (list, tuple)):
            for i, v in enumerate(node):
                walk(v, f"{path}[{i}]")
    walk(cfg, "")
    return found
