"""TEST INFRASTRUCTURE - recipe that compiles the reference's OWN voxel_layer sources, unmodified and in
place (/root/reference/mmdet3d/ops/voxel/src/*.{cpp,cu}), into oracle/_ref/voxel_layer_ref*.so.

The result is used only by tests (to pin oracle/sst_oracle.py and the CUDA path against the real reference
kernels on a GPU box).  Nothing is copied into the repo; oracle/_ref/ is git-ignored but travels with gpurun.
Runs only where /root/reference exists (the build container).  `python -m oracle.build_ref`
"""
import glob
import os
import sys

REF_SRC = os.path.join(os.environ.get("SST_REFERENCE_ROOT", "/root/reference"), "mmdet3d/ops/voxel/src")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def built():
    return sorted(glob.glob(os.path.join(OUT, "voxel_layer_ref*.so")))


def build(verbose=False):
    if built():
        return built()[0]
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "6")
    from torch.utils.cpp_extension import load
    srcs = [os.path.join(REF_SRC, f) for f in ("voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp",
                                               "scatter_points_cuda.cu", "voxelization_cuda.cu")]
    load(name="voxel_layer_ref", sources=srcs, extra_cflags=["-DWITH_CUDA", "-O2"],
         extra_cuda_cflags=["-DWITH_CUDA", "-O2", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return built()[0] if built() else None


def load_module():
    """Import the prebuilt extension (no compilation; works on the GPU box where the sources are absent)."""
    so = built()
    if not so:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)
    spec = importlib.util.spec_from_file_location("voxel_layer_ref", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
