"""oracle/build_ref.py — TEST INFRASTRUCTURE ONLY.

Compiles the UNMODIFIED reference dcn CUDA extension from the sources where they lie
under /root/reference (never copied into this repo) into oracle/_ref/ (git-ignored,
travels to the GPU box with gpurun).  It is the GPU-side oracle for DCN forward/backward
parity and the "reference CUDA dcn path" timed beside the product in bench.py.

Sources: /root/reference/basicsr/models/ops/dcn/src/{deform_conv_ext.cpp,
deform_conv_cuda.cpp, deform_conv_cuda_kernel.cu}; flags as the reference's setup.py:98-135
(-DWITH_CUDA and the three -D__CUDA_NO_HALF* defines), arch forced to sm_100a.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "deform_conv_ext_ref"
SRC = "/root/reference/basicsr/models/ops/dcn/src"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    """Build if the reference tree is present and the .so is missing. Returns the .so path or None."""
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(name=NAME,
         sources=[os.path.join(SRC, f) for f in
                  ("deform_conv_ext.cpp", "deform_conv_cuda.cpp", "deform_conv_cuda_kernel.cu")],
         extra_cflags=["-DWITH_CUDA", "-O2"],
         extra_cuda_cflags=["-DWITH_CUDA", "-D__CUDA_NO_HALF_OPERATORS__",
                            "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
         with_cuda=True, build_directory=OUT, is_python_module=False, verbose=verbose)
    return so_path() if os.path.exists(so_path()) else None


def load_ref():
    """Import the prebuilt extension module (needs torch; runs only on a CUDA box)."""
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    path = so_path()
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("built:" if p else "not built (no /root/reference):", p)
