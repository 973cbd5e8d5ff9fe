"""Build recipe for `oracle/_ref/quiptools_cuda*.so` -- the UNMODIFIED reference CUDA extension.

TEST INFRASTRUCTURE ONLY.  This compiles the reference's own kernel sources *where they lie* under
/root/reference/quip_cuda (quiptools_wrapper.cpp, e8p_gemv.cu, origin_order.cu) for sm_100a, with
outputs only into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  No reference
source is copied into this repository.  The reference's own build system (quip_cuda/setup.py) is
NOT run; this is the short recipe the task contract asks for.

The resulting module is used ONLY by
  * tests/ (-m gpu): dequantised weights bit-exact vs reference kernels K6-K8, mm outputs within
    tolerance vs K1-K3 (SURVEY.md section 8c "GPU-side cross-check"),
  * bench.py: an informational "kernel to beat" timing of the reference's recompiled kernels.
It is never imported by the product package `quip_for_all_b200`.

Run:  python oracle/build_ref.py        (needs /root/reference; ~2 min; no GPU required)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/quip_cuda"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "quiptools_cuda"


def ref_so_path():
    """Path of the built reference module, or None if it has not been built."""
    if not os.path.isdir(OUT_DIR):
        return None
    for f in sorted(os.listdir(OUT_DIR)):
        if f.startswith(NAME) and f.endswith(".so"):
            return os.path.join(OUT_DIR, f)
    return None


def load_ref_module():
    """Import oracle/_ref/quiptools_cuda.so (torch must be imported first). Returns None if absent."""
    path = ref_so_path()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        print(f"[build_ref] {REF_SRC} not present (GPU box?) -- using prebuilt files if any")
        return ref_so_path()
    if ref_so_path() is not None:
        return ref_so_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    cpp_extension.load(
        name=NAME,
        sources=[os.path.join(REF_SRC, f) for f in
                 ("quiptools_wrapper.cpp", "e8p_gemv.cu", "origin_order.cu")],
        extra_cflags=["-O2"],
        extra_cuda_cflags=["-O3", "-lineinfo"],
        build_directory=OUT_DIR,
        verbose=verbose,
        is_python_module=False,   # just produce the .so; we import it by path
    )
    # keep only the shared object (drop ninja scratch so the GPU snapshot stays small)
    for f in os.listdir(OUT_DIR):
        if f.endswith((".o", ".ninja", ".ninja_deps", ".ninja_log")) or f.startswith(".ninja"):
            try:
                os.remove(os.path.join(OUT_DIR, f))
            except OSError:
                pass
    return ref_so_path()


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("[build_ref]", p)
