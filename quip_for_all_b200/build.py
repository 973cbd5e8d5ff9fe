"""Build recipe for libquipb200.so (hand-written sm_100a CUDA, plain nvcc, no torch headers).

The library is built IN-TREE (quip_for_all_b200/lib/libquipb200.so) so that it travels to the GPU box
with the repository snapshot.  `nvcc` cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libquipb200.so")
SOURCES = ["api.cu", "decompress.cu", "hadamard.cu", "quantlinear.cu", "glue.cu", "decode_step.cu", "umma_gemm.cu",
           "rotate_batched.cu", "lm_tail.cu", "nearest.cu", "handoff.cu"]
# (hook for translation units that need relocatable device code; none at present)
RDC_SOURCES = set()
RDC_FLAGS = {}
OBJDIR = os.path.join(HERE, "lib", "obj")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "quip_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into lib/libquipb200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "quip_b200.h"))
    hdr_t = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        sp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), hdr_t):
            return obj, 0, ""
        cmd = [nvcc] + NVCC_FLAGS + RDC_FLAGS.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, sp]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, res.returncode, res.stdout

    # one translation unit per thread: the sources are independent (no -rdc)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for obj, rc, out in results:
        if verbose or rc != 0:
            sys.stderr.write(out)
        if rc != 0:
            raise RuntimeError("nvcc failed building %s:\n%s" % (obj, out[-4000:]))
    extra = []
    rdc_objs = [os.path.join(OBJDIR, f.replace(".cu", ".o")) for f in SOURCES if f in RDC_SOURCES]
    if rdc_objs:
        dlink = os.path.join(OBJDIR, "dlink.o")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-dlink", "-o", dlink] + rdc_objs
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("device link failed:\n" + res.stdout[-4000:])
        extra = [dlink]
    cmd = [nvcc, "-shared", "-o", LIB] + [r[0] for r in results] + extra
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed for libquipb200.so:\n" + res.stdout[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
