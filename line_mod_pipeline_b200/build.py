"""Builds liblmb200.so (all CUDA kernels + the C ABI) in-tree for sm_100a with nvcc.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "liblmb200.so")
SOURCES = ["kernels_frame.cu", "kernels_spread.cu", "kernels_color.cu", "kernels_depth.cu", "kernels_postmatch.cu", "kernels_match.cu", "kernels_epilogue.cu", "detector.cu", "extract.cpp", "persistence.cpp", "comm.cpp", "capi.cpp", "render.cpp", "microbench.cu", "sort_check.cpp"]
HEADERS = ["kernels.cuh", "detector.h", "sort_emul.h", os.path.join("..", "..", "include", "lmb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    extra = os.environ.get("LMB200_NVCC_EXTRA", "").split()
    if not force and not needs_build() and not extra:
        return SO
    objs = []
    odir = os.path.join(HERE, "build")
    os.makedirs(odir, exist_ok=True)
    log = []
    for s in SOURCES:
        o = os.path.join(odir, s.rsplit(".", 1)[0] + ".o")
        cmd = [nvcc()] + NVCC_FLAGS + extra + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError("nvcc failed on " + s)
        objs.append(o)
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs + ["-lz", "-ldl", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    with open(os.path.join(odir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
