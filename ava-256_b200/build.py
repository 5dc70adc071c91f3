"""In-tree build of the C-ABI CUDA library (sm_100a only).  No torch headers, no JIT cache: the .so lives next to
this file so it travels to the GPU box with the repo snapshot."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", n) for n in ("mvp_kernels.cu", "raydirs.cu", "epilogue.cu")]
# raydirs.cu mirrors the reference's utils extension, which is NOT built with -use_fast_math; epilogue.cu replaces
# eager PyTorch expressions and keeps their IEEE single-rounding arithmetic
NO_FAST_MATH = {"raydirs.cu", "epilogue.cu"}
HDR = [os.path.join(ROOT, "include", "mvpraymarch_b200.h"), os.path.join(HERE, "csrc", "epilogue_body.h")]
LIB = os.path.join(HERE, "libmvpraymarch_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-I" + os.path.join(ROOT, "include"),
]
FAST_MATH = ["-use_fast_math"]             # same math mode as the reference (extensions/mvpraymarch/setup.py:25)


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SRC + HDR + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MVP_NVCC_EXTRA", "").split()      # experiment knob, e.g. -DMVP_BWD_MINB=4
    objs = []
    for src in SRC:
        obj = os.path.join(HERE, "csrc", os.path.basename(src)[:-3] + ".o")
        fm = [] if os.path.basename(src) in NO_FAST_MATH else FAST_MATH
        subprocess.check_call([nvcc] + NVCC_FLAGS + fm + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
