"""In-tree build of the C-ABI CUDA library (sm_100a only).  No torch headers, no JIT cache: the .so lives next to
this file so it travels to the GPU box with the repo snapshot."""
import fcntl
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", n) for n in ("mvp_kernels.cu", "raydirs.cu", "epilogue.cu")]
# raydirs.cu mirrors the reference's utils extension, which is NOT built with -use_fast_math; epilogue.cu replaces
# eager PyTorch expressions and keeps their IEEE single-rounding arithmetic
NO_FAST_MATH = {"raydirs.cu", "epilogue.cu"}
HDR = [os.path.join(ROOT, "include", "mvpraymarch_b200.h"), os.path.join(HERE, "csrc", "epilogue_body.h"),
       os.path.join(HERE, "csrc", "raygen.h")]
LIB = os.path.join(HERE, "libmvpraymarch_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-diag-suppress=177",
    "-Xcompiler", "-fPIC",
    "-I" + os.path.join(ROOT, "include"),
]
FAST_MATH = ["-use_fast_math"]             # same math mode as the reference (extensions/mvpraymarch/setup.py:25)


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SRC + HDR + [os.path.abspath(__file__)])


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def have_nvcc():
    return os.path.exists(_nvcc()) or shutil.which(_nvcc()) is not None


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libmvpraymarch_b200.so.  Safe under torchrun: ranks serialise on a file lock (the first
    one builds, the others find the library up to date), objects go to a private temp directory and the finished .so is
    moved into place with os.replace, so no process can ever dlopen a half-written file."""
    if not force and not needs_build():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():          # another process built it while we waited
                return LIB
            nvcc = _nvcc()
            extra = os.environ.get("MVP_NVCC_EXTRA", "").split()      # experiment knob, e.g. -DMVP_BWD_MINB=4
            tmp = tempfile.mkdtemp(prefix=".build.", dir=HERE)
            try:
                objs = []
                for src in SRC:
                    obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
                    fm = [] if os.path.basename(src) in NO_FAST_MATH else FAST_MATH
                    subprocess.check_call([nvcc] + NVCC_FLAGS + fm + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
                    objs.append(obj)
                out = os.path.join(tmp, "lib.so")
                subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", out])
                os.replace(out, LIB)
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
