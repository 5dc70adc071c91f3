"""Build the UNMODIFIED reference CUDA extension (mvpraymarchlib) for sm_100 into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  This compiles the reference's own sources *where they lie* under
/root/reference/extensions/mvpraymarch (mvpraymarch.cpp, mvpraymarch_kernel.cu, bvh.cu; flags from its
setup.py:18-31 with only -arch=sm_70 -> sm_100 changed) and writes nothing but build outputs into
oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).  No reference source is copied into this repo.

The resulting pybind module is used by tests/ and bench.py (never by the product path) as:
  * the direct GPU parity comparator for our kernels, and
  * the generator of the committed golden vectors under tests/golden/ (tests/golden/make_golden.py).

/root/reference does not exist on the GPU box; there only the prebuilt oracle/_ref/mvpraymarchlib.so is used.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AVA256_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build_utils(verbose: bool = False) -> str:
    """Same for the reference's ray generator (extensions/utils -> utilslib); its setup.py has no -use_fast_math."""
    so = os.path.join(OUT, "utilslib", "utilslib.so")
    src = os.path.join(REF, "extensions", "utils")
    if not os.path.isdir(src):
        if os.path.exists(so):
            return so
        raise FileNotFoundError("reference sources not found at %s and no prebuilt %s" % (src, so))
    srcs = [os.path.join(src, f) for f in ("utils.cpp", "utils_kernel.cu")]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(os.path.dirname(so), exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load

    load(name="utilslib", sources=srcs, extra_include_paths=[os.path.join(REF, "extensions", "include")],
         extra_cuda_cflags=["-gencode", "arch=compute_100,code=sm_100", "-std=c++17", "-lineinfo"],
         extra_cflags=["-DNDEBUG"],   # as setuptools builds it; utils.cpp:67 only compiles with asserts disabled
         build_directory=os.path.dirname(so), verbose=verbose, is_python_module=False)
    return so


def build(verbose: bool = False) -> str:
    """Returns the path of the built module, building it if /root/reference is present."""
    so = os.path.join(OUT, "mvpraymarchlib.so")
    src = os.path.join(REF, "extensions", "mvpraymarch")
    if not os.path.isdir(src):
        if os.path.exists(so):
            return so
        raise FileNotFoundError("reference sources not found at %s and no prebuilt %s" % (src, so))
    srcs = [os.path.join(src, f) for f in ("mvpraymarch.cpp", "mvpraymarch_kernel.cu", "bvh.cu")]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load

    load(
        name="mvpraymarchlib",
        sources=srcs,
        extra_include_paths=[os.path.join(REF, "extensions", "include")],
        extra_cuda_cflags=["-use_fast_math", "-gencode", "arch=compute_100,code=sm_100", "-std=c++17", "-lineinfo"],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,
    )
    return so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    print(build_utils(verbose="-v" in sys.argv))
