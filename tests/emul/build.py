"""TEST INFRASTRUCTURE: builds tests/emul/libepilogue_emul.so (host build of csrc/epilogue_body.h) with g++."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "epilogue_emul.cpp")
BODY = os.path.join(ROOT, "ava-256_b200", "csrc", "epilogue_body.h")
LIB = os.path.join(HERE, "libepilogue_emul.so")


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(f) for f in (SRC, BODY, __file__)):
        return LIB
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I" + cuda_inc,
                           "-I" + os.path.dirname(BODY), SRC, "-o", LIB])
    return LIB


KSRC = [os.path.join(HERE, "mvp_emul.cpp"), os.path.join(HERE, "cuda_emul.cpp")]
KDEP = KSRC + [os.path.join(HERE, "cuda_emul.h"), os.path.join(ROOT, "ava-256_b200", "csrc", "mvp_kernels.cu"),
               os.path.join(ROOT, "ava-256_b200", "csrc", "raygen.h"),
               os.path.join(ROOT, "include", "mvpraymarch_b200.h")]
KLIB = os.path.join(HERE, "libmvp_emul.so")


def build_kernels(defines=(), opt="-O1"):
    """libmvp_emul[_<defines>].so: the product kernels' source compiled for the host on the CPU emulation of cuda_emul.h.
    `defines` selects one of the kernels' build-time variants, e.g. ("MVP_LIST_REUSE=1",)."""
    tag = "".join("_" + d.replace("=", "").replace("MVP_", "").lower() for d in defines) + ("" if opt == "-O1" else "_" + opt.strip("-").lower())
    lib = KLIB[:-3] + tag + ".so"
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(f) for f in KDEP + [__file__]):
        return lib
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++20", opt, "-g", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-Wno-attributes",
                           "-Wno-unknown-pragmas", "-I" + cuda_inc, "-I" + HERE, "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "ava-256_b200", "csrc")] +
                          ["-D" + d for d in defines] + KSRC + ["-o", lib])
    return lib


ALIB = os.path.join(HERE, "libaux_emul.so")


def build_aux():
    """libaux_emul.so: raydirs.cu + epilogue.cu compiled for the host on the CPU emulation."""
    src = [os.path.join(HERE, "aux_emul.cpp"), os.path.join(HERE, "cuda_emul.cpp")]
    dep = src + [os.path.join(HERE, "cuda_emul.h"), BODY, os.path.join(ROOT, "ava-256_b200", "csrc", "raydirs.cu"),
                 os.path.join(ROOT, "ava-256_b200", "csrc", "raygen.h"), os.path.join(ROOT, "ava-256_b200", "csrc", "epilogue.cu"), os.path.join(ROOT, "include", "mvpraymarch_b200.h")]
    if os.path.exists(ALIB) and all(os.path.getmtime(ALIB) >= os.path.getmtime(f) for f in dep + [__file__]):
        return ALIB
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-Wno-attributes",
                           "-Wno-unknown-pragmas", "-I" + cuda_inc, "-I" + HERE, "-I" + os.path.dirname(BODY),
                           "-I" + os.path.join(ROOT, "include")] + src + ["-o", ALIB])
    return ALIB


if __name__ == "__main__":
    print(build())
    print(build_kernels())
    print(build_aux())
