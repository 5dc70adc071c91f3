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


if __name__ == "__main__":
    print(build())
