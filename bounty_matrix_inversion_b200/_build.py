"""In-tree build of libbmi_tfhe.so (CUDA kernels + C ABI) for sm_100a."""
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.environ.get("BMI_TFHE_LIB") or os.path.join(CSRC, "libbmi_tfhe.so")   # override: kernel-variant experiments
SOURCES = ["engine.cu", "client.cpp"]
HEADERS = ["field.cuh", "ntt.cuh", "kernels.cuh", "split.cuh", "host_common.h", "../../include/bmi_tfhe.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def stale():
    if os.environ.get("BMI_TFHE_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-lcudart"]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
