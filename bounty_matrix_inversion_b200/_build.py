"""In-tree build of libbmi_tfhe.so (CUDA kernels + C ABI) for sm_100a.

The kernel launchers of every polynomial size are one translation unit each (launch_inst.cu with
-DBMI_INST_L=<log2 N>), compiled in parallel and linked with the context / C-ABI unit and the host client.
"""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.environ.get("BMI_TFHE_LIB") or os.path.join(CSRC, "libbmi_tfhe.so")   # override: kernel-variant experiments
OBJ = os.path.join(CSRC, "build")
SIZES = (10, 11, 12, 13, 14)
SOURCES = ["engine.cu", "client.cpp", "launch_inst.cu"]
HEADERS = ["field.cuh", "ntt.cuh", "kernels.cuh", "split.cuh", "split2.cuh", "leveled.cuh", "launchers.cuh", "host_common.h", "../../include/bmi_tfhe.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def stale():
    if os.environ.get("BMI_TFHE_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, lib=None, extra=()):
    """extra: additional nvcc flags (e.g. -DBMI_... overrides for kernel-variant experiments, with `lib` naming the output)"""
    lib = lib or LIB
    if not force and lib == LIB and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    tag = "" if lib == LIB else "_" + os.path.splitext(os.path.basename(lib))[0]
    flags = NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else [])
    units = [("engine.cu", f"engine{tag}.o", []), ("client.cpp", f"client{tag}.o", [])]
    units += [("launch_inst.cu", f"launch_L{L}{tag}.o", [f"-DBMI_INST_L={L}"]) for L in SIZES]

    def compile_unit(unit):
        src, obj, defs = unit
        out = os.path.join(OBJ, obj)
        r = subprocess.run([nvcc] + flags + defs + ["-c", "-o", out, src], cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} {defs}:\n{r.stdout}\n{r.stderr}")
        return out, r.stderr

    with ThreadPoolExecutor(max_workers=len(units)) as pool:
        results = list(pool.map(compile_unit, units))
    if verbose:
        for _, log in results:
            print(log)
    subprocess.check_call([nvcc, "-shared", "-o", lib] + [o for o, _ in results] + ["-lcudart"], cwd=CSRC)
    return lib


if __name__ == "__main__":
    print(build(force=True, verbose=True))
