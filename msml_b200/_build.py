"""In-tree build of libmsml_b200.so (sm_100a only) with plain nvcc — no torch extension machinery,
because the boundary is a C ABI (include/msml_b200.h) bound with ctypes.

    python -m msml_b200._build [--force]

Objects go to msml_b200/csrc/build/, the library to msml_b200/libmsml_b200.so (git-ignored, but
it travels to the GPU box with the gpurun snapshot).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libmsml_b200.so")
SOURCES = ["capi.cu", "fm_gate.cu", "fm_cat.cu", "fm_mask.cu", "dap.cu", "pfc_sample.cu", "head.cu", "bn_act.cu", "optim.cu", "seg_loss.cu", "comm.cu", "fm_peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
HEADERS = [os.path.join(CSRC, h) for h in ("common.cuh", "tc_gemm.cuh", "seg_loss_kernels.cuh", "pfc_sgd_kernels.cuh", "fm_cat_kernels.cuh", "bn_act_kernels.cuh", "fm_gate_kernels.cuh", "pfc_sample_kernels.cuh", "dap_kernels.cuh", "head_small_kernels.cuh", "fm_mask_kernels.cuh", "accum_kernels.cuh", "sgd_flat_kernels.cuh", "fm_peer_kernels.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "msml_b200.h")]


def _digest(paths):
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    stamp_path = os.path.join(OBJ, "stamp.txt")
    stamp = _digest([os.path.join(CSRC, s) for s in SOURCES] + HEADERS)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_path) and open(stamp_path).read() == stamp:
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_path, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
