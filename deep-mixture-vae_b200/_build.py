"""Build recipe for libdmvae_b200.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun snapshots)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdmvae_b200.so")
SOURCES = ["abi.cu", "elbo.cu", "reparam.cu", "gemm_f32.cu", "gemm_tc.cu", "gemm_chain.cu", "moe.cu"]
HEADERS = ["common.cuh", "epilogue.cuh", "tc_device.cuh", "philox.cuh", "elbo_mma.cuh", os.path.join("..", "..", "include", "dmvae_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _digest() -> str:
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for d in SOURCES + HEADERS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    """Content hash (not mtime): a gpurun snapshot copy must not trigger a rebuild on the GPU box."""
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".sha256"):
        return True
    with open(LIB + ".sha256") as f:
        return f.read().strip() != _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu to an object and link libdmvae_b200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        log.append("==== %s ====\n%s" % (s, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % s)
        objs.append(obj)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    with open(LIB + ".sha256", "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
