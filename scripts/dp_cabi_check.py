#!/usr/bin/env python
"""Data parallel through the C ABI ALONE - no torch anywhere in this script.  `world` processes (one GPU each) exchange
dmvae_dp_alloc's 64-byte IPC handles through multiprocessing queues, map each other's buffer with dmvae_dp_open, and run
dmvae_dp_barrier + dmvae_dp_reduce_adam on the mapped pointers; the summed gradient's Adam update must land in every
replica (fp32 master + bf16 operand copy).      python scripts/dp_cabi_check.py [world=2]"""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _cudart():
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    raise OSError("libcudart not found")


def worker(rank, world, qs, res):
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaSetDevice(rank) == 0
    # the ctypes binding of include/dmvae_b200.h, loaded without importing torch
    import importlib.util
    pkg = os.path.join(ROOT, "deep-mixture-vae_b200")
    spec = importlib.util.spec_from_file_location("dmvae_b200", os.path.join(pkg, "__init__.py"), submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dmvae_b200"] = mod
    spec.loader.exec_module(mod)
    from dmvae_b200 import _abi
    assert "torch" not in sys.modules
    lib = _abi.load()
    ctx = C.c_void_p()
    _abi.check(lib.dmvae_ctx_create(rank, C.byref(ctx)))
    n = 1 << 16
    pad_off = 10 * n
    nbytes = pad_off + 32 * 8 * 4                       # params | grads | bf16 copy | barrier pads
    local = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    _abi.check(lib.dmvae_dp_alloc(ctx, nbytes, C.byref(local), handle))
    for r in range(world):
        if r != rank:
            qs[r].put((rank, bytes(handle)))
    bases = [0] * world
    bases[rank] = local.value
    for _ in range(world - 1):
        src, hb = qs[rank].get()
        ph = (C.c_ubyte * 64).from_buffer_copy(hb)
        pp = C.c_void_p()
        _abi.check(lib.dmvae_dp_open(ctx, ph, C.byref(pp)))
        bases[src] = pp.value
    VP = C.c_void_p * world
    P, G = VP(*bases), VP(*[b + 4 * n for b in bases])
    Bf, pads = VP(*[b + 8 * n for b in bases]), VP(*[b + pad_off for b in bases])
    p0 = np.random.RandomState(0).randn(n).astype(np.float32)
    grads = [np.random.RandomState(10 + r).randn(n).astype(np.float32) for r in range(world)]
    H2D, D2H = 1, 2
    assert rt.cudaMemcpy(bases[rank], p0.ctypes.data, 4 * n, H2D) == 0
    assert rt.cudaMemcpy(bases[rank] + 4 * n, grads[rank].ctypes.data, 4 * n, H2D) == 0
    # Adam slots of the owned shard, barrier epochs: ordinary device memory
    per = (n // world + 3) // 4 * 4
    b, e = rank * per, min(n, (rank + 1) * per)
    m, v, ep = C.c_void_p(), C.c_void_p(), C.c_void_p()
    for ptr, sz in ((m, 4 * per), (v, 4 * per), (ep, 4 * 32)):
        assert rt.cudaMalloc(C.byref(ptr), C.c_size_t(sz)) == 0
        assert rt.cudaMemset(ptr, 0, C.c_size_t(sz)) == 0
    lr_t = 0.002 * np.sqrt(1 - 0.999) / (1 - 0.9)
    _abi.check(lib.dmvae_dp_barrier(ctx, rank, world, pads, ep, 0, None))          # every replica is filled
    _abi.check(lib.dmvae_dp_reduce_adam(ctx, rank, world, G, P, Bf, None, None, 0, m, v, n, b, e, C.c_float(lr_t), None,
                                        C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), 0, None))
    _abi.check(lib.dmvae_dp_barrier(ctx, rank, world, pads, ep, 1, None))          # every shard is written everywhere
    assert rt.cudaDeviceSynchronize() == 0
    got = np.empty(n, np.float32)
    assert rt.cudaMemcpy(got.ctypes.data, bases[rank], 4 * n, D2H) == 0
    gs = np.zeros(n, np.float32)
    for r in range(world):                              # rank order, like the kernel
        gs = gs + grads[r]
    mm = (1 - 0.9) * gs
    vv = (1 - 0.999) * gs * gs
    ref = p0 - np.float32(lr_t) * mm / (np.sqrt(vv) + np.float32(1e-8))
    err = float(np.abs(got - ref).max())
    res.put((rank, err))
    _abi.check(lib.dmvae_dp_barrier(ctx, rank, world, pads, ep, 2, None))          # nobody unmaps while a peer still reads
    assert rt.cudaDeviceSynchronize() == 0
    for r in range(world):
        if r != rank:
            _abi.check(lib.dmvae_dp_close(ctx, C.c_void_p(bases[r])))
    lib.dmvae_ctx_destroy(ctx)


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    mp.set_start_method("spawn")
    qs = [mp.Queue() for _ in range(world)]
    res = mp.Queue()
    ps = [mp.Process(target=worker, args=(r, world, qs, res)) for r in range(world)]
    for p in ps:
        p.start()
    errs = dict(res.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    ok = all(e < 1e-6 for e in errs.values()) and all(p.exitcode == 0 for p in ps)
    print("C-ABI data parallel (IPC handles, no torch), world %d: max |param - reference| per rank %s -> %s"
          % (world, {r: "%.1e" % e for r, e in sorted(errs.items())}, "PASS" if ok else "FAIL"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
