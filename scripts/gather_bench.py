#!/usr/bin/env python
"""Host->device transfer of one shuffled minibatch: dmvae_gather_rows reading pinned host memory (zero-copy) against a
plain cudaMemcpyAsync of a contiguous slice of the same size.   python scripts/gather_bench.py [--rows 4096] [--D 784]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmvae_b200 import _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--D", type=int, default=784)
    ap.add_argument("--n", type=int, default=65536)
    args = ap.parse_args()
    lib = _abi.load()
    ctx = C.c_void_p()
    _abi.check(lib.dmvae_ctx_create(0, C.byref(ctx)))
    host = torch.from_numpy(np.random.randint(0, 2, size=(args.n, args.D)).astype(np.uint8)).pin_memory()
    perm = torch.from_numpy(np.random.permutation(args.n).astype(np.int32)).cuda()
    dst = torch.empty(args.rows, args.D, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    sp = C.c_void_p(st.cuda_stream)
    rb = args.D

    def t(fn, iters=20):
        for _ in range(3):
            fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters

    nb = args.n // args.rows
    us_g = t(lambda i: _abi.check(lib.dmvae_gather_rows(ctx, host.data_ptr(), rb, perm.data_ptr() + 4 * (i % nb) * args.rows,
                                                        dst.data_ptr(), rb, args.rows, rb, sp)))
    us_c = t(lambda i: dst.copy_(host[(i % nb) * args.rows:(i % nb + 1) * args.rows], non_blocking=True))
    mb = args.rows * args.D / 1e6
    print("rows %d x %d B (%.2f MB): gather from pinned host %.1f us (%.1f GB/s), cudaMemcpyAsync contiguous %.1f us (%.1f GB/s)"
          % (args.rows, args.D, mb, us_g, mb / us_g * 1e3, us_c, mb / us_c * 1e3))
    dev = host.cuda()
    us_d = t(lambda i: _abi.check(lib.dmvae_gather_rows(ctx, dev.data_ptr(), rb, perm.data_ptr() + 4 * (i % nb) * args.rows,
                                                        dst.data_ptr(), rb, args.rows, rb, sp)))
    print("gather from a device-resident copy: %.1f us" % us_d)


if __name__ == "__main__":
    main()
