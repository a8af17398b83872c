#!/usr/bin/env python
"""Kernel timeline of Engine.run_epoch (the end-to-end path: pinned host rows -> device gather -> captured step): per
step the start-to-start interval, the gather kernel's duration and the idle gaps on the compute stream.
    python scripts/epoch_timeline.py [--config 2] [--steps 24]"""
import argparse
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import configs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--full", action="store_true", help="print every kernel of the last two steps")
    ap.add_argument("--packed", action="store_true", help="binarised rows at one bit per element (dmvae_gather_rows_bits)")
    a = ap.parse_args()
    cfg = configs.CONFIGS[a.config]
    B = cfg["batch"]
    eng = configs.make_engine(cfg, B)
    opt = eng.optimizer("train", 0.002)
    N = a.steps * B
    X = configs.synth_inputs(cfg, N)
    kw = dict(x_scale=configs.x_scale(cfg))
    if a.packed:
        from dmvae_b200.includes.utils import _pack_bits
        X, kw = _pack_bits(X), dict(x_scale=1.0, packed_D=cfg["D"])
    host = torch.from_numpy(X).pin_memory()
    rs = np.random.RandomState(0)
    for _ in range(2):
        eng.run_epoch(host, B, opt, perm=rs.permutation(N), **kw)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.run_epoch(host, B, opt, perm=rs.permutation(N), **kw)
        torch.cuda.synchronize()
    out = os.path.join(ROOT, "gpurun_out", "epoch_trace.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    prof.export_chrome_trace(out)
    ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    ticks = [e["ts"] for e in ev if "step_tick" in e["name"]]
    gath = [e for e in ev if "gather_" in e["name"]]
    print("%s: %d steps, %d gathers" % (cfg["name"], len(ticks), len(gath)))
    iv = np.diff(ticks)
    print("step start-to-start us: median %.1f  min %.1f  max %.1f  first five %s" %
          (np.median(iv), iv.min(), iv.max(), np.round(iv[:5], 1)))
    print("gather kernel us: median %.1f  min %.1f  max %.1f" %
          (np.median([g["dur"] for g in gath]), min(g["dur"] for g in gath), max(g["dur"] for g in gath)))
    print("epoch span %.1f us = %.1f us/step" % (ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"], (ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) / len(ticks)))
    if a.full:
        t0 = ticks[-3]
        short = lambda e: e["name"].replace("(anonymous namespace)::", "").split("(")[0][:50]
        for e in ev:
            if e["ts"] >= t0 - 5 and e["ts"] < ticks[-1]:
                print("%8.1f %7.1f  s%-3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), short(e)))


if __name__ == "__main__":
    main()
