#!/usr/bin/env python
"""Short program for ncu: a few captured training steps of one BASELINE config (python scripts/ncu_step.py --config N
[--steps 3]).  Under ncu every kernel node of the replayed graph is profiled (cold cache, serialised: compare shares)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import configs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    cfg = configs.CONFIGS[args.config]
    B = cfg["batch"]
    eng = configs.make_engine(cfg, B)
    opt = eng.optimizer("train", 0.002)
    xs = torch.from_numpy(configs.synth_inputs(cfg, B)).cuda()
    ys = None
    if cfg["model"] == "dmoe":
        ys = torch.nn.functional.one_hot(torch.arange(B) % cfg["output_dim"], cfg["output_dim"]).float().cuda()
    for i in range(2 + args.steps):                      # eager step, capture, then `steps` replays
        if i == 2:                                       # ncu --profile-from-start off: only the replays are profiled
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        if ys is None:
            eng.train_step(xs, B, opt)
        else:
            eng.moe_step(xs, ys, B, opt, graph=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ok loss %.3f launches %d" % (float(eng.loss_out[3]), eng.launches()))


if __name__ == "__main__":
    main()
