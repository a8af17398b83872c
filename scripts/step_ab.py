#!/usr/bin/env python
"""A/B of the captured training step under environment switches: python scripts/step_ab.py [--config 2] VAR=VAL ...
Each switch set runs in a fresh process (the switches are read at import / first launch); prints ms per step."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "scripts"))
import configs as CFG
cfg = CFG.CONFIGS[%d]; B = cfg["batch"]
eng = CFG.make_engine(cfg, B); opt = eng.optimizer("train", 0.002)
xs = torch.from_numpy(CFG.synth_inputs(cfg, B)).cuda()
for _ in range(10): eng.train_step(xs, B, opt)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): eng.train_step(xs, B, opt)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 200)
print("%%.4f ms/step  loss %%.3f" %% (best, float(eng.loss_out[3])))
'''


def main():
    args = sys.argv[1:]
    config = 2
    if args and args[0] == "--config":
        config = int(args[1])
        args = args[2:]
    sets = [""] + args
    for sset in sets:
        env = dict(os.environ)
        for kv in sset.split(","):
            if "=" in kv:
                k, v = kv.split("=", 1)
                env[k] = v
        out = subprocess.run([sys.executable, "-c", CODE % (ROOT, ROOT, config)], env=env, capture_output=True, text=True)
        print("%-50s %s" % (sset or "(default)", (out.stdout.strip().splitlines() or [out.stderr.strip()[-300:]])[-1]), flush=True)


if __name__ == "__main__":
    main()
