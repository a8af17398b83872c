#!/usr/bin/env python
"""Kernel timeline of the captured training step (torch.profiler / CUPTI): per-kernel start, duration and stream for
one CUDA-graph replay, so that gaps and overlap between the dgrad chain and the side stream can be read off.

    python scripts/step_timeline.py [--config 2] [--rows N] [--summary]

Single GPU: as above.  Data parallel: torchrun --nproc-per-node N scripts/step_timeline.py (rank 0 prints its own
timeline, which then includes the cross-GPU barriers and the exchange kernel)."""
import argparse
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import configs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--summary", action="store_true", help="also print the per-kernel-name totals")
    args = ap.parse_args()
    cfg = configs.CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    B = args.rows or (cfg["batch"] // world if cfg["scaling"] == "strong" else cfg["batch"])
    eng = configs.make_engine(cfg, B)
    if world > 1:
        from dmvae_b200.dp import DataParallel
        DataParallel(eng, mode="auto")
    opt = eng.optimizer("train", 0.002)
    xs = torch.from_numpy(configs.synth_inputs(cfg, B, seed=1 + rank)).cuda()
    ys = None
    if cfg["model"] == "dmoe":
        ys = torch.nn.functional.one_hot(torch.arange(B) % cfg["output_dim"], cfg["output_dim"]).float().cuda()

    def step():
        if ys is None:
            eng.train_step(xs, B, opt)
        else:
            eng.moe_step(xs, ys, B, opt)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        if rank != 0:
            torch.cuda.synchronize()
            os._exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", "step_trace.json")
    prof.export_chrome_trace(out)
    ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    if not ev:
        print("no kernel records")
        return
    # last replay only
    step_starts = [e["ts"] for e in ev if "step_tick" in e["name"]]
    t0 = step_starts[-1] if step_starts else ev[len(ev) * 2 // 3]["ts"]
    rows = [e for e in ev if e["ts"] >= t0]
    end = max(e["ts"] + e["dur"] for e in rows)
    print("%s rows %d: step span %.1f us, %d kernels, sum of durations %.1f us" %
          (cfg["name"], B, end - t0, len(rows), sum(e["dur"] for e in rows)))
    short = lambda e: e["name"].replace("(anonymous namespace)::", "").split("(")[0][:60]
    for e in rows:
        print("%8.1f %7.1f  s%-3s %s  grid=%s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), short(e), e["args"].get("grid", "")))
    if args.summary:
        tot = {}
        for e in rows:
            k = short(e)
            n, d = tot.get(k, (0, 0.0))
            tot[k] = (n + 1, d + e["dur"])
        print("per kernel name: launches, total us")
        for k, (n, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            print("  %-62s %3d %8.1f" % (k, n, d))
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
