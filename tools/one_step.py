#!/usr/bin/env python
"""One eager CFG-batched VideoUNet forward (the body of one sampler step) at the bench shapes, nothing else: the target
of the ncu launch list / full captures under profiles/ (`ncu ... python tools/one_step.py [--stage 2] [--reps N]`)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hi3d_official_b200 import configs, spec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, default=1)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = configs.build_engine(args.stage, device=dev)
    spec.synth_fill_(model, seed=0, fast=True)
    wl = bench.workload(args.stage)
    plan = model.model.diffusion_model.get_plan(2 * bench.T_FRAMES, wl["h"], wl["h"], bench.T_FRAMES)
    torch.cuda.synchronize()
    for _ in range(args.reps):
        for s in plan.steps:
            s()
    torch.cuda.synchronize()
    print(f"one_step: stage {args.stage}, {len(plan.steps)} ops x {args.reps}")


if __name__ == "__main__":
    main()
