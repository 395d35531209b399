#!/usr/bin/env python
"""Algorithmic FLOPs of the hot path, from the layer plan alone -- no GPU, no weights.

Two counts per CFG-batched UNet forward (N = 32 = 2 CFG x 16 frames):
  * reference graph: what the reference's modules execute (incl. the single-token cross-attention q / out projections
    that this implementation folds into one per-sample row);
  * this implementation: the GEMM + attention-core work of the launch plan (what bench.py's roofline uses).
Plus the per-level split the microbenchmarks refer to.  `python tools/count_flops.py [--stage 1|2]`."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hi3d_official_b200 import configs, spec  # noqa: E402


def count(stage: int):
    kw = configs.UNET_STAGE1 if stage == 1 else configs.UNET_STAGE2
    cfg = spec.UNetConfig.from_kwargs(**kw)
    plan = spec.unet_plan(cfg)
    latent = bench.workload(stage)["h"]
    N, T = 2 * bench.T_FRAMES, bench.T_FRAMES
    per = collections.OrderedDict()

    def add(key, fl):
        per[key] = per.get(key, 0.0) + fl

    folded = 0.0
    for blk in plan.input_blocks + [plan.middle] + plan.output_blocks:
        for L in blk:
            hw = (latent // L.ds) ** 2
            M = N * hw
            lvl = f"ds{L.ds}"
            if L.kind in ("conv_in", "down", "up"):
                mo = M // 4 if L.kind == "down" else (M * 4 if L.kind == "up" else M)
                # nearest-x2 + conv3x3 runs as four 2x2 parity convs on the source grid: 4/9 of the reference FLOPs
                ref = 2.0 * mo * L.cout * 9 * L.cin
                add(f"{lvl} conv_in/down/up", ref)
                if L.kind == "up":
                    folded += ref * (1.0 - 4.0 / 9.0)
            elif L.kind == "res":
                add(f"{lvl} resblock conv3x3", 2.0 * M * L.cout * 9 * L.cin + 2.0 * M * L.cout * 9 * L.cout)
                if L.cin != L.cout:
                    add(f"{lvl} resblock skip 1x1", 2.0 * M * L.cout * L.cin)
                add(f"{lvl} resblock conv3d(3,1,1)", 2 * (2.0 * M * L.cout * 3 * L.cout))
            elif L.kind == "attn":
                C = L.cin
                add(f"{lvl} transformer proj_in/out", 2.0 * M * C * C * 2)
                add(f"{lvl} attention qkv/out projections", 2 * 2.0 * M * C * 4 * C)
                add(f"{lvl} GEGLU feed-forward (x2 + ff_in)", 2 * 2.0 * M * C * 12 * C + 2.0 * M * C * 12 * C)
                add(f"{lvl} spatial attention core", 4.0 * N * hw * hw * C)
                add(f"{lvl} temporal attention core", 4.0 * N * hw * T * C)
                # reference only: attn2 to_q and to_out on every token of both blocks (one key/value token -> folded)
                xa = 2 * 2 * 2.0 * M * C * C
                add(f"{lvl} cross-attention q/out (reference only)", xa)
                folded += xa
    total_ref = sum(per.values())
    return per, total_ref, total_ref - folded


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, default=0, help="1 | 2 (default: both)")
    args = ap.parse_args()
    for st in ([1, 2] if args.stage == 0 else [args.stage]):
        per, ref, ours = count(st)
        print(f"== stage {st} ({bench.workload(st)['name']}), one CFG-batched UNet forward")
        for k, v in per.items():
            print(f"  {k:52s} {v / 1e9:12.1f} GFLOP  {100 * v / ref:5.1f} %")
        print(f"  reference graph: {ref:.4e} FLOP   this implementation (after folds): {ours:.4e} FLOP   "
              f"bench.unet_step_flops: {bench.unet_step_flops(st, bench.workload(st)['h']):.4e}")


if __name__ == "__main__":
    main()
