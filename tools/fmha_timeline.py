#!/usr/bin/env python
"""Clock-stamp timeline of one CTA of the lean split FMHA kernel (variant 5): where a half-tile pipeline spends its period.
    python tools/fmha_timeline.py [--L 16384 --heads 5 --n-img 32]
Events per key tile j (cycles relative to the first stamp): softmax warp of quarter 0, half h: S ready / exps done / P
announced; MMA warp, half h: P seen / operands ready (K of tile j+1 landed) / MMAs + commits issued."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hi3d_official_b200 import _native, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=16384)
ap.add_argument("--heads", type=int, default=5)
ap.add_argument("--n-img", type=int, default=32)
a = ap.parse_args()
lib = _native.load()
C = a.heads * 64
qkv = torch.randn(a.n_img * a.L, 3 * C, device="cuda", dtype=torch.float16)
out = torch.empty(a.n_img * a.L, C, device="cuda", dtype=torch.float16)
dbg = torch.zeros(128, device="cuda", dtype=torch.int64)
_native.check(lib.hi3d_attention_tc5_set_exp_emulation(1), "emu")
_native.check(lib.hi3d_attention_tc5_set_variant(5), "variant")
_native.check(lib.hi3d_attention_tc5_set_debug_buffer(dbg.data_ptr()), "dbg")
for _ in range(2):
    ops.attention_d64(qkv, a.n_img, a.L, a.heads, out, engine="tc5")
torch.cuda.synchronize()
_native.check(lib.hi3d_attention_tc5_set_debug_buffer(None), "dbg")
_native.check(lib.hi3d_attention_tc5_set_variant(2), "variant")
t = dbg.cpu().view(8, 16)
t0 = int(t[0][t[0] > 0].min())
names = ["sm0 S ready", "sm0 exps done", "sm0 P announced", "sm1 S ready", "sm1 exps done", "sm1 P announced",
         "mma P0 seen", "mma K(j+1) ok (0)", "mma 0 issued", "mma P1 seen", "mma K(j+1) ok (1)", "mma 1 issued"]
print("tile " + " ".join(f"{n:>18s}" for n in names))
for j in range(8):
    print(f"{16 + j:4d} " + " ".join(f"{int(t[j][e]) - t0:18d}" for e in range(12)))
per = (int(t[7][0]) - int(t[0][0])) / 7
print(f"period per tile (half 0): {per:.0f} cycles")
for e0, e1, what in ((0, 1, "sm0: S ready -> exps done"), (1, 2, "sm0: exps done -> P announced"), (2, 6, "P0 announced -> mma sees it"),
                     (6, 7, "mma: wait K(j+1)"), (7, 8, "mma: issue P V + S + commits"), (3, 4, "sm1: S ready -> exps done"),
                     (5, 9, "P1 announced -> mma sees it"), (10, 11, "mma (half 1): issue")):
    d = [int(t[j][e1]) - int(t[j][e0]) for j in range(8)]
    print(f"{what:36s} mean {sum(d) / 8:7.0f}  min {min(d):6d}  max {max(d):6d}")
d = [int(t[j + 1][0]) - int(t[j][8]) for j in range(7)]
print(f"{'mma 0 issued -> sm0 S(j+1) ready':36s} mean {sum(d) / 7:7.0f}  min {min(d):6d}  max {max(d):6d}")
d = [int(t[j + 1][3]) - int(t[j][11]) for j in range(7)]
print(f"{'mma 1 issued -> sm1 S(j+1) ready':36s} mean {sum(d) / 7:7.0f}  min {min(d):6d}  max {max(d):6d}")
