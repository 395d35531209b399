#!/usr/bin/env python
"""One spatial-attention launch (plus warm-up) for an ncu capture of the FMHA kernel alone:
    ncu --set full --import-source on -k regex:fmha --launch-skip 2 --launch-count 1 -o x python tools/one_attn.py"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hi3d_official_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=16384)
ap.add_argument("--heads", type=int, default=5)
ap.add_argument("--n-img", type=int, default=4)
a = ap.parse_args()
C = a.heads * 64
qkv = torch.randn(a.n_img * a.L, 3 * C, device="cuda", dtype=torch.float16)
out = torch.empty(a.n_img * a.L, C, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.attention_d64(qkv, a.n_img, a.L, a.heads, out, engine="tc5")
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
