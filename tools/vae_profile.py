#!/usr/bin/env python
"""Per-launch timing of the AutoencoderKL decoder plan (CUDA events around every plan step, host kept ahead of the GPU):
    python tools/vae_profile.py [--latent 128] [--frames 1]
prints the decode time of `frames` frames decoded together, per kernel class and per GEMM shape."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hi3d_official_b200 import configs, ops, spec  # noqa: E402
from hi3d_official_b200.vae import AutoencoderKL  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--latent", type=int, default=128)
ap.add_argument("--frames", type=int, default=1)
a = ap.parse_args()
dd = dict(configs._VAE_DD) if hasattr(configs, "_VAE_DD") else dict(
    attn_type="vanilla-xformers", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
    ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
ae = AutoencoderKL(embed_dim=4, ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, monitor="val/rec_loss")
ae.load_state_dict(spec.synth_state_dict(spec.vae_param_shapes(spec.VAEConfig.from_ddconfig(dd, 4)), seed=3), strict=True)
ae = ae.cuda().half()
z = torch.randn(a.frames, 4, a.latent, a.latent, device="cuda", dtype=torch.float16)
for _ in range(2):
    out = ae.decode(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = ae.decode(z)
e1.record()
torch.cuda.synchronize()
print(f"decode({a.frames} x 4 x {a.latent} x {a.latent}) = {e0.elapsed_time(e1) / 3:.2f} ms  ({e0.elapsed_time(e1) / 3 / a.frames:.2f} ms per frame)")
plan = ae._plan("dec", a.frames, a.latent, a.latent)
torch.cuda._sleep(int(1.2e8))
recs = []
for s in plan.steps:
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(); s(); a1.record()
    recs.append((s, a0, a1))
torch.cuda.synchronize()
cls, shapes = {}, {}
for s, a0, a1 in recs:
    ms = a0.elapsed_time(a1)
    if isinstance(s, ops.Gemm):
        k = "gemm"
        pp = s.p
        key = f"mode{pp.mode} M={pp.M} N={pp.N} K={pp.K}" + (" up" if pp.out_up else "")
        d = shapes.setdefault(key, [0.0, 0, 0.0])
        d[0] += ms; d[1] += 1; d[2] += s.flops
    else:
        k = getattr(s, "kind", None) or getattr(getattr(s, "__func__", s), "__name__", "other")
    d = cls.setdefault(k, [0.0, 0])
    d[0] += ms; d[1] += 1
tot = sum(v[0] for v in cls.values())
print(f"plan: {len(plan.steps)} launches, sum {tot:.2f} ms, {plan.flops / 1e12:.2f} TFLOP")
for k, v in sorted(cls.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {v[0]:8.2f} ms  x{v[1]}")
for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0]):
    print(f"  {v[0]:8.2f} ms x{v[1]:3d} {v[2] / v[0] / 1e9 if v[0] else 0:7.1f} TFLOP/s  {k}")
