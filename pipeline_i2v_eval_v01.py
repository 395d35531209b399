#!/usr/bin/env python
"""Stage-1 entry point, same CLI as the reference's pipeline_i2v_eval_v01.py (:39-45) on the B200 engine.

    python pipeline_i2v_eval_v01.py --denoise_config configs/inference-v01.yaml --denoise_checkpoint ckpts/first_stage.pt \
        --image_path demo/15_out.png --output_dir outputs/15_out --elevation 0  [--cond cond.pt | --synthetic]

The hot path (25-step fused Euler-EDM over VideoUNet + VAE decode) runs here; the conditioner towers (rembg, OpenCLIP,
aesthetic MLP) are outside it: pass their output as --cond (torch.save({'c': .., 'uc': ..}) from the reference's
`conditioner.get_unconditional_conditioning`, pipeline_i2v_eval_v01.py:74-78), or --synthetic for seeded stand-ins.
Without a checkpoint file the seeded synthetic weights of spec.synth_fill_ are used (said loudly).
"""
import argparse
import os
import random

import torch

from hi3d_official_b200 import configs, spec
from hi3d_official_b200.engine import create_model
from hi3d_official_b200.util import get_obj_from_str


def load_model(config_path, ckpt, stage):
    if os.path.exists(config_path):
        model = create_model(config_path)
    else:
        print(f"[hi3d-b200] {config_path} not found: using the built-in copy of the stage-{stage} inference config")
        cfg = (configs.stage1_config() if stage == 1 else configs.stage2_config())["model"]
        model = get_obj_from_str(cfg["target"])(**cfg["params"])
    if os.path.exists(ckpt):
        model.init_from_ckpt(ckpt)
        model = model.cuda().half()
    else:
        print(f"[hi3d-b200] checkpoint {ckpt} not found: SEEDED SYNTHETIC WEIGHTS (outputs are not images)")
        model = model.cuda().half()
        spec.synth_fill_(model, seed=0, fast=True)
    return model


def synthetic_cond(stage, T, h, device, seed):
    g = torch.Generator().manual_seed(seed)
    adm, cc = (768, 4) if stage == 1 else (512, 13)
    c = dict(crossattn=torch.randn(1, 1, 1024, generator=g), vector=torch.randn(1, adm, generator=g),
             concat=(torch.randn(T, cc, h, h, generator=g) * 0.18).half())
    c = {k: v.to(device) for k, v in c.items()}
    uc = dict(crossattn=torch.zeros_like(c["crossattn"]), vector=c["vector"].clone(), concat=torch.zeros_like(c["concat"]))
    return c, uc


def save_frames(frames, out_dir, name):
    """frames: (T, 3, H, W) in [-1, 1] -> <out_dir>/<name>.pt (+ PNGs when Pillow is importable)."""
    os.makedirs(out_dir, exist_ok=True)
    torch.save(frames.cpu(), os.path.join(out_dir, name + ".pt"))
    try:
        from PIL import Image
        for t, f in enumerate(frames):
            arr = ((f.float().clamp(-1, 1) + 1) * 127.5).permute(1, 2, 0).byte().cpu().numpy()
            Image.fromarray(arr).save(os.path.join(out_dir, f"{name}_{t:02d}.png"))
    except ImportError:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--denoise_config", type=str, default="configs/inference-v01.yaml")
    ap.add_argument("--denoise_checkpoint", type=str, default="ckpts/first_stage.pt")
    ap.add_argument("--image_path", type=str, default="demo/15_out.png")
    ap.add_argument("--output_dir", type=str, default="outputs/15_out")
    ap.add_argument("--elevation", type=int, default=0)
    ap.add_argument("--cond", type=str, default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--seed", type=int, default=None)
    params = ap.parse_args()
    seed = random.randint(0, 65535) if params.seed is None else params.seed      # v01:33-34
    torch.manual_seed(seed)
    model = load_model(params.denoise_config, params.denoise_checkpoint, 1)
    T, h = model.num_samples, 64                                                 # 16 frames, 512^2 / 8
    if params.cond:
        d = torch.load(params.cond, map_location="cuda")
        c, uc = d["c"], d["uc"]
    elif params.synthetic:
        c, uc = synthetic_cond(1, T, h, "cuda", seed)
    else:
        raise SystemExit("the conditioner towers are outside the B200 hot path: pass --cond <file> or --synthetic")
    randn = torch.randn(T, 4, h, h, device="cuda")                               # v01:91
    with torch.no_grad():
        frames = model.sample_stage1(c, uc, randn)                               # v01:92-94
    save_frames(frames, os.path.join(params.output_dir, "first_step"), "first")
    print(f"[hi3d-b200] wrote {T} frames {tuple(frames.shape[1:])} to {params.output_dir}/first_step (seed {seed})")


if __name__ == "__main__":
    main()
