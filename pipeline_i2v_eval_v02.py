#!/usr/bin/env python
"""Stage-2 (1024^2 refiner) entry point, same CLI as the reference's pipeline_i2v_eval_v02.py (:38-44) on the B200
engine.  Reads <output_dir>/first_step/first.pt (stage-1 frames written by pipeline_i2v_eval_v01.py), up-samples them
to 1024^2, VAE-encodes each frame (posterior sample, CPU RNG like the reference), runs the 25-step re-noise/blend loop
of pipeline_i2v_eval_v02.py:127-135 on the fused sampler and decodes.  Conditioning: --cond / --synthetic as in v01."""
import argparse
import os
import random

import torch
import torch.nn.functional as F

from pipeline_i2v_eval_v01 import load_model, save_frames, synthetic_cond


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--denoise_config", type=str, default="configs/inference-v02.yaml")
    ap.add_argument("--denoise_checkpoint", type=str, default="ckpts/second_stage.pt")
    ap.add_argument("--image_path", type=str, default="demo/15_out.png")
    ap.add_argument("--output_dir", type=str, default="outputs/15_out")
    ap.add_argument("--elevation", type=int, default=0)
    ap.add_argument("--cond", type=str, default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--seed", type=int, default=None)
    params = ap.parse_args()
    seed = random.randint(0, 65535) if params.seed is None else params.seed
    torch.manual_seed(seed)
    model = load_model(params.denoise_config, params.denoise_checkpoint, 2)
    T, h = model.num_samples, 128
    first = os.path.join(params.output_dir, "first_step", "first.pt")
    if os.path.exists(first):
        frames = torch.load(first).cuda().float()
        frames = F.interpolate(frames, size=(8 * h, 8 * h), mode="bilinear", align_corners=False)   # cv2.resize, v02:186
    elif params.synthetic:
        frames = torch.rand(T, 3, 8 * h, 8 * h, device="cuda") * 2 - 1
    else:
        raise SystemExit(f"{first} not found (run pipeline_i2v_eval_v01.py first) and --synthetic not given")
    if params.cond:
        d = torch.load(params.cond, map_location="cuda")
        c, uc = d["c"], d["uc"]
    elif params.synthetic:
        c, uc = synthetic_cond(2, T, h, "cuda", seed)
    else:
        raise SystemExit("the conditioner towers are outside the B200 hot path: pass --cond <file> or --synthetic")
    with torch.no_grad():
        init_latents = torch.randn(T, 4, h, h, device="cuda")                                      # v02:93
        z = torch.cat([model.encode_first_stage(frames[t:t + 1].half()) for t in range(T)], 0)     # v02:96-101
        out = model.sample_stage2(c, uc, init_latents, z.float())                                  # v02:103-137
    save_frames(out, os.path.join(params.output_dir, "second_step_video"), "second")
    print(f"[hi3d-b200] wrote {T} frames {tuple(out.shape[1:])} to {params.output_dir}/second_step_video (seed {seed})")


if __name__ == "__main__":
    main()
