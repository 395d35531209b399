#!/usr/bin/env python
"""bench.py -- Hi3D denoising hot path on B200.

One "step" = one orbital video of the workload: 25 Euler-EDM sampler steps of the CFG-batched VideoUNet
(N = 2 x 16 frames) followed by the AutoencoderKL decode of the 16 frames.  Metric = multi-view frames / s.
Default workload (N=1): BASELINE.json configs[1], first-stage 16 x 512 x 512 (latents 16 x 4 x 64 x 64), fp16
(the reference's inference dtype and the parity dtype).  `--stage 2` selects configs[2].
With --gpus N > 1 (torchrun) every rank runs its own video (BASELINE configs[4] style data parallel, weak scaling,
no data-path collective); time = max over ranks, value = all videos / that time.

`--shard frames` instead shards the 16 frames of ONE video over the ranks (strong scaling, NCCL exchanges in the step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--stage 1|2] [--shard videos|frames] [--impl reference]

stdout carries exactly one line, the JSON (libraries that print to fd 1 are redirected to stderr).  Keys beyond the
contract: `unet_ms_per_sampler_step` (one eager UNet forward with the host kept ahead of the GPU = the cost of a sampler
step inside the graph), `kernel_breakdown` (CUDA events around every launch of one instrumented forward: per kernel class
ms / share / TFLOP/s / GB/s, the top GEMM shapes, and the sum of launches), `roofline` (all GEMM launches of that forward
against the measured sustained bf16 peak of MEASURED_PEAKS.json).

--impl reference times the oracle port of the reference's CPU path (the reference itself is pure Python and cannot
travel to the GPU box) on the host cores (<= 16 threads: more oversubscribe the small sample), on a bounded sample, and
prints the same JSON line with impl=reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "multi-view frames/sec (16f, 25-step EDM)"
T_FRAMES = 16
NUM_STEPS = 25


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------------------
def workload(stage: int):
    if stage == 1:
        return dict(name="first-stage inference-v01 16x512x512, 25 EDM steps", h=64, cc=4, adm=768, max_scale=2.5)
    return dict(name="second-stage inference-v02 16x1024x1024 depth-concat, 25 EDM steps", h=128, cc=13, adm=512,
                max_scale=2.0)


def make_host_inputs(stage: int, seed: int, pin: bool):
    import torch
    wl = workload(stage)
    g = torch.Generator().manual_seed(1234 + seed)
    h = wl["h"]
    t = dict(randn=torch.randn(T_FRAMES, 4, h, h, generator=g),
             crossattn=torch.randn(1, 1, 1024, generator=g),
             vector=torch.randn(1, wl["adm"], generator=g),
             concat=(torch.randn(T_FRAMES, wl["cc"], h, h, generator=g) * 0.18).half())
    if stage == 2:
        t["z"] = torch.randn(T_FRAMES, 4, h, h, generator=g) * 0.18
    if pin:
        t = {k: v.pin_memory() for k, v in t.items()}
    return t


def to_cond(dev_t):
    import torch
    c = dict(crossattn=dev_t["crossattn"], vector=dev_t["vector"], concat=dev_t["concat"])
    uc = dict(crossattn=torch.zeros_like(c["crossattn"]), vector=c["vector"], concat=torch.zeros_like(c["concat"]))
    return c, uc


def run_video(model, stage: int, dev_t, shard=None):
    c, uc = to_cond(dev_t)
    if stage == 1:
        return model.sample_stage1(c, uc, dev_t["randn"].clone(), shard=shard)
    return model.sample_stage2(c, uc, dev_t["randn"].clone(), dev_t["z"], shard=shard)


def shard_frames(t: dict, rank: int, world: int) -> dict:
    """This rank's frames of the per-frame tensors (BASELINE configs[3]: one video, frames sharded over GPUs)."""
    tl = T_FRAMES // world
    sl = slice(rank * tl, (rank + 1) * tl)
    return {k: (v[sl].contiguous() if v.shape[0] == T_FRAMES else v) for k, v in t.items()}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
# (not measurable live): mean over the 10 hi3d_gemm_tc5 launches of input block 1 (64x64 level: conv3x3, temporal conv,
# proj_in, qkv, attention out, GEGLU, ff2 ...), each 84-420 MB algorithmic.  Stage 2 was not captured.
NCU_GEMM_TRAFFIC = {1: 292.8e6}
NCU_GEMM_TRAFFIC_NOTE = ("bytes per launch, mean of the 10 top-level GEMM launches in profiles/r01_ncu_full_stage1_final.txt "
                         "(ncu --set full, caches flushed per pass); `achieved` averages all 300 GEMM launches of a forward")


def kernel_breakdown(model, stage: int, dev_t, peaks):
    """One extra, instrumented sampler step: CUDA events around every launch of the UNet plan on the launching
    stream -> per kernel-class time, algorithmic FLOPs / bytes, achieved rate."""
    import torch
    from hi3d_official_b200 import ops
    unet = model.model.diffusion_model
    wl = workload(stage)
    plan = unet.get_plan(2 * T_FRAMES, wl["h"], wl["h"], T_FRAMES)
    recs = []
    # queue ~60 ms of spinning first so the host gets ahead of the GPU: otherwise the short kernels (norms, small GEMMs)
    # are bracketed together with the idle time the GPU spends waiting for their launch
    torch.cuda._sleep(int(1.2e8))
    for s in plan.steps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s(); e1.record()
        recs.append((s, e0, e1))
    torch.cuda.synchronize()
    # the whole UNet forward back to back (host far ahead of the GPU) = what one sampler step costs inside the graph
    torch.cuda._sleep(int(4e7))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(3):
        for s in plan.steps:
            s()
    f1.record()
    torch.cuda.synchronize()
    fwd_ms = f0.elapsed_time(f1) / 3
    cls = {}
    for s, e0, e1 in recs:
        ms = e0.elapsed_time(e1)
        if isinstance(s, ops.Gemm):
            k, fl, by = "gemm(conv/linear)", s.flops, 0.0
        else:
            k, fl, by = getattr(s, "kind", "other"), getattr(s, "flops", 0.0), getattr(s, "bytes", 0.0)
        d = cls.setdefault(k, dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
        d["ms"] += ms; d["launches"] += 1; d["flops"] += fl; d["bytes"] += by
    tot = sum(d["ms"] for d in cls.values())
    # per GEMM shape (rows mode, M, N, K, activation): where the implicit-GEMM time goes
    shapes = {}
    for s, e0, e1 in recs:
        if isinstance(s, ops.Gemm):
            pp = s.p
            key = f"mode{pp.mode} M={pp.M} N={pp.N} K={pp.K} act={pp.act}" + (" up" if pp.out_up else "") + \
                  (" s2" if pp.stride == 2 else "")
            d = shapes.setdefault(key, dict(ms=0.0, launches=0, flops=0.0))
            d["ms"] += e0.elapsed_time(e1); d["launches"] += 1; d["flops"] += s.flops
    top = sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:14]
    out = {}
    out["gemm_shapes_top"] = {k: dict(ms=round(d["ms"], 3), launches=d["launches"], tflops=round(d["flops"] / d["ms"] / 1e9, 1))
                              for k, d in top}
    for k, d in cls.items():
        out[k] = dict(ms=round(d["ms"], 3), share=round(d["ms"] / tot, 4), launches=d["launches"])
        if d["flops"]:
            out[k]["tflops"] = round(d["flops"] / d["ms"] / 1e9, 1)
        if d["bytes"]:
            out[k]["gbs"] = round(d["bytes"] / d["ms"] / 1e6, 1)
            out[k]["hbm_frac"] = round(d["bytes"] / d["ms"] / 1e6 / peaks["hbm"], 4)
    g = cls.get("gemm(conv/linear)")
    roof = None
    if g:
        ach = g["flops"] / g["ms"] / 1e9
        roof = dict(bound="tensor", kernel="hi3d_gemm (implicit-GEMM conv/linear, all launches of one UNet step)",
                    achieved=round(ach, 1), peak=peaks["tf_sust"], unit="TFLOP/s", frac=round(ach / peaks["tf_sust"], 4),
                    peak_source=f"{peaks['src']} bf16_tflops_sustained", traffic=NCU_GEMM_TRAFFIC.get(stage),
                    traffic_note=NCU_GEMM_TRAFFIC_NOTE if stage in NCU_GEMM_TRAFFIC else None,
                    flops_per_step=g["flops"], avg_launch_ms=round(g["ms"] / g["launches"], 4))
    out["sum_of_launches_ms"] = round(tot, 3)
    return out, roof, fwd_ms


def cpu_baseline_sample(stage: int, latent: int, steps: int, warmup: int):
    """Oracle port (oracle/hi3d_oracle.py: plain-PyTorch fp32 restatement of the reference path) on the host cores:
    `steps` timed Euler steps (CFG-batched UNet forward + guider + Euler) at a reduced latent size."""
    import torch
    from hi3d_official_b200 import configs, spec
    from oracle import hi3d_oracle as O
    # bounded thread count: on the 128-core GPU host the 8x8-latent sample is all tiny ops and 128 OpenMP threads made one
    # step take minutes (oversubscription); 16 threads is what the reported `cores` says
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    kw = configs.UNET_STAGE1 if stage == 1 else configs.UNET_STAGE2
    cfg = spec.UNetConfig.from_kwargs(**kw)
    sd = spec.synth_state_dict(spec.unet_param_shapes(cfg), seed=1)
    wl = workload(stage)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(T_FRAMES, 4, latent, latent, generator=g)
    c = dict(crossattn=torch.randn(1, 1, 1024, generator=g), vector=torch.randn(1, wl["adm"], generator=g),
             concat=torch.randn(T_FRAMES, wl["cc"], latent, latent, generator=g) * 0.18)
    uc = dict(crossattn=torch.zeros(1, 1, 1024), vector=c["vector"], concat=torch.zeros_like(c["concat"]))
    scale = O.guider_scale(T_FRAMES, wl["max_scale"])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.time()
            O.euler_step(sd, x, 10.0, 8.0, c, uc, scale, num_video_frames=T_FRAMES)
            if i >= warmup:
                times.append(time.time() - t0)
    return sum(times) / len(times)


def unet_step_flops(stage: int, latent: int) -> float:
    """Algorithmic FLOPs of one CFG-batched UNet forward (GEMM + attention cores), from the plan description."""
    # analytic: use the same counting as _Plan.flops without touching the GPU
    from hi3d_official_b200 import configs, spec
    kw = configs.UNET_STAGE1 if stage == 1 else configs.UNET_STAGE2
    cfg = spec.UNetConfig.from_kwargs(**kw)
    plan = spec.unet_plan(cfg)
    N, T = 2 * T_FRAMES, T_FRAMES
    fl = 0.0
    for blk in plan.input_blocks + [plan.middle] + plan.output_blocks:
        for L in blk:
            hw = (latent // L.ds) ** 2
            M = N * hw
            if L.kind in ("conv_in", "down", "up"):
                mo = M // 4 if L.kind == "down" else (M * 4 if L.kind == "up" else M)
                fl += 2.0 * mo * L.cout * 9 * L.cin
            elif L.kind == "res":
                fl += 2.0 * M * L.cout * 9 * L.cin + 2.0 * M * L.cout * 9 * L.cout
                if L.cin != L.cout:
                    fl += 2.0 * M * L.cout * L.cin
                fl += 2 * (2.0 * M * L.cout * 3 * L.cout)
            elif L.kind == "attn":
                C = L.cin
                per_blk = 2.0 * M * C * (3 * C + C + 8 * C + 4 * C)       # qkv, out, ff1, ff2
                fl += 2.0 * M * C * C * 2                                 # proj_in / proj_out
                fl += per_blk + per_blk + 2.0 * M * C * 12 * C            # spatial, temporal, + ff_in
                fl += 4.0 * N * hw * hw * C + 4.0 * N * hw * T * C        # attention cores
    return fl


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--stage", type=int, default=1, choices=(1, 2))
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--engine", default=os.environ.get("HI3D_ENGINE", "tc5"), choices=("mma", "tc5"))
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--shard", default="videos", choices=("videos", "frames"),
                    help="N > 1: 'videos' = one video per GPU (weak scaling, default); 'frames' = ONE video with its 16 "
                         "frames sharded over the GPUs (strong scaling; K/V all-gather + halo + GN all-reduce per layer)")
    ap.add_argument("--cpu-latent", type=int, default=8, help="latent size of the bounded CPU sample")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: anything a library writes to fd 1 in between (NCCL prints its version
    # banner there) goes to stderr instead
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(_real_stdout, (json.dumps(line) + "\n").encode())
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    wl = workload(args.stage)
    config = {"workload": wl["name"], "frames": T_FRAMES, "sampler_steps": NUM_STEPS, "latent": [T_FRAMES, 4, wl["h"], wl["h"]],
              "cfg_batch": 2 * T_FRAMES, "vae_decode_in_step": True, "parallelism": f"dp{world} (one video per GPU)",
              "l2": "activations >> L2 (UNet working set > 1 GB per step); no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        lat = args.cpu_latent
        t = cpu_baseline_sample(args.stage, lat, args.steps, args.warmup)
        ratio = unet_step_flops(args.stage, wl["h"]) / unet_step_flops(args.stage, lat)
        fps = T_FRAMES / (NUM_STEPS * t * ratio)
        cores = min(os.cpu_count() or 1, 16)     # threads cpu_baseline_sample() actually uses
        sample = (f"{args.steps} timed Euler steps (CFG-batched full-width VideoUNet fwd, fp32, oracle port) at {lat}x{lat} "
                  f"latents = {t:.2f} s/step; projected to {wl['h']}x{wl['h']} latents by the UNet FLOP ratio {ratio:.1f} "
                  f"x 25 steps (VAE decode not included)")
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from hi3d_official_b200 import _native, configs, spec
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout = the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    model = configs.build_engine(args.stage, device=dev)
    spec.synth_fill_(model, seed=0, fast=True)
    model.model.diffusion_model.set_engine(args.engine)
    model.first_stage_model.set_engine(args.engine)
    peaks = _peaks()

    frames_mode = args.shard == "frames" and world > 1
    shard = (rank, world) if frames_mode else None
    host = make_host_inputs(args.stage, seed=0 if frames_mode else rank, pin=True)
    if frames_mode:
        if T_FRAMES % world:
            raise SystemExit(f"--shard frames needs 16 % world == 0, got {world}")
        host = {k: v.pin_memory() for k, v in shard_frames(host, rank, world).items()}
        config["parallelism"] = f"frames sharded over {world} GPUs ({T_FRAMES // world} per GPU), NCCL all-gather/halo/all-reduce"
    dev_t = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    n_local = T_FRAMES // world if frames_mode else T_FRAMES
    out_host = torch.empty(n_local, 3, wl["h"] * 8, wl["h"] * 8, dtype=torch.float16).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _native.launch_count()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _native.launch_count() - l0

    def step_resident():
        run_video(model, args.stage, dev_t, shard)

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        frames = run_video(model, args.stage, d, shard)
        out_host.copy_(frames, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3) if args.warmup else 0):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches = timed(step_resident, args.steps)
    clk = clocks.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    breakdown = roof = None
    unet_ms = None
    if rank == 0 and not args.no_breakdown and not frames_mode:
        breakdown, roof, unet_ms = kernel_breakdown(model, args.stage, dev_t, peaks)
    cpu_b = None
    if rank == 0 and world == 1 and not os.environ.get("HI3D_SKIP_CPU_BASELINE"):
        lat = args.cpu_latent
        t = cpu_baseline_sample(args.stage, lat, 1, 1)
        ratio = unet_step_flops(args.stage, wl["h"]) / unet_step_flops(args.stage, lat)
        cpu_b = {"value": T_FRAMES / (NUM_STEPS * t * ratio), "unit": "frames/s", "cores": min(os.cpu_count() or 1, 16), "kind": "port",
                 "sample": f"1 timed Euler step after 1 warm-up (full-width VideoUNet, fp32 oracle port) at {lat}x{lat} latents = {t:.2f} s/step, "
                           f"projected to {wl['h']}x{wl['h']} by UNet FLOP ratio {ratio:.1f} x 25 steps"}
    if rank == 0:
        videos = 1 if frames_mode else world
        fps = T_FRAMES * args.steps * videos / (ms / 1e3)
        fps_e2e = T_FRAMES * args.steps * videos / (ms_e2e / 1e3)
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if frames_mode else "weak",
                "vs_baseline": None, "dtype": "fp16", "data": "synthetic", "config": config, "engine": args.engine,
                "unet_ms_per_sampler_step": unet_ms, "clocks": clk,
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "roofline": roof, "kernel_breakdown": breakdown, "cpu_baseline": cpu_b}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
