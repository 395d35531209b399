#!/usr/bin/env python
"""Per-kernel microbenchmarks on representative Hi3D shapes (stage-1: N=32 CFG x frames, 64^2 latents;
stage-2: 128^2).  CUDA-event timing, L2 flushed between iterations by cycling through enough distinct buffers
or an explicit flush write.  Also the target for `ncu` captures (one launch per shape with --once).

    python tools/microbench.py [--engine mma|tc5|both] [--stage 1|2] [--once] [--only gemm|attn|norm]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hi3d_official_b200 import ops, pack  # noqa: E402

DEV = "cuda"
H = torch.float16
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def timeit(fn, iters=10, warm=3, once=False):
    if once:
        fn(); torch.cuda.synchronize(); return float("nan")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=DEV) * scale).to(H)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="both")
    ap.add_argument("--stage", type=int, default=1)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    engines = ["mma", "tc5"] if args.engine == "both" else [args.engine]
    hw = 64 if args.stage == 1 else 128
    N, T = 32, 16
    res = []

    def report(name, ms, flops=0.0, bytes_=0.0):
        d = dict(name=name, ms=round(ms, 4))
        if flops:
            d["tflops"] = round(flops / ms / 1e9, 1)
            d["frac_tensor_burst"] = round(flops / ms / 1e9 / PEAKS["bf16_tflops"], 3)
        if bytes_:
            d["gbs"] = round(bytes_ / ms / 1e6, 1)
            d["frac_hbm"] = round(bytes_ / ms / 1e6 / PEAKS["hbm_gbs"], 3)
        res.append(d)
        print(json.dumps(d), flush=True)

    if args.only in ("", "gemm"):
        for ds, C in ((1, 320), (2, 640), (4, 1280)):
            h = hw // ds
            M = N * h * h
            x = rnd(N, h, h, C)
            out = torch.empty(M, C, dtype=H, device=DEV)
            emb = rnd(N, C)
            w3 = pack.pack_conv2d(torch.randn(C, C, 3, 3, device=DEV) * (9 * C) ** -0.5)
            wt = pack.pack_conv3d_t(torch.randn(C, C, 3, 1, 1, device=DEV) * (3 * C) ** -0.5)
            w1, b1 = pack.pack_geglu(torch.randn(8 * C, C, device=DEV) * C ** -0.5, torch.zeros(8 * C, device=DEV))
            w2 = rnd(C, 4 * C, scale=(4 * C) ** -0.5)
            wqkv = rnd(3 * C, C, scale=C ** -0.5)
            bias = torch.zeros(C, device=DEV)
            ffh = torch.empty(M, 4 * C, dtype=H, device=DEV)
            qkv = torch.empty(M, 3 * C, dtype=H, device=DEV)
            x2 = x.view(M, C)
            for e in engines:
                g = ops.Gemm(ops.conv_taps([x]), w3, out, M, mode=ops.ROWS_CONV2D, geom=dict(Ho=h, Wo=h, Hs=h, Ws=h), bias=bias,
                             rowbias=emb, rb_div=h * h, rb_mod=N, engine=e)
                report(f"{e} conv3x3 C={C} {h}x{h} (M={M})", timeit(g, once=args.once), g.flops)
                g = ops.Gemm(ops.temporal_taps(x2), wt, out, M, mode=ops.ROWS_TEMPORAL, geom=dict(Ho=h * h, Wo=1, T=T), bias=bias,
                             residual=x2, blend_x=x2, alpha=0.5, engine=e)
                report(f"{e} tconv(3,1,1) C={C} {h}x{h}", timeit(g, once=args.once), g.flops)
                g = ops.Gemm([ops.SegSpec(x2)], w1, ffh, M, bias=b1, act=ops.ACT_GEGLU, engine=e)
                report(f"{e} geglu C->8C C={C} M={M}", timeit(g, once=args.once), g.flops)
                g = ops.Gemm([ops.SegSpec(ffh)], w2, out, M, bias=bias, residual=x2, engine=e)
                report(f"{e} ff2 4C->C C={C} M={M}", timeit(g, once=args.once), g.flops)
                g = ops.Gemm([ops.SegSpec(x2)], wqkv, qkv, M, engine=e)
                report(f"{e} qkv C->3C C={C} M={M}", timeit(g, once=args.once), g.flops)
                wo = rnd(C, C, scale=C ** -0.5)
                out2 = torch.empty(M, C, dtype=H, device=DEV)
                g = ops.Gemm([ops.SegSpec(x2)], wo, out2, M, bias=bias, residual=out, engine=e)
                report(f"{e} proj C->C +res C={C} M={M}", timeit(g, once=args.once), g.flops, 6.0 * M * C)
    if args.only in ("", "attn"):
        for ds, C in ((1, 320), (2, 640), (4, 1280)):
            h = hw // ds
            L, M, heads = h * h, N * h * h, C // 64
            qkv = rnd(M, 3 * C)
            out = torch.empty(M, C, dtype=H, device=DEV)
            for e in engines:
                from hi3d_official_b200 import _native
                for var in ((0, 1, 2, 3, 4, 6) if e == "tc5" else (0,)):   # shared rows / split / split lean / lean + any-order / lean + ping-pong
                    for emu in (((0, 6, 1, 5, 2, 7, 3, 4) if var == 2 else ((0, 1, 5, 2) if var == 6 else (0, 1, 2))) if e == "tc5" else (0,)):   # quarters of the exps on the FMA pipe
                        _native.load().hi3d_attention_tc5_set_variant(var)
                        _native.load().hi3d_attention_tc5_set_exp_emulation(emu)
                        report(f"{e} spatial attention L={L} heads={heads} variant={var} emu={emu}/4",
                               timeit(lambda: ops.attention_d64(qkv, N, L, heads, out, engine=e), once=args.once),
                               4.0 * N * L * L * C)
                _native.load().hi3d_attention_tc5_set_variant(2)
                _native.load().hi3d_attention_tc5_set_exp_emulation(1)
            report(f"temporal attention S={L} heads={heads}",
                   timeit(lambda: ops.temporal_attention_d64(qkv, 2, T, L, heads, out), once=args.once), 4.0 * N * L * T * C,
                   8.0 * M * C)
    if args.only in ("", "vae", "attn"):
        for L in ((hw * 8 // 8) ** 2 // 1,):          # (H/8)^2 of the stage's image size: 4096 (512^2) / 16384 (1024^2)
            for n_img in (1, 4):
                qkv = rnd(n_img * L, 1536)
                out = torch.empty(n_img * L, 512, dtype=H, device=DEV)
                report(f"VAE attention d=512 flash L={L} n={n_img}",
                       timeit(lambda: ops.attention_d512(qkv, n_img, L, out), once=args.once), 4.0 * n_img * L * L * 512)
    if args.only in ("", "norm"):
        for ds, C in ((1, 320), (2, 640), (4, 1280), (1, 640)):
            h = hw // ds
            M = N * h * h
            x = rnd(M, C)
            y = torch.empty_like(x)
            g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
            ws = ops.groupnorm_ws(N, DEV)
            report(f"groupnorm+silu C={C} {h}x{h} (spatial)", timeit(lambda: ops.groupnorm_silu(x, None, N, h * h, g, b, 1e-5, True, y, ws),
                                                                   once=args.once), 0, 6.0 * M * C)
            report(f"groupnorm+silu C={C} {h}x{h} (temporal)",
                   timeit(lambda: ops.groupnorm_silu(x, None, 2, T * h * h, g, b, 1e-5, True, y, ws), once=args.once), 0, 6.0 * M * C)
            report(f"layernorm C={C} M={M}", timeit(lambda: ops.layernorm(x, g, b, y, M), once=args.once), 0, 4.0 * M * C)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"microbench_stage{args.stage}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
