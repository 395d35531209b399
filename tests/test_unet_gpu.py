"""VideoUNet (CUDA launch plan) against the fp32 oracle on the same seeded weights/inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from hi3d_official_b200 import spec  # noqa: E402
from hi3d_official_b200.unet import VideoUNet  # noqa: E402
from oracle import hi3d_oracle as O  # noqa: E402

KW_S1 = dict(adm_in_channels=768, num_classes="sequential", use_checkpoint=True, in_channels=8, out_channels=4,
             model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
             num_head_channels=64, use_linear_in_transformer=True, transformer_depth=1, context_dim=1024,
             spatial_transformer_attn_type="softmax-xformers", extra_ff_mix_layer=True, use_spatial_context=True,
             merge_strategy="learned_with_images", video_kernel_size=[3, 1, 1])


def report(a, b, name):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    tol = 1e-2 + 1e-3 * b.abs()
    frac = float((err > tol).float().mean())
    print(f"[{name}] max|err| {float(err.max()):.3e} mean|err| {float(err.mean()):.3e} ref mean|x| "
          f"{float(b.abs().mean()):.3e} max|x| {float(b.abs().max()):.3e} frac outside(1e-2,1e-3) {frac:.2e}")
    return float(err.max()), frac


def build(kw, seed=1):
    cfg = spec.UNetConfig.from_kwargs(**kw)
    sd = spec.synth_state_dict(spec.unet_param_shapes(cfg), seed=seed)
    net = VideoUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().half()
    return net, {k: v.cuda() for k, v in sd.items()}


def inputs(N, cin, hw, adm, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, cin, hw, hw, generator=g).cuda()
    ctx = torch.randn(N // T, 1, 1024, generator=g).cuda()
    ctx[0] = 0                                              # uc half: zeroed CLIP embedding
    y = torch.randn(N // T, adm, generator=g).cuda()
    t = torch.full((N,), 0.7).cuda()
    return x, ctx, y, t


@pytest.mark.parametrize("engine", ["mma", "tc5"])
@pytest.mark.parametrize("mc,T,hw", [(64, 4, 16), (64, 16, 8), (128, 8, 16), (64, 16, 32)])
def test_small_unet_vs_oracle(mc, T, hw, engine):
    kw = dict(KW_S1, model_channels=mc)
    net, sd = build(kw)
    net.set_engine(engine)
    N = 2 * T
    x, ctx, y, t = inputs(N, 8, hw, 768, T)
    out = net(x, timesteps=t, context=ctx, y=y, num_video_frames=T, image_only_indicator=torch.zeros(2, T).cuda())
    ref = O.unet_forward(sd, x, t, ctx, y, num_video_frames=T)
    assert out.shape == ref.shape and out.dtype == torch.float16
    mx, frac = report(out, ref, f"unet mc={mc} T={T} hw={hw}")
    assert torch.isfinite(out).all()
    assert frac < 1e-3 and mx < 5e-2


@pytest.mark.parametrize("engine", ["mma", "tc5"])
def test_full_width_unet_vs_oracle_small_latents(engine):
    """Stage-1 architecture at full width (1.52 B params), 16x16 latents, T=16 (CFG batch 32)."""
    net, sd = build(KW_S1)
    net.set_engine(engine)
    T, N = 16, 32
    x, ctx, y, t = inputs(N, 8, 16, 768, T, seed=3)
    out = net(x, timesteps=t, context=ctx, y=y, num_video_frames=T)
    ref = O.unet_forward(sd, x, t, ctx, y, num_video_frames=T)
    mx, frac = report(out, ref, "unet full width 16x16")
    assert frac < 1e-3 and mx < 5e-2
