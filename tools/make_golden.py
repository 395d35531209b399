#!/usr/bin/env python
"""Generates tests/golden/*.pt by running the UNMODIFIED reference modules (imported read-only from
/root/reference via oracle/ref_import.py) on seeded synthetic weights + inputs, CPU fp32.  The reference tree
does not exist on the GPU box, so these small fixtures are what pins parity there.  Re-run only here:

    python tools/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import as R  # noqa: E402
from hi3d_official_b200 import spec  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
UNET_KW = dict(R.UNET_S1, model_channels=64)
UNET2_KW = dict(R.UNET_S2, model_channels=64)
VAE_DD = dict(R.VAE_DD, ch=64)


def cond(T, cc, adm, hw, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, 4, hw, hw, generator=g)
    c = dict(crossattn=torch.randn(1, 1, 1024, generator=g), vector=torch.randn(1, adm, generator=g),
             concat=torch.randn(T, cc, hw, hw, generator=g) * 0.18)
    uc = dict(crossattn=torch.zeros(1, 1, 1024), vector=c["vector"].clone(), concat=torch.zeros(T, cc, hw, hw))
    return x, c, uc


@torch.no_grad()
def video_decoder_fixture():
    """the unmodified temporal_ae.VideoDecoder (time_mode 'conv-only', video_kernel_size [3, 1, 1]) run as
    Decoder.forward(z, timesteps=T) on seeded synthetic weights, 2 clips x 4 frames, 16x16 latents."""
    R.setup()
    from sgm.modules.autoencoding.temporal_ae import VideoDecoder
    dd = dict(VAE_DD, attn_type="vanilla")
    cfg = spec.VAEConfig.from_ddconfig(VAE_DD, 4)
    sd = spec.synth_state_dict(spec.video_decoder_param_shapes(cfg, (3, 1, 1)), seed=2)
    ref = VideoDecoder(**dd, video_kernel_size=[3, 1, 1], time_mode="conv-only").eval()
    ref.load_state_dict({k[len("decoder."):]: v for k, v in sd.items()}, strict=True)
    T = 4
    g = torch.Generator().manual_seed(8)
    z = torch.randn(2 * T, 4, 16, 16, generator=g)
    out = ref(z, timesteps=T)
    fix = dict(ddconfig=VAE_DD, seed=2, T=T, z=z, dec=out, video_kernel_size=[3, 1, 1])
    torch.save(fix, os.path.join(OUT, "vae_video_ch64.pt"))
    print("video decoder", tuple(out.shape), float(out.abs().mean()))


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    if "--only-video" in sys.argv:
        video_decoder_fixture()
        return
    torch.manual_seed(0)
    # ---- stage-1 style UNet (8 input channels), T=4, 16x16 latents: denoiser outputs at 3 sigmas + 3-step sampler
    for tag, kw, cc, adm, scale in (("s1", UNET_KW, 4, 768, 2.5), ("s2", UNET2_KW, 13, 512, 2.0)):
        T, hw = 4, 16
        ref = R.build_unet(**kw)
        sd = spec.synth_state_dict(spec.unet_param_shapes(spec.UNetConfig.from_kwargs(**kw)), seed=1)
        ref.load_state_dict(sd, strict=True)
        net, den = R.wrap(ref), R.build_denoiser()
        x, c, uc = cond(T, cc, adm, hw, seed=11)
        kwm = dict(image_only_indicator=torch.zeros(2, T), num_video_frames=T)
        smp = R.build_sampler(num_steps=3, max_scale=scale, num_frames=T)
        fix = dict(unet_kwargs=kw, seed=1, T=T, hw=hw, x=x, c=c, uc=uc, max_scale=scale, denoised={}, euler={})
        for sigma in (700.0, 10.0, 0.5):
            xs = x * (1 + sigma ** 2) ** 0.5
            s = torch.full((T,), sigma)
            d = smp.denoise(xs, lambda i, sg, cc_: den(net, i, sg, cc_, **kwm), s, c, uc)     # CFG-combined D(x, sigma)
            fix["denoised"][sigma] = d
            fix["euler"][sigma] = smp.sampler_step(s, s * 0.7, lambda i, sg, cc_: den(net, i, sg, cc_, **kwm), xs, c, uc)
        fix["sampled3"] = smp(lambda i, sg, cc_: den(net, i, sg, cc_, **kwm), x.clone(), cond=c, uc=uc)
        torch.save(fix, os.path.join(OUT, f"unet_{tag}_mc64.pt"))
        print(tag, {k: float(v.abs().mean()) for k, v in fix["denoised"].items()}, float(fix["sampled3"].abs().mean()))
    # ---- VAE ch=64: mode-encode, sampled encode (CPU RNG, as the reference draws it), decode
    ae = R.build_vae(sample=True, **{k: v for k, v in VAE_DD.items()})
    sdv = spec.synth_state_dict(spec.vae_param_shapes(spec.VAEConfig.from_ddconfig(VAE_DD, 4)), seed=2)
    ae.load_state_dict(sdv, strict=True)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    z_in = torch.randn(2, 4, 16, 16, generator=g)
    torch.manual_seed(77)
    z_sampled = ae.encode(img)
    torch.manual_seed(77)
    noise = torch.randn(2, 4, 16, 16)
    ae.regularization.sample = False
    fix = dict(ddconfig=VAE_DD, seed=2, img=img, z_in=z_in, noise=noise, z_mode=ae.encode(img), z_sampled=z_sampled,
               dec=ae.decode(z_in))
    torch.save(fix, os.path.join(OUT, "vae_ch64.pt"))
    print("vae", float(fix["z_mode"].abs().mean()), float(fix["dec"].abs().mean()))
    video_decoder_fixture()


if __name__ == "__main__":
    main()
