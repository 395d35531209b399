"""Pins the oracle restatement (oracle/hi3d_oracle.py) and the param specs (hi3d_official_b200/spec.py)
against the UNMODIFIED reference modules.  Runs only where /root/reference exists (the build container);
on the GPU box the committed fixtures in tests/golden/ play this role (tests/test_golden.py)."""
import pytest
import torch

from oracle import hi3d_oracle as O
from oracle import ref_import as R
from hi3d_official_b200 import spec

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")

SMALL = dict(model_channels=64, channel_mult=[1, 2, 4, 4], adm_in_channels=768)


def _inputs(cin_cat=4, adm=768, hw=16, T=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, 4, hw, hw, generator=g)
    c = dict(crossattn=torch.randn(1, 1, 1024, generator=g), vector=torch.randn(1, adm, generator=g),
             concat=torch.randn(T, cin_cat, hw, hw, generator=g) * 0.18)
    uc = dict(crossattn=torch.zeros(1, 1, 1024), vector=c["vector"].clone(), concat=torch.zeros(T, cin_cat, hw, hw))
    return x, c, uc


@pytest.fixture(scope="module")
def small_unet():
    torch.manual_seed(0)
    ref = R.build_unet(**SMALL)
    cfg = spec.UNetConfig.from_kwargs(**dict(R.UNET_S1, **SMALL))
    shapes = spec.unet_param_shapes(cfg)
    ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert dict(shapes) == ref_shapes
    sd = spec.synth_state_dict(shapes, seed=1)
    ref.load_state_dict(sd, strict=True)
    return ref, sd


def test_unet_param_spec_full_size_matches_reference_on_meta():
    R.setup()
    from sgm.modules.diffusionmodules.video_model import VideoUNet
    for kw in (R.UNET_S1, R.UNET_S2):
        with torch.device("meta"):
            ref = VideoUNet(**kw)
        ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        mine = spec.unet_param_shapes(spec.UNetConfig.from_kwargs(**kw))
        assert dict(mine) == ref_shapes
        assert sum(torch.Size(s).numel() for s in mine.values()) in (1524623082, 1524321322)


def test_unet_forward_matches_reference(small_unet):
    ref, sd = small_unet
    T = 4
    x, c, uc = _inputs(T=T)
    xin = torch.cat([torch.cat([x, x]), torch.cat([uc["concat"], c["concat"]])], 1)
    t = torch.full((2 * T,), 0.7)
    ctx = torch.cat([uc["crossattn"], c["crossattn"]])
    y = torch.cat([uc["vector"], c["vector"]])
    with torch.no_grad():
        a = ref(xin, timesteps=t, context=ctx, y=y, num_video_frames=T, image_only_indicator=torch.zeros(2, T))
        b = O.unet_forward(sd, xin, t, ctx, y, num_video_frames=T)
    assert a.abs().mean() > 1e-2
    torch.testing.assert_close(b, a, rtol=1e-4, atol=2e-4)


def test_sampler_matches_reference(small_unet):
    ref, sd = small_unet
    T, steps = 4, 3
    x, c, uc = _inputs(T=T, seed=3)
    smp = R.build_sampler(num_steps=steps, max_scale=2.5, num_frames=T)
    den = R.build_denoiser()
    net = R.wrap(ref)
    kw = dict(image_only_indicator=torch.zeros(2, T), num_video_frames=T)
    with torch.no_grad():
        a = smp(lambda inp, s, cc: den(net, inp, s, cc, **kw), x.clone(), cond=c, uc=uc)
        b = O.sample(sd, x.clone(), c, uc, num_steps=steps, max_scale=2.5, num_frames=T)
    torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-3)


def test_sampler_constants():
    s = O.edm_sigmas(25)
    ref = R.build_sampler().discretization(25, device="cpu")
    torch.testing.assert_close(s, ref, rtol=0, atol=0)
    assert abs(float(s[0]) - 700.0001) < 1e-3 and float(s[-1]) == 0.0 and abs(float(s[-2]) - 0.002) < 1e-6
    cs = O.vscaling_edm_cnoise(torch.tensor(700.0))
    assert abs(float(cs[3]) - 1.6377701) < 1e-6 and abs(float(cs[2]) - 1.4285699e-03) < 1e-9
    assert abs(O.v02_alpha(1) - 0.984126) < 1e-6 and O.v02_alpha(0) == 1.0


def test_single_key_cross_attention_is_constant(small_unet):
    """attn2 with one context token == to_out(to_v(ctx)) for every query."""
    _, sd = small_unet
    pre = "input_blocks.1.1.transformer_blocks.0.attn2."
    x = torch.randn(3, 10, 64)
    ctx = torch.randn(3, 1, 1024)
    full = O.cross_attention(sd, pre, x, ctx, 1)
    const = torch.nn.functional.linear(torch.nn.functional.linear(ctx, sd[pre + "to_v.weight"]),
                                       sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])
    torch.testing.assert_close(full, const.expand_as(full), rtol=1e-5, atol=1e-6)


def test_vae_matches_reference():
    torch.manual_seed(0)
    ref = R.build_vae(sample=False, ch=32, ch_mult=[1, 2, 4, 4])
    cfg = spec.VAEConfig.from_ddconfig(dict(R.VAE_DD, ch=32), 4)
    shapes = spec.vae_param_shapes(cfg)
    ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert dict(shapes) == ref_shapes
    sd = spec.synth_state_dict(shapes, seed=2)
    ref.load_state_dict(sd, strict=True)
    img = torch.rand(2, 3, 64, 64) * 2 - 1
    with torch.no_grad():
        za = ref.encode(img)
        zb = O.vae_encode(sd, img, scale_factor=1.0)
        torch.testing.assert_close(zb, za, rtol=1e-4, atol=1e-4)
        z = torch.randn(2, 4, 8, 8)
        torch.testing.assert_close(O.vae_decode(sd, z, scale_factor=1.0), ref.decode(z), rtol=1e-4, atol=2e-4)
        # sampled posterior: the reference draws CPU randn (distributions.py:37-41)
        ref.regularization.sample = True
        torch.manual_seed(7)
        zs = ref.encode(img)
        torch.manual_seed(7)
        noise = torch.randn(2, 4, 8, 8)
        torch.testing.assert_close(O.vae_encode(sd, img, noise=noise, scale_factor=1.0), zs, rtol=1e-4, atol=1e-4)


def test_vae_full_size_spec_on_meta():
    R.setup()
    from sgm.modules.diffusionmodules.model import Decoder, Encoder
    with torch.device("meta"):
        e, d = Encoder(**R.VAE_DD), Decoder(**R.VAE_DD)
    ref = {"encoder." + k: tuple(v.shape) for k, v in e.state_dict().items()}
    ref.update({"decoder." + k: tuple(v.shape) for k, v in d.state_dict().items()})
    mine = {k: v for k, v in spec.vae_param_shapes(spec.VAEConfig.from_ddconfig(R.VAE_DD, 4)).items()
            if not k.startswith(("quant_conv", "post_quant_conv"))}
    assert mine == ref


@pytest.mark.parametrize("vks", [[3, 1, 1], 3])
def test_video_decoder_oracle_matches_reference(vks):
    """the oracle restatement of temporal_ae.VideoDecoder (time_mode 'conv-only'; kernel (3,1,1) as in
    SVD and the full 3x3x3 default) against the unmodified reference class, seeded synthetic weights."""
    R.setup()
    from sgm.modules.autoencoding.temporal_ae import VideoDecoder
    torch.manual_seed(0)
    dd = dict(R.VAE_DD, ch=32, ch_mult=[1, 2, 4, 4], attn_type="vanilla")
    ref = VideoDecoder(**dd, video_kernel_size=vks, time_mode="conv-only").eval()
    g = torch.Generator().manual_seed(11)
    sd = {}
    for k, v in ref.state_dict().items():
        if k.endswith("mix_factor"):
            sd[k] = torch.full_like(v, 0.3)
        elif v.ndim == 1 and ("norm" in k or "in_layers.0" in k or "out_layers.0" in k) and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif v.ndim == 1:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:                       # incl. the zero_module'd out_layers conv of every time_stack (F8)
            fan_in = v[0].numel()
            sd[k] = torch.randn(v.shape, generator=g) * fan_in ** -0.5
    ref.load_state_dict(sd, strict=True)
    T = 3
    z = torch.randn(2 * T, 4, 8, 8, generator=g)
    with torch.no_grad():
        a = ref(z, timesteps=T)
        b = O.vae_video_decoder({"decoder." + k: v for k, v in sd.items()}, z, T)
    assert a.shape == (2 * T, 3, 64, 64)
    torch.testing.assert_close(b, a, rtol=1e-4, atol=2e-4)
    # the temporal branch matters: shuffling the frames changes more than a permutation of the output
    with torch.no_grad():
        a2 = ref(z.flip(0), timesteps=T).flip(0)
    assert (a2 - a).abs().max() > 1e-3


@pytest.mark.parametrize("vks", [[3, 1, 1], 3])
def test_video_decoder_param_spec_matches_reference_on_meta(vks):
    R.setup()
    from sgm.modules.autoencoding.temporal_ae import VideoDecoder
    with torch.device("meta"):
        ref = VideoDecoder(**dict(R.VAE_DD, attn_type="vanilla"), video_kernel_size=vks, time_mode="conv-only")
    ref_shapes = {"decoder." + k: tuple(v.shape) for k, v in ref.state_dict().items()}
    mine = spec.video_decoder_param_shapes(spec.VAEConfig.from_ddconfig(R.VAE_DD, 4), vks)
    assert dict(mine) == ref_shapes
