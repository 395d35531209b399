"""Architecture description of the Hi3D hot-path networks, derived from constructor kwargs.

One place that knows the block topology of `VideoUNet` (reference: sgm/modules/diffusionmodules/
video_model.py:84-440) and of `AutoencoderKL`'s Encoder/Decoder (sgm/modules/diffusionmodules/
model.py:487-748): it yields (i) the reference `state_dict` key names + shapes so
that checkpoints load unchanged, and (ii) a flat block plan that the CUDA executor in `unet.py` /
`vae.py` compiles into kernel launches.  Also holds the seeded synthetic-weight recipe used wherever no
checkpoint exists (there are none offline): deterministic per key, generated on the CPU generator so the
very same tensors can be given to the reference (build container), the oracle and the CUDA path.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Shapes = "OrderedDict[str, Tuple[int, ...]]"


# ------------------------------------------------------------------------------------------------
# VideoUNet
# ------------------------------------------------------------------------------------------------
@dataclass
class UNetConfig:
    in_channels: int = 8
    model_channels: int = 320
    out_channels: int = 4
    num_res_blocks: int = 2
    attention_resolutions: Sequence[int] = (4, 2, 1)
    channel_mult: Sequence[int] = (1, 2, 4, 4)
    num_head_channels: int = 64
    transformer_depth: int = 1
    context_dim: int = 1024
    adm_in_channels: int = 768
    video_kernel_size: Sequence[int] = (3, 1, 1)
    max_ddpm_temb_period: int = 10000
    num_classes: Optional[str] = "sequential"

    @staticmethod
    def from_kwargs(**kw) -> "UNetConfig":
        """Accepts the full ctor kwargs of the reference VideoUNet (video_model.py:85-118); rejects the
        variants the Hi3D configs never use instead of silently computing something else."""
        def need(name, allowed):
            if name in kw and kw[name] not in allowed:
                raise NotImplementedError(f"VideoUNet({name}={kw[name]!r}) is outside the Hi3D hot path; "
                                          f"supported: {allowed}")
        need("dims", (2,)); need("use_scale_shift_norm", (False,)); need("resblock_updown", (False,))
        need("conv_resample", (True,)); need("time_downup", (False,)); need("num_classes", ("sequential",))
        need("use_linear_in_transformer", (True,)); need("extra_ff_mix_layer", (True,))
        need("use_spatial_context", (True,)); need("merge_strategy", ("learned_with_images", "learned"))
        need("disable_temporal_crossattention", (False,)); need("dropout", (0, 0.0)); need("num_heads", (-1,))
        need("spatial_transformer_attn_type", ("softmax", "softmax-xformers"))   # synonyms of the native kernel
        need("time_context_dim", (None,)); need("transformer_depth_middle", (None,))
        td = kw.get("transformer_depth", 1)
        if not isinstance(td, int):
            td = list(td)
            if len(set(td)) != 1:
                raise NotImplementedError("per-level transformer_depth")
            td = td[0]
        vks = kw.get("video_kernel_size", 3)
        vks = [vks] * 3 if isinstance(vks, int) else list(vks)
        if list(vks) != [3, 1, 1]:
            raise NotImplementedError(f"video_kernel_size={vks}; the Hi3D configs use [3, 1, 1]")
        if "context_dim" not in kw or kw["context_dim"] is None:
            raise AssertionError("context_dim is required (video_model.py:120)")
        if kw.get("num_head_channels", -1) == -1:
            raise NotImplementedError("num_heads-style head split; Hi3D uses num_head_channels=64")
        return UNetConfig(
            in_channels=kw["in_channels"], model_channels=kw["model_channels"], out_channels=kw["out_channels"],
            num_res_blocks=kw["num_res_blocks"], attention_resolutions=tuple(kw["attention_resolutions"]),
            channel_mult=tuple(kw.get("channel_mult", (1, 2, 4, 8))), num_head_channels=kw["num_head_channels"],
            transformer_depth=td, context_dim=kw["context_dim"], adm_in_channels=kw["adm_in_channels"],
            video_kernel_size=tuple(vks), max_ddpm_temb_period=kw.get("max_ddpm_temb_period", 10000),
            num_classes=kw.get("num_classes"))


@dataclass
class Layer:
    kind: str            # 'conv_in' | 'res' | 'attn' | 'down' | 'up'
    name: str            # state-dict prefix, e.g. 'input_blocks.1.0.'
    cin: int = 0
    cout: int = 0
    ds: int = 1          # spatial down-sampling factor of the *input* of this layer


@dataclass
class UNetPlan:
    cfg: UNetConfig
    input_blocks: List[List[Layer]] = field(default_factory=list)
    middle: List[Layer] = field(default_factory=list)
    output_blocks: List[List[Layer]] = field(default_factory=list)
    skip_channels: List[int] = field(default_factory=list)   # channels pushed on `hs` by each input block


def unet_plan(cfg: UNetConfig) -> UNetPlan:
    """Block topology, following the constructor walk of video_model.py:186-440."""
    mc = cfg.model_channels
    plan = UNetPlan(cfg)
    plan.input_blocks.append([Layer("conv_in", "input_blocks.0.0.", cfg.in_channels, mc, 1)])
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            i = len(plan.input_blocks)
            layers = [Layer("res", f"input_blocks.{i}.0.", ch, mult * mc, ds)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(Layer("attn", f"input_blocks.{i}.1.", ch, ch, ds))
            plan.input_blocks.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            i = len(plan.input_blocks)
            plan.input_blocks.append([Layer("down", f"input_blocks.{i}.0.", ch, ch, ds)])
            ds *= 2
            chans.append(ch)
    plan.skip_channels = list(chans)
    plan.middle = [Layer("res", "middle_block.0.", ch, ch, ds), Layer("attn", "middle_block.1.", ch, ch, ds),
                   Layer("res", "middle_block.2.", ch, ch, ds)]
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            j = len(plan.output_blocks)
            layers = [Layer("res", f"output_blocks.{j}.0.", ch + ich, mc * mult, ds)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(Layer("attn", f"output_blocks.{j}.{len(layers)}.", ch, ch, ds))
            if level and i == cfg.num_res_blocks:
                layers.append(Layer("up", f"output_blocks.{j}.{len(layers)}.", ch, ch, ds))
                ds //= 2
            plan.output_blocks.append(layers)
    return plan


def _res_shapes(s, pre, cin, cout, emb):
    def rb(q, ci, co, k):
        s[q + "in_layers.0.weight"] = (ci,); s[q + "in_layers.0.bias"] = (ci,)
        s[q + "in_layers.2.weight"] = (co, ci) + k; s[q + "in_layers.2.bias"] = (co,)
        s[q + "emb_layers.1.weight"] = (co, emb); s[q + "emb_layers.1.bias"] = (co,)
        s[q + "out_layers.0.weight"] = (co,); s[q + "out_layers.0.bias"] = (co,)
        s[q + "out_layers.3.weight"] = (co, co) + k; s[q + "out_layers.3.bias"] = (co,)
        if ci != co:
            s[q + "skip_connection.weight"] = (co, ci) + (1,) * len(k); s[q + "skip_connection.bias"] = (co,)
    rb(pre, cin, cout, (3, 3))
    rb(pre + "time_stack.", cout, cout, (3, 1, 1))
    s[pre + "time_mixer.mix_factor"] = (1,)


def _attn_shapes(s, pre, c, ctx, depth):
    def xattn(q, qd, cd):
        s[q + "to_q.weight"] = (qd, qd); s[q + "to_k.weight"] = (qd, cd); s[q + "to_v.weight"] = (qd, cd)
        s[q + "to_out.0.weight"] = (qd, qd); s[q + "to_out.0.bias"] = (qd,)

    def ff(q):
        s[q + "net.0.proj.weight"] = (8 * c, c); s[q + "net.0.proj.bias"] = (8 * c,)
        s[q + "net.2.weight"] = (c, 4 * c); s[q + "net.2.bias"] = (c,)

    def ln(q):
        s[q + ".weight"] = (c,); s[q + ".bias"] = (c,)
    ln(pre + "norm")
    s[pre + "proj_in.weight"] = (c, c); s[pre + "proj_in.bias"] = (c,)
    for d in range(depth):
        q = pre + f"transformer_blocks.{d}."
        xattn(q + "attn1.", c, c); ff(q + "ff."); xattn(q + "attn2.", c, ctx)
        ln(q + "norm1"); ln(q + "norm2"); ln(q + "norm3")
    s[pre + "proj_out.weight"] = (c, c); s[pre + "proj_out.bias"] = (c,)
    for d in range(depth):
        q = pre + f"time_stack.{d}."
        ln(q + "norm_in"); ff(q + "ff_in."); xattn(q + "attn1.", c, c); ff(q + "ff.")
        ln(q + "norm2"); xattn(q + "attn2.", c, ctx); ln(q + "norm1"); ln(q + "norm3")
    s[pre + "time_pos_embed.0.weight"] = (4 * c, c); s[pre + "time_pos_embed.0.bias"] = (4 * c,)
    s[pre + "time_pos_embed.2.weight"] = (c, 4 * c); s[pre + "time_pos_embed.2.bias"] = (c,)
    s[pre + "time_mixer.mix_factor"] = (1,)


def unet_param_shapes(cfg: UNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Reference state_dict keys/shapes (order is not significant)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    mc, emb = cfg.model_channels, cfg.model_channels * 4
    s["time_embed.0.weight"] = (emb, mc); s["time_embed.0.bias"] = (emb,)
    s["time_embed.2.weight"] = (emb, emb); s["time_embed.2.bias"] = (emb,)
    s["label_emb.0.0.weight"] = (emb, cfg.adm_in_channels); s["label_emb.0.0.bias"] = (emb,)
    s["label_emb.0.2.weight"] = (emb, emb); s["label_emb.0.2.bias"] = (emb,)
    plan = unet_plan(cfg)
    for blk in plan.input_blocks + [plan.middle] + plan.output_blocks:
        for L in blk:
            if L.kind == "conv_in":
                s[L.name + "weight"] = (L.cout, L.cin, 3, 3); s[L.name + "bias"] = (L.cout,)
            elif L.kind == "res":
                _res_shapes(s, L.name, L.cin, L.cout, emb)
            elif L.kind == "attn":
                _attn_shapes(s, L.name, L.cin, cfg.context_dim, cfg.transformer_depth)
            elif L.kind == "down":
                s[L.name + "op.weight"] = (L.cout, L.cin, 3, 3); s[L.name + "op.bias"] = (L.cout,)
            elif L.kind == "up":
                s[L.name + "conv.weight"] = (L.cout, L.cin, 3, 3); s[L.name + "conv.bias"] = (L.cout,)
    s["out.0.weight"] = (mc,); s["out.0.bias"] = (mc,)
    s["out.2.weight"] = (cfg.out_channels, mc, 3, 3); s["out.2.bias"] = (cfg.out_channels,)
    return s


# ------------------------------------------------------------------------------------------------
# AutoencoderKL
# ------------------------------------------------------------------------------------------------
@dataclass
class VAEConfig:
    ch: int = 128
    ch_mult: Sequence[int] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    in_channels: int = 3
    out_ch: int = 3
    z_channels: int = 4
    embed_dim: int = 4
    double_z: bool = True

    @staticmethod
    def from_ddconfig(ddconfig: dict, embed_dim: int = 4) -> "VAEConfig":
        if list(ddconfig.get("attn_resolutions", [])):
            raise NotImplementedError("attn_resolutions != [] is outside the Hi3D hot path")
        if ddconfig.get("attn_type", "vanilla") not in ("vanilla", "vanilla-xformers"):
            raise NotImplementedError(f"attn_type={ddconfig['attn_type']}")
        if float(ddconfig.get("dropout", 0.0)) != 0.0:
            raise NotImplementedError("dropout")
        return VAEConfig(ch=ddconfig["ch"], ch_mult=tuple(ddconfig["ch_mult"]),
                         num_res_blocks=ddconfig["num_res_blocks"], in_channels=ddconfig["in_channels"],
                         out_ch=ddconfig["out_ch"], z_channels=ddconfig["z_channels"], embed_dim=embed_dim,
                         double_z=bool(ddconfig.get("double_z", True)))


def vae_param_shapes(cfg: VAEConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Encoder (model.py:487-575), Decoder (model.py:604-713), quant convs (autoencoder.py:453-458)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def rb(q, ci, co):
        s[q + "norm1.weight"] = (ci,); s[q + "norm1.bias"] = (ci,)
        s[q + "conv1.weight"] = (co, ci, 3, 3); s[q + "conv1.bias"] = (co,)
        s[q + "norm2.weight"] = (co,); s[q + "norm2.bias"] = (co,)
        s[q + "conv2.weight"] = (co, co, 3, 3); s[q + "conv2.bias"] = (co,)
        if ci != co:
            s[q + "nin_shortcut.weight"] = (co, ci, 1, 1); s[q + "nin_shortcut.bias"] = (co,)

    def attn(q, c):
        s[q + "norm.weight"] = (c,); s[q + "norm.bias"] = (c,)
        for n in ("q", "k", "v", "proj_out"):
            s[q + n + ".weight"] = (c, c, 1, 1); s[q + n + ".bias"] = (c,)

    def conv(q, ci, co, k=3):
        s[q + ".weight"] = (co, ci, k, k); s[q + ".bias"] = (co,)
    nres = len(cfg.ch_mult)
    # encoder
    conv("encoder.conv_in", cfg.in_channels, cfg.ch)
    in_mult = (1,) + tuple(cfg.ch_mult)
    bi = cfg.ch
    for lvl in range(nres):
        bi, bo = cfg.ch * in_mult[lvl], cfg.ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks):
            rb(f"encoder.down.{lvl}.block.{b}.", bi, bo)
            bi = bo
        if lvl != nres - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", bi, bi)
    rb("encoder.mid.block_1.", bi, bi); attn("encoder.mid.attn_1.", bi); rb("encoder.mid.block_2.", bi, bi)
    s["encoder.norm_out.weight"] = (bi,); s["encoder.norm_out.bias"] = (bi,)
    conv("encoder.conv_out", bi, (2 if cfg.double_z else 1) * cfg.z_channels)
    # decoder
    bi = cfg.ch * cfg.ch_mult[-1]
    conv("decoder.conv_in", cfg.z_channels, bi)
    rb("decoder.mid.block_1.", bi, bi); attn("decoder.mid.attn_1.", bi); rb("decoder.mid.block_2.", bi, bi)
    for lvl in reversed(range(nres)):
        bo = cfg.ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks + 1):
            rb(f"decoder.up.{lvl}.block.{b}.", bi, bo)
            bi = bo
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", bi, bi)
    s["decoder.norm_out.weight"] = (bi,); s["decoder.norm_out.bias"] = (bi,)
    conv("decoder.conv_out", bi, cfg.out_ch)
    nz = (2 if cfg.double_z else 1)
    conv("quant_conv", nz * cfg.z_channels, nz * cfg.embed_dim, 1)
    conv("post_quant_conv", cfg.embed_dim, cfg.z_channels, 1)
    return s


def video_decoder_param_shapes(cfg: VAEConfig, video_kernel_size=(3, 1, 1)) -> "OrderedDict[str, Tuple[int, ...]]":
    """temporal_ae.VideoDecoder, time_mode 'conv-only' (temporal_ae.py:18-108, 293-349): the 2-D
    decoder's keys (prefix `decoder.` kept for symmetry with vae_param_shapes) plus, per ResnetBlock, a `time_stack`
    ResBlock(dims=3, no emb) and a `mix_factor`, and the `time_mix_conv` of the AE3DConv output conv.  The order is the
    reference's registration order (mix_factor parameter after time_stack; buffers n/a for 'learned')."""
    k3 = tuple(video_kernel_size) if not isinstance(video_kernel_size, int) else (video_kernel_size,) * 3
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def rb(q, ci, co):
        s[q + "mix_factor"] = (1,)
        s[q + "norm1.weight"] = (ci,); s[q + "norm1.bias"] = (ci,)
        s[q + "conv1.weight"] = (co, ci, 3, 3); s[q + "conv1.bias"] = (co,)
        s[q + "norm2.weight"] = (co,); s[q + "norm2.bias"] = (co,)
        s[q + "conv2.weight"] = (co, co, 3, 3); s[q + "conv2.bias"] = (co,)
        if ci != co:
            s[q + "nin_shortcut.weight"] = (co, ci, 1, 1); s[q + "nin_shortcut.bias"] = (co,)
        t = q + "time_stack."
        s[t + "in_layers.0.weight"] = (co,); s[t + "in_layers.0.bias"] = (co,)
        s[t + "in_layers.2.weight"] = (co, co) + k3; s[t + "in_layers.2.bias"] = (co,)
        s[t + "out_layers.0.weight"] = (co,); s[t + "out_layers.0.bias"] = (co,)
        s[t + "out_layers.3.weight"] = (co, co) + k3; s[t + "out_layers.3.bias"] = (co,)

    def attn(q, c):
        s[q + "norm.weight"] = (c,); s[q + "norm.bias"] = (c,)
        for n in ("q", "k", "v", "proj_out"):
            s[q + n + ".weight"] = (c, c, 1, 1); s[q + n + ".bias"] = (c,)
    nres = len(cfg.ch_mult)
    bi = cfg.ch * cfg.ch_mult[-1]
    s["decoder.conv_in.weight"] = (bi, cfg.z_channels, 3, 3); s["decoder.conv_in.bias"] = (bi,)
    rb("decoder.mid.block_1.", bi, bi); attn("decoder.mid.attn_1.", bi); rb("decoder.mid.block_2.", bi, bi)
    for lvl in reversed(range(nres)):
        bo = cfg.ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks + 1):
            rb(f"decoder.up.{lvl}.block.{b}.", bi, bo)
            bi = bo
        if lvl != 0:
            s[f"decoder.up.{lvl}.upsample.conv.weight"] = (bi, bi, 3, 3); s[f"decoder.up.{lvl}.upsample.conv.bias"] = (bi,)
    s["decoder.norm_out.weight"] = (bi,); s["decoder.norm_out.bias"] = (bi,)
    s["decoder.conv_out.weight"] = (cfg.out_ch, bi, 3, 3); s["decoder.conv_out.bias"] = (cfg.out_ch,)
    s["decoder.conv_out.time_mix_conv.weight"] = (cfg.out_ch, cfg.out_ch) + k3
    s["decoder.conv_out.time_mix_conv.bias"] = (cfg.out_ch,)
    return s


# ------------------------------------------------------------------------------------------------
# Seeded synthetic weights
# ------------------------------------------------------------------------------------------------
_ZERO_INIT_SUFFIXES = ("out_layers.3.weight", "proj_out.weight", "out.2.weight")


def synth_state_dict(shapes: "Dict[str, Tuple[int, ...]]", seed: int = 0, dtype=torch.float32,
                     device="cpu", gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic stand-in weights (no checkpoints are available offline).

    Per-key CPU generator seeded with sha1(key) ^ seed, so any subset of keys reproduces bit-exactly
    everywhere.  Matrices/convs ~ N(0, 1/fan_in) (half that std for the reference's zero-initialised
    layers, which a default-init model would leave at exactly 0 and make every parity test vacuous);
    norm scales ~ 1 + 0.1 N; biases ~ 0.02 N; mix_factor ~ 0.5 N.
    """
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, shp in shapes.items():
        h = int.from_bytes(hashlib.sha1(k.encode()).digest()[:6], "little") ^ (seed * 0x9E3779B1)
        g = torch.Generator(device="cpu").manual_seed(h & 0x7FFFFFFFFFFF)
        x = torch.randn(tuple(shp), generator=g, dtype=torch.float32)
        out[k] = _synth_rule(k, x, gain).to(device=device, dtype=dtype)
    return out


def _synth_rule(k: str, x: torch.Tensor, gain: float = 1.0) -> torch.Tensor:
    """Maps a standard-normal tensor to the synthetic distribution of parameter `k`."""
    shp = x.shape
    if k.endswith("mix_factor"):
        return x * 0.5
    if len(shp) == 1:
        return 1.0 + 0.1 * x if k.endswith(".weight") else 0.02 * x
    fan_in = 1
    for d in shp[1:]:
        fan_in *= d
    std = gain / (fan_in ** 0.5)
    if k.endswith(_ZERO_INIT_SUFFIXES):
        std *= 0.5
    return x * std


@torch.no_grad()
def synth_fill_(module: torch.nn.Module, seed: int = 0, fast: bool = False) -> torch.nn.Module:
    """Fill every parameter of `module` in place with the synthetic recipe.  fast=False reproduces
    `synth_state_dict` bit-exactly (CPU generator per key, needed wherever results are compared with the
    oracle); fast=True draws on the parameter's own device (benchmarks: same distribution, different bits)."""
    if fast:
        dev_gens = {}
    for k, p in module.state_dict().items():
        # engine-level prefixes are stripped so the same keys give the same tensors as the per-network dicts
        kk = k
        for pre in ("model.diffusion_model.", "first_stage_model."):
            if kk.startswith(pre):
                kk = kk[len(pre):]
        if fast:
            g = dev_gens.get(p.device)
            if g is None:
                g = dev_gens[p.device] = torch.Generator(device=p.device).manual_seed(1234 + seed)
            x = torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)
        else:
            h = int.from_bytes(hashlib.sha1(kk.encode()).digest()[:6], "little") ^ (seed * 0x9E3779B1)
            g = torch.Generator(device="cpu").manual_seed(h & 0x7FFFFFFFFFFF)
            x = torch.randn(tuple(p.shape), generator=g, dtype=torch.float32)
        p.copy_(_synth_rule(kk, x).to(device=p.device, dtype=p.dtype))
    for m in module.modules():          # packed weights / launch plans derived from the old values are stale
        if hasattr(m, "_packed"):
            m._packed, m._plans = None, {}
    return module
