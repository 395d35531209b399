"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (Hi3D `sgm`) from
/root/reference on the build container so that (a) the oracle restatement in
`oracle/hi3d_oracle.py` can be validated against it and (b) golden fixtures under
`tests/golden/` can be generated (see `tools/make_golden.py`).

On the GPU box /root/reference does not exist; there the byte-for-byte copy of the needed modules staged by
`oracle/build_ref.py` under `oracle/_ref/` (git-ignored, travels with the snapshot) is used instead -- only by
`bench.py`'s reference arm / cpu_baseline, never by the product and never by a `-m gpu` parity test.
Stubs follow SURVEY.md App. D: pytorch_lightning / omegaconf / kornia / open_clip are only
needed at module-import time of files we never execute.
"""
import os
import sys
import types

import torch.nn as nn

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("HI3D_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "sgm")) and os.path.isdir(os.path.join(_STAGED, "sgm")):
    REF_ROOT = _STAGED


def is_staged_copy() -> bool:
    return REF_ROOT == _STAGED


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "sgm"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


_done = False


def setup():
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _stub("pytorch_lightning", LightningModule=nn.Module)
    _stub("omegaconf", ListConfig=list, OmegaConf=dict)
    _stub("kornia")
    _stub("open_clip")
    # make sure OUR drop-in aliases are not shadowing the real reference
    for k in [k for k in sys.modules if k == "sgm" or k.startswith("sgm.")]:
        if not getattr(sys.modules[k], "__file__", "").startswith(REF_ROOT):
            del sys.modules[k]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _done = True


UNET_S1 = dict(
    adm_in_channels=768, num_classes="sequential", use_checkpoint=False, in_channels=8, out_channels=4,
    model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
    num_head_channels=64, use_linear_in_transformer=True, transformer_depth=1, context_dim=1024,
    spatial_transformer_attn_type="softmax", extra_ff_mix_layer=True, use_spatial_context=True,
    merge_strategy="learned_with_images", video_kernel_size=[3, 1, 1],
)
UNET_S2 = dict(UNET_S1, adm_in_channels=512, in_channels=17)
VAE_DD = dict(attn_type="vanilla", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
              ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def build_unet(**overrides):
    """Reference VideoUNet (sgm/modules/diffusionmodules/video_model.py:84) with SDPA attention."""
    setup()
    from sgm.modules.diffusionmodules.video_model import VideoUNet
    kw = dict(UNET_S1)
    kw.update(overrides)
    kw["use_checkpoint"] = False
    if kw.get("spatial_transformer_attn_type") == "softmax-xformers":
        kw["spatial_transformer_attn_type"] = "softmax"
    return VideoUNet(**kw).eval()


def build_vae(sample=True, **dd_overrides):
    """Reference AutoencoderKL (sgm/models/autoencoder.py:508) with vanilla (SDPA) attention."""
    setup()
    from sgm.models.autoencoder import AutoencoderKL
    dd = dict(VAE_DD)
    dd.update(dd_overrides)
    ae = AutoencoderKL(embed_dim=4, monitor="val/rec_loss", lossconfig={"target": "torch.nn.Identity"}, ddconfig=dd)
    ae.regularization.sample = sample
    return ae.eval()


def build_sampler(num_steps=25, max_scale=2.5, num_frames=16, sigma_max=700.0, device="cpu"):
    setup()
    from sgm.modules.diffusionmodules.sampling import EulerEDMSampler
    return EulerEDMSampler(
        num_steps=num_steps, device=device, verbose=False,
        discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization",
                               "params": {"sigma_max": sigma_max}},
        guider_config={"target": "sgm.modules.diffusionmodules.guiders.LinearPredictionGuider",
                       "params": {"num_frames": num_frames, "max_scale": max_scale, "min_scale": 1.0}})


def build_denoiser():
    setup()
    from sgm.modules.diffusionmodules.denoiser import Denoiser
    return Denoiser({"target": "sgm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise"})


def wrap(unet):
    setup()
    from sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    return OpenAIWrapper(unet)
