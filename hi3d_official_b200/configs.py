"""The two inference configurations of the reference (configs/inference-v01.yaml / inference-v02.yaml) as plain
dicts with the reference's own `target:` strings, restricted to the keys the inference path reads.  The real
YAML files load through `engine.create_model` unchanged; these dicts exist because the reference tree does not
travel to the GPU box."""
from __future__ import annotations

import copy

_VAE_DD = dict(attn_type="vanilla-xformers", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3,
               ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)

UNET_STAGE1 = dict(adm_in_channels=768, num_classes="sequential", use_checkpoint=True, in_channels=8, out_channels=4,
                   model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                   num_head_channels=64, use_linear_in_transformer=True, transformer_depth=1, context_dim=1024,
                   spatial_transformer_attn_type="softmax-xformers", extra_ff_mix_layer=True,
                   use_spatial_context=True, merge_strategy="learned_with_images", video_kernel_size=[3, 1, 1])
UNET_STAGE2 = dict(UNET_STAGE1, adm_in_channels=512, in_channels=17)


def _model(stage: int) -> dict:
    unet = UNET_STAGE1 if stage == 1 else UNET_STAGE2
    return {
        "target": "vtdm.vtdm_gen_v01.VideoLDM" if stage == 1 else "vtdm.vtdm_gen_stage2_degradeImage.VideoLDM",
        "params": {
            "input_key": "video", "scale_factor": 0.18215, "log_keys": "caption", "num_samples": 16,
            "en_and_decode_n_samples_a_time": 16 if stage == 1 else 1,
            "disable_first_stage_autocast": True,
            "denoiser_config": {"target": "sgm.modules.diffusionmodules.denoiser.Denoiser", "params": {
                "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise"}}},
            "network_config": {"target": "sgm.modules.diffusionmodules.video_model.VideoUNet",
                               "params": copy.deepcopy(unet)},
            "conditioner_config": {"target": "sgm.modules.GeneralConditioner", "params": {"emb_models": []}},
            "first_stage_config": {"target": "sgm.models.autoencoder.AutoencoderKL", "params": {
                "embed_dim": 4, "monitor": "val/rec_loss", "ddconfig": copy.deepcopy(_VAE_DD),
                "lossconfig": {"target": "torch.nn.Identity"}}},
            "sampler_config": {"target": "sgm.modules.diffusionmodules.sampling.EulerEDMSampler", "params": {
                "num_steps": 25, "verbose": False,
                "discretization_config": {"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization",
                                          "params": {"sigma_max": 700.0}},
                "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.LinearPredictionGuider",
                                  "params": {"num_frames": 16, "max_scale": 2.5 if stage == 1 else 2.0,
                                             "min_scale": 1.0}}}},
        },
    }


def stage1_config() -> dict:
    return {"model": _model(1)}


def stage2_config() -> dict:
    return {"model": _model(2)}


def build_engine(stage: int = 1, device="cuda", unet_overrides=None, vae_overrides=None, num_steps=None,
                 num_frames=None):
    """Instantiate the engine from the dict config directly on `device` (fp16), without weights."""
    import torch
    from .util import get_obj_from_str
    cfg = _model(stage)
    p = cfg["params"]
    if unet_overrides:
        p["network_config"]["params"].update(unet_overrides)
    if vae_overrides:
        p["first_stage_config"]["params"]["ddconfig"].update(vae_overrides)
    if num_steps is not None:
        p["sampler_config"]["params"]["num_steps"] = num_steps
    if num_frames is not None:
        p["num_samples"] = num_frames
        p["sampler_config"]["params"]["guider_config"]["params"]["num_frames"] = num_frames
    with torch.device(device):
        model = get_obj_from_str(cfg["target"])(**p)
    return model.half()
