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


def _emb_models(stage: int) -> list:
    """conditioner_config.params.emb_models of configs/inference-v0{1,2}.yaml (:52-112 / :52-114)."""
    clip = {"is_trainable": False, "input_key": "cond_frames_without_noise", "ucg_rate": 0.1,
            "target": "sgm.modules.encoders.modules.FrozenOpenCLIPImagePredictionEmbedder",
            "params": {"n_cond_frames": 1, "n_copies": 1, "open_clip_embedding_config": {
                "target": "sgm.modules.encoders.modules.FrozenOpenCLIPImageEmbedder",
                "params": {"version": "ckpts/open_clip_pytorch_model.bin", "freeze": True}}}}

    def ts(key):
        return {"is_trainable": False, "input_key": key, "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND",
                "params": {"outdim": 256}}
    vae = {"input_key": "cond_frames", "is_trainable": False, "ucg_rate": 0.1,
           "target": "sgm.modules.encoders.modules.VideoPredictionEmbedderWithEncoder",
           "params": {"disable_encoder_autocast": True, "n_cond_frames": 1, "n_copies": 16 if stage == 1 else 1, "is_ae": True,
                      "encoder_config": {"target": "sgm.models.autoencoder.AutoencoderKLModeOnly", "params": {
                          "embed_dim": 4, "monitor": "val/rec_loss", "ddconfig": copy.deepcopy(_VAE_DD),
                          "lossconfig": {"target": "torch.nn.Identity"}}}}}
    if stage == 1:
        aes = {"is_trainable": False, "input_key": "video", "ucg_rate": 0.0, "target": "vtdm.encoders.AesEmbedder"}
        return [clip, aes, ts("elevation"), vae, ts("cond_aug")]
    depth = {"is_trainable": False, "input_key": "cond_frames", "ucg_rate": 0.0, "target": "vtdm.encoders.DepthEmbedder",
             "params": {"shuffle_size": 3}}
    return [clip, ts("elevation"), depth, vae, ts("cond_aug")]


def _model(stage: int, conditioner: bool = False) -> dict:
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
            # benches / parity tests feed ready-made conditioning: no embedders unless asked for (each run would carry a
            # second VAE encoder for the cond-frame latents)
            "conditioner_config": {"target": "sgm.modules.GeneralConditioner",
                                   "params": {"emb_models": _emb_models(stage) if conditioner else []}},
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


def stage1_config(conditioner: bool = False) -> dict:
    return {"model": _model(1, conditioner)}


def stage2_config(conditioner: bool = False) -> dict:
    return {"model": _model(2, conditioner)}


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
