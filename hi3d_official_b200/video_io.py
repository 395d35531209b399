"""Host-side video I/O and checkpoint surgery of the Hi3D pipelines -- same names and semantics as the
reference helpers, no GPU work:

  tensor2vid, export_to_video      vtdm/util.py:12-49  (called at pipeline_i2v_eval_v01.py:96-98,129 / v02.py:139-141)
  widen_svd_state_dict             tool_make_init_svd_to_vid2vid.py:40-61: turns an SVD (stage-1 shaped, 8 input channels,
                                   768-d vector conditioning) UNet state dict into the stage-2 layout -- input conv widened to
                                   17 channels [x(4) | depth(9, zero-initialised) | cond latent(4)], label_emb input 768 -> 512
                                   ([zeros(256) | columns 512..767], i.e. the aesthetic slot is zeroed, elevation dropped,
                                   cond_aug kept).

`export_to_video` writes mp4 through OpenCV like the reference's default branch; GIFs go through Pillow because imageio (the
reference's GIF writer) is not part of this image; the `use_cv2=False` torchvision branch is kept when torchvision is present.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch


def tensor2vid(video: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> List[np.ndarray]:
    """(i, c, f, h, w) in [-1, 1] -> list of i*f uint8 (h, w, c) frames.  Like the reference, `video` is modified in place."""
    m = torch.tensor(mean, device=video.device, dtype=video.dtype).reshape(1, -1, 1, 1, 1)
    s = torch.tensor(std, device=video.device, dtype=video.dtype).reshape(1, -1, 1, 1, 1)
    video = video.mul_(s).add_(m)
    video.clamp_(0, 1)
    i, c, f, h, w = video.shape
    images = video.permute(0, 2, 3, 4, 1).reshape(i * f, h, w, c)
    return [(im.float().cpu().numpy() * 255).astype("uint8") for im in images.unbind(0)]


def export_to_video(video_frames: List[np.ndarray], output_video_path: Optional[str] = None, save_to_gif: bool = False,
                    use_cv2: bool = True, fps: int = 8) -> str:
    h, w, c = video_frames[0].shape
    if save_to_gif:
        from PIL import Image
        if output_video_path.endswith("mp4"):
            output_video_path = output_video_path[:-3] + "gif"
        frames = [Image.fromarray(f) for f in video_frames]
        frames[0].save(output_video_path, save_all=True, append_images=frames[1:], duration=int(round(1000 / fps)), loop=0)
        return output_video_path
    if use_cv2:
        import cv2
        writer = cv2.VideoWriter(output_video_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
        for f in video_frames:
            writer.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
        writer.release()
        return output_video_path
    import torchvision
    frames = list(video_frames)
    duration = math.ceil(len(frames) / fps)
    frames += [frames[-1]] * (duration * fps - len(frames))
    torchvision.io.write_video(output_video_path, torch.from_numpy(np.stack(frames, 0)), fps=fps, options={"crf": "17"})
    return output_video_path


def read_video_frames(path: str) -> List[np.ndarray]:
    """mp4 -> list of RGB uint8 frames (pipeline_i2v_eval_v02.py:169-176 reads first.mp4 with imageio; OpenCV here)."""
    import cv2
    cap = cv2.VideoCapture(path)
    out = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        out.append(cv2.cvtColor(fr, cv2.COLOR_BGR2RGB))
    cap.release()
    return out


def widen_svd_state_dict(svd_sd: Dict[str, torch.Tensor], scratch_sd: Dict[str, torch.Tensor],
                         verbose: bool = False) -> Dict[str, torch.Tensor]:
    """tool_make_init_svd_to_vid2vid.py:40-61.  `svd_sd`: checkpoint with stage-1 shapes (keys as in `scratch_sd`, e.g.
    `model.diffusion_model.*`); `scratch_sd`: state dict of the freshly built stage-2 model (target shapes; supplies every key
    the checkpoint lacks).  Returns the state dict to load with strict=True."""
    out: Dict[str, torch.Tensor] = {}
    for k, tgt in scratch_sd.items():
        if k in svd_sd:
            w = svd_sd[k].clone()
            if "label_emb.0.0.weight" in k:
                assert w.shape[1] == 768, f"{k}: expected 768 vector-conditioning inputs, got {w.shape[1]}"
                w = torch.cat([torch.zeros_like(w[:, :256]), w[:, 512:]], 1)
            if "diffusion_model.input_blocks.0.0.weight" in k or k == "input_blocks.0.0.weight":
                parts = [w[:, :4]] + [torch.zeros_like(w[:, :3]) for _ in range(3)] + [w[:, 4:]]
                w = torch.cat(parts, 1)
            if verbose:
                print(f"These weights are from svd: {k}")
        else:
            if verbose:
                print(f"These weights are newly added: {k}")
            w = tgt.clone()
        if tuple(w.shape) != tuple(tgt.shape):
            raise ValueError(f"{k}: widened shape {tuple(w.shape)} != stage-2 shape {tuple(tgt.shape)}")
        out[k] = w
    return out
