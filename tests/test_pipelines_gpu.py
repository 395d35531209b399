"""The two entry points (pipeline_i2v_eval_v0{1,2}.py, reference CLI flags) executed end to end at smoke size (--tiny:
reduced-width model, 2 sampler steps): stage 1 writes first_step/first.mp4 + first.pt, stage 2 reads them, VAE-encodes,
runs the re-noise loop, decodes and writes second_step_video/second.mp4.  One run with seeded stand-in conditioning
(--synthetic), one through the model's own GeneralConditioner with the third-party towers' outputs supplied (--towers)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, script), *args], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("mode", ["synthetic", "towers"])
def test_pipelines_run_end_to_end(tmp_path, mode):
    out = str(tmp_path / "out")
    extra = ["--synthetic"]
    if mode == "towers":
        import cv2
        import numpy as np
        img = (np.random.RandomState(0).rand(200, 160, 3) * 255).astype("uint8")
        cv2.imwrite(str(tmp_path / "in.png"), img)
        g = torch.Generator().manual_seed(0)
        torch.save({"clip": torch.randn(1, 1024, generator=g), "aes": torch.tensor([[5.2]]),
                    "depth": torch.rand(16, 24, 24, generator=g)}, str(tmp_path / "towers.pt"))
        extra = ["--towers", str(tmp_path / "towers.pt"), "--image_path", str(tmp_path / "in.png")]
    so = _run("pipeline_i2v_eval_v01.py", "--tiny", "--output_dir", out, "--seed", "1", "--elevation", "10", *extra)
    assert "first.mp4" in so
    first = torch.load(os.path.join(out, "first_step", "first.pt"))
    assert first.shape == (16, 3, 128, 128) and torch.isfinite(first.float()).all()
    assert os.path.getsize(os.path.join(out, "first_step", "first.mp4")) > 1000
    so = _run("pipeline_i2v_eval_v02.py", "--tiny", "--output_dir", out, "--seed", "1", *extra)
    assert "second.mp4" in so
    second = torch.load(os.path.join(out, "second_step_video", "second.pt"))
    assert second.shape == (16, 3, 256, 256) and torch.isfinite(second.float()).all()
    assert os.path.getsize(os.path.join(out, "second_step_video", "second.mp4")) > 1000
