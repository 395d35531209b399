"""TEST / BASELINE INFRASTRUCTURE -- recipe that stages the UNMODIFIED reference's own Python modules for this path
into `oracle/_ref/` so that `bench.py --impl reference` (and `cpu_baseline`) can time THE REFERENCE ITSELF on the GPU
box's host cores (`cpu_baseline.kind = "reference"`), where /root/reference does not exist.

    python -m oracle.build_ref            # in the build container (needs /root/reference); idempotent

The reference is pure Python: "building" it means importing exactly the modules the hot path needs
(VideoUNet, Denoiser, EulerEDMSampler, LinearPredictionGuider, OpenAIWrapper, AutoencoderKL) in a scratch interpreter,
listing every module file that import pulled in from the reference tree, and copying those files -- byte for byte,
with their package `__init__.py`s -- under `oracle/_ref/`.  Nothing is edited; a manifest with SHA-256 sums is written
next to them.  `oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but is NOT
gpurun-ignored, so it travels to the GPU box exactly like the built `.so`.  `__graft_entry__.build()` calls this when
/root/reference is present.  `oracle/ref_import.py` resolves the reference root as /root/reference, else oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("HI3D_REFERENCE_ROOT", "/root/reference")

_PROBE = r"""
import json, os, sys
sys.path.insert(0, {repo!r})
from oracle import ref_import as R
R.setup()
from sgm.modules.diffusionmodules.video_model import VideoUNet            # noqa
from sgm.modules.diffusionmodules.sampling import EulerEDMSampler          # noqa
from sgm.modules.diffusionmodules.denoiser import Denoiser                 # noqa
from sgm.modules.diffusionmodules.denoiser_scaling import VScalingWithEDMcNoise   # noqa
from sgm.modules.diffusionmodules.discretizer import EDMDiscretization     # noqa
from sgm.modules.diffusionmodules.guiders import LinearPredictionGuider    # noqa
from sgm.modules.diffusionmodules.wrappers import OpenAIWrapper            # noqa
from sgm.models.autoencoder import AutoencoderKL                           # noqa
from sgm.modules.autoencoding.temporal_ae import VideoDecoder              # noqa
root = os.path.realpath(R.REF_ROOT) + os.sep
files = sorted({{os.path.realpath(m.__file__) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(root)}})
print("FILES=" + json.dumps([f[len(root):] for f in files]))
"""


def available() -> bool:
    return os.path.isdir(os.path.join(SRC, "sgm"))


def is_fresh() -> bool:
    man = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    try:
        m = json.load(open(man))
        return all(os.path.exists(os.path.join(DST, f)) for f in m["files"])
    except Exception:
        return False


def build(force: bool = False) -> str:
    if not available():
        if is_fresh():
            return DST
        raise RuntimeError(f"reference tree not found at {SRC} and no staged copy under {DST}")
    if is_fresh() and not force:
        return DST
    env = dict(os.environ, HI3D_REFERENCE_ROOT=SRC)
    r = subprocess.run([sys.executable, "-c", _PROBE.format(repo=os.path.dirname(HERE))], capture_output=True, text=True, env=env)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("FILES=")]
    if r.returncode != 0 or not line:
        raise RuntimeError(f"probing the reference imports failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    files = json.loads(line[0][6:])
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    sums = {}
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        sums[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": files, "sha256": sums,
               "note": "byte-for-byte copies of the unmodified reference modules imported by the hot path; not tracked by git"},
              open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
