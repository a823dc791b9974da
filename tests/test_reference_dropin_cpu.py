"""Drop-in boundary, exercised against the REFERENCE's own registry and config loader:

``PIPELINE.register_obj(simseg_b200.pipeline.clip_b200)`` -> the reference's ``update_cfg`` with its own YAML ->
``build_from_cfg(cfg.model.name, cfg, PIPELINE)`` (``tools/seg_evaluation.py:194,212``) must construct OUR model, and its
state-dict keys must be the reference model's.  Runs only where ``/root/reference`` exists (the build container); in a
subprocess because the reference keeps its config in a process-global that freezes after one load.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import json, os, sys, tempfile
ROOT, REF = sys.argv[1], sys.argv[2]
sys.path[:0] = [REF, os.path.join(ROOT, "oracle", "shims"), ROOT]
import torch
import simseg.core                                  # must precede simseg.models (SURVEY 8c)
from simseg.core import cfg, update_cfg
from simseg.models import PIPELINE
from simseg.utils import build_from_cfg
from simseg.tasks.clip.config import task_cfg_init_fn, update_clip_config
import simseg_b200.pipeline as b200

out = {}
# --- INTEGRATION.md 1, variant (a): register under a new name, select it with model.name
PIPELINE.register_obj(b200.clip_b200)
out["registered"] = sorted(PIPELINE.obj_dict.keys())
try:
    PIPELINE.register_obj(b200.clip)                # same __name__ as the reference's entry: must collide
    out["collision"] = False
except KeyError:
    out["collision"] = True
tmp = tempfile.mkdtemp()
from transformers import BertConfig
BertConfig().save_pretrained(os.path.join(tmp, "bert-base-uncased"))
os.chdir(tmp)
update_cfg(task_cfg_init_fn, os.path.join(REF, "configs/clip/simseg.vit-s.yaml"),
           ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False", "transforms.input_size=288",
            "model.name=clip_b200", "loss.global_reduce=False"],   # the reference's own NCE needs a process group otherwise
           preprocess_fn=update_clip_config)
ours = build_from_cfg(cfg.model.name, cfg, PIPELINE)
out["ours_type"] = type(ours).__module__ + "." + type(ours).__name__
theirs = build_from_cfg("clip", cfg, PIPELINE)
out["theirs_type"] = type(theirs).__module__ + "." + type(theirs).__name__
# --- variant (b): replace the entry in place without touching the reference tree
PIPELINE._obj_dict["clip"] = b200.clip
swapped = build_from_cfg("clip", cfg, PIPELINE)
out["swapped_type"] = type(swapped).__module__ + "." + type(swapped).__name__
skip = ("pooler.", "position_ids", "token_type_ids")     # HF-version artefacts the reference never trains (SURVEY 8c)
ko = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
kt = {k: tuple(v.shape) for k, v in theirs.state_dict().items() if not any(s in k for s in skip)}
out["only_ours"] = sorted(set(ko) - set(kt))
out["only_theirs"] = sorted(set(kt) - set(ko))
out["shape_mismatch"] = sorted(k for k in ko if k in kt and ko[k] != kt[k])
out["n_keys"] = len(ko)
out["n_params_ours"] = sum(p.numel() for p in ours.parameters())
out["n_params_theirs"] = sum(p.numel() for n, p in theirs.named_parameters() if "pooler." not in n)
# a reference checkpoint loads (strict on our side), incl. the attribute path the tool's pos-embed interpolation reaches into
sd = {k: v for k, v in theirs.state_dict().items() if not any(s in k for s in skip)}
missing, unexpected = ours.load_state_dict(sd, strict=True)
out["load_ok"] = not missing and not unexpected
vis = ours.image_encoder.model.model
out["num_patches"] = vis.patch_embed.num_patches
out["pos_embed"] = list(vis.pos_embed.shape)
out["temperature"] = float(ours.loss.temperature)
out["has_cfg"] = ours.cfg is cfg
# the model has no CPU path: calling it without a GPU must fail loudly, not fall back
try:
    ours.forward_image_feature(torch.zeros(1, 3, 288, 288))
    out["cpu_call"] = "ran"
except Exception as e:
    out["cpu_call"] = type(e).__name__
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "simseg")), reason="needs the reference tree (build container only)")
def test_build_through_reference_registry_and_cfg():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=600,
                       env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(line[0][7:])
    assert "clip_b200" in out["registered"] and "clip" in out["registered"]
    assert out["collision"] is True
    assert out["ours_type"] == "simseg_b200.pipeline.CLIPModel" == out["swapped_type"]
    assert out["theirs_type"] == "simseg.models.pipelines.clip.CLIPModel"
    assert out["only_ours"] == [] and out["only_theirs"] == [] and out["shape_mismatch"] == [], out
    assert out["n_keys"] > 340 and out["n_params_ours"] == out["n_params_theirs"]
    assert out["load_ok"] and out["has_cfg"]
    assert out["num_patches"] == 324 and out["pos_embed"] == [1, 325, 384]       # transforms.input_size=288 (simseg.vit-s.yaml:70-77)
    assert abs(out["temperature"] - 0.02) < 1e-9
    assert out["cpu_call"] in ("SimsegError", "RuntimeError", "AssertionError"), out["cpu_call"]
