"""Config surface of the hot path: the keys ``CLIPModel`` reads, loaded from the same
``configs/clip/*.yaml`` files the reference ships, with ``a.b.c=value`` overrides like
``simseg/core/config.py:143-179``.  When the reference package is importable its own ``update_cfg`` can be
used instead — ``CLIPModel`` only needs attribute access."""
from __future__ import annotations

import ast
import copy
import os
from typing import Iterable, Optional

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "clip")


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(d):
    if isinstance(d, dict):
        return AttrDict({k: _wrap(v) for k, v in d.items()})
    return d


def default_cfg() -> AttrDict:
    """Task defaults for the keys on the path (``simseg/tasks/clip/config.py:115-173``)."""
    return _wrap({
        "transforms": {"input_size": 224},
        "model": {
            "name": "clip", "max_length": 25, "freeze_cnn_bn": False, "syncbn": True, "interpolate_pos_embed": False,
            "image_encoder": {"name": "timm_modelzoo", "tag": "vit_base_patch16_224_in21k", "embedding_dim": 768,
                              "pretrained": True, "trainable": True},
            "text_encoder": {"name": "huggingface_modelzoo", "tag": "bert-base-uncased", "embedding_dim": 768,
                             "pretrained": True, "trainable": True, "target_token_idx": 0},
            "projection": {"name": "simple", "dim": 512, "text_projector_trainable": True,
                           "image_projector_trainable": True},
            "pool": {"name": "identity", "loda": {"image_k": 5, "text_k": 5}},
        },
        "loss": {"name": "NCE", "global_reduce": True, "group_size": -1, "smoothing": 0.0,
                 "nce_loss": {"gather_backward": False}, "temperature": {"name": "constant", "value": 0.02}},
        # tasks/clip/config.py:44-49 (the shipped YAMLs override weight_decay with 0.001, simseg.vit-s.yaml:36)
        "optim": {"name": "torch.optim.AdamW", "param": {"betas": (0.9, 0.98), "eps": 1e-6, "weight_decay": 0.1},
                  "lr": {"init": 1e-4}, "param_group_rules": {}, "grad_clip": {}},
        "data": {"batch_size": 1024},
        "dist": {"name": "torch", "fp16": True},
    })


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_constructor("tag:yaml.org,2002:python/tuple", lambda l, n: tuple(l.construct_sequence(n)))


def _merge(a: dict, b: AttrDict):
    for k, v in a.items():
        if isinstance(v, dict) and isinstance(b.get(k), dict):
            _merge(v, b[k])
        else:
            b[k] = _wrap(v)          # keys outside the hot path (runner, data, transforms...) are carried along


def load_cfg(yaml_file: Optional[str] = None, overrides: Iterable[str] = ()) -> AttrDict:
    cfg = copy.deepcopy(default_cfg())
    if yaml_file:
        path = yaml_file if os.path.exists(yaml_file) else os.path.join(CONFIG_DIR, yaml_file)
        with open(path) as f:
            _merge(yaml.load(f, Loader=_Loader) or {}, cfg)
    for ov in overrides:
        key, val = ov.split("=", 1)
        try:
            val = ast.literal_eval(val)
        except (ValueError, SyntaxError):
            pass
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node:
                raise KeyError(f"Non-existent config key: {key}")        # core/config.py:194-195 behaviour
            node = node[p]
        if parts[-1] not in node:
            raise KeyError(f"Non-existent config key: {key}")
        node[parts[-1]] = val
    return cfg
