"""Config plumbing: an attribute-dict standing in for omegaconf.DictConfig and a small composer that reads
Hydra-layout yaml trees (``config.yaml`` + ``model/``, ``audio/``, ``optimizer/`` groups with ``# @package model``
headers), i.e. the reference's ``configuration/`` directory loads unchanged (train.py:28, config.yaml:1-28).
Hydra/omegaconf themselves are not required."""
import copy
import importlib
import os
import re

import yaml

from . import label_sets

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configuration")


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _set_path(cfg, dotted, value):
    keys = dotted.split(".")
    d = cfg
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def _get_path(cfg, dotted):
    d = cfg
    for k in dotted.split("."):
        d = d[k]
    return d


_INTERP = re.compile(r"^\$\{([^}]+)\}$")


def _resolve(node, root):
    if isinstance(node, dict):
        for k in list(node):
            node[k] = _resolve(node[k], root)
    elif isinstance(node, list):
        node[:] = [_resolve(v, root) for v in node]
    elif isinstance(node, str):
        m = _INTERP.match(node.strip())
        if m:
            try:
                return _resolve(copy.deepcopy(_get_path(root, m.group(1))), root)
            except (KeyError, TypeError):
                return node
    return node


def compose(config_dir=None, overrides=(), resolve_labels=True):
    """Hydra-style composition.  ``overrides``: ``["model=jasper", "model.mid_layers=20", "optimizer=..."]``.
    Returns the full config (``cfg.model`` is what the model constructors take, train.py:33)."""
    config_dir = config_dir or CONFIG_DIR
    with open(os.path.join(config_dir, "config.yaml")) as f:
        root = yaml.safe_load(f)
    groups = {}
    for item in root.pop("defaults", []):
        groups.update(item)
    values = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        if key in groups and "." not in key:
            groups[key] = val
        else:
            values.append((key, yaml.safe_load(val)))
    cfg = {}
    for group, choice in groups.items():
        path = os.path.join(config_dir, group, "%s.yaml" % choice)
        with open(path) as f:
            text = f.read()
        m = re.search(r"^#\s*@package\s+(\S+)", text, re.M)
        package = m.group(1) if m else group
        body = yaml.safe_load(text) or {}
        _merge(cfg.setdefault(package, {}) if package != "_global_" else cfg, body)
    root.pop("hydra", None)
    _merge(cfg, root)
    for key, val in values:
        _set_path(cfg, key, val)
    _resolve(cfg, cfg)
    if resolve_labels and isinstance(cfg.get("model", {}).get("labels"), str):          # train.py:30-31
        labels = list(label_sets.labels_map[cfg["model"]["labels"]])
        cfg["model"]["labels"] = labels
        if isinstance(cfg["model"].get("decoder"), dict) and "labels" in cfg["model"]["decoder"]:
            cfg["model"]["decoder"]["labels"] = labels
    return to_attr(cfg)


_TARGET_ALIASES = {
    "decoder.GreedyDecoder": "wav2letter_pytorch_b200.decoder.GreedyDecoder",
    "decoder.PrefixBeamSearchLMDecoder": "wav2letter_pytorch_b200.decoder.PrefixBeamSearchLMDecoder",
    "novograd.Novograd": "wav2letter_pytorch_b200.novograd.Novograd",
}


def instantiate(cfg, **kwargs):
    """hydra.utils.instantiate for the ``_target_`` nodes the reference uses (decoder, optimizer, scheduler):
    base_asr_models.py:22,73-76.  The reference's module paths map onto this package."""
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    target = _TARGET_ALIASES.get(target, target)
    mod, _, name = target.rpartition(".")
    fn = getattr(importlib.import_module(mod), name)
    cfg.update(kwargs)
    return fn(**cfg)
