"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules for golden-vector generation.

This file never ships on a product path.  It only works where ``/root/reference`` is mounted (the
build container); the GPU box has no such directory, so nothing in ``-m gpu`` tests, ``smoke()`` or
``bench.py`` may call it.  ``oracle/gen_golden.py`` uses it to freeze outputs of the reference's own
code into ``tests/golden/*.npz``.

The reference cannot be imported as-is here: pytorch_lightning, hydra, librosa,
python-Levenshtein and soundfile are absent, and ``data/__init__.py`` pulls in ``data_loader.py`` which
dies on ``scipy.signal.hamming``.  Five tiny ``sys.modules`` stand-ins are enough for
``wav2letter.py``, ``jasper.py``, ``base_asr_models.py``, ``decoder.py`` and ``novograd.py`` to import
verbatim and run on CPU.
"""
import importlib
import os
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = os.environ.get("W2L_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "wav2letter.py"))


class AttrDict(dict):
    """Stand-in for omegaconf.DictConfig: attribute access + ``.get`` (what the reference uses)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _instantiate(cfg, **kwargs):
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    mod, _, name = target.rpartition(".")
    fn = getattr(importlib.import_module(mod), name)
    cfg.update(kwargs)
    return fn(**cfg)


def install_stubs():
    if "pytorch_lightning" not in sys.modules:
        ptl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def log_dict(self, *a, **k):
                pass

            def optimizers(self):
                return self._stub_optimizer

        ptl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = ptl
    if "hydra" not in sys.modules:
        hydra = types.ModuleType("hydra")
        hutils = types.ModuleType("hydra.utils")
        hutils.instantiate = _instantiate
        hydra.utils = hutils
        sys.modules["hydra"] = hydra
        sys.modules["hydra.utils"] = hutils
    if "librosa" not in sys.modules:
        sys.modules["librosa"] = types.ModuleType("librosa")
    if "Levenshtein" not in sys.modules:
        lev = types.ModuleType("Levenshtein")
        lev.distance = _levenshtein
        sys.modules["Levenshtein"] = lev
    if "data" not in sys.modules or not hasattr(sys.modules["data"], "__w2l_stub__"):
        data = types.ModuleType("data")
        data.__path__ = [os.path.join(REFERENCE_ROOT, "data")]
        data.__w2l_stub__ = True
        sys.modules["data"] = data


def load_reference():
    """Returns a namespace with the reference modules imported verbatim from REFERENCE_ROOT."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    for name in ("base_asr_models", "wav2letter", "jasper", "decoder", "novograd"):
        setattr(ns, name, importlib.import_module(name))
    ns.label_sets = importlib.import_module("data.label_sets")
    return ns


def load_reference_features():
    """``data.data_loader`` of the reference, imported verbatim.  Extra stand-ins it needs here: ``scipy.signal.hamming`` & co
    (moved to scipy.signal.windows in current scipy), an empty ``soundfile``, and ``librosa.filters.mel`` -- librosa is absent,
    so the filterbank comes from the oracle's restatement of its published algorithm (the only part of the feature path whose
    parity is therefore unpinned; everything downstream of the filter weights is the reference's own code)."""
    import scipy.signal
    import scipy.signal.windows as W
    from oracle.w2l_oracle import mel_filterbank_slaney

    install_stubs()
    for n in ("hamming", "hann", "blackman", "bartlett"):
        if not hasattr(scipy.signal, n):
            setattr(scipy.signal, n, getattr(W, n))
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    filters = types.ModuleType("librosa.filters")
    filters.mel = lambda sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw: mel_filterbank_slaney(sr, n_fft, n_mels, fmin, fmax)
    sys.modules["librosa"].filters = filters
    sys.modules["librosa.filters"] = filters
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("data.data_loader")


def reference_model_cfg(model="wav2letter", mid_layers=None, dropout=None, labels="english_lowercase",
                        jasper_blocks=None):
    """Compose ``cfg.model`` the way Hydra would from the reference's yaml files (PyYAML only)."""
    import yaml

    ref = load_reference()
    cfgdir = os.path.join(REFERENCE_ROOT, "configuration")
    with open(os.path.join(cfgdir, "model", model + ".yaml")) as f:
        m = yaml.safe_load(f)
    with open(os.path.join(cfgdir, "audio", "standard_16k.yaml")) as f:
        m.update(yaml.safe_load(f))
    with open(os.path.join(cfgdir, "optimizer", "exp_lr_optimizer.yaml")) as f:
        m.update(yaml.safe_load(f))
    lab = list(ref.label_sets.labels_map[labels])
    m.update(input_size=64, labels=lab, decoder={"_target_": "decoder.GreedyDecoder", "labels": lab})
    if mid_layers is not None:
        m["mid_layers"] = mid_layers
    if jasper_blocks is not None:
        m["jasper_blocks"] = jasper_blocks
    if dropout is not None:
        for l in m.get("layers", []):
            l["dropout"] = dropout
        for l in m.get("jasper_blocks", []):
            l["dropout"] = dropout
    return to_attr(m)
