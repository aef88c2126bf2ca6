"""Import harness for the UNMODIFIED reference (build container only)  --  TEST INFRASTRUCTURE.

``/root/reference`` is a pure-Python repository; its CSS path runs on CPU once four
modules that are imported at module top but never called on the numeric path are
stubbed (librosa, soundfile, omegaconf, sounddevice -- SURVEY.md section 8c).  Nothing is
copied: the reference is put on ``sys.path`` and imported where it lies.

The reference does not exist on the GPU box; nothing that runs there (``-m gpu`` tests,
``smoke()``, ``bench.py``) may import this module.  It is used by
``tests/golden/make_golden.py`` (fixture generation) and by the ``not gpu`` pinning tests,
which skip when ``available()`` is False.
"""
from __future__ import annotations

import os
import sys
import types
import warnings
from typing import Dict

import numpy as np

REFERENCE_ROOT = os.environ.get("NSF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "css", "css.py"))


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path symbols."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("librosa", "soundfile", "omegaconf", "sounddevice"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["omegaconf"], "OmegaConf"):
        sys.modules["omegaconf"].OmegaConf = object
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import css.css as css_mod                                     # noqa: E402
        from css.css_with_conformer.utils import mvdr_util            # noqa: E402
        from css.training import conformer_wrapper as cw              # noqa: E402
        from css.training import losses                               # noqa: E402
        from utils import numpy_utils                                 # noqa: E402
    ns = types.SimpleNamespace(css=css_mod, mvdr_util=mvdr_util, cw=cw, losses=losses, numpy_utils=numpy_utils)
    _loaded = ns
    return ns


def build_separator(weights: Dict[str, np.ndarray]):
    """Reference ConformerCssWrapper (conformer_wrapper.py:51) holding the given weights."""
    import torch
    from .css_oracle import net_dims
    ns = load()
    d = net_dims(weights)
    # single-channel models have no IPD pairs (conformer_v1.0_sc.yaml: extractor_conf.ipd_index = '')
    ext = ns.cw.ExtractorCfg(ipd_index='') if d.in_features == 257 else ns.cw.ExtractorCfg()
    cfg = ns.cw.ConformerCssCfg(extractor_conf=ext, nnet_conf=ns.cw.NnetCfg(conformer_conf=ns.cw.ConformerCfg(
        attention_dim=d.d_model, attention_heads=d.n_heads, num_blocks=d.n_blocks,
        linear_units=d.d_ff, kernel_size=d.kernel_size, dropout_rate=0.0),
        in_features=d.in_features))
    sep = ns.cw.ConformerCssWrapper(cfg).eval()
    sd = sep.state_dict()
    new = {}
    for k, v in sd.items():
        if k.endswith(".K"):
            new[k] = v                     # keep the reference's own STFT kernels
        else:
            new[k] = torch.from_numpy(np.asarray(weights[k])).reshape(v.shape).to(v.dtype)
    sep.load_state_dict(new)
    return sep


def sample_mixture(n_samples: int = 160000, offset: int = 0) -> np.ndarray:
    """The reference's bundled 10 s 7-ch training example, decoded per
    css/training/simulated_dataset.py:134-151,247-251: int16 [160000, 7] / mixture_scale."""
    import glob
    import json
    d = os.path.join(REFERENCE_ROOT, "sample_data", "css_train_set")
    meta_file = glob.glob(os.path.join(d, "*.json"))[0]
    meta = json.load(open(meta_file))
    col = meta["columns"]["mixture"]
    raw = np.fromfile(meta_file[:-len(".json")] + ".mixture", dtype=col["dtype"]).reshape(col["shape"])
    scale = eval(meta["columns"]["mixture_scale"]["values"])
    scale = float(np.asarray(scale).reshape(-1)[0])
    x = raw[offset:offset + n_samples].astype(np.float32) / np.float32(scale)
    return x, raw[offset:offset + n_samples], scale
