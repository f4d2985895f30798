"""Import shim that instantiates the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (in the build container, where /root/reference is
mounted) to pin the numpy oracle and to generate ``tests/golden/*.npz``.  Nothing in the product, in ``-m gpu``
tests, in ``smoke()`` or in ``bench.py`` imports this file: /root/reference does not exist on the GPU box.

What is shimmed (no reference file is edited):
  * ``np.float``                                   (mae_vit.py:40 uses the alias removed in numpy 1.24)
  * stub modules ``timm.models.vision_transformer`` (mae_vit.py:21, shadowed by the local PatchEmbed),
    ``pytorch_pretrained_bert(.file_utils)``       (bert_backbone.py:35, tracker :16), ``easydict`` (config.py:1),
    ``matplotlib.pyplot``, ``thop``
  * ``BertModel.from_pretrained`` -> random-init BertModel of the right size; ``torch.load('pretrain/...')`` -> {}
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("UVLT_REFERENCE_ROOT", "/root/reference")


class _EasyDict(dict):
    """Minimal attr-dict with the behaviour lib/config/uvltrack/config.py relies on."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) else x for x in v)
        super().__setattr__(k, v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Make ``import lib....`` resolve to the reference tree with the stubs in place (idempotent)."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REF_ROOT, "lib")):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    if not hasattr(np, "float"):
        np.float = float  # noqa: NPY001  (reference mae_vit.py:40)
    if "easydict" not in sys.modules:
        _stub("easydict", EasyDict=_EasyDict)
    if "timm" not in sys.modules:
        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.vision_transformer", PatchEmbed=object)
    if "pytorch_pretrained_bert" not in sys.modules:
        class _Tok:  # tracker imports the name only
            @classmethod
            def from_pretrained(cls, *a, **k):
                return cls()
        _stub("pytorch_pretrained_bert", BertTokenizer=_Tok)
        _stub("pytorch_pretrained_bert.file_utils", cached_path=lambda p, **k: p, WEIGHTS_NAME="pytorch_model.bin",
              CONFIG_NAME="bert_config.json")
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
    if "thop" not in sys.modules:
        _stub("thop", profile=None, clever_format=None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def build_reference_model(arch="base", template_size=128, search_size=256, state_dict=None, seed=0):
    """Build ``registry.MODELS['uvltrack'](cfg)`` from the reference (eval mode, CPU, fp32).

    ``state_dict``: reference-format tensors (numpy or torch) loaded with strict=True minus unused keys.
    """
    install()
    import torch
    from lib.config.uvltrack import config as ref_config
    from lib.models.backbones import bert_backbone

    cfg = ref_config.cfg
    ref_config.update_config_from_file(os.path.join(REF_ROOT, "experiments/uvltrack", f"baseline_{arch}.yaml"))
    cfg.DATA.TEMPLATE.SIZE = template_size
    cfg.DATA.SEARCH.SIZE = search_size
    cfg.TEST.TEMPLATE_SIZE = template_size
    cfg.TEST.SEARCH_SIZE = search_size

    dims = {"base": (768, 12, 12, 3072), "large": (1024, 24, 16, 4096)}[arch]

    def _from_pretrained(cls, *a, **k):
        bc = bert_backbone.BertConfig(30522, hidden_size=dims[0], num_hidden_layers=dims[1],
                                      num_attention_heads=dims[2], intermediate_size=dims[3])
        return cls(bc)

    orig_fp = bert_backbone.BertModel.from_pretrained
    orig_load = torch.load
    bert_backbone.BertModel.from_pretrained = classmethod(_from_pretrained)
    torch.load = lambda path, *a, **k: {"model": {}} if str(path).startswith("pretrain") else orig_load(path, *a, **k)
    try:
        torch.manual_seed(seed)
        from lib import registry
        import lib.models.uvltrack.uvltrack  # noqa: F401  (registers MODELS['uvltrack'])
        model = registry.MODELS["uvltrack"](cfg)
    finally:
        bert_backbone.BertModel.from_pretrained = orig_fp
        torch.load = orig_load
    if state_dict is not None:
        sd = {k: torch.as_tensor(np.asarray(v)) if not torch.is_tensor(v) else v for k, v in state_dict.items()}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        if unexpected:
            raise KeyError(f"unexpected keys in synthetic state_dict: {unexpected[:5]}")
        model._uvlt_missing = list(missing)
    model.eval()
    return model, cfg


def nested_tensor(ids, mask):
    install()
    from lib.utils.misc import NestedTensor
    return NestedTensor(ids, mask)
