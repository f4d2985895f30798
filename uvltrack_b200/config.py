"""Config tree for the hot path: attribute dict + YAML overlay with the reference's semantics
(lib/config/uvltrack/config.py:7-147 defaults, :169-187 overlay: an unknown key raises ``ValueError``).

Only the keys the per-frame path and the tracker read are given defaults (SURVEY.md section 5 "Config / flags");
training-only sections of a reference yaml (TRAIN.*, DATA.TRAIN ...) are accepted and stored verbatim so that an
unmodified ``experiments/uvltrack/*.yaml`` loads.
"""
from __future__ import annotations

import copy


class AttrDict(dict):
    """dict with attribute access, nested dicts converted on assignment."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def default_cfg() -> AttrDict:
    """Defaults of the keys the hot path reads, with the reference's default values."""
    return AttrDict({
        "MODEL": {
            "HIDDEN_DIM": 384, "NUM_OBJECT_QUERIES": 1, "POSITION_EMBEDDING": "sine", "PREDICT_MASK": False,
            "LEARNABLE_POSITION": False,
            "BACKBONE": {
                "TYPE": "mae_vit", "DROP_PATH_RATE": 0.0, "PRETRAINED_PATH": "", "FUSION_LAYER": [8, 9, 10, 11],
                "CONT_LOSS_LAYER": [4, 5, 6, 7, 8, 9, 10, 11], "TXT_TOKEN_MODE": "token",
                "LANGUAGE": {"IMPLEMENT": "pytorch", "TYPE": "bert-base-uncased", "PATH": "", "VOCAB_PATH": "",
                             "BERT": {"LR": 10e-5, "ENC_NUM": 12, "HIDDEN_DIM": 256, "MAX_QUERY_LEN": 40}},
            },
            "HEAD": {"TYPE": "anchor_free", "HEAD_DIM": 384, "CLS_TOKENIZE": True, "OFFSET_SIGMOID": True,
                     "JOINT_CLS": False, "DROP": 0.0, "SOFTMAX_ONE": False, "GROUNDING_DILATION": 1,
                     "CONTRASTIVE_CONV": False},
        },
        "TRAIN": {"CONT_WEIGHT": 1.0},
        "DATA": {"MEAN": [0.485, 0.456, 0.406], "STD": [0.229, 0.224, 0.225],
                 "SEARCH": {"SIZE": 320, "FACTOR": 5.0}, "TEMPLATE": {"SIZE": 128, "FACTOR": 2.0}},
        "TEST": {"MODE": "NL", "TEMPLATE_FACTOR": 2.0, "TEMPLATE_SIZE": 128, "SEARCH_FACTOR": 5.0, "SEARCH_SIZE": 320,
                 "EPOCH": 500, "THRESHOLD": 0.5, "UPDATE_INTERVAL": 100000},
    })


# sections of a reference yaml that only training reads: stored without key checking
_PASSTHROUGH = {("TRAIN",), ("DATA",)}


def _overlay(base: AttrDict, exp: dict, path=()):
    for k, v in exp.items():
        if k not in base:
            if any(path[:len(p)] == p for p in _PASSTHROUGH):
                base[k] = copy.deepcopy(v)
                continue
            raise ValueError("{} not exist in config.py".format(".".join(path + (k,))))
        if isinstance(v, dict) and isinstance(base[k], dict):
            _overlay(base[k], v, path + (k,))
        else:
            base[k] = copy.deepcopy(v)


def update_config(cfg: AttrDict, exp: dict) -> AttrDict:
    _overlay(cfg, exp)
    return cfg


def update_config_from_file(cfg: AttrDict, filename: str) -> AttrDict:
    """lib/config/uvltrack/config.py:183-187."""
    import yaml

    with open(filename) as f:
        return update_config(cfg, yaml.safe_load(f))


def baseline_cfg(arch: str = "base", template_size: int = 128, search_size: int = 256, mode: str = "NLBBOX") -> AttrDict:
    """The hot-path content of experiments/uvltrack/baseline_{base,large}.yaml (values from :12,22,73-89,117-123 of
    the base yaml and the large yaml's differing lines), with the crop sizes overridable."""
    large = arch == "large"
    cfg = default_cfg()
    update_config(cfg, {
        "MODEL": {
            "HIDDEN_DIM": 1024 if large else 768,
            "BACKBONE": {
                "TYPE": "modality_unified_feature_extractor",
                "PRETRAINED_PATH": f"pretrain/mae_pretrain_vit_{arch}.pth",
                "FUSION_LAYER": list(range(12, 24)) if large else list(range(6, 12)),
                "CONT_LOSS_LAYER": list(range(8, 24)) if large else list(range(3, 12)),
                "TXT_TOKEN_MODE": "cls",
                "LANGUAGE": {"TYPE": "pretrain/bert-large-uncased" if large else "pretrain/bert"},
            },
            "HEAD": {"TYPE": "modality_adaptive_box_head", "HEAD_DIM": 256, "OFFSET_SIGMOID": True,
                     "CLS_TOKENIZE": False, "JOINT_CLS": False, "SOFTMAX_ONE": True},
        },
        "TRAIN": {"CONT_WEIGHT": 1.0},
        "DATA": {"SEARCH": {"SIZE": search_size, "FACTOR": 5.0 if large else 4.0},
                 "TEMPLATE": {"SIZE": template_size, "FACTOR": 2.0}},
        "TEST": {"MODE": mode, "EPOCH": 300, "SEARCH_FACTOR": 5.0 if large else 4.0, "SEARCH_SIZE": search_size,
                 "TEMPLATE_FACTOR": 2.0, "TEMPLATE_SIZE": template_size, "UPDATE_INTERVAL": 20},
    })
    return cfg


class TrackerParams:
    """lib/test/utils/params.py:5-25."""

    def get(self, name, *default):
        if len(default) > 1:
            raise ValueError("Can only give one default value.")
        return getattr(self, name) if not default else getattr(self, name, default[0])

    def has(self, name):
        return hasattr(self, name)


def parameters(cfg: AttrDict, checkpoint=None, debug=0) -> TrackerParams:
    """lib/test/parameter/uvltrack.py:21-47 without the environment / yaml lookup: tracker params from a cfg tree."""
    p = TrackerParams()
    p.cfg = cfg
    p.template_factor = cfg.TEST.TEMPLATE_FACTOR
    p.template_size = cfg.TEST.TEMPLATE_SIZE
    p.search_factor = cfg.TEST.SEARCH_FACTOR
    p.search_size = cfg.TEST.SEARCH_SIZE
    p.grounding_size = cfg.TEST.SEARCH_SIZE
    p.checkpoint = checkpoint
    p.save_all_boxes = False
    p.debug = debug
    return p
