"""uvltrack_b200: UVLTrack's per-frame forward hot path on NVIDIA B200 (sm_100a), behind the reference's operator API.

    from uvltrack_b200 import registry, config
    model = registry.MODELS['uvltrack'](config.baseline_cfg('base'))

Importing the package does not need a GPU; building a model does (there is no CPU path).
"""
from . import config, registry  # noqa: F401
from .misc import NestedTensor  # noqa: F401
from .weights import ModelDims, synthetic_inputs, synthetic_state_dict  # noqa: F401
from . import model as _model  # noqa: F401  (registers MODELS / BACKBONES / HEADS)

__all__ = ["config", "registry", "NestedTensor", "ModelDims", "synthetic_inputs", "synthetic_state_dict"]
