"""Model-level operator API of the reference, backed by the sm_100a engine.

``registry.MODELS['uvltrack'](cfg)`` returns a :class:`UVLTrack` with the reference's four entry points and output
dict keys (lib/models/uvltrack/uvltrack.py:8-57).  The backbone and head objects registered under the reference's
names share the model's engine: the arithmetic lives in ``libuvlt_sm100.so``, these classes only carry configuration
and marshal tensors.
"""
from __future__ import annotations

from . import registry
from .engine import Engine
from .weights import ModelDims


class ModalityUnifiedFeatureExtractor:
    """Handle for the backbone half (lib/models/backbones/modality_unified_feature_extractor.py:11-77)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.dims = ModelDims.from_cfg(cfg)
        self.engine = None  # bound by UVLTrack

    def __call__(self, template, search, text, flag):
        if self.engine is None:
            raise RuntimeError("backbone is not bound to an engine: build it through MODELS['uvltrack'](cfg)")
        return self.engine.backbone(template, search, text, flag, want_logits=True)

    forward = __call__


class ModalityAdaptiveBoxHead:
    """Handle for the head half (lib/models/heads/modality_adaptive_box_head.py:10-149, built as in
    lib/models/heads/__init__.py:4-13)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.feat_sz = int(cfg.DATA.SEARCH.SIZE / 16)
        self.channel = cfg.MODEL.HEAD.HEAD_DIM
        self.softmax_one = cfg.MODEL.HEAD.SOFTMAX_ONE
        self.engine = None

    def forward(self, out_dict, test=False):
        """ModalityAdaptiveBoxHead.forward (modality_adaptive_box_head.py:62-94): updates and returns ``out_dict`` with
        cls_score / cls_score_test / bbox_map / pred_boxes / cont_score / prompts.  The test-time branch: the prompt is
        in ``out_dict['prompt']`` as UVLTrack.forward_test injects it (uvltrack.py:43); without a prompt the reference
        head derives one from the batch's own template / context masks, which is what MODELS['uvltrack'].forward does."""
        if self.engine is None:
            raise RuntimeError("head is not bound to an engine: build it through MODELS['uvltrack'](cfg)")
        if out_dict.get("prompt") is None:
            raise NotImplementedError("head.forward without out_dict['prompt'] is the training branch: call "
                                      "MODELS['uvltrack'](cfg).forward(template, search, text, template_mask, context_mask, flag)")
        out_dict.update(self.engine.head(out_dict["search"], out_dict["prompt"], out_dict["flag"]))
        return out_dict

    __call__ = forward

    def forward_prompt(self, out_dict):
        if self.engine is None:
            raise RuntimeError("head is not bound to an engine: build it through MODELS['uvltrack'](cfg)")
        return self.engine.forward_prompt(out_dict.get("tokens"), out_dict["flag"], out_dict["template_mask"],
                                          out_dict["context_mask"], out_dict.get("_text_mask"))


@registry.BACKBONES.register("modality_unified_feature_extractor")
def build_modality_unified_feature_extractor(cfg):
    return ModalityUnifiedFeatureExtractor(cfg)


@registry.HEADS.register("modality_adaptive_box_head")
def build_modality_adaptive_box_head(cfg):
    return ModalityAdaptiveBoxHead(cfg)


class UVLTrack:
    """lib/models/uvltrack/uvltrack.py:8-45 on one B200."""

    def __init__(self, backbone, box_head, max_batch=1):
        self.backbone = backbone
        self.box_head = box_head
        self.dims = backbone.dims
        self.engine = Engine(self.dims, max_batch=max_batch)
        backbone.engine = self.engine
        box_head.engine = self.engine
        self.training = False

    # nn.Module-flavoured conveniences the tracker / callers use
    def load_state_dict(self, state_dict, strict=False):
        return self.engine.load_state_dict(state_dict, strict=strict)

    def cuda(self, device=None):
        return self

    def eval(self):
        self.training = False
        return self

    def forward(self, template, search, text, template_mask, context_mask, flag):
        return self.engine.forward_train(template, search, text, template_mask, context_mask, flag, want_logits=True)

    __call__ = forward

    def forward_prompt_init(self, template, search, text, template_mask, context_mask, flag):
        info = self.engine.backbone(template, search, text, flag, clone=False)
        return self.engine.forward_prompt(None, info["flag"], template_mask, context_mask, info["_text_mask"])

    def forward_prompt(self, out_dict, template_mask, context_mask):
        out_dict["template_mask"] = template_mask
        out_dict["context_mask"] = context_mask
        return self.box_head.forward_prompt(out_dict)

    def forward_test(self, template, search, text, prompt, flag):
        return self.engine.forward_test(template, search, text, prompt, flag, want_logits=True)


@registry.MODELS.register("uvltrack")
def build_model(cfg, max_batch=1):
    backbone = registry.BACKBONES[cfg.MODEL.BACKBONE.TYPE](cfg)
    head = registry.HEADS[cfg.MODEL.HEAD.TYPE](cfg)
    return UVLTrack(backbone, head, max_batch=max_batch)
