"""CUDA path vs the numpy oracle evaluated live on seeded inputs, at shapes the golden files do not cover
(other crop sizes -> other token counts / tile tails; the 'mean' text token; softmax_one off)."""
import numpy as np
import pytest
import torch

from util import BF16_MAX_ABS, BF16_REL_L2, max_abs, rel_l2

from oracle import uvlt_oracle as O
from uvltrack_b200 import NestedTensor, config, registry
from uvltrack_b200.weights import ModelDims, synthetic_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("z,x,B,mode,variant", [
    (64, 128, 4, "MIXED", "default"),      # n = 1+16+64(+40): single q tile, tails everywhere
    (128, 384, 2, "MIXED", "default"),     # n = 1+64+576(+40) = 641 / 681 (the '384' stress shape)
    (64, 128, 3, "MIXED", "mean_token"),   # TXT_TOKEN_MODE 'mean'
    (64, 128, 3, "MIXED", "no_softmax_one"),
])
def test_forward_test_vs_oracle(z, x, B, mode, variant):
    dims = ModelDims.base(z, x)
    cfg = config.baseline_cfg("base", z, x)
    if variant == "mean_token":
        dims.txt_token_mode = "mean"
        cfg.MODEL.BACKBONE.TXT_TOKEN_MODE = "mean"
    if variant == "no_softmax_one":
        dims.softmax_one = False
        cfg.MODEL.HEAD.SOFTMAX_ONE = False
    # cls_sharpen=1 (default 4): the sharpening gain of the synthetic cls tower multiplies the bf16 feature error into
    # the score map; the peaked map it buys is only needed by the argmax / box tests, not by this tensor comparison
    sd = synthetic_state_dict(dims, seed=11, cls_sharpen=1.0)
    inp = synthetic_inputs(dims, B, mode, seed=11)
    model = registry.MODELS["uvltrack"](cfg, max_batch=B)
    model.load_state_dict(sd)
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    out = model.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    ref = O.forward_test(sd, dims, inp["template"], inp["search"], inp["ids"], inp["text_mask"], inp["prompt"],
                         inp["flag"].reshape(-1), want_logits=True)
    for k in ("search", "template", "text", "vis_token", "logits", "cont_score", "bbox_map", "cls_score_test"):
        print(f"[{z}/{x} B={B} {variant}] {k:16s} rel_l2={rel_l2(out[k].cpu().numpy(), ref[k]):.3e} "
              f"max_abs={max_abs(out[k].cpu().numpy(), ref[k]):.3e}")
    # unbounded features / logits: rel-L2 <= 1e-2; maps in [0, 1]: max-abs <= 1e-2 (tests/util.py)
    for k in ("search", "template", "text", "vis_token", "txt_token", "logits", "cont_score"):
        got, want = out[k].cpu().numpy(), ref[k]
        # TXT_TOKEN_MODE 'mean' divides by the number of real text tokens: 0/0 = NaN for a BBOX-mode sequence in the
        # reference as well (modality_unified_feature_extractor.py:80-81); the NaN rows must coincide
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), k
        assert rel_l2(np.where(nan, 0, got), np.where(nan, 0, want)) < BF16_REL_L2, k
    for k in ("bbox_map", "cls_score_test"):
        assert max_abs(out[k].cpu().numpy(), ref[k]) < BF16_MAX_ABS, k
    assert out["cont_score"].shape[-1] == (3 if dims.softmax_one else 2)
    # prompter on random target masks
    rng = np.random.default_rng(5)
    tm = rng.random((B, dims.nz)) < 0.3
    cm = rng.random((B, dims.nx)) < 0.2
    info = O.backbone(sd, dims, inp["template"], inp["search"], inp["ids"], inp["text_mask"], inp["flag"].reshape(-1),
                      want_logits=False)
    pref = O.forward_prompt(sd, dims, info, tm, cm)
    pgot = model.forward_prompt_init(T(inp["template"]), T(inp["search"]), text, T(tm), T(cm), T(inp["flag"]))
    assert rel_l2(pgot.cpu().numpy(), pref) < BF16_REL_L2
    model.engine.close()


def test_prompter_edge_masks():
    """Empty target mask (grounding init: all zeros) and full target mask."""
    dims = ModelDims.base(64, 128)
    sd = synthetic_state_dict(dims, seed=12)
    inp = synthetic_inputs(dims, 2, "NLBBOX", seed=12)
    model = registry.MODELS["uvltrack"](config.baseline_cfg("base", 64, 128), max_batch=2)
    model.load_state_dict(sd)
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    info = O.backbone(sd, dims, inp["template"], inp["search"], inp["ids"], inp["text_mask"], inp["flag"].reshape(-1),
                      want_logits=False)
    for fill in (False, True):
        tm = np.full((2, dims.nz), fill)
        cm = np.full((2, dims.nx), fill)
        pref = O.forward_prompt(sd, dims, info, tm, cm)
        pgot = model.forward_prompt_init(T(inp["template"]), T(inp["search"]), text, T(tm), T(cm), T(inp["flag"]))
        assert torch.isfinite(pgot).all()
        assert rel_l2(pgot.cpu().numpy(), pref) < BF16_REL_L2
    model.engine.close()
