"""Parity of the CUDA path (through the C ABI) against the committed reference outputs (tests/golden) and the oracle.

Tolerances (north_star: 1e-2 for the bf16 path; boxes +-1 px):
  * every compared tensor: rel-L2 <= 1e-2
  * box maps / boxes (normalised coordinates, what the +-1 px requirement is about): max-abs <= 1e-2/2.56 -> 1 px of 256
  * cls score map in [0,1]: max-abs <= 3e-2 (the synthetic weights put a x4 gain on the last cls conv to make the map
    peaked, SURVEY.md H4, which multiplies the bf16 feature error by the same factor; rel-L2 stays <= 1e-2)
"""
import numpy as np
import pytest
import torch

from util import BF16_REL_L2, dims_of, golden_cases, load_golden, max_abs, rel_l2

from uvltrack_b200 import NestedTensor, config, registry
from uvltrack_b200.weights import synthetic_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu

PX = 1.0 / 256.0  # one pixel of the 256^2 search crop in normalised coordinates
CLS_MAX_ABS = 3e-2        # sharpened synthetic cls tower (x4 on top of x2 conv gains), 12-layer model
CLS_MAX_ABS_LARGE = 4e-2  # 24 layers accumulate ~sqrt(2) more bf16 rounding noise in the features


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module", params=golden_cases())
def case(request):
    g, meta = load_golden(request.param)
    dims = dims_of(meta)
    sd = synthetic_state_dict(dims, seed=meta["weight_seed"])
    inp = synthetic_inputs(dims, meta["batch"], meta["mode"], seed=meta["input_seed"])
    cfg = config.baseline_cfg(meta["arch"], meta["template_size"], meta["search_size"])
    model = registry.MODELS["uvltrack"](cfg, max_batch=max(4, meta["batch"]))
    ignored = model.load_state_dict(sd)
    assert ignored == []  # the synthetic state_dict holds exactly the tensors of the hot path
    yield g, meta, dims, inp, model
    model.engine.close()


def _check_head(out, g, prefix="", large=False):
    assert rel_l2(out["cls_score_test"].cpu().numpy(), g[prefix + ("cls" if prefix else "cls_score_test")]) < BF16_REL_L2
    assert max_abs(out["cls_score_test"].cpu().numpy(), g[prefix + ("cls" if prefix else "cls_score_test")]) < \
        (CLS_MAX_ABS_LARGE if large else CLS_MAX_ABS)
    for k in ("bbox_map", "cont_score"):
        assert rel_l2(out[k].cpu().numpy(), g[prefix + k]) < BF16_REL_L2, k
    assert max_abs(out["bbox_map"].cpu().numpy(), g[prefix + "bbox_map"]) < 2.56 * PX


def test_forward_test_matches_reference(case):
    g, meta, dims, inp, model = case
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    out = model.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    torch.cuda.synchronize()
    assert model.engine.last_launch_count > 50  # the CUDA path ran (kernels counted by the library)
    _check_head(out, g, large=meta["arch"] == "large")
    for k in ("logits", "vis_token", "txt_token"):
        assert rel_l2(out[k].cpu().numpy(), g[k]) < BF16_REL_L2, k
    assert rel_l2(out["search"].cpu().numpy()[:, ::8, ::4], g["search_sub"]) < BF16_REL_L2
    assert rel_l2(out["template"].cpu().numpy()[:, ::8, ::4], g["template_sub"]) < BF16_REL_L2
    assert rel_l2(out["text"].cpu().numpy()[:, ::4, ::4], g["text_sub"]) < BF16_REL_L2
    assert rel_l2(np.linalg.norm(out["search"].cpu().numpy(), axis=-1), g["search_rownorm"]) < BF16_REL_L2 / 2
    # output dict surface of the reference (a9 + a12)
    for k in ("search", "template", "text", "vis_token", "txt_token", "flag", "logits", "cls_score", "cls_score_test",
              "bbox_map", "pred_boxes", "cont_score", "prompts", "prompt"):
        assert k in out, k
    S, B = dims.feat_size, meta["batch"]
    assert out["cls_score_test"].shape == (B, S, S) and out["bbox_map"].shape == (B, S * S, 4)
    assert out["pred_boxes"].shape == (B, 1, 4) and out["cont_score"].shape == (B, S * S, 3)
    assert out["logits"].shape == (B, len(dims.cont_loss_layers), S, S)
    # pred_boxes: identical cell unless the reference's own top-2 margin is inside the bf16 noise
    ref_score = g["cls_score_test"].reshape(B, -1) * torch.from_numpy(g["cont_score"]).softmax(-1)[..., 0].numpy()
    for b in range(B):
        top2 = np.sort(ref_score[b])[-2:]
        if top2[1] - top2[0] > 2e-2:
            assert max_abs(out["pred_boxes"][b].cpu().numpy(), g["pred_boxes"][b]) < PX


def test_track_decode_matches_reference(case):
    """Hanning-window merge / argmax / gather on the device vs the reference's host code (stored in the golden)."""
    g, meta, dims, inp, model = case
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    model.engine.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    S = dims.feat_size
    window = torch.from_numpy(np.outer(np.hanning(S), np.hanning(S)).flatten()).cuda()
    rows = model.engine.track_decode(window)[:meta["batch"]].cpu().numpy()
    for b in range(meta["batch"]):
        ref = g["track"][b]
        if ref[6] > 2e-2:  # reference's top-1 / top-2 margin of the merged map
            assert int(rows[b, 5]) == int(ref[5])
            assert np.abs(rows[b, :4] - ref[:4]).max() < PX  # +-1 px of the 256-px crop
            assert abs(rows[b, 4] - ref[4]) < 1e-2
        # whatever cell was picked, the row must be self-consistent with the engine's own maps
        j = int(rows[b, 5])
        assert 0 <= j < S * S and j // S not in (0, S - 1) and j % S not in (0, S - 1)  # border cells have zero window


def test_graph_and_direct_launch_are_bit_identical(case):
    g, meta, dims, inp, model = case
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    args = (T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    model.engine.set_option("graph", 0)
    a = model.engine.forward_test(*args)
    model.engine.set_option("graph", 1)
    b = model.engine.forward_test(*args)
    c = model.engine.forward_test(*args)
    for k in ("tokens", "cls_score_test", "bbox_map", "cont_score"):
        assert torch.equal(a[k], b[k]) and torch.equal(b[k], c[k]), k


def test_prompter_and_training_forward(case):
    g, meta, dims, inp, model = case
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    tm, cm = T(g["template_mask"]), T(g["context_mask"])
    prompt = model.forward_prompt_init(T(inp["template"]), T(inp["search"]), text, tm, cm, T(inp["flag"]))
    assert rel_l2(prompt.cpu().numpy(), g["prompt_init"]) < BF16_REL_L2
    out = model.forward(T(inp["template"]), T(inp["search"]), text, tm, cm, T(inp["flag"]))
    assert out["cont_score"].shape[-1] == 2
    _check_head(out, g, prefix="train_", large=meta["arch"] == "large")
    # forward_prompt on the dict returned by forward_test == forward_prompt_init on the same inputs
    o2 = model.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    p2 = model.forward_prompt(o2, tm, cm)
    assert torch.equal(p2, prompt)


def test_skip_text_is_exact_in_bbox_mode(case):
    """flag 0 masks every text key (modality_unified_feature_extractor.py:47): dropping the BERT branch and the text
    rows must not change a single bit of the image-side outputs."""
    g, meta, dims, inp, model = case
    if meta["mode"] != "BBOX":
        pytest.skip("only defined when every flag is 0")
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    args = (T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    full = model.engine.forward_test(*args)
    fast = model.engine.forward_test(*args, skip_text=True)
    for k in ("cls_score_test", "bbox_map", "cont_score", "pred_boxes", "search", "template", "vis_token"):
        assert torch.equal(full[k], fast[k]), k
    assert model.engine.last_launch_count < 120


def test_text_cache_is_exact(case):
    """the BERT-only layers depend on the (constant) text alone; running them once
    (uvlt_text_encode) and restoring their rows every frame must not change a single bit."""
    g, meta, dims, inp, model = case
    if meta["mode"] == "BBOX":
        pytest.skip("no text branch in BBOX mode")
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    args = (T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    full = model.engine.forward_test(*args)
    n_full = model.engine.last_launch_count
    model.engine.text_encode(text, T(inp["flag"]))
    fast = model.engine.forward_test(*args, text_cached=True)
    assert model.engine.last_launch_count < n_full - 30   # embedding + 6 x 7 BERT kernels are gone
    for k in ("cls_score_test", "bbox_map", "cont_score", "pred_boxes", "tokens"):
        assert torch.equal(full[k], fast[k]), k
    with pytest.raises(RuntimeError):
        model.engine.forward_test(*args, text_cached=True, want_logits=True)


def test_batch_independence_and_order(case):
    """Sequences never interact: sequence b of a batch equals the same sequence run alone, bit for bit,
    and permuting the batch permutes the outputs."""
    g, meta, dims, inp, model = case
    B = meta["batch"]
    if B < 2:
        pytest.skip("needs a batch")
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    full = model.engine.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    perm = list(reversed(range(B)))
    p_text = NestedTensor(T(inp["ids"][perm]), T(inp["text_mask"][perm]))
    pout = model.engine.forward_test(T(inp["template"][perm]), T(inp["search"][perm]), p_text, T(inp["prompt"][perm]),
                                     T(inp["flag"][perm]))
    for k in ("cls_score_test", "bbox_map", "cont_score", "tokens"):
        assert torch.equal(pout[k], full[k][perm]), k
    b = B - 1
    one = model.engine.forward_test(T(inp["template"][b:b + 1]), T(inp["search"][b:b + 1]),
                                    NestedTensor(T(inp["ids"][b:b + 1]), T(inp["text_mask"][b:b + 1])),
                                    T(inp["prompt"][b:b + 1]), T(inp["flag"][b:b + 1]))
    for k in ("cls_score_test", "bbox_map", "cont_score", "tokens"):
        assert torch.equal(one[k][0], full[k][b]), k


def test_error_paths(case):
    g, meta, dims, inp, model = case
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    big = model.engine.max_batch + 1
    with pytest.raises(RuntimeError, match="batch out of range"):
        model.engine.forward_test(torch.zeros(big, 3, dims.template_size, dims.template_size).cuda(),
                                  torch.zeros(big, 3, dims.search_size, dims.search_size).cuda(),
                                  NestedTensor(torch.zeros(big, 40, dtype=torch.int64), torch.zeros(big, 40)),
                                  torch.zeros(big, 3, dims.embed_dim).cuda(), torch.zeros(big, 1, dtype=torch.int64))
    with pytest.raises(ValueError):
        model.engine.forward_test(T(inp["template"]), T(inp["search"]),
                                  NestedTensor(T(inp["ids"][:, :30]), T(inp["text_mask"][:, :30])), T(inp["prompt"]),
                                  T(inp["flag"]))


def test_batch_independence_across_gemm_kernels():
    """Batch 8 at 256/256 runs fc1 / fc2 / proj through the persistent CTA-pair GEMM (M = 4424 >= 4096) while a
    sequence run alone on the same engine uses the one-tile-per-CTA kernels.  Both accumulate every output element over
    K in the same order (k-blocks of 64 in sequence, fp32 in TMEM), so a sequence inside the batch must equal the same
    sequence alone bit for bit; and the batch must agree with the oracle-pinned small-batch path at all."""
    from uvltrack_b200.weights import ModelDims

    dims = ModelDims.base(256, 256)
    cfg = config.baseline_cfg("base", 256, 256)
    model = registry.MODELS["uvltrack"](cfg, max_batch=8)
    model.load_state_dict(synthetic_state_dict(dims, seed=0))
    inp = synthetic_inputs(dims, 8, "MIXED", seed=11)
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    full = model.engine.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    assert all(torch.isfinite(full[k]).all() for k in ("cls_score_test", "bbox_map", "tokens"))
    for b in (0, 5, 7):
        one = model.engine.forward_test(T(inp["template"][b:b + 1]), T(inp["search"][b:b + 1]),
                                        NestedTensor(T(inp["ids"][b:b + 1]), T(inp["text_mask"][b:b + 1])),
                                        T(inp["prompt"][b:b + 1]), T(inp["flag"][b:b + 1]))
        for k in ("cls_score_test", "bbox_map", "cont_score", "tokens"):
            assert torch.equal(one[k][0], full[k][b]), (b, k, float((one[k][0] - full[k][b]).abs().max()))
    model.engine.close()


@pytest.mark.parametrize("arch,z,x,B,mode", [
    ("base", 256, 256, 32, "NLBBOX"),    # BASELINE.json configs[2]: the CTA-pair GEMMs at M = 17696, attention at n = 553
    ("large", 384, 384, 8, "NLBBOX"),    # BASELINE.json configs[3]: UVLTrack-L, n = 1193, 24 layers
])
def test_forward_at_baseline_configs_vs_torch_oracle(arch, z, x, B, mode):
    """Model-level parity at the two BASELINE configurations the golden files do not reach (they stop at batch 3),
    against the PyTorch-CPU restatement of the reference forward (oracle/uvlt_oracle_torch.py, pinned to the goldens by
    tests/test_oracle_torch.py).  Tolerances = north_star's bf16 figures (tests/util.py): rel-L2 <= 1e-2 on features /
    logits, max-abs <= 1e-2 on the maps in [0, 1] -- with the UNSHARPENED cls tower (cls_sharpen = 1), so the 1e-2 bound on
    cls_score_test is checked at these shapes as well."""
    from util import BF16_MAX_ABS

    from oracle import uvlt_oracle_torch as OT
    from uvltrack_b200.weights import ModelDims

    dims = (ModelDims.base if arch == "base" else ModelDims.large)(z, x)
    sd = synthetic_state_dict(dims, seed=21, cls_sharpen=1.0)
    inp = synthetic_inputs(dims, B, mode, seed=21)
    model = registry.MODELS["uvltrack"](config.baseline_cfg(arch, z, x), max_batch=B)
    model.load_state_dict(sd)
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    out = model.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    got = {k: out[k].cpu().numpy() for k in ("search", "template", "text", "vis_token", "txt_token", "logits", "cont_score",
                                             "bbox_map", "cls_score_test", "pred_boxes")}
    model.engine.close()
    torch.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0))))
    C = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    with torch.no_grad():
        ref = OT.forward_test(OT.to_torch(sd), dims, C(inp["template"]), C(inp["search"]), C(inp["ids"]), C(inp["text_mask"]),
                              C(inp["prompt"]), C(inp["flag"].reshape(-1)), want_logits=True)
    ref = {k: v.numpy() for k, v in ref.items() if torch.is_tensor(v)}
    for k in ("search", "template", "text", "vis_token", "txt_token", "logits", "cont_score"):
        e = rel_l2(got[k], ref[k])
        print(f"[{arch} {z}/{x} B={B}] {k:14s} rel_l2={e:.3e}")
        assert e < BF16_REL_L2, (k, e)
    for k in ("bbox_map", "cls_score_test"):
        e = max_abs(got[k], ref[k])
        print(f"[{arch} {z}/{x} B={B}] {k:14s} max_abs={e:.3e}")
        assert e < BF16_MAX_ABS, (k, e)
    # every sequence's box is its own argmax row (batch rows do not mix)
    for b in range(B):
        j = int(np.argmax((got["cls_score_test"][b].reshape(-1) *
                           np.exp(got["cont_score"][b][:, 0] - np.log(np.exp(got["cont_score"][b]).sum(-1))))))
        assert np.allclose(got["pred_boxes"][b, 0], got["bbox_map"][b, j])


def test_head_operator_entry_matches_forward_test(case):
    """registry.HEADS['modality_adaptive_box_head'](cfg).forward(backbone_info) (modality_adaptive_box_head.py:62-94,
    test branch) on the backbone's own output == the head half of forward_test, bit for bit."""
    g, meta, dims, inp, model = case
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    full = model.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
    info = model.backbone(T(inp["template"]), T(inp["search"]), text, T(inp["flag"]))
    for k in ("search", "template", "text", "vis_token", "txt_token", "flag"):
        assert k in info
    info["prompt"] = T(inp["prompt"])                       # what UVLTrack.forward_test injects (uvltrack.py:43)
    out = model.box_head.forward(dict(info))
    for k in ("cls_score", "cls_score_test", "bbox_map", "pred_boxes", "cont_score", "prompts"):
        assert torch.equal(out[k], full[k]), k
    # external features (not the engine's own stream): same maps
    model.backbone(T(inp["template"]), T(inp["search"] * 0.5), text, T(inp["flag"]))   # overwrite the engine's stream
    out2 = model.box_head.forward({"search": full["search"].clone(), "prompt": T(inp["prompt"]), "flag": T(inp["flag"])})
    for k in ("cls_score_test", "bbox_map", "cont_score"):
        assert torch.equal(out2[k], full[k]), k
    with pytest.raises(NotImplementedError):
        model.box_head.forward({"search": full["search"], "flag": T(inp["flag"])})
