"""Host-side mirror of the reference interface: registries, config overlay, crops, masks, box mapping, tokenizer."""
import numpy as np
import pytest

import uvltrack_b200 as u
from uvltrack_b200 import config, preprocess as pp
from uvltrack_b200.tracker import WordPieceTokenizer, extract_token_from_nlp
from uvltrack_b200.weights import ModelDims, sincos_pos_embed, synthetic_inputs, synthetic_state_dict

from oracle import uvlt_oracle as O


def test_registry_names_match_reference():
    assert "uvltrack" in u.registry.MODELS
    assert "modality_unified_feature_extractor" in u.registry.BACKBONES
    assert "modality_adaptive_box_head" in u.registry.HEADS
    with pytest.raises(AssertionError):
        u.registry.MODELS.register("uvltrack", object())


def test_config_overlay_semantics():
    cfg = config.baseline_cfg("base")
    assert cfg.MODEL.BACKBONE.FUSION_LAYER == [6, 7, 8, 9, 10, 11] and cfg.TEST.UPDATE_INTERVAL == 20
    with pytest.raises(ValueError, match="not exist in config.py"):
        config.update_config(cfg, {"MODEL": {"HEAD": {"BOGUS": 1}}})
    config.update_config(cfg, {"TRAIN": {"ANYTHING_TRAINING_ONLY": 3}})  # training sections pass through
    d = ModelDims.from_cfg(cfg)
    assert (d.embed_dim, d.depth, d.num_heads, d.fusion_start, d.nz, d.nx, d.n_tokens) == (768, 12, 12, 6, 64, 256, 361)
    dl = ModelDims.from_cfg(config.baseline_cfg("large"))
    assert (dl.embed_dim, dl.depth, dl.num_heads, dl.fusion_start, len(dl.cont_loss_layers)) == (1024, 24, 16, 12, 16)
    cfg.MODEL.HEAD.CLS_TOKENIZE = True
    with pytest.raises(NotImplementedError):
        ModelDims.from_cfg(cfg)
    p = config.parameters(config.baseline_cfg("base"))
    assert (p.template_size, p.search_size, p.template_factor, p.search_factor) == (128, 256, 2.0, 4.0)


def test_sample_target_geometry():
    rng = np.random.default_rng(0)
    im = rng.integers(0, 255, (120, 160, 3), dtype=np.uint8)
    crop, rf, box = pp.sample_target(im, [60, 40, 20, 10], 4.0, 64)
    assert crop.shape == (64, 64, 3) and crop.dtype == np.uint8
    side = int(np.ceil(np.sqrt(200) * 4.0))
    assert rf == 64 / side
    assert np.allclose(box, [0.5 - 20 / side / 2, 0.5 - 10 / side / 2, 20 / side, 10 / side])
    # a crop hanging over the image border is zero padded
    crop2, _, _ = pp.sample_target(im, [-30, -30, 20, 20], 2.0, 32)
    assert (crop2 == 0).all()
    with pytest.raises(Exception, match="Too small"):
        pp.sample_target(im, [10, 10, 0, 0], 4.0, 64)
    # identity case: crop == resize target -> pixels copied verbatim
    crop3, rf3, _ = pp.sample_target(im, [50, 30, 16, 16], 4.0, 64)
    assert rf3 == 1.0 and np.array_equal(crop3, im[6:70, 26:90])


def test_anno2mask_and_box_mapping():
    m = pp.anno2mask(np.array([[0.25, 0.25, 0.5, 0.5]], dtype=np.float32), 8).reshape(8, 8)
    assert m[2:6, 2:6].all() and m.sum() == 16
    tiny = pp.anno2mask(np.array([[0.51, 0.51, 0.01, 0.01]], dtype=np.float32), 8).reshape(8, 8)
    assert tiny.sum() == 1 and tiny[4, 4]  # the centre cell is always set
    state = [100.0, 50.0, 40.0, 20.0]
    box = pp.map_box_back(state, [128.0, 128.0, 40.0, 20.0], 1.0, 256)
    assert box == [100.0, 50.0, 40.0, 20.0]  # a centred prediction keeps the state
    assert pp.clip_box([-50, -50, 20, 20], 480, 640, margin=10) == [0, 0, 10, 10]
    assert pp.clip_box([700, 500, 20, 20], 480, 640, margin=10) == [630, 470, 10, 10]
    assert pp.map_box_back(state, [1, 2, 3, 4], 0.5, 256) == O.map_box_back(state, [1, 2, 3, 4], 0.5, 256)
    assert np.array_equal(pp.hanning_window(16), O.hanning_window(16))


def test_wordpiece_tokenizer(tmp_path):
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "the", "dog", "##s", "run", "##ning", ",", "white"]
    path = tmp_path / "vocab.txt"
    path.write_text("\n".join(vocab) + "\n")
    tok = WordPieceTokenizer(str(path))
    assert tok.tokenize("The white dogs, running zebra") == ["the", "white", "dog", "##s", ",", "run", "##ning", "[UNK]"]
    ids, mask = extract_token_from_nlp(tok, "the dog", 8)
    assert ids == [2, 4, 5, 3, 0, 0, 0, 0] and mask == [1, 1, 1, 1, 0, 0, 0, 0]
    ids, mask = extract_token_from_nlp(tok, "dog " * 30, 8)
    assert len(ids) == 8 and ids[0] == 2 and ids[-1] == 3 and sum(mask) == 8


def test_synthetic_weights_are_deterministic_and_complete():
    d = ModelDims.base(64, 128)
    a = synthetic_state_dict(d, seed=5)
    b = synthetic_state_dict(d, seed=5)
    assert list(a) == list(b) and all(np.array_equal(a[k], b[k]) for k in a)
    assert a["backbone.vit.pos_embed_z"].shape == (1, 16, 768) and a["backbone.vit.pos_embed_x"].shape == (1, 64, 768)
    assert not any("encoder.layer.6." in k for k in a)  # BERT layers >= fusion_start are discarded by the reference
    pe = sincos_pos_embed(768, 4)
    assert pe.shape == (16, 768) and np.allclose(pe[0, :192], 0) and np.allclose(pe[0, 192:384], 1)  # sin(0), cos(0)
    inp = synthetic_inputs(d, 3, "MIXED", seed=1)
    assert inp["flag"].reshape(-1).tolist() == [0, 1, 2] and inp["text_mask"][0].sum() == 0 and inp["ids"][1, 0] == 101


def test_opencv_linear_resize_fixed_point_model():
    """The arithmetic csrc/track.cuh implements for cv2.resize(INTER_LINEAR) on 8-bit images (11-bit fixed-point taps,
    x clamps index and weight, y clamps only the row index, VResizeLinear rounding) restated in numpy and pinned to the
    OpenCV build in this image, for upscales, downscales and the identity."""
    import cv2

    F = np.float32

    def taps(ssize, dsize, clamp):
        scale = 1.0 / (dsize / ssize)
        idx, a = np.zeros(dsize, np.int64), np.zeros((dsize, 2), np.int64)
        for d in range(dsize):
            f = F((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = F(f - F(s))
            if clamp and s < 0:
                f, s = F(0), 0
            if clamp and s >= ssize - 1:
                f, s = F(0), ssize - 1
            idx[d] = s
            a[d] = (int(np.rint(F(F(1.0) - f) * F(2048))), int(np.rint(f * F(2048))))
        return idx, a

    def resize(src, out):
        h, w, _ = src.shape
        xi, xa = taps(w, out, True)
        yi, ya = taps(h, out, False)
        s = src.astype(np.int64)
        rows = s[:, xi, :] * xa[:, 0][None, :, None] + s[:, np.minimum(xi + 1, w - 1), :] * xa[:, 1][None, :, None]
        s0, s1 = rows[np.clip(yi, 0, h - 1)], rows[np.clip(yi + 1, 0, h - 1)]
        b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
        return ((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)

    rng = np.random.default_rng(0)
    for sz in (33, 100, 196, 255, 256, 257, 511, 777):
        src = rng.integers(0, 256, (sz, sz, 3), dtype=np.uint8)
        for out in (128, 256):
            assert np.array_equal(resize(src, out), cv2.resize(src, (out, out))), (sz, out)


def test_search_window_contains_every_pixel_sample_target_reads():
    """BatchTracker.track() uploads only pp.search_window() of each frame.  Property: blanking everything OUTSIDE the
    window never changes the crop (boxes inside, over the border, outside the frame, tiny and huge)."""
    rng = np.random.default_rng(3)
    H, W = 120, 160
    frame = rng.integers(1, 256, size=(H, W, 3), dtype=np.uint8)  # no zeros: padding is distinguishable
    n_none = 0
    for i in range(400):
        w, h = rng.uniform(1.0, 90.0), rng.uniform(1.0, 90.0)
        x, y = rng.uniform(-80.0, W + 40.0), rng.uniform(-80.0, H + 40.0)
        if i % 7 == 0:  # exact .5 origins: Python's round() is half-to-even
            x, y, w, h = float(int(x)) + 0.5, float(int(y)), float(int(w) + 1), float(int(h) + 1)
        factor = [2.0, 4.0, 5.0][i % 3]
        box = [x, y, w, h]
        side = int(np.ceil(np.sqrt(w * h) * factor))
        x1, y1 = int(round(x + 0.5 * w - side * 0.5)), int(round(y + 0.5 * h - side * 0.5))
        win = pp.search_window(box, factor, H, W)
        if x1 + side <= 0 or y1 + side <= 0 or x1 >= W - 1 or y1 >= H - 1:
            # Window entirely outside the frame.  The reference's own slicing is not meaningful here (a negative stop
            # index wraps around, an empty slice reaches copyMakeBorder); clip_box(margin=10) keeps every tracked box
            # >= 10 px inside the frame, so Tracker.track() never gets here.  The window must simply be empty or tiny.
            n_none += win is None
            continue
        full, rf, _ = pp.sample_target(frame, box, factor, 64)
        assert win is not None, (i, box)
        xa, ya, xb, yb = win
        assert 0 <= xa < xb <= W and 0 <= ya < yb <= H
        masked = np.zeros_like(frame)
        masked[ya:yb, xa:xb] = frame[ya:yb, xa:xb]
        part, rf2, _ = pp.sample_target(masked, box, factor, 64)
        assert rf == rf2 and np.array_equal(full, part), (i, box, win)
    assert n_none > 0
    assert pp.search_window([10.0, 10.0, 0.0, 5.0], 4.0, H, W) is None   # degenerate box: "Too small bounding box."


def test_gemm_planner_rules_are_pinned():
    """uvlt_gemm_plan (host only): the kernel-selection rules encode B200 measurements (profiles/r01_gemm_2sm.md,
    DESIGN.md section 4); pin them for the benchmark shapes so that a change of rule is a conscious one."""
    import ctypes as C

    from uvltrack_b200 import _cabi

    lib = _cabi.load()

    def plan(M, N, K, f32=0, act=0, split_k=0, groups=1):
        p = (C.c_int32 * 4)()
        assert lib.uvlt_gemm_plan(M, N, K, groups, f32, act, split_k, p) == 0
        return list(p)[:3]   # [CTA-pair kernel, tile width, split-K]

    n = 513  # tokens per sequence at 256/256, BBOX
    # batch 1: narrow tiles for parallelism, fc2 cut six ways, head conv 0 nine ways
    assert plan(n, 2304, 768) == [0, 64, 1] and plan(n, 768, 768, 1) == [0, 32, 1]
    assert plan(n, 3072, 768, 0, 1) == [0, 64, 1] and plan(n, 768, 3072, 1, 0, 1) == [0, 64, 6]
    p = (C.c_int32 * 4)()
    for B, hs in ((1, 9), (2, 4), (4, 2), (8, 1)):
        assert lib.uvlt_gemm_plan(B * 256, 1024, 6912, 1, 0, 2, 0, p) == 0 and p[3] == hs
    # split-K fades out as the grid fills; never combined with the CTA-pair kernel
    assert [plan(B * n, 768, 3072, 1, 0, 1)[2] for B in (1, 2, 4, 8, 32)] == [6, 4, 2, 1, 1]
    # batch 32: every layer GEMM on the persistent CTA-pair kernel
    for N, K, f32, act in ((2304, 768, 0, 0), (768, 768, 1, 0), (3072, 768, 0, 1), (768, 3072, 1, 0)):
        assert plan(32 * n, N, K, f32, act, int(K > 2048)) == [1, 256, 1]
    # batch 16: proj would leave most clusters idle in its second round of tiles -> one-CTA 128-wide tiles
    assert plan(16 * n, 768, 768, 1) == [0, 128, 1] and plan(16 * n, 3072, 768, 0, 1) == [1, 256, 1]
    # below M = 4096, grouped GEMMs and N % 256 != 0 never use the pair kernel
    assert plan(4 * n, 2304, 768)[0] == 0 and plan(32 * 256, 128, 2304, 0, 2, 0, groups=4)[0] == 0
    assert plan(32 * n, 320, 768)[0] == 0
    assert lib.uvlt_gemm_plan(0, 8, 8, 1, 0, 0, 0, p) != 0


def test_runtime_switch_defaults_are_pinned():
    """The kernel switches resolved from the environment must default to the measured-best configuration (DESIGN.md
    section 4): PDL on, TMA-multicast GEMM off (10-25 % slower at M = 513), CTA-pair GEMM by rule, automatic attention
    generation, key-split attention on, packed-arithmetic softmax.  Run in a subprocess with the UVLT_* variables removed
    (the switches are read once per process)."""
    import os
    import subprocess
    import sys

    code = ("import ctypes as C; from uvltrack_b200 import _cabi; lib = _cabi.load(); o = (C.c_int32 * 6)(); "
            "assert lib.uvlt_runtime_switches(o) == 0; print(list(o))")
    env = {k: v for k, v in os.environ.items() if not k.startswith("UVLT_")}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "[1, 0, 1, 0, 1, 2]"
    r = subprocess.run([sys.executable, "-c", code], env=dict(env, UVLT_MULTICAST="1", UVLT_ATTN_V="3", UVLT_ATTN_POLY="4"),
                       capture_output=True, text=True, cwd=root)
    assert r.stdout.strip().splitlines()[-1] == "[1, 1, 1, 3, 1, 4]"
