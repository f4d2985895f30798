#!/usr/bin/env python
"""Generate tests/golden/preproc.npz and tests/golden/tokenizer.json from the UNMODIFIED reference (build container only).

    python oracle/make_golden_preproc.py

TEST INFRASTRUCTURE ONLY.  Calls the reference's own host-side functions of the tracker path on seeded inputs and stores
their outputs, so that the repo's ports (uvltrack_b200/preprocess.py, tracker.py) and the device kernels
(crop_resize_kernel, box_update in csrc/track.cuh) are pinned to the reference in-tree:

  sample_target, grounding_resize      /root/reference/lib/train/data/processing_utils.py:60-141,159-243
  UVLTrack.map_box_back, .anno2mask,   /root/reference/lib/test/tracker/uvltrack.py:167-194
  UVLTrack.grounding (box arithmetic)  /root/reference/lib/test/tracker/uvltrack.py:45-62
  UVLTrack.extract_token_from_nlp      /root/reference/lib/test/tracker/uvltrack.py:196-233
  clip_box                             /root/reference/lib/utils/box_ops.py:117-126

Frames are regenerated from seeds by the tests (make_frame below is imported from here by nobody: the tests carry their
own copy of the three-line generator, checked by the stored frame checksums).  Crops are stored as a SHA-256 plus an
16x-subsampled copy (bit-exact comparison without megabytes of fixtures).

The tokenizer fixture: the reference builds pytorch_pretrained_bert.BertTokenizer(vocab, do_lower_case=True)
(tracker :39), a package that is not in this image; its BasicTokenizer + WordpieceTokenizer algorithm is the one
`transformers.BertTokenizer` (installed) implements with its defaults, so the expected token ids are produced by
transformers on the committed mini vocabulary tests/golden/mini_vocab.txt and then pushed through the reference's own
extract_token_from_nlp for the [CLS]/[SEP]/padding logic.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_frame(seed, H, W):
    """Blocky random texture (8x8 cells) with a fine random overlay: exercises both flat areas and per-pixel detail."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(0, 255, ((H + 7) // 8, (W + 7) // 8, 3), dtype=np.uint8).repeat(8, axis=0).repeat(8, axis=1)[:H, :W]
    fine = rng.integers(0, 64, (H, W, 3), dtype=np.uint8)
    return np.ascontiguousarray((bg // 4 * 3 + fine).astype(np.uint8))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


MINI_VOCAB = """[PAD] [unused0] [UNK] [CLS] [SEP] [MASK] the a an of on in at to and with is are left right red blue white black
green small big man woman person dog cat car bike ball ##s ##ing ##ed ##er ##est ##ly run walk play jump stand sit
near behind front top bottom corner cafe naive resume uber strasse fast slow one two three 1 2 3 ##1 ##2 ##3 , . ! ? ' "
- ( ) / : ; & # 中 国 人 的 車 ##a ##b ##c ##d ##e ##f ##g ##h ##i ##k ##l ##m ##n ##o ##p ##r ##t ##u ##w ##y b c d e f g
h i k l m n o p r s t u w y track follow look find object target player girl boy shirt hat wearing holding moving
street road field sky water tree build ##ing ##s house window door second first third last next ##th ##nd ##rd""".split()

QUERIES = [
    "the red car on the left",
    "A man wearing a white shirt, holding a ball!",
    "Café naïve résumé über Straße",
    "dog's ball (left-top corner)",
    "中国人的車 near the tree",
    "running jumping walked fastest slowly",
    "player #3 & player #12",
    "tab\tseparated\nnew line\x00null�rep​zero",
    "xylophone quizzically zzz",
    "",
    "   ",
    " ".join(["the big red dog and the small blue cat"] * 8),
    "It's 2nd; the 3rd: 1st/last?",
    "HELLO World THE Man",
    "école mañana Ångström",
]


def main():
    import torch
    from transformers import BertTokenizer  # before the shim installs its `timm` stub (transformers probes for timm)

    ref_shim.install()
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference tracker hard-codes .cuda()
    from lib.test.tracker.uvltrack import UVLTrack as RefTracker
    from lib.train.data.processing_utils import grounding_resize, sample_target
    from lib.utils.box_ops import clip_box

    os.makedirs(GOLDEN, exist_ok=True)
    out = {}
    rng = np.random.default_rng(20261017)

    # ---------------- sample_target ----------------
    frames = [(11, 480, 640), (12, 360, 500), (13, 720, 1280), (14, 97, 131)]
    st_cases = []
    for fseed, H, W in frames:
        boxes = [[W * 0.4, H * 0.4, W * 0.1, H * 0.12], [-20.0, -15.0, 60.0, 50.0], [W - 30.0, H - 25.0, 70.0, 44.0],
                 [W * 0.5, -40.0, 33.3, 90.7], [-300.0, H * 0.5, 20.0, 20.0], [W + 200.0, H + 200.0, 40.0, 30.0],
                 [10.5, 20.25, 3.0, 2.0], [0.0, 0.0, float(W), float(H)], [W * 0.3, H * 0.3, 1.2, 0.9]]
        for _ in range(6):
            w, h = rng.uniform(4, 0.6 * W), rng.uniform(4, 0.6 * H)
            boxes.append([float(rng.uniform(-0.2 * W, W)), float(rng.uniform(-0.2 * H, H)), float(w), float(h)])
        for box in boxes:
            for factor, osz in ((2.0, 128), (4.0, 256), (5.0, 320), (4.0, 384)):
                st_cases.append((fseed, H, W, box, factor, osz))
    keep = rng.permutation(len(st_cases))[:120]
    st_meta, st_rf, st_bbox, st_sub, st_sha = [], [], [], [], []
    frame_cache = {}
    for idx in sorted(keep.tolist()):
        fseed, H, W, box, factor, osz = st_cases[idx]
        if fseed not in frame_cache:
            frame_cache[fseed] = make_frame(fseed, H, W)
        im = frame_cache[fseed]
        crop, rf, _amask, bbox = sample_target(im, list(box), factor, output_sz=osz, return_bbox=True)
        crop2, rf2, _ = sample_target(im, list(box), factor, output_sz=osz)   # the call track() makes
        assert rf == rf2 and np.array_equal(crop, crop2)
        st_meta.append([fseed, H, W, factor, osz] + list(box))
        st_rf.append(rf)
        st_bbox.append(bbox.numpy().reshape(4))
        st_sha.append(sha(crop))
        st_sub.append(np.pad(crop[::16, ::16], ((0, 24 - crop[::16, ::16].shape[0]), (0, 24 - crop[::16, ::16].shape[1]), (0, 0))))
    out["st_meta"] = np.array(st_meta, dtype=np.float64)
    out["st_rf"] = np.array(st_rf, dtype=np.float64)
    out["st_bbox"] = np.array(st_bbox, dtype=np.float32)
    out["st_sub"] = np.array(st_sub, dtype=np.uint8)
    out["st_sha"] = np.array(st_sha)
    out["frame_sha"] = np.array([sha(frame_cache[s]) if s in frame_cache else "" for s, _, _ in frames])
    out["frame_meta"] = np.array(frames, dtype=np.int64)
    # the "Too small bounding box." contract
    try:
        sample_target(frame_cache[11], [10.0, 10.0, 0.0, 5.0], 4.0, output_sz=256)
        raised = False
    except Exception as e:  # noqa: BLE001
        raised = "Too small" in str(e)
    assert raised

    # ---------------- grounding_resize ----------------
    gr_meta, gr_sha, gr_sub, gr_top = [], [], [], []
    for fseed, H, W in [(21, 480, 640), (22, 640, 480), (23, 333, 333), (24, 100, 719), (25, 1080, 1920), (26, 201, 200)]:
        im = make_frame(fseed, H, W)
        for osz in (256, 320, 384):
            pad, _box, _att, _m, top = grounding_resize(im, osz, torch.tensor([0., 0., 0., 0.]), None)
            assert pad.shape == (osz, osz, 3)
            gr_meta.append([fseed, H, W, osz])
            gr_sha.append(sha(pad))
            gr_sub.append(np.pad(pad[::16, ::16], ((0, 24 - pad[::16, ::16].shape[0]), (0, 24 - pad[::16, ::16].shape[1]), (0, 0))))
            gr_top.append(top)
    out["gr_meta"] = np.array(gr_meta, dtype=np.int64)
    out["gr_sha"] = np.array(gr_sha)
    out["gr_sub"] = np.array(gr_sub, dtype=np.uint8)
    out["gr_top"] = np.array(gr_top, dtype=np.int64)

    # ---------------- map_box_back + clip_box (the state update of track(), :123-125) ----------------
    n = 400
    states = np.stack([rng.uniform(-50, 700, n), rng.uniform(-50, 500, n), rng.uniform(1, 400, n), rng.uniform(1, 300, n)], 1)
    preds32 = rng.uniform(0, 1, (n, 4)).astype(np.float32)          # network box (cx, cy, w, h) in crop units
    ssz = rng.choice([256, 320, 384], n)
    crop_sz = np.ceil(np.sqrt(states[:, 2] * states[:, 3]) * 4.0)
    rfs = ssz / crop_sz
    HW = np.stack([rng.choice([480, 360, 720, 97], n), rng.choice([640, 500, 1280, 131], n)], 1)
    mbb, clipped = [], []
    for i in range(n):
        fake = types.SimpleNamespace(state=states[i].tolist(), params=types.SimpleNamespace(search_size=int(ssz[i])))
        pred_box = (torch.from_numpy(preds32[i]) * int(ssz[i]) / float(rfs[i])).tolist()   # tracker :123
        m = RefTracker.map_box_back(fake, pred_box, float(rfs[i]))
        mbb.append(m)
        clipped.append(clip_box(m, int(HW[i, 0]), int(HW[i, 1]), margin=10))
    out["bx_state"] = states
    out["bx_pred32"] = preds32
    out["bx_ssz"] = ssz.astype(np.int64)
    out["bx_rf"] = rfs
    out["bx_hw"] = HW.astype(np.int64)
    out["bx_mapped"] = np.array(mbb, dtype=np.float64)
    out["bx_clipped"] = np.array(clipped, dtype=np.float64)

    # ---------------- anno2mask ----------------
    am_boxes, am_sizes, am_masks = [], [], []
    for size in (8, 16, 20, 24):
        b = 24
        wh = rng.uniform(0.01, 0.7, (b, 2))
        xy = rng.uniform(0.0, 0.99, (b, 2)) * (1 - wh)
        boxes = np.concatenate([xy, wh], 1).astype(np.float32)
        boxes[0] = [0.5 - 0.5 / 4, 0.5 - 0.25 / 4, 1 / 4, 0.5 / 4]       # centred template box (factor 2 crop)
        boxes[1] = [0.49, 0.49, 1e-3, 1e-3]                              # smaller than a cell: only the centre cell
        boxes[2] = [0.0, 0.0, 0.999, 0.999]
        mask = RefTracker.anno2mask(None, torch.from_numpy(boxes), size).numpy()
        am_boxes.append(boxes)
        am_sizes.append(size)
        am_masks.append(np.pad(mask, ((0, 0), (0, 24 * 24 - size * size))))
    out["am_boxes"] = np.array(am_boxes, dtype=np.float32)
    out["am_sizes"] = np.array(am_sizes, dtype=np.int64)
    out["am_masks"] = np.array(am_masks, dtype=bool)

    # ---------------- the box arithmetic of Tracker.grounding (:58-62) ----------------
    g_in, g_hw, g_out = [], [], []
    for i in range(40):
        H, W = int(rng.choice([480, 360, 720, 1080, 333])), int(rng.choice([640, 500, 1280, 1920, 333]))
        pb = rng.uniform(0.05, 0.95, 4).astype(np.float32)
        im = np.zeros((H, W, 3), dtype=np.uint8)
        fake = types.SimpleNamespace(
            params=types.SimpleNamespace(grounding_size=320, template_size=128, search_size=256),
            preprocessor=types.SimpleNamespace(process=lambda a: torch.zeros(1)),
            cfg=types.SimpleNamespace(MODEL=types.SimpleNamespace(BACKBONE=types.SimpleNamespace(
                LANGUAGE=types.SimpleNamespace(BERT=types.SimpleNamespace(MAX_QUERY_LEN=40))))),
            extract_token_from_nlp=lambda nlp, n: (torch.zeros(1, n).long(), torch.zeros(1, n).long()),
            network=types.SimpleNamespace(forward=lambda *a, pb=pb: {"pred_boxes": torch.from_numpy(pb).view(1, 1, 4).clone()}))
        res = RefTracker.grounding(fake, im, {"language": "x"})
        g_in.append(pb)
        g_hw.append([H, W])
        g_out.append(res["pred_boxes"])
    out["gd_pred32"] = np.array(g_in, dtype=np.float32)
    out["gd_hw"] = np.array(g_hw, dtype=np.int64)
    out["gd_box"] = np.array(g_out, dtype=np.float64)

    np.savez_compressed(os.path.join(GOLDEN, "preproc.npz"), **out)
    print("preproc.npz:", {k: v.shape for k, v in out.items()})

    # ---------------- tokenizer ----------------
    vocab_path = os.path.join(GOLDEN, "mini_vocab.txt")
    seen, vocab = set(), []
    for t in MINI_VOCAB:
        if t not in seen:
            seen.add(t)
            vocab.append(t)
    with open(vocab_path, "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    tok = BertTokenizer(vocab_path, do_lower_case=True)
    cases = []
    for q in QUERIES:
        fake = types.SimpleNamespace(tokenizer=tok)
        ids, mask = RefTracker.extract_token_from_nlp(fake, q, 40)
        cases.append({"query": q, "tokens": tok.tokenize(q), "ids": ids[0].tolist(), "mask": mask[0].tolist()})
    with open(os.path.join(GOLDEN, "tokenizer.json"), "w", encoding="utf-8") as f:
        json.dump({"vocab": "mini_vocab.txt", "seq_length": 40, "cases": cases}, f, ensure_ascii=True, indent=1)
    print("tokenizer.json:", len(cases), "queries")


if __name__ == "__main__":
    main()
