#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/, prints oracle-vs-reference errors

For each case the reference modules (imported from /root/reference through oracle/ref_shim.py) are loaded with the
seeded synthetic state_dict of ``uvltrack_b200.weights`` and run on the seeded inputs of ``synthetic_inputs``; the
reference outputs are stored (small tensors in full, the feature maps subsampled).  The same run checks the numpy
oracle against the reference so that a drifted oracle is caught here, where the reference exists.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, uvlt_oracle as O  # noqa: E402
from uvltrack_b200.weights import ModelDims, synthetic_inputs, synthetic_state_dict  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name, arch, template, search, batch, mode, weight seed, input seed
CASES = [
    ("base_z128_x256_b3_mixed", "base", 128, 256, 3, "MIXED", 0, 0),
    ("base_z256_x256_b1_bbox", "base", 256, 256, 1, "BBOX", 0, 1),
    ("base_z128_x256_b2_nlbbox", "base", 128, 256, 2, "NLBBOX", 1, 2),
    ("large_z128_x256_b1_nlbbox", "large", 128, 256, 1, "NLBBOX", 0, 3),
]


def dims_for(arch, z, x):
    return ModelDims.base(z, x) if arch == "base" else ModelDims.large(z, x)


def anno_masks(dims, batch, seed):
    """Random target boxes -> template/context masks in the way Tracker.anno2mask builds them
    (lib/test/tracker/uvltrack.py:183-194)."""
    rng = np.random.default_rng(500 + seed)

    def one(size):
        m = np.zeros((batch, size, size), dtype=bool)
        for b in range(batch):
            w, h = rng.uniform(0.2, 0.5, 2)
            x0, y0 = rng.uniform(0.1, 0.9 - w), rng.uniform(0.1, 0.9 - h)
            bb = np.array([x0, y0, x0 + w, y0 + h]) * size
            c = np.arange(size) + 0.5
            xm = (c > bb[0]) & (c < bb[2])
            ym = (c > bb[1]) & (c < bb[3])
            m[b] = ym[:, None] & xm[None, :]
            m[b, int((bb[1] + bb[3]) / 2), int((bb[0] + bb[2]) / 2)] = True
        return m.reshape(batch, -1)

    return one(dims.template_size // 16), one(dims.search_size // 16)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)), float(np.abs(a - b).max())


def main():
    import torch

    torch.set_num_threads(os.cpu_count())
    torch.Tensor.cuda = lambda self, *a, **k: self  # reference prompter hard-codes .cuda() (heads/utils.py:96)
    os.makedirs(GOLDEN, exist_ok=True)
    report = {}
    for name, arch, z, x, B, mode, wseed, iseed in CASES:
        t0 = time.time()
        dims = dims_for(arch, z, x)
        sd = synthetic_state_dict(dims, seed=wseed)
        inp = synthetic_inputs(dims, B, mode, seed=iseed)
        tmask, cmask = anno_masks(dims, B, iseed)
        model, _ = ref_shim.build_reference_model(arch, z, x, state_dict=sd)
        leftover = [k for k in model._uvlt_missing if not any(s in k for s in (
            "vit.norm.", "pooler.", "prompter.q.", "prompter.kv.", "prompter.proj.", "prompter.norm.", "coodinate",
            "num_batches_tracked")) and not ("bert.encoder.layer." in k and int(k.split("layer.")[1].split(".")[0]) >= dims.fusion_start)]
        assert not leftover, leftover[:5]
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
        text = ref_shim.nested_tensor(T(inp["ids"]), T(inp["text_mask"]))
        flag = T(inp["flag"])
        with torch.no_grad():
            ref = model.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), flag)
            ref = {k: (v.numpy().copy() if torch.is_tensor(v) else v) for k, v in ref.items()}
            ref_prompt = model.forward_prompt_init(T(inp["template"]), T(inp["search"]), text, T(tmask), T(cmask),
                                                   flag).numpy().copy()
            ref_train = model.forward(T(inp["template"]), T(inp["search"]), text, T(tmask), T(cmask), flag)
            ref_train = {k: (v.numpy().copy() if torch.is_tensor(v) else v) for k, v in ref_train.items()}
        # tracker merge exactly as lib/test/tracker/uvltrack.py:116-121 (CPU torch, float64 window)
        S = dims.feat_size
        window = O.hanning_window(S)
        trk = []
        for b in range(B):
            pred_boxes = torch.from_numpy(ref["bbox_map"][b]).view(-1, 4)
            pred_cls = torch.from_numpy(ref["cls_score_test"][b]).view(-1)
            pred_cont = torch.from_numpy(ref["cont_score"][b:b + 1]).softmax(-1)[:, :, 0].view(-1)
            merge = pred_cls * window * pred_cont
            j = int(torch.argmax(merge))
            top2 = torch.topk(merge, 2).values
            trk.append(list(pred_boxes[j].numpy()) + [float((pred_cls * pred_cont)[j]), j,
                                                      float(top2[0] - top2[1])])
        trk = np.array(trk, dtype=np.float64)

        # ---- oracle check (here, where the reference exists) ----
        orc = O.forward_test(sd, dims, inp["template"], inp["search"], inp["ids"], inp["text_mask"], inp["prompt"],
                             inp["flag"].reshape(-1), want_logits=True)
        errs = {}
        for k in ("search", "template", "text", "vis_token", "txt_token", "logits", "cls_score_test", "bbox_map",
                  "cont_score", "pred_boxes"):
            errs[k] = rel(orc[k], ref[k])
        info = O.backbone(sd, dims, inp["template"], inp["search"], inp["ids"], inp["text_mask"],
                          inp["flag"].reshape(-1), want_logits=False)
        errs["prompt_init"] = rel(O.forward_prompt(sd, dims, info, tmask, cmask), ref_prompt)
        otr = O.forward_train(sd, dims, inp["template"], inp["search"], inp["ids"], inp["text_mask"], tmask, cmask,
                              inp["flag"].reshape(-1))
        for k in ("cont_score", "bbox_map", "pred_boxes"):
            errs["train_" + k] = rel(otr[k], ref_train[k])
        for b in range(B):
            box, score, j = O.track_decode(orc["cls_score_test"][b], orc["cont_score"][b], orc["bbox_map"][b], window)
            assert j == int(trk[b, 5]), (name, b, j, trk[b])
        report[name] = {k: {"rel_l2": v[0], "max_abs": v[1]} for k, v in errs.items()}
        worst = max(v[0] for v in errs.values())
        print(f"{name}: oracle vs reference worst rel_l2 = {worst:.2e}  ({time.time() - t0:.1f}s)")
        for k, v in errs.items():
            print(f"    {k:20s} rel_l2={v[0]:.2e} max_abs={v[1]:.2e}")
        assert worst < 2e-4, "oracle drifted from the reference"

        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"),
            meta=json.dumps(dict(arch=arch, template_size=z, search_size=x, batch=B, mode=mode, weight_seed=wseed,
                                 input_seed=iseed)),
            cls_score_test=ref["cls_score_test"], bbox_map=ref["bbox_map"], cont_score=ref["cont_score"],
            pred_boxes=ref["pred_boxes"], logits=ref["logits"], vis_token=ref["vis_token"],
            txt_token=ref["txt_token"], search_sub=ref["search"][:, ::8, ::4], template_sub=ref["template"][:, ::8, ::4],
            text_sub=ref["text"][:, ::4, ::4], search_rownorm=np.linalg.norm(ref["search"], axis=-1),
            template_mask=tmask, context_mask=cmask, prompt_init=ref_prompt,
            train_cont_score=ref_train["cont_score"], train_bbox_map=ref_train["bbox_map"],
            train_pred_boxes=ref_train["pred_boxes"], train_cls=ref_train["cls_score_test"], track=trk)
        del model, sd
    with open(os.path.join(GOLDEN, "oracle_vs_reference.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
