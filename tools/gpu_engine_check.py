#!/usr/bin/env python
"""Engine bring-up on a B200: forward_test / forward_train / forward_prompt vs the committed golden vectors, with
per-tensor error tables and quick timings.  Diagnostics only; the graded checks live in tests/."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)), float(np.abs(a - b).max())


def main():
    import torch

    from uvltrack_b200 import NestedTensor, config, registry
    from uvltrack_b200.weights import ModelDims, synthetic_inputs, synthetic_state_dict

    only = sys.argv[1:] or None
    gdir = os.path.join(ROOT, "tests", "golden")
    for fn in sorted(os.listdir(gdir)):
        if not fn.endswith(".npz") or (only and not any(o in fn for o in only)):
            continue
        g = np.load(os.path.join(gdir, fn))
        meta = json.loads(str(g["meta"]))
        arch, z, x, B = meta["arch"], meta["template_size"], meta["search_size"], meta["batch"]
        dims = ModelDims.base(z, x) if arch == "base" else ModelDims.large(z, x)
        t0 = time.time()
        sd = synthetic_state_dict(dims, seed=meta["weight_seed"])
        inp = synthetic_inputs(dims, B, meta["mode"], seed=meta["input_seed"])
        cfg = config.baseline_cfg(arch, z, x)
        model = registry.MODELS["uvltrack"](cfg, max_batch=max(B, 4))
        model.load_state_dict(sd)
        print(f"=== {fn}: weights+engine ready in {time.time() - t0:.1f}s", flush=True)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
        text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
        for use_graph in (0, 1):
            model.engine.set_option("graph", use_graph)
            out = model.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))
            torch.cuda.synchronize()
            print(f"  -- graph={use_graph} launches={model.engine.last_launch_count}")
            for k, gk in (("cls_score_test", "cls_score_test"), ("bbox_map", "bbox_map"), ("cont_score", "cont_score"),
                          ("pred_boxes", "pred_boxes"), ("logits", "logits"), ("vis_token", "vis_token"),
                          ("txt_token", "txt_token")):
                r = rel(out[k].cpu().numpy(), g[gk])
                print(f"     {k:16s} rel_l2={r[0]:.3e} max_abs={r[1]:.3e}")
            r = rel(out["search"].cpu().numpy()[:, ::8, ::4], g["search_sub"])
            print(f"     {'search_sub':16s} rel_l2={r[0]:.3e} max_abs={r[1]:.3e}")
            r = rel(out["text"].cpu().numpy()[:, ::4, ::4], g["text_sub"])
            print(f"     {'text_sub':16s} rel_l2={r[0]:.3e} max_abs={r[1]:.3e}")
        # tracker merge
        win = torch.from_numpy(np.outer(np.hanning(dims.feat_size), np.hanning(dims.feat_size)).flatten()).cuda()
        dec = model.engine.track_decode(win)[:B].cpu().numpy()
        print("     track rows (engine) :", np.round(dec, 4).tolist())
        print("     track rows (golden) :", np.round(g["track"][:, :6], 4).tolist(), "margins", g["track"][:, 6].tolist())
        # prompter / training forward
        pr = model.forward_prompt_init(T(inp["template"]), T(inp["search"]), text, T(g["template_mask"]),
                                       T(g["context_mask"]), T(inp["flag"]))
        r = rel(pr.cpu().numpy(), g["prompt_init"])
        print(f"     {'prompt_init':16s} rel_l2={r[0]:.3e} max_abs={r[1]:.3e}")
        tr = model.forward(T(inp["template"]), T(inp["search"]), text, T(g["template_mask"]), T(g["context_mask"]),
                           T(inp["flag"]))
        for k in ("cont_score", "bbox_map", "pred_boxes"):
            r = rel(tr[k].cpu().numpy(), g["train_" + k])
            print(f"     train_{k:10s} rel_l2={r[0]:.3e} max_abs={r[1]:.3e}")
        # skip-text path (only meaningful when every flag is 0)
        if meta["mode"] == "BBOX":
            out = model.engine.forward_test(T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]),
                                            skip_text=True)
            for k in ("cls_score_test", "bbox_map", "cont_score"):
                r = rel(out[k].cpu().numpy(), g[k])
                print(f"     skip_text {k:12s} rel_l2={r[0]:.3e} max_abs={r[1]:.3e}")
        # timing: device-resident forward_test + decode
        model.engine.set_option("graph", 1)
        tm, sr, pm, fl = T(inp["template"]), T(inp["search"]), T(inp["prompt"]), T(inp["flag"])
        for skip in ((False, True) if meta["mode"] == "BBOX" else (False,)):
            for _ in range(5):
                model.engine.forward_test(tm, sr, text, pm, fl, skip_text=skip, clone=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 50
            e0.record()
            for _ in range(n):
                model.engine.forward_test(tm, sr, text, pm, fl, skip_text=skip, clone=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"  timing: forward_test B={B} skip_text={skip}: {ms * 1000:.0f} us/step -> {B / ms * 1000:.0f} frames/s")
        model.engine.close()
        del model, sd


if __name__ == "__main__":
    main()
