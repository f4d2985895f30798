#!/usr/bin/env python
"""Diagnostic (not collected by pytest): per-frame box error of the CUDA tracker against the teacher-forced oracle,
for judging the +-1 px tolerance of test_tracker_gpu.py under a kernel change.

    python tests/diag_box_error.py [BBOX|NLBBOX] [n_frames]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from util import synthetic_sequence  # noqa: E402

from oracle import uvlt_oracle as O  # noqa: E402
from uvltrack_b200 import config, preprocess as pp  # noqa: E402
from uvltrack_b200.tracker import get_tracker_class  # noqa: E402
from uvltrack_b200.weights import ModelDims, synthetic_state_dict  # noqa: E402


def main(mode="NLBBOX", n=40):
    z, x = 128, 256
    dims = ModelDims.base(z, x)
    sd = synthetic_state_dict(dims, seed=0)
    frames, gts = synthetic_sequence(n + 1, seed=4)
    cfg = config.baseline_cfg("base", z, x, mode=mode)
    params = config.parameters(cfg)
    params.state_dict = sd
    params.cfg.TEST.UPDATE_INTERVAL = 10
    params.cfg.TEST.THRESHOLD = 0.05
    tracker = get_tracker_class()(params, "synthetic")
    info = {"init_bbox": gts[0]}
    if mode == "NLBBOX":
        info["text_ids"] = [101, 2023, 3899, 2003, 2652, 102]
    tracker.initialize(frames[0], info)
    bt = tracker._bt
    window = pp.hanning_window(dims.feat_size)
    ids, mask, flag = bt.ids.cpu().numpy(), bt.text_mask.cpu().numpy(), bt.flag.cpu().numpy()
    template = bt.template.cpu().numpy()
    errs = []
    for t in range(1, n + 1):
        state_before = list(bt.state[0])
        prompt_before = bt.prompt.cpu().numpy().copy()
        tracker.track(frames[t])
        crop, rf, _ = pp.sample_target(frames[t], state_before, params.search_factor, x)
        ref = O.forward_test(sd, dims, template, pp.normalize_image(crop), ids, mask, prompt_before, flag)
        box, score, j = O.track_decode(ref["cls_score_test"][0], ref["cont_score"][0], ref["bbox_map"][0], window)
        merged = ref["cls_score_test"][0].reshape(-1).astype(np.float64) * window * \
            O.softmax(ref["cont_score"][0])[:, 0].astype(np.float64)
        top2 = np.sort(merged)[-2:]
        row = bt.out_np[0]
        if top2[1] - top2[0] > 2e-2 and int(row[5]) == j:
            errs.append(float(np.abs(row[:4] - box).max()) * x)
    errs = np.array(errs)
    print(f"{mode}: {len(errs)} decisive frames, box error in crop px: max {errs.max():.3f} mean {errs.mean():.3f} "
          f"p90 {np.percentile(errs, 90):.3f}  HEAD_SPLITS={os.environ.get('UVLT_HEAD_SPLITS', 'auto')} "
          f"SPLITK={os.environ.get('UVLT_NO_SPLITK', '-')}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "NLBBOX", int(sys.argv[2]) if len(sys.argv) > 2 else 40)
