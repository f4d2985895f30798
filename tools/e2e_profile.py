#!/usr/bin/env python
"""Where does a BatchTracker.track() call spend its time?  (run under gpurun)  python tools/e2e_profile.py [B] [mode]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uvltrack_b200 import config, preprocess as pp
from uvltrack_b200.synthetic import synthetic_sequence
from uvltrack_b200.tracker import BatchTracker
from uvltrack_b200.weights import ModelDims, synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = sys.argv[2] if len(sys.argv) > 2 else "BBOX"
z = x = 256
dims = ModelDims.base(z, x)
cfg = config.baseline_cfg("base", z, x, mode=mode)
params = config.parameters(cfg)
params.state_dict = synthetic_state_dict(dims, seed=0)
bt = BatchTracker(params, batch=B)
n = 120
seqs = [synthetic_sequence(n + 1, seed=b) for b in range(B)]
infos = [{"init_bbox": s[1][0], "text_ids": [101, 2023, 3899, 102]} for s in seqs]
bt.initialize([s[0][0] for s in seqs], infos)
for t in range(1, 20):
    bt.track([s[0][t] for s in seqs])
torch.cuda.synchronize()
T = {"crop": 0.0, "engine": 0.0, "post": 0.0, "total": 0.0}
S = params.search_size
cnt = 0
for t in range(21, n):   # avoid frames 20, 40.. (prompt update) by timing them separately
    images = [s[0][t] for s in seqs]
    t0 = time.perf_counter()
    bt.frame_id += 1
    rf = [0.0] * B
    for b, image in enumerate(images):
        crop, rf[b], _ = pp.sample_target(image, bt.state[b], params.search_factor, S)
        bt.crops_np[b] = crop
    t1 = time.perf_counter()
    bt.engine.track_frame_host(bt.crops, bt.template, bt.ids, bt.text_mask, bt.prompt, bt.flag, bt.window_dev, bt.out, B,
                               has_cont=bt.has_cont, skip_text=bt.skip_text, max_score=bt.max_score_dev, snapshot=bt.snapshot)
    t2 = time.perf_counter()
    for b, image in enumerate(images):
        H, W = image.shape[:2]
        row = bt.out_np[b]
        pred_box = (row[:4] * np.float32(S) / np.float32(rf[b])).tolist()
        bt.state[b] = pp.clip_box(pp.map_box_back(bt.state[b], pred_box, rf[b], S), H, W, margin=10)
    t3 = time.perf_counter()
    T["crop"] += t1 - t0; T["engine"] += t2 - t1; T["post"] += t3 - t2; T["total"] += t3 - t0
    cnt += 1
print({k: round(v / cnt * 1e6, 1) for k, v in T.items()}, "us per step, B =", B)
# engine call split: launch-only (no sync) vs sync
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    bt.engine.track_frame_host(bt.crops, bt.template, bt.ids, bt.text_mask, bt.prompt, bt.flag, bt.window_dev, bt.out, B,
                               has_cont=bt.has_cont, skip_text=bt.skip_text, max_score=bt.max_score_dev, snapshot=bt.snapshot)
t1 = time.perf_counter()
print("engine call back-to-back us:", round((t1 - t0) / 50 * 1e6, 1))
