#!/usr/bin/env python
"""Where does a BatchTracker.track() call (device-preprocess path) spend its time?  python tools/e2e_profile.py [B] [mode]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uvltrack_b200 import config
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
n = 100
seqs = [synthetic_sequence(n + 1, seed=b) for b in range(B)]
infos = [{"init_bbox": s[1][0], "text_ids": [101, 2023, 3899, 102]} for s in seqs]
bt.initialize([s[0][0] for s in seqs], infos)
for t in range(1, 20):
    bt.track([s[0][t] for s in seqs])
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(21, 61):
    bt.track([s[0][t] for s in seqs])
t1 = time.perf_counter()
print(f"track(): {(t1 - t0) / 40 * 1e6:.1f} us per step, B = {B}")
# pieces
H, W = seqs[0][0][0].shape[:2]
images = [s[0][30] for s in seqs]
t0 = time.perf_counter()
for _ in range(40):
    for b, image in enumerate(images):
        np.copyto(bt.frames_np[b], image)
t1 = time.perf_counter()
print(f"  np.copyto of {B} frames: {(t1 - t0) / 40 * 1e6:.1f} us")
total = bt.frames.numel()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(40):
    for b in range(B):
        bt.engine.upload_frames(bt.frames[b], b * H * W * 3, total)
    torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"  upload calls + sync: {(t1 - t0) / 40 * 1e6:.1f} us")
t0 = time.perf_counter()
for _ in range(40):
    bt.engine.track_frame_image_host(bt.frames, bt.state_dev, params.search_factor, bt.template, bt.ids, bt.text_mask,
                                     bt.prompt, bt.flag, bt.window_dev, bt.out10, B, has_cont=bt.has_cont,
                                     skip_text=bt.skip_text, max_score=bt.max_score_dev, snapshot=bt.snapshot,
                                     text_cached=bt.text_cached, uploaded=True)
t1 = time.perf_counter()
print(f"  engine call (no upload): {(t1 - t0) / 40 * 1e6:.1f} us")
t0 = time.perf_counter()
for _ in range(40):
    rows = []
    for b in range(B):
        row = bt.out10_np[b]
        st = row[:4].tolist()
        rows.append((row[4:8].astype(np.float32), float(np.float32(row[8]))))
t1 = time.perf_counter()
print(f"  host post: {(t1 - t0) / 40 * 1e6:.1f} us")
