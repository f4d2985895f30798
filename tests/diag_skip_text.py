#!/usr/bin/env python
"""Diagnostic (not collected): where do forward_test(full) and forward_test(skip_text) differ in BBOX mode?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uvltrack_b200.misc import NestedTensor  # noqa: E402
from uvltrack_b200.model import build_model  # noqa: E402
from uvltrack_b200 import config  # noqa: E402
from uvltrack_b200.weights import ModelDims, synthetic_state_dict  # noqa: E402

z = x = 256
dims = ModelDims.base(z, x)
cfg = config.baseline_cfg("base", z, x, mode="BBOX")
from uvltrack_b200.weights import synthetic_inputs  # noqa: E402

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = build_model(cfg, max_batch=MB)
model.load_state_dict(synthetic_state_dict(dims, seed=0), strict=False)
for seed in range(4):
    inp = synthetic_inputs(dims, 1, "BBOX", seed=seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    args = (T(inp["template"]), T(inp["search"]), NestedTensor(T(inp["ids"]), T(inp["text_mask"])), T(inp["prompt"]), T(inp["flag"]))
    full = model.engine.forward_test(*args)
    fast = model.engine.forward_test(*args, skip_text=True)
    a, b = full["tokens"][:, :dims.n_visual], fast["tokens"][:, :dims.n_visual]
    bad = (a != b)
    rows = bad.any(-1)[0].nonzero().flatten().tolist()
    print(f"max_batch {MB} synthetic_inputs seed {seed}: {int(bad.sum())} differing elements in rows {rows[:12]}{'...' if len(rows) > 12 else ''} "
          f"max abs {float((a - b).abs().max()):.3e}; head equal: {torch.equal(full['bbox_map'], fast['bbox_map'])}", flush=True)
for seed in range(0):
    g = torch.Generator().manual_seed(seed)
    tmpl = torch.randn(1, 3, z, z, generator=g).cuda()
    srch = torch.randn(1, 3, x, x, generator=g).cuda()
    ids = torch.zeros(1, dims.text_len, dtype=torch.int64).cuda()
    mask = torch.ones(1, dims.text_len).cuda()
    prompt = torch.randn(1, 3, dims.embed_dim, generator=g).cuda()
    flag = torch.zeros(1, dtype=torch.int64).cuda()
    args = (tmpl, srch, NestedTensor(ids, mask), prompt, flag)
    full = model.engine.forward_test(*args)
    fast = model.engine.forward_test(*args, skip_text=True)
    a, b = full["tokens"][:, :dims.n_visual], fast["tokens"][:, :dims.n_visual]
    bad = (a != b)
    rows = bad.any(-1)[0].nonzero().flatten().tolist()
    print(f"seed {seed}: {int(bad.sum())} differing elements in rows {rows[:12]}{'...' if len(rows) > 12 else ''} "
          f"max abs {float((a - b).abs().max()):.3e}; head equal: {torch.equal(full['bbox_map'], fast['bbox_map'])}", flush=True)
