#!/usr/bin/env python
"""One attention launch shape for ncu captures: python tools/attn_one.py B n [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uvltrack_b200 import _cabi
lib = _cabi.load()
B, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H = 12
qkv = torch.randn(B, n, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, n, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(reps):
    _cabi.check(lib.uvlt_op_attention(qkv.data_ptr(), None, out.data_ptr(), B, n, H, None, 0, None), "attn")
torch.cuda.synchronize()
print("ok")
