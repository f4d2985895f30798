#!/usr/bin/env python
"""In-kernel timeline of attention2 (needs the TRACE build: make -C uvltrack_b200/csrc trace):
    python tools/attn_trace.py B n
CTA (0,0,0): tags 0x2ww warp start; 0x3xx / 0x4xx MMA thread of slot A / B (0x1i QK(i) issued, 0x2i p_full(i) seen,
0x3i PV(i) issued); 0x5xx / 0x6xx softmax warp 0 of slot A / B (0x1i s_full(i) seen, 0x2i S in registers, 0x3i max + first
chunk done, 0x4i pv_done(i-1) seen, 0x5i P stored, 0x60 epilogue done)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uvltrack_b200 import _cabi
_cabi.LIB_PATH = os.path.join(ROOT, "uvltrack_b200", "libuvlt_sm100_trace.so")
lib = _cabi.load()
B, n = int(sys.argv[1]), int(sys.argv[2])
H = 12
qkv = torch.randn(B, n, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, n, H * 64, device="cuda", dtype=torch.bfloat16)
def fn():
    _cabi.check(lib.uvlt_op_attention(qkv.data_ptr(), None, out.data_ptr(), B, n, H, None, 0, None), "attn")
buf = (C.c_ulonglong * (3 * 4096))()
for _ in range(3):
    fn()
torch.cuda.synchronize()
lib.uvlt_debug_trace(buf, 4096)
fn()
torch.cuda.synchronize()
cnt = lib.uvlt_debug_trace(buf, 4096)
r = np.frombuffer(buf, dtype=np.uint64)[: 3 * cnt].reshape(cnt, 3)
rows = [(int(t), int(c)) for t, c, _ in r.tolist() if t != 0xffff]
t0 = min(c for _, c in rows)
def name(tag):
    if 0x200 <= tag < 0x300: return f"warp {tag - 0x200} start"
    who = {3: "MMA A", 4: "MMA B", 5: "SM  A", 6: "SM  B"}[tag >> 8]
    ph, i = (tag >> 4) & 0xf, tag & 0xf
    if tag >> 8 in (3, 4): what = {1: "QK issued", 2: "p_full seen", 3: "PV issued"}[ph]
    else: what = {1: "s_full seen", 2: "S in regs", 3: "max+chunk0", 4: "pv_done seen", 5: "P stored", 6: "epilogue done"}[ph]
    return f"{who} {what} [{i}]"
for tag, c in sorted(rows, key=lambda x: x[1]):
    print(f"{c - t0:8d}  {name(tag)}")
