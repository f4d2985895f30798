#!/usr/bin/env python
"""In-kernel timeline of attention3 (UVLT_ATTN_V=3; needs the TRACE build: make -C uvltrack_b200/csrc trace):
    python tools/attn3_trace.py B n
CTA 0: tags 0x3xx / 0x4xx MMA issuer of slot 0 / 1 (0x1k QK issued, 0x2k p_full seen, 0x3k PV issued; k = the slot's running
block count mod 16, across work items); 0x5xx softmax warp 0 (slot 0, first half of the columns): 0x1k s_full seen, 0x2k
block reference known (scores loaded, maximum exchanged with the other half), 0x3k pv_done of the previous block seen,
0x4k P stored, 0x60 work item stored.  Each thread records at most 40 points (the first ~10 key blocks)."""
import ctypes as C, os, sys
os.environ.setdefault("UVLT_ATTN_V", "3")
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
    elif ph == 6: return who + " epilogue: " + {0: "item stored", 1: "last PV done", 2: "row sums exchanged", 3: "staging rows written", 4: "fenced", 5: "slot barrier passed"}.get(i, str(i))
    else: what = {1: "s_full seen", 2: "reference known", 3: "pv_done seen", 4: "P stored", 7: "turn acquired", 8: "exponentials done"}[ph]
    return f"{who} {what} [{i}]"
for tag, c in sorted(rows, key=lambda x: x[1]):
    print(f"{c - t0:8d}  {name(tag)}")
