#!/usr/bin/env python
"""In-kernel timeline of the attention kernel (TRACE=1 build): python tools/attn_trace.py B n"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uvltrack_b200 import _cabi
lib = _cabi.load()
B, n = int(sys.argv[1]), int(sys.argv[2])
H = 12
qkv = torch.randn(B, n, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, n, H * 64, device="cuda", dtype=torch.bfloat16)
buf = (C.c_ulonglong * (3 * 4096))()
def fn():
    _cabi.check(lib.uvlt_op_attention(qkv.data_ptr(), None, out.data_ptr(), B, n, H, None, 0, _cabi.current_stream()), "attn")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fn(); s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(3):
            fn()
    g.replay(); s.synchronize()
    lib.uvlt_debug_trace(buf, 4096)
    g.replay(); s.synchronize()
    nrec = lib.uvlt_debug_trace(buf, 4096)
r = np.frombuffer(buf, dtype=np.uint64)[: 3 * nrec].reshape(nrec, 3)
sub = {0: "wait s_full", 1: "s_full passed", 2: "chunk0 done, wait pv_done", 3: "pv_done passed", 4: "P written", 5: "p_full arrived",
       8: "QK issued", 9: "PV issued"}
fixed = {0x1f0: "setup done", 0x1f1: "pdl wait passed", 0x1f8: "epilogue stored", 0x1f9: "teardown"}
# records come in per-thread groups terminated by a (0xffff, clk_at_flush, ns_at_flush) marker: ns = ns_f - (clk_f - clk) / GHz
GHZ = 1.9
out_rows = []
grp = []
for tag, clk, ns in r.tolist():
    if tag == 0xffff:
        out_rows += [(t, c, ns - (clk - c) / GHZ) for t, c in grp]
        grp = []
    else:
        grp.append((tag, clk))
r = np.array(out_rows, dtype=np.float64).reshape(-1, 3)
t0 = r[:, 2].min()
for i in np.argsort(r[:, 2], kind="stable"):
    tag, clk, ns = int(r[i][0]), int(r[i][1]), r[i][2]
    name = fixed.get(tag) or f"blk {(tag - 0x200) // 16}: {sub.get((tag - 0x200) % 16, hex(tag))}"
    print(f"{(ns - t0) / 1e3:9.2f} us  clk {clk % 10**8:9d}  {name}")
