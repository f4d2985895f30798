#!/usr/bin/env python
"""In-kernel timeline of the GEMM (needs a TRACE=1 build): python tools/gemm_trace.py M N K bn"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uvltrack_b200 import _cabi
lib = _cabi.load()
M, N, K, bn = [int(v) for v in sys.argv[1:5]]
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
def fn():
    _cabi.check(lib.uvlt_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), None, out.data_ptr(), M, N, K, act, 0, bn,
                                 _cabi.current_stream()), "gemm")
s = torch.cuda.Stream()
buf = (C.c_ulonglong * (3 * 4096))()
with torch.cuda.stream(s):
    fn(); s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(6):
            fn()
    g.replay(); s.synchronize()
    lib.uvlt_debug_trace(buf, 4096)
    g.replay(); s.synchronize()
    n = lib.uvlt_debug_trace(buf, 4096)
r = np.frombuffer(buf, dtype=np.uint64)[: 3 * n].reshape(n, 3)
names = {0x100: "cta start", 0x101: "setup done", 0x102: "pdl wait passed", 0x103: "first stage landed",
         0x104: "last mma issued", 0x105: "accumulator ready", 0x106: "epilogue stored", 0x107: "teardown", 0x108: "stage1 done", 0x109: "first tmem ld done", 0x10a: "stage2 group0 loaded", 0x10b: "stage2 group0 stored"}
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
order = np.argsort(r[:, 2], kind="stable")
prev_clk = None
for i in order:
    tag, clk, ns = int(r[i][0]), int(r[i][1]), r[i][2]
    print(f"{(ns - t0) / 1e3:9.2f} us  clk {clk % 10**9:10d}  {names.get(tag, hex(tag))}")
