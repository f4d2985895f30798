#!/usr/bin/env python
"""Microbenchmarks of the two tensor-core kernels through the C ABI (run under gpurun).  Each case is captured as a
CUDA graph of `reps` back-to-back launches on one stream and timed with CUDA events on that stream.

    python tools/kernel_sweep.py gemm  [B ...]     # per-layer GEMM shapes at n=513/553 tokens per sequence
    python tools/kernel_sweep.py attn  [B ...]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uvltrack_b200 import _cabi  # noqa: E402

lib = _cabi.load()


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        s.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            s.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3 / reps)
    return best


def gemm(Bs, n=513, D=768):
    Hd = 4 * D
    for B in Bs:
        M = B * n
        a = torch.randn(M, Hd, device="cuda").to(torch.bfloat16)
        w = (torch.randn(max(3 * D, Hd), Hd, device="cuda") * 0.02).to(torch.bfloat16)
        bias = torch.zeros(Hd, device="cuda")
        out_b = torch.empty(M, Hd, device="cuda", dtype=torch.bfloat16)
        out_f = torch.zeros(M, D, device="cuda")
        for name, N_, K_, act, f32 in (("qkv", 3 * D, D, 0, 0), ("proj", D, D, 0, 1), ("fc1", Hd, D, 1, 0), ("fc2", D, Hd, 0, 1)):
            row = []
            for bn in [int(v) for v in os.environ.get("SWEEP_BNS", "128,256,512,0").split(",")]:
                def fn():
                    _cabi.check(lib.uvlt_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                                 out_f.data_ptr() if f32 else None,
                                                 out_f.data_ptr() if f32 else out_b.data_ptr(), M, N_, K_, act, f32, bn,
                                                 _cabi.current_stream()), "gemm")
                t = timed(fn)
                row.append(f"bn{bn or 'auto'}: {t * 1e6:7.2f} us {2.0 * M * N_ * K_ / t / 1e12:7.1f} TF")
            print(f"B={B:3d} M={M:6d} {name:5s} N={N_:5d} K={K_:5d} | " + " | ".join(row), flush=True)


def attn(Bs, ns=(513, 553, 321, 361), H=12):
    if os.environ.get("SWEEP_NS"):  # e.g. SWEEP_NS=1193 SWEEP_H=16 for UVLTrack-L 384
        ns = [int(v) for v in os.environ["SWEEP_NS"].split(",")]
    H = int(os.environ.get("SWEEP_H", H))
    for B in Bs:
        for n in ns:
            qkv = torch.randn(B, n, 3 * H * 64, device="cuda").to(torch.bfloat16)
            out = torch.empty(B, n, H * 64, device="cuda", dtype=torch.bfloat16)

            def fn():
                _cabi.check(lib.uvlt_op_attention(qkv.data_ptr(), None, out.data_ptr(), B, n, H, None, 0,
                                                  _cabi.current_stream()), "attn")
            t = timed(fn)
            print(f"B={B:3d} n={n:4d} attn {t * 1e6:8.2f} us {4.0 * B * H * n * n * 64 / t / 1e12:7.1f} TF", flush=True)


def ln(Bs, n=513, D=768):
    """LayerNorm chain: a trivial kernel, i.e. the per-launch floor inside a graph (compare UVLT_PDL=0 / 1)."""
    for B in Bs:
        x = torch.randn(B, n, D, device="cuda")
        g = torch.ones(D, device="cuda")
        be = torch.zeros(D, device="cuda")
        dst = torch.empty(B * n, D, device="cuda", dtype=torch.bfloat16)

        def fn():
            _cabi.check(lib.uvlt_op_layernorm(x.data_ptr(), n * D, 0, n, None, None, 0, 0, dst.data_ptr(), g.data_ptr(),
                                              be.data_ptr(), 1e-6, B, D, _cabi.current_stream()), "ln")
        t = timed(fn, reps=50)
        print(f"B={B:3d} rows={B * n:6d} layernorm {t * 1e6:7.2f} us  ({B * n * D * 6 / t / 1e9:7.1f} GB/s)  PDL={os.environ.get('UVLT_PDL', '1')}",
              flush=True)


def splitk(Bs, n=513, D=768):
    """fp32-output split-K GEMMs: the ViT fc2 shape and the head's first conv GEMM, (BN, splits) grid."""
    for B in Bs:
        for name, M, N_, K_, grid in (("fc2", B * n, D, 4 * D, (1, 2, 3, 4, 6, 8)),
                                      ("head0", B * 256, 1024, 9 * D, (1, 2, 4, 6, 9, 12))):
            a = torch.randn(M, K_, device="cuda").to(torch.bfloat16)
            w = (torch.randn(N_, K_, device="cuda") * 0.02).to(torch.bfloat16)
            out = torch.zeros(M, N_, device="cuda")
            part = torch.zeros(max(grid), M, N_, device="cuda")
            used = C.c_int(0)
            for bn in (32, 64, 128):
                os.environ["UVLT_SPLITK_BN"] = str(bn)
                row = []
                for sp in grid:
                    def fn():
                        _cabi.check(lib.uvlt_op_gemm_splitk(a.data_ptr(), w.data_ptr(), None, None, out.data_ptr(),
                                                            part.data_ptr(), M, N_, K_, sp, C.byref(used),
                                                            _cabi.current_stream()), "splitk")
                    row.append(f"s{sp}: {timed(fn) * 1e6:6.2f}")
                print(f"B={B:3d} {name:5s} M={M:5d} N={N_:5d} K={K_:5d} bn{bn:3d} | " + " | ".join(row) + " us", flush=True)
            os.environ.pop("UVLT_SPLITK_BN", None)


if __name__ == "__main__":
    which = sys.argv[1]
    Bs = [int(x) for x in sys.argv[2:]] or [1, 32]
    {"gemm": gemm, "attn": attn, "ln": ln, "splitk": splitk}[which](Bs)
