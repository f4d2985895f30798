#!/usr/bin/env python
"""Kernel bring-up on a B200: runs every op-level case in its own subprocess (a device trap in one case must not
poison the rest) and prints one PASS/FAIL line per case.  Usage (on the GPU box):
    python tools/gpu_bringup.py            # all cases
    python tools/gpu_bringup.py --case gemm_qkv_bn128
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _lib():
    from uvltrack_b200 import _cabi

    lib = C.CDLL(_cabi.LIB_PATH)
    for name, (res, args) in _cabi.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


def _err(lib):
    return lib.uvlt_last_error().decode()


def _stats(name, got, ref, tol):
    import torch

    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel_l2 = (diff.norm() / (ref.norm() + 1e-12)).item()
    max_abs = diff.max().item()
    bad = torch.isnan(got).any().item() or torch.isinf(got).any().item()
    ok = (not bad) and rel_l2 < tol
    print(f"{'PASS' if ok else 'FAIL'} {name}: rel_l2={rel_l2:.3e} max_abs={max_abs:.3e} ref_max={denom:.3e} nan_inf={bad}",
          flush=True)
    if not ok:
        # where are the errors? (rows / cols summary helps to diagnose swizzle / descriptor mistakes)
        d2 = diff.reshape(-1, diff.shape[-1])
        rows = (d2.max(dim=1).values > 10 * tol * denom).nonzero().flatten()
        cols = (d2.max(dim=0).values > 10 * tol * denom).nonzero().flatten()
        print(f"   bad rows: n={rows.numel()} first={rows[:16].tolist()}  bad cols: n={cols.numel()} first={cols[:16].tolist()}")
        print("   got[0,:8]=", got.reshape(-1, got.shape[-1])[0, :8].tolist())
        print("   ref[0,:8]=", ref.reshape(-1, ref.shape[-1])[0, :8].tolist())
    return ok


def case_gemm(M, N, K, bn, act=0, bias=True, resid=False, out_f32=False):
    import torch

    lib = _lib()
    torch.manual_seed(0)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * (1.0 / K ** 0.5)).to(torch.bfloat16)
    b = torch.randn(N, device=dev) if bias else None
    R = torch.randn(M, N, device=dev) if resid else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    if resid and out_f32:
        out.copy_(R)  # in-place residual, as the engine uses it
        rptr = out.data_ptr()
    else:
        rptr = R.data_ptr() if resid else None
    ref = A.float() @ W.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    if resid:
        ref = ref + R
    rc = lib.uvlt_op_gemm(A.data_ptr(), W.data_ptr(), b.data_ptr() if bias else None, rptr, out.data_ptr(), M, N, K,
                          act, int(out_f32), bn, None)
    if rc:
        print("FAIL launch:", _err(lib))
        return False
    torch.cuda.synchronize()
    # timing
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not (resid and out_f32):
        ev0.record()
        for _ in range(20):
            lib.uvlt_op_gemm(A.data_ptr(), W.data_ptr(), b.data_ptr() if bias else None, rptr, out.data_ptr(), M, N,
                             K, act, int(out_f32), bn, None)
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1000 / 20
        print(f"   time {us:.1f} us/launch  {2.0 * M * N * K / us / 1e6:.1f} TFLOP/s")
    return _stats(f"gemm M={M} N={N} K={K} bn={bn} act={act} resid={resid} f32={out_f32}", out, ref, 1e-2 if not out_f32 else 2e-3)


def case_gemm_grouped():
    import torch

    lib = _lib()
    torch.manual_seed(0)
    G, M, N, K = 4, 512, 128, 2304
    A = (torch.randn(G, M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(G, N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(G, N, device="cuda")
    out = torch.zeros(M, G * N, device="cuda", dtype=torch.bfloat16)
    rc = lib.uvlt_op_gemm_grouped(A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), G, M, N, K, 2, G * N, N,
                                  0, None)
    if rc:
        print("FAIL launch:", _err(lib))
        return False
    torch.cuda.synchronize()
    ref = torch.relu(torch.einsum("gmk,gnk->mgn", A.float(), W.float()) + b[None]).reshape(M, G * N)
    return _stats("gemm_grouped", out, ref, 1e-2)


def case_attn(B, n, H, masked, vt):
    import torch

    lib = _lib()
    torch.manual_seed(0)
    D = H * 64
    qkv = (torch.randn(B, n, 3 * D, device="cuda")).to(torch.bfloat16)
    bias = None
    if masked:
        bias = torch.zeros(B, n, device="cuda")
        bias[:, : min(65, n // 2)] = -1e10  # cls + template keys ignored (flag == 1)
        if B > 1:
            bias[1] = 0
            bias[1, n - 7:] = -10000.0  # BERT-style additive padding mask
    out = torch.zeros(B, n, D, device="cuda", dtype=torch.bfloat16)
    vt_t, n_pad = None, 0
    if vt:
        n_pad = (n + 7) // 8 * 8
        vt_t = torch.zeros(B, D, n_pad, device="cuda", dtype=torch.bfloat16)
        vt_t[:, :, :n] = qkv[:, :, 2 * D:].transpose(1, 2)
    rc = lib.uvlt_op_attention(qkv.data_ptr(), bias.data_ptr() if masked else None, out.data_ptr(), B, n, H,
                               vt_t.data_ptr() if vt else None, n_pad, None)
    if rc:
        print("FAIL launch:", _err(lib))
        return False
    torch.cuda.synchronize()
    q, k, v = qkv.float().reshape(B, n, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if masked:
        s = s + bias[:, None, None, :]
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, n, D)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(20):
        lib.uvlt_op_attention(qkv.data_ptr(), bias.data_ptr() if masked else None, out.data_ptr(), B, n, H,
                              vt_t.data_ptr() if vt else None, n_pad, None)
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1000 / 20
    print(f"   time {us:.1f} us/launch  {4.0 * B * H * n * n * 64 / us / 1e6:.2f} TFLOP/s")
    return _stats(f"attn B={B} n={n} H={H} masked={masked} vt={vt}", out, ref, 1.5e-2)


def case_ln():
    import torch

    lib = _lib()
    torch.manual_seed(0)
    B, r0, r1, D = 3, 321, 40, 768
    s0 = torch.randn(B, r0, D, device="cuda") * 2 + 0.3
    s1 = torch.randn(B, r1, D, device="cuda")
    a0, a1 = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    g, be = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    dst = torch.zeros(B, r0 + r1, D, device="cuda")
    dstb = torch.zeros(B * (r0 + r1), D, device="cuda", dtype=torch.bfloat16)
    rc = lib.uvlt_op_layernorm(s0.data_ptr(), r0, s1.data_ptr(), r1, a0.data_ptr(), a1.data_ptr(), r0,
                               dst.data_ptr(), 1, dstb.data_ptr(), g.data_ptr(), be.data_ptr(), 1e-6, B, D, None)
    if rc:
        print("FAIL launch:", _err(lib))
        return False
    torch.cuda.synchronize()
    x = torch.cat([s0 + a0, s1 + a1], dim=1)
    ref = torch.nn.functional.layer_norm(x, (D,), g, be, 1e-6)
    ok = _stats("ln pre-norm stream", dst, x, 1e-6)
    ok &= _stats("ln bf16 out", dstb.reshape(B, r0 + r1, D), ref, 5e-3)
    return ok


CASES = {
    "gemm_qkv_bn128": lambda: case_gemm(361, 2304, 768, 128),
    "gemm_qkv_bn64": lambda: case_gemm(361, 2304, 768, 64),
    "gemm_qkv_bn32": lambda: case_gemm(361, 2304, 768, 32),
    "gemm_fc1_gelu": lambda: case_gemm(361, 3072, 768, 0, act=1),
    "gemm_fc2_resid": lambda: case_gemm(361, 768, 3072, 0, resid=True, out_f32=True),
    "gemm_small_m": lambda: case_gemm(40, 768, 768, 0, resid=True, out_f32=True),
    "gemm_big": lambda: case_gemm(11552, 2304, 768, 128),
    "gemm_big_fc2": lambda: case_gemm(11552, 768, 3072, 128, resid=True, out_f32=True),
    "gemm_head_k6912": lambda: case_gemm(256, 1024, 6912, 0, act=2),
    "gemm_grouped": case_gemm_grouped,
    "attn_361_mn": lambda: case_attn(2, 361, 12, True, False),
    "attn_361_vt": lambda: case_attn(2, 361, 12, True, True),
    "attn_321_mn": lambda: case_attn(1, 321, 12, False, False),
    "attn_40_mn": lambda: case_attn(2, 40, 12, True, False),
    "attn_553_mn": lambda: case_attn(2, 553, 12, True, False),
    "attn_128_mn": lambda: case_attn(1, 128, 12, False, False),
    "attn_1193_mn": lambda: case_attn(1, 1193, 16, False, False),
    "attn_b32_mn": lambda: case_attn(32, 361, 12, True, False),
    "ln": case_ln,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None, help="comma separated prefixes")
    args = ap.parse_args()
    if args.case:
        import torch

        ok = CASES[args.case]()
        torch.cuda.synchronize()
        sys.exit(0 if ok else 1)
    names = list(CASES)
    if args.only:
        pre = args.only.split(",")
        names = [n for n in names if any(n.startswith(p) for p in pre)]
    results = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", n], capture_output=True, text=True, timeout=180)
            out = (r.stdout + r.stderr).strip()
            results[n] = r.returncode == 0
        except subprocess.TimeoutExpired as e:
            out = f"TIMEOUT {e}"
            results[n] = False
        print(f"=== {n} ({time.time() - t0:.1f}s) rc_ok={results[n]}\n{out[-2500:]}", flush=True)
    print("SUMMARY", {k: ("ok" if v else "FAIL") for k, v in results.items()})
    sys.exit(0 if all(results.values()) else 1)


if __name__ == "__main__":
    main()
