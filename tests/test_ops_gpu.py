"""Operator-level parity through the C ABI: the tcgen05 GEMM (every epilogue, M/N tails), the fused attention kernel
(every sequence length of the path incl. tile tails and both mask dialects) and the row kernels, against fp32 torch."""
import numpy as np
import pytest
import torch

from uvltrack_b200 import _cabi

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("M,N,K,bn,act,resid,f32", [
    (361, 2304, 768, 128, 0, False, False), (361, 2304, 768, 64, 0, False, False), (361, 2304, 768, 32, 0, False, False),
    (513, 3072, 768, 0, 1, False, False), (513, 768, 3072, 0, 0, True, True), (40, 768, 768, 0, 0, True, True),
    (1, 768, 768, 0, 0, False, True), (129, 96, 64, 32, 2, False, False), (17696, 2304, 768, 128, 0, False, False),
    (256, 1024, 6912, 0, 2, False, False),
    # 128 x 256 tiles (large-M throughput configuration): every epilogue flavour, M tail, two-pass fp32 epilogue
    (1300, 2304, 768, 256, 0, False, False), (1300, 3072, 768, 256, 1, False, False), (1300, 768, 3072, 256, 0, True, True),
    (200, 1024, 128, 256, 2, False, False),
    # bn = 512: the CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 pair tile): every epilogue flavour, odd number of
    # M tiles (the pair's second CTA is all out of bounds), M tail, one k-block, K longer than the ring
    (1300, 2304, 768, 512, 0, False, False), (1300, 3072, 768, 512, 1, False, False), (1300, 768, 3072, 512, 0, True, True),
    (513, 768, 768, 512, 0, True, True), (100, 256, 64, 512, 2, False, False), (17696, 2304, 768, 512, 0, False, False),
    (256, 1024, 6912, 512, 2, False, False),
])
def test_gemm(M, N, K, bn, act, resid, f32):
    lib = _cabi.load()
    torch.manual_seed(0)
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda") if resid else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    rptr = None
    if resid:
        out.copy_(R)  # in place, as the engine uses it
        rptr = out.data_ptr()
    ref = A.float() @ W.float().t() + bias
    ref = torch.nn.functional.gelu(ref) if act == 1 else (torch.relu(ref) if act == 2 else ref)
    if resid:
        ref = ref + R
    _cabi.check(lib.uvlt_op_gemm(A.data_ptr(), W.data_ptr(), bias.data_ptr(), rptr, out.data_ptr(), M, N, K, act, int(f32),
                                 bn, None))
    torch.cuda.synchronize()
    assert rel(out.float(), ref) < (2e-5 if f32 else 4e-3)


def test_gemm_rejects_bad_shapes():
    lib = _cabi.load()
    a = torch.zeros(8, 100, device="cuda", dtype=torch.bfloat16)
    assert lib.uvlt_op_gemm(a.data_ptr(), a.data_ptr(), None, None, a.data_ptr(), 8, 8, 100, 0, 0, 0, None) != 0
    assert b"K % 64" in lib.uvlt_last_error() or b"multiple" in lib.uvlt_last_error()


@pytest.mark.parametrize("B,n,H,masked", [
    (2, 361, 12, True), (1, 321, 12, False), (2, 40, 12, True), (2, 553, 12, True), (1, 513, 12, True),
    (1, 128, 12, False), (1, 129, 12, True), (3, 81, 12, True), (1, 1193, 16, False), (2, 681, 16, True), (32, 361, 12, True),
])
def test_attention(B, n, H, masked):
    lib = _cabi.load()
    torch.manual_seed(1)
    D = H * 64
    qkv = torch.randn(B, n, 3 * D, device="cuda").to(torch.bfloat16)
    bias = None
    if masked:
        bias = torch.zeros(B, n, device="cuda")
        bias[0, : min(65, n // 2)] = -1e10               # cls + template keys filled (ViT dialect)
        if B > 1:
            bias[1, n - 7:] = -10000.0                    # BERT additive padding mask
        if B > 2:
            bias[2, :] = -10000.0                         # every key padded (BBOX-mode text): plain softmax
    out = torch.zeros(B, n, D, device="cuda", dtype=torch.bfloat16)
    _cabi.check(lib.uvlt_op_attention(qkv.data_ptr(), bias.data_ptr() if masked else None, out.data_ptr(), B, n, H, None,
                                      0, None))
    torch.cuda.synchronize()
    q, k, v = qkv.float().reshape(B, n, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if masked:
        s = s + bias[:, None, None, :]
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, n, D)
    assert torch.isfinite(out.float()).all()
    assert rel(out.float(), ref) < 6e-3


def test_layernorm_in_place_modes():
    lib = _cabi.load()
    torch.manual_seed(2)
    B, N, D, off, rows = 3, 361, 768, 321, 40
    x = torch.randn(B, N, D, device="cuda") * 2 + 0.3
    g, be = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    a0, a1 = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    # mode 1: add modal embeddings, write back, emit bf16 LN
    x1 = x.clone()
    dst = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
    _cabi.check(lib.uvlt_op_layernorm(x1.data_ptr(), N * D, 0, N, a0.data_ptr(), a1.data_ptr(), 321, 1, dst.data_ptr(),
                                      g.data_ptr(), be.data_ptr(), 1e-6, B, D, None))
    want = x.clone()
    want[:, :321] += a0
    want[:, 321:] += a1
    assert torch.equal(x1, want)
    ref = torch.nn.functional.layer_norm(want, (D,), g, be, 1e-6)
    assert rel(dst.float().reshape(B, N, D), ref) < 3e-3
    # mode 2 on a row range: post-LN replaces the text rows only
    x2 = x.clone()
    dst2 = torch.zeros(B * rows, D, device="cuda", dtype=torch.bfloat16)
    _cabi.check(lib.uvlt_op_layernorm(x2.data_ptr(), N * D, off, rows, None, None, 0, 2, dst2.data_ptr(), g.data_ptr(),
                                      be.data_ptr(), 1e-12, B, D, None))
    ref2 = torch.nn.functional.layer_norm(x[:, off:], (D,), g, be, 1e-12)
    assert torch.equal(x2[:, :off], x[:, :off])
    assert rel(x2[:, off:], ref2) < 1e-5
    assert rel(dst2.float().reshape(B, rows, D), ref2) < 3e-3


def test_patch_im2col_uint8_equals_float():
    lib = _cabi.load()
    rng = np.random.default_rng(0)
    B, Hz, Hx, D = 2, 128, 256, 768
    zu = torch.from_numpy(rng.integers(0, 256, (B, Hz, Hz, 3), dtype=np.uint8)).cuda()
    xu = torch.from_numpy(rng.integers(0, 256, (B, Hx, Hx, 3), dtype=np.uint8)).cuda()
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    zf = ((zu.float().permute(0, 3, 1, 2) / 255.0) - mean) / std
    xf = ((xu.float().permute(0, 3, 1, 2) / 255.0) - mean) / std
    Np = 64 + 256
    cls = torch.randn(D, device="cuda")
    outs = []
    for use_u8 in (False, True):
        out = torch.zeros(B * Np, 768, device="cuda", dtype=torch.bfloat16)
        xs = torch.zeros(B, 1 + Np + 40, D, device="cuda")
        _cabi.check(lib.uvlt_op_patch_im2col(None if use_u8 else zf.contiguous().data_ptr(),
                                             None if use_u8 else xf.contiguous().data_ptr(),
                                             zu.data_ptr() if use_u8 else None, xu.data_ptr() if use_u8 else None,
                                             B, Hz, Hx, out.data_ptr(), cls.data_ptr(), xs.data_ptr(), (1 + Np + 40) * D, D,
                                             None))
        torch.cuda.synchronize()
        assert torch.equal(xs[:, 0], cls.expand(B, D))
        outs.append(out)
    # fp32 reference of the im2col itself
    ref = torch.cat([zf.reshape(B, 3, 8, 16, 8, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, 64, 768),
                     xf.reshape(B, 3, 16, 16, 16, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, 256, 768)], dim=1)
    assert torch.equal(outs[0].float(), ref.to(torch.bfloat16).float().reshape(B * Np, 768))
    assert (outs[0].float() - outs[1].float()).abs().max() <= 2.0 ** -6  # at most one bf16 ulp from fp32 op-order


def test_gemm_cluster_multicast_variant_matches():
    """The optional cluster / TMA-multicast GEMM (UVLT_MULTICAST=1, off by default because it measured slower) must stay
    correct: run the GEMM cases in a subprocess with the variant enabled."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, UVLT_MULTICAST="1")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-m", "gpu", "-k", "test_gemm and not multicast",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("poly", ["0", "1"])
def test_attention_second_generation_kernel_matches(poly):
    """attention2.cuh (P in tensor memory, TMEM-operand PV MMA, two slots per CTA; selected with UVLT_ATTN_V=2) must pass
    the same parity cases as the default kernel, with and without the FMA-pipe exponentials."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, UVLT_ATTN_V="2", UVLT_ATTN_POLY=poly)
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-m", "gpu", "-k", "test_attention and not second and not third",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("grid", ["0", "5"])
def test_attention_third_generation_kernel_matches(grid):
    """attention3.cuh (persistent CTAs over (batch, head, tile pair) work items, two threads per query row; selected with
    UVLT_ATTN_V=3) must pass the same parity cases as the default kernel.  UVLT_ATTN_SPLIT=0 sends every shape to it (by
    default small grids keep the key-split variant of the first kernel); UVLT_ATTN_GRID=5 caps the grid so that every CTA
    walks a long list of items: PAIR -> PAIR, PAIR -> LONE and LONE -> LONE hand-overs, Q / staging buffer reuse, barrier
    phases running across items."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, UVLT_ATTN_V="3", UVLT_ATTN_SPLIT="0", UVLT_ATTN_GRID=grid)
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-m", "gpu", "-k",
                        "test_attention and not second and not third", "-p", "no:cacheprovider"], env=env,
                       capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("M,splits", [(513, 4), (513, 0), (1026, 2), (4104, 0)])
def test_gemm_split_k(M, splits):
    """fc2 at small batch: K is cut in `splits` CTAs per tile; split 0 carries bias + residual, the other splits leave raw
    fp32 partials that the consumer (the engine's next LayerNorm) adds in a fixed order."""
    import ctypes as C

    lib = _cabi.load()
    torch.manual_seed(2)
    N, K = 768, 3072
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda")
    out = R.clone()
    partials = torch.full((5, M, N), float("nan"), device="cuda")
    used = C.c_int(0)
    _cabi.check(lib.uvlt_op_gemm_splitk(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), out.data_ptr(),
                                        partials.data_ptr(), M, N, K, splits, C.byref(used), None))
    torch.cuda.synchronize()
    s = used.value
    assert s == (splits or s) and 1 <= s <= 6
    if splits == 0:
        assert s == (6 if M == 513 else 1)     # the engine's rule: split only when the grid would leave SMs idle
    total = out.clone()
    for i in range(s - 1):
        total = total + partials[i]            # same order as the LayerNorm kernel
    ref = A.float() @ W.float().t() + bias + R
    assert rel(total, ref) < 2e-5
    if s < 6:
        assert torch.isnan(partials[s - 1:]).all()   # untouched
