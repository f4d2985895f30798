#!/usr/bin/env python
"""Benchmark of the UVLTrack per-frame hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py                                   # N=1: UVLTrack-B 256x256 template / 256x256 search, BBOX mode,
                                                      #      batch 1, synthetic 500-frame sequence (BASELINE configs[1])
    python bench.py --batch 32 --mode NLBBOX          # configs[2]
    torchrun ... bench.py --gpus 8 ...                # one process per GPU, sequences sharded, one all-gather at the end
    python bench.py --impl reference                  # the CPU arm: the unmodified reference's forward_test (oracle/_ref)

A "step" is one tracker frame for every sequence of the per-GPU batch.  One JSON line is printed by rank 0:
  value     frames/s (all GPUs) of forward_test + window merge/argmax with inputs resident in HBM (CUDA events)
  e2e       frames/s through the reference-facing call surface Tracker.track(): H2D of the raw uint8 frames from pinned
            host memory (the clips are page-locked, tracker.pinned_frames; --pageable-frames = ordinary numpy frames
            staged by host threads), device crop/resize + engine + box update, D2H of the [B,10] result rows, prompt updates
  roofline  the GEMM kernel (dominant: ~75% of the step) timed live on this step's shapes vs the bf16 tensor peak;
            roofline_attention is the same for the fused attention kernel
  configs   the other BASELINE.json configs measured in the same run: B=32 NL+BBOX per GPU (configs[2] at N=1,
            configs[4] = 32 sequences per GPU at N>1) and UVLTrack-L 384^2 B=8 (configs[3], N=1 only), each with its
            own value / e2e / e2e_phases / rooflines
  cpu_baseline  the reference's own forward_test (oracle/_ref, eager PyTorch CPU) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASELINE_FPS_3090 = 60.0  # profile_model.py, UVLTrack-B z128/x256, RTX 3090 (README.md:130-131)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="base", choices=["base", "large"])
    ap.add_argument("--template-size", type=int, default=256)
    ap.add_argument("--search-size", type=int, default=256)
    ap.add_argument("--mode", default="BBOX", choices=["BBOX", "NLBBOX"])
    ap.add_argument("--batch", type=int, default=1, help="sequences per GPU")
    ap.add_argument("--cpu-frames", type=int, default=None, help="frames of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs block (B=32 NL+BBOX, UVLTrack-L 384 B=8)")
    ap.add_argument("--pageable-frames", action="store_true",
                    help="e2e: hand track() ordinary (pageable) numpy frames, which it stages into pinned memory with "
                         "host threads; default: the frames live in page-locked memory (uvltrack_b200.tracker.pinned_frames)")
    return ap.parse_args()


def peaks():
    """Measured roofline denominators (driver-written MEASURED_PEAKS.json), else the recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                    "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def workload_name(a):
    return (f"UVLTrack-{'B' if a.arch == 'base' else 'L'} baseline_{a.arch} template {a.template_size}^2 / search "
            f"{a.search_size}^2 / 40-token text, {a.mode} mode, batch={a.batch} per GPU, synthetic sequence")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def gemm_flops_per_frame(d, skip_text):
    """Executed GEMM FLOPs (2*MACs) of one sequence-frame: patch embed + transformer linears + head conv GEMMs."""
    D, Hd, Nv, N, T, F0, L = d.embed_dim, d.mlp_hidden, d.n_visual, d.n_tokens, d.text_len, d.fusion_start, d.depth
    per_tok = 2 * (D * 3 * D + D * D + 2 * D * Hd)
    fl = (d.nz + d.nx) * 2 * 768 * D
    if skip_text:
        fl += L * Nv * per_tok
    else:
        fl += F0 * (Nv + T) * per_tok + (L - F0) * N * per_tok
    C, SS = d.head_channels, d.nx
    fl += 2 * SS * 9 * (D * 4 * C + 4 * (C * C // 2 + (C // 2) * (C // 4) + (C // 4) * (C // 8)))
    return fl


def attn_flops_per_frame(d, skip_text):
    Nv, N, T, F0, L, D = d.n_visual, d.n_tokens, d.text_len, d.fusion_start, d.depth, d.embed_dim
    if skip_text:
        return L * 4 * Nv * Nv * D
    return F0 * (4 * Nv * Nv * D + 4 * T * T * D) + (L - F0) * 4 * N * N * D


def kernel_rooflines(dims, B, skip_text, pk, tag="b1"):
    """Times the two tensor-core kernels on this step's shapes: each distinct (M,N,K) GEMM of a layer and the
    attention launch, replayed from a CUDA graph (device-bound timing, CUDA events on the launching stream)."""
    import torch

    from uvltrack_b200 import _cabi

    lib = _cabi.load()
    D, Hd, H = dims.embed_dim, dims.mlp_hidden, dims.num_heads
    n = dims.n_visual if skip_text else dims.n_tokens
    M = B * n
    dev = "cuda"
    a = torch.randn(M, Hd, device=dev).to(torch.bfloat16)
    w = torch.randn(max(3 * D, Hd), Hd, device=dev).to(torch.bfloat16) * 0.02
    bias = torch.zeros(Hd, device=dev)
    out_b = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16)
    out_f = torch.zeros(M, D, device=dev)
    qkv = torch.randn(B, n, 3 * D, device=dev).to(torch.bfloat16)
    att = torch.empty(B, n, D, device=dev, dtype=torch.bfloat16)
    shapes = [("qkv", 3 * D, D, 0, 0), ("proj", D, D, 0, 1), ("fc1", Hd, D, 1, 0), ("fc2", D, Hd, 0, 1)]
    reps = 20

    def timed(fn):
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
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            s.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    g_time, g_flops, per = 0.0, 0.0, {}
    import ctypes as C
    partials = torch.zeros(5, M, D, device=dev)
    used = C.c_int(1)
    for name, N_, K_, act, f32 in shapes:
        def fn(N_=N_, K_=K_, act=act, f32=f32, name=name):
            if name == "fc2":
                # the engine's own fc2 launch: split-K at small batch (the partials are summed by the next LayerNorm)
                _cabi.check(lib.uvlt_op_gemm_splitk(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out_f.data_ptr(),
                                                    out_f.data_ptr(), partials.data_ptr(), M, N_, K_, 0, C.byref(used),
                                                    _cabi.current_stream()), "uvlt_op_gemm_splitk")
                return
            _cabi.check(lib.uvlt_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out_f.data_ptr() if f32 else None,
                                         out_f.data_ptr() if f32 else out_b.data_ptr(), M, N_, K_, act, f32, 0,
                                         _cabi.current_stream()), "uvlt_op_gemm")
        t = timed(fn)
        fl = 2.0 * M * N_ * K_
        per[name] = {"us": round(t * 1e6, 2), "tflops": round(fl / t / 1e12, 1)}
        if name == "fc2":
            per[name]["splits"] = int(used.value)
        g_time += t
        g_flops += fl

    def afn():
        _cabi.check(lib.uvlt_op_attention(qkv.data_ptr(), None, att.data_ptr(), B, n, H, None, 0, _cabi.current_stream()),
                    "uvlt_op_attention")
    ta = timed(afn)
    fa = 4.0 * B * H * n * n * 64
    peak = pk["bf16_tflops"]
    # DRAM bytes per launch from the committed `ncu --set full` capture of these launches (profiles/r02_<tag>_*_full.md);
    # null when no capture of this shape is committed
    traffic, tsrc = profile_traffic(tag + "_gemm", "gemm_bf16")
    a_traffic, a_tsrc = profile_traffic(tag + "_attn", "attention")
    w_bytes = round((3 * D * D + D * D + 2 * D * Hd) * 2 / 4)
    act_bytes = round((M * D * 2 + M * 3 * D * 2 + M * D * 2 + M * D * 8 + M * D * 2 + M * Hd * 2 + M * Hd * 2 + M * D * 8) / 4)
    roof = {"bound": "tensor", "kernel": "GEMM (qkv+proj+fc1+fc2 of one layer, M=%d)" % M,
            "achieved": round(g_flops / g_time / 1e12, 2), "peak": peak, "unit": "TFLOP/s",
            "frac": round(g_flops / g_time / 1e12 / peak, 4), "traffic": traffic,
            "traffic_note": "mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), %s; algorithmic "
                            "bytes per launch (mean of the four): weights %d + activations %d"
                            % (tsrc or "no committed capture of this shape", w_bytes, act_bytes),
            "peak_source": pk["source"] + " (burst)",
            "per_shape": per, "avg_launch_us": round(g_time / 4 * 1e6, 2)}
    roof_a = {"bound": "tensor", "kernel": "attention kernel (B=%d, H=%d, n=%d)" % (B, H, n),
              "achieved": round(fa / ta / 1e12, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(fa / ta / 1e12 / peak, 4),
              "traffic": a_traffic, "traffic_note": a_tsrc or "no committed capture of this shape",
              "avg_launch_us": round(ta * 1e6, 2)}
    return roof, roof_a


# ---------------------------------------------------------------------------------------------------------------
def profile_traffic(tag, kernel_substr):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches) of the
    kernels whose name contains `kernel_substr`, parsed from the committed `ncu --set full` summary
    profiles/r02_<tag>_full.md (written by tools/ncu_summary.py from the capture of these very launches).  None when the
    file or the kernel is missing -- the bench never carries a literal copied from an old capture."""
    path = os.path.join(ROOT, "profiles", "r02_%s_full.md" % tag)
    if not os.path.exists(path):
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, cur = 0.0, 0, None
    for line in open(path):
        if line.startswith("## "):
            cur = line
            if kernel_substr in cur:
                n += 1
        elif cur and kernel_substr in cur and ("dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line):
            cells = [c.strip() for c in line.strip().strip("|").split("|")]
            try:
                tot += float(cells[1].replace(",", "")) * unit.get(cells[2], 1.0)
            except (ValueError, IndexError):
                return None, None
    return (round(tot / n) if n else None), os.path.relpath(path, ROOT)


def reference_root():
    """Where the UNMODIFIED reference tree can be imported from: /root/reference in the build container, the copy that
    oracle/make_ref.py places under oracle/_ref/ (git-ignored, shipped by gpurun) on the GPU box; None if neither."""
    for r in (os.environ.get("UVLT_REFERENCE_ROOT"), os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        if r and os.path.isdir(os.path.join(r, "lib", "models", "uvltrack")):
            return r
    return None


def cpu_frames_per_second(dims, mode, frames, B=1, warmup=1):
    """The reference's CPU path on the host cores.  Preferred: `UVLTrack.forward_test` of the UNMODIFIED reference
    modules (oracle/_ref, imported through oracle/ref_shim.py; eager PyTorch on the ATen CPU kernels, all host threads)
    + the numpy window merge of Tracker.track -> kind "reference".  Without oracle/_ref: the op-for-op PyTorch-CPU
    restatement (oracle/uvlt_oracle_torch.py) -> kind "port"; if torch's CPU path fails too, the numpy port.
    Returns (frames/s, seconds, description, threads, kind)."""
    from oracle import uvlt_oracle as O
    from uvltrack_b200.weights import synthetic_inputs, synthetic_state_dict

    sd = synthetic_state_dict(dims, seed=0)
    inp = synthetic_inputs(dims, B, mode, seed=0)
    window = O.hanning_window(dims.feat_size)
    args = (inp["template"], inp["search"], inp["ids"], inp["text_mask"], inp["prompt"], inp["flag"].reshape(-1))
    kind = "port"
    impl, threads = "numpy fp32 port (oracle/uvlt_oracle.py), BLAS on all host cores", os.cpu_count()
    fwd = lambda: O.forward_test(sd, dims, *args, want_logits=True)  # noqa: E731
    try:
        import torch

        fwd_torch = None
        root = reference_root()
        if root is not None:
            try:
                os.environ["UVLT_REFERENCE_ROOT"] = root
                from oracle import ref_shim

                ref_shim.REF_ROOT = root
                model, _ = ref_shim.build_reference_model(dims.arch, dims.template_size, dims.search_size, state_dict=sd)
                T = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
                text = ref_shim.nested_tensor(T(inp["ids"]), T(inp["text_mask"]))
                t_in = (T(inp["template"]), T(inp["search"]), text, T(inp["prompt"]), T(inp["flag"]))

                def fwd_torch():
                    with torch.no_grad():
                        out = model.forward_test(*t_in)
                    return {k: out[k].numpy() for k in ("cls_score_test", "cont_score", "bbox_map")}

                kind = "reference"
                impl = "UNMODIFIED reference modules (%s), UVLTrack.forward_test, eager PyTorch CPU fp32" % os.path.relpath(root, ROOT)
            except Exception as e:  # fall through to the restatement
                fwd_torch = None
                impl_err = " [reference import failed: %s: %s]" % (type(e).__name__, str(e)[:80])
            else:
                impl_err = ""
        else:
            impl_err = " [oracle/_ref absent]"
        if fwd_torch is None:
            from oracle import uvlt_oracle_torch as OT

            sdt = OT.to_torch(sd)
            t_args = tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in args)

            def fwd_torch():
                out = OT.forward_test(sdt, dims, *t_args, want_logits=True)
                return {k: out[k].numpy() for k in ("cls_score_test", "cont_score", "bbox_map")}

            impl = "PyTorch-CPU fp32 restatement of the reference forward (oracle/uvlt_oracle_torch.py)" + impl_err

        # torchrun exports OMP_NUM_THREADS=1; this arm is the only CPU work of the job (the other ranks exit), so give
        # it the host: logical CPUs of this process, or half of them (hyper-threads), whichever runs a frame faster
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        best = None
        for n in sorted({ncpu, max(1, ncpu // 2)}, reverse=True):
            torch.set_num_threads(n)
            fwd_torch()
            t0 = time.perf_counter()
            fwd_torch()
            dt1 = time.perf_counter() - t0
            if best is None or dt1 < best[0]:
                best = (dt1, n)
        torch.set_num_threads(best[1])
        fwd = fwd_torch
        threads = torch.get_num_threads()
        impl += ", %d ATen threads" % threads
    except Exception as e:  # the baseline must not take the benchmark down
        kind = "port"
        impl += " [torch CPU path unavailable: %s]" % type(e).__name__

    def one():
        out = fwd()
        for b in range(B):
            O.track_decode(out["cls_score_test"][b], out["cont_score"][b], out["bbox_map"][b], window)

    for _ in range(max(warmup, 1)):  # thread pools, page faults
        one()
    t0 = time.perf_counter()
    for _ in range(frames):
        one()
    dt = time.perf_counter() - t0
    return B * frames / dt, dt, impl, threads, kind


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path (see cpu_frames_per_second), on the repo
    arm's workload, metric, unit, steps and warm-up.  Each step is one frame of the batch-1 workload (bounded sample: the
    CPU arm takes ~0.1-0.3 s per frame; steps above 100 are cut and the cut is stated).  Under torchrun only rank 0 runs."""
    from uvltrack_b200.weights import ModelDims

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims = (ModelDims.base if a.arch == "base" else ModelDims.large)(a.template_size, a.search_size)
    steps = min(a.steps, 100)
    t_start = time.perf_counter()
    fps, dt, impl, threads, kind = cpu_frames_per_second(dims, a.mode, steps, a.batch, warmup=min(a.warmup, 5))
    line = {
        "impl": "reference", "metric": "tracker FPS (frames/sec)", "value": round(fps, 3), "unit": "frames/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(dt / steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": primary_config(a, 1, False),
        "cpu_baseline": {"value": round(fps, 3), "unit": "frames/s", "cores": threads, "kind": kind,
                         "host_cpus": os.cpu_count(),
                         "sample": f"{steps} timed frames ({min(a.warmup, 5)} warm-up) of the same workload on the host of "
                                   f"rank 0, forward_test (incl. the per-layer contrastive logits the reference always "
                                   f"computes) + window merge: {impl}"},
        "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "timed_steps": steps, "wall_s": round(time.perf_counter() - t_start, 1),
        "note": "one CPU process on rank 0's host whatever --gpus says: compare with the N=1 line of the repo arm",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def primary_config(a, world, text_cached):
    """The `config` object of the JSON line: identical keys and values in both arms (--impl b200 / reference)."""
    return {"workload": workload_name(a), "sequences_per_gpu": a.batch, "total_sequences": a.batch * world,
            "model": "UVLTrack-%s baseline_%s" % ("B" if a.arch == "base" else "L", a.arch),
            "template_size": a.template_size, "search_size": a.search_size, "text_len": 40, "mode": a.mode,
            "l2": "not flushed: every step streams the bf16 weight set (273 MB for UVLTrack-B > 126 MB L2) and rotates 4 input frames"}


PAGEABLE_FRAMES = False  # --pageable-frames
SPIN_UP_S = 0.25         # device-loop time before the warm-up steps of each workload (clock ramp, see measure_workload)


def measure_workload(arch, z, x, mode, B, steps, warmup, rank, world, pk, with_rooflines=True, frame_pool=4):
    """One workload on this rank's GPU: device-resident `value` and end-to-end `e2e` through BatchTracker.track().
    Returns a dict of per-rank measurements already reduced over ranks (max time)."""
    import torch
    import torch.distributed as dist

    from uvltrack_b200 import NestedTensor, config, dp
    from uvltrack_b200.synthetic import synthetic_sequence
    from uvltrack_b200.tracker import BatchTracker, pinned_frames
    from uvltrack_b200.weights import ModelDims, synthetic_inputs, synthetic_state_dict

    dims = (ModelDims.base if arch == "base" else ModelDims.large)(z, x)
    cfg = config.baseline_cfg(arch, z, x, mode=mode)
    params = config.parameters(cfg)
    params.state_dict = synthetic_state_dict(dims, seed=0)
    tracker = BatchTracker(params, batch=B)
    eng = tracker.engine
    skip_text = mode == "BBOX"

    # ---- synthetic sequences: `frame_pool` distinct videos per rank shared by the B slots ----
    n_frames = steps + warmup + 1
    pool = [synthetic_sequence(n_frames, seed=rank * 1000 + k) for k in range(min(B, frame_pool))]
    if not PAGEABLE_FRAMES:
        # the decoded clips live in page-locked host memory (what a reader decoding into pinned buffers provides): every
        # step's H2D copy reads the frames where they are, no host thread copies pixels into a staging buffer
        pool = [(pinned_frames(frames), gts) for frames, gts in pool]
    seqs = [pool[b % len(pool)] for b in range(B)]
    infos = []
    rng = np.random.default_rng(7 + rank)
    for b in range(B):
        info = {"init_bbox": seqs[b][1][0]}
        if mode != "BBOX":
            k = int(rng.integers(3, 16))
            info["text_ids"] = [101] + rng.integers(1000, 30000, size=k).tolist() + [102]
        infos.append(info)
    tracker.initialize([s[0][0] for s in seqs], infos)

    # ================= value: device-resident forward_test + merge ================================================
    inp = synthetic_inputs(dims, B, "BBOX" if mode == "BBOX" else "NLBBOX", seed=rank)
    T = lambda v: torch.from_numpy(np.ascontiguousarray(v)).cuda()  # noqa: E731
    ring = [T(np.random.default_rng(100 + i).standard_normal((B, 3, dims.search_size, dims.search_size), dtype=np.float32))
            for i in range(4)]
    tmpl, prompt, flag = T(inp["template"]), T(inp["prompt"]), T(inp["flag"])
    text = NestedTensor(T(inp["ids"]), T(inp["text_mask"]))
    window = tracker.window_dev
    dec_out = torch.empty(B, 6, device="cuda")
    text_cached = not skip_text  # the tracker runs the language branch once per sequence (uvlt_text_encode)
    if text_cached:
        eng.text_encode(text, flag)

    def dev_step(i):
        eng.forward_test(tmpl, ring[i % len(ring)], text, prompt, flag, skip_text=skip_text, clone=False,
                         text_cached=text_cached)
        eng.lib.uvlt_track_decode(eng.h, window.data_ptr(), 1, None, None, dec_out.data_ptr(), None)

    # clock spin-up (not a timed step, reported as config.spin_up_ms): a fresh box idles at its lowest clocks and a few
    # millisecond-long warm-up steps are over before the SM clock has ramped -- the same binary measured 0.65 and 0.69
    # ms per step on consecutive boxes with only the W warm-up steps in front of a 14 ms timed region
    t_spin = time.perf_counter()
    i_spin = 0
    while time.perf_counter() - t_spin < SPIN_UP_S:
        dev_step(i_spin)
        i_spin += 1
        if i_spin % 8 == 0:
            torch.cuda.synchronize()
    for i in range(max(warmup, 3)):
        dev_step(i)
    eng.forward_test(tmpl, ring[0], text, prompt, flag, skip_text=skip_text, clone=False, text_cached=text_cached)
    launches_per_step = eng.last_launch_count + 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        dev_step(i)
    e1.record()
    torch.cuda.synchronize()
    dev_s = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([dev_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t.item())

    # ================= e2e: Tracker.track() from host frames =======================================================
    if text_cached:
        tracker.engine._text_owner = None  # the device-resident loop overwrote the engine's text cache
    frames_at = lambda t: [s[0][t] for s in seqs]  # noqa: E731
    for t_ in range(1, warmup + 1):
        # the last warm-up step neither prefetches nor commits its successor: no copy and no kernel of a timed step starts
        # before the timed region
        tracker.track(frames_at(t_), next_images=frames_at(t_ + 1) if t_ < warmup else None, commit_next=t_ < warmup)
    traj = np.zeros((B, steps, 4), dtype=np.float32)
    traj_dev = torch.zeros(B, steps, 4, device="cuda")
    dp.warmup_gather(traj_dev)  # NCCL builds its communicator lazily: not part of the run
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_launches = 0
    h2d0 = tracker.h2d_bytes
    ph0 = dict(tracker.phase_s)
    t0 = time.perf_counter()
    for i in range(steps):
        # the caller knows its next frames (a video file, a camera queue): they are staged and uploaded while this
        # step computes (double-buffered frame staging); every copy is still inside the timed region
        # ... and it commits to them (commit_next): step i + 1 is enqueued before this call returns, so the GPU does not idle
        # while the host turns the rows of step i into results
        res = tracker.track(frames_at(warmup + 1 + i), next_images=frames_at(warmup + 2 + i) if i + 1 < steps else None,
                            commit_next=True)
        e2e_launches += eng.last_launch_count
        for b in range(B):
            traj[b, i] = res[b]["target_bbox"]
    t_track = time.perf_counter()
    traj_dev.copy_(torch.from_numpy(traj))
    all_traj = dp.gather_trajectories(traj_dev, n_sequences=B * world)  # the ONE collective of the run
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_s = t1 - t0
    per_rank = None
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        # after the timed region: where each rank spent its loop (rank skew shows up as waiting in the final gather)
        mine = {"rank": rank, "track_loop_ms": round((t_track - t0) * 1e3, 3), "gather_wait_ms": round((t1 - t_track) * 1e3, 3),
                **{k: round((tracker.phase_s[k] - ph0[k]) * 1e3, 3) for k in ("stage_h2d", "engine")}}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    assert all_traj.shape[0] == B * world and bool(torch.isfinite(all_traj).all())
    phases = {k: round((tracker.phase_s[k] - ph0[k]) / steps * 1e3, 4) for k in tracker.phase_s}
    phases["python_loop_other"] = round(((t_track - t0) - sum(tracker.phase_s[k] - ph0[k] for k in ph0)) / steps * 1e3, 4)
    phases["final_gather_total_ms"] = round((t1 - t_track) * 1e3, 4)

    out = {
        "workload": f"UVLTrack-{'B' if arch == 'base' else 'L'} baseline_{arch} template {z}^2 / search {x}^2 / 40-token "
                    f"text, {mode} mode, batch={B} per GPU, synthetic sequences",
        "sequences_per_gpu": B, "total_sequences": B * world,
        "value": round(world * B * steps / dev_s, 2), "unit": "frames/s", "steps": steps, "warmup": warmup,
        "ms_per_step": round(dev_s / steps * 1e3, 4),
        "e2e": {"value": round(world * B * steps / e2e_s, 2), "unit": "frames/s",
                "h2d_bytes_per_step": int((tracker.h2d_bytes - h2d0) / steps),
                "frame_bytes_per_step": int(B * np.prod(seqs[0][0][0].shape)),
                "d2h_bytes_per_step": B * 80, "ms_per_step": round(e2e_s / steps * 1e3, 4)},
        "e2e_phases_ms_per_step": phases,
        "e2e_per_rank_totals_ms": per_rank,
        "launches_per_step": int(launches_per_step),
        "gpu_launches": int(launches_per_step * steps + e2e_launches),
        "spin_up_ms": int(SPIN_UP_S * 1e3), "skip_dead_text_branch": skip_text, "text_branch_cached_per_sequence": text_cached,
        "_dims": dims, "_skip_text": skip_text, "_dev_s": dev_s,
    }
    del tracker, eng
    return out


def attach_rooflines(w, B, pk, tag):
    """Live kernel timings of this workload's GEMM / attention shapes + the committed ncu traffic (rank 0 only)."""
    dims, skip_text = w["_dims"], w["_skip_text"]
    roof, roof_a = kernel_rooflines(dims, B, skip_text, pk, tag)
    g_fl, a_fl = gemm_flops_per_frame(dims, skip_text), attn_flops_per_frame(dims, skip_text)
    w["roofline"], w["roofline_attention"] = roof, roof_a
    w["step_model"] = {"gemm_gflop_per_frame": round(g_fl / 1e9, 2), "attention_gflop_per_frame": round(a_fl / 1e9, 2),
                       "achieved_tflops_whole_step": round((g_fl + a_fl) * B / (w["_dev_s"] / w["steps"]) / 1e12, 2)}


def strip_private(w):
    return {k: v for k, v in w.items() if not k.startswith("_")}


def run_b200(a):
    import torch
    import torch.distributed as dist

    from uvltrack_b200 import dp

    rank, world, local = dp.init_process_group()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    else:
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU path")
    pk = peaks()
    clocks = ClockSampler(local)
    clocks.start()

    # ---- primary workload: BASELINE configs[1] per GPU (or what the flags say) ----
    global PAGEABLE_FRAMES
    PAGEABLE_FRAMES = a.pageable_frames
    prim = measure_workload(a.arch, a.template_size, a.search_size, a.mode, a.batch, a.steps, a.warmup, rank, world, pk)

    # ---- the other BASELINE configs, measured in the same run (rank-synchronous: every rank runs them) ----
    extra = {}
    default_primary = (a.arch == "base" and a.batch == 1 and a.mode == "BBOX" and a.template_size == 256 and a.search_size == 256)
    if default_primary and not a.no_configs:
        st = max(5, min(a.steps, 40))
        key = "configs[2]" if world == 1 else "configs[4]"
        extra[key] = measure_workload("base", 256, 256, "NLBBOX", 32, st, min(a.warmup, 5), rank, world, pk)
        extra[key]["baseline_config"] = ("UVLTrack-B 256^2 NL+BBOX, batch=32 sequences, 1xB200" if world == 1 else
                                         "UVLTrack-B 256^2 NL+BBOX, %d independent synthetic sequences sharded across %dxB200 "
                                         "(32/GPU), NCCL box all-gather" % (32 * world, world))
        if world == 1:
            extra["configs[3]"] = measure_workload("large", 384, 384, "NLBBOX", 8, max(5, min(a.steps, 10)), 3, rank, world, pk)
            extra["configs[3]"]["baseline_config"] = "UVLTrack-L baseline_large 384^2 NL+BBOX, batch=8, 1xB200"
    clk = clocks.stop()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    attach_rooflines(prim, a.batch, pk, "b%d" % a.batch)
    for key, w in extra.items():
        attach_rooflines(w, w["sequences_per_gpu"], pk, "b32" if w["sequences_per_gpu"] == 32 else "l8")
    cfgd = primary_config(a, world, prim["text_branch_cached_per_sequence"])  # identical in both arms (driver: same_config)
    impl_notes = {"skip_dead_text_branch": prim["skip_dead_text_branch"],
                  "text_branch_cached_per_sequence": prim["text_branch_cached_per_sequence"],
                  "vs_baseline_note": "value / 60 FPS = the reference's RTX-3090 profile_model.py figure for UVLTrack-B "
                                      "(z128/x256, forward_test only); this workload is the heavier 256/256 shape"}
    line = {
        "metric": "tracker FPS (frames/sec)", "value": prim["value"], "unit": "frames/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": prim["ms_per_step"],
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": round(prim["value"] / world / BASELINE_FPS_3090, 2) if (a.arch == "base" and a.batch == 1) else None,
        "dtype": "bf16", "data": "synthetic", "config": cfgd, "implementation": impl_notes,
        "e2e": dict(prim["e2e"], frames="pageable numpy arrays, staged into pinned memory by host threads" if a.pageable_frames
                    else "page-locked host memory (uvltrack_b200.tracker.pinned_frames), copied by DMA only",
                    path="BatchTracker.track(): raw uint8 frames (480x640x3) in host memory -> whole frames of the NEXT step "
                    "uploaded on a copy stream into the engine's second frame buffer while this step computes (a step "
                    "without a prefetched frame uploads only the search window sample_target reads) -> "
                    "uvlt_track_frame_image_host (device "
                    "crop+resize bit-exact with cv2, forward_test, window merge, map_box_back / clip_box) -> D2H of the "
                    "[B,10] fp64 rows; prompt update every 20 frames; ONE final trajectory all-gather included"),
        "e2e_phases_ms_per_step": prim["e2e_phases_ms_per_step"],
        "e2e_per_rank_totals_ms": prim["e2e_per_rank_totals_ms"],
        "gpu_launches": prim["gpu_launches"] + sum(w["gpu_launches"] for w in extra.values()),
        "launches_per_step": prim["launches_per_step"],
        "spin_up_ms": prim["spin_up_ms"],  # untimed device loop in front of the W warm-up steps of each workload (clock ramp)
        "clocks": clk, "roofline": prim["roofline"], "roofline_attention": prim["roofline_attention"],
        "step_model": prim["step_model"],
        "configs": {k: strip_private(w) for k, w in extra.items()},
    }
    if not a.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only (the other arms: --impl reference)
        frames = a.cpu_frames or (60 if a.arch == "base" else 16)
        dims = prim["_dims"]
        fps, dt, impl, threads, kind = cpu_frames_per_second(dims, a.mode, frames, 1)
        line["cpu_baseline"] = {"value": round(fps, 3), "unit": "frames/s", "cores": threads, "kind": kind,
                                "host_cpus": os.cpu_count(),
                                "sample": f"{frames} frames (batch 1) of the same workload in {dt:.1f} s, forward_test + "
                                          f"window merge: {impl}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
