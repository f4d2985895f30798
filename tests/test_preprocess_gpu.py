"""Device-side sample_target / box update (csrc/track.cuh) against the host path the reference uses
(OpenCV crop + cv2.resize INTER_LINEAR, lib/train/data/processing_utils.py:159-243): the uint8 crop must be identical
bit for bit, and the tracker must produce the same boxes through either path."""
import ctypes as C

import numpy as np
import pytest
import torch

from uvltrack_b200 import _cabi, config, preprocess as pp
from uvltrack_b200.synthetic import synthetic_sequence
from uvltrack_b200.tracker import BatchTracker
from uvltrack_b200.weights import ModelDims, synthetic_state_dict

pytestmark = pytest.mark.gpu

BOXES = [
    (300.0, 200.0, 60.0, 40.0),      # interior, upscale-free (crop 196 -> 256: upscale)
    (5.0, 3.0, 50.0, 70.0),          # crosses the top-left border
    (600.0, 440.0, 60.0, 50.0),      # crosses the bottom-right border
    (200.0, 100.0, 300.0, 250.0),    # crop larger than the frame on every side (downscale)
    (310.5, 220.25, 64.0, 64.0),     # crop side exactly 256 (identity resize), half-integer centre
    (100.3, 50.7, 10.0, 10.0),       # tiny box: 40-pixel crop, strong upscale
    (0.0, 0.0, 640.0, 480.0),        # whole frame
    (639.0, 479.0, 10.0, 10.0),      # mostly outside
    (123.456, 234.567, 33.3, 77.7),
]


@pytest.mark.parametrize("factor,out", [(4.0, 256), (2.0, 128), (4.0, 384)])
def test_crop_resize_bit_exact_vs_opencv(factor, out):
    lib = _cabi.load()
    rng = np.random.default_rng(3)
    H, W = 480, 640
    B = len(BOXES)
    frames = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    state = np.array(BOXES, dtype=np.float64)
    d_frames, d_state = torch.from_numpy(frames).cuda(), torch.from_numpy(state).cuda()
    d_crops = torch.zeros(B, out, out, 3, dtype=torch.uint8, device="cuda")
    d_rf = torch.zeros(B, dtype=torch.float64, device="cuda")
    _cabi.check(lib.uvlt_op_crop_resize(d_frames.data_ptr(), H, W, d_state.data_ptr(), factor, out, d_crops.data_ptr(),
                                        d_rf.data_ptr(), B, None))
    torch.cuda.synchronize()
    got, rf = d_crops.cpu().numpy(), d_rf.cpu().numpy()
    for b in range(B):
        ref, ref_rf, _ = pp.sample_target(frames[b], BOXES[b], factor, out)
        assert rf[b] == ref_rf, (b, rf[b], ref_rf)
        diff = np.abs(got[b].astype(int) - ref.astype(int))
        assert diff.max() == 0, (b, BOXES[b], int(diff.max()), int((diff > 0).sum()))


def test_crop_rejects_too_small_box():
    lib = _cabi.load()
    frames = torch.zeros(1, 64, 64, 3, dtype=torch.uint8, device="cuda")
    state = torch.tensor([[10.0, 10.0, 0.0, 5.0]], dtype=torch.float64, device="cuda")
    crops = torch.ones(1, 32, 32, 3, dtype=torch.uint8, device="cuda")
    rf = torch.ones(1, dtype=torch.float64, device="cuda")
    _cabi.check(lib.uvlt_op_crop_resize(frames.data_ptr(), 64, 64, state.data_ptr(), 4.0, 32, crops.data_ptr(),
                                        rf.data_ptr(), 1, None))
    torch.cuda.synchronize()
    assert float(rf[0]) == 0.0 and int(crops.max()) == 0  # the host wrapper raises "Too small bounding box."


@pytest.mark.parametrize("mode", ["BBOX", "NLBBOX"])
def test_device_and_host_preprocessing_give_the_same_boxes(mode):
    z, x, n = 128, 256, 30
    dims = ModelDims.base(z, x)
    cfg = config.baseline_cfg("base", z, x, mode=mode)
    cfg.TEST.UPDATE_INTERVAL = 10
    cfg.TEST.THRESHOLD = 0.05
    params = config.parameters(cfg)
    params.state_dict = synthetic_state_dict(dims, seed=0)
    B = 2
    seqs = [synthetic_sequence(n + 1, seed=30 + b, box=(20.0 + 500 * b, 30.0 + 300 * b, 60.0, 40.0)) for b in range(B)]
    infos = [{"init_bbox": s[1][0], "text_ids": [101, 2023, 3899, 102]} for s in seqs]
    dev = BatchTracker(params, batch=B, device_preprocess=True)
    host = BatchTracker(params, batch=B, network=dev.network, device_preprocess=False)
    dev.initialize([s[0][0] for s in seqs], infos)
    host.initialize([s[0][0] for s in seqs], infos)
    for t in range(1, n + 1):
        images = [s[0][t] for s in seqs]
        rd = dev.track(images)
        rh = host.track(images)
        for b in range(B):
            # same crop bytes -> same network outputs -> same fp64 box arithmetic: identical, not merely close
            assert rd[b]["target_bbox"] == rh[b]["target_bbox"], (t, b, rd[b], rh[b])
            assert rd[b]["score"] == rh[b]["score"]


def test_window_upload_never_reads_stale_pixels():
    """track() uploads only the search window of each frame into the device staging buffer; every other pixel there is
    stale (earlier frames).  With independent noise frames any stale or missing pixel inside the crop changes the crop,
    hence the box: the device path must still equal the host path (full-frame OpenCV crop) exactly, also with windows
    that hang over the frame border."""
    z, x, n = 128, 256, 14
    dims = ModelDims.base(z, x)
    cfg = config.baseline_cfg("base", z, x, mode="BBOX")
    params = config.parameters(cfg)
    params.state_dict = synthetic_state_dict(dims, seed=0)
    B = 3
    rng = np.random.default_rng(7)
    boxes = [(5.0, 8.0, 50.0, 70.0), (560.0, 400.0, 70.0, 60.0), (300.0, 200.0, 90.0, 30.0)]
    frames0 = [rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8) for _ in range(B)]
    infos = [{"init_bbox": list(bx)} for bx in boxes]
    dev = BatchTracker(params, batch=B, device_preprocess=True)
    host = BatchTracker(params, batch=B, network=dev.network, device_preprocess=False)
    dev.initialize(frames0, infos)
    host.initialize(frames0, infos)
    uploaded = 0
    for t in range(1, n + 1):
        images = [rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8) for _ in range(B)]
        before = dev.h2d_bytes
        rd = dev.track(images)
        uploaded += dev.h2d_bytes - before
        rh = host.track(images)
        for b in range(B):
            assert rd[b]["target_bbox"] == rh[b]["target_bbox"], (t, b, rd[b], rh[b])
    assert 0 < uploaded < n * B * 480 * 640 * 3  # windows, not whole frames


def test_prefetched_frames_give_identical_tracks():
    """track(images, next_images=...) stages and uploads the next step's frames into the engine's second staging buffer
    while the current step runs; the boxes must be identical to the plain one-frame-at-a-time path, including when the
    caller breaks its promise and passes different frames than it announced."""
    z, x, B, n = 128, 256, 2, 8
    dims = ModelDims.base(z, x)
    cfg = config.baseline_cfg("base", z, x, mode="BBOX")
    params = config.parameters(cfg)
    params.state_dict = synthetic_state_dict(dims, seed=0)
    seqs = [synthetic_sequence(n + 1, seed=70 + b) for b in range(B)]
    infos = [{"init_bbox": s[1][0]} for s in seqs]
    tracks = {}
    from uvltrack_b200.tracker import PinnedFrame, pinned_frames

    pinned = [pinned_frames(s[0]) for s in seqs]  # the same clips in page-locked memory: uploaded without staging
    assert all(isinstance(f, PinnedFrame) and np.array_equal(f, g) for fs, s in zip(pinned, seqs) for f, g in zip(fs, s[0]))
    for mode in ("plain", "prefetch", "broken_promise", "pinned_plain", "pinned_prefetch", "committed", "pinned_committed"):
        bt = BatchTracker(params, batch=B)
        bt.initialize([s[0][0] for s in seqs], infos)
        out = []
        src = pinned if mode.startswith("pinned") else [s[0] for s in seqs]
        commit = mode.endswith("committed")  # the next step is enqueued before track() returns
        for t in range(1, n + 1):
            cur = [f[t] for f in src]
            nxt = [f[t + 1] for f in src] if t < n else None
            if mode in ("plain", "pinned_plain"):
                res = bt.track(cur)
            elif mode in ("prefetch", "pinned_prefetch") or commit:
                res = bt.track(cur, next_images=nxt, commit_next=commit)
            else:  # announce frame t+1 of the OTHER sequence order: the tracker must notice and stage `cur` itself
                res = bt.track(cur, next_images=nxt[::-1] if nxt else None)
            out.append([r["target_bbox"] for r in res])
        tracks[mode] = np.array(out)
        bt.engine.close()
    assert np.array_equal(tracks["plain"], tracks["prefetch"])
    assert np.array_equal(tracks["plain"], tracks["broken_promise"])
    assert np.array_equal(tracks["plain"], tracks["pinned_plain"])
    assert np.array_equal(tracks["plain"], tracks["pinned_prefetch"])
    assert np.array_equal(tracks["plain"], tracks["committed"])
    assert np.array_equal(tracks["plain"], tracks["pinned_committed"])


def test_committed_next_frames_must_be_honoured():
    """track(..., commit_next=True) has already advanced the device state with the announced frames: a next call with
    other frames must fail loudly, and initialize() must drop the pending step."""
    z, x, B = 128, 256, 2
    dims = ModelDims.base(z, x)
    cfg = config.baseline_cfg("base", z, x, mode="BBOX")
    params = config.parameters(cfg)
    params.state_dict = synthetic_state_dict(dims, seed=0)
    seqs = [synthetic_sequence(4, seed=80 + b) for b in range(B)]
    infos = [{"init_bbox": s[1][0]} for s in seqs]
    bt = BatchTracker(params, batch=B)
    bt.initialize([s[0][0] for s in seqs], infos)
    bt.track([s[0][1] for s in seqs], next_images=[s[0][2] for s in seqs], commit_next=True)
    with pytest.raises(ValueError):
        bt.track([s[0][3] for s in seqs])
    bt.initialize([s[0][0] for s in seqs], infos)
    a = bt.track([s[0][1] for s in seqs], next_images=[s[0][2] for s in seqs], commit_next=True)
    bt.initialize([s[0][0] for s in seqs], infos)          # drops the committed step
    b = bt.track([s[0][1] for s in seqs])
    assert [r["target_bbox"] for r in a] == [r["target_bbox"] for r in b]
    bt.engine.close()
