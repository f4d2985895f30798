"""Evaluation harness (uvltrack_b200/evaluation.py) against the reference's conventions (lib/test/evaluation/running.py):
result file names and formats, skip-if-exists resume, per-sequence exception swallowing; on the GPU, the batched
scheduler against one-sequence-at-a-time runs."""
import os

import numpy as np
import pytest

from uvltrack_b200 import config
from uvltrack_b200 import evaluation as ev


class _FakeTracker:
    """Moves the box by (+1.6, +0.4) per frame; fails on demand."""

    def __init__(self, params, dataset_name):
        self.params, self.state, self.n = params, None, 0

    def initialize(self, image, info):
        self.state = list(info["init_bbox"])
        assert "seq_name" in info and "language" in info

    def track(self, image, info=None):
        self.n += 1
        if getattr(self.params, "fail_at", None) == self.n:
            raise RuntimeError("boom")
        self.state = [self.state[0] + 1.6, self.state[1] + 0.4, self.state[2], self.state[3]]
        return {"target_bbox": list(self.state)}


def _params(mode="BBOX"):
    cfg = config.baseline_cfg("base", 128, 256, mode=mode)
    return config.parameters(cfg)


def _seq(name, n, dataset="synth"):
    frames = [np.zeros((48, 64, 3), np.uint8) for _ in range(n)]
    return ev.Sequence(name, frames, dataset, [[10.0, 20.0, 30.5, 12.25]] * n, language="a thing")


def test_result_files_format_and_resume(tmp_path, capsys):
    params = _params()
    tr = ev.Tracker("uvltrack", "baseline_base", "synth", params, str(tmp_path), tracker_class=_FakeTracker)
    seq = _seq("seq_a", 5)
    fps = ev.run_sequence(seq, tr)
    assert fps is not None and fps > 0
    sub = f"synth_BBOX_{int(params.cfg.TEST.EPOCH):03d}"
    box_file = tmp_path / sub / "seq_a.txt"
    lines = box_file.read_text().strip().split("\n")
    assert len(lines) == 5
    assert lines[0] == "10\t20\t30\t12"          # the init box, truncated to int, tab separated (running.py:17-19)
    assert lines[2] == "13\t20\t30\t12"          # 10 + 2 * 1.6 = 13.2 -> 13 ; 20 + 0.8 -> 20
    times = (tmp_path / sub / "seq_a_time.txt").read_text().strip().split("\n")
    assert len(times) == 5 and all(len(t.split(".")[1]) == 6 for t in times)   # "%f"
    # second run: results exist -> skipped, file untouched (running.py:104-116)
    mtime = os.path.getmtime(box_file)
    assert ev.run_sequence(seq, tr) is None
    assert "FPS: -1" in capsys.readouterr().out
    assert os.path.getmtime(box_file) == mtime


def test_sequence_failures_are_swallowed_unless_debugging(tmp_path):
    params = _params()
    params.fail_at = 2
    tr = ev.Tracker("uvltrack", "baseline_base", "synth", params, str(tmp_path), tracker_class=_FakeTracker)
    assert ev.run_sequence(_seq("bad", 4), tr, debug=False) is None        # printed, not raised (running.py:124-128)
    assert not (tmp_path / f"synth_BBOX_{int(params.cfg.TEST.EPOCH):03d}" / "bad.txt").exists()
    with pytest.raises(RuntimeError):
        ev.run_sequence(_seq("bad", 4), tr, debug=True)
    ev.run_dataset([_seq("s1", 3), _seq("s2", 2)], [tr])                   # both fail at frame 2 / none -> no crash


class _FakeBatchTracker:
    """BatchTracker stand-in (CPU): sequence 'tiny' fails with the reference's 'Too small bounding box.' at frame 3."""

    def __init__(self, params, batch=1):
        self.B, self.state, self.failed, self.n, self.names = batch, None, None, 0, None
        self.committed = None  # frames promised with commit_next: the next call must pass these very objects

    def initialize(self, images, infos):
        self.state = [list(i["init_bbox"]) for i in infos]
        self.names = [i["seq_name"] for i in infos]
        self.failed = [None] * self.B
        self.n = 0
        self.committed = None

    def track(self, images, raise_on_failure=True, next_images=None, commit_next=False):
        if self.committed is not None:  # the contract of BatchTracker.track(commit_next=True)
            assert len(images) == len(self.committed) and all(a is b for a, b in zip(images, self.committed))
        self.committed = list(next_images) if (commit_next and next_images is not None) else None
        self.n += 1
        out = []
        for b in range(self.B):
            if self.names[b] == "tiny" and self.n >= 3:
                self.failed[b] = "Too small bounding box."
            if self.failed[b]:
                assert not raise_on_failure
                out.append({"target_bbox": self.state[b], "failed": True, "error": self.failed[b]})
                continue
            self.state[b] = [self.state[b][0] + 1.0] + self.state[b][1:]
            out.append({"target_bbox": list(self.state[b])})
        return out


def test_batched_scheduler_isolates_failures(tmp_path, monkeypatch):
    """a failing sequence (tiny box, unreadable frame) must cost that sequence only -- the reference's
    run_sequence swallows the exception per sequence (running.py:124-128); healthy co-batched sequences are saved and the
    rest of the shard goes on."""
    import uvltrack_b200.tracker as trk

    monkeypatch.setattr(trk, "BatchTracker", _FakeBatchTracker)
    params = _params()
    bad_frames = _seq("unreadable", 5)
    bad_frames.frames[2] = str(tmp_path / "does_not_exist.jpg")
    first_bad = _seq("nofirst", 4)
    first_bad.frames[0] = str(tmp_path / "missing0.jpg")
    seqs = [_seq("ok_a", 6), _seq("tiny", 6), bad_frames, _seq("ok_b", 5), first_bad, _seq("ok_c", 3)]
    tr = ev.Tracker("uvltrack", "baseline_base", "synth", params, str(tmp_path))
    done = ev.run_dataset_batched(seqs, tr, batch=2)
    assert sorted(done) == ["ok_a", "ok_b", "ok_c"]
    assert [len(done[k]) for k in ("ok_a", "ok_b", "ok_c")] == [6, 5, 3]
    sub = tmp_path / f"synth_BBOX_{int(params.cfg.TEST.EPOCH):03d}"
    assert sorted(f.name for f in sub.iterdir() if not f.name.endswith("_time.txt")) == ["ok_a.txt", "ok_b.txt", "ok_c.txt"]


@pytest.mark.gpu
def test_batched_scheduler_matches_sequential(tmp_path, monkeypatch):
    # the batched run uses an engine of capacity 2, the sequential one capacity 1: their split-K factors differ, and with
    # random weights a last-bit difference can flip an argmax.  Without split-K every row is bit-identical across engines,
    # which is what lets this test compare the SCHEDULER (batching, padding, per-sequence bookkeeping) exactly.
    monkeypatch.setenv("UVLT_SPLITK", "0")
    from uvltrack_b200.synthetic import synthetic_sequence
    from uvltrack_b200.weights import ModelDims, synthetic_state_dict

    z, x = 128, 256
    params = _params()
    params.state_dict = synthetic_state_dict(ModelDims.base(z, x), seed=0)
    lens = [9, 6, 4]
    seqs = []
    for i, n in enumerate(lens):
        frames, gts = synthetic_sequence(n, seed=50 + i)
        seqs.append(ev.Sequence(f"v{i}", frames, "synth", gts))
    tr_b = ev.Tracker("uvltrack", "baseline_base", "synth", params, str(tmp_path / "batched"))
    done = ev.run_dataset_batched(seqs, tr_b, batch=2)
    assert sorted(done) == ["v0", "v1", "v2"] and [len(done[f"v{i}"]) for i in range(3)] == lens
    tr_s = ev.Tracker("uvltrack", "baseline_base", "synth", params, str(tmp_path / "single"))
    ev.run_dataset(seqs, [tr_s], debug=True)
    sub = f"synth_BBOX_{int(params.cfg.TEST.EPOCH):03d}"
    for i in range(3):
        a = np.loadtxt(tmp_path / "batched" / sub / f"v{i}.txt")
        b = np.loadtxt(tmp_path / "single" / sub / f"v{i}.txt")
        assert a.shape == (lens[i], 4) and np.abs(a - b).max() <= 1     # integer pixels on disk
    # resume: nothing left to do
    assert ev.run_dataset_batched(seqs, tr_b, batch=2) == {}
