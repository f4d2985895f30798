"""Evaluation harness around the tracker: the caller side of the hot path.

Mirrors the reference's ``lib/test/evaluation``: ``Sequence`` (data.py:21-150), the ``Tracker`` wrapper with
``run_sequence`` (tracker.py:25-152), ``run_sequence`` / ``run_dataset`` with skip-if-exists resume, per-sequence
exception swallowing and the on-disk result format (running.py:11-172):

    <results_dir>/<dataset>_<TEST.MODE>_<TEST.EPOCH:03d>/<seq>.txt        one "x\\ty\\tw\\th" line per frame, integers
    <results_dir>/<dataset>_<TEST.MODE>_<TEST.EPOCH:03d>/<seq>_time.txt   one "%f" seconds per frame

The reference parallelises over OS processes, one model copy each (running.py:145-172).  The B200-native addition is
``run_dataset_batched``: B sequences advance in lock step through ONE engine call per frame (``BatchTracker``); sequences
of different lengths share a batch (a finished sequence keeps receiving its last frame and its rows are dropped).
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict
from typing import List, Optional, Sequence as Seq

import numpy as np


class Sequence:
    """One video: frame paths (or in-memory uint8 RGB arrays), ground-truth boxes [x, y, w, h], optional language."""

    def __init__(self, name, frames, dataset, ground_truth_rect, language: str = "", object_ids=None, init_extra=None):
        self.init_extra = dict(init_extra or {})  # e.g. {"text_ids": [...]} when no BERT vocabulary file is available
        self.name = name
        self.frames = frames
        self.dataset = dataset
        self.ground_truth_rect = np.asarray(ground_truth_rect, dtype=np.float64).reshape(-1, 4)
        self.language = language
        self.object_ids = object_ids

    def init_info(self) -> dict:
        return dict(self.init_extra, init_bbox=[float(v) for v in self.ground_truth_rect[0]])

    def frame_info(self, frame_num: int) -> dict:
        return {}

    def __len__(self):
        return len(self.frames)


def read_image(frame):
    """lib/test/evaluation/tracker.py:262-273: path -> RGB uint8 (arrays pass through)."""
    if isinstance(frame, np.ndarray):
        return frame
    import cv2

    im = cv2.imread(frame)
    if im is None:
        raise FileNotFoundError(frame)
    return cv2.cvtColor(im, cv2.COLOR_BGR2RGB)


def results_subdir(params, dataset: str) -> str:
    return f"{dataset}_{params.cfg.TEST.MODE}_{int(params.cfg.TEST.EPOCH):03d}"


class Tracker:
    """Wrapper that builds one tracker object per sequence (lib/test/evaluation/tracker.py:25-84).

    ``tracker_class`` is what ``get_tracker_class()`` returns; ``params`` what ``parameters(yaml_name)`` returns."""

    def __init__(self, name: str, parameter_name: str, dataset_name: str, params, results_dir: str, tracker_class=None,
                 run_id=None):
        self.name, self.parameter_name, self.dataset_name, self.run_id = name, parameter_name, dataset_name, run_id
        self.params = params
        self.results_dir = results_dir
        if tracker_class is None:
            from .tracker import get_tracker_class

            tracker_class = get_tracker_class()
        self.tracker_class = tracker_class

    def create_tracker(self, params):
        return self.tracker_class(params, self.dataset_name)

    def run_sequence(self, seq: Sequence, debug=None) -> dict:
        self.params.debug = getattr(self.params, "debug", 0) if debug is None else debug
        tracker = self.create_tracker(self.params)
        return self._track_sequence(tracker, seq, seq.init_info())

    def _track_sequence(self, tracker, seq: Sequence, init_info: dict) -> dict:
        """tracker.py:86-152: output['target_bbox'][i], output['time'][i] for every frame; frame 0 = the init box."""
        output = {"target_bbox": [], "time": []}
        image = read_image(seq.frames[0])
        t0 = time.time()
        init_info = dict(init_info, seq_name=seq.name, language=seq.language)
        tracker.initialize(image, init_info)
        output["target_bbox"].append(init_info["init_bbox"])
        output["time"].append(time.time() - t0)
        prev = OrderedDict()
        for frame_num in range(1, len(seq.frames)):
            image = read_image(seq.frames[frame_num])
            t0 = time.time()
            info = seq.frame_info(frame_num)
            info["previous_output"] = prev
            info["seq_name"] = seq.name
            out = tracker.track(image, info)
            prev = OrderedDict(out)
            output["target_bbox"].append(out["target_bbox"])
            output["time"].append(time.time() - t0)
        return output


def save_tracker_output(results_dir: str, subdir: str, seq_name: str, output: dict) -> None:
    """running.py:11-90 (single-object mode): integer boxes, tab separated; times as %f."""
    base_dir = os.path.join(results_dir, subdir)
    os.makedirs(base_dir, exist_ok=True)
    base = os.path.join(base_dir, seq_name)
    if output.get("target_bbox"):
        np.savetxt(base + ".txt", np.array(output["target_bbox"]).astype(int), delimiter="\t", fmt="%d")
    if output.get("time"):
        np.savetxt(base + "_time.txt", np.array(output["time"]).astype(float), delimiter="\t", fmt="%f")


def results_exist(results_dir: str, subdir: str, seq_name: str) -> bool:
    f = os.path.join(results_dir, subdir, seq_name + ".txt")
    return os.path.isfile(f) and os.path.getsize(f) > 0


def run_sequence(seq: Sequence, tracker: Tracker, debug=False) -> Optional[float]:
    """running.py:93-142.  Returns the FPS (None when skipped or failed)."""
    subdir = results_subdir(tracker.params, seq.dataset)
    if results_exist(tracker.results_dir, subdir, seq.name):
        print("FPS: {}".format(-1))
        return None
    print("Tracker: {} {} {} ,  Sequence: {}".format(tracker.name, tracker.parameter_name, tracker.run_id, seq.name))
    if debug:
        output = tracker.run_sequence(seq, debug=debug)
    else:
        try:
            output = tracker.run_sequence(seq, debug=debug)
        except Exception as e:  # the reference swallows per-sequence failures when not debugging (running.py:124-128)
            print(e)
            return None
    fps = len(output["time"]) / max(sum(output["time"]), 1e-12)
    print("FPS: {}".format(fps))
    save_tracker_output(tracker.results_dir, subdir, seq.name, output)
    return fps


def run_dataset(dataset: Seq[Sequence], trackers: Seq[Tracker], debug=False) -> None:
    """running.py:145-172, sequential mode (the process pool of the reference is replaced by run_dataset_batched)."""
    print("Evaluating {:4d} trackers on {:5d} sequences".format(len(trackers), len(dataset)))
    for seq in dataset:
        for tr in trackers:
            run_sequence(seq, tr, debug=debug)
    print("Done")


def run_dataset_batched(dataset: Seq[Sequence], tracker: Tracker, batch: int, rank: int = 0, world: int = 1) -> dict:
    """All sequences of this rank's shard, `batch` at a time through one BatchTracker.  Same files as run_sequence.
    Returns {seq.name: boxes [T, 4]} for the sequences it ran."""
    from .dp import shard_range
    from .tracker import BatchTracker

    subdir = results_subdir(tracker.params, dataset[0].dataset) if len(dataset) else ""
    lo, hi = shard_range(len(dataset), rank, world)
    todo = [s for s in list(dataset)[lo:hi] if not results_exist(tracker.results_dir, subdir, s.name)]
    todo.sort(key=len, reverse=True)  # similar lengths share a batch
    done = {}
    bt = None

    def run_group(group: List[Sequence]):
        """One batch of sequences through the BatchTracker.  Returns one output dict per sequence, None for a sequence
        that failed (unreadable frame, 'Too small bounding box.'): like the reference's run_sequence (running.py:124-128)
        a failure costs that sequence only -- the lane is frozen and the sequences sharing the batch run to their end."""
        nonlocal bt
        pad = batch - len(group)
        seqs = group + [group[-1]] * pad  # a short last group is padded with a copy whose rows are dropped
        if bt is None:
            bt = BatchTracker(tracker.params, batch=batch)
        t0 = time.time()
        infos = [dict(s.init_info(), language=s.language, seq_name=s.name) for s in seqs]
        bt.initialize([read_image(s.frames[0]) for s in seqs], infos)
        outs = [{"target_bbox": [s.init_info()["init_bbox"]], "time": [time.time() - t0]} for s in group]
        dead = [None] * len(seqs)
        last = [None] * len(seqs)
        next_cache = [None]
        n_max = max(len(s) for s in group)
        for t in range(1, n_max):
            t0 = time.time()
            frames = []
            cached = next_cache[0][1] if next_cache[0] is not None and next_cache[0][0] == t else None
            for b, s in enumerate(seqs):
                try:
                    if cached is not None:
                        # the very array objects handed to the tracker as `next_images` (it recognises them) -- for every
                        # lane, also one that failed in the step before: the step was committed with these frames, and
                        # what a frozen lane is shown does not matter
                        last[b] = cached[b]
                    elif dead[b] is None:
                        last[b] = read_image(s.frames[min(t, len(s) - 1)])
                except Exception as e:  # an unreadable frame ends this sequence only
                    dead[b] = f"{type(e).__name__}: {e}"
                if last[b] is None:
                    last[b] = read_image(s.frames[0])
                frames.append(last[b])
            # the frames of step t+1 are known: let the tracker upload them while step t computes
            nxt = None
            if t + 1 < n_max and all(d is None for d in dead):
                try:
                    nxt = [read_image(s.frames[min(t + 1, len(s) - 1)]) for s in seqs]
                    next_cache[0] = (t + 1, nxt)
                except Exception:
                    nxt = None
            # ... and commit to them: step t+1 is enqueued behind step t before the host waits for step t
            res = bt.track(frames, raise_on_failure=False, next_images=nxt, commit_next=nxt is not None)
            dt = (time.time() - t0) / len(group)
            for b, s in enumerate(group):
                if dead[b] is None and res[b].get("failed"):
                    dead[b] = res[b].get("error", "failed")
                if dead[b] is None and t < len(s):
                    outs[b]["target_bbox"].append(res[b]["target_bbox"])
                    outs[b]["time"].append(dt)
        for b, s in enumerate(group):
            if dead[b] is not None:
                print(f"{s.name}: {dead[b]}")
                outs[b] = None
        return outs

    for i in range(0, len(todo), batch):
        group: List[Sequence] = todo[i:i + batch]
        try:
            outs = run_group(group)
        except Exception as e:  # e.g. a first frame that cannot be read: isolate the culprit by running the group one by one
            print(e)
            outs = []
            for s in group:
                try:
                    outs.extend(run_group([s]))
                except Exception as e1:
                    print(f"{s.name}: {e1}")
                    outs.append(None)
        for s, o in zip(group, outs):
            if o is None:
                continue  # skipped, like a failed sequence of the reference: no result file, the rest of the shard goes on
            save_tracker_output(tracker.results_dir, subdir, s.name, o)
            done[s.name] = np.array(o["target_bbox"], dtype=np.float64)
    return done
