"""Synthetic video sequences for the tracker-level tests and bench.py."""
import numpy as np


def synthetic_sequence(n_frames, seed=0, H=480, W=640, box=(300.0, 200.0, 60.0, 40.0)):
    """Synthetic video: static textured background, a textured rectangle moving on a smooth
    path.  Returns (list of uint8 RGB frames, list of ground-truth [x, y, w, h])."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(0, 255, (H // 8, W // 8, 3), dtype=np.uint8).repeat(8, axis=0).repeat(8, axis=1)
    w, h = int(box[2]), int(box[3])
    tex = rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
    frames, gts = [], []
    for t in range(n_frames):
        x = int(box[0] + 80 * np.sin(t / 25.0) + 0.3 * t)
        y = int(box[1] + 60 * np.sin(t / 17.0))
        x = max(0, min(W - w, x))
        y = max(0, min(H - h, y))
        f = bg.copy()
        f[y:y + h, x:x + w] = tex
        frames.append(f)
        gts.append([float(x), float(y), float(w), float(h)])
    return frames, gts
