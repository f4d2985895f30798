"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Tolerances.  north_star: "outputs match the reference PyTorch forward on identical inputs within 1e-3 fp32 /
# 1e-2 bf16, and box coordinates match to +-1 px".  The CUDA path computes GEMM operands in bf16 with fp32
# accumulation, fp32 residual stream and fp32 LayerNorm / softmax statistics, so the bf16 figure applies:
#   * maps in [0, 1] (cls_score, bbox_map, pred_boxes):        max-abs <= 1e-2
#   * features / logits (unbounded):                            rel-L2 <= 1e-2
BF16_MAX_ABS = 1e-2
BF16_REL_L2 = 1e-2
# oracle (fp32 numpy) vs reference (fp32 torch): different BLAS summation order only
F32_REL_L2 = 1e-4


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def max_abs(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


def golden_cases(include_large=True):
    names = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(("base_", "large_")))
    return [n for n in names if include_large or not n.startswith("large")]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(g["meta"]))
    return g, meta


def dims_of(meta):
    from uvltrack_b200.weights import ModelDims

    mk = ModelDims.base if meta["arch"] == "base" else ModelDims.large
    return mk(meta["template_size"], meta["search_size"])


from uvltrack_b200.synthetic import synthetic_sequence  # noqa: E402,F401  (shared with bench.py)
