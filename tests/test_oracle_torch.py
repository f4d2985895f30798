"""Pins the PyTorch-CPU restatement used as the benchmark's CPU baseline (oracle/uvlt_oracle_torch.py) to the golden
vectors of the unmodified reference and to the numpy oracle."""
import numpy as np
import pytest

from util import F32_REL_L2, dims_of, golden_cases, load_golden, rel_l2

from oracle import uvlt_oracle as O
from oracle import uvlt_oracle_torch as OT
from uvltrack_b200.weights import synthetic_inputs, synthetic_state_dict


@pytest.mark.parametrize("name", golden_cases(include_large=False))
def test_torch_port_matches_reference_and_numpy_oracle(name):
    g, meta = load_golden(name)
    dims = dims_of(meta)
    sd = synthetic_state_dict(dims, seed=meta["weight_seed"])
    inp = synthetic_inputs(dims, meta["batch"], meta["mode"], seed=meta["input_seed"])
    args = (inp["template"], inp["search"], inp["ids"], inp["text_mask"], inp["prompt"], inp["flag"].reshape(-1))
    out = {k: (v.numpy() if hasattr(v, "numpy") else v)
           for k, v in OT.forward_test(OT.to_torch(sd), dims, *args, want_logits=True).items()}
    for k in ("cls_score_test", "bbox_map", "cont_score", "pred_boxes", "logits", "vis_token"):
        assert rel_l2(out[k], g[k]) < F32_REL_L2, k
    assert rel_l2(out["search"][:, ::8, ::4], g["search_sub"]) < F32_REL_L2
    assert rel_l2(out["template"][:, ::8, ::4], g["template_sub"]) < F32_REL_L2
    ref = O.forward_test(sd, dims, *args, want_logits=True)
    for k in ("cls_score_test", "bbox_map", "cont_score", "search", "template", "logits"):
        assert rel_l2(out[k], ref[k]) < 2e-5, k
    # the tracker merge runs on the host in numpy for both arms
    window = O.hanning_window(dims.feat_size)
    for b in range(meta["batch"]):
        box, score, j = O.track_decode(out["cls_score_test"][b], out["cont_score"][b], out["bbox_map"][b], window)
        assert j == int(g["track"][b, 5]) and np.allclose(box, g["track"][b, :4], atol=1e-5)
