"""Model dimensions and seeded synthetic weights in the reference's ``state_dict`` key names.

No checkpoints, BERT vocabularies or datasets are reachable from this environment, so every parity
test, the smoke test and the benchmark use weights drawn here from ``numpy.random.default_rng(seed)`` -- bit-identical
on every host -- and fed, under the reference key names (``backbone.vit.blocks.3.attn.qkv.weight`` ...,
lib/test/tracker/uvltrack.py:24), both to the oracle / reference modules and to the CUDA engine.

Initialisers follow the reference where it matters for scale (xavier-uniform ViT linears mae_vit.py:137-164, N(0, 0.02)
BERT bert_backbone.py:512-523, default Conv2d) but biases, LayerNorm affine and BatchNorm statistics are randomised
so that bias epilogues and BatchNorm folding are actually exercised, and the last 1x1 conv of the classification
tower is scaled up so the score map has a clear peak (a flat map makes argmax parity meaningless).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class ModelDims:
    """Static shape description of one UVLTrack variant (what ``build_model(cfg)`` derives from the yaml)."""

    arch: str = "base"
    embed_dim: int = 768
    num_heads: int = 12
    depth: int = 12
    mlp_hidden: int = 3072
    template_size: int = 128
    search_size: int = 256
    text_len: int = 40
    fusion_layers: List[int] = field(default_factory=lambda: [6, 7, 8, 9, 10, 11])
    cont_loss_layers: List[int] = field(default_factory=lambda: [3, 4, 5, 6, 7, 8, 9, 10, 11])
    head_channels: int = 256
    vocab_size: int = 30522
    max_position: int = 512
    softmax_one: bool = True
    offset_sigmoid: bool = True
    txt_token_mode: str = "cls"

    @property
    def fusion_start(self) -> int:
        return min(self.fusion_layers)

    @property
    def nz(self) -> int:
        return (self.template_size // 16) ** 2

    @property
    def nx(self) -> int:
        return (self.search_size // 16) ** 2

    @property
    def feat_size(self) -> int:
        return self.search_size // 16

    @property
    def n_visual(self) -> int:
        return 1 + self.nz + self.nx

    @property
    def n_tokens(self) -> int:
        return self.n_visual + self.text_len

    @staticmethod
    def base(template_size=128, search_size=256) -> "ModelDims":
        return ModelDims(template_size=template_size, search_size=search_size)

    @staticmethod
    def large(template_size=128, search_size=256) -> "ModelDims":
        # experiments/uvltrack/baseline_large.yaml:73-76,89: HIDDEN_DIM 1024, FUSION_LAYER 12..23, CONT_LOSS_LAYER 8..23
        return ModelDims(arch="large", embed_dim=1024, num_heads=16, depth=24, mlp_hidden=4096,
                         template_size=template_size, search_size=search_size,
                         fusion_layers=list(range(12, 24)),
                         cont_loss_layers=list(range(8, 24)))

    @staticmethod
    def from_cfg(cfg) -> "ModelDims":
        """From a reference-style cfg tree (lib/config/uvltrack/config.py); 'base'/'large' picked as the reference
        does, by substring of MODEL.BACKBONE.PRETRAINED_PATH (modality_unified_feature_extractor.py:20,30)."""
        path = cfg.MODEL.BACKBONE.PRETRAINED_PATH
        if "base" in path:
            d = ModelDims.base()
        elif "large" in path:
            d = ModelDims.large()
        else:
            raise ValueError(f"cannot infer architecture from PRETRAINED_PATH={path!r}")
        if int(cfg.MODEL.HIDDEN_DIM) != d.embed_dim:
            raise ValueError("MODEL.HIDDEN_DIM does not match the backbone architecture")
        d.template_size = int(cfg.DATA.TEMPLATE.SIZE)
        d.search_size = int(cfg.DATA.SEARCH.SIZE)
        d.text_len = int(cfg.MODEL.BACKBONE.LANGUAGE.BERT.MAX_QUERY_LEN)
        d.fusion_layers = [int(i) for i in cfg.MODEL.BACKBONE.FUSION_LAYER]
        d.cont_loss_layers = [int(i) for i in cfg.MODEL.BACKBONE.CONT_LOSS_LAYER]
        d.head_channels = int(cfg.MODEL.HEAD.HEAD_DIM)
        d.softmax_one = bool(cfg.MODEL.HEAD.SOFTMAX_ONE)
        d.offset_sigmoid = bool(cfg.MODEL.HEAD.OFFSET_SIGMOID)
        d.txt_token_mode = str(cfg.MODEL.BACKBONE.TXT_TOKEN_MODE)
        if bool(cfg.MODEL.HEAD.CLS_TOKENIZE) or bool(cfg.MODEL.HEAD.JOINT_CLS):
            raise NotImplementedError("CLS_TOKENIZE / JOINT_CLS are false in every shipped UVLTrack yaml and are not "
                                      "implemented by the sm_100a engine")
        fl = sorted(d.fusion_layers)
        if fl != list(range(fl[0], d.depth)):
            raise NotImplementedError("FUSION_LAYER must be a suffix range of the block indices")
        return d


# ---------------------------------------------------------------------------------------------------------------
# fixed 2-D sin-cos position embedding (mae_vit.py:33-78): float64 maths, stored fp32
# ---------------------------------------------------------------------------------------------------------------
def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(dim // 2, dtype=np.float64)
    omega /= dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1).astype(np.float64), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed(dim: int, grid_size: int) -> np.ndarray:
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_size, grid_size)  # w goes first
    emb = np.concatenate([_sincos_1d(dim // 2, grid[0]), _sincos_1d(dim // 2, grid[1])], axis=1)
    return emb.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
def synthetic_state_dict(dims: ModelDims, seed: int = 0, cls_sharpen: float = 4.0,
                         reg_gain: float = 2.0) -> "OrderedDict[str, np.ndarray]":
    """Reference-format ``state_dict`` (fp32 numpy) for every parameter/buffer reachable from the hot path.

    ``cls_sharpen``: gain on the last 1x1 conv of the classification tower (a peaked score map).
    ``reg_gain``: gain on the 3x3 convs of the offset / size towers relative to nn.Conv2d's default initialisation (the cls
    tower always gets 2).  The golden files were generated with 2 (x16 over the four layers); 1 = the reference's own
    initialisation scale, used by the +-1 px box tests."""
    rng = np.random.default_rng(seed)
    D, Hd = dims.embed_dim, dims.mlp_hidden
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def normal(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def xavier(out_f, in_f, shape=None):
        a = math.sqrt(6.0 / (in_f + out_f))
        return rng.uniform(-a, a, size=shape or (out_f, in_f)).astype(np.float32)

    def ln(prefix, wname="weight", bname="bias"):
        sd[f"{prefix}.{wname}"] = (1.0 + 0.1 * rng.standard_normal(D)).astype(np.float32)
        sd[f"{prefix}.{bname}"] = normal((D,), 0.05)

    # ---- ViT (mae_vit.py:102-164) ----
    v = "backbone.vit"
    sd["backbone.logit_scale"] = np.array(math.log(1 / 0.07), dtype=np.float32)
    sd[f"{v}.cls_token"] = normal((1, 1, D), 0.02)
    sd[f"{v}.pos_embed_z"] = sincos_pos_embed(D, dims.template_size // 16)[None]
    sd[f"{v}.pos_embed_x"] = sincos_pos_embed(D, dims.search_size // 16)[None]
    sd[f"{v}.modal_embed"] = normal((2, D), 0.02)
    sd[f"{v}.patch_embed.proj.weight"] = xavier(D, 768, (D, 3, 16, 16))
    sd[f"{v}.patch_embed.proj.bias"] = normal((D,), 0.02)
    for i in range(dims.depth):
        b = f"{v}.blocks.{i}"
        ln(f"{b}.norm1")
        sd[f"{b}.attn.qkv.weight"] = xavier(3 * D, D)
        sd[f"{b}.attn.qkv.bias"] = normal((3 * D,), 0.02)
        sd[f"{b}.attn.proj.weight"] = xavier(D, D)
        sd[f"{b}.attn.proj.bias"] = normal((D,), 0.02)
        ln(f"{b}.norm2")
        sd[f"{b}.mlp.fc1.weight"] = xavier(Hd, D)
        sd[f"{b}.mlp.fc1.bias"] = normal((Hd,), 0.02)
        sd[f"{b}.mlp.fc2.weight"] = xavier(D, Hd)
        sd[f"{b}.mlp.fc2.bias"] = normal((D,), 0.02)

    # ---- BERT: embeddings + the first min(FUSION_LAYER) encoder layers (modality_unified_feature_extractor.py:28) ----
    t = "backbone.bert"
    sd[f"{t}.embeddings.word_embeddings.weight"] = normal((dims.vocab_size, D), 0.02)
    sd[f"{t}.embeddings.position_embeddings.weight"] = normal((dims.max_position, D), 0.02)
    sd[f"{t}.embeddings.token_type_embeddings.weight"] = normal((2, D), 0.02)
    ln(f"{t}.embeddings.LayerNorm")
    for i in range(dims.fusion_start):
        b = f"{t}.encoder.layer.{i}"
        for nm in ("query", "key", "value"):
            sd[f"{b}.attention.self.{nm}.weight"] = normal((D, D), 0.02)
            sd[f"{b}.attention.self.{nm}.bias"] = normal((D,), 0.02)
        sd[f"{b}.attention.output.dense.weight"] = normal((D, D), 0.02)
        sd[f"{b}.attention.output.dense.bias"] = normal((D,), 0.02)
        ln(f"{b}.attention.output.LayerNorm")
        sd[f"{b}.intermediate.dense.weight"] = normal((Hd, D), 0.02)
        sd[f"{b}.intermediate.dense.bias"] = normal((Hd,), 0.02)
        sd[f"{b}.output.dense.weight"] = normal((D, Hd), 0.02)
        sd[f"{b}.output.dense.bias"] = normal((D,), 0.02)
        ln(f"{b}.output.LayerNorm")

    # ---- box head (modality_adaptive_box_head.py:27-53, heads/utils.py:126-130) ----
    h = "box_head"
    C = dims.head_channels
    sd[f"{h}.logit_scale"] = np.array(math.log(1 / 0.07), dtype=np.float32)
    chans = [D, C, C // 2, C // 4, C // 8]
    for tower, n_out in (("conv_cls", 1), ("conv_offset", 2), ("conv_bbox", 2), ("conv_bbox_grounding", 2)):
        for j in range(4):
            cin, cout = chans[j], chans[j + 1]
            bound = 1.0 / math.sqrt(cin * 9)
            gain = 2.0 if tower == "conv_cls" else reg_gain
            sd[f"{h}.{tower}.{j}.0.weight"] = rng.uniform(-bound, bound, (cout, cin, 3, 3)).astype(np.float32) * np.float32(gain)
            sd[f"{h}.{tower}.{j}.0.bias"] = rng.uniform(-bound, bound, (cout,)).astype(np.float32)
            sd[f"{h}.{tower}.{j}.1.weight"] = rng.uniform(0.5, 1.5, (cout,)).astype(np.float32)
            sd[f"{h}.{tower}.{j}.1.bias"] = normal((cout,), 0.1) + np.float32(0.1)
            sd[f"{h}.{tower}.{j}.1.running_mean"] = normal((cout,), 0.1)
            sd[f"{h}.{tower}.{j}.1.running_var"] = rng.uniform(0.5, 1.5, (cout,)).astype(np.float32)
        bound = 1.0 / math.sqrt(chans[4])
        w = rng.uniform(-bound, bound, (n_out, chans[4], 1, 1)).astype(np.float32)
        if tower == "conv_cls":
            w *= np.float32(cls_sharpen)
        sd[f"{h}.{tower}.4.weight"] = w
        sd[f"{h}.{tower}.4.bias"] = rng.uniform(-bound, bound, (n_out,)).astype(np.float32)

    # ---- prompter (heads/utils.py:23-43): only the parameters its forward touches ----
    p = f"{h}.prompter"
    sd[f"{p}.logit_scale"] = np.array(math.log(1 / 0.07), dtype=np.float32)
    sd[f"{p}.query_embed.weight"] = normal((3, D), 1.0) * np.float32(0.5)
    sd[f"{p}.mlp.fc1.weight"] = xavier(Hd, D)
    sd[f"{p}.mlp.fc1.bias"] = normal((Hd,), 0.02)
    sd[f"{p}.mlp.fc2.weight"] = xavier(D, Hd)
    sd[f"{p}.mlp.fc2.bias"] = normal((D,), 0.02)
    return sd


def synthetic_inputs(dims: ModelDims, batch: int, mode: str, seed: int = 0):
    """Seeded ``forward_test`` inputs.  mode: 'BBOX' (flag 0), 'NL' (1), 'NLBBOX' (2),
    or 'MIXED' (flags cycle 0,1,2 across the batch).  Returns a dict of numpy arrays."""
    rng = np.random.default_rng(1000 + seed)
    z, x, T, D = dims.template_size, dims.search_size, dims.text_len, dims.embed_dim
    out = {
        "template": rng.standard_normal((batch, 3, z, z), dtype=np.float32),
        "search": rng.standard_normal((batch, 3, x, x), dtype=np.float32),
        "prompt": rng.standard_normal((batch, 3, D), dtype=np.float32),
    }
    flags = {"BBOX": [0], "NL": [1], "NLBBOX": [2], "MIXED": [0, 1, 2]}[mode]
    flag = np.array([flags[i % len(flags)] for i in range(batch)], dtype=np.int64).reshape(batch, 1)
    ids = np.zeros((batch, T), dtype=np.int64)
    mask = np.zeros((batch, T), dtype=np.float32)
    for b in range(batch):
        if flag[b, 0] == 0:
            continue  # tracker BBOX mode feeds zero ids and a zero mask (lib/test/tracker/uvltrack.py:81-85)
        k = int(rng.integers(3, 16))
        ids[b, 0] = 101
        ids[b, 1:1 + k] = rng.integers(1000, 30000, size=k)
        ids[b, 1 + k] = 102
        mask[b, :k + 2] = 1.0
    out.update(ids=ids, text_mask=mask, flag=flag)
    return out
