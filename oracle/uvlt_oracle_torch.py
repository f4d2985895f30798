"""CPU baseline: the same restatement as ``uvlt_oracle.py`` written with PyTorch CPU ops (fp32, eager).

TEST INFRASTRUCTURE, NOT PRODUCT -- same rules as ``uvlt_oracle.py``: only ``tests/`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.

Why it exists: the reference IS eager PyTorch; on a CPU its forward runs through ATen's multithreaded CPU kernels
(oneDNN / MKL GEMMs, vectorised softmax / GELU / LayerNorm, direct convolutions).  The numpy port computes the same
numbers but spends most of its time in single-threaded elementwise passes, which understates the reference's CPU speed by
about 4x.  This file restates the forward_test path op for op with the ``torch.nn.functional`` calls the reference's
modules make, so that the CPU arm of the benchmark is as fast as the reference's own CPU path; it is pinned to the
numpy oracle and to the golden vectors by ``tests/test_oracle_torch.py``.  Only ``forward_test`` (the timed workload) is
restated here; the prompter / training branches live in the numpy oracle.

Every function cites the reference lines it follows (paths relative to the reference checkout).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def to_torch(sd):
    """reference-format state_dict of fp32 ndarrays -> CPU tensors (shared memory, no copy)."""
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def patchify(sd, template, search):
    """PatchEmbed.forward + MaskedAutoencoderViT.patchify (mae_vit.py:80-100, 203-215)."""
    w, b = sd["backbone.vit.patch_embed.proj.weight"], sd["backbone.vit.patch_embed.proj.bias"]
    z = F.conv2d(template, w, b, stride=16).flatten(2).transpose(1, 2) + sd["backbone.vit.pos_embed_z"]
    x = F.conv2d(search, w, b, stride=16).flatten(2).transpose(1, 2) + sd["backbone.vit.pos_embed_x"]
    cls = sd["backbone.vit.cls_token"].expand(x.shape[0], -1, -1)
    return torch.cat([cls, z, x], dim=1)


def bert_embedding(sd, ids, text_mask):
    """BertModel.embedding / BertEmbeddings.forward (bert_backbone.py:740-750, 260-274), eval mode."""
    p = "backbone.bert.embeddings."
    T = ids.shape[1]
    e = F.embedding(ids, sd[p + "word_embeddings.weight"]) + sd[p + "position_embeddings.weight"][:T][None] \
        + sd[p + "token_type_embeddings.weight"][0][None, None]
    e = F.layer_norm(e, e.shape[-1:], sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], 1e-12)
    ext = ((1.0 - text_mask) * -10000.0)[:, None, None, :]
    return e, ext


def cat_mask(nz, nx, text_mask, flag):
    """ModalityUnifiedFeatureExtractor.cat_mask (modality_unified_feature_extractor.py:43-50). True = ignore key."""
    B = flag.shape[0]
    f = flag.reshape(B, 1)
    x_mask = torch.ones(B, nx)
    z_mask = torch.ones(B, nz) * (f != 1)
    c_mask = torch.ones(B, 1) * (f != 1)
    t_mask = text_mask * (f != 0)
    mask = ~torch.cat([c_mask, z_mask, x_mask, t_mask], dim=1).bool()
    visual = ~torch.cat([c_mask, z_mask, x_mask], dim=1).bool()
    return mask, visual


def vit_block(sd, i, x, key_ignore, heads):
    """Block.forward / Attention.forward (block.py:29-32, 47-61): pre-LN eps 1e-6, masked keys FILLED with -1e10."""
    p = f"backbone.vit.blocks.{i}."
    B, n, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, n, 3, heads, C // heads)
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    attn = (q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)
    attn = attn.masked_fill(key_ignore[:, None, None, :], -1e10).softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, n, C)
    x = x + F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def bert_layer(sd, i, t, ext_mask, heads):
    """BertLayer.forward (bert_backbone.py:383-394, 299-325, 335-339, 363-366, 376-380): post-LN, additive mask."""
    p = f"backbone.bert.encoder.layer.{i}."
    B, T, C = t.shape
    dh = C // heads

    def split(y):
        return y.reshape(B, T, heads, dh).permute(0, 2, 1, 3)

    q = split(F.linear(t, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"]))
    k = split(F.linear(t, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"]))
    v = split(F.linear(t, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"]))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh) + ext_mask
    ctx = (s.softmax(dim=-1) @ v).permute(0, 2, 1, 3).reshape(B, T, C)
    a = F.layer_norm(F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]) + t,
                     (C,), sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], 1e-12)
    h = F.gelu(F.linear(a, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    return F.layer_norm(F.linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]) + a, (C,),
                        sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)


def txt_token(txt, text_mask, mode):
    """generate_txt_token (modality_unified_feature_extractor.py:79-83)."""
    if mode == "mean":
        m = text_mask[..., None]
        return (txt * m).sum(1, keepdim=True) / m.sum(1, keepdim=True)
    return txt[:, :1]


def backbone_logits(sd, dims, img, txt, text_mask, flag):
    """contractive_learning (modality_unified_feature_extractor.py:85-93)."""
    vis, x = img[:, :1], img[:, 1 + dims.nz:]
    tt = txt_token(txt, text_mask, dims.txt_token_mode)
    scale = sd["backbone.logit_scale"].exp()
    xn = F.normalize(x, dim=-1)
    vl = scale * (xn @ F.normalize(vis, dim=-1).transpose(-2, -1))
    tl = scale * (xn @ F.normalize(tt, dim=-1).transpose(-2, -1))
    group = torch.stack([vl, tl, (vl + tl) / 2], dim=1)
    return group[torch.arange(flag.shape[0]), flag.reshape(-1)]


def backbone(sd, dims, template, search, ids, text_mask, flag, want_logits=True):
    """ModalityUnifiedFeatureExtractor.forward (modality_unified_feature_extractor.py:52-77), forward_joint
    (mae_vit.py:193-200: the modal-embedding add stays in the residual stream)."""
    H = dims.num_heads
    img = patchify(sd, template, search)
    txt, ext = bert_embedding(sd, ids, text_mask)
    mask, visual = cat_mask(dims.nz, dims.nx, text_mask, flag)
    modal = sd["backbone.vit.modal_embed"]
    nv = dims.n_visual
    logits = []
    for i in range(dims.depth):
        if i in dims.fusion_layers:
            emb = vit_block(sd, i, torch.cat([img + modal[0], txt + modal[1]], dim=1), mask, H)
            img, txt = emb[:, :nv], emb[:, nv:]
        else:
            img = vit_block(sd, i, img, visual, H)
            txt = bert_layer(sd, i, txt, ext, H)
        if want_logits and i in dims.cont_loss_layers:
            logits.append(backbone_logits(sd, dims, img, txt, text_mask, flag))
    out = {"search": img[:, 1 + dims.nz:], "template": img[:, 1:1 + dims.nz], "text": txt, "vis_token": img[:, :1],
           "txt_token": txt_token(txt, text_mask, dims.txt_token_mode), "flag": flag.reshape(-1)}
    if want_logits:
        S = dims.feat_size
        out["logits"] = torch.stack(logits, dim=1).reshape(img.shape[0], -1, S, S)
    return out


def tower(sd, name, x):
    """conv() x 4 + Conv2d 1x1 (heads/utils.py:126-130, modality_adaptive_box_head.py:25-47): Conv3x3(pad 1) +
    BatchNorm2d (eval) + ReLU."""
    for j in range(4):
        p = f"box_head.{name}.{j}."
        x = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1)
        x = F.batch_norm(x, sd[p + "1.running_mean"], sd[p + "1.running_var"], sd[p + "1.weight"], sd[p + "1.bias"],
                         training=False, eps=1e-5)
        x = F.relu(x)
    return F.conv2d(x, sd[f"box_head.{name}.4.weight"], sd[f"box_head.{name}.4.bias"])


def box_head(sd, dims, info, prompt):
    """ModalityAdaptiveBoxHead.forward + contractive_learning (test branch) + convert2bbox
    (modality_adaptive_box_head.py:62-94, 108-119, 140-148)."""
    search, flag = info["search"], info["flag"]
    B, SS, D = search.shape
    S = dims.feat_size
    c = sd["box_head.logit_scale"].exp() * (F.normalize(search, dim=-1) @ F.normalize(prompt, dim=-1).transpose(-2, -1))
    zero = torch.zeros_like(c[:, :, :1])
    if dims.softmax_one:
        mid = torch.cat([c[:, :, 1:], zero], dim=-1).max(-1, keepdim=True)[0]
        cont = torch.cat([c[:, :, :1], mid, zero], dim=-1)
    else:
        cont = torch.cat([c[:, :, :1], c[:, :, 1:].max(-1, keepdim=True)[0]], dim=-1)
    x = search.transpose(-2, -1).reshape(B, D, S, S)
    cls_map = tower(sd, "conv_cls", x).sigmoid()[:, 0]
    off = tower(sd, "conv_offset", x)
    off = off.sigmoid() if dims.offset_sigmoid else off
    size_tr = tower(sd, "conv_bbox", x).sigmoid()
    size_gr = tower(sd, "conv_bbox_grounding", x).sigmoid()
    size = torch.stack([size_tr, size_gr, size_tr], dim=1)[torch.arange(B), flag]
    gy, gx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    coord = torch.stack([gx.reshape(-1), gy.reshape(-1)]).float()[None]
    if not dims.offset_sigmoid:
        coord = coord + 0.5
    score = cls_map.reshape(B, -1) * cont.softmax(-1)[:, :, 0]
    idx = score.argmax(-1)
    ctr = (coord + off.reshape(B, 2, -1)) / S
    bbox_map = torch.cat([ctr, size.reshape(B, 2, -1)], dim=1).transpose(1, 2)
    out = dict(info)
    out.update(cls_score=cls_map, cls_score_test=cls_map, bbox_map=bbox_map,
               pred_boxes=bbox_map[torch.arange(B), idx][:, None], cont_score=cont, prompts=prompt)
    return out


@torch.no_grad()
def forward_test(sd, dims, template, search, ids, text_mask, prompt, flag, want_logits=False):
    """UVLTrack.forward_test (lib/models/uvltrack/uvltrack.py:41-45).  `sd` from to_torch(); inputs are numpy arrays
    (as for the numpy oracle) or CPU tensors; returns a dict of CPU tensors."""
    t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(dt)
    flag_t = t(flag, torch.int64).reshape(-1)
    info = backbone(sd, dims, t(template), t(search), t(ids, torch.int64), t(text_mask), flag_t, want_logits)
    return box_head(sd, dims, info, t(prompt))
