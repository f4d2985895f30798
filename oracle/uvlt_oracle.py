"""CPU oracle: a numpy (fp32) restatement of UVLTrack's per-frame forward hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module, and only as the checker / CPU baseline.  The product
(``uvltrack_b200``) never routes through it: without the CUDA library it fails loudly.

Parity status: the reference ships NO golden vectors or tests for this path, so the oracle is pinned
against the reference itself instead -- ``oracle/make_golden.py`` imports the unmodified reference modules from
/root/reference (through ``oracle/ref_shim.py``), runs them on seeded inputs/weights and stores their outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against those vectors.

Every function cites the reference lines it follows (paths relative to the reference checkout).  Weights arrive as the
reference ``state_dict`` (key -> fp32 ndarray).  All arithmetic is fp32 like the reference's eager PyTorch path, except
where the reference itself promotes to float64 (the tracker's window merge).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf as _erf

F32 = np.float32


# ---------------------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------------------
def linear(x, w, b=None):
    """nn.Linear: x @ w.T + b."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y.astype(F32, copy=False)


def gelu(x):
    """exact erf GELU (nn.GELU default, lib/models/backbones/utils.py:50; bert_backbone.py:118-124)."""
    return (x * F32(0.5) * (F32(1.0) + _erf(x / F32(math.sqrt(2.0))).astype(F32))).astype(F32)


def layer_norm(x, g, b, eps):
    """nn.LayerNorm / BertLayerNorm (bert_backbone.py:240-244): biased variance, eps inside the sqrt."""
    u = x.mean(-1, keepdims=True, dtype=F32)
    d = x - u
    s = (d * d).mean(-1, keepdims=True, dtype=F32)
    return (g * (d / np.sqrt(s + F32(eps))) + b).astype(F32)


def softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def l2_normalize(x, eps=1e-12):
    """F.normalize(dim=-1): x / max(||x||, eps)."""
    n = np.sqrt((x * x).sum(-1, keepdims=True, dtype=F32))
    return (x / np.maximum(n, F32(eps))).astype(F32)


# ---------------------------------------------------------------------------------------------------------------
# backbone
# ---------------------------------------------------------------------------------------------------------------
def patch_embed(sd, img):
    """PatchEmbed.forward (mae_vit.py:95-100): Conv2d(3, D, 16, stride 16) -> flatten(2).transpose(1, 2)."""
    w = sd["backbone.vit.patch_embed.proj.weight"]
    b = sd["backbone.vit.patch_embed.proj.bias"]
    B, C, H, W = img.shape
    g = H // 16
    cols = img.reshape(B, C, g, 16, W // 16, 16).transpose(0, 2, 4, 1, 3, 5).reshape(B, g * (W // 16), C * 256)
    return linear(cols.astype(F32), w.reshape(w.shape[0], -1), b)


def patchify(sd, template, search):
    """MaskedAutoencoderViT.patchify (mae_vit.py:203-215): [cls | z + pos_z | x + pos_x]."""
    z = patch_embed(sd, template) + sd["backbone.vit.pos_embed_z"]
    x = patch_embed(sd, search) + sd["backbone.vit.pos_embed_x"]
    cls = np.broadcast_to(sd["backbone.vit.cls_token"], (x.shape[0], 1, x.shape[-1]))
    return np.concatenate([cls, z, x], axis=1).astype(F32)


def bert_embedding(sd, ids, text_mask):
    """BertModel.embedding + BertEmbeddings.forward (bert_backbone.py:740-750, 260-274); eval mode (dropout off)."""
    p = "backbone.bert.embeddings."
    T = ids.shape[1]
    e = sd[p + "word_embeddings.weight"][ids] + sd[p + "position_embeddings.weight"][:T][None] \
        + sd[p + "token_type_embeddings.weight"][0][None, None]
    e = layer_norm(e.astype(F32), sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], 1e-12)
    ext = ((F32(1.0) - text_mask.astype(F32)) * F32(-10000.0))[:, None, None, :]
    return e, ext.astype(F32)


def cat_mask(nz, nx, text_mask, flag):
    """ModalityUnifiedFeatureExtractor.cat_mask (modality_unified_feature_extractor.py:43-50). True = ignore key."""
    B = flag.shape[0]
    f = flag.reshape(B, 1)
    x_mask = np.ones((B, nx), F32)
    z_mask = np.ones((B, nz), F32) * (f != 1)
    c_mask = np.ones((B, 1), F32) * (f != 1)
    t_mask = text_mask.astype(F32) * (f != 0)
    mask = ~np.concatenate([c_mask, z_mask, x_mask, t_mask], axis=1).astype(bool)
    visual = ~np.concatenate([c_mask, z_mask, x_mask], axis=1).astype(bool)
    return mask, visual


def vit_attention(sd, pre, x, key_ignore, heads):
    """Attention.forward (block.py:47-61): masked keys are FILLED with -1e10 (not added)."""
    B, n, C = x.shape
    qkv = linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, n, 3, heads, C // heads)
    q, k, v = (qkv[:, :, i].transpose(0, 2, 1, 3) for i in range(3))
    attn = (q @ k.transpose(0, 1, 3, 2)) * F32((C // heads) ** -0.5)
    if key_ignore is not None:
        attn = np.where(key_ignore[:, None, None, :], F32(-1e10), attn)
    attn = softmax(attn.astype(F32))
    o = (attn @ v).transpose(0, 2, 1, 3).reshape(B, n, C)
    return linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def vit_block(sd, i, x, key_ignore, heads):
    """Block.forward (block.py:29-32): pre-LN (eps 1e-6), LayerScale/DropPath are identities at eval."""
    p = f"backbone.vit.blocks.{i}."
    x = x + vit_attention(sd, p + "attn.", layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6),
                          key_ignore, heads)
    h = gelu(linear(layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6),
                    sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return (x + linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])).astype(F32)


def bert_layer(sd, i, t, ext_mask, heads):
    """BertLayer.forward (bert_backbone.py:383-394, 299-325, 335-339, 363-366, 376-380): post-LN, additive mask."""
    p = f"backbone.bert.encoder.layer.{i}."
    B, T, C = t.shape
    dh = C // heads

    def split(y):
        return y.reshape(B, T, heads, dh).transpose(0, 2, 1, 3)

    q = split(linear(t, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"]))
    k = split(linear(t, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"]))
    v = split(linear(t, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"]))
    s = (q @ k.transpose(0, 1, 3, 2)) / F32(math.sqrt(dh)) + ext_mask
    ctx = (softmax(s.astype(F32)) @ v).transpose(0, 2, 1, 3).reshape(B, T, C)
    a = layer_norm(linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]) + t,
                   sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], 1e-12)
    h = gelu(linear(a, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    return layer_norm(linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]) + a,
                      sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)


def txt_token(txt, text_mask, mode):
    """generate_txt_token (modality_unified_feature_extractor.py:79-83)."""
    if mode == "mean":
        m = text_mask.astype(F32)[..., None]
        return ((txt * m).sum(1, keepdims=True) / m.sum(1, keepdims=True)).astype(F32)
    return txt[:, :1]


def backbone_logits(sd, dims, img, txt, text_mask, flag):
    """contractive_learning (modality_unified_feature_extractor.py:85-93)."""
    nz, nx = dims.nz, dims.nx
    vis, x = img[:, :1], img[:, 1 + nz:]
    tt = txt_token(txt, text_mask, dims.txt_token_mode)
    scale = np.exp(sd["backbone.logit_scale"].astype(F32))
    xn = l2_normalize(x)
    vl = scale * (xn @ l2_normalize(vis).transpose(0, 2, 1))
    tl = scale * (xn @ l2_normalize(tt).transpose(0, 2, 1))
    group = np.stack([vl, tl, (vl + tl) / F32(2)], axis=1)
    return group[np.arange(flag.shape[0]), flag.reshape(-1)].astype(F32)


def backbone(sd, dims, template, search, ids, text_mask, flag, want_logits=True):
    """ModalityUnifiedFeatureExtractor.forward (modality_unified_feature_extractor.py:52-77)."""
    H = dims.num_heads
    img = patchify(sd, template.astype(F32), search.astype(F32))
    txt, ext = bert_embedding(sd, ids, text_mask)
    mask, visual = cat_mask(dims.nz, dims.nx, text_mask, flag)
    modal = sd["backbone.vit.modal_embed"]
    nv = dims.n_visual
    logits = []
    for i in range(dims.depth):
        if i in dims.fusion_layers:
            # forward_joint (mae_vit.py:193-200): the modal-embedding add stays in the residual stream
            emb = np.concatenate([img + modal[0], txt + modal[1]], axis=1).astype(F32)
            emb = vit_block(sd, i, emb, mask, H)
            img, txt = emb[:, :nv], emb[:, nv:]
        else:
            img = vit_block(sd, i, img, visual, H)
            txt = bert_layer(sd, i, txt, ext, H)
        if want_logits and i in dims.cont_loss_layers:
            logits.append(backbone_logits(sd, dims, img, txt, text_mask, flag))
    out = {
        "search": img[:, 1 + dims.nz:], "template": img[:, 1:1 + dims.nz], "text": txt, "vis_token": img[:, :1],
        "txt_token": txt_token(txt, text_mask, dims.txt_token_mode), "flag": flag.reshape(-1),
        "tokens": np.concatenate([img, txt], axis=1),
    }
    if want_logits:
        S = dims.feat_size
        out["logits"] = np.stack(logits, axis=1).reshape(img.shape[0], -1, S, S)
    return out


# ---------------------------------------------------------------------------------------------------------------
# box head
# ---------------------------------------------------------------------------------------------------------------
def conv3x3_bn_relu(sd, pre, x):
    """conv() of heads/utils.py:126-130: Conv2d(3x3, pad 1) + BatchNorm2d (eval, eps 1e-5) + ReLU.  x: [B,C,S,S]."""
    w, b = sd[pre + "0.weight"], sd[pre + "0.bias"]
    B, C, S, _ = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    cols = np.stack([xp[:, :, ky:ky + S, kx:kx + S] for ky in range(3) for kx in range(3)], axis=2)  # B,C,9,S,S
    cols = cols.transpose(0, 3, 4, 1, 2).reshape(B * S * S, C * 9)
    y = cols @ w.reshape(w.shape[0], -1).T + b
    g, be = sd[pre + "1.weight"], sd[pre + "1.bias"]
    mu, var = sd[pre + "1.running_mean"], sd[pre + "1.running_var"]
    y = (y - mu) / np.sqrt(var + F32(1e-5)) * g + be
    y = np.maximum(y, 0).astype(F32)
    return y.reshape(B, S, S, -1).transpose(0, 3, 1, 2)


def tower(sd, name, x):
    for j in range(4):
        x = conv3x3_bn_relu(sd, f"box_head.{name}.{j}.", x)
    w, b = sd[f"box_head.{name}.4.weight"], sd[f"box_head.{name}.4.bias"]
    B, C, S, _ = x.shape
    y = x.transpose(0, 2, 3, 1).reshape(-1, C) @ w.reshape(w.shape[0], C).T + b
    return y.reshape(B, S, S, -1).transpose(0, 3, 1, 2).astype(F32)


def cont_score_test(sd, dims, search, prompt):
    """ModalityAdaptiveBoxHead.contractive_learning, test branch (modality_adaptive_box_head.py:140-148)."""
    scale = np.exp(sd["box_head.logit_scale"].astype(F32))
    c = scale * (l2_normalize(search) @ l2_normalize(prompt).transpose(0, 2, 1))
    c = c.astype(F32)
    zero = np.zeros_like(c[:, :, :1])
    if dims.softmax_one:
        mid = np.concatenate([c[:, :, 1:], zero], axis=-1).max(-1, keepdims=True)
        return np.concatenate([c[:, :, :1], mid, zero], axis=-1)
    return np.concatenate([c[:, :, :1], c[:, :, 1:].max(-1, keepdims=True)], axis=-1)


def cont_score_train(sd, dims, search, prompt):
    """training branch (modality_adaptive_box_head.py:132-137): two columns."""
    scale = np.exp(sd["box_head.logit_scale"].astype(F32))
    c = (scale * (l2_normalize(search) @ l2_normalize(prompt).transpose(0, 2, 1))).astype(F32)
    if dims.softmax_one:
        zero = np.zeros_like(c[:, :, :1])
        mid = np.concatenate([c[:, :, 1:], zero], axis=-1).max(-1, keepdims=True)
    else:
        mid = c[:, :, 1:].max(-1, keepdims=True)
    return np.concatenate([c[:, :, :1], mid], axis=-1)


def box_head(sd, dims, info, prompt=None, cont_score=None):
    """ModalityAdaptiveBoxHead.forward + convert2bbox (modality_adaptive_box_head.py:62-94, 108-119).
    CLS_TOKENIZE / JOINT_CLS are false in every shipped yaml."""
    search, flag = info["search"], info["flag"]
    B, SS, D = search.shape
    S = dims.feat_size
    if cont_score is None:
        cont_score = cont_score_test(sd, dims, search, prompt)
    x = search.transpose(0, 2, 1).reshape(B, D, S, S)
    cls_map = sigmoid(tower(sd, "conv_cls", x))[:, 0]
    off = tower(sd, "conv_offset", x)
    off = sigmoid(off) if dims.offset_sigmoid else off
    size_tr = sigmoid(tower(sd, "conv_bbox", x))
    size_gr = sigmoid(tower(sd, "conv_bbox_grounding", x))
    size = np.stack([size_tr, size_gr, size_tr], axis=1)[np.arange(B), flag]
    # coodinate buffer (:54-59): channel 0 = column index, channel 1 = row index (+0.5 without offset sigmoid)
    gy, gx = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    coord = np.stack([gx.reshape(-1), gy.reshape(-1)]).astype(F32)[None]
    if not dims.offset_sigmoid:
        coord = coord + F32(0.5)
    prob0 = softmax(cont_score)[:, :, 0]
    score = cls_map.reshape(B, -1) * prob0
    idx = score.argmax(-1)
    ctr = (coord + off.reshape(B, 2, -1)) / F32(S)
    bbox_map = np.concatenate([ctr, size.reshape(B, 2, -1)], axis=1).transpose(0, 2, 1).astype(F32)
    out = dict(info)
    out.update(cls_score=cls_map, cls_score_test=cls_map, bbox_map=bbox_map,
               pred_boxes=bbox_map[np.arange(B), idx][:, None], cont_score=cont_score.astype(F32), prompts=prompt)
    return out


def forward_test(sd, dims, template, search, ids, text_mask, prompt, flag, want_logits=False):
    """UVLTrack.forward_test (lib/models/uvltrack/uvltrack.py:41-45)."""
    info = backbone(sd, dims, template, search, ids, text_mask, flag, want_logits)
    return box_head(sd, dims, info, prompt=prompt.astype(F32))


# ---------------------------------------------------------------------------------------------------------------
# prompter
# ---------------------------------------------------------------------------------------------------------------
def prompter(sd, dims, tem, tem_mask, ctx, ctx_mask, token, flag):
    """DistributionBasedCrossAttention.forward (heads/utils.py:45-99); dropout off."""
    p = "box_head.prompter."
    B = ctx.shape[0]
    src_ = np.repeat(sd[p + "query_embed.weight"][None], B, axis=0).astype(F32)
    src_[:, 0] = src_[:, 0] + token
    tgt = np.concatenate([tem, ctx], axis=1)
    tgt_mask = np.concatenate([tem_mask, ctx_mask], axis=1).astype(bool)[:, None]
    scale = np.exp(sd[p + "logit_scale"].astype(F32))
    sim = (l2_normalize(token)[:, None] @ l2_normalize(tgt).transpose(0, 2, 1) * scale).astype(F32)

    NEG = F32(-1e20)
    tgt_score = softmax(np.where(~tgt_mask, NEG, sim))
    tgt_token = tgt_score @ tgt
    bgd_logit = np.where(tgt_mask, NEG, sim)
    bgd_score = softmax(bgd_logit)
    # divide_background (:45-55)
    values = np.sort(bgd_score, axis=-1)
    m = np.cumsum(values, axis=-1, dtype=F32) < F32(0.25)
    thr = np.where(m, F32(1.0), values).min(-1, keepdims=True)
    dis_mask = bgd_score >= thr
    bgd_score2 = softmax(np.where(dis_mask, NEG, bgd_logit))
    dis_score = softmax(np.where(~dis_mask, NEG, bgd_logit))
    bgd_token = bgd_score2 @ tgt
    dis_token = dis_score @ tgt
    src = (np.concatenate([tgt_token, dis_token, bgd_token], axis=1) + src_).astype(F32)
    h = gelu(linear(src, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    src = linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]) + src
    sel = np.stack([src, src_, src], axis=1)
    return sel[np.arange(B), flag.reshape(-1)].astype(F32)


def select_token(info):
    """token_group[bid, flag] (modality_adaptive_box_head.py:97-102)."""
    vis, txt, flag = info["vis_token"], info["txt_token"], info["flag"]
    group = np.concatenate([vis, txt, (vis + txt) / F32(2)], axis=1)
    return group[np.arange(flag.shape[0]), flag]


def forward_prompt(sd, dims, info, template_mask, context_mask, ctx_rot=0):
    """box_head.forward_prompt (modality_adaptive_box_head.py:96-106)."""
    ctx = np.roll(info["search"], -ctx_rot, axis=0) if ctx_rot else info["search"]
    return prompter(sd, dims, info["template"], template_mask, ctx, context_mask, select_token(info), info["flag"])


def forward_prompt_init(sd, dims, template, search, ids, text_mask, template_mask, context_mask, flag):
    """UVLTrack.forward_prompt_init (lib/models/uvltrack/uvltrack.py:26-31)."""
    info = backbone(sd, dims, template, search, ids, text_mask, flag, want_logits=False)
    return forward_prompt(sd, dims, info, template_mask, context_mask)


def forward_train(sd, dims, template, search, ids, text_mask, template_mask, context_mask, flag):
    """UVLTrack.forward (lib/models/uvltrack/uvltrack.py:18-24) with the training branch of contractive_learning
    (modality_adaptive_box_head.py:123-137): context = search rolled by B//2."""
    info = backbone(sd, dims, template, search, ids, text_mask, flag, want_logits=False)
    B = flag.shape[0]
    prompt = forward_prompt(sd, dims, info, template_mask, context_mask, ctx_rot=B // 2)
    cs = cont_score_train(sd, dims, info["search"], prompt)
    return box_head(sd, dims, info, prompt=prompt, cont_score=cs)


# ---------------------------------------------------------------------------------------------------------------
# tracker post-processing (host side of the reference)
# ---------------------------------------------------------------------------------------------------------------
def hanning_window(S):
    """window_prior (lib/test/tracker/uvltrack.py:64-68): float64 outer product, exact zeros on the border."""
    h = np.hanning(S)
    return np.outer(h, h).flatten()


def track_decode(cls_map, cont_score, bbox_map, window, has_cont=True):
    """Tracker.track merge (lib/test/tracker/uvltrack.py:116-121) for one sequence: fp32 maps promoted to float64 by the
    numpy window.  Returns (box[4] fp32, score fp32, argmax index)."""
    cls = cls_map.reshape(-1).astype(F32)
    cont = softmax(cont_score.astype(F32))[:, 0] if has_cont else np.ones_like(cls)
    merge = cls.astype(np.float64) * window * cont.astype(np.float64)
    j = int(np.argmax(merge))
    return bbox_map.reshape(-1, 4)[j], F32(cls[j] * cont[j]), j


def map_box_back(state, pred_box, resize_factor, search_size):
    """UVLTrack.map_box_back (lib/test/tracker/uvltrack.py:167-173)."""
    cx_prev, cy_prev = state[0] + 0.5 * state[2], state[1] + 0.5 * state[3]
    cx, cy, w, h = pred_box
    half = 0.5 * search_size / resize_factor
    return [cx + (cx_prev - half) - 0.5 * w, cy + (cy_prev - half) - 0.5 * h, w, h]


def clip_box(box, H, W, margin=0):
    """lib/utils/box_ops.py:117-126."""
    x1, y1, w, h = box
    x2, y2 = x1 + w, y1 + h
    x1 = min(max(0, x1), W - margin)
    x2 = min(max(margin, x2), W)
    y1 = min(max(0, y1), H - margin)
    y2 = min(max(margin, y2), H)
    return [x1, y1, max(margin, x2 - x1), max(margin, y2 - y1)]
