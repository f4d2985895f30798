"""Python face of the sm_100a engine (``libuvlt_sm100.so``): owns one ``uvlt_handle`` and turns torch CUDA tensors
into the raw device pointers the C ABI takes.  PyTorch is plumbing here (device memory, streams); every arithmetic
operation of the hot path runs in the hand-written kernels behind the ABI.  There is no fallback: a missing library or
a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _cabi
from .weights import ModelDims


class _DevArray:
    """Zero-copy view of engine-owned device memory for ``torch.as_tensor`` (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def _view(ptr, shape):
    import torch

    return torch.as_tensor(_DevArray(ptr, shape), device="cuda")


class Engine:
    """One engine per process/GPU (the reference also builds one network per tracker process,
    lib/test/tracker/uvltrack.py:23-28)."""

    def __init__(self, dims: ModelDims, max_batch: int = 1):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("uvltrack_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _cabi.load()
        self.dims = dims
        self.max_batch = int(max_batch)
        cfg = _cabi.UvltConfig()
        cfg.embed_dim, cfg.num_heads, cfg.depth, cfg.mlp_hidden = dims.embed_dim, dims.num_heads, dims.depth, dims.mlp_hidden
        cfg.template_size, cfg.search_size, cfg.text_len = dims.template_size, dims.search_size, dims.text_len
        cfg.fusion_start, cfg.head_channels = dims.fusion_start, dims.head_channels
        cfg.vocab_size, cfg.max_position, cfg.max_batch = dims.vocab_size, dims.max_position, self.max_batch
        cfg.softmax_one, cfg.offset_sigmoid = int(dims.softmax_one), int(dims.offset_sigmoid)
        cfg.txt_token_mean = int(dims.txt_token_mode == "mean")
        cfg.num_cont_layers = len(dims.cont_loss_layers)
        for i, l in enumerate(dims.cont_loss_layers):
            cfg.cont_layers[i] = int(l)
        h = C.c_void_p()
        _cabi.check(self.lib.uvlt_create(C.byref(cfg), C.byref(h)), "uvlt_create")
        self.h = h
        self._loaded = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.uvlt_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = False):
        """Reference-format state_dict (torch tensors or numpy arrays).  Unknown / unused keys are skipped like
        ``load_state_dict(strict=False)`` does; a missing required tensor raises from ``uvlt_finalize_weights``."""
        ignored = []
        for k, v in state_dict.items():
            if hasattr(v, "detach"):
                v = v.detach().cpu().numpy()
            a = np.ascontiguousarray(np.asarray(v), dtype=np.float32)
            shape = (C.c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
            rc = self.lib.uvlt_set_weight(self.h, k.encode(), a.ctypes.data_as(C.c_void_p), shape, max(a.ndim, 1))
            if rc == 2:
                ignored.append(k)
            elif rc != 0:
                _cabi.check(rc, f"uvlt_set_weight({k})")
        if strict and ignored:
            raise KeyError(f"keys not used by the hot path: {ignored[:8]}")
        _cabi.check(self.lib.uvlt_finalize_weights(self.h), "uvlt_finalize_weights")
        self._loaded = True
        return ignored

    def set_option(self, name: str, value: int):
        _cabi.check(self.lib.uvlt_set_option(self.h, name.encode(), int(value)), f"uvlt_set_option({name})")

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.uvlt_last_launch_count(self.h))

    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _f32(t):
        import torch

        return t.to(device="cuda", dtype=torch.float32).contiguous()

    @staticmethod
    def _i64(t):
        import torch

        return t.to(device="cuda", dtype=torch.int64).contiguous()

    @staticmethod
    def _u8(t):
        import torch

        return t.to(device="cuda", dtype=torch.uint8).contiguous()

    def _prep_text(self, text, flag, B):
        ids = self._i64(text.tensors).reshape(B, -1)
        mask = self._f32(text.mask).reshape(B, -1)
        flag = self._i64(flag).reshape(-1)
        if ids.shape[1] != self.dims.text_len or flag.numel() != B:
            raise ValueError("text / flag shapes do not match the engine configuration")
        return ids, mask, flag

    def _outputs(self, out: _cabi.UvltOutputs, head: bool, clone: bool) -> Dict[str, "object"]:
        d, B = self.dims, out.batch
        S, nz, nv = d.feat_size, d.nz, d.n_visual
        tokens = _view(out.tokens, (B, d.n_tokens, d.embed_dim))
        if clone:
            tokens = tokens.clone()
        res = {
            "tokens": tokens,
            "search": tokens[:, 1 + nz:nv], "template": tokens[:, 1:1 + nz], "text": tokens[:, nv:],
            "vis_token": tokens[:, :1],
        }
        if out.logits:
            lg = _view(out.logits, (B, len(d.cont_loss_layers), S, S))
            res["logits"] = lg.clone() if clone else lg
        if head:
            for key, ptr, shape in (("cls_score_test", out.cls_score, (B, S, S)),
                                    ("bbox_map", out.bbox_map, (B, S * S, 4)),
                                    ("pred_boxes", out.pred_boxes, (B, 1, 4)),
                                    ("cont_score", out.cont_score, (B, S * S, out.cont_cols)),
                                    ("prompts", out.prompts, (B, 3, d.embed_dim))):
                t = _view(ptr, shape)
                res[key] = t.clone() if clone else t
            res["cls_score"] = res["cls_score_test"]  # JOINT_CLS false (modality_adaptive_box_head.py:87)
        return res

    def _txt_token(self, res, mask):
        if self.dims.txt_token_mode == "mean":  # modality_unified_feature_extractor.py:80-81 (tiny, output assembly)
            m = mask.unsqueeze(-1)
            return (res["text"] * m).sum(dim=1, keepdim=True) / m.sum(dim=1, keepdim=True)
        return res["text"][:, :1]

    # ------------------------------------------------------------------------------------------------------
    def text_encode(self, text, flag):
        """BERT embedding + BERT-only layers once per sequence (constant text): fills the engine's text cache that
        ``text_cached=True`` forwards restore instead of re-running the branch every frame."""
        B = text.tensors.shape[0]
        ids, mask, fl = self._prep_text(text, flag, B)
        _cabi.check(self.lib.uvlt_text_encode(self.h, _cabi.ptr(ids), _cabi.ptr(mask), _cabi.ptr(fl), B,
                                              _cabi.current_stream()), "uvlt_text_encode")
        self._text_owner = None  # whoever relies on the cache re-claims it (BatchTracker does)

    def forward_test(self, template, search, text, prompt, flag, want_logits=False, skip_text=False, clone=True,
                     text_cached=False):
        """UVLTrack.forward_test (lib/models/uvltrack/uvltrack.py:41-45)."""
        import torch

        B = search.shape[0]
        tmpl, srch, prompt = self._f32(template), self._f32(search), self._f32(prompt)
        ids, mask, fl = self._prep_text(text, flag, B)
        out = _cabi.UvltOutputs()
        flags = ((_cabi.WANT_LOGITS if want_logits else 0) | (_cabi.SKIP_TEXT if skip_text else 0) |
                 (_cabi.TEXT_CACHED if text_cached else 0))
        _cabi.check(self.lib.uvlt_forward_test(self.h, _cabi.ptr(tmpl), _cabi.ptr(srch), _cabi.ptr(ids), _cabi.ptr(mask),
                                               _cabi.ptr(prompt), _cabi.ptr(fl), B, flags, C.byref(out),
                                               _cabi.current_stream()), "uvlt_forward_test")
        res = self._outputs(out, head=True, clone=clone)
        res["txt_token"] = self._txt_token(res, mask)
        res["flag"] = fl
        res["prompt"] = prompt
        res["_text_mask"] = mask
        return res

    def head(self, search, prompt, flag, clone=True):
        """ModalityAdaptiveBoxHead.forward, test branch (modality_adaptive_box_head.py:62-94,140-148) on backbone features
        ``search`` [B, Nx, D] (None: the token stream the last backbone call left in the engine)."""
        prompt = self._f32(prompt)
        B = prompt.shape[0]
        fl = self._i64(flag).reshape(-1)
        srch = None if search is None else self._f32(search)
        if srch is not None and tuple(srch.shape) != (B, self.dims.nx, self.dims.embed_dim):
            raise ValueError("search tokens must be [B, Nx, D]")
        out = _cabi.UvltOutputs()
        _cabi.check(self.lib.uvlt_head(self.h, _cabi.ptr(srch), _cabi.ptr(prompt), _cabi.ptr(fl), B, C.byref(out),
                                       _cabi.current_stream()), "uvlt_head")
        res = self._outputs(out, head=True, clone=clone)
        return {k: res[k] for k in ("cls_score", "cls_score_test", "bbox_map", "pred_boxes", "cont_score", "prompts")}

    def backbone(self, template, search, text, flag, want_logits=False, clone=True):
        """ModalityUnifiedFeatureExtractor.forward (modality_unified_feature_extractor.py:52-77)."""
        B = search.shape[0]
        tmpl, srch = self._f32(template), self._f32(search)
        ids, mask, fl = self._prep_text(text, flag, B)
        out = _cabi.UvltOutputs()
        _cabi.check(self.lib.uvlt_backbone(self.h, _cabi.ptr(tmpl), _cabi.ptr(srch), _cabi.ptr(ids), _cabi.ptr(mask),
                                           _cabi.ptr(fl), B, _cabi.WANT_LOGITS if want_logits else 0, C.byref(out),
                                           _cabi.current_stream()), "uvlt_backbone")
        res = self._outputs(out, head=False, clone=clone)
        res["txt_token"] = self._txt_token(res, mask)
        res["flag"] = fl
        res["_text_mask"] = mask
        return res

    def forward_train(self, template, search, text, template_mask, context_mask, flag, want_logits=False):
        """UVLTrack.forward (lib/models/uvltrack/uvltrack.py:18-24)."""
        B = search.shape[0]
        tmpl, srch = self._f32(template), self._f32(search)
        ids, mask, fl = self._prep_text(text, flag, B)
        tm, cm = self._u8(template_mask).reshape(B, -1), self._u8(context_mask).reshape(B, -1)
        out = _cabi.UvltOutputs()
        _cabi.check(self.lib.uvlt_forward_train(self.h, _cabi.ptr(tmpl), _cabi.ptr(srch), _cabi.ptr(ids),
                                                _cabi.ptr(mask), _cabi.ptr(fl), _cabi.ptr(tm), _cabi.ptr(cm), B,
                                                _cabi.WANT_LOGITS if want_logits else 0, C.byref(out),
                                                _cabi.current_stream()), "uvlt_forward_train")
        res = self._outputs(out, head=True, clone=True)
        res["txt_token"] = self._txt_token(res, mask)
        res["flag"] = fl
        res["template_mask"], res["context_mask"] = template_mask, context_mask
        res["_text_mask"] = mask
        return res

    def forward_prompt(self, tokens, flag, template_mask, context_mask, text_mask=None):
        """box_head.forward_prompt (modality_adaptive_box_head.py:96-106) on a [B, N, D] token stream
        (``None`` = the stream left by the last forward on this engine)."""
        import torch

        fl = self._i64(flag).reshape(-1)
        B = fl.numel()
        tm, cm = self._u8(template_mask).reshape(B, -1), self._u8(context_mask).reshape(B, -1)
        if tm.shape[1] != self.dims.nz or cm.shape[1] != self.dims.nx:
            raise ValueError("template/context mask sizes do not match the engine configuration")
        tk = None if tokens is None else self._f32(tokens)
        tmask = None if text_mask is None else self._f32(text_mask)
        out = torch.empty(B, 3, self.dims.embed_dim, device="cuda", dtype=torch.float32)
        _cabi.check(self.lib.uvlt_forward_prompt(self.h, _cabi.ptr(tk), _cabi.ptr(fl), _cabi.ptr(tmask), _cabi.ptr(tm),
                                                 _cabi.ptr(cm), B, _cabi.ptr(out), _cabi.current_stream()),
                    "uvlt_forward_prompt")
        return out

    def track_decode(self, window, has_cont=True, max_score=None, snapshot=None):
        """Tracker.track merge (lib/test/tracker/uvltrack.py:116-121) on the last forward.  Returns a [B,6] tensor."""
        import torch

        out = torch.empty(self.max_batch, 6, device="cuda", dtype=torch.float32)
        _cabi.check(self.lib.uvlt_track_decode(self.h, _cabi.ptr(window), int(has_cont), _cabi.ptr(max_score),
                                               _cabi.ptr(snapshot), _cabi.ptr(out), _cabi.current_stream()),
                    "uvlt_track_decode")
        return out

    def track_frame_host(self, search_u8_pinned, template, ids, text_mask, prompt, flag, window, out_pinned, batch,
                         has_cont=True, skip_text=False, max_score=None, snapshot=None, text_cached=False):
        """One tracker step from pinned host memory (uint8 crops in, [B,6] rows out); synchronises the stream."""
        _cabi.check(self.lib.uvlt_track_frame_host(
            self.h, C.c_void_p(search_u8_pinned.data_ptr()), _cabi.ptr(template), _cabi.ptr(ids), _cabi.ptr(text_mask),
            _cabi.ptr(prompt), _cabi.ptr(flag), _cabi.ptr(window), int(batch),
            (_cabi.SKIP_TEXT if skip_text else 0) | (_cabi.TEXT_CACHED if text_cached else 0),
            int(has_cont), _cabi.ptr(max_score), _cabi.ptr(snapshot), C.c_void_p(out_pinned.data_ptr()),
            _cabi.current_stream()), "uvlt_track_frame_host")
        return out_pinned

    def track_frame_image_host(self, frames_pinned, state_dev, search_factor, template, ids, text_mask, prompt, flag,
                               window, out_pinned, batch, has_cont=True, skip_text=False, max_score=None, snapshot=None,
                               text_cached=False, uploaded=False):
        """One tracker step from the raw frames (pinned uint8 [B,H,W,3]): crop + resize, forward, merge and box update
        all on the device; ``state_dev`` (fp64 [B,4]) is updated in place, ``out_pinned`` (fp64 [B,10]) receives the rows.
        Synchronises the stream."""
        B, H, W, _ = frames_pinned.shape
        _cabi.check(self.lib.uvlt_track_frame_image_host(
            self.h, None if uploaded else C.c_void_p(frames_pinned.data_ptr()), int(H), int(W), _cabi.ptr(state_dev),
            float(search_factor),
            _cabi.ptr(template), _cabi.ptr(ids), _cabi.ptr(text_mask), _cabi.ptr(prompt), _cabi.ptr(flag),
            _cabi.ptr(window), int(batch), (_cabi.SKIP_TEXT if skip_text else 0) | (_cabi.TEXT_CACHED if text_cached else 0),
            int(has_cont), _cabi.ptr(max_score), _cabi.ptr(snapshot), C.c_void_p(out_pinned.data_ptr()),
            _cabi.current_stream()), "uvlt_track_frame_image_host")
        return out_pinned

    def upload_frames(self, pinned_piece, dst_offset: int, total_bytes: int):
        """Enqueue the H2D copy of one piece of the next step's frames (see uvlt_upload_frames)."""
        _cabi.check(self.lib.uvlt_upload_frames(self.h, C.c_void_p(pinned_piece.data_ptr()), int(dst_offset),
                                                int(pinned_piece.numel() * pinned_piece.element_size()), int(total_bytes),
                                                _cabi.current_stream()), "uvlt_upload_frames")
