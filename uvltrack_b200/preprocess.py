"""Host-side pre/post-processing of the tracker (numpy + OpenCV), with the reference's arithmetic.

``sample_target`` follows lib/train/data/processing_utils.py:159-243 (the branch the tracker uses: no mask), the box
helpers follow lib/utils/box_ops.py and lib/test/tracker/uvltrack.py:167-194.  The image normalisation of
``Preprocessor_wo_mask`` (lib/test/tracker/tracker_utils.py:20-29) is NOT here: for search crops it is fused into the
patch-embedding kernel (raw uint8 crops go to the device), for the template it runs once per sequence.
"""
from __future__ import annotations

import math

import numpy as np


def sample_target(im: np.ndarray, target_bb, search_area_factor: float, output_sz: int):
    """Square crop of side ceil(sqrt(w*h) * factor) centred on the box, zero padded, resized (cv2 bilinear).
    Returns (crop uint8 [output_sz, output_sz, 3], resize_factor, normalised box [x, y, w, h] inside the crop)."""
    import cv2

    x, y, w, h = [float(v) for v in target_bb]
    crop_sz = math.ceil(math.sqrt(w * h) * search_area_factor)
    if crop_sz < 1:
        raise Exception("Too small bounding box.")
    x1 = int(round(x + 0.5 * w - crop_sz * 0.5))
    x2 = x1 + crop_sz
    y1 = int(round(y + 0.5 * h - crop_sz * 0.5))
    y2 = y1 + crop_sz
    x1_pad = max(0, -x1)
    x2_pad = max(x2 - im.shape[1] + 1, 0)
    y1_pad = max(0, -y1)
    y2_pad = max(y2 - im.shape[0] + 1, 0)
    crop = im[y1 + y1_pad:y2 - y2_pad, x1 + x1_pad:x2 - x2_pad, :]
    padded = cv2.copyMakeBorder(crop, y1_pad, y2_pad, x1_pad, x2_pad, cv2.BORDER_CONSTANT)
    bbox = np.array([0.5 - w / crop_sz / 2, 0.5 - h / crop_sz / 2, w / crop_sz, h / crop_sz], dtype=np.float32)
    resize_factor = output_sz / crop_sz
    resized = cv2.resize(padded, (output_sz, output_sz))
    return resized, resize_factor, bbox


def search_window(target_bb, search_area_factor: float, frame_h: int, frame_w: int, slack: int = 2):
    """Rectangle [xa, xb) x [ya, yb) of the frame that contains every pixel ``sample_target`` reads for this box
    (the crop window x1..x2, y1..y2 of processing_utils.py:183-199 clipped to the frame), padded by `slack` pixels
    against the rounding of the window origin.  None when the window lies entirely outside the frame (the crop is all
    padding) or the box is degenerate.  Used by BatchTracker.track() to upload only this part of each frame."""
    x, y, w, h = [float(v) for v in target_bb]
    side = math.ceil(math.sqrt(max(w * h, 0.0)) * search_area_factor)
    if side < 1:
        return None
    xa = max(int(math.floor(x + 0.5 * w - 0.5 * side)) - slack, 0)
    ya = max(int(math.floor(y + 0.5 * h - 0.5 * side)) - slack, 0)
    xb = min(int(math.floor(x + 0.5 * w - 0.5 * side)) + side + slack + 2, frame_w)
    yb = min(int(math.floor(y + 0.5 * h - 0.5 * side)) + side + slack + 2, frame_h)
    if xb <= xa or yb <= ya:
        return None
    return xa, ya, xb, yb


def grounding_resize(im: np.ndarray, output_sz: int):
    """Aspect-preserving resize of the whole frame to fit output_sz, centre padded with zeros
    (lib/train/data/processing_utils.py:60-141, image part only)."""
    import cv2

    h, w = im.shape[:2]
    if w > h:
        ow, oh = output_sz, int(output_sz * h / w)
    else:
        oh, ow = output_sz, int(output_sz * w / h)
    img = cv2.resize(im, (ow, oh), interpolation=cv2.INTER_LINEAR)  # the reference passes PIL's BILINEAR (== 2)...
    y1 = y2 = int((output_sz - oh) / 2)
    x1 = x2 = int((output_sz - ow) / 2)
    if y1 + y2 + oh != output_sz:
        y1 += 1
    if x1 + x2 + ow != output_sz:
        x1 += 1
    return cv2.copyMakeBorder(img, y1, y2, x1, x2, cv2.BORDER_CONSTANT, value=(0, 0, 0))


def grounding_box(pred_cxcywh, frame_h: int, frame_w: int):
    """The box arithmetic of Tracker.grounding (lib/test/tracker/uvltrack.py:58-62): the network's normalised
    (cx, cy, w, h) of the padded square frame -> [x, y, w, h] in frame pixels.  fp32 tensor arithmetic
    (`pred * max(h, w)`, box_cxcywh_to_xywh), then Python floats; the padding offset of the shorter side is removed."""
    p = np.asarray(pred_cxcywh, dtype=np.float32).reshape(4) * np.float32(max(frame_h, frame_w))
    half = np.float32(0.5)
    box = [float(np.float32(p[0] - half * p[2])), float(np.float32(p[1] - half * p[3])), float(p[2]), float(p[3])]
    dx, dy = min(0, (frame_w - frame_h) / 2), min(0, (frame_h - frame_w) / 2)
    box[0] = box[0] + dx
    box[1] = box[1] + dy
    return box


def normalize_image(img_u8: np.ndarray) -> np.ndarray:
    """Preprocessor_wo_mask.process (lib/test/tracker/tracker_utils.py:25-29) on the host: HWC uint8 -> [1,3,H,W] fp32."""
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape(1, 3, 1, 1)
    std = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape(1, 3, 1, 1)
    x = img_u8.astype(np.float32).transpose(2, 0, 1)[None]
    return ((x / np.float32(255.0)) - mean) / std


def anno2mask(box_xywh: np.ndarray, size: int) -> np.ndarray:
    """Tracker.anno2mask (lib/test/tracker/uvltrack.py:183-194): normalised [x, y, w, h] boxes [B,4] -> bool [B, size*size]
    (cells whose centre lies inside the box, plus the cell under the box centre)."""
    box = np.asarray(box_xywh, dtype=np.float32).reshape(-1, 4)
    xyxy = np.stack([box[:, 0], box[:, 1], box[:, 0] + box[:, 2], box[:, 1] + box[:, 3]], axis=1) * np.float32(size)
    cood = (np.arange(size, dtype=np.float32) + np.float32(0.5))[None]
    xm = (cood > xyxy[:, 0:1]) & (cood < xyxy[:, 2:3])
    ym = (cood > xyxy[:, 1:2]) & (cood < xyxy[:, 3:4])
    mask = ym[:, :, None] & xm[:, None, :]
    cx = ((xyxy[:, 0] + xyxy[:, 2]) / 2).astype(np.int64)
    cy = ((xyxy[:, 1] + xyxy[:, 3]) / 2).astype(np.int64)
    mask[np.arange(box.shape[0]), cy, cx] = True
    return mask.reshape(box.shape[0], -1)


def map_box_back(state, pred_box, resize_factor: float, search_size: int):
    """lib/test/tracker/uvltrack.py:167-173."""
    cx_prev, cy_prev = state[0] + 0.5 * state[2], state[1] + 0.5 * state[3]
    cx, cy, w, h = pred_box
    half_side = 0.5 * search_size / resize_factor
    cx_real = cx + (cx_prev - half_side)
    cy_real = cy + (cy_prev - half_side)
    return [cx_real - 0.5 * w, cy_real - 0.5 * h, w, h]


def clip_box(box, H, W, margin=0):
    """lib/utils/box_ops.py:117-126."""
    x1, y1, w, h = box
    x2, y2 = x1 + w, y1 + h
    x1 = min(max(0, x1), W - margin)
    x2 = min(max(margin, x2), W)
    y1 = min(max(0, y1), H - margin)
    y2 = min(max(margin, y2), H)
    w = max(margin, x2 - x1)
    h = max(margin, y2 - y1)
    return [x1, y1, w, h]


def hanning_window(map_size: int) -> np.ndarray:
    """window_prior (lib/test/tracker/uvltrack.py:64-68): float64 outer product of np.hanning, flattened."""
    h = np.hanning(map_size)
    return np.outer(h, h).flatten()
