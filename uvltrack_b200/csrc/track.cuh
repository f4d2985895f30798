// Device-side pre/post-processing of Tracker.track, so that a tracker step never leaves the GPU:
//   crop_resize_kernel : sample_target (lib/train/data/processing_utils.py:159-243, the tracker's branch without mask):
//                        square crop of side ceil(sqrt(w*h)*factor) around the current box, zero padded, resized to the
//                        search size with OpenCV's INTER_LINEAR fixed-point arithmetic for 8-bit images -- bit-exact with
//                        cv2.resize (tests/test_preprocess_gpu.py), so the uint8 crop equals the reference's.
//   box_update()       : pred_box * search_size / resize_factor, map_box_back (lib/test/tracker/uvltrack.py:167-173),
//                        clip_box(margin=10) (lib/utils/box_ops.py:117-126) in the reference's precisions
//                        (fp32 tensor arithmetic, then Python floats = fp64).
// The per-sequence box state lives in device memory as fp64 [B, 4] (x, y, w, h in frame pixels).
#pragma once
#include "common.cuh"

namespace uvlt {

struct CropGeom {
  int crop_sz;    // side of the square crop in frame pixels (0 => "Too small bounding box.")
  int x1, y1;     // top-left corner of the crop in the frame (may be negative)
  double resize_factor;
};

// processing_utils.py:177-192.  Python's round() is round-half-even == rint().
__device__ __forceinline__ CropGeom crop_geometry(const double* st, double factor, int out_sz) {
  CropGeom g;
  const double x = st[0], y = st[1], w = st[2], h = st[3];
  const double side = ceil(__dmul_rn(sqrt(__dmul_rn(w, h)), factor));
  g.crop_sz = (side >= 1.0 && side < 1.0e9) ? static_cast<int>(side) : 0;  // NaN / huge boxes are rejected as well
  const double half = __dmul_rn(static_cast<double>(g.crop_sz), 0.5);
  g.x1 = static_cast<int>(rint(__dadd_rn(__dadd_rn(x, __dmul_rn(0.5, w)), -half)));
  g.y1 = static_cast<int>(rint(__dadd_rn(__dadd_rn(y, __dmul_rn(0.5, h)), -half)));
  g.resize_factor = g.crop_sz > 0 ? static_cast<double>(out_sz) / static_cast<double>(g.crop_sz) : 0.0;
  return g;
}

// cv::resize INTER_LINEAR, 8-bit: source index and the two 11-bit fixed-point taps of destination index d.
// (OpenCV resize.cpp: fx = (float)((d + 0.5) * scale - 0.5); the x axis clamps the index AND the weight at the borders,
//  the y axis clamps only the row indices.)
__device__ __forceinline__ void linear_tap(int d, double scale, int ssize, bool clamp_weight, int* s, int* a0, int* a1) {
  float f = static_cast<float>(__dadd_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), -0.5));
  int si = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(si));
  if (clamp_weight) {
    if (si < 0) { f = 0.f; si = 0; }
    if (si >= ssize - 1) { f = 0.f; si = ssize - 1; }
  }
  *s = si;
  *a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  *a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

struct CropParams {
  const uint8_t* frames;   // [B, H, W, 3] RGB
  int H, W;
  const double* state;     // [B, 4]
  double factor;           // search_factor
  int out_sz;              // search_size
  uint8_t* crops;          // [B, out_sz, out_sz, 3]
  double* resize_factor;   // [B]
};

// one thread per output pixel (3 channels)
static __global__ void __launch_bounds__(256) crop_resize_kernel(const CropParams p) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const CropGeom g = crop_geometry(p.state + 4 * b, p.factor, p.out_sz);
  if (pix == 0) p.resize_factor[b] = g.resize_factor;
  if (pix >= p.out_sz * p.out_sz) return;
  const int dy = pix / p.out_sz, dx = pix - dy * p.out_sz;
  uint8_t* dst = p.crops + (static_cast<long long>(b) * p.out_sz * p.out_sz + pix) * 3;
  if (g.crop_sz == 0) { dst[0] = dst[1] = dst[2] = 0; return; }
  const double scale = 1.0 / (static_cast<double>(p.out_sz) / static_cast<double>(g.crop_sz));
  int sx, ax0, ax1, sy, by0, by1;
  linear_tap(dx, scale, g.crop_sz, true, &sx, &ax0, &ax1);
  linear_tap(dy, scale, g.crop_sz, false, &sy, &by0, &by1);
  const int sx1 = min(sx + 1, g.crop_sz - 1);
  const int ry0 = min(max(sy, 0), g.crop_sz - 1), ry1 = min(max(sy + 1, 0), g.crop_sz - 1);
  const uint8_t* img = p.frames + static_cast<long long>(b) * p.H * p.W * 3;
  // a crop pixel (yy, xx) is frame pixel (y1 + yy, x1 + xx) when that lies inside [0, H-2] x [0, W-2], else the zero
  // border (the reference pads one column / row more than needed on the far side: `x2 - im.shape[1] + 1`, :186-190)
  auto fetch = [&](int yy, int xx, int c) -> int {
    const int Y = g.y1 + yy, X = g.x1 + xx;
    if (Y < 0 || X < 0 || Y > p.H - 2 || X > p.W - 2) return 0;
    return img[(static_cast<long long>(Y) * p.W + X) * 3 + c];
  };
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int r0 = fetch(ry0, sx, c) * ax0 + fetch(ry0, sx1, c) * ax1;  // horizontal pass (int, <= 255 * 2048)
    const int r1 = fetch(ry1, sx, c) * ax0 + fetch(ry1, sx1, c) * ax1;
    const int v = (((by0 * (r0 >> 4)) >> 16) + ((by1 * (r1 >> 4)) >> 16) + 2) >> 2;  // VResizeLinear<uchar>
    dst[c] = static_cast<uint8_t>(min(max(v, 0), 255));
  }
}

// lib/test/tracker/uvltrack.py:123-125 + :167-173 + box_ops.py:117-126
__device__ __forceinline__ void box_update(const float4 net, double rf, int search_size, int H, int W, double* st) {
  const float S = static_cast<float>(search_size);
  const float rf32 = static_cast<float>(rf);
  // (pred_boxes * search_size / resize_factor): fp32 tensor arithmetic, then .tolist() -> fp64
  const double cx = static_cast<double>(__fdiv_rn(__fmul_rn(net.x, S), rf32));
  const double cy = static_cast<double>(__fdiv_rn(__fmul_rn(net.y, S), rf32));
  const double w = static_cast<double>(__fdiv_rn(__fmul_rn(net.z, S), rf32));
  const double h = static_cast<double>(__fdiv_rn(__fmul_rn(net.w, S), rf32));
  const double cx_prev = __dadd_rn(st[0], __dmul_rn(0.5, st[2]));
  const double cy_prev = __dadd_rn(st[1], __dmul_rn(0.5, st[3]));
  const double half_side = __ddiv_rn(__dmul_rn(0.5, static_cast<double>(search_size)), rf);
  const double cx_real = __dadd_rn(cx, __dadd_rn(cx_prev, -half_side));
  const double cy_real = __dadd_rn(cy, __dadd_rn(cy_prev, -half_side));
  double x1 = __dadd_rn(cx_real, -__dmul_rn(0.5, w));
  double y1 = __dadd_rn(cy_real, -__dmul_rn(0.5, h));
  const double margin = 10.0;
  double x2 = __dadd_rn(x1, w), y2 = __dadd_rn(y1, h);
  x1 = fmin(fmax(0.0, x1), static_cast<double>(W) - margin);
  x2 = fmin(fmax(margin, x2), static_cast<double>(W));
  y1 = fmin(fmax(0.0, y1), static_cast<double>(H) - margin);
  y2 = fmin(fmax(margin, y2), static_cast<double>(H));
  st[0] = x1;
  st[1] = y1;
  st[2] = fmax(margin, __dadd_rn(x2, -x1));
  st[3] = fmax(margin, __dadd_rn(y2, -y1));
}

// box_update alone, one thread per sequence (operator-level entry for the parity fixtures of tests/golden/preproc.npz)
static __global__ void box_update_kernel(const float* net, const double* rf, int search_size, int H, int W, double* state,
                                         int B) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float4 bb = make_float4(net[4 * b], net[4 * b + 1], net[4 * b + 2], net[4 * b + 3]);
  box_update(bb, rf[b], search_size, H, W, state + 4 * b);
}

// Tracker.anno2mask (lib/test/tracker/uvltrack.py:183-194): normalised [x, y, w, h] boxes -> cells of a size x size grid
// whose centre lies strictly inside the box, plus the cell under the box centre.  fp32 arithmetic in the reference's
// order (box_xywh_to_xyxy then * size; torch's .long() truncates toward zero), so the mask is bit-identical.
static __global__ void anno2mask_kernel(const float* boxes, int size, uint8_t* mask, int B) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || cell >= size * size) return;
  const float fs = static_cast<float>(size);
  const float x1 = __fmul_rn(boxes[4 * b], fs), y1 = __fmul_rn(boxes[4 * b + 1], fs);
  const float x2 = __fmul_rn(__fadd_rn(boxes[4 * b], boxes[4 * b + 2]), fs);
  const float y2 = __fmul_rn(__fadd_rn(boxes[4 * b + 1], boxes[4 * b + 3]), fs);
  const int r = cell / size, c = cell - r * size;
  const float cx = static_cast<float>(c) + 0.5f, cy = static_cast<float>(r) + 0.5f;
  bool in = (cx > x1) && (cx < x2) && (cy > y1) && (cy < y2);
  const long long ccx = static_cast<long long>(__fdiv_rn(__fadd_rn(x1, x2), 2.0f));
  const long long ccy = static_cast<long long>(__fdiv_rn(__fadd_rn(y1, y2), 2.0f));
  // the reference indexes mask[b, cy, cx] with Python semantics: a negative index wraps once, anything else out of range
  // raises; in-range indices are the only ones a box inside its crop produces
  const long long wx = ccx < 0 ? ccx + size : ccx, wy = ccy < 0 ? ccy + size : ccy;
  in = in || (wx == c && wy == r);
  mask[static_cast<long long>(b) * size * size + cell] = in ? 1 : 0;
}

// Preprocessor_wo_mask.process (lib/test/tracker/tracker_utils.py:25-29): uint8 HWC crop -> fp32 [3, S, S],
// ((x / 255) - mean) / std in fp32 in that order.
// grounding_resize (lib/train/data/processing_utils.py:60-141, image part): the WHOLE frame resized without changing its
// aspect ratio so that its longer side is out_sz (cv2.resize INTER_LINEAR, the same 8-bit fixed-point arithmetic as
// crop_resize_kernel, but with separate x / y scales and no border: every tap lies inside the frame), centred in an
// out_sz x out_sz canvas of zeros.  NL-mode first frame of Tracker.initialize (lib/test/tracker/uvltrack.py:45-62).
struct GroundParams {
  const uint8_t* frames;   // [B, H, W, 3] RGB
  int H, W;
  int out_sz;
  uint8_t* out;            // [B, out_sz, out_sz, 3]
};

static __global__ void __launch_bounds__(256) grounding_resize_kernel(const GroundParams p) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= p.out_sz * p.out_sz) return;
  const int dy = pix / p.out_sz, dx = pix - dy * p.out_sz;
  // processing_utils.py:83-100 (Python: int(output_sz * h / w) = truncation of a double quotient; pads int((sz - o) / 2),
  // the first one a pixel larger when the difference is odd)
  int ow, oh;
  if (p.W > p.H) {
    ow = p.out_sz;
    oh = static_cast<int>(static_cast<double>(p.out_sz * p.H) / static_cast<double>(p.W));
  } else {
    oh = p.out_sz;
    ow = static_cast<int>(static_cast<double>(p.out_sz * p.W) / static_cast<double>(p.H));
  }
  int y1 = (p.out_sz - oh) / 2, x1 = (p.out_sz - ow) / 2;
  if (2 * y1 + oh != p.out_sz) y1 += 1;
  if (2 * x1 + ow != p.out_sz) x1 += 1;
  uint8_t* dst = p.out + (static_cast<long long>(b) * p.out_sz * p.out_sz + pix) * 3;
  const int ry = dy - y1, rx = dx - x1;
  if (ry < 0 || ry >= oh || rx < 0 || rx >= ow || ow < 1 || oh < 1) { dst[0] = dst[1] = dst[2] = 0; return; }
  const double scale_x = 1.0 / (static_cast<double>(ow) / static_cast<double>(p.W));
  const double scale_y = 1.0 / (static_cast<double>(oh) / static_cast<double>(p.H));
  int sx, ax0, ax1, sy, by0, by1;
  linear_tap(rx, scale_x, p.W, true, &sx, &ax0, &ax1);
  linear_tap(ry, scale_y, p.H, false, &sy, &by0, &by1);
  const int sx1 = min(sx + 1, p.W - 1);
  const int ry0 = min(max(sy, 0), p.H - 1), ry1 = min(max(sy + 1, 0), p.H - 1);
  const uint8_t* img = p.frames + static_cast<long long>(b) * p.H * p.W * 3;
  const uint8_t* row0 = img + static_cast<long long>(ry0) * p.W * 3;
  const uint8_t* row1 = img + static_cast<long long>(ry1) * p.W * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int r0 = row0[sx * 3 + c] * ax0 + row0[sx1 * 3 + c] * ax1;
    const int r1 = row1[sx * 3 + c] * ax0 + row1[sx1 * 3 + c] * ax1;
    const int v = (((by0 * (r0 >> 4)) >> 16) + ((by1 * (r1 >> 4)) >> 16) + 2) >> 2;  // VResizeLinear<uchar>
    dst[c] = static_cast<uint8_t>(min(max(v, 0), 255));
  }
}

static __global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* crops, float* out, int S, int B) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || pix >= S * S) return;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const uint8_t* src = crops + (static_cast<long long>(b) * S * S + pix) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[(static_cast<long long>(b) * 3 + c) * S * S + pix] =
        __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(src[c]), 255.0f), mean[c]), stdv[c]);
}

}  // namespace uvlt
