// rpn_ops.cu — RPN anchor <-> ground-truth matching for a whole batch in two launches.
//
// Replaces the (B, G, A)-shaped ATen passes behind detectron2's pairwise_iou + Matcher(allow_low_quality_matches)
// and cubercnn's ignore-region IoA (cubercnn/modeling/proposal_generator/rpn.py:93-105 label_and_sample_anchors,
// :286-330; detectron2 Matcher.__call__ / set_low_quality_matches_): ~25 element-wise / reduction kernels over
// B*G*A = 26 M elements per step become two passes that keep the G boxes of an image in shared memory.
//   pass 1: per (image, anchor): best GT (first maximal IoU over valid GTs), max IoA over ignore regions,
//           per-GT running maximum over anchors (warp-reduced, then atomicMax on the fp32 bit pattern, IoU >= 0)
//   pass 2: per (image, anchor): label = IoU >= fg_thresh, or any valid GT whose maximum this anchor attains
//           (low-quality matches); per-GT first arg-max anchor via atomicMin
// The IoU arithmetic uses explicit round-to-nearest intrinsics in the operation order of the torch formulation
// (omni3d_b200/cubercnn/rpn.py pairwise_iou / pairwise_ioa), so results are bit-identical to it.
#include "c3d_common.cuh"

namespace c3d {

struct PairQ { float iou, ioa; };

__device__ __forceinline__ PairQ gt_anchor_quality(const float4 g, const float4 a, const float area_g, const float area_a) {
  const float w = fmaxf(__fsub_rn(fminf(g.z, a.z), fmaxf(g.x, a.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(g.w, a.w), fmaxf(g.y, a.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  PairQ q;
  q.iou = inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_g, area_a), inter)) : 0.f;
  q.ioa = inter > 0.f ? __fdiv_rn(inter, area_a) : 0.f;
  return q;
}
__device__ __forceinline__ float box_area4(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

constexpr int kMaxG = 512;

__global__ void anchor_match_init_kernel(int* __restrict__ rowmax_bits, int* __restrict__ best_idx, int n, int A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { rowmax_bits[i] = 0; best_idx[i] = A; }
}

template <int PASS>
__global__ void __launch_bounds__(256)
anchor_match_kernel(const float4* __restrict__ anchors, int A, const float4* __restrict__ gt, const unsigned char* __restrict__ gt_valid,
                    const unsigned char* __restrict__ gt_ign, int G, float fg_thresh, long long* __restrict__ matched_idx,
                    float* __restrict__ matched_iou, signed char* __restrict__ labels, float* __restrict__ max_ioa,
                    int* __restrict__ rowmax_bits, int* __restrict__ best_idx) {
  __shared__ float4 sg[kMaxG];
  __shared__ float sarea[kMaxG];
  __shared__ unsigned char sflag[kMaxG];          // bit0 valid, bit1 ignore
  __shared__ float srow[kMaxG];
  const int b = blockIdx.y;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float4 v = gt[(size_t)b * G + g];
    sg[g] = v;
    sarea[g] = box_area4(v);
    sflag[g] = (gt_valid[(size_t)b * G + g] ? 1 : 0) | (gt_ign[(size_t)b * G + g] ? 2 : 0);
    if (PASS == 2) srow[g] = __int_as_float(rowmax_bits[(size_t)b * G + g]);
  }
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = a < A;
  const float4 an = in ? anchors[a] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float area_a = box_area4(an);
  const int lane = threadIdx.x & 31;
  if (PASS == 1) {
    float best = -1.f, ioa = 0.f;
    int bi = 0;
    for (int g = 0; g < G; ++g) {
      const PairQ q = gt_anchor_quality(sg[g], an, sarea[g], area_a);
      const unsigned char f = sflag[g];
      const float v = (f & 1) ? q.iou : -1.f;
      if (v > best) { best = v; bi = g; }
      if (f & 2) ioa = fmaxf(ioa, q.ioa);
      if (f & 1) {                              // running maximum of this GT over all anchors
        float m = in ? q.iou : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0.f) atomicMax(rowmax_bits + (size_t)b * G + g, __float_as_int(m));
      }
    }
    if (in) {
      matched_idx[(size_t)b * A + a] = bi;
      matched_iou[(size_t)b * A + a] = fmaxf(best, 0.f);
      max_ioa[(size_t)b * A + a] = ioa;
    }
  } else {
    if (!in) return;
    const float best = matched_iou[(size_t)b * A + a];
    bool pos = false;
    // matched_iou was clamped at 0: an image without valid GT has best = -1 -> label 0 either way
    bool any_valid = false;
    for (int g = 0; g < G; ++g) {
      if (!(sflag[g] & 1)) continue;
      any_valid = true;
      const PairQ q = gt_anchor_quality(sg[g], an, sarea[g], area_a);
      if (q.iou == srow[g]) {
        pos = true;
        atomicMin(best_idx + (size_t)b * G + g, a);
      }
    }
    labels[(size_t)b * A + a] = (signed char)((pos || (any_valid && best >= fg_thresh)) ? 1 : 0);
  }
}

}  // namespace c3d

using namespace c3d;

extern "C" int32_t c3d_anchor_match(const float* anchors, int64_t A, const float* gt_boxes, const uint8_t* gt_valid,
                                    const uint8_t* gt_ign, int32_t B, int32_t G, float fg_thresh, int64_t* matched_idx,
                                    float* matched_iou, int8_t* labels, float* max_ioa, int32_t* best_idx,
                                    int32_t* rowmax_ws, void* stream) {
  if (!anchors || !gt_boxes || !gt_valid || !gt_ign || !matched_idx || !matched_iou || !labels || !max_ioa || !best_idx ||
      !rowmax_ws)
    return set_error(C3D_EINVAL, "anchor_match: null pointer");
  if (G < 1 || G > kMaxG) return set_error(C3D_EINVAL, "anchor_match: G=%d outside [1, %d]", G, kMaxG);
  if (A < 1 || A > 0x7fffffffLL || B < 1) return set_error(C3D_EINVAL, "anchor_match: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = B * G;
  anchor_match_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(rowmax_ws, best_idx, n, (int)A);
  dim3 grid((unsigned)((A + 255) / 256), (unsigned)B);
  anchor_match_kernel<1><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(anchors), (int)A,
                                              reinterpret_cast<const float4*>(gt_boxes), gt_valid, gt_ign, G, fg_thresh,
                                              reinterpret_cast<long long*>(matched_idx), matched_iou,
                                              reinterpret_cast<signed char*>(labels), max_ioa, rowmax_ws, best_idx);
  anchor_match_kernel<2><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(anchors), (int)A,
                                              reinterpret_cast<const float4*>(gt_boxes), gt_valid, gt_ign, G, fg_thresh,
                                              reinterpret_cast<long long*>(matched_idx), matched_iou,
                                              reinterpret_cast<signed char*>(labels), max_ioa, rowmax_ws, best_idx);
  return check_launch("anchor_match");
}

// ------------------------------------------------------------------------------------------------------------
// RPN losses over all B*A anchors in one pass each way (replaces ~80 ATen launches over (B,A[,4]) tensors):
//   objectness: BCE-with-logits against the IoU-ness target of the matched GT, weighted by that target, positives only
//   localisation: L1 between predicted and encoded GT deltas, weighted by the target, positives only
// (cubercnn/modeling/proposal_generator/rpn.py:108-218 losses(); formulas as omni3d_b200/cubercnn/rpn.py RPNWithIgnore.losses).
// acc[6] += {sum cls, sum loc, #pos, #neg, sum sigmoid over pos, sum sigmoid over non-pos}
namespace c3d {

struct RpnPos { float target; float gd[4]; };

__device__ __forceinline__ RpnPos rpn_targets(const float4 a, const float4 g, float wx, float wy, float ww, float wh) {
  RpnPos r;
  const float iw = fmaxf(fminf(a.z, g.z) - fmaxf(a.x, g.x), 0.f), ih = fmaxf(fminf(a.w, g.w) - fmaxf(a.y, g.y), 0.f);
  const float inter = iw * ih;
  r.target = inter / ((a.z - a.x) * (a.w - a.y) + (g.z - g.x) * (g.w - g.y) - inter);
  const float sw = a.z - a.x, sh = a.w - a.y, scx = a.x + 0.5f * sw, scy = a.y + 0.5f * sh;
  const float tw = g.z - g.x, th = g.w - g.y, tcx = g.x + 0.5f * tw, tcy = g.y + 0.5f * th;
  r.gd[0] = wx * (tcx - scx) / sw; r.gd[1] = wy * (tcy - scy) / sh;
  r.gd[2] = ww * logf(tw / sw); r.gd[3] = wh * logf(th / sh);
  return r;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
rpn_loss_kernel(const float* __restrict__ logits, const float4* __restrict__ deltas, const signed char* __restrict__ labels,
                const long long* __restrict__ matched_idx, const float4* __restrict__ gt, const float4* __restrict__ anchors,
                int B, int A, int G, float wx, float wy, float ww, float wh, float* __restrict__ acc,
                const float* __restrict__ g_cls, const float* __restrict__ g_loc, float* __restrict__ dlogits,
                float4* __restrict__ ddeltas) {
  const long long total = (long long)B * A;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float gc = BWD ? *g_cls : 0.f, gl = BWD ? *g_loc : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int lab = labels[i];
    const float x = logits[i];
    const float sig = 1.f / (1.f + __expf(-x));
    const bool pos = lab == 1;
    float dl = 0.f;
    float4 dd = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!BWD) {
      s[2] += pos ? 1.f : 0.f; s[3] += lab == 0 ? 1.f : 0.f;
      s[4] += pos ? sig : 0.f; s[5] += pos ? 0.f : sig;
    }
    if (pos) {
      const int b = (int)(i / A), a = (int)(i - (long long)b * A);
      const RpnPos t = rpn_targets(anchors[a], gt[(size_t)b * G + matched_idx[i]], wx, wy, ww, wh);
      const float4 d = deltas[i];
      const float e[4] = {d.x - t.gd[0], d.y - t.gd[1], d.z - t.gd[2], d.w - t.gd[3]};
      if (!BWD) {
        const float bce = fmaxf(x, 0.f) - x * t.target + log1pf(__expf(-fabsf(x)));
        s[0] += bce * t.target;
        s[1] += (fabsf(e[0]) + fabsf(e[1]) + fabsf(e[2]) + fabsf(e[3])) * t.target;
      } else {
        const float tt = t.target;
        dl = gc * (sig - tt) * tt;
        auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
        dd = make_float4(gl * tt * sgn(e[0]), gl * tt * sgn(e[1]), gl * tt * sgn(e[2]), gl * tt * sgn(e[3]));
      }
    }
    if (BWD) { dlogits[i] = dl; ddeltas[i] = dd; }
  }
  if (!BWD) {
    __shared__ float sm[6][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      float v = s[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sm[k][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) v += sm[threadIdx.x][j];
      atomicAdd(acc + threadIdx.x, v);
    }
  }
}

}  // namespace c3d

extern "C" int32_t c3d_rpn_loss_fwd(const float* logits, const float* deltas, const int8_t* labels, const int64_t* matched_idx,
                                    const float* gt_boxes, const float* anchors, int32_t B, int64_t A, int32_t G,
                                    const float* weights4_host, float* acc6, void* stream) {
  if (!logits || !deltas || !labels || !matched_idx || !gt_boxes || !anchors || !weights4_host || !acc6)
    return set_error(C3D_EINVAL, "rpn_loss: null pointer");
  if (B < 1 || A < 1 || A > 0x7fffffffLL || G < 1) return set_error(C3D_EINVAL, "rpn_loss: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(acc6, 0, 6 * sizeof(float), st);
  if (e != cudaSuccess) return set_error(C3D_ECUDA, "rpn_loss memset: %s", cudaGetErrorString(e));
  long long blocks = ((long long)B * A + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  rpn_loss_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(logits, reinterpret_cast<const float4*>(deltas),
      reinterpret_cast<const signed char*>(labels), reinterpret_cast<const long long*>(matched_idx),
      reinterpret_cast<const float4*>(gt_boxes), reinterpret_cast<const float4*>(anchors), B, (int)A, G, weights4_host[0],
      weights4_host[1], weights4_host[2], weights4_host[3], acc6, nullptr, nullptr, nullptr, nullptr);
  return check_launch("rpn_loss_fwd");
}

extern "C" int32_t c3d_rpn_loss_bwd(const float* logits, const float* deltas, const int8_t* labels, const int64_t* matched_idx,
                                    const float* gt_boxes, const float* anchors, int32_t B, int64_t A, int32_t G,
                                    const float* weights4_host, const float* g_cls, const float* g_loc, float* dlogits,
                                    float* ddeltas, void* stream) {
  if (!logits || !deltas || !labels || !matched_idx || !gt_boxes || !anchors || !weights4_host || !g_cls || !g_loc || !dlogits ||
      !ddeltas)
    return set_error(C3D_EINVAL, "rpn_loss_bwd: null pointer");
  if (B < 1 || A < 1 || A > 0x7fffffffLL || G < 1) return set_error(C3D_EINVAL, "rpn_loss_bwd: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = ((long long)B * A + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  rpn_loss_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(logits, reinterpret_cast<const float4*>(deltas),
      reinterpret_cast<const signed char*>(labels), reinterpret_cast<const long long*>(matched_idx),
      reinterpret_cast<const float4*>(gt_boxes), reinterpret_cast<const float4*>(anchors), B, (int)A, G, weights4_host[0],
      weights4_host[1], weights4_host[2], weights4_host[3], nullptr, g_cls, g_loc, dlogits,
      reinterpret_cast<float4*>(ddeltas));
  return check_launch("rpn_loss_bwd");
}

// ------------------------------------------------------------------------------------------------------------
// RPN proposal decoding for the top-k candidates of one FPN level (all images): Box2BoxTransform.apply_deltas +
// clip to the image + finite / min-size filter, written at a column offset of the concatenated candidate arrays
// (detectron2 find_top_rpn_proposals as used by cubercnn/modeling/proposal_generator/rpn.py:221-284).
// Replaces ~35 ATen launches per level.  Arithmetic in the operation order of the torch formulation with explicit
// round-to-nearest intrinsics (no FMA contraction), so boxes equal the torch-CUDA result bit for bit.
namespace c3d {

__global__ void rpn_decode_kernel(const long long* __restrict__ topk_idx, const float* __restrict__ topk_score,
                                  const float4* __restrict__ deltas, const float4* __restrict__ anchors,
                                  const float* __restrict__ hw, int B, int K, long long in_stride, long long A, float wx, float wy, float ww,
                                  float wh, float scale_clamp, float min_size, float level, int col0, int Ktot,
                                  float4* __restrict__ boxes, float* __restrict__ key, float* __restrict__ lvl,
                                  int* __restrict__ nvalid, int* __restrict__ maxc_bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K, k = i - b * K;
  const long long a = topk_idx[(long long)b * in_stride + k];
  const float4 d = deltas[(size_t)b * A + a];
  const float4 an = anchors[a];
  const float w = __fsub_rn(an.z, an.x), h = __fsub_rn(an.w, an.y);
  const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
  const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
  float dw = __fdiv_rn(d.z, ww), dh = __fdiv_rn(d.w, wh);
  dw = dw > scale_clamp ? scale_clamp : dw;             // torch.clamp(max=): NaN stays NaN (filtered below)
  dh = dh > scale_clamp ? scale_clamp : dh;
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  const float s = topk_score[(long long)b * in_stride + k];
  const bool finite = isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(s);
  const float H = hw[2 * b], W = hw[2 * b + 1];
  // torch.minimum(boxes.clamp(min=0), lim): NaN propagates through both (irrelevant: filtered by `finite`)
  x1 = fminf(fmaxf(x1, 0.f), W); y1 = fminf(fmaxf(y1, 0.f), H);
  x2 = fminf(fmaxf(x2, 0.f), W); y2 = fminf(fmaxf(y2, 0.f), H);
  const bool keep = finite && (__fsub_rn(x2, x1) > min_size) && (__fsub_rn(y2, y1) > min_size);
  const size_t o = (size_t)b * Ktot + col0 + k;
  boxes[o] = make_float4(x1, y1, x2, y2);
  key[o] = keep ? s : -INFINITY;
  lvl[o] = level;
  if (keep) {
    atomicAdd(nvalid + b, 1);
    atomicMax(maxc_bits + b, __float_as_int(fmaxf(fmaxf(x1, y1), fmaxf(x2, y2))));   // coordinates are >= 0
  }
}

}  // namespace c3d

extern "C" int32_t c3d_rpn_decode_level(const int64_t* topk_idx, const float* topk_score, int64_t in_stride, const float* deltas,
                                        const float* anchors, const float* image_hw, int32_t B, int32_t K, int64_t A,
                                        const float* weights4_host, float scale_clamp, float min_size, int32_t level,
                                        int32_t col0, int32_t Ktot, float* boxes, float* key, float* lvl, int32_t* nvalid,
                                        float* maxc, void* stream) {
  if (!topk_idx || !topk_score || !deltas || !anchors || !image_hw || !weights4_host || !boxes || !key || !lvl || !nvalid || !maxc)
    return set_error(C3D_EINVAL, "rpn_decode: null pointer");
  if (B < 1 || K < 1 || A < 1 || col0 < 0 || col0 + K > Ktot) return set_error(C3D_EINVAL, "rpn_decode: bad sizes");
  const int total = B * K;
  rpn_decode_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(topk_idx), topk_score, reinterpret_cast<const float4*>(deltas),
      reinterpret_cast<const float4*>(anchors), image_hw, B, K, in_stride > 0 ? in_stride : K, A, weights4_host[0], weights4_host[1], weights4_host[2],
      weights4_host[3], scale_clamp, min_size, (float)level, col0, Ktot, reinterpret_cast<float4*>(boxes), key, lvl, nvalid,
      reinterpret_cast<int*>(maxc));
  return check_launch("rpn_decode");
}
