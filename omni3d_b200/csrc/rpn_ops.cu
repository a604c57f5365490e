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
