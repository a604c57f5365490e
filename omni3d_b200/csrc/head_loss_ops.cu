// head_loss_ops.cu — loss assembly of the box head and the glue around the cube head's fused loss kernel.
//
//   box_loss_fwd / bwd    FastRCNNOutputs.losses (cubercnn/modeling/roi_heads/fast_rcnn.py:145-194, box_reg_loss :196-260):
//                         mean softmax cross-entropy over the sampled RoIs + L1 on the foreground RoIs' class-specific deltas
//                         (Box2BoxTransform weights (10,10,5,5)), normalised by the number of sampled RoIs, plus the three
//                         logged accuracies — straight from the fused predictor GEMM's fp32 output rows
//                         [K+1 scores | 4K deltas | pad].  One warp per RoI; replaces ~40 ATen launches and their tape.
//   cube_gather           roi_heads.py:409-461 + :372-404: per-RoI gather of the predicted class's 13 raw head outputs from the
//                         fused cube-predictor GEMM output and assembly of the 28 constants c3d_cube_loss_fwd/bwd read
//                         (box, intrinsics scaled to the network input, virtual->real depth factor, dimension prior, GT).
//   cube_reduce_fwd / bwd safely_reduce_losses (roi_heads.py:932-941) of the 6 loss columns over finite, valid RoIs + the logged
//                         statistics (:690-743), and the matching gradient of the per-RoI rows.
//   cube_scatter          gradient of the gather: d raw13 -> the predicted class's columns of d pred (zero elsewhere).
#include <stdint.h>
#include "c3d_common.cuh"

namespace c3d {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// get_deltas of Box2BoxTransform (detectron2) for one (proposal, gt) pair
__device__ __forceinline__ float4 box_deltas(const float4 s, const float4 t, const float4 w) {
  const float sw = s.z - s.x, sh = s.w - s.y, scx = s.x + 0.5f * sw, scy = s.y + 0.5f * sh;
  const float tw = t.z - t.x, th = t.w - t.y, tcx = t.x + 0.5f * tw, tcy = t.y + 0.5f * th;
  return make_float4(w.x * (tcx - scx) / sw, w.y * (tcy - scy) / sh, w.z * logf(tw / sw), w.w * logf(th / sh));
}

// acc[0] sum CE over valid, [1] sum L1 over fg, [2] #valid, [3] #fg, [4] #(argmax == class) over valid,
// [5] #(argmax == class) over fg, [6] #(argmax == K) over fg
template <bool BWD>
__global__ void box_loss_kernel(const float* __restrict__ pred, int ld, const long long* __restrict__ classes,
                                const unsigned char* __restrict__ valid, const float4* __restrict__ boxes,
                                const float4* __restrict__ gt_boxes, int R, int K, float4 w, float* __restrict__ acc,
                                const float* __restrict__ g /*[2] dL/dloss_cls, dL/dloss_box*/, float* __restrict__ dpred) {
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  float a[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float inv_n = 0.f, gc = 0.f, gb = 0.f;
  if (BWD) { inv_n = 1.f / fmaxf(acc[2], 1.f); gc = g[0] * inv_n; gb = g[1] * inv_n; }
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < R; r += gridDim.x * warps) {
    const float* row = pred + (size_t)r * ld;
    const bool v = valid[r] != 0;
    const long long c = classes[r];
    const int cc = (int)(c < 0 ? 0 : c);                    // torch path: cls.clamp(min=0) for the (masked) CE
    const bool fg = v && c >= 0 && c < K;
    // softmax statistics over the K+1 scores
    float m = -INFINITY;
    for (int k = lane; k <= K; k += 32) m = fmaxf(m, row[k]);
    int am = 0x7fffffff;                                    // first arg-max (torch.argmax)
    const float mm = warp_max(m);
    for (int k = lane; k <= K; k += 32) if (row[k] == mm) am = min(am, k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) am = min(am, __shfl_xor_sync(0xffffffffu, am, o));
    float s = 0.f;
    for (int k = lane; k <= K; k += 32) s += __expf(row[k] - mm);
    s = warp_sum(s);
    const float lse = mm + __logf(s);
    if (!BWD) {
      if (lane == 0) {
        if (v) { a[0] += lse - row[cc]; a[2] += 1.f; a[4] += am == cc ? 1.f : 0.f; }
        if (fg) {
          const float4 t = box_deltas(boxes[r], gt_boxes[r], w);
          const float* d = row + (K + 1) + 4 * cc;
          a[1] += fabsf(d[0] - t.x) + fabsf(d[1] - t.y) + fabsf(d[2] - t.z) + fabsf(d[3] - t.w);
          a[3] += 1.f; a[5] += am == cc ? 1.f : 0.f; a[6] += am == K ? 1.f : 0.f;
        }
      }
    } else {
      float* drow = dpred + (size_t)r * ld;
      for (int k = lane; k < ld; k += 32) {
        float dv = 0.f;
        if (k <= K) { if (v) dv = gc * (__expf(row[k] - lse) - (k == cc ? 1.f : 0.f)); }
        drow[k] = dv;
      }
      __syncwarp();
      if (fg && lane < 4) {
        const float4 t = box_deltas(boxes[r], gt_boxes[r], w);
        const float tt = lane == 0 ? t.x : (lane == 1 ? t.y : (lane == 2 ? t.z : t.w));
        const float diff = row[(K + 1) + 4 * cc + lane] - tt;
        drow[(K + 1) + 4 * cc + lane] = gb * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
      }
    }
  }
  if (!BWD && lane == 0) {
#pragma unroll
    for (int i = 0; i < 7; ++i) if (a[i] != 0.f) atomicAdd(acc + i, a[i]);
  }
}

// ---- cube head glue --------------------------------------------------------------------------------------------------
// pred (n, ld) fp32 = [deltas 2K | dims 3K | pose 6K | z K | uncert K | pad] (the fused predictor's column order)
// raw13 = delta x,y | z | dims W,H,L | pose6 | uncert;  aux28 as documented in cube_loss.cu
__global__ void cube_gather_kernel(const float* __restrict__ pred, int ld, const long long* __restrict__ classes,
                                   const float4* __restrict__ boxes, const float* __restrict__ meta /*[B][12]: h,w,ratio,K(9)*/,
                                   const float* __restrict__ priors /*[K][3] (W,H,L mean)*/, const float* __restrict__ gt3 /*[n][9]*/,
                                   const float* __restrict__ gtR /*[n][9]*/, int n, int per_image, int K, float virtual_focal,
                                   float* __restrict__ raw /*[n][13]*/, float* __restrict__ aux /*[n][28]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long c = classes[i];
  c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
  const float* p = pred + (size_t)i * ld;
  float* r = raw + (size_t)i * 13;
  r[0] = p[2 * c]; r[1] = p[2 * c + 1];
  r[2] = p[11 * K + c];
#pragma unroll
  for (int t = 0; t < 3; ++t) r[3 + t] = p[2 * K + 3 * c + t];
#pragma unroll
  for (int t = 0; t < 6; ++t) r[6 + t] = p[5 * K + 6 * c + t];
  r[12] = p[12 * K + c];
  const float* m = meta + (size_t)(i / per_image) * 12;
  const float h = m[0], ratio = m[2];
  float* a = aux + (size_t)i * 28;
  const float4 b = boxes[i];
  a[0] = b.x; a[1] = b.y; a[2] = b.z; a[3] = b.w;
  // roi_heads.py:374-378: K / ratio with K[2,2] = 1;  :380-404: virtual -> real = (h * fy) / (virtual_focal * (h * ratio))
  a[4] = m[3] / ratio; a[5] = m[7] / ratio; a[6] = m[5] / ratio; a[7] = m[8] / ratio;
  a[8] = (h * m[7]) / (virtual_focal * (h * ratio));
#pragma unroll
  for (int t = 0; t < 3; ++t) a[9 + t] = priors[3 * c + t];
#pragma unroll
  for (int t = 0; t < 6; ++t) a[12 + t] = gt3[(size_t)i * 9 + t];
#pragma unroll
  for (int t = 0; t < 9; ++t) a[18 + t] = gtR[(size_t)i * 9 + t];
  a[27] = 0.f;
}

// rows (n,10) -> out[0..5] = mean over finite & valid rows of columns 0..5 (column 5 additionally needs < inf, which
// finite already implies), out[6..10] = masked means of |z-gz|, dims err, xy err, (|z-gz| < 0.2), conf; cnt[0..5] = counts
__global__ void cube_reduce_kernel(const float* __restrict__ rows, const unsigned char* __restrict__ valid, int n,
                                   float* __restrict__ sums /*[12]*/, float* __restrict__ cnts /*[7]*/) {
  float s[11], c[7];
#pragma unroll
  for (int k = 0; k < 11; ++k) s[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 7; ++k) c[k] = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!valid[i]) continue;
    const float* r = rows + (size_t)i * 10;
#pragma unroll
    for (int k = 0; k < 6; ++k) { const float v = r[k]; if (isfinite(v)) { s[k] += v; c[k] += 1.f; } }
    s[6] += r[6]; s[7] += r[7]; s[8] += r[8]; s[9] += r[6] < 0.20f ? 1.f : 0.f; s[10] += r[9];
    c[6] += 1.f;
  }
#pragma unroll
  for (int k = 0; k < 11; ++k) { s[k] = warp_sum(s[k]); }
#pragma unroll
  for (int k = 0; k < 7; ++k) { c[k] = warp_sum(c[k]); }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 11; ++k) if (s[k] != 0.f) atomicAdd(sums + k, s[k]);
#pragma unroll
    for (int k = 0; k < 7; ++k) if (c[k] != 0.f) atomicAdd(cnts + k, c[k]);
  }
}

// d rows[i][k] = g[k] / max(cnt[k], 1) for finite, valid entries of the 6 loss columns (0 elsewhere);
// then d raw13 (from cube_loss_bwd) is scattered into the predicted class's columns of d pred by cube_scatter_kernel
__global__ void cube_reduce_bwd_kernel(const float* __restrict__ rows, const unsigned char* __restrict__ valid, int n,
                                       const float* __restrict__ cnts, const float* __restrict__ g /*[6]*/,
                                       float* __restrict__ drows /*[n][6]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool v = valid[i] != 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float x = rows[(size_t)i * 10 + k];
    drows[(size_t)i * 6 + k] = (v && isfinite(x)) ? g[k] / fmaxf(cnts[k], 1.f) : 0.f;
  }
}

__global__ void cube_scatter_kernel(const float* __restrict__ draw /*[n][13]*/, const long long* __restrict__ classes, int n, int K,
                                    int ld, float* __restrict__ dpred /*[n][ld]*/) {
  const int i = blockIdx.x;
  float* d = dpred + (size_t)i * ld;
  for (int k = threadIdx.x; k < ld; k += blockDim.x) d[k] = 0.f;
  __syncthreads();
  if (threadIdx.x < 13) {
    long long c = classes[i];
    c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
    const int t = threadIdx.x;
    int col;
    if (t < 2) col = 2 * (int)c + t;
    else if (t == 2) col = 11 * K + (int)c;
    else if (t < 6) col = 2 * K + 3 * (int)c + (t - 3);
    else if (t < 12) col = 5 * K + 6 * (int)c + (t - 6);
    else col = 12 * K + (int)c;
    d[col] = draw[(size_t)i * 13 + t];
  }
}

}  // namespace c3d

using namespace c3d;

extern "C" int32_t c3d_box_loss_fwd(const float* pred, int32_t ld, const int64_t* classes, const uint8_t* valid, const float* boxes,
                                    const float* gt_boxes, int32_t R, int32_t K, const float* weights4_host, float* acc7,
                                    void* stream) {
  if (!pred || !classes || !valid || !boxes || !gt_boxes || !weights4_host || !acc7 || K < 1 || ld < 5 * K + 1)
    return set_error(C3D_EINVAL, "box_loss_fwd: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(acc7, 0, 8 * sizeof(float), st);
  if (e != cudaSuccess) return set_error(C3D_ECUDA, "box_loss_fwd: %s", cudaGetErrorString(e));
  if (R <= 0) return C3D_OK;
  const float4 w = make_float4(weights4_host[0], weights4_host[1], weights4_host[2], weights4_host[3]);
  int blocks = (R + 7) / 8;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  box_loss_kernel<false><<<blocks, 256, 0, st>>>(pred, ld, reinterpret_cast<const long long*>(classes), valid,
                                                  reinterpret_cast<const float4*>(boxes), reinterpret_cast<const float4*>(gt_boxes), R,
                                                  K, w, acc7, nullptr, nullptr);
  return check_launch("box_loss_fwd");
}

extern "C" int32_t c3d_box_loss_bwd(const float* pred, int32_t ld, const int64_t* classes, const uint8_t* valid, const float* boxes,
                                    const float* gt_boxes, int32_t R, int32_t K, const float* weights4_host, const float* acc7,
                                    const float* g2, float* dpred, void* stream) {
  if (!pred || !classes || !valid || !boxes || !gt_boxes || !weights4_host || !acc7 || !g2 || !dpred || K < 1 || ld < 5 * K + 1)
    return set_error(C3D_EINVAL, "box_loss_bwd: bad args");
  if (R <= 0) return C3D_OK;
  const float4 w = make_float4(weights4_host[0], weights4_host[1], weights4_host[2], weights4_host[3]);
  int blocks = (R + 7) / 8;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  box_loss_kernel<true><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, ld, reinterpret_cast<const long long*>(classes), valid, reinterpret_cast<const float4*>(boxes),
      reinterpret_cast<const float4*>(gt_boxes), R, K, w, const_cast<float*>(acc7), g2, dpred);
  return check_launch("box_loss_bwd");
}

extern "C" int32_t c3d_cube_gather(const float* pred, int32_t ld, const int64_t* classes, const float* boxes, const float* meta12,
                                   const float* priors, const float* gt3, const float* gtR, int32_t n, int32_t per_image, int32_t K,
                                   float virtual_focal, float* raw13, float* aux28, void* stream) {
  if (!pred || !classes || !boxes || !meta12 || !priors || !gt3 || !gtR || !raw13 || !aux28 || per_image < 1 || ld < 13 * K)
    return set_error(C3D_EINVAL, "cube_gather: bad args");
  if (n <= 0) return C3D_OK;
  cube_gather_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, ld, reinterpret_cast<const long long*>(classes), reinterpret_cast<const float4*>(boxes), meta12, priors, gt3, gtR, n,
      per_image, K, virtual_focal, raw13, aux28);
  return check_launch("cube_gather");
}

extern "C" int32_t c3d_cube_reduce_fwd(const float* rows10, const uint8_t* valid, int32_t n, float* sums12, float* cnts8, void* stream) {
  if (!rows10 || !valid || !sums12 || !cnts8) return set_error(C3D_EINVAL, "cube_reduce_fwd: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(sums12, 0, 12 * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(cnts8, 0, 8 * sizeof(float), st);
  if (e != cudaSuccess) return set_error(C3D_ECUDA, "cube_reduce_fwd: %s", cudaGetErrorString(e));
  if (n <= 0) return C3D_OK;
  int blocks = (n + 255) / 256;
  if (blocks > kNumSMs) blocks = kNumSMs;
  cube_reduce_kernel<<<blocks, 256, 0, st>>>(rows10, valid, n, sums12, cnts8);
  return check_launch("cube_reduce_fwd");
}

extern "C" int32_t c3d_cube_reduce_bwd(const float* rows10, const uint8_t* valid, int32_t n, const float* cnts8, const float* g6,
                                       float* drows6, void* stream) {
  if (!rows10 || !valid || !cnts8 || !g6 || !drows6) return set_error(C3D_EINVAL, "cube_reduce_bwd: bad args");
  if (n <= 0) return C3D_OK;
  cube_reduce_bwd_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows10, valid, n, cnts8, g6, drows6);
  return check_launch("cube_reduce_bwd");
}

extern "C" int32_t c3d_cube_scatter(const float* draw13, const int64_t* classes, int32_t n, int32_t K, int32_t ld, float* dpred,
                                    void* stream) {
  if (!draw13 || !classes || !dpred || ld < 13 * K) return set_error(C3D_EINVAL, "cube_scatter: bad args");
  if (n <= 0) return C3D_OK;
  cube_scatter_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(draw13, reinterpret_cast<const long long*>(classes), n, K, ld,
                                                                         dpred);
  return check_launch("cube_scatter");
}
