// roi_align.cu — multi-level ROIAlign (aligned=True, adaptive sampling) on NHWC bf16 feature maps.
//
// Replaces detectron2 ROIPooler / ROIAlignV2 as used by the box pooler (roi_heads.py:267) and the cube
// pooler (roi_heads.py:362, built :166-171): 7x7 bins, sampling_ratio 0, FPN levels p2..p6 chosen by
// floor(4 + log2(sqrt(area)/224 + 1e-8)) clamped to [2,6].
// One warp per output bin; each lane owns 8 consecutive channels, so every bilinear tap is a single
// fully-coalesced 16-byte-per-lane read of the pixel's channel vector (no tensor cores: gather work).
// Backward scatters with vector fp32 atomics into per-level fp32 gradient maps.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "c3d_common.cuh"

namespace c3d {
using bf16 = __nv_bfloat16;

struct RoiLevels {
  const bf16* feat[5];
  float* grad[5];
  int H[5], W[5];
  float scale[5];
  int num_levels;
  int num_images;
};

struct Tap { int y0, y1, x0, x1; float w1, w2, w3, w4; bool valid; };

__device__ __forceinline__ Tap make_tap(float y, float x, int H, int W) {
  Tap t;
  t.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  t.y0 = yl; t.y1 = yh; t.x0 = xl; t.x1 = xh;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
  return t;
}

__device__ __forceinline__ void acc8(float (&a)[8], const bf16* p, float w) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); a[2 * i] += w * f.x; a[2 * i + 1] += w * f.y; }
}

// rois: [R][6] = (batch, level, x1, y1, x2, y2) fp32
template <bool BWD>
__global__ void roi_align_kernel(RoiLevels L, const float* __restrict__ rois, int R, int C, int PH, int PW,
                                 bf16* __restrict__ out, const bf16* __restrict__ dout) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long nbins = (long long)R * PH * PW;
  for (long long bin = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); bin < nbins;
       bin += (long long)gridDim.x * warps_per_block) {
    const int pw = (int)(bin % PW), ph = (int)((bin / PW) % PH), r = (int)(bin / ((long long)PW * PH));
    const float* roi = rois + (size_t)r * 6;
    // a NaN / Inf box (diverging step) must not become an out-of-range level or image index: such RoIs pool zeros
    const float fb = roi[0], fl = roi[1];
    const bool sane = fb >= 0.f && (L.num_images <= 0 || fb < (float)L.num_images) && fl >= 0.f && fl < (float)L.num_levels &&
                      isfinite(roi[2]) && isfinite(roi[3]) && isfinite(roi[4]) && isfinite(roi[5]);
    if (!sane) {
      if (!BWD)
        for (int c = lane * 8; c < C; c += 256) *reinterpret_cast<uint4*>(out + (size_t)bin * C + c) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const int b = (int)fb, lvl = (int)fl;
    const float sc = L.scale[lvl];
    const int H = L.H[lvl], W = L.W[lvl];
    const float sw = roi[2] * sc - 0.5f, sh = roi[3] * sc - 0.5f;
    const float rw = roi[4] * sc - 0.5f - sw, rh = roi[5] * sc - 0.5f - sh;
    const float bh = rh / PH, bw = rw / PW;
    const int gh = (int)ceilf(rh / PH), gw = (int)ceilf(rw / PW);
    const float cnt = fmaxf((float)(gh * gw), 1.f);
    const bf16* base = L.feat[lvl] + (size_t)b * H * W * C;
    float* gbase = BWD ? L.grad[lvl] + (size_t)b * H * W * C : nullptr;
    for (int c = lane * 8; c < C; c += 256) {
      float a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = 0.f;
      float g[8];
      if (BWD) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(dout + (size_t)bin * C + c));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); g[2 * i] = f.x / cnt; g[2 * i + 1] = f.y / cnt; }
      }
      for (int iy = 0; iy < gh; ++iy) {
        const float y = sh + ph * bh + (iy + 0.5f) * bh / (float)gh;
        for (int ix = 0; ix < gw; ++ix) {
          const float x = sw + pw * bw + (ix + 0.5f) * bw / (float)gw;
          Tap t = make_tap(y, x, H, W);
          if (!t.valid) continue;
          const size_t o1 = ((size_t)t.y0 * W + t.x0) * C + c, o2 = ((size_t)t.y0 * W + t.x1) * C + c;
          const size_t o3 = ((size_t)t.y1 * W + t.x0) * C + c, o4 = ((size_t)t.y1 * W + t.x1) * C + c;
          if (!BWD) {
            acc8(a, base + o1, t.w1); acc8(a, base + o2, t.w2); acc8(a, base + o3, t.w3); acc8(a, base + o4, t.w4);
          } else {
            const size_t offs[4] = {o1, o2, o3, o4};
            const float ws[4] = {t.w1, t.w2, t.w3, t.w4};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4* dst = reinterpret_cast<float4*>(gbase + offs[j]);
              atomicAdd(dst, make_float4(g[0] * ws[j], g[1] * ws[j], g[2] * ws[j], g[3] * ws[j]));
              atomicAdd(dst + 1, make_float4(g[4] * ws[j], g[5] * ws[j], g[6] * ws[j], g[7] * ws[j]));
            }
          }
        }
      }
      if (!BWD) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(a[2 * i] / cnt, a[2 * i + 1] / cnt);
        *reinterpret_cast<uint4*>(out + (size_t)bin * C + c) = u;
      }
    }
  }
}

// Backward with separable tap weights.  The samples of a bin form a product grid (gh x gw) and a bilinear weight is
// wy * wx, validity is (y valid) && (x valid): the gradient a bin sends to pixel (r, q) is
//     g / count * (sum_iy wy_iy(r)) * (sum_ix wx_ix(q)).
// The two 1-D tables are accumulated across the warp's lanes (lane l holds row/column base+l, up to 32 each) and
// broadcast with shuffles, so a bin issues rows x cols vector atomics instead of 4 * gh * gw (about half for the
// common 2x2 grids, a quarter for 4x4).  Bins wider than 32 feature pixels fall back to per-sample scatter.
__device__ __forceinline__ bool axis_tap(float v, int n, int* lo, int* hi, float* wl, float* wh) {
  if (v < -1.0f || v > (float)n) return false;
  if (v <= 0.f) v = 0.f;
  int l = (int)v, h;
  if (l >= n - 1) { h = l = n - 1; v = (float)l; } else h = l + 1;
  const float fl = v - l;
  *lo = l; *hi = h; *wh = fl; *wl = 1.f - fl;
  return true;
}

__global__ void roi_align_bwd_sep_kernel(RoiLevels L, const float* __restrict__ rois, int R, int C, int PH, int PW,
                                         const bf16* __restrict__ dout) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long nbins = (long long)R * PH * PW;
  for (long long bin = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); bin < nbins;
       bin += (long long)gridDim.x * warps_per_block) {
    const int pw = (int)(bin % PW), ph = (int)((bin / PW) % PH), r = (int)(bin / ((long long)PW * PH));
    const float* roi = rois + (size_t)r * 6;
    // a NaN / Inf box (diverging step) must not become an out-of-range level or image index: such RoIs pool zeros
    const float fb = roi[0], fl = roi[1];
    const bool sane = fb >= 0.f && (L.num_images <= 0 || fb < (float)L.num_images) && fl >= 0.f && fl < (float)L.num_levels &&
                      isfinite(roi[2]) && isfinite(roi[3]) && isfinite(roi[4]) && isfinite(roi[5]);
    if (!sane) continue;
    const int b = (int)fb, lvl = (int)fl;
    const float sc = L.scale[lvl];
    const int H = L.H[lvl], W = L.W[lvl];
    const float sw = roi[2] * sc - 0.5f, sh = roi[3] * sc - 0.5f;
    const float rw = roi[4] * sc - 0.5f - sw, rh = roi[5] * sc - 0.5f - sh;
    const float bh = rh / PH, bw = rw / PW;
    const int gh = (int)ceilf(rh / PH), gw = (int)ceilf(rw / PW);
    const float cnt = fmaxf((float)(gh * gw), 1.f);
    float* gbase = L.grad[lvl] + (size_t)b * H * W * C;
    // 1-D weight tables, one entry per lane
    int ybase = -1, xbase = -1, ymax = -1, xmax = -1;
    float wy = 0.f, wx = 0.f;
    bool fits = true;
    for (int iy = 0; iy < gh; ++iy) {
      int lo, hi; float wl, wh;
      if (!axis_tap(sh + ph * bh + (iy + 0.5f) * bh / (float)gh, H, &lo, &hi, &wl, &wh)) continue;
      if (ybase < 0) ybase = lo;
      if (hi - ybase > 31) { fits = false; break; }
      ymax = hi;
      if (lane == lo - ybase) wy += wl;
      if (lane == hi - ybase) wy += wh;
    }
    for (int ix = 0; ix < gw && fits; ++ix) {
      int lo, hi; float wl, wh;
      if (!axis_tap(sw + pw * bw + (ix + 0.5f) * bw / (float)gw, W, &lo, &hi, &wl, &wh)) continue;
      if (xbase < 0) xbase = lo;
      if (hi - xbase > 31) { fits = false; break; }
      xmax = hi;
      if (lane == lo - xbase) wx += wl;
      if (lane == hi - xbase) wx += wh;
    }
    if (fits && (ybase < 0 || xbase < 0)) continue;          // no valid sample: nothing to scatter
    for (int c = lane * 8; c < ((C + 255) / 256) * 256; c += 256) {   // uniform trip count (shuffles below)
      const bool cin = c < C;
      float g[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = 0.f;
      if (cin) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(dout + (size_t)bin * C + c));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); g[2 * i] = f.x / cnt; g[2 * i + 1] = f.y / cnt; }
      }
      if (fits) {
        const int nr = ymax - ybase + 1, nc = xmax - xbase + 1;
        for (int rr = 0; rr < nr; ++rr) {
          const float wr = __shfl_sync(0xffffffffu, wy, rr);
          if (wr == 0.f) continue;
          for (int q = 0; q < nc; ++q) {
            const float w = wr * __shfl_sync(0xffffffffu, wx, q);
            if (w == 0.f || !cin) continue;
            float4* dst = reinterpret_cast<float4*>(gbase + ((size_t)(ybase + rr) * W + (xbase + q)) * C + c);
            atomicAdd(dst, make_float4(g[0] * w, g[1] * w, g[2] * w, g[3] * w));
            atomicAdd(dst + 1, make_float4(g[4] * w, g[5] * w, g[6] * w, g[7] * w));
          }
        }
      } else if (cin) {
        for (int iy = 0; iy < gh; ++iy) {
          const float y = sh + ph * bh + (iy + 0.5f) * bh / (float)gh;
          for (int ix = 0; ix < gw; ++ix) {
            const float x = sw + pw * bw + (ix + 0.5f) * bw / (float)gw;
            Tap t = make_tap(y, x, H, W);
            if (!t.valid) continue;
            const size_t offs[4] = {((size_t)t.y0 * W + t.x0) * C + c, ((size_t)t.y0 * W + t.x1) * C + c,
                                    ((size_t)t.y1 * W + t.x0) * C + c, ((size_t)t.y1 * W + t.x1) * C + c};
            const float ws[4] = {t.w1, t.w2, t.w3, t.w4};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4* dst = reinterpret_cast<float4*>(gbase + offs[j]);
              atomicAdd(dst, make_float4(g[0] * ws[j], g[1] * ws[j], g[2] * ws[j], g[3] * ws[j]));
              atomicAdd(dst + 1, make_float4(g[4] * ws[j], g[5] * ws[j], g[6] * ws[j], g[7] * ws[j]));
            }
          }
        }
      }
    }
  }
}

static int32_t run(bool bwd, const c3d_roi_levels* lv, const float* rois, int R, int C, int PH, int PW, void* out,
                   const void* dout, cudaStream_t st) {
  if (!lv || lv->num_levels < 1 || lv->num_levels > 5 || C % 8 != 0) return set_error(C3D_EINVAL, "roi_align: bad args");
  if (R == 0) return C3D_OK;
  RoiLevels L;
  L.num_levels = lv->num_levels;
  L.num_images = lv->num_images;
  for (int i = 0; i < 5; ++i) {
    L.feat[i] = (const bf16*)lv->feat[i]; L.grad[i] = (float*)lv->grad[i];
    L.H[i] = lv->H[i]; L.W[i] = lv->W[i]; L.scale[i] = lv->scale[i];
  }
  long long nbins = (long long)R * PH * PW;
  long long blocks = (nbins + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  static const bool per_sample = getenv("C3D_ROI_PER_SAMPLE") != nullptr;
  if (bwd && !per_sample) roi_align_bwd_sep_kernel<<<(unsigned)blocks, 256, 0, st>>>(L, rois, R, C, PH, PW, (const bf16*)dout);
  else if (bwd) roi_align_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(L, rois, R, C, PH, PW, nullptr, (const bf16*)dout);
  else roi_align_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(L, rois, R, C, PH, PW, (bf16*)out, nullptr);
  return check_launch("roi_align");
}
}  // namespace c3d

extern "C" int32_t c3d_roi_align_fwd(const c3d_roi_levels* levels, const float* rois, int32_t R, int32_t C,
                                     int32_t pooled_h, int32_t pooled_w, void* out, void* stream) {
  if (!rois && R > 0) return c3d::set_error(C3D_EINVAL, "roi_align: null rois");
  return c3d::run(false, levels, rois, R, C, pooled_h, pooled_w, out, nullptr, (cudaStream_t)stream);
}
extern "C" int32_t c3d_roi_align_bwd(const c3d_roi_levels* levels, const float* rois, int32_t R, int32_t C,
                                     int32_t pooled_h, int32_t pooled_w, const void* dout, void* stream) {
  if (!rois && R > 0) return c3d::set_error(C3D_EINVAL, "roi_align: null rois");
  return c3d::run(true, levels, rois, R, C, pooled_h, pooled_w, nullptr, dout, (cudaStream_t)stream);
}
