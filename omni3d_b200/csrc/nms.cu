// nms.cu — batched greedy NMS for all images of a batch in two launches, fully on the device.
//
// Replaces torchvision.ops.nms (reached through detectron2 batched_nms in find_top_rpn_proposals,
// SURVEY.md A.3; configs/Base.yaml:51-54) whose CUDA path copies the suppression mask to the host and
// scans it there (one host sync per image).  Boxes arrive sorted by score (descending) per image and
// already shifted by the per-level "coordinate trick" offsets, so IoU > thr (same fp32 formula as
// torchvision's devIoU) reproduces its keep list exactly.
//   kernel 1: 64x64 tiles of the upper-triangular suppression bit matrix (one uint64 per row/tile);
//   kernel 2: one warp per image walks the rows in score order keeping the live "removed" bitset in
//             registers (4 words per lane, up to 8192 candidates), writes the first `max_keep` survivors.
#include "c3d_common.cuh"

namespace c3d {

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  float inter = width * height;
  float sa = (a.z - a.x) * (a.w - a.y);
  float sb = (b.z - b.x) * (b.w - b.y);
  return (inter / (sa + sb - inter)) > thr;
}

__device__ __forceinline__ float4 shift(float4 b, float off) {
  return make_float4(b.x + off, b.y + off, b.z + off, b.w + off);
}

// cats (may be null): per-box category as float; maxc: per-image max coordinate.  Images with
// 4*nvalid <= trick_max_numel use torchvision's coordinate trick (boxes shifted by cat*(maxc+1)), larger
// ones plain per-category suppression — exactly torchvision.ops.batched_nms's two code paths.
__global__ void nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ nvalid, int n, int words,
                                float thr, const float* __restrict__ cats, const float* __restrict__ maxc,
                                int trick_max_numel, unsigned long long* __restrict__ mask) {
  const int b = blockIdx.z, row_blk = blockIdx.y, col_blk = blockIdx.x;
  if (col_blk < row_blk) return;
  const int nv = nvalid[b];
  const int row0 = row_blk * 64, col0 = col_blk * 64;
  if (row0 >= nv || col0 >= nv) return;
  __shared__ float4 cb[64];
  __shared__ float cc[64];
  const int t = threadIdx.x;
  const float4* bx = boxes + (size_t)b * n;
  const float* cx = cats ? cats + (size_t)b * n : nullptr;
  const bool trick = cx && (4 * nv <= trick_max_numel);
  const float scale = trick ? (maxc[b] + 1.0f) : 0.f;
  if (col0 + t < nv) {
    float c = cx ? cx[col0 + t] : 0.f;
    cc[t] = c;
    cb[t] = trick ? shift(bx[col0 + t], c * scale) : bx[col0 + t];
  }
  __syncthreads();
  const int i = row0 + t;
  if (i < nv) {
    const float myc = cx ? cx[i] : 0.f;
    const float4 me = trick ? shift(bx[i], myc * scale) : bx[i];
    unsigned long long bits = 0;
    const int ncol = min(64, nv - col0);
    const int start = (row_blk == col_blk) ? t + 1 : 0;
    for (int j = start; j < ncol; ++j)
      // different categories never suppress each other: per-category mode by definition, coordinate-trick mode because
      // the shifted boxes are disjoint (IoU 0 <= thr) — skip the IoU arithmetic for ~(levels-1)/levels of the pairs
      if (cc[j] == myc && iou_gt(me, cb[j], thr)) bits |= 1ULL << j;
    mask[((size_t)b * n + i) * words + col_blk] = bits;
  }
}

__global__ void nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ nvalid, int n,
                                int words, int max_keep, int* __restrict__ keep_idx, int* __restrict__ keep_cnt) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int nv = nvalid[b];
  unsigned long long r0 = 0, r1 = 0, r2 = 0, r3 = 0;     // removed bits: word w lives on lane w%32, slot w/32
  int cnt = 0;
  int* out = keep_idx + (size_t)b * max_keep;
  for (int i = 0; i < nv && cnt < max_keep; ++i) {
    const int w = i >> 6, slot = w >> 5, owner = w & 31;
    unsigned long long word = slot == 0 ? r0 : (slot == 1 ? r1 : (slot == 2 ? r2 : r3));
    word = __shfl_sync(0xffffffffu, word, owner);
    if (!((word >> (i & 63)) & 1ULL)) {
      if (lane == 0) out[cnt] = i;
      ++cnt;
      const unsigned long long* row = mask + ((size_t)b * n + i) * words;
      // only words >= w were written by kernel 1 (upper triangle)
      int ww = lane;
      if (ww >= w && ww < words) r0 |= row[ww];
      ww = lane + 32; if (ww >= w && ww < words) r1 |= row[ww];
      ww = lane + 64; if (ww >= w && ww < words) r2 |= row[ww];
      ww = lane + 96; if (ww >= w && ww < words) r3 |= row[ww];
    }
  }
  for (int k = cnt + lane; k < max_keep; k += 32) out[k] = -1;
  if (lane == 0) keep_cnt[b] = cnt;
}

}  // namespace c3d

extern "C" size_t c3d_nms_workspace_bytes(int32_t B, int32_t n) {
  if (B < 0 || n < 0) return 0;
  size_t words = (size_t)(n + 63) / 64;
  return (size_t)B * n * words * 8 + 256;
}

extern "C" int32_t c3d_nms_batched(const float* boxes, const int32_t* nvalid, const float* cats, const float* maxc,
                                   int32_t trick_max_numel, int32_t B, int32_t n, float iou_thresh,
                                   int32_t max_keep, int32_t* keep_idx, int32_t* keep_cnt, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  using namespace c3d;
  if (B == 0 || n == 0) return C3D_OK;
  if (!boxes || !nvalid || !keep_idx || !keep_cnt || !workspace) return set_error(C3D_EINVAL, "nms: null pointer");
  if (n > 8192) return set_error(C3D_EINVAL, "nms: at most 8192 candidates per image (got %d)", n);
  if (workspace_bytes < c3d_nms_workspace_bytes(B, n)) return set_error(C3D_EWORKSPACE, "nms: workspace too small");
  const int words = (n + 63) / 64;
  cudaStream_t st = (cudaStream_t)stream;
  // rows whose diagonal tile is skipped (beyond nvalid) are never read; no memset needed because the scan
  // only reads words >= i/64 of rows i < nvalid, all of which kernel 1 writes when col0 < nvalid.
  dim3 grid(words, words, B);
  if (cats && !maxc) return set_error(C3D_EINVAL, "nms: cats given without maxc");
  nms_mask_kernel<<<grid, 64, 0, st>>>((const float4*)boxes, nvalid, n, words, iou_thresh, cats, maxc,
                                       trick_max_numel, (unsigned long long*)workspace);
  nms_scan_kernel<<<B, 32, 0, st>>>((const unsigned long long*)workspace, nvalid, n, words, max_keep, keep_idx,
                                    keep_cnt);
  return check_launch("nms_batched");
}
