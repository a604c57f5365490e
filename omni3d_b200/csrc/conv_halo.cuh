// conv_halo.cuh — "thin-channel" stride-1 convolutions (Cin, Cout <= 32) on tcgen05 with a rolling halo of input
// rows in shared memory.  Included by conv_tc.cu (same translation unit: shares the tensor-map encoder).
//
// Why a second kernel: for the 640x640 stem (7x7, 3->16) and level0 (3x3, 16->16) of DLA-34
// (cubercnn/modeling/backbone/dla.py:287-297) the tap-per-TMA-box implicit GEMM of conv_tc_kernel moves
// taps x 128 px x Cin x 2 B through L2->smem per 128 output pixels (196 KB for the stem) while the math is tiny:
// the layer is L2->SM bandwidth bound at ~20x its HBM roofline.  Here every input row is brought to shared memory
// ONCE per 128-pixel strip, as planes of 8 channels ([plane][pixel][8 ch], 16 B per pixel), and the tensor core
// reads its A operand *in place* through no-swizzle (INTERLEAVE) UMMA descriptors:
//   K-major canonical form  ((8,m),(T,2)) : ((16 B, SBO), (1, LBO))        (cute/atom/mma_traits_sm100.hpp:192-197)
//   rows = consecutive pixels (16 B apart, SBO = 128 B per group of 8), the two 8-wide K chunks of one MMA are
//   either two channel planes (LBO = plane bytes) or — for the 8-channel stem — two adjacent taps (LBO = 16 B).
// A filter tap is therefore just a different descriptor start address (+kw*16 B, other ring slot for kh): no
// im2col copy exists anywhere.  The CTA walks down a strip, so each new output row costs ONE new input row of TMA.
//
// The weight gradient uses the same resident rows as an MN-major operand: M = 16 pixel shifts (= kw) x 8 channels,
// K = pixels, N = Cout from dY staged the same way; one TMEM block per (kh, plane) accumulates over every row the
// CTA visits and is flushed once with fp32 atomics.
#pragma once

namespace c3d {

struct HaloParams {
  int N, H, W, Cin, Cout, KH, KW, pad;
  int P;                      // input channel planes (Cin / 8)
  int BW;                     // pixels per ring-slot row (multiple of 8)
  int R;                      // ring slots
  int ksteps;                 // MMAs (K = 16) per output row
  int strips;                 // ceil(W / 128)
  int rows_per_chunk, chunks_per_col, total_chunks;
  const bf16* w;              // [Cout][KH][KW][Cin] bf16
  const float* bias;
  int relu, out_fp32;
  void* out;
  long long out_pix_stride, out_img_stride, out_h_stride, out_w_stride, out_off;
  float* stats;               // [gridDim.x][2][Cout] or null
  // weight gradient only
  int PO;                     // output channel planes (Cout / 8)
  int RD;                     // dY ring slots
  float* dw;
  int oihw;
};

// ------------------------------------------------------------------------------------------------------------
// forward / data-gradient: y[n,y,x,:] = sum_taps W[:,kh,kw,:] . x[n, y+kh-pad, x+kw-pad, :]
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.  CH = Cout / 16 (1 or 2).
// descriptor halves: lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version 1 (bit 46) | layout none
__device__ __forceinline__ uint64_t halo_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; }
__host__ __device__ constexpr uint32_t halo_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }

// KS / CIN > 0 fix the filter size and input channels at compile time: the MMA-issuing thread is a single lane whose
// instruction latency bounds the kernel (measured: ~6 cycles per dependent instruction), so its per-MMA address
// arithmetic must fold to immediates.  KS = CIN = 0 is the generic (slow) fallback.
template <int CH, int KS, int CIN>
__global__ void __launch_bounds__(192, CH == 1 ? 3 : 2)
conv_halo_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const HaloParams P) {
  constexpr int kCout = CH * 16;
  constexpr int kAcc = 4;                                         // TMEM accumulator ring (rows in flight MMA -> epilogue)
  constexpr uint32_t kTmemCols = 128;
  const int KH = KS > 0 ? KS : P.KH, KW = KH;
  const int Cin = CIN > 0 ? CIN : P.Cin;
  const int BW = KS > 1 ? 136 : P.BW;
  const int NP = Cin >> 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int plane_bytes = BW * 16;
  const int slot_bytes = NP * plane_bytes;
  const int wimg_bytes = P.ksteps * kCout * 32;
  uint8_t* wimg = smem;
  uint8_t* ring = smem + ((wimg_bytes + 127) & ~127);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)P.R * slot_bytes);
  uint64_t* empty_bar = full_bar + P.R;
  uint64_t* tfull_bar = empty_bar + P.R;
  uint64_t* tempty_bar = tfull_bar + kAcc;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + kAcc);
  float* red = reinterpret_cast<float*>(tmem_ptr + 4);            // [4 warps][2][kCout]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KWP = (KW + 1) >> 1;                                  // tap pairs per filter row (Cin == 8 mode)
  const bool pair_mode = (Cin == 8);
  const int PQ = Cin >= 16 ? (Cin >> 4) : 1;

  // weight image: [kstep][Cout/8][2 K-chunks][8 rows][8 elems] = canonical no-swizzle K-major B operand
  for (int idx = threadIdx.x; idx < P.ksteps * kCout * 2; idx += blockDim.x) {
    const int j = idx & 1, co = (idx >> 1) % kCout, s = (idx >> 1) / kCout;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (pair_mode) {
      const int kh = s / KWP, kw = 2 * (s - kh * KWP) + j;
      if (kw < KW) v = *reinterpret_cast<const uint4*>(P.w + ((size_t)(co * KH + kh) * KW + kw) * 8);
    } else {
      const int tap = s / PQ, q = s - tap * PQ;
      v = *reinterpret_cast<const uint4*>(P.w + ((size_t)co * KH * KW + tap) * Cin + (2 * q + j) * 8);
    }
    *reinterpret_cast<uint4*>(wimg + (size_t)s * kCout * 32 + (co >> 3) * 256 + j * 128 + (co & 7) * 16) = v;
  }
  ptx::fence_proxy_async();

  if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&tmap_x);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P.R; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < kAcc; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kTmemCols>(tmem_ptr);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t slot = 0, par = 0;
      for (int c = blockIdx.x; c < P.total_chunks; c += gridDim.x) {
        const int col = c / P.chunks_per_col, cy = c - col * P.chunks_per_col;
        const int n = col / P.strips, xs = col - n * P.strips;
        const int y0 = cy * P.rows_per_chunk;
        const int rows = min(P.rows_per_chunk, P.H - y0);
        const int x0 = xs * 128 - P.pad;
        for (int i = 0; i < rows + KH - 1; ++i) {
          ptx::mbar_wait(&empty_bar[slot], par ^ 1u);
          ptx::mbar_expect_tx(&full_bar[slot], (uint32_t)slot_bytes);
          uint8_t* dst = ring + (size_t)slot * slot_bytes;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            ptx::tma_load_4d(dst + p * plane_bytes, &tmap_x, &full_bar[slot], p * 8, x0, y0 - P.pad + i, n);
          if (++slot == (uint32_t)P.R) { slot = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, kCout, 0, 0);
      const uint32_t R = (uint32_t)P.R;
      const uint32_t slot_u = (uint32_t)slot_bytes >> 4;                         // descriptor address units (16 B)
      const uint32_t a_lo0 = (ptx::smem_u32(ring) >> 4) | ((pair_mode ? 1u : ((uint32_t)plane_bytes >> 4)) << 16);
      const uint32_t b_lo0 = (ptx::smem_u32(wimg) >> 4) | (8u << 16);           // LBO 128 B
      constexpr uint32_t a_hi = halo_desc_hi(128), b_hi = halo_desc_hi(256);
      uint32_t slot0 = 0;                      // ring slot of the first input row of the current output row
      uint32_t rslot = 0, rpar = 0;            // next ring slot to wait for
      int acc = 0; uint32_t acc_phase = 0;
      for (int c = blockIdx.x; c < P.total_chunks; c += gridDim.x) {
        const int cy = c % P.chunks_per_col;
        const int y0 = cy * P.rows_per_chunk;
        const int rows = min(P.rows_per_chunk, P.H - y0);
        for (int j = 0; j < rows; ++j) {
          const int need = j == 0 ? KH : 1;
          for (int t = 0; t < need; ++t) {
            ptx::mbar_wait(&full_bar[rslot], rpar);
            if (++rslot == R) { rslot = 0; rpar ^= 1u; }
          }
          ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
          ptx::tcgen05_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)acc * kCout;
          uint32_t sl = slot0;
          int s = 0;
#pragma unroll
          for (int kh = 0; kh < KH; ++kh) {
            const uint32_t row_lo = a_lo0 + sl * slot_u;
            if (pair_mode) {
#pragma unroll
              for (int pp = 0; pp < KWP; ++pp, ++s)
                ptx::umma_bf16(tacc, halo_desc(row_lo + (uint32_t)(2 * pp), a_hi),
                               halo_desc(b_lo0 + (uint32_t)(s * kCout * 2), b_hi), idesc, s != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int kw = 0; kw < KW; ++kw)
#pragma unroll
                for (int q = 0; q < PQ; ++q, ++s)
                  ptx::umma_bf16(tacc, halo_desc(row_lo + (uint32_t)(2 * q) * ((uint32_t)plane_bytes >> 4) + (uint32_t)kw, a_hi),
                                 halo_desc(b_lo0 + (uint32_t)(s * kCout * 2), b_hi), idesc, s != 0 ? 1u : 0u);
            }
            if (++sl == R) sl = 0;
          }
          ptx::umma_commit(&empty_bar[slot0]);                   // the oldest row of the window is no longer needed
          ptx::umma_commit(&tfull_bar[acc]);
          if (++slot0 == R) slot0 = 0;
          if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
        }
        for (int t = 0; t < KH - 1; ++t) {                       // rows only the finished chunk used
          ptx::umma_commit(&empty_bar[slot0]);
          if (++slot0 == R) slot0 = 0;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    float s1[kCout], s2[kCout];
#pragma unroll
    for (int i = 0; i < kCout; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    for (int c = blockIdx.x; c < P.total_chunks; c += gridDim.x) {
      const int col = c / P.chunks_per_col, cy = c - col * P.chunks_per_col;
      const int n = col / P.strips, xs = col - n * P.strips;
      const int y0 = cy * P.rows_per_chunk;
      const int rows = min(P.rows_per_chunk, P.H - y0);
      const int x = xs * 128 + m;
      const bool valid = x < P.W;
      for (int j = 0; j < rows; ++j) {
        const long long pix = (long long)n * P.out_img_stride + (long long)(y0 + j) * P.out_h_stride +
                              (long long)x * P.out_w_stride + P.out_off;
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kCout;
        uint32_t v[CH][16];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) ptx::tmem_ld_32x32b_x16(tacc + (uint32_t)(ch * 16), v[ch]);
        ptx::tmem_ld_wait();
        // accumulator is in registers: hand the TMEM buffer back before the stores
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[ch][i]);
          if (P.stats) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float xv = valid ? f[i] : 0.f;
              s1[ch * 16 + i] += xv;
              s2[ch * 16 + i] += xv * xv;
            }
          }
          if (valid) {
            const int c0 = ch * 16;
            if (P.bias) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += __ldg(P.bias + c0 + i);
            }
            if (P.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (P.out_fp32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(P.out) + pix * P.out_pix_stride + c0);
#pragma unroll
              for (int i = 0; i < 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            } else {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t*>(&h);
              }
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(P.out) + pix * P.out_pix_stride + c0);
              op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
    }
    if (P.stats) {          // one partial-statistics row per CTA (BatchNorm batch statistics, fp32 partials)
#pragma unroll
      for (int i = 0; i < kCout; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
          s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kCout; ++i) { red[(q * 2 + 0) * kCout + i] = s1[i]; red[(q * 2 + 1) * kCout + i] = s2[i]; }
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      if (m < 2 * kCout) {
        const int which = m / kCout, ci = m - which * kCout;
        float a = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) a += red[(w4 * 2 + which) * kCout + ci];
        P.stats[(size_t)blockIdx.x * 2 * kCout + which * kCout + ci] = a;
      }
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------
// weight gradient: dW[co][kh][kw][ci] += sum_{n,y,x} dY[n,y,x,co] * X[n, y+kh-pad, x+kw-pad, ci]
// One input row X[i] meets the KH gradient rows dY[i-KH+1 .. i] that use it: those rows sit in CONSECUTIVE ring slots
// (the first KH-1 slots are mirrored behind the ring, so a window never wraps), which makes (row, Cout) one long N
// dimension of a single MMA:  D[m = (kw shift, ci)][n = (dY row, co)] += X_row^T . dY_window,  K = 16 pixels.
// TMEM block b of plane p (Cout columns at (p*KH + b)*Cout) holds kh = KH-1-b.  At the first/last rows of a chunk the
// window is clipped to the chunk's own rows (narrower N, shifted block), so every (row, kh) pair is counted once.
template <int KS, int CIN>
__global__ void __launch_bounds__(192, 4)
conv_halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                       const HaloParams P, const uint32_t tmem_cols_pow2) {
  const int KH = KS > 0 ? KS : P.KH;
  const int NP = CIN > 0 ? (CIN >> 3) : P.P;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  constexpr int plane_bytes = 144 * 16;                     // P.BW == 144 always for the weight gradient
  const int slot_bytes = NP * plane_bytes;
  constexpr int dy_plane_bytes = 128 * 16;
  const int dy_slot_bytes = P.PO * dy_plane_bytes;
  uint8_t* ring = smem;
  uint8_t* dyring = ring + (size_t)P.R * slot_bytes;                        // RD slots + KH-1 mirror slots
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(dyring + (size_t)(P.RD + KH - 1) * dy_slot_bytes);
  uint64_t* empty_bar = full_bar + P.R;
  uint64_t* dfull_bar = empty_bar + P.R;
  uint64_t* dempty_bar = dfull_bar + P.RD;
  uint64_t* done_bar = dempty_bar + P.RD;
  uint64_t* zero_bar = done_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(zero_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&tmap_x); ptx::prefetch_tensormap(&tmap_dy); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P.R; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < P.RD; ++s) { ptx::mbar_init(&dfull_bar[s], 1); ptx::mbar_init(&dempty_bar[s], 1); }
    ptx::mbar_init(done_bar, 1);
    ptx::mbar_init(zero_bar, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (tmem_cols_pow2 == 512) ptx::tmem_alloc<512>(tmem_ptr);
    else if (tmem_cols_pow2 == 256) ptx::tmem_alloc<256>(tmem_ptr);
    else ptx::tmem_alloc<128>(tmem_ptr);
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int ncols = NP * KH * P.Cout;                        // multiple of 16

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t slot = 0, par = 0, ds = 0, dpar = 0;
      for (int c = blockIdx.x; c < P.total_chunks; c += gridDim.x) {
        const int col = c / P.chunks_per_col, cy = c - col * P.chunks_per_col;
        const int n = col / P.strips, xs = col - n * P.strips;
        const int y0 = cy * P.rows_per_chunk;
        const int rows = min(P.rows_per_chunk, P.H - y0);
        const int x0 = xs * 128;
        for (int i = 0; i < rows + KH - 1; ++i) {
          if (i < rows) {                                   // dY row i is first needed together with input row i
            ptx::mbar_wait(&dempty_bar[ds], dpar ^ 1u);
            const bool mirror = ds < (uint32_t)(KH - 1);
            ptx::mbar_expect_tx(&dfull_bar[ds], (uint32_t)dy_slot_bytes * (mirror ? 2u : 1u));
            uint8_t* dd = dyring + (size_t)ds * dy_slot_bytes;
            for (int p = 0; p < P.PO; ++p) {
              ptx::tma_load_4d(dd + p * dy_plane_bytes, &tmap_dy, &dfull_bar[ds], p * 8, x0, y0 + i, n);
              if (mirror)
                ptx::tma_load_4d(dd + (size_t)P.RD * dy_slot_bytes + p * dy_plane_bytes, &tmap_dy, &dfull_bar[ds], p * 8, x0,
                                 y0 + i, n);
            }
            if (++ds == (uint32_t)P.RD) { ds = 0; dpar ^= 1u; }
          }
          ptx::mbar_wait(&empty_bar[slot], par ^ 1u);
          ptx::mbar_expect_tx(&full_bar[slot], (uint32_t)slot_bytes);
          uint8_t* dst = ring + (size_t)slot * slot_bytes;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            ptx::tma_load_4d(dst + p * plane_bytes, &tmap_x, &full_bar[slot], p * 8, x0 - P.pad, y0 - P.pad + i, n);
          if (++slot == (uint32_t)P.R) { slot = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc0 = ptx::make_idesc_bf16(128, 0, 1, 1);
      const uint32_t R = (uint32_t)P.R, RD = (uint32_t)P.RD;
      const uint32_t slot_u = (uint32_t)slot_bytes >> 4, dy_slot_u = (uint32_t)dy_slot_bytes >> 4;
      // A: MN-major, MN chunk (pixel shift) stride 16 B ("SBO"), K group (8 px) stride 128 B ("LBO")
      const uint32_t a_lo0 = (ptx::smem_u32(ring) >> 4) | (8u << 16);
      // B: MN-major, MN chunk (8 output channels; planes of a row, then the next row's) stride = dY plane
      const uint32_t b_lo0 = (ptx::smem_u32(dyring) >> 4) | (8u << 16);
      constexpr uint32_t a_hi = halo_desc_hi(16), b_hi = halo_desc_hi(dy_plane_bytes);
      uint32_t xslot = 0, xpar = 0;            // input-row FIFO
      uint32_t dwait = 0, dwpar = 0;           // next dY slot to wait for
      uint32_t dlo = 0;                        // ring slot of the oldest dY row still in use
      ptx::mbar_wait(zero_bar, 0);             // accumulators were zeroed by the epilogue warps
      ptx::tcgen05_fence_after();
      for (int c = blockIdx.x; c < P.total_chunks; c += gridDim.x) {
        const int cy = c % P.chunks_per_col;
        const int y0 = cy * P.rows_per_chunk;
        const int rows = min(P.rows_per_chunk, P.H - y0);
        for (int i = 0; i < rows + KH - 1; ++i) {
          if (i < rows) {
            ptx::mbar_wait(&dfull_bar[dwait], dwpar);
            if (++dwait == RD) { dwait = 0; dwpar ^= 1u; }
          }
          ptx::mbar_wait(&full_bar[xslot], xpar);
          ptx::tcgen05_fence_after();
          const int jlo = max(0, i - KH + 1), jhi = min(rows - 1, i);
          const int nvalid = jhi - jlo + 1;
          const int blo = KH - 1 - i + jlo;                  // TMEM block of dY row jlo (kh = i - jlo)
          const uint32_t idesc = idesc0 | ((uint32_t)(nvalid * P.Cout) >> 3) << 17;
          const uint32_t dy_lo = b_lo0 + dlo * dy_slot_u;    // dlo is the slot of row jlo (rows below jlo are released)
          const uint32_t row_lo = a_lo0 + xslot * slot_u;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const uint32_t tacc = tmem_base + (uint32_t)((p * KH + blo) * P.Cout);
#pragma unroll
            for (int t = 0; t < 8; ++t)
              ptx::umma_bf16(tacc, halo_desc(row_lo + (uint32_t)(p * (plane_bytes >> 4) + t * 16), a_hi),
                             halo_desc(dy_lo + (uint32_t)(t * 16), b_hi), idesc, 1u);
          }
          ptx::umma_commit(&empty_bar[xslot]);
          if (++xslot == R) { xslot = 0; xpar ^= 1u; }
          if (i >= KH - 1) {                                 // dY row i-KH+1 has met its last input row
            ptx::umma_commit(&dempty_bar[dlo]);
            if (++dlo == RD) dlo = 0;
          }
        }
      }
      ptx::umma_commit(done_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int kw = m >> 3, cil = m & 7;
    for (int c0 = 0; c0 < ncols; c0 += 16) ptx::tmem_st_32x32b_x16_fill(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, 0u);
    ptx::tmem_st_wait();
    ptx::tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(zero_bar);
    ptx::mbar_wait(done_bar, 0);
    ptx::tcgen05_fence_after();
    if (q * 4 < P.KW) {                      // warps whose 4 pixel shifts include a real tap
      for (int p = 0; p < NP; ++p)
        for (int b = 0; b < KH; ++b) {
          const int kh = KH - 1 - b;
          const int ci = p * 8 + cil;
          for (int c0 = 0; c0 < P.Cout; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((p * KH + b) * P.Cout + c0), v);
            ptx::tmem_ld_wait();
            if (kw < P.KW) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int co = c0 + i;
                const size_t off = P.oihw ? (((size_t)co * P.Cin + ci) * P.KH + kh) * P.KW + kw
                                          : (((size_t)co * P.KH + kh) * P.KW + kw) * P.Cin + ci;
                atomicAdd(P.dw + off, __uint_as_float(v[i]));
              }
            }
          }
        }
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    if (tmem_cols_pow2 == 512) ptx::tmem_dealloc<512>(tmem_base);
    else if (tmem_cols_pow2 == 256) ptx::tmem_dealloc<256>(tmem_base);
    else ptx::tmem_dealloc<128>(tmem_base);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
static bool halo_enabled() {
  static const bool on = getenv("C3D_CONV_NO_HALO") == nullptr;
  return on;
}
// pure function of the descriptor (c3d_conv2d_tiles must agree with c3d_conv2d_fwd)
static bool halo_fwd_eligible(const c3d_conv_desc* d) {
  if (!halo_enabled()) return false;
  if (d->stride != 1 || d->KH != d->KW || !(d->KH & 1) || 2 * d->pad != d->KH - 1) return false;
  if (d->out_h > 0 || d->out_w > 0 || d->add_mode != 0) return false;
  if (!(d->Cin == 8 || d->Cin == 16 || d->Cin == 32)) return false;
  if (!(d->Cout == 16 || d->Cout == 32)) return false;
  if (d->x_pix_stride != 0 && d->x_pix_stride != d->Cin) return false;
  if ((d->W < 128 && d->Cin != 8) || d->KH > 7) return false;   // NHWC8 input exists only for this kernel
  return true;
}
static bool halo_wgrad_eligible(const c3d_conv_desc* d) {
  if (!halo_enabled()) return false;
  if (d->stride != 1 || d->KH != d->KW || !(d->KH & 1) || 2 * d->pad != d->KH - 1) return false;
  if (!(d->Cin == 8 || d->Cin == 16 || d->Cin == 32)) return false;
  if (!(d->Cout == 16 || d->Cout == 32)) return false;
  if ((d->x_pix_stride != 0 && d->x_pix_stride != d->Cin) || (d->y_pix_stride != 0 && d->y_pix_stride != d->Cout)) return false;
  if ((d->W < 128 && d->Cin != 8) || d->KH > 7) return false;
  if (d->KH * (d->Cin / 8) * d->Cout > 512) return false;
  return true;
}
// ring depth: the TMA rows are only 2-4 KB, so hiding ~1.5 us of L2/HBM latency at ~40 B/ns per SM needs tens of rows in
// flight; C3D_HALO_PREFETCH overrides the number of rows beyond the filter window
static int halo_ring_rows(int window, int slot_bytes, int budget) {
  static const char* env = getenv("C3D_HALO_PREFETCH");
  int pf = budget / slot_bytes - window;
  if (pf > 48) pf = 48;
  if (env) pf = atoi(env);
  if (pf < 2) pf = 2;
  return window + pf;
}
static int halo_fwd_ctas_per_sm(int Cout) { return Cout == 16 ? 3 : 2; }     // = __launch_bounds__ of the instances
static void halo_chunking(const c3d_conv_desc* d, HaloParams* P, int ctas_per_sm, int* grid) {
  P->strips = (d->W + 127) / 128;
  P->rows_per_chunk = d->H < 32 ? d->H : 32;
  P->chunks_per_col = (d->H + P->rows_per_chunk - 1) / P->rows_per_chunk;
  P->total_chunks = d->N * P->strips * P->chunks_per_col;
  const int slots = kNumSMs * ctas_per_sm;
  *grid = P->total_chunks < slots ? P->total_chunks : slots;
}
static int halo_fwd_grid(const c3d_conv_desc* d) {
  HaloParams P; int grid;
  halo_chunking(d, &P, halo_fwd_ctas_per_sm(d->Cout), &grid);
  return grid;
}
static CUresult halo_tensormap(PFN_encodeTiled enc, CUtensorMap* m, const void* base, int C, int W, int H, int N, int boxw) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  cuuint32_t box[4] = {8, (cuuint32_t)boxw, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <int CH, int KS, int CIN>
static int32_t launch_halo_fwd_inst(const CUtensorMap& mx, const HaloParams& P, int grid, size_t smem, cudaStream_t st) {
  auto kern = conv_halo_fwd_kernel<CH, KS, CIN>;
  static size_t cur = 0;
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(C3D_ECUDA, "halo smem attr: %s", cudaGetErrorString(e));
    cur = smem;
  }
  kern<<<grid, 192, smem, st>>>(mx, P);
  return check_launch("conv_halo_fwd_kernel");
}

static int32_t launch_halo_fwd(const c3d_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                               float* stats, cudaStream_t st) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(C3D_ECUDA, "cuTensorMapEncodeTiled unavailable");
  HaloParams P;
  memset(&P, 0, sizeof(P));
  int grid;
  const int ctas = halo_fwd_ctas_per_sm(d->Cout);
  halo_chunking(d, &P, ctas, &grid);
  P.N = d->N; P.H = d->H; P.W = d->W; P.Cin = d->Cin; P.Cout = d->Cout; P.KH = d->KH; P.KW = d->KW; P.pad = d->pad;
  P.P = d->Cin / 8;
  P.BW = (128 + d->KW - 1 + 7) / 8 * 8;
  P.R = halo_ring_rows(d->KH, P.P * P.BW * 16, ctas == 3 ? 48 * 1024 : 64 * 1024);
  P.ksteps = d->Cin == 8 ? d->KH * ((d->KW + 1) / 2) : d->KH * d->KW * (d->Cin / 16);
  P.w = static_cast<const bf16*>(w); P.bias = bias; P.relu = d->relu; P.out_fp32 = d->out_fp32; P.out = y;
  P.out_pix_stride = d->y_pix_stride ? d->y_pix_stride : d->Cout;
  if (d->y_img_stride) {
    P.out_img_stride = d->y_img_stride; P.out_h_stride = d->y_h_stride; P.out_w_stride = d->y_w_stride; P.out_off = d->y_offset;
  } else {
    P.out_img_stride = (long long)d->H * d->W; P.out_h_stride = d->W; P.out_w_stride = 1; P.out_off = 0;
  }
  P.stats = stats;
  CUtensorMap mx;
  CUresult r = halo_tensormap(enc, &mx, x, d->Cin, d->W, d->H, d->N, P.BW);
  if (r != CUDA_SUCCESS) return set_error(C3D_ECUDA, "encode halo x tensormap failed: %d", (int)r);
  const int wimg = (P.ksteps * d->Cout * 32 + 127) & ~127;
  const size_t smem = 128 + (size_t)wimg + (size_t)P.R * P.P * P.BW * 16 + (size_t)(2 * P.R + 8) * 8 + 16 +
                      4 * 2 * d->Cout * sizeof(float) + 64;
  if (smem > 200 * 1024) return set_error(C3D_EINVAL, "halo conv: smem %zu too large", smem);
#define C3D_HALO_F(ch, ks, cin) \
  if (d->Cout == ch * 16 && d->KH == ks && d->Cin == cin) return launch_halo_fwd_inst<ch, ks, cin>(mx, P, grid, smem, st);
  C3D_HALO_F(1, 7, 8)
  C3D_HALO_F(1, 3, 16)
  C3D_HALO_F(1, 3, 32)
  C3D_HALO_F(2, 7, 8)
  C3D_HALO_F(2, 3, 16)
  C3D_HALO_F(2, 3, 32)
#undef C3D_HALO_F
  if (d->Cout == 16) return launch_halo_fwd_inst<1, 0, 0>(mx, P, grid, smem, st);
  return launch_halo_fwd_inst<2, 0, 0>(mx, P, grid, smem, st);
}

template <int KS, int CIN>
static int32_t launch_halo_wgrad_inst(const CUtensorMap& mx, const CUtensorMap& mdy, const HaloParams& P, int grid, size_t smem,
                                      uint32_t tcols, cudaStream_t st) {
  auto kern = conv_halo_wgrad_kernel<KS, CIN>;
  static size_t cur = 0;
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(C3D_ECUDA, "halo wgrad smem attr: %s", cudaGetErrorString(e));
    cur = smem;
  }
  kern<<<grid, 192, smem, st>>>(mx, mdy, P, tcols);
  return check_launch("conv_halo_wgrad_kernel");
}

static int32_t launch_halo_wgrad(const c3d_conv_desc* d, const void* x, const void* dy, float* dw, int oihw,
                                 cudaStream_t st) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(C3D_ECUDA, "cuTensorMapEncodeTiled unavailable");
  HaloParams P;
  memset(&P, 0, sizeof(P));
  const int cols = d->KH * (d->Cin / 8) * d->Cout;
  const uint32_t tcols = cols <= 128 ? 128u : (cols <= 256 ? 256u : 512u);
  P.N = d->N; P.H = d->H; P.W = d->W; P.Cin = d->Cin; P.Cout = d->Cout; P.KH = d->KH; P.KW = d->KW; P.pad = d->pad;
  P.P = d->Cin / 8; P.PO = d->Cout / 8;
  P.BW = 144;
  P.R = halo_ring_rows(0, P.P * P.BW * 16, 16 * 1024);             // input rows are a plain FIFO here
  P.RD = halo_ring_rows(d->KH, P.PO * 128 * 16, 0);                // window of KH rows + 2 in flight
  P.dw = dw; P.oihw = oihw;
  const size_t smem = 128 + (size_t)P.R * P.P * P.BW * 16 + (size_t)(P.RD + d->KH - 1) * P.PO * 128 * 16 +
                      (size_t)(2 * P.R + 2 * P.RD + 2) * 8 + 16 + 64;
  if (smem > 200 * 1024) return set_error(C3D_EINVAL, "halo wgrad: smem %zu too large", smem);
  int ctas = (int)(512u / tcols);                                  // co-resident CTAs: TMEM blocks, shared memory
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  if (ctas > by_smem) ctas = by_smem;
  if (ctas < 1) ctas = 1;
  int grid;
  halo_chunking(d, &P, ctas, &grid);
  CUtensorMap mx, mdy;
  CUresult r = halo_tensormap(enc, &mx, x, d->Cin, d->W, d->H, d->N, P.BW);
  if (r != CUDA_SUCCESS) return set_error(C3D_ECUDA, "encode halo x tensormap failed: %d", (int)r);
  r = halo_tensormap(enc, &mdy, dy, d->Cout, d->W, d->H, d->N, 128);
  if (r != CUDA_SUCCESS) return set_error(C3D_ECUDA, "encode halo dy tensormap failed: %d", (int)r);
  if (d->KH == 7 && d->Cin == 8) return launch_halo_wgrad_inst<7, 8>(mx, mdy, P, grid, smem, tcols, st);
  if (d->KH == 3 && d->Cin == 16) return launch_halo_wgrad_inst<3, 16>(mx, mdy, P, grid, smem, tcols, st);
  if (d->KH == 3 && d->Cin == 32) return launch_halo_wgrad_inst<3, 32>(mx, mdy, P, grid, smem, tcols, st);
  return launch_halo_wgrad_inst<0, 0>(mx, mdy, P, grid, smem, tcols, st);
}

}  // namespace c3d
