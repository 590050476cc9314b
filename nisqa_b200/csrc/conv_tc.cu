// conv_tc.cu - conv3 / conv4 of the AdaptCNN / StandardCNN (reference nisqa/NISQA_lib.py:696-700,
// 820-825) as an implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), with an
// error-compensated two-term FP16 split so that the result stays within fp32 rounding noise of
// the reference (plain TF32 / BF16 operands move MOS by 2e-3 / 1.7e-2):
//      a = a_hi + a_lo,  b = b_hi + b_lo      (fp16 parts: 11 + 11 significant bits)
//      a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi          (dropped a_lo*b_lo ~ 2^-22 |a b|)
// Issued as TWO MMAs per K-step: A_hi x [B_hi | B_lo] with N = 2*C_out (accumulator columns
// [0,C) = hi*hi, [C,2C) = hi*lo) and A_lo x B_hi with N = C_out into columns [0,C); the epilogue
// adds the two column halves.  Every A element is thus fetched from shared memory twice per tap
// instead of three times - operand fetch, not the MMA issue rate, bounds these kernels.
// Products are exact in the tensor core and accumulate in fp32 (TMEM).  The weights are
// pre-scaled by 2^S on the host (S chosen per layer so that max|w| 2^S <= 1024) to keep b_lo
// out of the fp16 subnormal range; the epilogue multiplies by 2^-S (exact).  Compared with
// 3xTF32 this halves shared memory, weight traffic and MMA time (kind::f16 has K = 16).
//
// GEMM view:  D[pos, co] = sum_{tap, ci} X[pos + off(tap), ci] * W[tap][co][ci]
//   M = flattened padded positions of G segments (row pitch W+1: one shared zero column per
//       row, one shared zero row between segments), N = 64 output channels, K = 9 * CIN.
//   A (activations) is staged ONCE per CTA in shared memory in the canonical no-swizzle K-major
//   UMMA layout [ci/8][row][8 halves] (core matrix = 8 rows x 16 B contiguous, SBO = 128 B,
//   LBO = plane).  The 9 taps are the SAME tile addressed with a row-shifted start address, so
//   im2col is never materialised.  B (weights, pre-split on the host) streams tap by tap with
//   cp.async.bulk + mbarrier into a 2-stage ring.  Accumulators: 2 M-tiles x 64 fp32 columns
//   of TMEM.  One elected thread issues the MMAs; 4 warps run the epilogue (tcgen05.ld ->
//   bias + ReLU (+ adaptive / 2x2 max-pool through a shared-memory staging tile) -> channels-last
//   global store).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nisqa {

#ifdef NISQA_TC_TIMING
__device__ long long g_tc_timing[8 * 8192];
#endif
__device__ __forceinline__ void tc_stamp(int slot, int who) {
#ifdef NISQA_TC_TIMING
  if ((int)threadIdx.x == who && blockIdx.x < 8192) g_tc_timing[blockIdx.x * 8 + slot] = clock64();
#endif
}

// ------------------------------------------------------------------ configuration
enum { TC_POOL_NONE = 0, TC_POOL_ADAPT = 1, TC_POOL_2X2 = 2 };

template <int H_, int W_, int CIN_, int COUT_, int POOL_, int POW_, int NSTAGE_, bool CENTER_ = false>
struct TcCfg {
  static constexpr int H = H_, W = W_, CIN = CIN_, COUT = COUT_, POOL = POOL_, POW = POW_;
  static constexpr bool CENTER = CENTER_;         // conv6 of the AdaptCNN: kernel (3,3), padding (1,0) on a
                                                  // 3-wide map == the padded conv evaluated at column 1 only
  static constexpr int P = W + 1;                 // row pitch: W interior columns + 1 shared zero column
  static constexpr int BLK = (H + 1) * P;         // rows per segment: H interior rows + 1 shared zero row
  static constexpr int G = 256 / BLK;             // segments per CTA (2 M-tiles of 128 rows)
  static constexpr int HALO = P + 1;              // |row offset| of the farthest tap
  static constexpr int AROWS = ((256 + 2 * HALO) | 1);    // odd: conflict-free staging stores
  static constexpr int NCH = CIN / 8;             // 16-byte K chunks (8 halves)
  static constexpr int A_BYTES = NCH * AROWS * 16;        // per hi / lo
  static constexpr int B_HALF = NCH * COUT * 16;          // per hi / lo
  static constexpr int B_STAGE = 2 * B_HALF;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int TMEM_COLS = (4 * COUT <= 64) ? 64 : (4 * COUT <= 128 ? 128 : 256);   // 2 M-tiles x 2*COUT
  static constexpr int HO = (POOL == TC_POOL_NONE) ? H : H / 2;
  static constexpr int STG_STRIDE = COUT + 4;     // floats per staged row (conflict-free float4)
  static constexpr int OFF_A_HI = 0;
  static constexpr int OFF_A_LO = A_BYTES;
  static constexpr int OFF_B = 2 * A_BYTES;
  static constexpr int OFF_BAR = OFF_B + NSTAGE * B_STAGE;
  static constexpr int SMEM_BYTES = OFF_BAR + 16 * NSTAGE + 32;
  static constexpr int MINB_SMEM = (SMEM_BYTES <= 56 * 1024) ? 4 : (SMEM_BYTES <= 74 * 1024) ? 3 : (SMEM_BYTES <= 112 * 1024 ? 2 : 1);
  static constexpr int MINB = (MINB_SMEM * TMEM_COLS <= 512) ? MINB_SMEM : 512 / TMEM_COLS;
  // D=f32, A=B=f16, both K-major, M=128; N = 2*COUT ([b_hi|b_lo]) and N = COUT (b_hi only)
  static constexpr uint32_t IDESC_2N = (1u << 4) | ((uint32_t)((2 * COUT) >> 3) << 17) | ((128u >> 4) << 24);
  static constexpr uint32_t IDESC_1N = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((128u >> 4) << 24);
  static_assert(POOL == TC_POOL_NONE || G * H * W * STG_STRIDE * 4 <= 2 * A_BYTES + NSTAGE * B_STAGE,
                "pool staging tile must fit in the A+B region");
  static_assert(A_BYTES % 16 == 0 && B_STAGE % 16 == 0 && CIN % 16 == 0 && COUT % 32 == 0, "shape");
  static_assert(G >= 1 && MINB * TMEM_COLS <= 512, "tile / TMEM budget");
};

template <class C>
__global__ void __launch_bounds__(192, C::MINB)
conv_tc_kernel(const float* __restrict__ in /*[seg][H][W][CIN] fp32*/,
               const __half* __restrict__ wtc /*[9][CIN/8][hi co | lo co][8] fp16, scaled by 2^S*/,
               const float* __restrict__ bias, float out_scale /*2^-S*/,
               float* __restrict__ out, int n_seg) {
  constexpr int H = C::H, W = C::W, CIN = C::CIN, COUT = C::COUT, P = C::P, BLK = C::BLK, G = C::G;
  constexpr int HALO = C::HALO, AROWS = C::AROWS, NCH = C::NCH, NS = C::NSTAGE;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a_hi = sbase + C::OFF_A_HI, a_lo = sbase + C::OFF_A_LO, b_base = sbase + C::OFF_B;
  const uint32_t bar_full = sbase + C::OFF_BAR;          // [NS]
  const uint32_t bar_empty = bar_full + 8 * NS;          // [NS]
  const uint32_t bar_acc = bar_full + 16 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 16 * NS + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seg0 = blockIdx.x * G;
  tc_stamp(0, 0);

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
  if (tid == 32) {
#pragma unroll
    for (int i = 0; i < NS; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  tc_stamp(1, 0);

  // ---- weight producer: the first NS taps are in flight while the activation tile is staged
  if (tid == 160) {
#pragma unroll
    for (int t = 0; t < NS; ++t) {
      mbar_expect_tx(bar_full + 8 * t, C::B_STAGE);
      bulk_g2s(b_base + t * C::B_STAGE, wtc + (size_t)t * (C::B_STAGE / 2), C::B_STAGE, bar_full + 8 * t);
    }
  }

  // ---- stage A: channels-last fp32 global -> [ci/8][row][8 halves] hi / lo, zero halo.
  //      4 items per thread per round so that 8 independent 16-byte loads are in flight.
  {
    uint4* ah = reinterpret_cast<uint4*>(smem + C::OFF_A_HI);
    uint4* al = reinterpret_cast<uint4*>(smem + C::OFF_A_LO);
    constexpr int ITEMS = AROWS * NCH;
    for (int it0 = tid; it0 < ITEMS; it0 += 192 * 4) {
      float4 va[4], vb[4];
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * 192;
        va[u] = make_float4(0.f, 0.f, 0.f, 0.f); vb[u] = va[u]; dst[u] = -1;
        if (it < ITEMS) {
          const int c8 = it % NCH, b = it / NCH;
          dst[u] = c8 * AROWS + b;
          const int r = b - HALO;
          if (r >= 0 && r < G * BLK) {
            const int s = r / BLK, q = r - s * BLK;
            const int hh = q / P, ww = q - hh * P;
            if (hh >= 1 && ww >= 1 && seg0 + s < n_seg) {
              const float4* src = reinterpret_cast<const float4*>(
                  in + ((size_t)(seg0 + s) * (H * W) + (hh - 1) * W + (ww - 1)) * CIN + c8 * 8);
              va[u] = __ldg(src); vb[u] = __ldg(src + 1);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (dst[u] >= 0) {
          uint4 hi, lo;
          split8(va[u], vb[u], hi, lo);
          ah[dst[u]] = hi; al[dst[u]] = lo;
        }
      }
    }
  }
  fence_proxy_async();          // generic-proxy stores -> visible to the tensor-core (async) proxy
  __syncthreads();
  tc_stamp(2, 0);

  if (warp == 5) {
    // ===== weight producer (one lane) =====
    if (lane == 0) {
      for (int t = NS; t < 9; ++t) {
        const int s = t % NS;
        mbar_wait(bar_empty + 8 * s, ((t / NS) - 1) & 1);     // MMAs of tap t-NS have drained the stage
        mbar_expect_tx(bar_full + 8 * s, C::B_STAGE);
        bulk_g2s(b_base + s * C::B_STAGE, wtc + (size_t)t * (C::B_STAGE / 2), C::B_STAGE, bar_full + 8 * s);
      }
    }
  } else if (warp == 4) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      tc_fence_after();
      for (int t = 0; t < 9; ++t) {
        const int s = t % NS;
        mbar_wait(bar_full + 8 * s, (t / NS) & 1);
        if (t == 0) tc_stamp(6, 128);
        tc_fence_after();
        const int tapoff = (t / 3 - 1) * P + (t % 3 - 1);
        const uint32_t bst = b_base + s * C::B_STAGE;       // [ci/8][2*COUT rows: hi then lo][8 halves]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t row = (uint32_t)(HALO + mt * 128 + tapoff);
          const uint32_t d = tmem + mt * (2 * COUT);
#pragma unroll
          for (int ks = 0; ks < CIN / 16; ++ks) {
            const uint32_t aoff = ((uint32_t)(2 * ks) * AROWS + row) * 16;
            const uint64_t dah = make_desc(a_hi + aoff, AROWS * 16, 128);
            const uint64_t dal = make_desc(a_lo + aoff, AROWS * 16, 128);
            const uint64_t db = make_desc(bst + (uint32_t)(2 * ks) * (2 * COUT * 16), 2 * COUT * 16, 128);
            umma_f16(d, dah, db, C::IDESC_2N, (t | ks) != 0);     // [0,C) += hi*hi ; [C,2C) += hi*lo
            umma_f16(d + COUT, dal, db, C::IDESC_1N, 1);          // [C,2C) += lo*hi (the small accumulator, see conv_split.cu)
          }
        }
        umma_commit(bar_empty + 8 * s);          // stage s may be refilled once these MMAs retire
      }
      umma_commit(bar_acc);                      // all accumulators final
      tc_stamp(7, 128);
    }
  } else {
    // ===== epilogue: warps 0..3 <-> TMEM lanes 32w..32w+31 =====
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    tc_stamp(3, 0);
    float* stg = reinterpret_cast<float*>(smem);          // reuses the A/B region (all MMAs retired)
    constexpr int WOUT = C::CENTER ? 1 : W;
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int r = mt * 128 + warp * 32 + lane;
      const int s = r / BLK, q = r - s * BLK;
      const int hh = q / P, ww = q - hh * P;
      bool valid = (s < G) && hh >= 1 && ww >= 1 && (seg0 + s < n_seg);
      if (C::CENTER) valid = valid && (ww == 2);
      const int h = hh - 1, w = C::CENTER ? 0 : ww - 1;
#pragma unroll 1
      for (int part = 0; part < COUT / 32; ++part) {
        float v[32], v2[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + mt * (2 * COUT) + part * 32, v);
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + mt * (2 * COUT) + COUT + part * 32, v2);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += v2[j];
        if (valid) {
          float* dst = (C::POOL == TC_POOL_NONE)
                           ? out + ((size_t)(seg0 + s) * (H * WOUT) + h * WOUT + w) * COUT + part * 32
                           : stg + ((s * H + h) * W + w) * C::STG_STRIDE + part * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = fmaxf(fmaf(v[4 * j + 0], out_scale, __ldg(bias + part * 32 + 4 * j + 0)), 0.f);
            o.y = fmaxf(fmaf(v[4 * j + 1], out_scale, __ldg(bias + part * 32 + 4 * j + 1)), 0.f);
            o.z = fmaxf(fmaf(v[4 * j + 2], out_scale, __ldg(bias + part * 32 + 4 * j + 2)), 0.f);
            o.w = fmaxf(fmaf(v[4 * j + 3], out_scale, __ldg(bias + part * 32 + 4 * j + 3)), 0.f);
            reinterpret_cast<float4*>(dst)[j] = o;
          }
        }
      }
    }
    if constexpr (C::POOL != TC_POOL_NONE) {
      asm volatile("bar.sync 1, 128;" ::: "memory");      // epilogue warps only
      constexpr int POW = C::POW, HO = H / 2, C4 = COUT / 4;
      for (int it = tid; it < G * HO * POW * C4; it += 128) {
        const int c4 = it % C4;
        int rest = it / C4;
        const int pw = rest % POW; rest /= POW;
        const int ph = rest % HO;
        const int s = rest / HO;
        if (seg0 + s >= n_seg) continue;
        int x0, x1;
        if (C::POOL == TC_POOL_ADAPT) { x0 = (pw * W) / POW; x1 = ((pw + 1) * W + POW - 1) / POW; }
        else { x0 = 2 * pw; x1 = 2 * pw + 2; }
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);       // post-ReLU values are >= 0
        for (int hy = 2 * ph; hy < 2 * ph + 2; ++hy)
          for (int x = x0; x < x1; ++x) {
            const float4 t = *reinterpret_cast<const float4*>(stg + ((s * H + hy) * W + x) * C::STG_STRIDE + c4 * 4);
            m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y); m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
          }
        *reinterpret_cast<float4*>(out + ((size_t)(seg0 + s) * (HO * POW) + ph * POW + pw) * COUT + c4 * 4) = m;
      }
    }
  }
  tc_stamp(4, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
  tc_stamp(5, 0);
}

#ifndef NISQA_TC_NS4
#define NISQA_TC_NS4 2
#endif
#ifndef NISQA_TC_NS5
#define NISQA_TC_NS5 2
#endif
#ifndef NISQA_TC_NS3
#define NISQA_TC_NS3 4
#endif
// layers 2..6; std_mode selects the StandardCNN geometry (W 8/4/2, MaxPool2d(2))
//                      H   W  CIN COUT POOL           POW NSTAGE CENTER
using TcConv2A = TcCfg<24, 7, 16, 32, TC_POOL_ADAPT, 5, 9>;
using TcConv3A = TcCfg<12, 5, 32, 64, TC_POOL_NONE, 0, NISQA_TC_NS3>;
using TcConv4A = TcCfg<12, 5, 64, 64, TC_POOL_ADAPT, 3, NISQA_TC_NS4>;
using TcConv5A = TcCfg<6, 3, 64, 64, TC_POOL_NONE, 0, NISQA_TC_NS5>;
using TcConv6A = TcCfg<6, 3, 64, 64, TC_POOL_NONE, 0, NISQA_TC_NS5, true>;
using TcConv2S = TcCfg<24, 8, 16, 32, TC_POOL_2X2, 4, 9>;
using TcConv3S = TcCfg<12, 4, 32, 64, TC_POOL_NONE, 0, 4>;
using TcConv4S = TcCfg<12, 4, 64, 64, TC_POOL_2X2, 2, 2>;
using TcConv5S = TcCfg<6, 2, 64, 64, TC_POOL_NONE, 0, 2>;

template <class C>
static void launch_tc(cudaStream_t st, const float* in, const __half* wtc, const float* b, float scale,
                      float* out, int n_seg) {
  static unsigned long long configured = 0;
  if (first_launch_on_device(configured)) {
    cudaFuncSetAttribute(conv_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  }
  conv_tc_kernel<C><<<(n_seg + C::G - 1) / C::G, 192, C::SMEM_BYTES, st>>>(in, wtc, b, scale, out, n_seg);
}

void launch_conv_tc(cudaStream_t st, int std_mode, int layer, const float* in, const void* wtc,
                    const float* b, float out_scale, float* out, int n_seg) {
  const __half* w = reinterpret_cast<const __half*>(wtc);
  if (!std_mode) {
    switch (layer) {
      case 2: launch_tc<TcConv2A>(st, in, w, b, out_scale, out, n_seg); break;
      case 3: launch_tc<TcConv3A>(st, in, w, b, out_scale, out, n_seg); break;
      case 4: launch_tc<TcConv4A>(st, in, w, b, out_scale, out, n_seg); break;
      case 5: launch_tc<TcConv5A>(st, in, w, b, out_scale, out, n_seg); break;
      default: launch_tc<TcConv6A>(st, in, w, b, out_scale, out, n_seg); break;
    }
  } else {
    switch (layer) {
      case 2: launch_tc<TcConv2S>(st, in, w, b, out_scale, out, n_seg); break;
      case 3: launch_tc<TcConv3S>(st, in, w, b, out_scale, out, n_seg); break;
      case 4: launch_tc<TcConv4S>(st, in, w, b, out_scale, out, n_seg); break;
      default: launch_tc<TcConv5S>(st, in, w, b, out_scale, out, n_seg); break;   // conv5 and conv6 share a geometry
    }
  }
}

#ifdef NISQA_TC_TIMING
int tc_timing_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_tc_timing, sizeof(long long) * n);
}
#endif

}  // namespace nisqa
