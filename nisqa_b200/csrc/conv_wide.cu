// conv_wide.cu - EXPERIMENTAL (round-2 candidate, not on the default path, not yet run on a GPU):
// conv3..conv6 (C_out = 64) on planes like conv_split.cu, but with the GEMM transposed so that ONE MMA covers
// all 256 positions of the CTA's tile:
//
//     D[m, n] = sum_{tap, ci} Wst[m, ci] * X[n + off(tap), ci]      m = 0..127, n = 0..255
//       rows m = 0..63   : W_hi of output channel m        rows m = 64..127 : W_lo of output channel m - 64
//       A operand (M = 128) = the tap's weight block  [ci/8][hi co | lo co][8 halves]  (conv_split's B block)
//       B operand (N = 256) = the activation tile, row-shifted per tap, X_hi then X_lo (two MMAs per K-step)
//     out[n, c] = relu((D[c, n] + D[64 + c, n]) * 2^-S + bias[c])      (= all four hi/lo product terms)
//
// Why: round 1 measured ~107 cycles per tcgen05.mma for the N = 128 / 64 instructions of conv_split.cu
// whatever the operand layout or accumulator order - conv4 issues 144 of them per CTA
// (15.4 k cycles).  Here conv4 issues 72 MMAs of N = 256 (math floor 128 cycles each = 9.2 k cycles).
// The price is an epilogue that transposes: TMEM lanes are weight rows, columns are positions, and the hi and
// lo weight rows of a channel sit in different lane quarters.  Warps of lane quarters 0/1 store their rows into
// a shared-memory tile S[position][channel], warps of quarters 2/3 add theirs (each element is touched by one
// thread per phase), and a second pass over S on all threads applies scale / bias / ReLU, max-pools, splits and
// stores 16-byte chunks (coalesced; the zero rows / columns of the planes are never written).
//
// Selected with nisqa_set_option("conv_wide", layer mask); tests/test_gpu_parity.py::test_conv_wide_candidate
// only runs when NISQA_EXPERIMENTAL=1.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nisqa {

namespace {

enum { WD_POOL_NONE = 0, WD_POOL_ADAPT = 1, WD_POOL_2X2 = 2 };

template <int ROWB>
__device__ __forceinline__ size_t wd_plane_off(int g, int c) {        // == split_off of conv_split.cu
  const size_t o = (size_t)g * ROWB + (size_t)c * 16;
  return o ^ ((o >> 3) & (size_t)(ROWB - 16));
}

template <int H_, int W_, int CIN_, int POOL_, int POW_, int NSTAGE_, bool F32OUT_ = false, bool CENTER_ = false>
struct WdCfg {
  static constexpr int H = H_, W = W_, CIN = CIN_, COUT = 64, POOL = POOL_, POW = POW_;
  static constexpr bool CENTER = CENTER_, OUT_SPLIT = !F32OUT_;
  static constexpr int EPW = 8, NT = (EPW + 2) * 32;
  static constexpr int P = W + 1, BLK = (H + 1) * P, G = 256 / BLK, HALO = P + 1;
  static constexpr int AROWS = 256 + 2 * HALO, ROWB = CIN * 2;
  static constexpr uint32_t LAYOUT = (ROWB == 128) ? 2u : (ROWB == 64 ? 4u : 6u);
  static constexpr int X_BYTES = ((AROWS + 7) * ROWB + 1023) & ~1023;      // activation tile, per hi / lo
  static constexpr int NCH = CIN / 8;
  static constexpr int W_STAGE = 2 * NCH * COUT * 16;                       // one tap: [ci/8][128 rows][16 B]
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int TMEM_COLS = 256;
  static constexpr int HO = (POOL == WD_POOL_NONE) ? H : H / 2;
  static constexpr int WO = (POOL == WD_POOL_NONE) ? (CENTER ? 1 : W) : POW;
  static constexpr int OP = WO + 1, OBLK = (HO + 1) * OP, OROWB = COUT * 2;
  static constexpr int S_STRIDE = COUT + 4;                                 // floats per position row of S
  static constexpr int OFF_X_HI = 0, OFF_X_LO = X_BYTES, OFF_W = 2 * X_BYTES;
  static constexpr int OFF_BAR = OFF_W + NSTAGE * W_STAGE;
  static constexpr int SMEM_BYTES = OFF_BAR + 16 * NSTAGE + 32 + 1024;
  static constexpr int MINB = (SMEM_BYTES <= 112 * 1024) ? 2 : 1;           // TMEM: 2 x 256 columns
  static constexpr uint32_t IDESC = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);   // f32 acc, f16 x f16, M 128, N 256
  static_assert(G * BLK * S_STRIDE * 4 <= OFF_BAR, "transpose tile must fit in the operand region");
  static_assert(ROWB == 64 || ROWB == 128, "C_in 32 / 64");
  static_assert(HALO <= kSplitLead && G >= 1, "geometry");
  static_assert(!CENTER || F32OUT_, "centre-column variant is the last layer");
  static_assert(OUT_SPLIT || POOL == WD_POOL_NONE, "fp32 output is not pooled");
};

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
conv_wide_kernel(const unsigned char* __restrict__ in_hi, const unsigned char* __restrict__ in_lo,
                 const __half* __restrict__ wtc /*[9][CIN/8][hi co | lo co][8] fp16, scaled by 2^S*/,
                 const float* __restrict__ bias, float out_scale /*2^-S*/,
                 unsigned char* __restrict__ out_hi, unsigned char* __restrict__ out_lo,
                 float* __restrict__ out_f32, int n_seg) {
  constexpr int H = C::H, W = C::W, CIN = C::CIN, COUT = C::COUT, P = C::P, BLK = C::BLK, G = C::G;
  constexpr int HALO = C::HALO, NS = C::NSTAGE, ROWB = C::ROWB, EPW = C::EPW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t x_hi = sbase + C::OFF_X_HI, x_lo = sbase + C::OFF_X_LO, w_base = sbase + C::OFF_W;
  const uint32_t bar_full = sbase + C::OFF_BAR, bar_empty = bar_full + 8 * NS;
  const uint32_t bar_acc = bar_full + 16 * NS, bar_x = bar_acc + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 16 * NS + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seg0 = blockIdx.x * G;
  const int g0 = kSplitLead + seg0 * BLK - HALO;
  const uint32_t sh = (uint32_t)(g0 & 7);

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
  if (tid == 32) {
#pragma unroll
    for (int i = 0; i < NS; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_x, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == EPW + 1) {
    // ===== producer (one lane): activation tile (2 bulk copies), then the weights tap by tap =====
    if (lane == 0) {
      constexpr uint32_t X_COPY = (uint32_t)C::AROWS * ROWB;
      mbar_expect_tx(bar_x, 2 * X_COPY);
      bulk_g2s(x_hi + sh * ROWB, in_hi + (size_t)g0 * ROWB, X_COPY, bar_x);
      bulk_g2s(x_lo + sh * ROWB, in_lo + (size_t)g0 * ROWB, X_COPY, bar_x);
      for (int t = 0; t < 9; ++t) {
        const int s = t % NS;
        if (t >= NS) mbar_wait(bar_empty + 8 * s, ((t / NS) - 1) & 1);
        mbar_expect_tx(bar_full + 8 * s, C::W_STAGE);
        bulk_g2s(w_base + s * C::W_STAGE, wtc + (size_t)t * (C::W_STAGE / 2), C::W_STAGE, bar_full + 8 * s);
      }
    }
  } else if (warp == EPW) {
    // ===== MMA issuer (one lane): 2 MMAs of M = 128, N = 256 per K-step =====
    if (lane == 0) {
      mbar_wait(bar_x, 0);
      for (int t = 0; t < 9; ++t) {
        const int s = t % NS;
        mbar_wait(bar_full + 8 * s, (t / NS) & 1);
        tc_fence_after();
        const int tapoff = (t / 3 - 1) * P + (t % 3 - 1);
        const uint32_t row = sh + (uint32_t)(HALO + tapoff);          // first of the 256 tile rows of this tap
        const uint32_t wst = w_base + s * C::W_STAGE;
#pragma unroll
        for (int ks = 0; ks < CIN / 16; ++ks) {
          const uint64_t dw = make_desc(wst + (uint32_t)(2 * ks) * (2 * COUT * 16), 2 * COUT * 16, 128);   // A: 128 weight rows
          const uint32_t xoff = row * ROWB + (uint32_t)ks * 32;
          umma_f16(tmem, dw, make_desc_swz(x_hi + xoff, 8 * ROWB, C::LAYOUT), C::IDESC, (t | ks) != 0);  // += W * x_hi
          umma_f16(tmem, dw, make_desc_swz(x_lo + xoff, 8 * ROWB, C::LAYOUT), C::IDESC, 1);              // += W * x_lo
        }
        umma_commit(bar_empty + 8 * s);
      }
      umma_commit(bar_acc);
    }
  } else {
    // ===== epilogue part 1 (8 warps): transpose TMEM [weight row][position] -> S[position][channel] =====
    // warp w: lane quarter q = w & 3 (rows 32q..32q+31), positions [128 (w >> 2), +128)
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    float* S = reinterpret_cast<float*>(smem);              // reuses the operand region (all MMAs retired)
    const int q = warp & 3, n_base = (warp >> 2) * 128;
    const int ch = (q & 1) * 32 + lane;                      // output channel of this thread's TMEM row
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + n_base;
    constexpr int NLIVE = G * BLK;                           // tile rows that map to plane rows
    if (q < 2) {                                             // W_hi rows: plain stores
#pragma unroll 1
      for (int n0 = 0; n0 < 128; n0 += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(trow + n0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n_base + n0 + j < NLIVE) S[(n_base + n0 + j) * C::S_STRIDE + ch] = __uint_as_float(r[j]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");           // the 8 epilogue warps
    if (q >= 2) {                                            // W_lo rows: add to what the hi warps stored
#pragma unroll 1
      for (int n0 = 0; n0 < 128; n0 += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(trow + n0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n_base + n0 + j < NLIVE) S[(n_base + n0 + j) * C::S_STRIDE + ch] += __uint_as_float(r[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();                     // S complete, accumulators read, every role done
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);

  // ===== epilogue part 2 (all threads): scale / bias / ReLU, (max-pool), (split), coalesced stores =====
  const float* S = reinterpret_cast<const float*>(smem);
  constexpr int C8 = COUT / 8;
  auto affine8 = [&](float4& a, float4& b, int c8) {         // relu(x * 2^-S + bias); monotone, so it commutes with max
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8) + 1);
    a.x = fmaxf(fmaf(a.x, out_scale, b0.x), 0.f); a.y = fmaxf(fmaf(a.y, out_scale, b0.y), 0.f);
    a.z = fmaxf(fmaf(a.z, out_scale, b0.z), 0.f); a.w = fmaxf(fmaf(a.w, out_scale, b0.w), 0.f);
    b.x = fmaxf(fmaf(b.x, out_scale, b1.x), 0.f); b.y = fmaxf(fmaf(b.y, out_scale, b1.y), 0.f);
    b.z = fmaxf(fmaf(b.z, out_scale, b1.z), 0.f); b.w = fmaxf(fmaf(b.w, out_scale, b1.w), 0.f);
  };
  if constexpr (C::POOL != WD_POOL_NONE) {
    constexpr int POW = C::POW, HO = H / 2;
    for (int it = tid; it < G * HO * POW * C8; it += C::NT) {
      const int c8 = it % C8;
      int rest = it / C8;
      const int pw = rest % POW; rest /= POW;
      const int ph = rest % HO;
      const int s = rest / HO;
      if (seg0 + s >= n_seg) continue;
      int x0, x1;
      if (C::POOL == WD_POOL_ADAPT) { x0 = (pw * W) / POW; x1 = ((pw + 1) * W + POW - 1) / POW; }
      else { x0 = 2 * pw; x1 = 2 * pw + 2; }
      float4 ma = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), mb = ma;     // raw accumulators: any sign
      for (int hy = 2 * ph; hy < 2 * ph + 2; ++hy)
        for (int x = x0; x < x1; ++x) {
          const float4* t = reinterpret_cast<const float4*>(S + (s * BLK + (hy + 1) * P + (x + 1)) * C::S_STRIDE + c8 * 8);
          const float4 ta = t[0], tb = t[1];
          ma.x = fmaxf(ma.x, ta.x); ma.y = fmaxf(ma.y, ta.y); ma.z = fmaxf(ma.z, ta.z); ma.w = fmaxf(ma.w, ta.w);
          mb.x = fmaxf(mb.x, tb.x); mb.y = fmaxf(mb.y, tb.y); mb.z = fmaxf(mb.z, tb.z); mb.w = fmaxf(mb.w, tb.w);
        }
      affine8(ma, mb, c8);
      uint4 hi, lo;
      split8(ma, mb, hi, lo);
      const int g = kSplitLead + (seg0 + s) * C::OBLK + (ph + 1) * C::OP + (pw + 1);
      const size_t o = wd_plane_off<C::OROWB>(g, c8);
      *reinterpret_cast<uint4*>(out_hi + o) = hi;
      *reinterpret_cast<uint4*>(out_lo + o) = lo;
    }
  } else {
    constexpr int WOUT = C::CENTER ? 1 : W;
    for (int it = tid; it < G * H * WOUT * C8; it += C::NT) {
      const int c8 = it % C8;
      int rest = it / C8;
      const int w = rest % WOUT; rest /= WOUT;
      const int h = rest % H;
      const int s = rest / H;
      if (seg0 + s >= n_seg) continue;
      const int ww = C::CENTER ? 2 : w + 1;
      const float4* t = reinterpret_cast<const float4*>(S + (s * BLK + (h + 1) * P + ww) * C::S_STRIDE + c8 * 8);
      float4 a = t[0], b = t[1];
      affine8(a, b, c8);
      if constexpr (C::OUT_SPLIT) {
        uint4 hi, lo;
        split8(a, b, hi, lo);
        const int g = kSplitLead + (seg0 + s) * C::OBLK + (h + 1) * C::OP + (w + 1);
        const size_t o = wd_plane_off<C::OROWB>(g, c8);
        *reinterpret_cast<uint4*>(out_hi + o) = hi;
        *reinterpret_cast<uint4*>(out_lo + o) = lo;
      } else {
        float4* dst = reinterpret_cast<float4*>(out_f32 + (((size_t)(seg0 + s) * H + h) * WOUT + w) * COUT + c8 * 8);
        dst[0] = a; dst[1] = b;
      }
    }
  }
}

//                     H   W  CIN POOL           POW NSTAGE F32OUT CENTER
using WdConv3A = WdCfg<12, 5, 32, WD_POOL_NONE, 0, 4>;
using WdConv4A = WdCfg<12, 5, 64, WD_POOL_ADAPT, 3, 2>;
using WdConv5A = WdCfg<6, 3, 64, WD_POOL_NONE, 0, 2>;
using WdConv6A = WdCfg<6, 3, 64, WD_POOL_NONE, 0, 2, true, true>;
using WdConv3S = WdCfg<12, 4, 32, WD_POOL_NONE, 0, 4>;
using WdConv4S = WdCfg<12, 4, 64, WD_POOL_2X2, 2, 2>;
using WdConv5S = WdCfg<6, 2, 64, WD_POOL_NONE, 0, 2>;
using WdConv6S = WdCfg<6, 2, 64, WD_POOL_NONE, 0, 2, true>;

template <class C>
void launch_wd(cudaStream_t st, const unsigned char* ih, const unsigned char* il, const __half* w, const float* b,
               float scale, unsigned char* oh, unsigned char* ol, float* of, int n_seg) {
  static unsigned long long configured = 0;
  if (first_launch_on_device(configured)) {
    cudaFuncSetAttribute(conv_wide_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  }
  conv_wide_kernel<C><<<(n_seg + C::G - 1) / C::G, C::NT, C::SMEM_BYTES, st>>>(ih, il, w, b, scale, oh, ol, of, n_seg);
}

}  // namespace

// layer 3..6 (C_out = 64); same plane contract as launch_conv_split
void launch_conv_wide(cudaStream_t st, int std_mode, int layer, const void* in_hi, const void* in_lo,
                      const void* wtc, const float* b, float out_scale, void* out_hi, void* out_lo,
                      float* out_f32, int n_seg) {
  const __half* w = reinterpret_cast<const __half*>(wtc);
  const unsigned char* ih = static_cast<const unsigned char*>(in_hi);
  const unsigned char* il = static_cast<const unsigned char*>(in_lo);
  unsigned char* oh = static_cast<unsigned char*>(out_hi);
  unsigned char* ol = static_cast<unsigned char*>(out_lo);
  if (!std_mode) {
    switch (layer) {
      case 3: launch_wd<WdConv3A>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
      case 4: launch_wd<WdConv4A>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
      case 5: launch_wd<WdConv5A>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
      default: launch_wd<WdConv6A>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
    }
  } else {
    switch (layer) {
      case 3: launch_wd<WdConv3S>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
      case 4: launch_wd<WdConv4S>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
      case 5: launch_wd<WdConv5S>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
      default: launch_wd<WdConv6S>(st, ih, il, w, b, out_scale, oh, ol, out_f32, n_seg); break;
    }
  }
}

}  // namespace nisqa
