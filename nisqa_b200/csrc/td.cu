// td.cu - time-dependency + pooling kernels.
//   adapt arch  : Linear 384->64 + LayerNorm (reference nisqa/NISQA_lib.py:989-991),
//                 2x post-norm encoder layer with 1-head attention (lib:1025-1040),
//                 5 (or 1) PoolAttFF heads (lib:1171-1183, fan-out lib:260-268).
//   standard    : fc_out 768->20 (lib:832-834), BiLSTM (lib:925-943), PoolLastStepBi
//                 (lib:1107-1115).
// All of it is row-local fp32 work on [n_seg, 64] matrices (3.5 % of the FLOPs of a clip):
// one thread owns one time step (row); weights sit in shared memory as [k][out] so the
// inner loop is a warp-broadcast LDS.128 per 4 FMAs.  Clips are ragged: every kernel works
// on the valid rows only, so no key-padding mask exists (masked keys in the reference
// contribute exactly 0 after softmax).
#include "common.cuh"

namespace nisqa {

constexpr int kRows = 128;       // rows (threads) per CTA in the row-thread kernels
constexpr int kXS = 65;          // padded row stride of the per-thread smem row

// acc[j] += sum_{k<kc} xrow[k] * ws[k*NCOL + j]
template <int NCOL>
__device__ __forceinline__ void rowgemm(float (&acc)[NCOL], const float* xrow, const float* ws, int kc) {
  static_assert(NCOL % 4 == 0, "NCOL");
#pragma unroll 4
  for (int k = 0; k < kc; ++k) {
    const float xv = xrow[k];
    const float4* w4 = reinterpret_cast<const float4*>(ws + k * NCOL);
#pragma unroll
    for (int q = 0; q < NCOL / 4; ++q) {
      const float4 w = w4[q];
      acc[q * 4 + 0] = fmaf(xv, w.x, acc[q * 4 + 0]);
      acc[q * 4 + 1] = fmaf(xv, w.y, acc[q * 4 + 1]);
      acc[q * 4 + 2] = fmaf(xv, w.z, acc[q * 4 + 2]);
      acc[q * 4 + 3] = fmaf(xv, w.w, acc[q * 4 + 3]);
    }
  }
}

// nn.LayerNorm(64): biased variance, eps 1e-5
__device__ __forceinline__ void layernorm64(float (&v)[64], const float* __restrict__ gamma,
                                            const float* __restrict__ beta) {
  float mean = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) mean += v[j];
  mean *= (1.0f / 64.0f);
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) { const float d = v[j] - mean; var = fmaf(d, d, var); }
  const float rstd = 1.0f / sqrtf(var * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
  for (int j = 0; j < 64; ++j) v[j] = (v[j] - mean) * rstd * __ldg(gamma + j) + __ldg(beta + j);
}

__device__ __forceinline__ void stage_f4(float* dst, const float* __restrict__ src, int n_floats) {
  const float4* s = reinterpret_cast<const float4*>(src);
  for (int i = threadIdx.x; i < n_floats / 4; i += blockDim.x) reinterpret_cast<float4*>(dst)[i] = __ldg(s + i);
}

// ---------------------------------------------------------------------------------------
// out[row][0..NOUT) = (LN?)( in[row][0..K) @ WT[K][NOUT] + bias )   K % 64 == 0
template <int NOUT, bool LN>
__global__ void __launch_bounds__(kRows)
linear_rows_kernel(const float* __restrict__ in, int K, const float* __restrict__ WT,
                   const float* __restrict__ bias, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ out, int n_rows) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                       // [kRows][kXS]
  float* ws = sm + kRows * kXS;         // [64][NOUT]
  const int row0 = blockIdx.x * kRows, tid = threadIdx.x;
  float acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = __ldg(bias + j);
  for (int k0 = 0; k0 < K; k0 += 64) {
    __syncthreads();
    for (int i = tid; i < kRows * 64; i += kRows) {
      const int r = i >> 6, k = i & 63;
      xs[r * kXS + k] = (row0 + r < n_rows) ? __ldg(in + (size_t)(row0 + r) * K + k0 + k) : 0.f;
    }
    stage_f4(ws, WT + (size_t)k0 * NOUT, 64 * NOUT);
    __syncthreads();
    rowgemm<NOUT>(acc, xs + tid * kXS, ws, 64);
  }
  if (row0 + tid >= n_rows) return;
  if constexpr (LN) {
    static_assert(!LN || NOUT == 64, "LayerNorm width");
    layernorm64(acc, gamma, beta);
  }
  float* o = out + (size_t)(row0 + tid) * NOUT;
#pragma unroll
  for (int q = 0; q < NOUT / 4; ++q)
    reinterpret_cast<float4*>(o)[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
}

// ---------------------------------------------------------------------------------------
// qkv[row] = [ (Wq x + bq)/8 | Wk x + bk | Wv x + bv ]   (1/sqrt(64) folded into Wq,bq on the host)
__global__ void __launch_bounds__(kRows)
qkv_kernel(const float* __restrict__ x, const float* __restrict__ WT3 /*[3][64][64]*/,
           const float* __restrict__ b3 /*[192]*/, float* __restrict__ qkv, int n_rows) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;
  float* ws = sm + kRows * kXS;         // [64][64]
  const int row0 = blockIdx.x * kRows, tid = threadIdx.x;
  for (int i = tid; i < kRows * 64; i += kRows) {
    const int r = i >> 6, k = i & 63;
    xs[r * kXS + k] = (row0 + r < n_rows) ? __ldg(x + (size_t)(row0 + r) * 64 + k) : 0.f;
  }
  for (int part = 0; part < 3; ++part) {
    __syncthreads();
    stage_f4(ws, WT3 + part * 4096, 4096);
    __syncthreads();
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = __ldg(b3 + part * 64 + j);
    rowgemm<64>(acc, xs + tid * kXS, ws, 64);
    if (row0 + tid < n_rows) {
      float4* o = reinterpret_cast<float4*>(qkv + (size_t)(row0 + tid) * 192 + part * 64);
#pragma unroll
      for (int q = 0; q < 16; ++q) o[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// One encoder layer minus the QKV projection, for 128 queries of one clip per CTA:
//   flash-style softmax(q k^T) v over the clip's own keys (online max/sum, keys in blocks of 8)
//   -> out_proj -> +x -> LN1 -> FFN(ReLU) -> + -> LN2          (reference lib:1032-1038)
struct SaLayerParams {
  const float* WoT; const float* bo;         // [64][64] k-major, [64]
  const float* W1T; const float* b1;
  const float* W2T; const float* b2;
  const float* ln1_g; const float* ln1_b;
  const float* ln2_g; const float* ln2_b;
};

constexpr int kKeyTile = 64;
constexpr int kSaSmemFloats = 3 * 4096 + 2 * kKeyTile * 64 + kRows * kXS;

__global__ void __launch_bounds__(kRows, 2)
sa_layer_kernel(const float* __restrict__ x_in, const float* __restrict__ qkv,
                const ClipDesc* __restrict__ clips, int n_clips, const int* __restrict__ qtile_prefix,
                SaLayerParams P, float* __restrict__ x_out) {
  extern __shared__ __align__(16) float sm[];
  float* wo = sm; float* w1 = sm + 4096; float* w2 = sm + 8192;
  float* ks = sm + 12288;                      // [64 keys][64]
  float* vs = ks + kKeyTile * 64;
  float* xs = vs + kKeyTile * 64;              // per-thread rows, stride 65
  const int tid = threadIdx.x;
  const int c = upper_slot(qtile_prefix, n_clips, blockIdx.x);
  const ClipDesc cd = clips[c];
  const int S = cd.n_seg;
  const int q0 = (blockIdx.x - __ldg(qtile_prefix + c)) * kRows;
  const int qi = q0 + tid;
  const bool active = qi < S;
  const size_t rowg = (size_t)cd.seg_off + (active ? qi : 0);

  stage_f4(wo, P.WoT, 4096); stage_f4(w1, P.W1T, 4096); stage_f4(w2, P.W2T, 4096);

  float q[64], o[64];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + rowg * 192);
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float4 t = __ldg(qp + i); q[4*i] = t.x; q[4*i+1] = t.y; q[4*i+2] = t.z; q[4*i+3] = t.w; }
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) o[j] = 0.f;
  float m = -INFINITY, l = 0.f;

  for (int j0 = 0; j0 < S; j0 += kKeyTile) {
    __syncthreads();
    for (int i = tid; i < kKeyTile * 16; i += kRows) {       // float4 granules of k and v rows
      const int kr = i >> 4, g = i & 15;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + kr < S) {
        const float4* src = reinterpret_cast<const float4*>(qkv + ((size_t)cd.seg_off + j0 + kr) * 192);
        kv = __ldg(src + 16 + g); vv = __ldg(src + 32 + g);
      }
      reinterpret_cast<float4*>(ks)[i] = kv; reinterpret_cast<float4*>(vs)[i] = vv;
    }
    __syncthreads();
    const int nk = min(kKeyTile, S - j0);
    for (int jb = 0; jb < nk; jb += 8) {
      float s[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4* kr = reinterpret_cast<const float4*>(ks + (jb + u) * 64);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const float4 kk = kr[g];
          a0 = fmaf(q[4*g], kk.x, a0); a1 = fmaf(q[4*g+1], kk.y, a1);
          a2 = fmaf(q[4*g+2], kk.z, a2); a3 = fmaf(q[4*g+3], kk.w, a3);
        }
        s[u] = (jb + u < nk) ? (a0 + a1) + (a2 + a3) : -INFINITY;
      }
      float bm = s[0];
#pragma unroll
      for (int u = 1; u < 8; ++u) bm = fmaxf(bm, s[u]);
      const float mn = fmaxf(m, bm);
      const float sc = expf(m - mn);          // m == -inf on the first block: expf(-inf) = 0
      l *= sc;
#pragma unroll
      for (int j = 0; j < 64; ++j) o[j] *= sc;
      m = mn;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float p = expf(s[u] - mn);      // masked tail: expf(-inf) = 0
        l += p;
        const float4* vr = reinterpret_cast<const float4*>(vs + (jb + u) * 64);
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const float4 vv = vr[g];
          o[4*g] = fmaf(p, vv.x, o[4*g]); o[4*g+1] = fmaf(p, vv.y, o[4*g+1]);
          o[4*g+2] = fmaf(p, vv.z, o[4*g+2]); o[4*g+3] = fmaf(p, vv.w, o[4*g+3]);
        }
      }
    }
  }
  const float inv = 1.0f / l;
  float* xrow = xs + tid * kXS;
#pragma unroll
  for (int j = 0; j < 64; ++j) xrow[j] = o[j] * inv;

  // out_proj + residual + LN1   (q[] reused as the accumulator, o[] as the residual stream)
#pragma unroll
  for (int j = 0; j < 64; ++j) q[j] = __ldg(P.bo + j);
  rowgemm<64>(q, xrow, wo, 64);
  {
    const float4* xp = reinterpret_cast<const float4*>(x_in + rowg * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float4 t = __ldg(xp + i); o[4*i] = t.x + q[4*i]; o[4*i+1] = t.y + q[4*i+1]; o[4*i+2] = t.z + q[4*i+2]; o[4*i+3] = t.w + q[4*i+3]; }
  }
  layernorm64(o, P.ln1_g, P.ln1_b);
  // FFN
#pragma unroll
  for (int j = 0; j < 64; ++j) { xrow[j] = o[j]; q[j] = __ldg(P.b1 + j); }
  rowgemm<64>(q, xrow, w1, 64);
#pragma unroll
  for (int j = 0; j < 64; ++j) { xrow[j] = fmaxf(q[j], 0.f); q[j] = __ldg(P.b2 + j); }
  rowgemm<64>(q, xrow, w2, 64);
#pragma unroll
  for (int j = 0; j < 64; ++j) o[j] += q[j];
  layernorm64(o, P.ln2_g, P.ln2_b);
  if (active) {
    float4* op = reinterpret_cast<float4*>(x_out + rowg * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) op[i] = make_float4(o[4*i], o[4*i+1], o[4*i+2], o[4*i+3]);
  }
}

// ---------------------------------------------------------------------------------------
// PoolAttFF logits: logit[row][h] = w2_h . relu(W1_h x + b1_h) + b2_h      (lib:1173)
struct PoolHeadParams {      // device pointers, heads concatenated
  const float* W1T;   // [n_heads][64 k][128 j]
  const float* b1;    // [n_heads][128]
  const float* w2;    // [n_heads][128]
  const float* b2;    // [n_heads]
  const float* w3;    // [n_heads][64]
  const float* b3;    // [n_heads]
};

__global__ void __launch_bounds__(kRows)
pool_logits_kernel(const float* __restrict__ x, PoolHeadParams P, int n_heads,
                   float* __restrict__ logits, int n_rows) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;
  float* ws = sm + kRows * kXS;         // [64][128]
  const int row0 = blockIdx.x * kRows, tid = threadIdx.x;
  for (int i = tid; i < kRows * 64; i += kRows) {
    const int r = i >> 6, k = i & 63;
    xs[r * kXS + k] = (row0 + r < n_rows) ? __ldg(x + (size_t)(row0 + r) * 64 + k) : 0.f;
  }
  for (int h = 0; h < n_heads; ++h) {
    __syncthreads();
    stage_f4(ws, P.W1T + (size_t)h * 64 * 128, 64 * 128);
    __syncthreads();
    float logit = __ldg(P.b2 + h);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = __ldg(P.b1 + h * 128 + half * 64 + j);
      // columns [half*64, half*64+64) of the [64][128] tile: row stride 128
      const float* xrow = xs + tid * kXS;
#pragma unroll 4
      for (int k = 0; k < 64; ++k) {
        const float xv = xrow[k];
        const float4* w4 = reinterpret_cast<const float4*>(ws + k * 128 + half * 64);
#pragma unroll
        for (int qd = 0; qd < 16; ++qd) {
          const float4 w = w4[qd];
          acc[qd*4] = fmaf(xv, w.x, acc[qd*4]); acc[qd*4+1] = fmaf(xv, w.y, acc[qd*4+1]);
          acc[qd*4+2] = fmaf(xv, w.z, acc[qd*4+2]); acc[qd*4+3] = fmaf(xv, w.w, acc[qd*4+3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 64; ++j) logit = fmaf(__ldg(P.w2 + h * 128 + half * 64 + j), fmaxf(acc[j], 0.f), logit);
    }
    if (row0 + tid < n_rows) logits[(size_t)(row0 + tid) * n_heads + h] = logit;
  }
}

// softmax over the clip's time steps, weighted sum of x, Linear 64->1   (lib:1177-1181)
// grid = n_clips, block = 64 * n_heads; thread (h, d)
__global__ void pool_final_kernel(const float* __restrict__ x, const float* __restrict__ logits,
                                  const ClipDesc* __restrict__ clips, PoolHeadParams P, int n_heads,
                                  float* __restrict__ scores) {
  __shared__ float red[5 * 64];
  const ClipDesc cd = clips[blockIdx.x];
  const int S = cd.n_seg;
  const int h = threadIdx.x >> 6, d = threadIdx.x & 63;
  if (S <= 0) { if (d == 0) scores[blockIdx.x * n_heads + h] = __int_as_float(0x7fc00000); return; }
  const float* lg = logits + (size_t)cd.seg_off * n_heads + h;
  float mx = -INFINITY;
  for (int t = d; t < S; t += 64) mx = fmaxf(mx, __ldg(lg + (size_t)t * n_heads));
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) { if (d < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
  mx = red[h * 64];
  __syncthreads();
  float sum = 0.f;
  for (int t = d; t < S; t += 64) sum += expf(__ldg(lg + (size_t)t * n_heads) - mx);
  red[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) { if (d < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  sum = red[h * 64];
  __syncthreads();
  const float* xb = x + (size_t)cd.seg_off * 64 + d;
  float acc = 0.f;
  for (int t = 0; t < S; ++t) acc = fmaf(expf(__ldg(lg + (size_t)t * n_heads) - mx), __ldg(xb + (size_t)t * 64), acc);
  acc = (acc / sum) * __ldg(P.w3 + h * 64 + d);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) { if (d < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (d == 0) scores[blockIdx.x * n_heads + h] = red[h * 64] + __ldg(P.b3 + h);
}

// ---------------------------------------------------------------------------------------
// BiLSTM(20 -> 128), one CTA per (clip, direction), 512 threads = 512 gate rows.  The four gates
// of hidden unit j live in one lane quad (thread t: unit t>>2, gate t&3 in PyTorch order i,f,g,o),
// so the gate exchange is four shuffles, the cell update is replicated in the quad, and a step
// needs ONE barrier (h and x are double buffered).  W_hh row: first 64 taps in registers, last 64
// in shared memory [k][512].
struct LstmParams {
  const float* w_ih;   // [2][512][20]
  const float* w_hh;   // [2][512][128]
  const float* b;      // [2][512]   (bias_ih + bias_hh)
  const float* w_pool; // [256]
};
constexpr int kLstmSmemFloats = 64 * 512 + 2 * 128 + 2 * 32 + 128;

__global__ void __launch_bounds__(512, 1)
lstm_kernel(const float* __restrict__ feats /*[n_seg][20]*/, const ClipDesc* __restrict__ clips,
            LstmParams P, float* __restrict__ td_out /*[n_seg][256]*/, float* __restrict__ partial /*[n_clips][2]*/) {
  extern __shared__ __align__(16) float sm[];
  float* whs = sm;                   // [64][512]  taps 64..127, indexed by thread
  float* hbuf = sm + 64 * 512;       // [2][128]
  float* xbuf = hbuf + 256;          // [2][32] input of the current / next step (20 used)
  float* red = xbuf + 64;            // [128]
  const int clip = blockIdx.x >> 1, dir = blockIdx.x & 1;
  const ClipDesc cd = clips[clip];
  const int S = cd.n_seg;
  const int t = threadIdx.x, lane = t & 31;
  const int unit = t >> 2, gate = t & 3;
  const int grow = gate * 128 + unit;          // row of the PyTorch gate matrices
  if (S <= 0) { if (t == 0) partial[clip * 2 + dir] = 0.f; return; }

  float wr[64], wi[20];
  {
    const float* wrow = P.w_hh + ((size_t)dir * 512 + grow) * 128;
#pragma unroll
    for (int k = 0; k < 64; ++k) wr[k] = __ldg(wrow + k);
    for (int k = 0; k < 64; ++k) whs[k * 512 + t] = __ldg(wrow + 64 + k);
    const float* irow = P.w_ih + ((size_t)dir * 512 + grow) * 20;
#pragma unroll
    for (int k = 0; k < 20; ++k) wi[k] = __ldg(irow + k);
  }
  const float bias = __ldg(P.b + dir * 512 + grow);
  if (t < 256) hbuf[t] = 0.f;
  if (t < 64) xbuf[t] = 0.f;
  float cstate = 0.f, hlast = 0.f;
  const float* fb = feats + (size_t)cd.seg_off * 20;
  __syncthreads();
  if (t < 20) xbuf[t] = __ldg(fb + (size_t)(dir ? S - 1 : 0) * 20 + t);
  __syncthreads();

  for (int step = 0; step < S; ++step) {
    const int tt = dir ? S - 1 - step : step;
    const float* h = hbuf + (step & 1) * 128;
    const float* xt = xbuf + (step & 1) * 32;
    // prefetch the next input row while this step computes
    float xnext = 0.f;
    if (t < 20 && step + 1 < S) xnext = __ldg(fb + (size_t)(dir ? S - 2 - step : step + 1) * 20 + t);
    float a0 = bias, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < 20; k += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(xt + k);
      a0 = fmaf(wi[k], xv.x, a0); a1 = fmaf(wi[k+1], xv.y, a1); a2 = fmaf(wi[k+2], xv.z, a2); a3 = fmaf(wi[k+3], xv.w, a3);
    }
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(h + k);
      a0 = fmaf(wr[k], hv.x, a0); a1 = fmaf(wr[k+1], hv.y, a1); a2 = fmaf(wr[k+2], hv.z, a2); a3 = fmaf(wr[k+3], hv.w, a3);
    }
#pragma unroll 8
    for (int k = 0; k < 64; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(h + 64 + k);
      a0 = fmaf(whs[(k) * 512 + t], hv.x, a0); a1 = fmaf(whs[(k+1) * 512 + t], hv.y, a1);
      a2 = fmaf(whs[(k+2) * 512 + t], hv.z, a2); a3 = fmaf(whs[(k+3) * 512 + t], hv.w, a3);
    }
    const float pre = (a0 + a1) + (a2 + a3);
    // gate nonlinearity: gate 2 is the cell candidate (tanh), the others sigmoid
    const float act = (gate == 2) ? tanhf(pre) : 1.0f / (1.0f + expf(-pre));
    const int q0 = lane & ~3;
    const float ig = __shfl_sync(0xffffffffu, act, q0);
    const float fg = __shfl_sync(0xffffffffu, act, q0 + 1);
    const float gg = __shfl_sync(0xffffffffu, act, q0 + 2);
    const float og = __shfl_sync(0xffffffffu, act, q0 + 3);
    cstate = fmaf(fg, cstate, ig * gg);
    hlast = og * tanhf(cstate);
    if (gate == 0) {
      hbuf[((step + 1) & 1) * 128 + unit] = hlast;
      td_out[((size_t)cd.seg_off + tt) * 256 + dir * 128 + unit] = hlast;
    }
    if (t < 20) xbuf[((step + 1) & 1) * 32 + t] = xnext;
    __syncthreads();
  }
  // PoolLastStepBi: this direction's final hidden state . w_pool half
  if (gate == 0) red[unit] = hlast * __ldg(P.w_pool + dir * 128 + unit);
  __syncthreads();
  if (t < 32) {
    float v = red[t] + red[t + 32] + red[t + 64] + red[t + 96];
    v = warp_sum(v);
    if (t == 0) partial[clip * 2 + dir] = v;
  }
}

// ---------------------------------------------------------------------------------------
// Batched BiLSTM (round 2): one CTA advances NB clips of one direction in lock step, so that the recurrent weights
// (512 x 128 fp32 = 256 KB: half in registers, half in shared memory) are fetched once per step for NB sequences
// instead of once per sequence - the step is FMA bound (2 x 148 x NB FMAs per thread) instead of LDS / barrier bound,
// and 256 clips x 2 directions fit on the chip in ONE wave (the one-sequence kernel above needed 3.5 waves of
// 987 serial steps).  256 threads: thread t owns hidden unit u = t >> 1 and the gate pair gp = t & 1 ((i, f) or
// (g, o), PyTorch row order i, f, g, o), i.e. two rows of W_hh / W_ih; the pair of lanes of a unit exchanges its four
// gate values with two shuffles per sequence and both update the (replicated) cell state.  Sequences of a group may
// have different lengths (clips are sorted by length on the host): a finished sequence keeps its state.
// `order` lists the clips of the pass by decreasing n_seg; group g = clips order[NB g .. NB g + NB).
template <int NB>
__global__ void __launch_bounds__(256, 1)
lstm_batched_kernel(const float* __restrict__ feats /*[n_seg][20]*/, const ClipDesc* __restrict__ clips,
                    const int* __restrict__ order, int n_clips, LstmParams P,
                    float* __restrict__ td_out /*[n_seg][256] or nullptr*/, float* __restrict__ partial /*[n_clips][2]*/) {
  extern __shared__ __align__(16) float sm[];
  float2* whs = reinterpret_cast<float2*>(sm);       // [64 k][256 t]: taps 64..127 of this thread's two rows
  float* hbuf = sm + 2 * 64 * 256;                   // [2][NB][128]
  float* xbuf = hbuf + 2 * NB * 128;                 // [2][NB][32]   (20 used)
  float* red = xbuf + 2 * NB * 32;                   // [NB][128]
  const int t = threadIdx.x, u = t >> 1, gp = t & 1;
  const int dir = blockIdx.x & 1, g0 = (blockIdx.x >> 1) * NB;
  int S[NB], clip[NB];
  const float* fb[NB];
  size_t seg_off[NB];
  int maxS = 0;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    clip[b] = (g0 + b < n_clips) ? __ldg(order + g0 + b) : -1;
    S[b] = 0; fb[b] = feats; seg_off[b] = 0;
    if (clip[b] >= 0) {
      const ClipDesc cd = clips[clip[b]];
      S[b] = cd.n_seg; seg_off[b] = (size_t)cd.seg_off; fb[b] = feats + (size_t)cd.seg_off * 20;
    }
    maxS = max(maxS, S[b]);
  }
  const int rowA = (2 * gp) * 128 + u, rowB = rowA + 128;
  float wrA[64], wrB[64], wiA[20], wiB[20];
  {
    const float* ra = P.w_hh + ((size_t)dir * 512 + rowA) * 128;
    const float* rb = P.w_hh + ((size_t)dir * 512 + rowB) * 128;
#pragma unroll
    for (int k = 0; k < 64; ++k) { wrA[k] = __ldg(ra + k); wrB[k] = __ldg(rb + k); }
    for (int k = 0; k < 64; ++k) whs[k * 256 + t] = make_float2(__ldg(ra + 64 + k), __ldg(rb + 64 + k));
    const float* ia = P.w_ih + ((size_t)dir * 512 + rowA) * 20;
    const float* ib = P.w_ih + ((size_t)dir * 512 + rowB) * 20;
#pragma unroll
    for (int k = 0; k < 20; ++k) { wiA[k] = __ldg(ia + k); wiB[k] = __ldg(ib + k); }
  }
  const float biasA = __ldg(P.b + dir * 512 + rowA), biasB = __ldg(P.b + dir * 512 + rowB);
  for (int i = t; i < 2 * NB * 128; i += 256) hbuf[i] = 0.f;
  for (int i = t; i < 2 * NB * 32; i += 256) xbuf[i] = 0.f;
  float cst[NB], hl[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) { cst[b] = 0.f; hl[b] = 0.f; }
  __syncthreads();
  // x of step 0: thread t < NB * 32 loads element (b = t >> 5, k = t & 31)
  const int xb_ = t >> 5, xk_ = t & 31;
  if (t < NB * 32 && xk_ < 20) {
    int Sb = 0; const float* f = feats;
#pragma unroll
    for (int b = 0; b < NB; ++b) if (b == xb_) { Sb = S[b]; f = fb[b]; }
    if (Sb > 0) xbuf[xb_ * 32 + xk_] = __ldg(f + (size_t)(dir ? Sb - 1 : 0) * 20 + xk_);
  }
  __syncthreads();

  for (int step = 0; step < maxS; ++step) {
    const float* h = hbuf + (step & 1) * NB * 128;
    const float* xt = xbuf + (step & 1) * NB * 32;
    float xnext = 0.f;
    bool xload = false;
    if (t < NB * 32 && xk_ < 20) {
      int Sb = 0; const float* f = feats;
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b == xb_) { Sb = S[b]; f = fb[b]; }
      if (step + 1 < Sb) { xnext = __ldg(f + (size_t)(dir ? Sb - 2 - step : step + 1) * 20 + xk_); xload = true; }
    }
    // every row keeps two partial sums (even / odd taps), whatever NB is: a clip's result must not depend on how many
    // clips share its CTA (alone == in a batch, bit for bit), and one 148-long dependent FMA chain per row would
    // leave the NB = 1 variant latency bound
    constexpr int PART = 2;
    float pA[NB][PART], pB[NB][PART];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int q = 0; q < PART; ++q) { pA[b][q] = q ? 0.f : biasA; pB[b][q] = q ? 0.f : biasB; }
#pragma unroll
    for (int k = 0; k < 20; k += 4) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 xv = *reinterpret_cast<const float4*>(xt + b * 32 + k);
        pA[b][0] = fmaf(wiA[k], xv.x, pA[b][0]); pA[b][1 % PART] = fmaf(wiA[k + 1], xv.y, pA[b][1 % PART]);
        pA[b][2 % PART] = fmaf(wiA[k + 2], xv.z, pA[b][2 % PART]); pA[b][3 % PART] = fmaf(wiA[k + 3], xv.w, pA[b][3 % PART]);
        pB[b][0] = fmaf(wiB[k], xv.x, pB[b][0]); pB[b][1 % PART] = fmaf(wiB[k + 1], xv.y, pB[b][1 % PART]);
        pB[b][2 % PART] = fmaf(wiB[k + 2], xv.z, pB[b][2 % PART]); pB[b][3 % PART] = fmaf(wiB[k + 3], xv.w, pB[b][3 % PART]);
      }
    }
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(h + b * 128 + k);
        pA[b][0] = fmaf(wrA[k], hv.x, pA[b][0]); pA[b][1 % PART] = fmaf(wrA[k + 1], hv.y, pA[b][1 % PART]);
        pA[b][2 % PART] = fmaf(wrA[k + 2], hv.z, pA[b][2 % PART]); pA[b][3 % PART] = fmaf(wrA[k + 3], hv.w, pA[b][3 % PART]);
        pB[b][0] = fmaf(wrB[k], hv.x, pB[b][0]); pB[b][1 % PART] = fmaf(wrB[k + 1], hv.y, pB[b][1 % PART]);
        pB[b][2 % PART] = fmaf(wrB[k + 2], hv.z, pB[b][2 % PART]); pB[b][3 % PART] = fmaf(wrB[k + 3], hv.w, pB[b][3 % PART]);
      }
    }
#pragma unroll 4
    for (int k = 0; k < 64; k += 4) {
      const float2 w0 = whs[(k) * 256 + t], w1 = whs[(k + 1) * 256 + t], w2 = whs[(k + 2) * 256 + t], w3 = whs[(k + 3) * 256 + t];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(h + b * 128 + 64 + k);
        pA[b][0] = fmaf(w0.x, hv.x, pA[b][0]); pA[b][1 % PART] = fmaf(w1.x, hv.y, pA[b][1 % PART]);
        pA[b][2 % PART] = fmaf(w2.x, hv.z, pA[b][2 % PART]); pA[b][3 % PART] = fmaf(w3.x, hv.w, pA[b][3 % PART]);
        pB[b][0] = fmaf(w0.y, hv.x, pB[b][0]); pB[b][1 % PART] = fmaf(w1.y, hv.y, pB[b][1 % PART]);
        pB[b][2 % PART] = fmaf(w2.y, hv.z, pB[b][2 % PART]); pB[b][3 % PART] = fmaf(w3.y, hv.w, pB[b][3 % PART]);
      }
    }
    float aA[NB], aB[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (PART == 4) { aA[b] = (pA[b][0] + pA[b][1 % PART]) + (pA[b][2 % PART] + pA[b][3 % PART]); aB[b] = (pB[b][0] + pB[b][1 % PART]) + (pB[b][2 % PART] + pB[b][3 % PART]); }
      else if (PART == 2) { aA[b] = pA[b][0] + pA[b][1 % PART]; aB[b] = pB[b][0] + pB[b][1 % PART]; }
      else { aA[b] = pA[b][0]; aB[b] = pB[b][0]; }
    }
    float* hn = hbuf + ((step + 1) & 1) * NB * 128;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      // gp == 0: (aA, aB) = pre-activations of (i, f); gp == 1: of (g, o)
      const float actA = gp ? tanhf(aA[b]) : 1.0f / (1.0f + expf(-aA[b]));
      const float actB = 1.0f / (1.0f + expf(-aB[b]));
      const float pA = __shfl_xor_sync(0xffffffffu, actA, 1), pB = __shfl_xor_sync(0xffffffffu, actB, 1);
      const float ig = gp ? pA : actA, fg = gp ? pB : actB, gg = gp ? actA : pA, og = gp ? actB : pB;
      if (step < S[b]) {                       // warp-uniform: S[b] is the same for every thread
        cst[b] = fmaf(fg, cst[b], ig * gg);
        hl[b] = og * tanhf(cst[b]);
        if (gp == 0) {
          hn[b * 128 + u] = hl[b];
          if (td_out) td_out[(seg_off[b] + (size_t)(dir ? S[b] - 1 - step : step)) * 256 + dir * 128 + u] = hl[b];
        }
      } else if (gp == 0) {
        hn[b * 128 + u] = hl[b];               // finished sequence: state carried along unchanged
      }
    }
    if (xload) xbuf[((step + 1) & 1) * NB * 32 + xb_ * 32 + xk_] = xnext;
    __syncthreads();
  }
  // PoolLastStepBi: this direction's final hidden state . w_pool half (lib:1107-1115)
  if (gp == 0) {
    const float w = __ldg(P.w_pool + dir * 128 + u);
#pragma unroll
    for (int b = 0; b < NB; ++b) red[b * 128 + u] = hl[b] * w;
  }
  __syncthreads();
  if (t < 32 * NB) {
    const int b = t >> 5, lane = t & 31;
    float v = red[b * 128 + lane] + red[b * 128 + lane + 32] + red[b * 128 + lane + 64] + red[b * 128 + lane + 96];
    v = warp_sum(v);
    int cb = -1;
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) if (bb == b) cb = clip[bb];
    if (lane == 0 && cb >= 0) partial[cb * 2 + dir] = v;      // clips without segments: 0 (lastbi_final writes NaN for them)
  }
}

__global__ void lastbi_final_kernel(const float* __restrict__ partial, const ClipDesc* __restrict__ clips,
                                    float bias, float* __restrict__ scores, int n_clips) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clips) return;
  scores[c] = (clips[c].n_seg > 0) ? (partial[2 * c] + partial[2 * c + 1]) + bias
                                   : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------ host launchers
constexpr int kRowSmem64 = (kRows * kXS + 64 * 64) * 4;
constexpr int kRowSmem20 = (kRows * kXS + 64 * 20) * 4;
constexpr int kRowSmem128 = (kRows * kXS + 64 * 128) * 4;

void launch_lin_ln(cudaStream_t st, const float* feats, const float* WT, const float* b,
                   const float* g, const float* be, float* out, int n_rows) {
  static unsigned long long cfg = 0;
  if (first_launch_on_device(cfg)) { cudaFuncSetAttribute(linear_rows_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem64); }
  linear_rows_kernel<64, true><<<(n_rows + kRows - 1) / kRows, kRows, kRowSmem64, st>>>(feats, 384, WT, b, g, be, out, n_rows);
}
void launch_fc20(cudaStream_t st, const float* feats, const float* WT, const float* b, float* out, int n_rows) {
  linear_rows_kernel<20, false><<<(n_rows + kRows - 1) / kRows, kRows, kRowSmem20, st>>>(feats, 768, WT, b, nullptr, nullptr, out, n_rows);
}
void launch_qkv(cudaStream_t st, const float* x, const float* WT3, const float* b3, float* qkv, int n_rows) {
  static unsigned long long cfg = 0;
  if (first_launch_on_device(cfg)) { cudaFuncSetAttribute(qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem64); }
  qkv_kernel<<<(n_rows + kRows - 1) / kRows, kRows, kRowSmem64, st>>>(x, WT3, b3, qkv, n_rows);
}
void launch_sa_layer(cudaStream_t st, const float* x_in, const float* qkv, const ClipDesc* clips,
                     int n_clips, const int* qtile_prefix, int n_qtiles, const SaLayerParams& P,
                     float* x_out) {
  static unsigned long long cfg = 0;
  const int smem = kSaSmemFloats * 4;
  if (first_launch_on_device(cfg)) { cudaFuncSetAttribute(sa_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
  sa_layer_kernel<<<n_qtiles, kRows, smem, st>>>(x_in, qkv, clips, n_clips, qtile_prefix, P, x_out);
}
void launch_pool_att(cudaStream_t st, const float* x, const ClipDesc* clips, int n_clips, int n_rows,
                     const PoolHeadParams& P, int n_heads, float* logits, float* scores) {
  static unsigned long long cfg = 0;
  if (first_launch_on_device(cfg)) { cudaFuncSetAttribute(pool_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem128); }
  pool_logits_kernel<<<(n_rows + kRows - 1) / kRows, kRows, kRowSmem128, st>>>(x, P, n_heads, logits, n_rows);
  pool_final_kernel<<<n_clips, 64 * n_heads, 0, st>>>(x, logits, clips, P, n_heads, scores);
}
void launch_pool_final(cudaStream_t st, const float* x, const float* logits, const ClipDesc* clips, int n_clips,
                       const PoolHeadParams& P, int n_heads, float* scores) {
  pool_final_kernel<<<n_clips, 64 * n_heads, 0, st>>>(x, logits, clips, P, n_heads, scores);
}
void launch_lstm(cudaStream_t st, const float* feats20, const ClipDesc* clips, int n_clips,
                 const LstmParams& P, float* td_out, float* partial, float pool_bias, float* scores) {
  static unsigned long long cfg = 0;
  const int smem = kLstmSmemFloats * 4;
  if (first_launch_on_device(cfg)) { cudaFuncSetAttribute(lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
  lstm_kernel<<<2 * n_clips, 512, smem, st>>>(feats20, clips, P, td_out, partial);
  lastbi_final_kernel<<<(n_clips + 127) / 128, 128, 0, st>>>(partial, clips, pool_bias, scores, n_clips);
}

template <int NB> constexpr int lstm_batched_smem() { return (2 * 64 * 256 + 2 * NB * 128 + 2 * NB * 32 + NB * 128) * 4; }

// `order`: device array of the pass's clip indices sorted by decreasing n_seg (host-built, run_pass); the batch
// width follows the number of sequences per SM: 148 SMs x NB sequences per direction pair of CTAs in one wave
void launch_lstm_batched(cudaStream_t st, const float* feats20, const ClipDesc* clips, const int* order, int n_clips,
                         const LstmParams& P, float* td_out, float* partial, float pool_bias, float* scores) {
  static unsigned long long cfg = 0;
  if (first_launch_on_device(cfg)) {
    cudaFuncSetAttribute(lstm_batched_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lstm_batched_smem<1>());
    cudaFuncSetAttribute(lstm_batched_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lstm_batched_smem<2>());
    cudaFuncSetAttribute(lstm_batched_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lstm_batched_smem<4>());
  }
  const int seqs = 2 * n_clips;
  if (seqs <= 148)
    lstm_batched_kernel<1><<<2 * n_clips, 256, lstm_batched_smem<1>(), st>>>(feats20, clips, order, n_clips, P, td_out, partial);
  else if (seqs <= 2 * 148)
    lstm_batched_kernel<2><<<2 * ((n_clips + 1) / 2), 256, lstm_batched_smem<2>(), st>>>(feats20, clips, order, n_clips, P, td_out, partial);
  else
    lstm_batched_kernel<4><<<2 * ((n_clips + 3) / 4), 256, lstm_batched_smem<4>(), st>>>(feats20, clips, order, n_clips, P, td_out, partial);
  lastbi_final_kernel<<<(n_clips + 127) / 128, 128, 0, st>>>(partial, clips, pool_bias, scores, n_clips);
}

}  // namespace nisqa
