// td_tiled.cu - time-dependency block + attention-pool logits of the adapt architecture as register-tiled fp32
// GEMMs (round 2; replaces the one-thread-per-row kernels of td.cu, which ran at 6 % occupancy):
//
//   td_in_kernel   : Linear 384->64 + LayerNorm (reference nisqa/NISQA_lib.py:989-991) and, fused behind it, the
//                    QKV projection of encoder layer 0 (in_proj of nn.MultiheadAttention, lib:1032)
//   td_sa_kernel   : one encoder layer for 64 queries of one clip (lib:1025-1040): softmax(q k^T) v over the clip's
//                    own keys (flash-style, keys in blocks of 64, online max / sum), out_proj, +x, LN1, FFN(ReLU),
//                    +, LN2 - and, fused behind it, either the NEXT layer's QKV projection or (last layer) the
//                    PoolAttFF logits of all heads  w2_h . relu(W1_h x + b1_h) + b2_h  (lib:1173)
//
// Every matrix product is a 64 x 64 x 64 tile product on the FFMA pipe: 256 threads = 16 (ty) x 16 (tx), a thread
// owns rows 4ty..4ty+3 and four columns, operands sit in shared memory, the A operand always row-major
// [row][k] with a padded leading dimension of 68 floats and read as float4 along k:
//   gemm_nn : B k-major [k][n]  (weights as packed by the engine, V), columns 4tx..4tx+3, LDS.128 along n
//   gemm_nt : B row-major [n][k] (K of q k^T), columns tx, tx+16, tx+32, tx+48 - consecutive lanes read consecutive
//             rows of pitch 68 floats = 4 banks apart: conflict-free LDS.128 along k
// A row's 64 columns live in the 16 lanes of one half-warp, so softmax / LayerNorm reductions are four xor-shuffles,
// and P (the softmax numerators) is written and re-read by the same half-warp: no block barrier inside a key block.
// Weights stream through two shared-memory buffers with cp.async (16 KB chunks, L2 resident), the next chunk in
// flight while the current one is multiplied.  fp32 throughout (parity: +-1e-4 on the scores).
#include "common.cuh"
#include "f32x2.cuh"

#ifndef NISQA_TD_PK
#define NISQA_TD_PK 0        // tile GEMMs with packed FFMA2 (same FMAs in the same order)
#endif

namespace nisqa {

struct SaLayerParams {
  const float* WoT; const float* bo; const float* W1T; const float* b1; const float* W2T;
  const float* b2; const float* ln1_g; const float* ln1_b; const float* ln2_g; const float* ln2_b;
};
struct PoolHeadParams { const float* W1T; const float* b1; const float* w2; const float* b2; const float* w3; const float* b3; };

namespace {

constexpr int kT = 64;            // tile edge (rows, columns, k chunk)
constexpr int kLd = 68;           // leading dimension of row-major tiles (floats): 16-byte aligned rows, 4 banks apart
constexpr int kTileF = kT * kLd;  // floats of a padded tile
constexpr int kNT = 256;

__device__ __forceinline__ void cp16(float* dst_smem, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int bytes = valid ? 16 : 0;                       // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 64 rows x 64 floats, global (row pitch src_ld floats) -> shared (row pitch dst_ld floats); rows >= rows_valid are
// zero-filled.  All 256 threads, 4 x 16 bytes each.
__device__ __forceinline__ void tile_load_async(float* dst, int dst_ld, const float* src, long long src_ld, int rows_valid, int tid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = tid + it * kNT;
    const int r = idx >> 4, c4 = idx & 15;
    const bool ok = r < rows_valid;
    cp16(dst + r * dst_ld + c4 * 4, src + (ok ? (long long)r * src_ld + c4 * 4 : 0), ok);
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

#if NISQA_TD_PK
// acc[i][j] += sum_k A[4ty+i][k] * B[k][4tx+j]   (packed FFMA2: the same fp32 FMA per element, in the same k order)
__device__ __forceinline__ void gemm_nn(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ B,
                                        int ldb, int ty, int tx) {
  const float* a0 = A + (4 * ty) * kLd;
  const float* b0 = B + 4 * tx;
  f2 c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = pk(acc[i][0], acc[i][1]); c[i][1] = pk(acc[i][2], acc[i][3]); }
#pragma unroll 2
  for (int k = 0; k < kT; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * kLd + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4*>(b0 + (k + kk) * ldb);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        c[i][0] = fma2(bc(av[kk]), pk(b[kk].x, b[kk].y), c[i][0]);
        c[i][1] = fma2(bc(av[kk]), pk(b[kk].z, b[kk].w), c[i][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 lo = upk(c[i][0]), hi = upk(c[i][1]);
    acc[i][0] = lo.x; acc[i][1] = lo.y; acc[i][2] = hi.x; acc[i][3] = hi.y;
  }
}

// acc[i][j] += sum_k A[4ty+i][k] * B[tx+16j][k]   (columns j, j+1 packed: FFMA2 with the A element broadcast)
__device__ __forceinline__ void gemm_nt(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ B,
                                        int ty, int tx) {
  const float* a0 = A + (4 * ty) * kLd;
  const float* b0 = B + tx * kLd;
  f2 c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = pk(acc[i][0], acc[i][1]); c[i][1] = pk(acc[i][2], acc[i][3]); }
#pragma unroll 2
  for (int k = 0; k < kT; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * kLd + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(b0 + (16 * j) * kLd + k);
    const f2 b01[4] = {pk(b[0].x, b[1].x), pk(b[0].y, b[1].y), pk(b[0].z, b[1].z), pk(b[0].w, b[1].w)};
    const f2 b23[4] = {pk(b[2].x, b[3].x), pk(b[2].y, b[3].y), pk(b[2].z, b[3].z), pk(b[2].w, b[3].w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        c[i][0] = fma2(bc(av[kk]), b01[kk], c[i][0]);
        c[i][1] = fma2(bc(av[kk]), b23[kk], c[i][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 lo = upk(c[i][0]), hi = upk(c[i][1]);
    acc[i][0] = lo.x; acc[i][1] = lo.y; acc[i][2] = hi.x; acc[i][3] = hi.y;
  }
}
#else
// acc[i][j] += sum_k A[4ty+i][k] * B[k][4tx+j]
__device__ __forceinline__ void gemm_nn(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ B,
                                        int ldb, int ty, int tx) {
  const float* a0 = A + (4 * ty) * kLd;
  const float* b0 = B + 4 * tx;
#pragma unroll 2
  for (int k = 0; k < kT; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * kLd + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4*>(b0 + (k + kk) * ldb);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[i][0] = fmaf(av[kk], b[kk].x, acc[i][0]);
        acc[i][1] = fmaf(av[kk], b[kk].y, acc[i][1]);
        acc[i][2] = fmaf(av[kk], b[kk].z, acc[i][2]);
        acc[i][3] = fmaf(av[kk], b[kk].w, acc[i][3]);
      }
    }
  }
}

// acc[i][j] += sum_k A[4ty+i][k] * B[tx+16j][k]
__device__ __forceinline__ void gemm_nt(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ B,
                                        int ty, int tx) {
  const float* a0 = A + (4 * ty) * kLd;
  const float* b0 = B + tx * kLd;
#pragma unroll 2
  for (int k = 0; k < kT; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * kLd + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(b0 + (16 * j) * kLd + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float s = acc[i][j];
        s = fmaf(a[i].x, b[j].x, s); s = fmaf(a[i].y, b[j].y, s);
        s = fmaf(a[i].z, b[j].z, s); s = fmaf(a[i].w, b[j].w, s);
        acc[i][j] = s;
      }
  }
}
#endif

// reductions over the 16 lanes (tx) that share a row
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float row_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// nn.LayerNorm(64) (biased variance, eps 1e-5) of the rows held as v[i][0..3] = columns 4tx..4tx+3
__device__ __forceinline__ void layernorm_rows(float (&v)[4][4], const float* __restrict__ gamma,
                                               const float* __restrict__ beta, int tx) {
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + tx), be = __ldg(reinterpret_cast<const float4*>(beta) + tx);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float mean = row_sum((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) * (1.0f / 64.0f);
    const float d0 = v[i][0] - mean, d1 = v[i][1] - mean, d2 = v[i][2] - mean, d3 = v[i][3] - mean;
    const float var = row_sum(fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3)) * (1.0f / 64.0f);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    v[i][0] = d0 * rstd * g.x + be.x; v[i][1] = d1 * rstd * g.y + be.y;
    v[i][2] = d2 * rstd * g.z + be.z; v[i][3] = d3 * rstd * g.w + be.w;
  }
}

__device__ __forceinline__ void store_rows_smem(float* X, const float (&v)[4][4], int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(X + (4 * ty + i) * kLd + 4 * tx) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
}

// QKV projection of a 64-row tile X (shared, row-major) -> qkv[row0 + r][192]; Wbuf: two [64][64] buffers, the
// first chunk (part 0) must already be in flight into Wbuf[b0] as the most recent cp.async group.
__device__ __forceinline__ void qkv_tail(const float* X, float* Wbuf, const float* __restrict__ WT3,
                                         const float* __restrict__ b3, float* __restrict__ qkv, long long row0,
                                         int rows_valid, int tid, int ty, int tx, int b0 = 0) {
#pragma unroll 1
  for (int part_ = 0; part_ < 3; ++part_) {
    const int part = part_;
    if (part + 1 < 3) tile_load_async(Wbuf + ((part + 1 + b0) & 1) * 4096, 64, WT3 + (part + 1) * 4096, 64, 64, tid);
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    float acc[4][4];
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b3 + part * 64) + tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
    gemm_nn(acc, X, Wbuf + ((part + b0) & 1) * 4096, 64, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * ty + i < rows_valid)
        *reinterpret_cast<float4*>(qkv + (row0 + 4 * ty + i) * 192 + part * 64 + 4 * tx) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();                        // buffer (part & 1) is refilled by the next iteration's prefetch
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
constexpr int kInSmemFloats = 2 * kTileF + 2 * 4096 + kTileF;

__global__ void __launch_bounds__(kNT, 2)
td_in_kernel(const float* __restrict__ feats /*[n][64 nk]*/, const float* __restrict__ WT /*[64 nk][64]*/, int nk,
             const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
             const float* __restrict__ qkvT /*[3][64][64]*/, const float* __restrict__ qkvb /*[192]*/,
             const float* __restrict__ pe /*[max_len][64] positional encoding or nullptr*/,
             const int* __restrict__ seg_clip, const ClipDesc* __restrict__ clips,
             float* __restrict__ x0 /*[n][64]*/, float* __restrict__ qkv /*[n][192]*/, int n_rows) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                       // [2][64][68]
  float* Ws = sm + 2 * kTileF;          // [2][64][64]
  float* Xs = Ws + 2 * 4096;            // [64][68]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long row0 = (long long)blockIdx.x * kT;
  const int rows_valid = (int)min((long long)kT, (long long)n_rows - row0);
  const int ld = 64 * nk;                 // 384 CNN features, or the 192 / 128 fused features of the double-ended model
  const float* src = feats + row0 * ld;

  tile_load_async(As, kLd, src, ld, rows_valid, tid);
  tile_load_async(Ws, 64, WT, 64, 64, tid);
  cp_commit();
  float acc[4][4];
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
  }
#pragma unroll 1
  for (int c = 0; c < nk; ++c) {
    if (c + 1 < nk) {
      tile_load_async(As + ((c + 1) & 1) * kTileF, kLd, src + (c + 1) * 64, ld, rows_valid, tid);
      tile_load_async(Ws + ((c + 1) & 1) * 4096, 64, WT + (size_t)(c + 1) * 4096, 64, 64, tid);
    } else {
      tile_load_async(Ws + ((c + 1) & 1) * 4096, 64, qkvT, 64, 64, tid);       // first QKV chunk rides behind the last k chunk
    }
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    gemm_nn(acc, As + (c & 1) * kTileF, Ws + (c & 1) * 4096, 64, ty, tx);
    __syncthreads();
  }
  layernorm_rows(acc, gamma, beta, tx);
  if (pe != nullptr) {
    // PositionalEncoding (lib:1042-1062): x[t] += pe[t], t = position of the segment inside its clip
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long row = row0 + min(4 * ty + i, rows_valid - 1);
      const int t = (int)(row - clips[__ldg(seg_clip + row)].seg_off);
      const float4 pv = __ldg(reinterpret_cast<const float4*>(pe + (size_t)t * 64) + tx);
      acc[i][0] += pv.x; acc[i][1] += pv.y; acc[i][2] += pv.z; acc[i][3] += pv.w;
    }
  }
  store_rows_smem(Xs, acc, ty, tx);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * ty + i < rows_valid)
      *reinterpret_cast<float4*>(x0 + (row0 + 4 * ty + i) * 64 + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  // QKV of layer 0: its first weight chunk is in flight into Ws[nk & 1]; qkv_tail syncs before reading Xs
  qkv_tail(Xs, Ws, qkvT, qkvb, qkv, row0, rows_valid, tid, ty, tx, nk & 1);
}

// ---------------------------------------------------------------------------------------------------------------
// shared memory: Qs | Ps | Kb[2] | Vb[2]  (Kb/Vb double as weight buffers after the attention loop)
constexpr int kSaSmemFloats = 2 * kTileF + 2 * kTileF + 2 * 4096;

__global__ void __launch_bounds__(kNT, 2)
td_sa_kernel(const float* __restrict__ x_in, const float* __restrict__ qkv, const ClipDesc* __restrict__ clips,
             int n_clips, const int* __restrict__ qtile_prefix /*64-row tiles*/, SaLayerParams P,
             float* __restrict__ x_out,
             const float* __restrict__ next_qkvT, const float* __restrict__ next_qkvb, float* __restrict__ qkv_next,
             PoolHeadParams H, int n_heads, float* __restrict__ logits) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                          // [64][68]  queries, later the row tile fed to the linears
  float* Ps = sm + kTileF;                 // [64][68]  softmax numerators, later scratch row tile
  float* Kb = Ps + kTileF;                 // [2][64][68]
  float* Vb = Kb + 2 * kTileF;             // [2][64][64]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int c = upper_slot(qtile_prefix, n_clips, blockIdx.x);
  const ClipDesc cd = clips[c];
  const int S = cd.n_seg;
  const int q0 = (blockIdx.x - __ldg(qtile_prefix + c)) * kT;
  const int rows_valid = min(kT, S - q0);
  const long long row0 = (long long)cd.seg_off + q0;
  const float* kv_base = qkv + (long long)cd.seg_off * 192;

  // ---- attention: Q tile, then key blocks of 64 (K row-major [key][d], V k-major [key][d])
  tile_load_async(Qs, kLd, qkv + row0 * 192, 192, rows_valid, tid);
  tile_load_async(Kb, kLd, kv_base + 64, 192, min(kT, S), tid);
  tile_load_async(Vb, 64, kv_base + 128, 192, min(kT, S), tid);
  cp_commit();
  float o[4][4];
  zero_acc(o);
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  const int n_kb = (S + kT - 1) / kT;
#pragma unroll 1
  for (int kb = 0; kb < n_kb; ++kb) {
    const int j0 = kb * kT;
    if (kb + 1 < n_kb) {
      const int nv = min(kT, S - (j0 + kT));
      tile_load_async(Kb + ((kb + 1) & 1) * kTileF, kLd, kv_base + (long long)(j0 + kT) * 192 + 64, 192, nv, tid);
      tile_load_async(Vb + ((kb + 1) & 1) * 4096, 64, kv_base + (long long)(j0 + kT) * 192 + 128, 192, nv, tid);
    }
    cp_commit();
    cp_wait<1>();
    __syncthreads();                               // K/V block kb (and Q) landed for every thread
    float s[4][4];
    zero_acc(s);
    gemm_nt(s, Qs, Kb + (kb & 1) * kTileF, ty, tx);           // s[i][j]: query 4ty+i, key j0 + tx + 16j
    const int nk = S - j0;                                     // keys >= nk of this block do not exist
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float bm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) { if (tx + 16 * j >= nk) s[i][j] = -INFINITY; bm = fmaxf(bm, s[i][j]); }
      bm = row_max(bm);                                        // every block holds >= 1 real key: bm is finite
      const float mn = fmaxf(m[i], bm);
      const float sc = expf(m[i] - mn);                        // first block: expf(-inf) = 0
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(s[i][j] - mn);                    // masked keys: expf(-inf) = 0
        ps += p;
        Ps[(4 * ty + i) * kLd + tx + 16 * j] = p;
      }
      l[i] = l[i] * sc + row_sum(ps);
      m[i] = mn;
      o[i][0] *= sc; o[i][1] *= sc; o[i][2] *= sc; o[i][3] *= sc;
    }
    __syncwarp();                                  // P rows 4ty..4ty+3 are written and read by this half-warp only
    gemm_nn(o, Ps, Vb + (kb & 1) * 4096, 64, ty, tx);         // o[i][j]: query 4ty+i, d = 4tx + j
    __syncthreads();                               // block kb consumed: its buffers are the prefetch target of kb + 2
  }
  // ---- out_proj + residual + LN1, FFN + residual + LN2; weights stream through Kb (as [64][64] chunks)
  float* Wb = Kb;                                  // two [64][64] weight buffers inside the K region (2 x 4096 <= 2 x kTileF)
  tile_load_async(Wb, 64, P.WoT, 64, 64, tid);
  cp_commit();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float inv = 1.0f / l[i];
    o[i][0] *= inv; o[i][1] *= inv; o[i][2] *= inv; o[i][3] *= inv;
  }
  store_rows_smem(Qs, o, ty, tx);                  // attention output tile (Q is no longer needed)
  tile_load_async(Wb + 4096, 64, P.W1T, 64, 64, tid);
  cp_commit();
  cp_wait<1>();
  __syncthreads();
  float v[4][4];
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(P.bo) + tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i][0] = bb.x; v[i][1] = bb.y; v[i][2] = bb.z; v[i][3] = bb.w; }
  }
  gemm_nn(v, Qs, Wb, 64, ty, tx);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = min(4 * ty + i, rows_valid - 1);                             // rows beyond the clip: any valid row
    const float4 xr = __ldg(reinterpret_cast<const float4*>(x_in + (row0 + r) * 64) + tx);
    v[i][0] += xr.x; v[i][1] += xr.y; v[i][2] += xr.z; v[i][3] += xr.w;
  }
  layernorm_rows(v, P.ln1_g, P.ln1_b, tx);
  store_rows_smem(Ps, v, ty, tx);                  // LN1 output: FFN input (and residual, kept in v)
  __syncthreads();                                 // Wb[0] (Wo) consumed, Ps complete
  tile_load_async(Wb, 64, P.W2T, 64, 64, tid);
  cp_commit();
  cp_wait<1>();                                    // W1 landed
  __syncthreads();
  float h[4][4];
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(P.b1) + tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) { h[i][0] = bb.x; h[i][1] = bb.y; h[i][2] = bb.z; h[i][3] = bb.w; }
  }
  gemm_nn(h, Ps, Wb + 4096, 64, ty, tx);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) h[i][j] = fmaxf(h[i][j], 0.f);
  store_rows_smem(Qs, h, ty, tx);                  // FFN hidden tile
  __syncthreads();                                 // Wb[1] (W1) consumed, Qs complete
  const bool last = next_qkvT == nullptr;
  // first chunk of the fused tail rides behind W2
  if (!last) tile_load_async(Wb + 4096, 64, next_qkvT, 64, 64, tid);
  else if (n_heads > 0) tile_load_async(Wb + 4096, 64, H.W1T, 128, 64, tid);      // (n_heads == 0: another pooling module follows)
  cp_commit();
  cp_wait<1>();                                    // W2 landed
  __syncthreads();
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(P.b2) + tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) { h[i][0] = bb.x; h[i][1] = bb.y; h[i][2] = bb.z; h[i][3] = bb.w; }
  }
  gemm_nn(h, Qs, Wb, 64, ty, tx);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] += h[i][j];
  layernorm_rows(v, P.ln2_g, P.ln2_b, tx);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * ty + i < rows_valid)
      *reinterpret_cast<float4*>(x_out + (row0 + 4 * ty + i) * 64 + 4 * tx) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  store_rows_smem(Ps, v, ty, tx);                  // layer output tile: input of the fused tail
  __syncthreads();                                 // Wb[0] (W2) consumed, Ps complete

  if (!last) {
    // ---- next layer's QKV projection (its part 0 is in flight into Wb[1])
#pragma unroll 1
    for (int part = 0; part < 3; ++part) {
      const int cur = (part + 1) & 1;
      if (part + 1 < 3) tile_load_async(Wb + (cur ^ 1) * 4096, 64, next_qkvT + (part + 1) * 4096, 64, 64, tid);
      cp_commit();
      cp_wait<1>();
      __syncthreads();
      float acc[4][4];
      const float4 bb = __ldg(reinterpret_cast<const float4*>(next_qkvb + part * 64) + tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
      gemm_nn(acc, Ps, Wb + cur * 4096, 64, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (4 * ty + i < rows_valid)
          *reinterpret_cast<float4*>(qkv_next + (row0 + 4 * ty + i) * 192 + part * 64 + 4 * tx) =
              make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();
    }
  } else {
    // ---- PoolAttFF logits of every head: chunk q = (head, half) of W1T [head][64 k][128 j]; chunk 0 is in flight
    const int n_chunks = 2 * n_heads;
    float part_logit[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int q = 0; q < n_chunks; ++q) {
      const int cur = (q + 1) & 1;
      if (q + 1 < n_chunks) {
        const int hn = (q + 1) >> 1, halfn = (q + 1) & 1;
        tile_load_async(Wb + (cur ^ 1) * 4096, 64, H.W1T + (size_t)hn * 64 * 128 + halfn * 64, 128, 64, tid);
      }
      cp_commit();
      cp_wait<1>();
      __syncthreads();
      const int hd = q >> 1, half = q & 1;
      float acc[4][4];
      const float4 bb = __ldg(reinterpret_cast<const float4*>(H.b1 + hd * 128 + half * 64) + tx);
      const float4 w2 = __ldg(reinterpret_cast<const float4*>(H.w2 + hd * 128 + half * 64) + tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
      gemm_nn(acc, Ps, Wb + cur * 4096, 64, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float t = part_logit[i];
        t = fmaf(w2.x, fmaxf(acc[i][0], 0.f), t); t = fmaf(w2.y, fmaxf(acc[i][1], 0.f), t);
        t = fmaf(w2.z, fmaxf(acc[i][2], 0.f), t); t = fmaf(w2.w, fmaxf(acc[i][3], 0.f), t);
        part_logit[i] = t;
      }
      if (half == 1) {
        const float b2 = __ldg(H.b2 + hd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float tot = row_sum(part_logit[i]);
          if (tx == 0 && 4 * ty + i < rows_valid) logits[(row0 + 4 * ty + i) * n_heads + hd] = tot + b2;
          part_logit[i] = 0.f;
        }
      }
      __syncthreads();
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Double-ended model (NISQA_DE, reference lib:272-424): time alignment of the reference clip's features to the degraded
// clip's (Alignment, lib:1228-1285) and feature fusion (Fusion, lib:1380-1417), for 64 degraded steps of one pair.
//   x = time_dependency output of the degraded clip (clip 2p), y = of the reference clip (clip 2p + 1)
//   score[i][j] : dot  x_i . y_j (AttDot, lib:1287-1296) | cosine similarity (AttCosine, lib:1298-1308, eps 1e-8) |
//                 -mean_d |x_id - y_jd| (AttDistance with its default norms, lib:1310-1323) | x_i . (W y_j + b) (AttLuong,
//                 lib:1344-1357) | v . tanh(Wq x_i + bq + Wy y_j + by) (AttBahdanau, att_dim 128, lib:1325-1342; its
//                 output bias shifts every score of a row alike and drops out of softmax / argmax); keys j >= n_wins_y masked
//   hard        : y_al[i] = y[argmax_j softmax(score[i])] = y[first maximal score] (ApplyHardAttention, lib:1359-1368)
//   soft        : y_al[i] = softmax_j(score[i]) . y (ApplySoftAttention, lib:1370-1378), online max / sum over key blocks
//   fuse        : [x, y_al, x - y_al] | [x + y_al, x - y_al] | [x, y_al]  ->  fused[row][64 * nf]
// Same 64 x 64 register tiling as td_sa_kernel; the y block serves as K (row-major, pitch 68) and as V.
enum { DE_ALIGN_DOT = 1, DE_ALIGN_COSINE = 2, DE_ALIGN_DISTANCE = 3, DE_ALIGN_LUONG = 4, DE_ALIGN_BAHD = 5 };      // = enum nisqa_de_align
enum { DE_FUSE_XY_MINUS = 0, DE_FUSE_PLUS_MINUS = 1, DE_FUSE_XY = 2 };

// s[i][j] -= sum_k |A[4ty+i][k] - B[tx+16j][k]|
__device__ __forceinline__ void absdiff_nt(float (&s)[4][4], const float* __restrict__ A, const float* __restrict__ B, int ty, int tx) {
  const float* a0 = A + (4 * ty) * kLd;
  const float* b0 = B + tx * kLd;
#pragma unroll 2
  for (int k = 0; k < kT; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * kLd + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(b0 + (16 * j) * kLd + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        s[i][j] -= (fabsf(a[i].x - b[j].x) + fabsf(a[i].y - b[j].y)) + (fabsf(a[i].z - b[j].z) + fabsf(a[i].w - b[j].w));
  }
}

constexpr int kDeSmemFloats = 2 * kTileF + 2 * kTileF + 2 * kT;      // Qs | Ps | Yb[2] | 1 / |y| of the two blocks
constexpr int kBhLd = 132;                                            // pitch of the [64][128] projections (AttBahdanau)
constexpr int kDeLuongFloats = 4096 + kTileF;                          // W^T | projected keys
constexpr int kDeBahdFloats = 64 * 128 + 2 * 64 * kBhLd;               // Wy^T | Wq x + bq | Wy y + by
// learned weights of AttLuong (wT [64][64] k-major, b [64]) / AttBahdanau (wqT, wyT [64][128] k-major, bq, by, v [128])
struct DeAlignParams { const float* wT; const float* b; const float* wqT; const float* bq; const float* wyT; const float* by; const float* v; };

__global__ void __launch_bounds__(kNT, 2)
de_align_kernel(const float* __restrict__ x_td /*[n_seg][64]*/, const ClipDesc* __restrict__ clips, int n_clips,
                const int* __restrict__ qtile_prefix /*64-row tiles*/, int align, int soft, int fuse, DeAlignParams A,
                float* __restrict__ fused /*[n_seg][64 nf]*/) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                          // [64][68] degraded rows x
  float* Ps = sm + kTileF;                 // [64][68] softmax numerators (soft attention)
  float* Yb = Ps + kTileF;                 // [2][64][68] reference rows y
  float* Yn = Yb + 2 * kTileF;             // [2][64] 1 / max(|y_j|, eps)
  float* Ex = Yn + 2 * kT;                 // AttLuong: W^T [64][64] | W y + b [64][68];  AttBahdanau: Wy^T [64][128] | XQ | YK [64][132]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int c = upper_slot(qtile_prefix, n_clips, blockIdx.x);
  if (c & 1) return;                       // reference clips are keys only
  const ClipDesc cd = clips[c], cr = clips[c + 1];
  const int Sx = cd.n_seg, Sy = cr.n_seg;
  if (Sx <= 0 || Sy <= 0) return;
  const int q0 = (blockIdx.x - __ldg(qtile_prefix + c)) * kT;
  const int rows_valid = min(kT, Sx - q0);
  const long long row0 = (long long)cd.seg_off + q0;
  const float* y_base = x_td + (long long)cr.seg_off * 64;
  const int nf = fuse == DE_FUSE_XY_MINUS ? 3 : 2;

  tile_load_async(Qs, kLd, x_td + row0 * 64, 64, rows_valid, tid);
  tile_load_async(Yb, kLd, y_base, 64, min(kT, Sy), tid);
  cp_commit();
  if (align == DE_ALIGN_LUONG) tile_load_async(Ex, 64, A.wT, 64, 64, tid);
  if (align == DE_ALIGN_BAHD) { tile_load_async(Ex, 128, A.wyT, 128, 64, tid); tile_load_async(Ex + 64, 128, A.wyT + 64, 128, 64, tid); }
  cp_commit();
  float o[4][4];
  zero_acc(o);
  if (align == DE_ALIGN_BAHD) {
    // XQ = Wq x + bq for the CTA's 64 degraded steps: two 64-column halves, Wq^T staged through Ps
    float* XQ = Ex + 64 * 128;
    cp_wait<0>();
    __syncthreads();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      tile_load_async(Ps, 64, A.wqT + half * 64, 128, 64, tid);
      cp_commit();
      cp_wait<0>();
      __syncthreads();
      float acc[4][4];
      const float4 bb = __ldg(reinterpret_cast<const float4*>(A.bq + half * 64) + tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
      gemm_nn(acc, Qs, Ps, 64, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(XQ + (4 * ty + i) * kBhLd + half * 64 + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();
    }
  }
  {
    float m[4], l[4], rq[4];
    int best[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; best[i] = 0; rq[i] = 1.f; }
    const int n_kb = (Sy + kT - 1) / kT;
#pragma unroll 1
    for (int kb = 0; kb < n_kb; ++kb) {
      const int j0 = kb * kT;
      if (kb + 1 < n_kb) tile_load_async(Yb + ((kb + 1) & 1) * kTileF, kLd, y_base + (long long)(j0 + kT) * 64, 64, min(kT, Sy - (j0 + kT)), tid);
      cp_commit();
      cp_wait<1>();
      __syncthreads();
      const float* Y = Yb + (kb & 1) * kTileF;
      if (align == DE_ALIGN_COSINE) {
        // nn.CosineSimilarity: x . y / (max(|x|, eps) max(|y|, eps)); four threads per key row, 16 features each
        const int r = tid >> 2, qd = tid & 3;
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Y + r * kLd + qd * 16 + k);
          ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        if (qd == 0) Yn[(kb & 1) * kT + r] = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
        if (kb == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float qs = 0.f;
            const float4 v = *reinterpret_cast<const float4*>(Qs + (4 * ty + i) * kLd + 4 * tx);
            qs = fmaf(v.x, v.x, qs); qs = fmaf(v.y, v.y, qs); qs = fmaf(v.z, v.z, qs); qs = fmaf(v.w, v.w, qs);
            rq[i] = 1.0f / fmaxf(sqrtf(row_sum(qs)), 1e-8f);
          }
        }
        __syncthreads();
      }
      float s[4][4];
      zero_acc(s);
      if (align == DE_ALIGN_LUONG) {
        // y' = W y + b for the block's keys, then x . y'
        float* Yp = Ex + 4096;
        float acc[4][4];
        const float4 bb = __ldg(reinterpret_cast<const float4*>(A.b) + tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
        gemm_nn(acc, Y, Ex, 64, ty, tx);
        store_rows_smem(Yp, acc, ty, tx);
        __syncthreads();
        gemm_nt(s, Qs, Yp, ty, tx);
      } else if (align == DE_ALIGN_BAHD) {
        const float* WyT = Ex;
        const float* XQ = Ex + 64 * 128;
        float* YK = Ex + 64 * 128 + 64 * kBhLd;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {                 // YK = Wy y + by for the block's keys
          float acc[4][4];
          const float4 bb = __ldg(reinterpret_cast<const float4*>(A.by + half * 64) + tx);
#pragma unroll
          for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
          gemm_nn(acc, Y, WyT + half * 64, 128, ty, tx);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(YK + (4 * ty + i) * kBhLd + half * 64 + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
#pragma unroll 1
        for (int a = 0; a < 128; a += 4) {
          const float4 vv = __ldg(reinterpret_cast<const float4*>(A.v + a));
          float4 xq[4], yk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xq[i] = *reinterpret_cast<const float4*>(XQ + (4 * ty + i) * kBhLd + a);
#pragma unroll
          for (int j = 0; j < 4; ++j) yk[j] = *reinterpret_cast<const float4*>(YK + (tx + 16 * j) * kBhLd + a);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float t = s[i][j];
              t = fmaf(vv.x, tanhf(xq[i].x + yk[j].x), t); t = fmaf(vv.y, tanhf(xq[i].y + yk[j].y), t);
              t = fmaf(vv.z, tanhf(xq[i].z + yk[j].z), t); t = fmaf(vv.w, tanhf(xq[i].w + yk[j].w), t);
              s[i][j] = t;
            }
        }
      } else if (align == DE_ALIGN_DISTANCE) {
        absdiff_nt(s, Qs, Y, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[i][j] *= (1.0f / 64.0f);
      } else {
        gemm_nt(s, Qs, Y, ty, tx);
        if (align == DE_ALIGN_COSINE) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float ry = Yn[(kb & 1) * kT + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i][j] = s[i][j] * rq[i] * ry;
          }
        }
      }
      const int nkeys = Sy - j0;
      if (!soft) {
        // first maximal score of the row: columns of a thread ascend with j and with the block; ties between threads
        // are resolved towards the smaller key index
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (tx + 16 * j < nkeys && s[i][j] > bv) { bv = s[i][j]; bi = j0 + tx + 16 * j; }
#pragma unroll
          for (int of = 8; of > 0; of >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, of);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, of);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          if (bv > m[i]) { m[i] = bv; best[i] = bi; }       // (strictly greater: an earlier block keeps a tie)
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float bm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 4; ++j) { if (tx + 16 * j >= nkeys) s[i][j] = -INFINITY; bm = fmaxf(bm, s[i][j]); }
          bm = row_max(bm);
          const float mn = fmaxf(m[i], bm);
          const float sc = expf(m[i] - mn);
          float ps = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float p = expf(s[i][j] - mn);
            ps += p;
            Ps[(4 * ty + i) * kLd + tx + 16 * j] = p;
          }
          l[i] = l[i] * sc + row_sum(ps);
          m[i] = mn;
          o[i][0] *= sc; o[i][1] *= sc; o[i][2] *= sc; o[i][3] *= sc;
        }
        __syncwarp();
        gemm_nn(o, Ps, Y, kLd, ty, tx);                      // o[i][d] += sum_j p[i][j] y[j][d]
      }
      __syncthreads();                                       // block kb consumed
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!soft) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(y_base + (long long)best[i] * 64) + tx);
        o[i][0] = v.x; o[i][1] = v.y; o[i][2] = v.z; o[i][3] = v.w;
      } else {
        const float inv = 1.0f / l[i];
        o[i][0] *= inv; o[i][1] *= inv; o[i][2] *= inv; o[i][3] *= inv;
      }
    }
  }
  // ---- fusion (lib:1402-1417)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (4 * ty + i >= rows_valid) continue;
    const float4 xv = *reinterpret_cast<const float4*>(Qs + (4 * ty + i) * kLd + 4 * tx);
    const float4 yv = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
    const float4 df = make_float4(xv.x - yv.x, xv.y - yv.y, xv.z - yv.z, xv.w - yv.w);
    float4* dst = reinterpret_cast<float4*>(fused + (row0 + 4 * ty + i) * (64 * nf)) + tx;
    if (fuse == DE_FUSE_XY_MINUS) { dst[0] = xv; dst[16] = yv; dst[32] = df; }
    else if (fuse == DE_FUSE_PLUS_MINUS) { dst[0] = make_float4(xv.x + yv.x, xv.y + yv.y, xv.z + yv.z, xv.w + yv.w); dst[16] = df; }
    else { dst[0] = xv; dst[16] = yv; }
  }
}

// scores of a double-ended pass: row 2p = the pair's score (NaN when either clip was skipped), row 2p + 1 = NaN
__global__ void de_finalize_kernel(const ClipDesc* __restrict__ clips, int n_clips, int n_out, float* __restrict__ scores) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clips) return;
  const bool bad = (c & 1) || clips[c].n_seg <= 0 || clips[c + 1].n_seg <= 0;
  if (bad) for (int h = 0; h < n_out; ++h) scores[c * n_out + h] = __int_as_float(0x7fc00000);
}


// ---------------------------------------------------------------------------------------------------------------
// Framewise models without convolutions (reference lib:504-583, user checkpoints): SkipCNN = BatchNorm2d(1) + flatten
// (+ Linear), DFF = BatchNorm2d(1) + flatten + 4 x (Linear + BatchNorm1d + ReLU).
//   seg_feats_kernel  : segment s -> row [768]: a * max(mel[frame0 + t][m], thr) + c at index m * 15 + t (x.view(-1, 720),
//                       lib:531 / 572), columns 720..767 zero (the rows feed 64-wide k chunks)
//   linear_tile_kernel: Y[n][N] = act(X[n][K] W^T + b), 64 x 64 output tiles, K in chunks of 64 (K, N multiples of 64),
//                       the same register-tiled fp32 product as the time-dependency block
__global__ void seg_feats_kernel(const float* __restrict__ mel, const int* __restrict__ seg_frame0,
                                 const float* __restrict__ seg_thr, const float* __restrict__ bn /*a, c*/, int n_seg,
                                 float* __restrict__ out /*[n_seg][768]*/) {
  const int s = blockIdx.x;
  if (s >= n_seg) return;
  const float* src = mel + (size_t)__ldg(seg_frame0 + s) * kMels;
  const float thr = __ldg(seg_thr + s), a = __ldg(bn), c = __ldg(bn + 1);
  for (int j = threadIdx.x; j < 768; j += blockDim.x) {
    float v = 0.f;
    if (j < kMels * kSegLen) {
      const int m = j / kSegLen, t = j - m * kSegLen;
      v = fmaf(a, fmaxf(__ldg(src + t * kMels + m), thr), c);
    }
    out[(size_t)s * 768 + j] = v;
  }
}

constexpr int kLinSmemFloats = 2 * kTileF + 2 * 4096;

__global__ void __launch_bounds__(kNT, 2)
linear_tile_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ WT /*[K][N]*/, const float* __restrict__ bias,
                   int relu, float* __restrict__ Y, int ldy, int n_rows, int K, int N) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                       // [2][64][68]
  float* Ws = sm + 2 * kTileF;          // [2][64][64]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long row0 = (long long)blockIdx.x * kT;
  const int n0 = blockIdx.y * kT;
  const int rows_valid = (int)min((long long)kT, (long long)n_rows - row0);
  const float* src = X + row0 * ldx;
  const int nk = K / kT;
  tile_load_async(As, kLd, src, ldx, rows_valid, tid);
  tile_load_async(Ws, 64, WT + n0, N, 64, tid);
  cp_commit();
  float acc[4][4];
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0) + tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
  }
#pragma unroll 1
  for (int c = 0; c < nk; ++c) {
    if (c + 1 < nk) {
      tile_load_async(As + ((c + 1) & 1) * kTileF, kLd, src + (c + 1) * 64, ldx, rows_valid, tid);
      tile_load_async(Ws + ((c + 1) & 1) * 4096, 64, WT + (size_t)(c + 1) * 64 * N + n0, N, 64, tid);
    }
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    gemm_nn(acc, As + (c & 1) * kTileF, Ws + (c & 1) * 4096, 64, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * ty + i < rows_valid) {
      float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(Y + (row0 + 4 * ty + i) * ldy + n0 + 4 * tx) = v;
    }
}

// ------------------------------------------------------------------ host launchers
void launch_td_in(cudaStream_t st, const float* feats, const float* WT, int nk, const float* b, const float* g, const float* be,
                  const float* qkvT, const float* qkvb, const float* pe, const int* seg_clip, const ClipDesc* clips,
                  float* x0, float* qkv, int n_rows) {
  static unsigned long long cfg = 0;
  const int smem = kInSmemFloats * 4;
  if (first_launch_on_device(cfg)) cudaFuncSetAttribute(td_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  td_in_kernel<<<(n_rows + kT - 1) / kT, kNT, smem, st>>>(feats, WT, nk, b, g, be, qkvT, qkvb, pe, seg_clip, clips, x0, qkv, n_rows);
}

void launch_de_align(cudaStream_t st, const float* x_td, const ClipDesc* clips, int n_clips, const int* qtile64_prefix,
                     int n_qtiles, int align, int soft, int fuse, const DeAlignParams& A, float* fused) {
  static unsigned long long cfg = 0;
  if (first_launch_on_device(cfg))
    cudaFuncSetAttribute(de_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kDeSmemFloats + kDeBahdFloats) * 4);
  const int smem = (kDeSmemFloats + (align == DE_ALIGN_LUONG ? kDeLuongFloats : align == DE_ALIGN_BAHD ? kDeBahdFloats : 0)) * 4;
  de_align_kernel<<<n_qtiles, kNT, smem, st>>>(x_td, clips, n_clips, qtile64_prefix, align, soft, fuse, A, fused);
}

void launch_seg_feats(cudaStream_t st, const float* mel, const int* seg_frame0, const float* seg_thr, const float* bn, int n_seg,
                      float* out) {
  if (n_seg > 0) seg_feats_kernel<<<n_seg, 256, 0, st>>>(mel, seg_frame0, seg_thr, bn, n_seg, out);
}

void launch_linear_tile(cudaStream_t st, const float* X, int ldx, const float* WT, const float* bias, int relu, float* Y, int ldy,
                        int n_rows, int K, int N) {
  static unsigned long long cfg = 0;
  const int smem = kLinSmemFloats * 4;
  if (first_launch_on_device(cfg)) cudaFuncSetAttribute(linear_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const dim3 grid((n_rows + kT - 1) / kT, N / kT);
  linear_tile_kernel<<<grid, kNT, smem, st>>>(X, ldx, WT, bias, relu, Y, ldy, n_rows, K, N);
}

void launch_de_finalize(cudaStream_t st, const ClipDesc* clips, int n_clips, int n_out, float* scores) {
  de_finalize_kernel<<<(n_clips + 127) / 128, 128, 0, st>>>(clips, n_clips, n_out, scores);
}

void launch_td_sa(cudaStream_t st, const float* x_in, const float* qkv, const ClipDesc* clips, int n_clips,
                  const int* qtile64_prefix, int n_qtiles, const SaLayerParams& P, float* x_out,
                  const float* next_qkvT, const float* next_qkvb, float* qkv_next,
                  const PoolHeadParams& H, int n_heads, float* logits) {
  static unsigned long long cfg = 0;
  const int smem = kSaSmemFloats * 4;
  if (first_launch_on_device(cfg)) cudaFuncSetAttribute(td_sa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  td_sa_kernel<<<n_qtiles, kNT, smem, st>>>(x_in, qkv, clips, n_clips, qtile64_prefix, P, x_out,
                                            next_qkvT, next_qkvb, qkv_next, H, n_heads, logits);
}

}  // namespace nisqa
