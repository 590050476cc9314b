// mma_probe.cu - GPU-box microbenchmark: what does one tcgen05.mma.kind::f16 cost, as a function of N, of the number
// of co-resident CTAs, and of the number of ISSUING THREADS inside one CTA?
//
// Round-2 first call (profiles/r02a_mma_probe.txt): from ONE issuing thread a M=128 K=16 MMA retires every ~142
// cycles whatever N is (32 .. 256) and whatever the accumulator pattern; two co-resident CTAs each get their own 142
// (the SM then retires one MMA per 71 cycles) until the math floor binds (N = 256: 251 per CTA = 2 x 128).  So the
// tensor pipe itself is not the limit of a small-N instruction stream - the per-CTA (or per-thread?) issue chain is.
// This version answers: do MMAs issued by DIFFERENT warps of the same CTA (into disjoint accumulators) overlap
// like MMAs of different CTAs do?  And how does it look with 3 / 4 small CTAs per SM?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nisqa_b200/csrc tools/mma_probe.cu -o /tmp/mma_probe
//   /tmp/mma_probe            (prints one table; a few seconds)
//
// Each CTA: `issuers` threads (lane 0 of warps 0..issuers-1) each issue REPS MMAs (A: 128 x 16 halves, B: N x 16
// halves, K-major without swizzle, zero-filled shared memory, private A/B copies per issuer) into their own TMEM
// columns, commit to their own mbarrier and wait; clock64 around issue -> completion.  Grid = 148 x ctas_per_sm.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "tc_ptx.cuh"

using namespace nisqa;

constexpr int AROWS = 128 + 64;
constexpr int A_BYTES = 2 * AROWS * 16, B_BYTES = 2 * 256 * 16;     // two 16-byte K chunks each

__global__ void __launch_bounds__(128)
probe_kernel(int n, int issuers, int reps, int tmem_cols, int swz, long long* out /*[grid][4][2]*/) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int per_issuer = swz ? (AROWS * 128 + 256 * 128) : (A_BYTES + B_BYTES);
  const uint32_t bar0 = sbase + issuers * per_issuer;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + issuers * per_issuer + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < issuers * per_issuer / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  if (tid == 32) { for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8 * i, 1); fence_barrier_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if ((tid & 31) == 0 && warp < issuers) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = sbase + warp * per_issuer;
    const uint32_t b_base = a_base + (swz ? AROWS * 128 : A_BYTES);
    const uint64_t db = swz ? make_desc_swz(b_base, 8 * 128, 2u) : make_desc(b_base, 256 * 16, 128);
    const uint32_t d = tmem + warp * n;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // a different (whole-row shifted) A start per MMA, like the tap shifts of the conv kernels
      const uint64_t da = swz ? make_desc_swz(a_base + (uint32_t)(r % 16) * 128 + (uint32_t)(r & 3) * 32, 8 * 128, 2u)
                              : make_desc(a_base + (uint32_t)((r * 3) % 64) * 16, AROWS * 16, 128);
      umma_f16(d, da, db, idesc, r > 0);
    }
    const long long t1 = clock64();
    umma_commit(bar0 + 8 * warp);
    mbar_wait(bar0 + 8 * warp, 0);
    const long long t2 = clock64();
    out[(blockIdx.x * 4 + warp) * 2 + 0] = t1 - t0;
    out[(blockIdx.x * 4 + warp) * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

static double median(std::vector<long long> v) {
  std::sort(v.begin(), v.end());
  return (double)v[v.size() / 2];
}

int main() {
  const int reps = 256;
  const int smem_max = 4 * (AROWS * 128 + 256 * 128) + 256 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  long long* d_out;
  cudaMalloc(&d_out, 148 * 4 * 4 * 2 * sizeof(long long));
  printf("M=128 K=16 kind::f16, %d MMAs per issuing thread; cycles per MMA of ONE issuer (issue loop / until retired), median;\n", reps);
  printf("SM rate = cycles per MMA seen by the SM = retired / (CTAs/SM x issuers);  floor = N/2\n");
  printf("%4s %6s %8s %8s | %9s %9s | %8s | %5s\n", "swz", "N", "CTAs/SM", "issuers", "issue", "retired", "SM rate", "floor");
  for (int swz = 0; swz <= 1; ++swz)
    for (int n : {64, 128, 256})
      for (int ctas : {1, 2, 3, 4})
        for (int issuers : {1, 2, 4}) {
          if (issuers * n > 512 / ctas) continue;                       // TMEM columns per CTA
          int cols = 32;
          while (cols < issuers * n) cols *= 2;
          if (cols * ctas > 512) continue;
          const int per_issuer = swz ? (AROWS * 128 + 256 * 128) : (A_BYTES + B_BYTES);
          const int smem = issuers * per_issuer + 256 + 1024;
          if ((size_t)smem * ctas > 220 * 1024) continue;
          const int grid = 148 * ctas;
          for (int it = 0; it < 3; ++it) probe_kernel<<<grid, 128, smem>>>(n, issuers, reps, cols, swz, d_out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          std::vector<long long> h(grid * 4 * 2);
          cudaMemcpy(h.data(), d_out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
          std::vector<long long> a, b;
          for (int i = 0; i < grid; ++i)
            for (int w = 0; w < issuers; ++w) { a.push_back(h[(i * 4 + w) * 2]); b.push_back(h[(i * 4 + w) * 2 + 1]); }
          const double ret = median(b) / reps;
          printf("%4d %6d %8d %8d | %9.1f %9.1f | %8.1f | %5d\n", swz, n, ctas, issuers, median(a) / reps, ret, ret / (ctas * issuers), n / 2);
        }
  cudaFree(d_out);
  return 0;
}
