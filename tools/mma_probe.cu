// mma_probe.cu - GPU-box microbenchmark: what does one tcgen05.mma.kind::f16 cost from one issuing thread,
// as a function of N, of the accumulator dependency pattern and of the number of co-resident CTAs?
//
// Round-1 finding this probe is meant to explain: in the conv kernels 144 MMAs of
// M=128, N=128/64, K=16 take ~107 cycles each, independent of operand layout / alignment / accumulator
// interleaving, while the math floor is 64 / 32 cycles.  If the per-instruction cost is flat in N, the
// kernels want fewer and larger MMAs (N = 256).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nisqa_b200/csrc tools/mma_probe.cu -o /tmp/mma_probe
//   /tmp/mma_probe            (prints one table; a few seconds)
//
// Each CTA: one thread issues REPS MMAs (A: 128 x 16 halves, B: N x 16 halves, both K-major without swizzle,
// zero-filled shared memory) into `chains` TMEM accumulators used round-robin, commits to an mbarrier and
// waits; clock64 around issue -> completion.  Grid = 148 x ctas_per_sm CTAs, all timed, median reported.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "tc_ptx.cuh"

using namespace nisqa;

__global__ void __launch_bounds__(128)
probe_kernel(int n, int chains, int reps, int a_stride_rows, long long* out /*[grid][2]*/) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  // layout: A [2 planes][AROWS rows][16 B], B [2 planes][256 rows][16 B], barrier, tmem slot
  constexpr int AROWS = 128 + 64;
  const uint32_t a_base = sbase, b_base = sbase + 2 * AROWS * 16, bar = b_base + 2 * 256 * 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * AROWS * 16 + 2 * 256 * 16 + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (2 * AROWS * 16 + 2 * 256 * 16) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512 / 2);          // 256 columns: two CTAs fit on an SM
  if (tid == 32) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t db = make_desc(b_base, 256 * 16, 128);
    const int cols_per_chain = n;                                   // chains * n <= 256
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int c = r % chains;
      // a different 16-byte aligned A start per MMA (like the tap shifts of the conv kernels)
      const uint64_t da = make_desc(a_base + (uint32_t)((r * a_stride_rows) % 64) * 16, AROWS * 16, 128);
      umma_f16(tmem + c * cols_per_chain, da, db, idesc, r >= chains);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2 + 0] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static double median(std::vector<long long> v) {
  std::sort(v.begin(), v.end());
  return (double)v[v.size() / 2];
}

int main() {
  const int reps = 256;
  const int smem = 2 * 192 * 16 + 2 * 256 * 16 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d_out;
  cudaMalloc(&d_out, 148 * 4 * 2 * sizeof(long long));
  printf("M=128 K=16 kind::f16, %d MMAs per CTA, cycles per MMA (issue loop / until retired), median over CTAs\n", reps);
  printf("%6s %7s %9s %8s | %10s %10s | floor\n", "N", "chains", "CTAs/SM", "Ashift", "issue", "retired");
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int n : {32, 64, 128, 256})
      for (int chains : {1, 2, 4})
        for (int shift : {0, 3}) {
          if (chains * n > 256) continue;
          const int grid = 148 * ctas;
          for (int it = 0; it < 3; ++it)       // warm, then the measured launch
            probe_kernel<<<grid, 128, smem>>>(n, chains, reps, shift, d_out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          std::vector<long long> h(grid * 2);
          cudaMemcpy(h.data(), d_out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
          std::vector<long long> a, b;
          for (int i = 0; i < grid; ++i) { a.push_back(h[2 * i]); b.push_back(h[2 * i + 1]); }
          printf("%6d %7d %9d %8d | %10.1f %10.1f | %d\n", n, chains, ctas, shift, median(a) / reps, median(b) / reps, n / 2);
        }
  cudaFree(d_out);
  return 0;
}
