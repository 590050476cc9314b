// conv1_cell.cuh - conv1(1->16) + BN + ReLU fused with the first max-pool for ONE pooled cell of one segment (all 16
// channels), shared by conv1_pool1_kernel (cnn.cu) and the fused conv1 + conv2 kernel (conv12.cu).
//   MODE 0 (adapt, reference lib:690-691): adaptive_max_pool2d 48x15 -> 24x7 : rows {2i,2i+1}, cols [2j,2j+3)
//   MODE 1 (standard, lib:813-814): MaxPool2d(2, stride 2, padding (0,1)) -> 24x8 : cols {2j-1,2j}
#pragma once
#include "common.cuh"

namespace nisqa {

// ws: [9][16] folded conv1 weights followed by the 16 biases (shared memory).  LDG: `mel` is global memory read through
// the read-only path; false: a shared-memory copy of the segment's 15 mel rows (f0 = 0).
template <int MODE, bool LDG = true>
__device__ __forceinline__ void conv1_cell(const float* __restrict__ mel, int f0, float thr, const float* ws,
                                           int ph, int pw, float (&res)[16]) {
  constexpr int NWC = (MODE == 0) ? 3 : 2;       // window columns
  constexpr int PC = NWC + 2;                    // patch columns
  const int r0 = 2 * ph - 1;                     // first patch row (mel index)
  const int c0 = (MODE == 0) ? 2 * pw - 1 : 2 * pw - 2;   // first patch col (frame in segment)
  float patch[4][PC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < PC; ++j) {
      const int r = r0 + i, t = c0 + j;
      float v = 0.f;                             // zero padding of the segment's own border
      if (r >= 0 && r < kMels && t >= 0 && t < kSegLen)
        v = fmaxf(LDG ? __ldg(mel + (size_t)(f0 + t) * kMels + r) : mel[(f0 + t) * kMels + r], thr);
      patch[i][j] = v;
    }

#pragma unroll
  for (int cq = 0; cq < 4; ++cq) {
    float acc[2][NWC][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NWC; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float4 w = *reinterpret_cast<const float4*>(ws + tap * 16 + cq * 4);
      const int ky = tap / 3, kx = tap % 3;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NWC; ++j) {
          const float a = patch[i + ky][j + kx];
          acc[i][j][0] = fmaf(a, w.x, acc[i][j][0]);
          acc[i][j][1] = fmaf(a, w.y, acc[i][j][1]);
          acc[i][j][2] = fmaf(a, w.z, acc[i][j][2]);
          acc[i][j][3] = fmaf(a, w.w, acc[i][j][3]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NWC; ++j) {
          const int col = c0 + 1 + j;            // conv output column of this window slot
          if (MODE == 0 || (col >= 0 && col < kSegLen)) m = fmaxf(m, acc[i][j][c]);
        }
      res[cq * 4 + c] = fmaxf(m + ws[144 + cq * 4 + c], 0.f);   // bias + ReLU commute with max
    }
  }
}

}  // namespace nisqa
