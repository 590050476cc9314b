// f32x2.cuh - packed fp32 arithmetic of sm_100 (PTX fma/add/sub/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2): two
// independent IEEE fp32 operations per instruction on an aligned register pair.  The fp32 pipe's lane rate is unchanged
// (measured 122 of 128 FMA lanes / clk / SM with FFMA2 vs 100 with scalar FFMA, profiles/r02o_ffma2_probe.txt) but
// every packed instruction frees an issue slot, which is what bounds the front-end FFT, conv1 and the fp32 tile GEMMs
// of the time-dependency block.  Every lane result is the correctly rounded fp32 result of the same scalar operation:
// a kernel that replaces fmaf(a, b, c) pairs by fma2 is bit-identical.  ptxas folds a pair built from one scalar
// (`pk(a, a)`) into the instruction's scalar-broadcast operand form (`R.F32`): no extra moves.
#pragma once
#include <cuda_runtime.h>

namespace nisqa {

typedef unsigned long long f2;      // two packed floats: .x in the low register, .y in the high one

__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 pk(float2 v) { return pk(v.x, v.y); }
__device__ __forceinline__ f2 bc(float a) { return pk(a, a); }
__device__ __forceinline__ float2 upk(f2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// complex product d * w of packed (re, im) values; `wr` = (w.x, w.y), `wi` = (-w.y, w.x)
__device__ __forceinline__ f2 cmul2(f2 d, f2 wr, f2 wi) {
  const float2 s = upk(d);
  return fma2(bc(s.y), wi, mul2(bc(s.x), wr));
}

}  // namespace nisqa
