// ffma2_probe.cu - issue / pipe rate of the packed fp32 instructions of sm_100 (fma / add / mul.rn.f32x2 -> FFMA2 / FADD2 /
// FMUL2) against their scalar forms, alone and mixed with integer work, with operands as a real kernel has them
// (distinct registers, scalar-broadcast form).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe.bin tools/ffma2_probe.cu && tools/ffma2_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }

constexpr int ITERS = 2048, N = 16;     // N packed values (or 2N scalars) per thread

// MODE 0: scalar FFMA x[i] = x[i]*y[i]+z[i] (distinct regs)   1: FFMA2 same
//      2: scalar FADD ring x[i] += x[i+1]                      3: FADD2 ring
//      4: scalar FMUL                                           5: FMUL2
//      6: scalar FFMA broadcast a*w+acc (conv/gemm shape)       7: FFMA2 broadcast (R.F32 operand form)
//      8: butterfly scalar (add, sub, cmul const)               9: butterfly packed
template <int MODE>
__global__ void k(float* out, float s) {
  float a[2 * N], y[2 * N];
  u64 p[N], q[N];
  for (int i = 0; i < 2 * N; ++i) { a[i] = threadIdx.x * 0.001f + i; y[i] = 1.0f + 1e-6f * (i + threadIdx.x); }
  for (int i = 0; i < N; ++i) { p[i] = pk(a[2 * i], a[2 * i + 1]); q[i] = pk(y[2 * i], y[2 * i + 1]); }
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 2 * N; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(y[i]), "f"(y[(i + 1) % (2 * N)]));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q[i]), "l"(q[(i + 1) % N]));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 2 * N; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(y[i]));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < N; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 2 * N; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(y[i]));
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < N; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
    } else if (MODE == 6) {
#pragma unroll
      for (int i = 0; i < 2 * N; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(y[i & 3]), "f"(y[4 + (i & 7)]));
    } else if (MODE == 7) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const u64 b = pk(y[i & 3], y[i & 3]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(b), "l"(q[i & 3]));
      }
    } else if (MODE == 8) {     // N/2 radix-2 butterflies with a constant twiddle on the difference: 4 add + 2 mul + 2 fma
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        float ax = a[2 * i], ay = a[2 * i + 1], bx = a[2 * i + 2], by = a[2 * i + 3];
        float sx, sy, dx, dy, tx, ty;
        asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(sx) : "f"(ax), "f"(bx));
        asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(sy) : "f"(ay), "f"(by));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dx) : "f"(ax), "f"(bx));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dy) : "f"(ay), "f"(by));
        asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(tx) : "f"(dy), "f"(0.38268343f));
        asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(ty) : "f"(dx), "f"(-0.38268343f));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(tx) : "f"(dx), "f"(0.92387953f));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(ty) : "f"(dy), "f"(0.92387953f));
        a[2 * i] = sx; a[2 * i + 1] = sy; a[2 * i + 2] = tx; a[2 * i + 3] = ty;
      }
    } else {                    // packed: add2, sub2, mul2 (broadcast d.x), fma2 (broadcast d.y)
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        u64 sm, df, t;
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(sm) : "l"(p[i]), "l"(p[i + 1]));
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(df) : "l"(p[i]), "l"(p[i + 1]));
        const float2 d = upk(df);
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk(d.x, d.x)), "l"(pk(0.92387953f, -0.38268343f)));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(pk(d.y, d.y)), "l"(pk(0.38268343f, 0.92387953f)));
        p[i] = sm; p[i + 1] = t;
      }
    }
  }
  float acc = 0.f;
  for (int i = 0; i < 2 * N; ++i) acc += a[i];
  for (int i = 0; i < N; ++i) { float2 v = upk(p[i]); acc += v.x + v.y; }
  if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
static void run(const char* name, int warps_per_smsp, double ops_per_iter /*scalar-equivalent fp32 ops per thread per iteration*/) {
  float* d; cudaMalloc(&d, 4);
  const int sms = 148;
  dim3 grid(sms), block(32 * 4 * warps_per_smsp);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, block>>>(d, 1.0001f);
  cudaEventRecord(e0);
  k<MODE><<<grid, block>>>(d, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * block.x * ITERS * ops_per_iter;
  printf("%-34s warps/SMSP %d: %.3f ms  %6.1f fp32 lane-ops / clk / SM (1.965 GHz; 128 = one scalar instruction per cycle and SMSP)\n",
         name, warps_per_smsp, ms, ops / (ms * 1e-3) / 1.965e9 / sms);
  cudaFree(d);
}

int main() {
  for (int w : {1, 4}) {
    run<0>("FFMA  distinct operands", w, 2 * N);
    run<1>("FFMA2 distinct operands", w, 2 * N);
    run<2>("FADD", w, 2 * N);
    run<3>("FADD2", w, 2 * N);
    run<4>("FMUL", w, 2 * N);
    run<5>("FMUL2", w, 2 * N);
    run<6>("FFMA  broadcast a*w+acc", w, 2 * N);
    run<7>("FFMA2 broadcast a*w+acc", w, 2 * N);
    run<8>("butterfly scalar (8 instr)", w, 8.0 * N / 2);
    run<9>("butterfly packed (4 instr)", w, 8.0 * N / 2);
  }
  return 0;
}
