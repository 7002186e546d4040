// Pipe-throughput microbenchmarks for the local-attention softmax design (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu && tools/ubench/pipes
// Every kernel: one CTA per SM, W warps, each thread runs R rounds of an unrolled body over 16 independent registers.
// Reports operations per clock per SM from clock64() deltas (independent of the SM clock).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }

// degree-3 exp2 on the FMA/ALU pipes: x <= 0 (softmax numerators); 2^x = 2^floor(x) * p(x - floor(x))
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                   // 1.5 * 2^23: round to nearest integer in the low mantissa bits
  const float j = t - 12582912.0f;
  const float f = x - j;                             // [-0.5, 0.5]
  float p = fmaf(f, 0.05550411f, 0.24022651f);
  p = fmaf(p, f, 0.69314718f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float tanh_f32(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void bench(float* out, long long* cyc, int rounds, float a, float b) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = -0.001f * float(threadIdx.x + i);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) r[i] = ex2(r[i]) - 1.0f;                                   // MUFU + FADD
      if (MODE == 1) r[i] = ex2(fmaf(r[i], a, b));                              // FFMA + MUFU (the softmax body)
      if (MODE == 2) r[i] = fmaf(r[i], a, b);                                   // FFMA only
      if (MODE == 3) r[i] = ex2_poly(fmaf(r[i], a, b));                         // polynomial exp2
      if (MODE == 5) r[i] = fmaxf(r[i] * a, b);                                 // FMUL + FMNMX
      if (MODE == 9) r[i] = __uint_as_float(ex2_f16x2(__float_as_uint(r[i])) & 0xbbffbbffu);     // 2 half exponentials per instruction
      if (MODE == 10) r[i] = __uint_as_float(ex2_bf16x2(__float_as_uint(r[i])) & 0xbf7fbf7fu);   // 2 bf16 exponentials per instruction
      if (MODE == 11) r[i] = __uint_as_float(tanh_bf16x2(__float_as_uint(r[i])));                // 2 bf16 tanh per instruction
      if (MODE == 12) r[i] = tanh_f32(r[i]) + a;                                                 // fp32 tanh
    }
    if (MODE == 4) {                                                            // F2FP only
#pragma unroll
      for (int i = 0; i < 16; i += 2) { acc ^= pack(r[i], r[i + 1]); r[i] = __uint_as_float(acc | 0x3f000000u); }
    }
    if (MODE == 6) {                                                            // softmax body: FFMA + MUFU + FADD + pack
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { r[i] = ex2(fmaf(r[i], a, b)); s += r[i]; }
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(r[i], r[i + 1]);
      r[0] += s * 1e-30f;
    }
    if (MODE == 7) {                                                            // same, 4 of 16 exponentials on the polynomial
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float x = fmaf(r[i], a, b); r[i] = (i & 3) == 3 ? ex2_poly(x) : ex2(x); s += r[i]; }
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(r[i], r[i + 1]);
      r[0] += s * 1e-30f;
    }
    if (MODE == 8) {                                                            // same, 8 of 16 on the polynomial
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float x = fmaf(r[i], a, b); r[i] = (i & 1) ? ex2_poly(x) : ex2(x); s += r[i]; }
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(r[i], r[i + 1]);
      r[0] += s * 1e-30f;
    }
  }
  const long long t1 = clock64();
  float s = __uint_as_float(acc & 0xff);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, float ops_per_elem) {
  const int rounds = 2000, sms = 148;
  float* out; long long* cyc;
  cudaMalloc(&out, sms * warps * 32 * sizeof(float));
  cudaMalloc(&cyc, sms * sizeof(long long));
  bench<MODE><<<sms, warps * 32>>>(out, cyc, rounds, 0.999f, -0.0001f);
  bench<MODE><<<sms, warps * 32>>>(out, cyc, rounds, 0.999f, -0.0001f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += double(h[i]); avg /= sms;
  const double elems = double(rounds) * 16 * warps * 32;
  printf("%-46s warps/SM %2d: %7.2f elements/clk/SM  (%6.2f cycles per warp-element-instr group; %s)\n", name, warps, elems / avg,
         avg / (double(rounds) * 16 * warps / 4.0), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {8, 16}) {
    printf("---- %d warps per SM (%d per SMSP)\n", w, w / 4);
    run<0>("MUFU.EX2 + FADD", w, 1);
    run<1>("FFMA + MUFU.EX2", w, 1);
    run<2>("FFMA", w, 1);
    run<5>("FMUL + FMNMX", w, 1);
    run<3>("FFMA + polynomial exp2 (FMA/ALU pipes)", w, 1);
    run<4>("F2FP.BF16 pack (per pair; elements = 2x)", w, 1);
    run<9>("ex2.approx.f16x2 (elements = 2x shown)", w, 1);
    run<10>("ex2.approx.ftz.bf16x2 (elements = 2x shown)", w, 1);
    run<11>("tanh.approx.bf16x2 (elements = 2x shown)", w, 1);
    run<12>("tanh.approx.f32 + FADD", w, 1);
    run<6>("softmax body: FFMA+EX2+FADD+pack/2", w, 1);
    run<7>("softmax body, 25% polynomial", w, 1);
    run<8>("softmax body, 50% polynomial", w, 1);
  }
  return 0;
}
