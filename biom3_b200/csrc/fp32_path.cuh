// Kernels of the fp32-class forward (precision 1): activations stay fp32 in HBM, every dense contraction runs on
// the tcgen05 GEMM as a three-product bf16 split (x = hi + lo, hi.Whi + hi.Wlo + lo.Whi, fp32 accumulate; see
// Params::split3), attention runs in fp32 on the CUDA cores.  Target: logits within 1e-4 relative of the fp32
// reference (BASELINE.json north_star); speed is secondary here, the bf16 path is the throughput mode.
//
// Split layout: a "[hi | lo]" buffer of an [M][K] fp32 matrix is bf16 [M][2K]: columns [0, K) hold bf16(x),
// columns [K, 2K) hold bf16(x - bf16(x)).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace f32p {

constexpr int DH = 32;
constexpr int WIN = 128;

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// 4 consecutive fp32 -> 4 hi at `hi_dst`, 4 lo at `hi_dst + K`
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi_dst, int K, float4 v) {
  __nv_bfloat16 h[4], l[4];
  split_bf16(v.x, h[0], l[0]);
  split_bf16(v.y, h[1], l[1]);
  split_bf16(v.z, h[2], l[2]);
  split_bf16(v.w, h[3], l[3]);
  *reinterpret_cast<uint2*>(hi_dst) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(hi_dst + K) = *reinterpret_cast<const uint2*>(l);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm (eps 1e-5, two-pass statistics like torch) + split: u fp32 [rows][D] -> a2 bf16 [rows][2D].
// One warp per row, D <= 1024, D % 128 == 0.
__global__ void __launch_bounds__(256)
ln_split_kernel(const float* __restrict__ u, const float* __restrict__ gamma, const float* __restrict__ beta,
                __nv_bfloat16* __restrict__ a2, int rows, int D) {
  const int lane = threadIdx.x & 31;
  const int nv = D / 128;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        v[i] = *reinterpret_cast<const float4*>(u + size_t(row) * D + (i * 32 + lane) * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    const float mean = warp_sum_f(s) / float(D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const float x0 = v[i].x - mean, x1 = v[i].y - mean, x2 = v[i].z - mean, x3 = v[i].w - mean;
        q += (x0 * x0 + x1 * x1) + (x2 * x2 + x3 * x3);
      }
    const float rstd = 1.0f / sqrtf(warp_sum_f(q) / float(D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
        const float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                     (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        store_split4(a2 + size_t(row) * 2 * D + col, D, o);
      }
  }
}

// h fp32 [rows][N] -> split(gelu_erf(h + bias)) bf16 [rows][2N]   (exact erf form, torch's default nn.GELU)
__global__ void __launch_bounds__(256)
bias_gelu_split_kernel(const float* __restrict__ h, const float* __restrict__ bias, __nv_bfloat16* __restrict__ h2,
                       size_t rows, int N) {
  const size_t n4 = rows * size_t(N / 4);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const size_t row = i / size_t(N / 4);
    const int col = int(i % size_t(N / 4)) * 4;
    const float4 x = *reinterpret_cast<const float4*>(h + row * N + col);
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col));
    auto g = [](float t) { return 0.5f * t * (1.0f + erff(t * 0.70710678118654752440f)); };
    const float4 o = make_float4(g(x.x + b.x), g(x.y + b.y), g(x.z + b.z), g(x.w + b.w));
    store_split4(h2 + row * 2 * N + col, N, o);
  }
}

// Windowed softmax attention, fp32.  qkv fp32 row-major [B*L][3*D] (column = which*D + h*32 + d), heads [0, NL).
// grid (L/128, NL, B), 128 threads: thread = one query row; K and V of one key window at a time in shared memory
// (every thread reads the same key -> broadcast), online softmax over chunks of 16 keys.  Windows w-1 and w+1 that
// fall off the sequence are skipped, which equals the reference's -FLT_MAX fill (their exp underflows to exactly 0).
// Output: split [hi | lo] of the [B*L][D] attention matrix, columns h*32 .. h*32+31.
__global__ void __launch_bounds__(128)
local_attention_f32_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ att2, int B, int H, int L,
                           float scale) {
  __shared__ __align__(16) float sk[WIN][DH];
  __shared__ __align__(16) float sv[WIN][DH];
  const int w = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * DH;
  const int nw = L / WIN;
  const int tid = threadIdx.x;
  const size_t row = size_t(b) * L + size_t(w) * WIN + tid;
  float q[DH], acc[DH];
  {
    const float* qp = qkv + row * 3 * D + h * DH;
#pragma unroll
    for (int i = 0; i < DH / 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(qp + 4 * i);
      q[4 * i] = t.x * scale; q[4 * i + 1] = t.y * scale; q[4 * i + 2] = t.z * scale; q[4 * i + 3] = t.w * scale;
    }
  }
#pragma unroll
  for (int i = 0; i < DH; ++i) acc[i] = 0.f;
  float mx = -INFINITY, den = 0.f;
  for (int kw = w - 1; kw <= w + 1; ++kw) {
    if (kw < 0 || kw >= nw) continue;
    __syncthreads();
    for (int i = tid; i < WIN * DH / 4; i += 128) {
      const int r = i / (DH / 4), c = (i % (DH / 4)) * 4;
      const float* src = qkv + (size_t(b) * L + size_t(kw) * WIN + r) * 3 * D + h * DH + c;
      *reinterpret_cast<float4*>(&sk[r][c]) = *reinterpret_cast<const float4*>(src + D);
      *reinterpret_cast<float4*>(&sv[r][c]) = *reinterpret_cast<const float4*>(src + 2 * D);
    }
    __syncthreads();
    for (int j0 = 0; j0 < WIN; j0 += 16) {
      float s[16];
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < DH / 4; ++i) {
          const float4 kk = *reinterpret_cast<const float4*>(&sk[j0 + j][4 * i]);
          d = fmaf(q[4 * i], kk.x, d);
          d = fmaf(q[4 * i + 1], kk.y, d);
          d = fmaf(q[4 * i + 2], kk.z, d);
          d = fmaf(q[4 * i + 3], kk.w, d);
        }
        s[j] = d;
        cm = fmaxf(cm, d);
      }
      const float nm = fmaxf(mx, cm);
      const float corr = expf(mx - nm);          // first chunk: exp(-inf) = 0
      mx = nm;
      den *= corr;
#pragma unroll
      for (int i = 0; i < DH; ++i) acc[i] *= corr;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pj = expf(s[j] - nm);
        den += pj;
#pragma unroll
        for (int i = 0; i < DH / 4; ++i) {
          const float4 vv = *reinterpret_cast<const float4*>(&sv[j0 + j][4 * i]);
          acc[4 * i] = fmaf(pj, vv.x, acc[4 * i]);
          acc[4 * i + 1] = fmaf(pj, vv.y, acc[4 * i + 1]);
          acc[4 * i + 2] = fmaf(pj, vv.z, acc[4 * i + 2]);
          acc[4 * i + 3] = fmaf(pj, vv.w, acc[4 * i + 3]);
        }
      }
    }
  }
  const float inv = 1.0f / den;
  __nv_bfloat16* dst = att2 + row * 2 * D + h * DH;
#pragma unroll
  for (int i = 0; i < DH / 4; ++i)
    store_split4(dst + 4 * i, D, make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv));
}

// Windowed softmax attention of the fp32-class mode on the tensor cores (round 2; the CUDA-core kernel above is kept as the
// unit-test reference).  Same input and output as local_attention_f32_kernel.  Every contraction is a three-product bf16
// split on mma.sync m16n8k16 with fp32 accumulators: x = hi + lo, hi = bf16(x), lo = bf16(x - hi), and
// a.b ~ a_hi.b_hi + a_hi.b_lo + a_lo.b_hi (what the tcgen05 GEMMs of this mode do, Params::split3); the softmax itself
// (scores, maxima, exponentials, sums, the output accumulator) stays fp32.
//   grid (L/128, NL, B), 256 threads, LAF_SMEM_BYTES of dynamic shared memory: one query window; warp w owns query rows 16 w .. 16 w + 15 (Q fragments in registers)
//   per key window (w - 1, w, w + 1 where they exist): K and V converted to (hi, lo) bf16 tiles in shared memory (64-byte
//   rows, 16-byte pieces XOR-swizzled for ldmatrix); per block of 64 keys S = Q K^T (48 MMAs per warp), online softmax,
//   P split in registers (the accumulator fragments of two key tiles are the A fragment of one k-step), O += P V.
constexpr int LAF_SMEM_BYTES = 4 * WIN * 64 + 2 * WIN * DH * 4 + 128;      // four bf16 tiles + raw fp32 K and V rows + alignment slack
__device__ __forceinline__ uint32_t f32_swz(int row, int chunk) { return uint32_t(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

__device__ __forceinline__ void split_pack2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = ptx::pack_bf16x2(x, y);
  lo = ptx::pack_bf16x2(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xffff0000u));
}

__global__ void __launch_bounds__(256, 2)
local_attention_f32_mma_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ att2, int B, int H, int L,
                               float scale) {
  extern __shared__ uint8_t laf_raw[];
  uint8_t* const laf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(laf_raw) + 127) & ~uintptr_t(127));
  uint8_t* const sKh = laf;                         // (hi, lo) bf16 tiles of the current key window, 8 KB each
  uint8_t* const sKl = laf + WIN * 64;
  uint8_t* const sVh = laf + 2 * WIN * 64;
  uint8_t* const sVl = laf + 3 * WIN * 64;
  uint8_t* const sRaw = laf + 4 * WIN * 64;         // fp32 K and V rows of the NEXT key window (cp.async), 16 KB each
  const int w = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * DH;
  const int nw = L / WIN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;
  // Q fragments (A operand, rows g and g + 8 of this warp's 16 rows), scaled like the reference scales q before the dot
  uint32_t qh[2][4], ql[2][4];
  {
    const float* q0 = qkv + (size_t(b) * L + size_t(w) * WIN + warp * 16 + g) * 3 * D + h * DH + 2 * t;
    const float* q1 = q0 + size_t(8) * 3 * D;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const float2 a0 = *reinterpret_cast<const float2*>(q0 + 16 * ks), a1 = *reinterpret_cast<const float2*>(q1 + 16 * ks);
      const float2 a2 = *reinterpret_cast<const float2*>(q0 + 16 * ks + 8), a3 = *reinterpret_cast<const float2*>(q1 + 16 * ks + 8);
      split_pack2(a0.x * scale, a0.y * scale, qh[ks][0], ql[ks][0]);
      split_pack2(a1.x * scale, a1.y * scale, qh[ks][1], ql[ks][1]);
      split_pack2(a2.x * scale, a2.y * scale, qh[ks][2], ql[ks][2]);
      split_pack2(a3.x * scale, a3.y * scale, qh[ks][3], ql[ks][3]);
    }
  }
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mx[2] = {-INFINITY, -INFINITY}, den[2] = {0.f, 0.f};      // rows g, g + 8
  const uint32_t kh0 = ptx::smem_u32(sKh), kl0 = ptx::smem_u32(sKl), vh0 = ptx::smem_u32(sVh), vl0 = ptx::smem_u32(sVl);
  // The fp32 rows of a key window travel global -> shared with cp.async while the previous window is being multiplied;
  // each thread then converts the pieces it fetched itself into the (hi, lo) tiles.
  const int kw_lo = max(w - 1, 0), kw_hi = min(w + 1, nw - 1);
  auto fetch = [&](int kw) {
    for (int i = tid; i < WIN * DH / 4; i += 256) {
      const int r = i >> 3, c4 = i & 7;                             // row, float4 index within the 32 features
      const float* src = qkv + (size_t(b) * L + size_t(kw) * WIN + r) * 3 * D + h * DH + 4 * c4;
      ptx::cp_async_16(ptx::smem_u32(sRaw) + i * 16, src + D);
      ptx::cp_async_16(ptx::smem_u32(sRaw) + WIN * DH * 4 + i * 16, src + 2 * D);
    }
    ptx::cp_async_commit();
  };
  fetch(kw_lo);
  for (int kw = kw_lo; kw <= kw_hi; ++kw) {
    ptx::cp_async_wait<0>();
    __syncthreads();                                                // everyone is done with the previous window's tiles
    for (int i = tid; i < WIN * DH / 4; i += 256) {
      const int r = i >> 3, c4 = i & 7;
      const float4 kk = *reinterpret_cast<const float4*>(sRaw + i * 16);
      const float4 vv = *reinterpret_cast<const float4*>(sRaw + WIN * DH * 4 + i * 16);
      const uint32_t off = f32_swz(r, c4 >> 1) + (c4 & 1) * 8;
      uint32_t h0, l0, h1, l1;
      split_pack2(kk.x, kk.y, h0, l0);
      split_pack2(kk.z, kk.w, h1, l1);
      *reinterpret_cast<uint2*>(sKh + off) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(sKl + off) = make_uint2(l0, l1);
      split_pack2(vv.x, vv.y, h0, l0);
      split_pack2(vv.z, vv.w, h1, l1);
      *reinterpret_cast<uint2*>(sVh + off) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(sVl + off) = make_uint2(l0, l1);
    }
    if (kw < kw_hi) fetch(kw + 1);                                  // this thread's raw pieces are consumed: refill them
    __syncthreads();
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {                                  // two blocks of 64 keys (half the score registers of 128)
      // S = Q K^T: 8 key tiles of 8
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        uint32_t bh[4], bl[4];                                       // b0, b1 of k-step 0, then of k-step 1
        const uint32_t off = f32_swz(64 * hb + 8 * j + (lane & 7), lane >> 3);
        ptx::ldmatrix_x4(kh0 + off, bh[0], bh[1], bh[2], bh[3]);
        ptx::ldmatrix_x4(kl0 + off, bl[0], bl[1], bl[2], bl[3]);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          ptx::mma_bf16_16816(s[j], ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], bh[2 * ks], bh[2 * ks + 1]);
          ptx::mma_bf16_16816(s[j], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bl[2 * ks], bl[2 * ks + 1]);
          ptx::mma_bf16_16816(s[j], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bh[2 * ks], bh[2 * ks + 1]);
        }
      }
      // online softmax over the block's 64 keys; rows g (c0, c1) and g + 8 (c2, c3)
      float corr[2], ml2[2];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) cm = fmaxf(cm, fmaxf(s[j][2 * hf], s[j][2 * hf + 1]));
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
        const float nm = fmaxf(mx[hf], cm);
        corr[hf] = exp2f((mx[hf] - nm) * LOG2E);                     // first block: exp(-inf) = 0
        mx[hf] = nm;
        ml2[hf] = nm * LOG2E;
        den[hf] *= corr[hf];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i][0] *= corr[0]; o[i][1] *= corr[0];
        o[i][2] *= corr[1]; o[i][3] *= corr[1];
      }
      float ps[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = exp2f(fmaf(s[j][0], LOG2E, -ml2[0]));
        s[j][1] = exp2f(fmaf(s[j][1], LOG2E, -ml2[0]));
        s[j][2] = exp2f(fmaf(s[j][2], LOG2E, -ml2[1]));
        s[j][3] = exp2f(fmaf(s[j][3], LOG2E, -ml2[1]));
        ps[0] += s[j][0] + s[j][1];
        ps[1] += s[j][2] + s[j][3];
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float v = ps[hf];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        den[hf] += v;
      }
      // O += P V: key k-steps of 16 = two key tiles; V fragments through ldmatrix.trans (V is stored key-major)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t ph[4], pl[4];
        split_pack2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
        split_pack2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
        split_pack2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
        split_pack2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
        for (int ep = 0; ep < 2; ++ep) {
          uint32_t vh[4], vl[4];
          const uint32_t off = f32_swz(64 * hb + kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * ep + (lane >> 4));
          ptx::ldmatrix_x4_trans(vh0 + off, vh[0], vh[1], vh[2], vh[3]);
          ptx::ldmatrix_x4_trans(vl0 + off, vl[0], vl[1], vl[2], vl[3]);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            ptx::mma_bf16_16816(o[2 * ep + q], pl[0], pl[1], pl[2], pl[3], vh[2 * q], vh[2 * q + 1]);
            ptx::mma_bf16_16816(o[2 * ep + q], ph[0], ph[1], ph[2], ph[3], vl[2 * q], vl[2 * q + 1]);
            ptx::mma_bf16_16816(o[2 * ep + q], ph[0], ph[1], ph[2], ph[3], vh[2 * q], vh[2 * q + 1]);
          }
        }
      }
    }
  }
  const float inv0 = 1.0f / den[0], inv1 = 1.0f / den[1];
  const size_t row0 = size_t(b) * L + size_t(w) * WIN + warp * 16 + g;
  __nv_bfloat16* d0 = att2 + row0 * 2 * D + h * DH + 2 * t;
  __nv_bfloat16* d1 = d0 + size_t(8) * 2 * D;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    uint32_t hi, lo;
    split_pack2(o[nt][0] * inv0, o[nt][1] * inv0, hi, lo);
    *reinterpret_cast<uint32_t*>(d0 + 8 * nt) = hi;
    *reinterpret_cast<uint32_t*>(d0 + 8 * nt + D) = lo;
    split_pack2(o[nt][2] * inv1, o[nt][3] * inv1, hi, lo);
    *reinterpret_cast<uint32_t*>(d1 + 8 * nt) = hi;
    *reinterpret_cast<uint32_t*>(d1 + 8 * nt + D) = lo;
  }
}

// The same attention on tcgen05 (round 2, default): mma.sync runs at about an eighth of the tcgen05 rate on this part and
// bounded the kernel above (19 M m16n8k16 per layer).  One CTA of 128 threads per query window, thread = query row = TMEM
// lane; up to four CTAs per SM overlap each other's phases (128 TMEM columns, 48 KB of shared memory each):
//   * the CTA converts the Q rows (scaled like the reference scales q) and, per key window, the K and V rows from fp32
//     (coalesced float4 loads, 8 threads per row) to (hi, lo) bf16 rows of shared-memory tiles in the 64-byte-swizzled layout the MMA descriptors expect
//     (what TMA writes in the bf16 kernel, attention.cuh); fence.proxy.async publishes them to the tensor core
//   * one elected thread issues, per block of 64 keys, S = Ql Kh^T + Qh Kl^T + Qh Kh^T (six 128 x 64 x 16 MMAs into 64 TMEM
//     columns) and, after the softmax, O += Pl Vh + Ph Vl + Ph Vh (twelve 128 x 32 x 16 MMAs, P read from TMEM, V as
//     stored, MN-major); tcgen05 operations of one thread execute in issue order, which orders S of the next block
//     behind the P V reads of this one
//   * softmax per row in fp32 against a lazily rescaled reference maximum (as in the bf16 kernel); P is split into bf16
//     hi (TMEM columns 0-31 of the S slot) and lo (columns 32-63): the split overwrites the scores in place
#ifndef BIOM3_F32_TC_CTAS
#define BIOM3_F32_TC_CTAS 3      // CTAs per SM the register budget is cut for (4 measured equal: the kernel is bound by its serial phases, not by occupancy)
#endif
constexpr int LAT_SMEM_BYTES = 6 * WIN * 64 + 1024;     // Qh, Ql, Kh, Kl, Vh, Vl tiles + alignment slack

// fp32 rows [128][32] at `src` (row stride `ld` floats) -> (hi, lo) bf16 tiles, 128 threads: thread i + 128 j takes the j-th
// float4 round, 8 threads per row, so every warp load covers four full 128-byte rows
__device__ __forceinline__ void split_tile_128(const float* __restrict__ src, size_t ld, float mul, uint8_t* th, uint8_t* tl, int tid) {
  float4 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = tid + 128 * j;
    v[j] = *reinterpret_cast<const float4*>(src + size_t(i >> 3) * ld + 4 * (i & 7));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = tid + 128 * j, r = i >> 3, c4 = i & 7;
    uint32_t h0, l0, h1, l1;
    split_pack2(v[j].x * mul, v[j].y * mul, h0, l0);
    split_pack2(v[j].z * mul, v[j].w * mul, h1, l1);
    const uint32_t off = f32_swz(r, c4 >> 1) + (c4 & 1) * 8;
    *reinterpret_cast<uint2*>(th + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(tl + off) = make_uint2(l0, l1);
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(128, BIOM3_F32_TC_CTAS)
local_attention_f32_tc_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ att2, int B, int H, int L, float scale) {
  extern __shared__ uint8_t lat_raw[];
  uint8_t* const sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lat_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* const sQh = sm;
  uint8_t* const sQl = sm + WIN * 64;
  uint8_t* const sKh = sm + 2 * WIN * 64;
  uint8_t* const sKl = sm + 3 * WIN * 64;
  uint8_t* const sVh = sm + 4 * WIN * 64;
  uint8_t* const sVl = sm + 5 * WIN * 64;
  __shared__ uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  const int w = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * DH;
  const int nw = L / WIN;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr float LAZY_LOG2 = 8.0f;
  constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  constexpr uint32_t IDESC_O = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    ptx::mbar_init(&bar_s, 1);
    ptx::mbar_init(&bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_s = tmem + ((uint32_t(warp) * 32u) << 16);          // this warp's lanes: S / P at columns 0-63
  const uint32_t t_o = t_s + 64;                                          // O at columns 64-95
  const size_t qrow = size_t(b) * L + size_t(w) * WIN + tid;
  split_tile_128(qkv + (size_t(b) * L + size_t(w) * WIN) * 3 * D + h * DH, size_t(3) * D, scale, sQh, sQl, tid);
  const uint32_t qh_lo = attn::local_desc_lo(ptx::smem_u32(sQh)), ql_lo = attn::local_desc_lo(ptx::smem_u32(sQl));
  const uint32_t kh_lo = attn::local_desc_lo(ptx::smem_u32(sKh)), kl_lo = attn::local_desc_lo(ptx::smem_u32(sKl));
  const uint32_t vh_lo = attn::local_desc_lo(ptx::smem_u32(sVh)), vl_lo = attn::local_desc_lo(ptx::smem_u32(sVl));
  auto issue_s = [&](int sb) {                         // one thread: S = Ql Kh^T + Qh Kl^T + Qh Kh^T for keys 64 sb ..
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ptx::umma_bf16(tmem, attn::local_desc(ql_lo + 2 * ks), attn::local_desc(kh_lo + sb * 256 + 2 * ks), IDESC_S, ks != 0);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ptx::umma_bf16(tmem, attn::local_desc(qh_lo + 2 * ks), attn::local_desc(kl_lo + sb * 256 + 2 * ks), IDESC_S, 1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ptx::umma_bf16(tmem, attn::local_desc(qh_lo + 2 * ks), attn::local_desc(kh_lo + sb * 256 + 2 * ks), IDESC_S, 1);
    ptx::umma_commit(&bar_s);
  };
  float m_ref = 0.f, rs = 0.f;
  uint32_t ph_s = 0, ph_o = 0;
  int blocks_done = 0;                                 // P V blocks issued so far for this item
  const int kw_lo = max(w - 1, 0), kw_hi = min(w + 1, nw - 1);
  for (int kw = kw_lo; kw <= kw_hi; ++kw) {
    // the previous window's MMAs must be done with the K / V tiles before they are rewritten
    if (blocks_done > 0) {
      ptx::mbar_wait(&bar_o, ph_o ^ 1);                // the last commit's phase (ph_o was flipped when it was issued)
      ptx::tc_fence_after();
    }
    const float* kbase = qkv + (size_t(b) * L + size_t(kw) * WIN) * 3 * D + D + h * DH;
    split_tile_128(kbase, size_t(3) * D, 1.0f, sKh, sKl, tid);
    split_tile_128(kbase + D, size_t(3) * D, 1.0f, sVh, sVl, tid);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      ptx::tc_fence_after();
      if (ptx::elect_one()) issue_s(0);
      __syncwarp();
    }
    for (int sb = 0; sb < 2; ++sb) {
      ptx::mbar_wait(&bar_s, ph_s);
      ph_s ^= 1;
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32(t_s, r0);
      ptx::tmem_ld_32x32(t_s + 32, r1);
      ptx::tmem_ld_wait();
      float bm = -INFINITY;
#pragma unroll
      for (int k = 0; k < 32; ++k) bm = fmaxf(bm, fmaxf(__uint_as_float(r0[k]), __uint_as_float(r1[k])));
      if (blocks_done == 0) {
        m_ref = bm;
        rs = 0.f;
      } else {
        const bool need = (bm - m_ref) * LOG2E > LAZY_LOG2;
        if (__any_sync(0xffffffffu, need)) {
          // every P V issued so far must have landed in O
          ptx::mbar_wait(&bar_o, ph_o ^ 1);
          ptx::tc_fence_after();
          const float f = need ? exp2f((m_ref - bm) * LOG2E) : 1.f;
          uint32_t ro[32];
          ptx::tmem_ld_32x32(t_o, ro);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) ro[k] = __float_as_uint(__uint_as_float(ro[k]) * f);
          ptx::tmem_st_32x32(t_o, ro);
          ptx::tmem_st_wait();
          rs *= f;
          if (need) m_ref = bm;
        }
      }
      const float ml = m_ref * LOG2E;
      uint32_t ph[32], pl[32];                         // P hi / lo as bf16 pairs: 64 keys -> 32 + 32 columns
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * k]), LOG2E, -ml)), p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * k + 1]), LOG2E, -ml));
        const float p2 = ex2_approx(fmaf(__uint_as_float(r1[2 * k]), LOG2E, -ml)), p3 = ex2_approx(fmaf(__uint_as_float(r1[2 * k + 1]), LOG2E, -ml));
        sum += (p0 + p1) + (p2 + p3);
        split_pack2(p0, p1, ph[k], pl[k]);
        split_pack2(p2, p3, ph[16 + k], pl[16 + k]);
      }
      rs += sum;
      ptx::tmem_st_32x32(t_s, ph);
      ptx::tmem_st_32x32(t_s + 32, pl);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // O (+)= Pl Vh + Ph Vl + Ph Vh over the block's 64 keys (four k-steps of 16)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16_ts(tmem + 64, tmem + 32 + ks * 8, attn::local_desc(vh_lo + sb * 256 + ks * 64), IDESC_O, (blocks_done | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16_ts(tmem + 64, tmem + ks * 8, attn::local_desc(vl_lo + sb * 256 + ks * 64), IDESC_O, 1);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16_ts(tmem + 64, tmem + ks * 8, attn::local_desc(vh_lo + sb * 256 + ks * 64), IDESC_O, 1);
          ptx::umma_commit(&bar_o);
          if (sb == 0) issue_s(1);                     // in issue order behind the P V reads of the S / P columns
        }
        __syncwarp();
      }
      ph_o ^= 1;
      ++blocks_done;
    }
  }
  // output: O / row sum, split, this thread's row
  ptx::mbar_wait(&bar_o, ph_o ^ 1);
  ptx::tc_fence_after();
  {
    uint32_t ro[32];
    ptx::tmem_ld_32x32(t_o, ro);
    ptx::tmem_ld_wait();
    const float inv = 1.0f / rs;
    __nv_bfloat16* dst = att2 + qrow * 2 * D + h * DH;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        split_pack2(__uint_as_float(ro[8 * c + 2 * i]) * inv, __uint_as_float(ro[8 * c + 2 * i + 1]) * inv, hw[i], lw[i]);
      *reinterpret_cast<uint4*>(dst + 8 * c) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(dst + D + 8 * c) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 128);
}

// Linear attention of the fp32-class mode on the tensor cores (round 2; the CUDA-core kernel below is kept as the unit-test
// reference).  Same input and output as linear_attention_f32_kernel, grid (H - NL, B), 256 threads, LINF_SMEM_BYTES of
// dynamic shared memory.  The softmaxes are fp32 and thread-per-token (each thread owns one token's 32 features); the two
// contractions (exp(k)^T v over the tokens, softmax(q) ctx over the features) are three-product bf16 splits on mma.sync
// with fp32 accumulators, like local_attention_f32_mma_kernel.
//   pass 1   per-feature maximum of k over the sequence (lane = feature, warps stride over the tokens)
//   pass 2   256 tokens at a time: thread = token computes exp(k - max) and splits it and v into (hi, lo) bf16 rows of its
//            warp's 32-row slice of the tiles; the warp then multiplies its own 32 tokens in (only __syncwarp between the
//            two), denominators through the same A fragments against a ones operand; partial ctx per warp -> reduced in
//            shared memory -> ctx^T / Z * q_scale as (hi, lo) bf16
//   pass 3   thread = token softmaxes its q row, splits it into the warp's tile slice; the warp multiplies by ctx
constexpr int LINF_TOK = 256;
constexpr int LINF_SMEM_BYTES = 4 * LINF_TOK * 64 + 128;
constexpr uint32_t ONES_BF16X2 = 0x3F803F80u;

__global__ void __launch_bounds__(256)
linear_attention_f32_mma_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ att2, int B, int H, int L,
                                int NL, float q_scale) {
  extern __shared__ uint8_t linf_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(linf_raw) + 127) & ~uintptr_t(127));
  uint8_t* sEh = tiles;                           // exp(k) (pass 2) / softmax(q) (pass 3), hi; [256 tokens][64 bytes], swizzled
  uint8_t* sEl = tiles + LINF_TOK * 64;
  uint8_t* sVh = tiles + 2 * LINF_TOK * 64;
  uint8_t* sVl = tiles + 3 * LINF_TOK * 64;
  __shared__ float red[8][DH];
  __shared__ float kmax[DH], zsum[DH];
  __shared__ __align__(128) uint8_t sCh[DH * 64], sCl[DH * 64];   // ctx^T [e][d] bf16 (hi, lo), swizzled rows
  const int h = NL + blockIdx.x, b = blockIdx.y;
  const int D = H * DH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;
  const float* base = qkv + size_t(b) * L * 3 * D + h * DH;       // + n*3D (+ D for k, + 2D for v)
  // pass 1
  {
    float m = -INFINITY;
    for (int n = warp; n < L; n += 8) m = fmaxf(m, base[size_t(n) * 3 * D + D + lane]);
    red[warp][lane] = m;
    __syncthreads();
    if (warp == 0) {
      float mm = red[0][lane];
#pragma unroll
      for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i][lane]);
      kmax[lane] = mm * LOG2E;
    }
    __syncthreads();
  }
  const int r0 = warp * 32;                       // this warp's rows of the tiles
  const uint32_t eh0 = ptx::smem_u32(sEh), el0 = ptx::smem_u32(sEl), vh0 = ptx::smem_u32(sVh), vl0 = ptx::smem_u32(sVl);
  // pass 2
  float acc[2][4][4], accd[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      accd[i][j] = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
    }
  for (int n0 = 0; n0 < L; n0 += LINF_TOK) {
    const int n = n0 + tid;
    const bool live = n < L;
    const float* row = base + size_t(live ? n : 0) * 3 * D;
#pragma unroll
    for (int c = 0; c < 4; ++c) {                 // 8 features per step: one 16-byte piece of each bf16 row
      float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, v0 = k0, v1 = k0;
      float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (live) {
        k0 = *reinterpret_cast<const float4*>(row + D + 8 * c);
        k1 = *reinterpret_cast<const float4*>(row + D + 8 * c + 4);
        v0 = *reinterpret_cast<const float4*>(row + 2 * D + 8 * c);
        v1 = *reinterpret_cast<const float4*>(row + 2 * D + 8 * c + 4);
        const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = exp2f(fmaf(kk[i], LOG2E, -kmax[8 * c + i]));
      }
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_pack2(e[2 * i], e[2 * i + 1], hw[i], lw[i]);
      const uint32_t off = f32_swz(tid, c);
      *reinterpret_cast<uint4*>(sEh + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(sEl + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      split_pack2(v0.x, v0.y, hw[0], lw[0]);
      split_pack2(v0.z, v0.w, hw[1], lw[1]);
      split_pack2(v1.x, v1.y, hw[2], lw[2]);
      split_pack2(v1.z, v1.w, hw[3], lw[3]);
      *reinterpret_cast<uint4*>(sVh + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(sVl + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    __syncwarp();
    uint32_t ah[2][2][4], al[2][2][4];             // [k-step][m-tile(d)][a0..a3]: exp(k)^T fragments
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t off = f32_swz(r0 + ks * 16 + (lane & 7) + 8 * (lane >> 4), 2 * mt + ((lane >> 3) & 1));
        ptx::ldmatrix_x4_trans(eh0 + off, ah[ks][mt][0], ah[ks][mt][1], ah[ks][mt][2], ah[ks][mt][3]);
        ptx::ldmatrix_x4_trans(el0 + off, al[ks][mt][0], al[ks][mt][1], al[ks][mt][2], al[ks][mt][3]);
      }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int ep = 0; ep < 2; ++ep) {
        uint32_t vh[4], vl[4];
        const uint32_t off = f32_swz(r0 + ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * ep + (lane >> 4));
        ptx::ldmatrix_x4_trans(vh0 + off, vh[0], vh[1], vh[2], vh[3]);
        ptx::ldmatrix_x4_trans(vl0 + off, vl[0], vl[1], vl[2], vl[3]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            ptx::mma_bf16_16816(acc[mt][2 * ep + q], al[ks][mt][0], al[ks][mt][1], al[ks][mt][2], al[ks][mt][3], vh[2 * q], vh[2 * q + 1]);
            ptx::mma_bf16_16816(acc[mt][2 * ep + q], ah[ks][mt][0], ah[ks][mt][1], ah[ks][mt][2], ah[ks][mt][3], vl[2 * q], vl[2 * q + 1]);
            ptx::mma_bf16_16816(acc[mt][2 * ep + q], ah[ks][mt][0], ah[ks][mt][1], ah[ks][mt][2], ah[ks][mt][3], vh[2 * q], vh[2 * q + 1]);
          }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        ptx::mma_bf16_16816(accd[mt], al[ks][mt][0], al[ks][mt][1], al[ks][mt][2], al[ks][mt][3], ONES_BF16X2, ONES_BF16X2);
        ptx::mma_bf16_16816(accd[mt], ah[ks][mt][0], ah[ks][mt][1], ah[ks][mt][2], ah[ks][mt][3], ONES_BF16X2, ONES_BF16X2);
      }
    }
    __syncwarp();
  }
  // per-warp partials -> shared (aliasing the tiles), reduced over the 8 warps
  __syncthreads();
  float* part = reinterpret_cast<float*>(tiles);                 // [8][32 d][32 e]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int d = 16 * mt + 8 * hf + g;
      if (t == 0) red[warp][d] = accd[mt][2 * hf];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float* dst = part + (size_t(warp) * DH + d) * DH + 8 * nt + 2 * t;
        dst[0] = acc[mt][nt][2 * hf];
        dst[1] = acc[mt][nt][2 * hf + 1];
      }
    }
  __syncthreads();
  if (tid < DH) {
    float z = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) z += red[i][tid];
    zsum[tid] = q_scale / z;
  }
  __syncthreads();
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5, e = i & 31;
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sum += part[(size_t(wv) * DH + d) * DH + e];
    sum *= zsum[d];
    const __nv_bfloat16 hi = __float2bfloat16_rn(sum);
    const uint32_t off = f32_swz(e, d >> 3) + (d & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sCh + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(sCl + off) = __float2bfloat16_rn(sum - __bfloat162float(hi));
  }
  __syncthreads();
  // pass 3
  uint32_t ch[4][4], cl[4][4];                     // ctx as B fragments: [n-tile(e)][b0 ks0, b1 ks0, b0 ks1, b1 ks1]
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const uint32_t off = f32_swz(8 * nt + (lane & 7), lane >> 3);
    ptx::ldmatrix_x4(ptx::smem_u32(sCh) + off, ch[nt][0], ch[nt][1], ch[nt][2], ch[nt][3]);
    ptx::ldmatrix_x4(ptx::smem_u32(sCl) + off, cl[nt][0], cl[nt][1], cl[nt][2], cl[nt][3]);
  }
  for (int n0 = 0; n0 < L; n0 += LINF_TOK) {
    const int n = n0 + tid;
    const bool live = n < L;
    const float* row = base + size_t(live ? n : 0) * 3 * D;
    float qv[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 x = live ? *reinterpret_cast<const float4*>(row + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      qv[4 * c] = x.x; qv[4 * c + 1] = x.y; qv[4 * c + 2] = x.z; qv[4 * c + 3] = x.w;
    }
    float qm = qv[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) qm = fmaxf(qm, qv[i]);
    float qs = 0.f;
    const float qml = qm * LOG2E;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      qv[i] = exp2f(fmaf(qv[i], LOG2E, -qml));
      qs += qv[i];
    }
    const float qi = 1.0f / qs;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_pack2(qv[8 * c + 2 * i] * qi, qv[8 * c + 2 * i + 1] * qi, hw[i], lw[i]);
      const uint32_t off = f32_swz(tid, c);
      *reinterpret_cast<uint4*>(sEh + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(sEl + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {               // 16 tokens each
      uint32_t qh[2][4], ql[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t off = f32_swz(r0 + mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * ks + (lane >> 4));
        ptx::ldmatrix_x4(eh0 + off, qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
        ptx::ldmatrix_x4(el0 + off, ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
      }
      float o[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          ptx::mma_bf16_16816(o[nt], ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], ch[nt][2 * ks], ch[nt][2 * ks + 1]);
          ptx::mma_bf16_16816(o[nt], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], cl[nt][2 * ks], cl[nt][2 * ks + 1]);
          ptx::mma_bf16_16816(o[nt], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], ch[nt][2 * ks], ch[nt][2 * ks + 1]);
        }
      }
      const int tok0 = n0 + r0 + mt * 16 + g;
      __nv_bfloat16* d0 = att2 + (size_t(b) * L + tok0) * 2 * D + h * DH + 2 * t;
      __nv_bfloat16* d1 = d0 + size_t(8) * 2 * D;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t hi, lo;
        if (tok0 < L) {
          split_pack2(o[nt][0], o[nt][1], hi, lo);
          *reinterpret_cast<uint32_t*>(d0 + 8 * nt) = hi;
          *reinterpret_cast<uint32_t*>(d0 + 8 * nt + D) = lo;
        }
        if (tok0 + 8 < L) {
          split_pack2(o[nt][2], o[nt][3], hi, lo);
          *reinterpret_cast<uint32_t*>(d1 + 8 * nt) = hi;
          *reinterpret_cast<uint32_t*>(d1 + 8 * nt + D) = lo;
        }
      }
    }
    __syncwarp();
  }
}

// Linear attention, fp32, heads [NL, H): q = softmax_d(q) * dh^-0.5, k = softmax_n(k), ctx = k^T v, out = q ctx.
// grid (H - NL, B), 256 threads.  Three passes over the head's [L][32] k / v / q columns (L2 resident).
__global__ void __launch_bounds__(256)
linear_attention_f32_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ att2, int B, int H, int L,
                            int NL, float q_scale) {
  __shared__ float red[8][DH];
  __shared__ float kmax[DH], kinv[DH];
  __shared__ float ctx[DH][DH + 1];
  __shared__ float part[8][DH][DH + 1];        // per-warp partial ctx (33 KB)
  __shared__ float zpart[8][DH];
  const int h = NL + blockIdx.x, b = blockIdx.y;
  const int D = H * DH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = qkv + size_t(b) * L * 3 * D + h * DH;       // + n*3D (+ D for k, + 2D for v)
  // pass 1: per-feature max of k over the sequence (lane = feature, warps stride over tokens)
  float m = -INFINITY;
  for (int n = warp; n < L; n += 8) m = fmaxf(m, base[size_t(n) * 3 * D + D + lane]);
  red[warp][lane] = m;
  __syncthreads();
  if (warp == 0) {
    float mm = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i][lane]);
    kmax[lane] = mm;
  }
  __syncthreads();
  // pass 2: ctx[d][e] = sum_n exp(k[n][d] - max_d) v[n][e], Z[d] = sum_n exp(.)   (lane = e; warp strides tokens)
  float c[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) c[d] = 0.f;
  float z = 0.f;                                // lane = d for the Z sum
  for (int n = warp; n < L; n += 8) {
    const float kv = expf(base[size_t(n) * 3 * D + D + lane] - kmax[lane]);   // lane = d
    const float vv = base[size_t(n) * 3 * D + 2 * D + lane];                  // lane = e
    z += kv;
#pragma unroll
    for (int d = 0; d < DH; ++d) c[d] = fmaf(__shfl_sync(0xffffffffu, kv, d), vv, c[d]);
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) part[warp][d][lane] = c[d];
  zpart[warp][lane] = z;
  __syncthreads();
  if (warp == 0) {
    float zz = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) zz += zpart[i][lane];
    kinv[lane] = 1.0f / zz;
  }
  __syncthreads();
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i / DH, e = i % DH;
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sum += part[wv][d][e];
    ctx[d][e] = sum * kinv[d];
  }
  __syncthreads();
  // pass 3: out[n][e] = sum_d softmax_d(q[n])[d] * q_scale * ctx[d][e]   (lane = d for the softmax, = e for the output)
  for (int n = warp; n < L; n += 8) {
    const float qv = base[size_t(n) * 3 * D + lane];
    float qm = qv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, o));
    const float qe = expf(qv - qm);
    const float qs = qe / warp_sum_f(qe) * q_scale;
    float o = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) o = fmaf(__shfl_sync(0xffffffffu, qs, d), ctx[d][lane], o);
    __nv_bfloat16 hi, lo;
    split_bf16(o, hi, lo);
    __nv_bfloat16* dst = att2 + (size_t(b) * L + n) * 2 * D + h * DH + lane;
    dst[0] = hi;
    dst[D] = lo;
  }
}

}  // namespace f32p
