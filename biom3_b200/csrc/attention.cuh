// Attention kernels for the ProteoScribe block (heads [0,NL) windowed softmax, heads [NL,H) linear).
// Semantics restated from the un-vendored LocalAttention / linear_attn blocks the reference calls through
// /root/reference/Stage3_source/cond_diff_transformer_layer.py:123-143,171.
//
// Input layout (written by the QKV GEMM epilogue): qkv bf16 [3][B][H][L][32]; output bf16 [B*L][H*32].
#pragma once
#include "ptx.cuh"

namespace attn {

constexpr int DH = 32;     // head dim (fixed by the kernels; checked at create time)
constexpr int WIN = 128;   // local window (fixed; checked at create time)

// ------------------------------------------------------------------------------------------------
// Local attention: query window w attends to key windows w-1, w, w+1 (those that exist), softmax
// over the real keys only (masked keys get exp() == 0 in the reference, i.e. they are skipped here).
// One CTA per (window, local head, sample); 8 warps x 16 query rows; mma.sync m16n8k16 bf16.
// smem rows are 64 B (32 bf16); the 16-byte chunk index is XOR-swizzled with (row >> 1) & 3 so that
// ldmatrix (8 rows x 16 B) is bank-conflict free.
// ------------------------------------------------------------------------------------------------
constexpr int LOCAL_SMEM_BYTES = 7 * WIN * 64;   // Q (1 window) + K (3) + V (3), 64 B per row

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return uint32_t(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

__global__ void __launch_bounds__(256)
local_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int H, int L,
                       float scale_log2e) {
  const int w = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nw = L / WIN;
  const int w_lo = max(w - 1, 0), w_hi = min(w + 1, nw - 1);
  const int nkeys = (w_hi - w_lo + 1) * WIN;
  const size_t head_stride = size_t(L) * DH;
  const size_t plane = size_t(B) * H * head_stride;
  const __nv_bfloat16* qg = qkv + (size_t(b) * H + h) * head_stride + size_t(w) * WIN * DH;
  const __nv_bfloat16* kg = qkv + plane + (size_t(b) * H + h) * head_stride + size_t(w_lo) * WIN * DH;
  const __nv_bfloat16* vg = kg + plane;

  extern __shared__ __align__(128) uint8_t attn_smem[];     // LOCAL_SMEM_BYTES, opt-in dynamic
  uint8_t* sQ = attn_smem;
  uint8_t* sK = sQ + WIN * 64;
  uint8_t* sV = sK + 3 * WIN * 64;

  const int tid = threadIdx.x;
  const uint32_t sq = ptx::smem_u32(sQ), sk = ptx::smem_u32(sK), sv = ptx::smem_u32(sV);
  for (int i = tid; i < WIN * 4; i += 256) {
    const int row = i >> 2, ch = i & 3;
    ptx::cp_async_16(sq + swz(row, ch), qg + row * DH + ch * 8);
  }
  for (int i = tid; i < nkeys * 4; i += 256) {
    const int row = i >> 2, ch = i & 3;
    ptx::cp_async_16(sk + swz(row, ch), kg + row * DH + ch * 8);
    ptx::cp_async_16(sv + swz(row, ch), vg + row * DH + ch * 8);
  }
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = warp * 16;

  // Q fragments for the two k-steps (d 0..15, 16..31)
  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int row = q0 + (lane & 7) + 8 * ((lane >> 3) & 1);
    const int ch = ks * 2 + (lane >> 4);
    ptx::ldmatrix_x4(sq + swz(row, ch), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
  }

  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kc = 0; kc < nkeys; kc += 64) {
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kb0, kb1, kb2, kb3;
      ptx::ldmatrix_x4(sk + swz(kc + nt * 8 + (lane & 7), lane >> 3), kb0, kb1, kb2, kb3);
      ptx::mma_bf16_16816(s[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kb0, kb1);
      ptx::mma_bf16_16816(s[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kb2, kb3);
    }
    float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      cm0 = fmaxf(cm0, fmaxf(s[nt][0], s[nt][1]));
      cm1 = fmaxf(cm1, fmaxf(s[nt][2], s[nt][3]));
    }
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
    const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
    const float corr0 = exp2f((m0 - mn0) * scale_log2e), corr1 = exp2f((m1 - mn1) * scale_log2e);
    m0 = mn0; m1 = mn1;
    const float ms0 = mn0 * scale_log2e, ms1 = mn1 * scale_log2e;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(fmaf(s[nt][0], scale_log2e, -ms0));
      s[nt][1] = exp2f(fmaf(s[nt][1], scale_log2e, -ms0));
      s[nt][2] = exp2f(fmaf(s[nt][2], scale_log2e, -ms1));
      s[nt][3] = exp2f(fmaf(s[nt][3], scale_log2e, -ms1));
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * corr0 + rs0;
    l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      o[dt][0] *= corr0; o[dt][1] *= corr0;
      o[dt][2] *= corr1; o[dt][3] *= corr1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {        // 16 keys per step
      const uint32_t a0 = ptx::pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      const uint32_t a1 = ptx::pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = ptx::pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      const uint32_t a3 = ptx::pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const int krow = kc + kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {      // two d-chunks (16 features) per ldmatrix.x4
        uint32_t v0, v1, v2, v3;
        ptx::ldmatrix_x4_trans(sv + swz(krow, dp * 2 + (lane >> 4)), v0, v1, v2, v3);
        ptx::mma_bf16_16816(o[dp * 2], a0, a1, a2, a3, v0, v1);
        ptx::mma_bf16_16816(o[dp * 2 + 1], a0, a1, a2, a3, v2, v3);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int D = H * DH;
  const size_t row0 = size_t(b) * L + size_t(w) * WIN + q0 + g;
  __nv_bfloat16* o0 = out + row0 * D + h * DH + 2 * t;
  __nv_bfloat16* o1 = o0 + size_t(8) * D;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    *reinterpret_cast<uint32_t*>(o0 + dt * 8) = ptx::pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
    *reinterpret_cast<uint32_t*>(o1 + dt * 8) = ptx::pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------------
// Linear attention (global heads): q <- softmax_d(q) * dh^-0.5 ; k <- softmax over the L tokens ;
// ctx = k^T v (32 x 32) ; out = q ctx.  One CTA per (global head, sample), 256 threads, fp32 math.
// ------------------------------------------------------------------------------------------------
constexpr int LIN_CHUNK = 128;

__global__ void __launch_bounds__(256)
linear_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int H, int L,
                        int NL, float q_scale) {
  const int h = NL + blockIdx.x, b = blockIdx.y;
  const size_t head_stride = size_t(L) * DH;
  const size_t plane = size_t(B) * H * head_stride;
  const __nv_bfloat16* qg = qkv + (size_t(b) * H + h) * head_stride;
  const __nv_bfloat16* kg = qg + plane;
  const __nv_bfloat16* vg = kg + plane;

  __shared__ float s_red[8][DH];
  __shared__ float s_max[DH];
  __shared__ float s_e[LIN_CHUNK][DH + 1];
  __shared__ __align__(16) float s_v[LIN_CHUNK][DH];
  __shared__ __align__(16) float s_ctx[DH][DH];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // pass 1: per-feature max of k over the sequence (lane = feature d)
  float mx = -INFINITY;
  for (int n = warp; n < L; n += 8) mx = fmaxf(mx, __bfloat162float(kg[size_t(n) * DH + lane]));
  s_red[warp][lane] = mx;
  __syncthreads();
  if (warp == 0) {
    float m = s_red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i][lane]);
    s_max[lane] = m;
  }
  __syncthreads();

  // pass 2: ctx_un[d][e] = sum_n exp(k[n][d] - max[d]) v[n][e];  den[d] = sum_n exp(...)
  const int d = tid >> 3, e0 = (tid & 7) * 4;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f, den = 0.f;
  for (int c0 = 0; c0 < L; c0 += LIN_CHUNK) {
    for (int i = tid; i < LIN_CHUNK * DH; i += 256) {
      const int r = i >> 5, c = i & 31;
      s_e[r][c] = __expf(__bfloat162float(kg[size_t(c0 + r) * DH + c]) - s_max[c]);
      s_v[r][c] = __bfloat162float(vg[size_t(c0 + r) * DH + c]);
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < LIN_CHUNK; ++r) {
      const float ev = s_e[r][d];
      const float4 vv = *reinterpret_cast<const float4*>(&s_v[r][e0]);
      acc0 = fmaf(ev, vv.x, acc0);
      acc1 = fmaf(ev, vv.y, acc1);
      acc2 = fmaf(ev, vv.z, acc2);
      acc3 = fmaf(ev, vv.w, acc3);
      den += ev;
    }
    __syncthreads();
  }
  const float inv = q_scale / den;       // fold the q scale (dh^-0.5) into ctx
  s_ctx[d][e0] = acc0 * inv;
  s_ctx[d][e0 + 1] = acc1 * inv;
  s_ctx[d][e0 + 2] = acc2 * inv;
  s_ctx[d][e0 + 3] = acc3 * inv;
  __syncthreads();

  // pass 3: out[n][:] = softmax_d(q[n][:]) . ctx   (one token per thread per iteration)
  const int D = H * DH;
  for (int n = tid; n < L; n += 256) {
    float qv[DH];
    const uint4* q4 = reinterpret_cast<const uint4*>(qg + size_t(n) * DH);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(q4 + i);
      const uint32_t w32[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(&w32[j]);
        qv[i * 8 + j * 2] = __bfloat162float(bb.x);
        qv[i * 8 + j * 2 + 1] = __bfloat162float(bb.y);
      }
    }
    float qm = qv[0];
#pragma unroll
    for (int i = 1; i < DH; ++i) qm = fmaxf(qm, qv[i]);
    float qs = 0.f;
#pragma unroll
    for (int i = 0; i < DH; ++i) {
      qv[i] = __expf(qv[i] - qm);
      qs += qv[i];
    }
    const float qinv = 1.f / qs;
    float ov[DH];
#pragma unroll
    for (int e = 0; e < DH; ++e) ov[e] = 0.f;
#pragma unroll
    for (int dd = 0; dd < DH; ++dd) {
      const float pq = qv[dd] * qinv;
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        const float4 cv = *reinterpret_cast<const float4*>(&s_ctx[dd][e]);
        ov[e] = fmaf(pq, cv.x, ov[e]);
        ov[e + 1] = fmaf(pq, cv.y, ov[e + 1]);
        ov[e + 2] = fmaf(pq, cv.z, ov[e + 2]);
        ov[e + 3] = fmaf(pq, cv.w, ov[e + 3]);
      }
    }
    uint4* o4 = reinterpret_cast<uint4*>(out + (size_t(b) * L + n) * D + h * DH);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o4[i] = make_uint4(ptx::pack_bf16x2(ov[8 * i], ov[8 * i + 1]), ptx::pack_bf16x2(ov[8 * i + 2], ov[8 * i + 3]),
                         ptx::pack_bf16x2(ov[8 * i + 4], ov[8 * i + 5]), ptx::pack_bf16x2(ov[8 * i + 6], ov[8 * i + 7]));
  }
}

}  // namespace attn
