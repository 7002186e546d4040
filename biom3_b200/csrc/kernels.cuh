// Bandwidth-bound kernels of the ProteoScribe step: embedding + conditioning + LayerNorm, the vocab
// head, the softmax / renormalise / Exp(1)-race categorical draw, the unmask scatter, and the small
// fp32 GEMM used once per model load / per prompt for the conditioning MLPs.
//
// Reference semantics:
//   embed + axial pos + per-layer additive conditioning  /root/reference/Stage3_source/cond_diff_transformer_layer.py:152-171
//   final LayerNorm + vocab head + permute                /root/reference/Stage3_source/cond_diff_transformer_layer.py:173-176
//   softmax(dim=1) + OneHotCategorical                    /root/reference/Stage3_source/transformer_training_helper.py:443-449
//   sample every position, argmax, unmask write           /root/reference/Stage3_source/sampling_analysis.py:251-256
#pragma once
#include <climits>
#include "ptx.cuh"

namespace k {

// Device-resident loop state: the persistent step loop never returns to the host.
struct DecodeCtl {
  int step;                    // current time index t (same for every sample during a decode)
  int start;                   // step the decode started at (index 0 of noise / trajectory)
  unsigned int done;           // block-completion counter used to advance `step` exactly once
  int pad;
  const float* noise;          // external Exp(1) noise [T][B*L][C] or nullptr -> Philox
  uint8_t* traj;               // trajectory [T][B][L] or nullptr
  unsigned long long seed;     // Philox seed when noise == nullptr
  const unsigned long long* group_seeds;   // or one seed per reference batch (group) of a fused launch: the draw for (sample b,
                               // position l) then uses seed group_seeds[b / group] at position (b % group) * L + l, i.e. exactly
                               // what a separate launch of that group with that seed would draw
  unsigned long long* stamps;  // [L][2] %globaltimer (ns) at the start of a step's first kernel and at the end of its last one
                               // (biom3_debug_copy "stamps": shows that consecutive graph replays leave no host-side gap)
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int MAXV = 8;   // up to D = 1024 (D / 128 float4 groups per lane)

// LayerNorm of one row held as nv float4 per lane; writes bf16.
__device__ __forceinline__ void ln_row_store(const float4 (&v)[MAXV], int nv, int D, int lane,
                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                             __nv_bfloat16* __restrict__ dst) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / float(D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / float(D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + col));
      const float o0 = (v[i].x - mean) * rstd * gm.x + bt.x;
      const float o1 = (v[i].y - mean) * rstd * gm.y + bt.y;
      const float o2 = (v[i].z - mean) * rstd * gm.z + bt.z;
      const float o3 = (v[i].w - mean) * rstd * gm.w + bt.w;
      *reinterpret_cast<uint2*>(dst + col) = make_uint2(ptx::pack_bf16x2(o0, o1), ptx::pack_bf16x2(o2, o3));
    }
}

// cvec[b][j][d] = Ttab[t_b][j][d] + Y[b][j][d]
__global__ void cond_build_kernel(const float* __restrict__ Ttab, const float* __restrict__ Y,
                                  const int* __restrict__ t_per_sample, const DecodeCtl* __restrict__ ctl,
                                  float* __restrict__ cvec, int B, int JD) {
  ptx::pdl_sync();
  const int b = blockIdx.y;
  const int t = t_per_sample ? t_per_sample[b] : ctl->step;
  if (!t_per_sample && ctl->stamps && blockIdx.x == 0 && b == 0 && threadIdx.x == 0) ctl->stamps[2 * t] = global_timer_ns();
  const float4* tt = reinterpret_cast<const float4*>(Ttab + size_t(t) * JD);
  const float4* yy = reinterpret_cast<const float4*>(Y + size_t(b) * JD);
  float4* cc = reinterpret_cast<float4*>(cvec + size_t(b) * JD);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < JD / 4; i += gridDim.x * blockDim.x) {
    const float4 a = __ldg(tt + i), c = __ldg(yy + i);
    cc[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

// u[b,l,:] = E[x] + (ax0[l / W] + ax1[l % W]) + cvec[b][0][:].  Writes the fp32 residual stream, its raw
// bf16 copy (A operand of the first QKV GEMM) and the row's (sum, sum of squares) for the folded LayerNorm
// (slot 0 of `parts`; the other slots are zeroed).
__global__ void __launch_bounds__(256)
embed_kernel(const uint8_t* __restrict__ state, const float* __restrict__ emb, const float* __restrict__ ax0,
             const float* __restrict__ ax1, const float* __restrict__ cvec, int cond_stride, float* __restrict__ u,
             __nv_bfloat16* __restrict__ ub, __nv_bfloat16* __restrict__ ulo, float* __restrict__ stats, int parts,
             int rows, int L, int W, int D) {
  ptx::pdl_sync();
  const int lane = threadIdx.x & 31;
  const int nv = D / 128;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
    const int b = row / L, l = row % L;
    const int tok = state[row];
    const float* e = emb + size_t(tok) * D;
    const float* p0 = ax0 + size_t(l / W) * D;
    const float* p1 = ax1 + size_t(l % W) * D;
    const float* cv = cvec + size_t(b) * cond_stride;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(e + col));
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(p0 + col));
        const float4 x2 = __ldg(reinterpret_cast<const float4*>(p1 + col));
        const float4 x3 = __ldg(reinterpret_cast<const float4*>(cv + col));
        const float4 v = make_float4((x0.x + (x1.x + x2.x)) + x3.x, (x0.y + (x1.y + x2.y)) + x3.y,
                                     (x0.z + (x1.z + x2.z)) + x3.z, (x0.w + (x1.w + x2.w)) + x3.w);
        const uint32_t h0 = ptx::pack_bf16x2(v.x, v.y), h1 = ptx::pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(ub + size_t(row) * D + col) = make_uint2(h0, h1);
        if (ulo)              // split residual stream: u = hi + lo, both bf16 (see gemm::EPI_BIAS_RESID_SPLIT)
          *reinterpret_cast<uint2*>(ulo + size_t(row) * D + col) =
              make_uint2(ptx::pack_bf16x2(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xffff0000u)),
                         ptx::pack_bf16x2(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xffff0000u)));
        else
          *reinterpret_cast<float4*>(u + size_t(row) * D + col) = v;
        s += (v.x + v.y) + (v.z + v.w);
        q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane < parts)
      *reinterpret_cast<float2*>(stats + (size_t(row) * parts + lane) * 2) = lane == 0 ? make_float2(s, q) : make_float2(0.f, 0.f);
  }
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ u, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ a, int rows, int D) {
  const int lane = threadIdx.x & 31;
  const int nv = D / 128;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
    float4 v[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) v[i] = *reinterpret_cast<const float4*>(u + size_t(row) * D + (i * 32 + lane) * 4);
    ln_row_store(v, nv, D, lane, gamma, beta, a + size_t(row) * D);
  }
}

// ---------------------------------------------------------------- Philox4x32-10 -> Exp(1)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// Exp(1) draw for (step, flat position, class) by inversion: q = -log1p(-v), v uniform on the open interval (0, 1).
// v = (w + 1/2) 2^-32 from one 32-bit Philox word, capped at the largest float below 1, so q is finite and strictly
// positive (1.2e-10 <= q <= 16.7): no draw can win the argmax(p / q) race by being zero, negative or infinite.  IEEE
// log1pf (not the fast-math logarithm) keeps the small draws — the ones that win races — accurate to an ulp.
// oracle/philox.py restates this stream on the CPU; biom3_debug_noise() exports it for the parity tests.
__device__ __forceinline__ float philox_exp1(unsigned long long seed, int step, int pos, int c) {
  const uint4 r = philox4x32_10(make_uint4(uint32_t(pos), uint32_t(step), uint32_t(c >> 2), 0u),
                                make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const float v = fminf(fmaf(float(w[c & 3]), 2.3283064365386963e-10f, 1.1641532182693481e-10f), 0.99999994f);
  return -log1pf(-v);
}

// Test hook (biom3_debug_noise): the Exp(1) draws head_kernel consumes at `step`, as noise[pos][c], pos = b * L + l
__global__ void debug_noise_kernel(unsigned long long seed, int step, int n_pos, int C, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pos * C) out[i] = philox_exp1(seed, step, i / C, i % C);
}

// ---------------------------------------------------------------- sampling paths on the device
// One uniformly random permutation of 0..L-1 per row: sort L Philox keys (64 bits, ties broken by index, so the result
// is a permutation for every seed) with a bitonic network in shared memory.  Replaces the per-sample
// `torch.randperm(args.diffusion_steps)` of /root/reference/run_ProteoScribe_sample.py:103-105 when the caller asks for
// device-side paths; not bit-compatible with torch's generator (like the on-device noise).
// One block per row; N = next power of two >= L (padding keys are +inf and sort to the end).
__global__ void __launch_bounds__(1024)
random_paths_kernel(unsigned long long seed, long long* __restrict__ path, int L, int N) {
  extern __shared__ unsigned long long rp_keys[];          // [N] keys, then [N] uint32 indices
  uint32_t* idx = reinterpret_cast<uint32_t*>(rp_keys + N);
  const int row = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < L) {
      const uint4 r = philox4x32_10(make_uint4(uint32_t(i), uint32_t(row), 0x70617468u, 0u), make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
      k = (static_cast<unsigned long long>(r.x) << 32) | r.y;
      if (k == ~0ull) k -= 1;                               // keep real keys below the padding
    }
    rp_keys[i] = k;
    idx[i] = uint32_t(i);
  }
  __syncthreads();
  for (int size = 2; size <= N; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < N / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));          // index of the lower element of pair i
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long ka = rp_keys[lo], kb = rp_keys[hi];
        const uint32_t ia = idx[lo], ib = idx[hi];
        const bool a_gt_b = ka > kb || (ka == kb && ia > ib);
        if (a_gt_b == up) {
          rp_keys[lo] = kb; rp_keys[hi] = ka;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < L; i += blockDim.x) path[size_t(row) * L + i] = static_cast<long long>(idx[i]);
}

// ---------------------------------------------------------------- categorical draw (one warp, lane = class)
// p = softmax(logits); p /= sum(p); token = argmax(p / q), ties -> lowest class id.
// The two sums run sequentially over c = 0..C-1 (the order the fp32 CPU reference accumulates in).
__device__ __forceinline__ int categorical_draw(float logit, float q, int lane, int C) {
  const bool valid = lane < C;
  float mx = valid ? logit : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = valid ? expf(logit - mx) : 0.f;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += __shfl_sync(0xffffffffu, e, c);
  float p = e / s;
  float s2 = 0.f;
  for (int c = 0; c < C; ++c) s2 += __shfl_sync(0xffffffffu, p, c);
  p = p / s2;
  float r = valid ? p / q : -INFINITY;
  int idx = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float r2 = __shfl_xor_sync(0xffffffffu, r, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (r2 > r || (r2 == r && i2 < idx)) { r = r2; idx = i2; }
  }
  return idx;
}

// Last-layer row compaction (decode only).  After the last block's attention every remaining operation of the step
// (out-proj, residual, LayerNorm, MLP, final norm, vocab head) is per token, and the sampler consumes only the
// tokens of the SELECTED list below (the reference computes all B*L and discards the rest,
// sampling_analysis.py:251-256).  This kernel copies the three per-token inputs of those operations — the attention
// output row and the (hi, lo) residual row — for list entry ti = b' * group + g into row ti of compact buffers, so
// the last layer's three GEMMs run on ceil256(B * group) rows instead of B * L.  Rows >= ntok are zero-filled
// padding up to the GEMM tile height.  One warp per row, 16-byte accesses.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __nv_bfloat16* __restrict__ att, const __nv_bfloat16* __restrict__ u_hi,
                   const __nv_bfloat16* __restrict__ u_lo, __nv_bfloat16* __restrict__ att_c,
                   __nv_bfloat16* __restrict__ hi_c, __nv_bfloat16* __restrict__ lo_c, const int* __restrict__ inv_path,
                   const DecodeCtl* __restrict__ ctl, int L, int D, int group, int ntok, int rows_c) {
  ptx::pdl_sync();
  const int lane = threadIdx.x & 31;
  const int step = ctl->step;
  const int n16 = D / 8;                           // 16-byte pieces per row
  for (int ti = blockIdx.x * 8 + (threadIdx.x >> 5); ti < rows_c; ti += gridDim.x * 8) {
    uint4* da = reinterpret_cast<uint4*>(att_c + size_t(ti) * D);
    uint4* dh = reinterpret_cast<uint4*>(hi_c + size_t(ti) * D);
    uint4* dl = reinterpret_cast<uint4*>(lo_c + size_t(ti) * D);
    if (ti < ntok) {
      const int b = ti / group;
      const int src = (b / group) * group + (ti % group);          // sample whose location is written (head_kernel)
      const size_t row = size_t(b) * L + inv_path[size_t(src) * L + step];
      const uint4* sa = reinterpret_cast<const uint4*>(att + row * D);
      const uint4* sh = reinterpret_cast<const uint4*>(u_hi + row * D);
      const uint4* sl = reinterpret_cast<const uint4*>(u_lo + row * D);
      for (int i = lane; i < n16; i += 32) {
        da[i] = sa[i];
        dh[i] = sh[i];
        dl[i] = sl[i];
      }
    } else {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (int i = lane; i < n16; i += 32) {
        da[i] = z;
        dh[i] = z;
        dl[i] = z;
      }
    }
  }
}

// Final LayerNorm + vocab head for a list of tokens; optionally writes logits [B][C][L] and/or draws
// and scatters tokens.  SELECTED mode (decode): token list = {(b', loc[b]) : b in group(b')},
// loc[b] = inv_path[b][step]  -> this IS the reference's B x B unmask write.  ALL mode: every token.
struct HeadArgs {
  const float* u;            // [rows][D] fp32 hidden after the last block, or nullptr when it is stored split:
  const __nv_bfloat16* u_hi; //   u = u_hi + u_lo, both bf16 [rows][D]
  const __nv_bfloat16* u_lo;
  const float* gamma; const float* beta;     // final LayerNorm
  const float* w_out;        // [C][D] fp32
  const float* b_out;        // [C]
  float* logits_out;         // [B][C][L] or nullptr
  uint8_t* state;            // [B][L] updated in place when sampling, or nullptr
  const int* inv_path;       // [B][L] (SELECTED mode)
  const DecodeCtl* ctl;      // step / noise / seed (sampling)
  int B, L, D, C, group;     // group = samples per reference batch (SELECTED mode); 0 -> ALL mode
  int compact;               // SELECTED mode: the hidden rows are compacted, entry ti lives in row ti (gather_rows_kernel)
};

__global__ void __launch_bounds__(256)
head_kernel(const HeadArgs a) {
  ptx::pdl_sync();
  extern __shared__ float s_w[];                 // [C][D] staged vocab head
  for (int i = threadIdx.x; i < a.C * a.D / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(a.w_out) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nv = a.D / 128;
  const int ntok = a.group > 0 ? a.B * a.group : a.B * a.L;
  const int step = a.ctl ? a.ctl->step : 0;
  for (int ti = blockIdx.x * 8 + (threadIdx.x >> 5); ti < ntok; ti += gridDim.x * 8) {
    int b, l;
    if (a.group > 0) {
      b = ti / a.group;
      const int src = (b / a.group) * a.group + (ti % a.group);     // sample whose location is written
      l = a.inv_path[size_t(src) * a.L + step];
    } else {
      b = ti / a.L;
      l = ti % a.L;
    }
    const size_t row = size_t(b) * a.L + l;        // token position (state / logits / noise index)
    const size_t hrow = a.compact ? size_t(ti) : row;   // row of the hidden state
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        if (a.u) {
          v[i] = *reinterpret_cast<const float4*>(a.u + hrow * a.D + (i * 32 + lane) * 4);
        } else {
          const uint2 h = *reinterpret_cast<const uint2*>(a.u_hi + hrow * a.D + (i * 32 + lane) * 4);
          const uint2 l = *reinterpret_cast<const uint2*>(a.u_lo + hrow * a.D + (i * 32 + lane) * 4);
          v[i] = make_float4(__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16),
                             __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u),
                             __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16),
                             __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u));
        }
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    const float mean = warp_sum(s) / float(a.D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        const float x0 = v[i].x - mean, x1 = v[i].y - mean, x2 = v[i].z - mean, x3 = v[i].w - mean;
        q += (x0 * x0 + x1 * x1) + (x2 * x2 + x3 * x3);
      }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / float(a.D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma + col));
        const float4 bt = __ldg(reinterpret_cast<const float4*>(a.beta + col));
        v[i] = make_float4((v[i].x - mean) * rstd * gm.x + bt.x, (v[i].y - mean) * rstd * gm.y + bt.y,
                           (v[i].z - mean) * rstd * gm.z + bt.z, (v[i].w - mean) * rstd * gm.w + bt.w);
      }
    float my_logit = 0.f;                        // lane c ends up holding logit c
    for (int c = 0; c < a.C; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
          const float4 w4 = *reinterpret_cast<const float4*>(s_w + size_t(c) * a.D + (i * 32 + lane) * 4);
          acc = fmaf(v[i].x, w4.x, acc);
          acc = fmaf(v[i].y, w4.y, acc);
          acc = fmaf(v[i].z, w4.z, acc);
          acc = fmaf(v[i].w, w4.w, acc);
        }
      acc = warp_sum(acc);
      if (lane == c) my_logit = acc + __ldg(a.b_out + c);
    }
    if (a.logits_out && lane < a.C) a.logits_out[(size_t(b) * a.C + lane) * a.L + l] = my_logit;
    if (a.state) {
      const int pos = int(row);
      float qn = 1.f;
      if (lane < a.C) {
        if (a.ctl->noise) {
          const size_t BLC = size_t(a.B) * a.L * a.C;
          qn = __ldg(a.ctl->noise + size_t(step - a.ctl->start) * BLC + size_t(pos) * a.C + lane);
        } else {
          if (a.ctl->group_seeds)
            qn = philox_exp1(a.ctl->group_seeds[b / a.group], step, (b % a.group) * a.L + l, lane);
          else
            qn = philox_exp1(a.ctl->seed, step, pos, lane);
        }
      }
      const int tok = categorical_draw(my_logit, qn, lane, a.C);
      if (lane == 0) a.state[row] = uint8_t(tok);
    }
  }
}

// K11 as the reference runs it: draw a token at EVERY position.  logits [B][C][L], noise [B*L][C]
// -> tok int64 [B][L].  One thread per position; bandwidth bound (logits + noise read once).
__global__ void __launch_bounds__(256)
sample_all_kernel(const float* __restrict__ logits, const float* __restrict__ noise, long long* __restrict__ tok,
                  int B, int L, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L, l = i % L;
  const float* lg = logits + size_t(b) * C * L + l;
  const float* qn = noise + size_t(i) * C;
  float x[32];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      x[c] = __ldg(lg + size_t(c) * L);
      mx = fmaxf(mx, x[c]);
    }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      x[c] = expf(x[c] - mx);
      s += x[c];
    }
  float s2 = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      x[c] = x[c] / s;
      s2 += x[c];
    }
  float best = -INFINITY;
  int bi = 0;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      const float r = (x[c] / s2) / __ldg(qn + c);
      if (r > best) { best = r; bi = c; }
    }
  tok[i] = bi;
}

// Same draw, staged: one block = 256 consecutive positions of one sequence (needs L % 256 == 0).  The block's logits
// tile [C][256] and its noise block [256][C] (contiguous in memory) are copied to shared memory with 16-byte
// cp.async, so every global access is a full coalesced line and ~2 KB * C per block is in flight without holding
// registers; the per-position math then reads shared memory conflict-free (logits: consecutive threads, consecutive
// words; noise: stride C words, odd for C = 29).  Three blocks per SM overlap one block's loads with another's math
// (29 exp + 87 IEEE divides per position: the kernel is as much ALU as HBM time; a persistent two-stage version with
// fewer resident warps measured slower).  Same operation order as sample_all_kernel: bit-identical tokens.
constexpr int SAMPLE_TILE = 256;
__global__ void __launch_bounds__(SAMPLE_TILE)
sample_all_tiled_kernel(const float* __restrict__ logits, const float* __restrict__ noise, long long* __restrict__ tok,
                        int B, int L, int C) {
  extern __shared__ __align__(16) float s_samp[];
  float* sl = s_samp;                           // [C][256]
  float* sn = s_samp + C * SAMPLE_TILE;         // [256][C]
  const int tiles_per_row = L / SAMPLE_TILE;
  const int b = blockIdx.x / tiles_per_row, l0 = (blockIdx.x % tiles_per_row) * SAMPLE_TILE;
  const size_t i0 = size_t(b) * L + l0;
  const float* lg = logits + size_t(b) * C * L + l0;
  const float* qn = noise + i0 * C;
  for (int v = threadIdx.x; v < C * (SAMPLE_TILE / 4); v += SAMPLE_TILE) {
    const int c = v / (SAMPLE_TILE / 4), part = v % (SAMPLE_TILE / 4);
    ptx::cp_async_16(ptx::smem_u32(sl + c * SAMPLE_TILE + part * 4), lg + size_t(c) * L + part * 4);
    ptx::cp_async_16(ptx::smem_u32(sn + v * 4), qn + size_t(v) * 4);
  }
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();
  const int t = threadIdx.x;
  float x[32];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      x[c] = sl[c * SAMPLE_TILE + t];
      mx = fmaxf(mx, x[c]);
    }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      x[c] = expf(x[c] - mx);
      sum += x[c];
    }
  float s2 = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      x[c] = x[c] / sum;
      s2 += x[c];
    }
  float best = -INFINITY;
  int bi = 0;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c < C) {
      const float r = (x[c] / s2) / sn[t * C + c];
      if (r > best) { best = r; bi = c; }
    }
  tok[i0 + t] = bi;
}

// K12: the reference unmask write in one launch, no scratch.  Block = one source sample b: find loc[b] = the first
// position with path[b][loc] == step (argmax of the match mask, sampling_analysis.py:254; 0 when nothing matches, as
// argmax of an all-false mask is), then for every sample b' of b's group: state[b'][loc[b]] = tok[b'][loc[b]].
__global__ void __launch_bounds__(256)
unmask_scan_kernel(const long long* __restrict__ tok, const long long* __restrict__ path, long long* __restrict__ state,
                   int B, int L, int group, int step) {
  __shared__ int s_loc;
  const int src = blockIdx.x;
  if (threadIdx.x == 0) s_loc = INT_MAX;
  __syncthreads();
  int best = INT_MAX;
  for (int l = threadIdx.x; l < L; l += blockDim.x)
    if (path[size_t(src) * L + l] == step) best = min(best, l);
  if (best != INT_MAX) atomicMin(&s_loc, best);
  __syncthreads();
  const int loc = s_loc == INT_MAX ? 0 : s_loc;
  const int g0 = (src / group) * group;
  for (int i = threadIdx.x; i < group; i += blockDim.x) {
    const size_t at = size_t(g0 + i) * L + loc;
    state[at] = tok[at];
  }
}

__global__ void inverse_path_kernel(const long long* __restrict__ path, int* __restrict__ inv, int B, int L, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L;
  const long long t = path[i];
  if (t >= 0 && t < L) inv[size_t(b) * L + t] = i % L;
  else atomicOr(err, 4);
}

// Inputs are range-checked where they enter the resident state, once per call: an out-of-range value sets a sticky bit in
// `err` (1: token id outside [0, limit), 2: time index outside [0, limit), 4: path entry outside [0, limit)) and is
// clamped so that nothing downstream indexes out of bounds; biom3_input_errors() reports and clears the bits.
__global__ void i64_to_u8_kernel(const long long* __restrict__ src, uint8_t* __restrict__ dst, int n, int limit, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long v = src[i];
  if (v < 0 || v >= limit) {
    atomicOr(err, 1);
    v = v < 0 ? 0 : limit - 1;
  }
  dst[i] = uint8_t(v);
}
__global__ void u8_to_i64_kernel(const uint8_t* __restrict__ src, long long* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
__global__ void i64_to_i32_kernel(const long long* __restrict__ src, int* __restrict__ dst, int n, int limit, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long v = src[i];
  if (v < 0 || v >= limit) {
    atomicOr(err, 2);
    v = v < 0 ? 0 : limit - 1;
  }
  dst[i] = int(v);
}

// Last kernel of a step: snapshot the state into the trajectory, then advance `step` exactly once
// (the last block to finish does it, after every block has read the old value).
__global__ void advance_kernel(DecodeCtl* ctl, const uint8_t* __restrict__ state, int n) {
  ptx::pdl_sync();
  const int step = ctl->step;
  if (ctl->traj) {
    uint8_t* dst = ctl->traj + size_t(step - ctl->start) * n;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n; i += gridDim.x * blockDim.x * 16)
      *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(state + i);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl->done, 1u) == gridDim.x - 1) {
      ctl->done = 0;
      if (ctl->stamps) ctl->stamps[2 * step + 1] = global_timer_ns();
      ctl->step = step + 1;
    }
  }
}

// ---------------------------------------------------------------- conditioning MLP pieces (fp32)
// te[t][i] = sin/cos((t / num_steps * 4000) * exp(-i * ln(1e4) / (half - 1)))   [L][D]
__global__ void time_embedding_kernel(float* __restrict__ te, int L, int D, float num_steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * D) return;
  const int t = i / D, c = i % D, half = D / 2;
  const float x = float(t) / num_steps * 4000.0f;
  const float e = logf(10000.0f) / float(half - 1);
  const int f = c < half ? c : c - half;
  const float arg = x * expf(float(f) * -e);
  te[i] = c < half ? sinf(arg) : cosf(arg);
}

// C[M][N] = act(A[M][K] . W[N][K]^T + bias[N]); act: 0 none, 1 softplus(beta 1, threshold 20).
// 64 x 64 tile, 16 x 16 threads, 4 x 4 outputs each.  Runs once per model load / per prompt.
__global__ void __launch_bounds__(256)
sgemm_bias_act_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                      float* __restrict__ Cm, int M, int N, int K, int act) {
  __shared__ float sa[16][64 + 1], sb[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      sa[c][r] = (m0 + r < M && k0 + c < K) ? A[size_t(m0 + r) * K + k0 + c] : 0.f;     // K need not be a multiple of 16
      sb[c][r] = (n0 + r < N && k0 + c < K) ? W[size_t(n0 + r) * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float v = acc[i][j] + bias[n];
        if (act == 1) v = v > 20.f ? v : log1pf(expf(v));
        else if (act == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));   // exact erf-GELU
        Cm[size_t(m) * N + n] = v;
      }
    }
}

// out[r][j][d] = in[r][d * depth + j]   (the reference's reshape(B,1,D,1,depth)[..., j] layout)
__global__ void cond_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int D,
                                      int depth) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t n = size_t(rows) * D * depth;
  if (i >= n) return;
  const int d = int(i % D);
  const int j = int((i / D) % depth);
  const size_t r = i / (size_t(D) * depth);
  out[i] = in[r * size_t(D) * depth + size_t(d) * depth + j];
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

}  // namespace k
