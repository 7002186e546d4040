// Attention kernels for the ProteoScribe block (heads [0,NL) windowed softmax, heads [NL,H) linear).
// Semantics restated from the un-vendored LocalAttention / linear_attn blocks the reference calls through
// /root/reference/Stage3_source/cond_diff_transformer_layer.py:123-143,171.
//
// Input layout (written by the QKV GEMM epilogue): qkv bf16 [3][B][H][L][32]; output bf16 [B*L][H*32].
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace attn {

constexpr int DH = 32;     // head dim (fixed by the kernels; checked at create time)
constexpr int WIN = 128;   // local window (fixed; checked at create time)

// 2^x as a single MUFU.EX2 (exp2f() adds denormal range handling: 3 more instructions per element)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 64-byte rows (32 bf16): the 16-byte chunk index is XOR-swizzled with (row >> 1) & 3 so that ldmatrix (8 rows x 16 B) is
// bank-conflict free
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return uint32_t(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }


// ------------------------------------------------------------------------------------------------
// Linear attention (global heads): q <- softmax_d(q) * dh^-0.5 ; k <- softmax over the L tokens ;
// ctx = k^T v (32 x 32) ; out = q ctx.  One CTA (4 warps) per (global head, sample).
//   phase A  each warp streams a quarter of the sequence in 32-row chunks (cp.async, 2 stages) and
//            accumulates exp(k - reference)^T v on the tensor cores (mma.sync, fp32 accumulators).  The softmax over
//            tokens is shift invariant, so the per-feature reference only has to stay within e^8 of the running
//            maximum: it moves (and the accumulators are rescaled) only when a chunk's maximum exceeds it by more
//            than 8 — after the first chunk practically never — which takes the per-chunk rescale of the 32 x 32
//            context out of the loop (the kernel is issue bound)
//   merge    the four partial (max, denominator, ctx) sets are combined in shared memory -> ctx^T bf16
//   phase B  each warp softmaxes its q rows in registers (quad shuffles) and multiplies by ctx
// Round 2 built the two phases as two kernels, twice, and kept this one (profiles/r02_linear_two_kernel.log): TMA-fed
// context CTAs per 512 rows with global partials merged by the last CTA to arrive plus an output kernel per 128 rows
// (37 + 19 us under ncu against 35 us for this kernel; step +0.25 ms); this kernel stopping after its merge plus an
// output kernel per 512 rows with the row sums taken from the tensor cores (19.7 + 21.5 us; step +0.02 ms).
// ------------------------------------------------------------------------------------------------
constexpr int LIN_CH = 32;                         // rows per chunk
constexpr float LIN_LAZY = 8.0f;                   // the column reference may trail the running maximum by e^8
constexpr float LOG2E = 1.4426950408889634f;
constexpr int LIN_STAGE_BYTES = 2 * LIN_CH * 64;   // one k + v chunk
constexpr int LIN_NST = 3;                         // k/v stages per warp: two chunks in flight while one is consumed (the kernel is
                                                   // bound by load latency, not bytes: round 1 had two stages = one chunk in flight)
constexpr int LIN_QST = 4;                         // q stages per warp in phase B (LIN_CH x 64 bytes each, inside the same space)
constexpr int LIN_WARP_BYTES = LIN_NST * LIN_STAGE_BYTES;
// the four warps' partial (ctx, max, denominator) sets alias the stage space once phase A is over: 4 CTAs per SM still fit
constexpr int LIN_SMEM_BYTES = 4 * LIN_WARP_BYTES + DH * 64;
static_assert(4 * DH * DH * 4 + 2 * 4 * DH * 4 <= 4 * LIN_WARP_BYTES, "partials must fit in the stage space");
static_assert(LIN_QST * LIN_CH * 64 <= LIN_WARP_BYTES, "q stages must fit in the warp's stage space");

__device__ __forceinline__ void lin_load_chunk(uint32_t dst, const __nv_bfloat16* src, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = lane + 32 * i;                   // 128 pieces of 16 B = 32 rows x 64 B
    ptx::cp_async_16(dst + swz(p >> 2, p & 3), src + p * 8);
  }
}

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u);
  return make_float2(__bfloat162float(b.x), __bfloat162float(b.y));
}

// QSOFT = false (the decode's form): q arrives already softmaxed over its features (the QKV GEMM epilogue did it from the
// fp32 accumulator, gemm::Params::qkv_chunk kind 1); phase B is then load -> MMA -> store.
template <bool QSOFT>
__global__ void __launch_bounds__(128)
linear_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int H, int L,
                        int NL, float q_scale, int reverse) {
  ptx::pdl_sync();
  const int h = NL + blockIdx.x, b = reverse ? int(gridDim.y) - 1 - int(blockIdx.y) : int(blockIdx.y);
  const size_t head_stride = size_t(L) * DH;
  const size_t plane = size_t(B) * H * head_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int rows_per_warp = L / 4;
  const int nchunks = rows_per_warp / LIN_CH;
  const __nv_bfloat16* qg = qkv + (size_t(b) * H + h) * head_stride + size_t(warp) * rows_per_warp * DH;
  const __nv_bfloat16* kg = qg + plane;
  const __nv_bfloat16* vg = kg + plane;

  extern __shared__ __align__(128) uint8_t lin_smem[];
  const uint32_t stage0 = ptx::smem_u32(lin_smem) + warp * LIN_WARP_BYTES;
  float* pctx = reinterpret_cast<float*>(lin_smem);                           // [4][32][32]: aliases the stages after phase A
  float* pmax = pctx + 4 * DH * DH;                                           // [4][32]
  float* pden = pmax + 4 * DH;                                                // [4][32]
  uint8_t* ctxT = lin_smem + 4 * LIN_WARP_BYTES;                              // [32 e][32 d] bf16, swizzled rows

  // ---------------------------------------------------------------- phase A
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
  float mrun[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
  float den[2][2] = {{0.f, 0.f}, {0.f, 0.f}};

  auto load_kv = [&](int c) {
    const uint32_t dst = stage0 + (c % LIN_NST) * LIN_STAGE_BYTES;
    lin_load_chunk(dst, kg + size_t(c) * LIN_CH * DH, lane);
    lin_load_chunk(dst + LIN_CH * 64, vg + size_t(c) * LIN_CH * DH, lane);
    ptx::cp_async_commit();
  };
  for (int c = 0; c < LIN_NST - 1 && c < nchunks; ++c) load_kv(c);
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t cur = stage0 + (c % LIN_NST) * LIN_STAGE_BYTES;
    // chunk c + NST - 1 goes into the stage chunk c - 1 was consumed from (the __syncwarp at the loop's end orders it)
    if (c + LIN_NST - 1 < nchunks) {
      load_kv(c + LIN_NST - 1);
      ptx::cp_async_wait<LIN_NST - 1>();
    } else if (c + 1 < nchunks) {                  // LIN_NST == 3: exactly one younger chunk is still in flight
      ptx::cp_async_wait<1>();
    } else {
      ptx::cp_async_wait<0>();
    }
    __syncwarp();
    const uint32_t sk = cur, sv = cur + LIN_CH * 64;
    uint32_t kr[2][2][4];                          // [k-step][m-tile(d)][a0..a3], raw k as A^T fragments
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ptx::ldmatrix_x4_trans(sk + swz(ks * 16 + (lane & 7) + 8 * (lane >> 4), 2 * mt + ((lane >> 3) & 1)),
                               kr[ks][mt][0], kr[ks][mt][1], kr[ks][mt][2], kr[ks][mt][3]);
    // running per-feature max: this thread owns features 16*mt + g (a0, a2) and 16*mt + 8 + g (a1, a3)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float cm = -INFINITY;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const float2 x0 = bf2_to_f2(kr[ks][mt][hf]), x1 = bf2_to_f2(kr[ks][mt][hf + 2]);
          cm = fmaxf(cm, fmaxf(fmaxf(x0.x, x0.y), fmaxf(x1.x, x1.y)));
        }
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
        // lazy reference (quad-uniform: the four lanes of a quad own the same features): -inf on the first chunk
        if (cm > mrun[mt][hf] + LIN_LAZY) {
          const float f = __expf(mrun[mt][hf] - cm);       // 0 on the first chunk
          mrun[mt][hf] = cm;
          den[mt][hf] *= f;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            acc[mt][nt][2 * hf] *= f;
            acc[mt][nt][2 * hf + 1] *= f;
          }
        }
        const float ml2 = mrun[mt][hf] * LOG2E;
        float dsum = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            const float2 x = bf2_to_f2(kr[ks][mt][hf + 2 * q2]);
            const float e0 = fast_ex2(fmaf(x.x, LOG2E, -ml2)), e1 = fast_ex2(fmaf(x.y, LOG2E, -ml2));
            dsum += e0 + e1;
            kr[ks][mt][hf + 2 * q2] = ptx::pack_bf16x2(e0, e1);
          }
        den[mt][hf] += dsum;
      }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int ep = 0; ep < 2; ++ep) {
        uint32_t v0, v1, v2, v3;
        ptx::ldmatrix_x4_trans(sv + swz(ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * ep + (lane >> 4)), v0, v1, v2, v3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          ptx::mma_bf16_16816(acc[mt][2 * ep], kr[ks][mt][0], kr[ks][mt][1], kr[ks][mt][2], kr[ks][mt][3], v0, v1);
          ptx::mma_bf16_16816(acc[mt][2 * ep + 1], kr[ks][mt][0], kr[ks][mt][1], kr[ks][mt][2], kr[ks][mt][3], v2, v3);
        }
      }
    __syncwarp();
  }
  // partial results of this warp -> shared (the partials alias the stages: every warp must be done reading its chunks)
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float dn = den[mt][hf];
      dn += __shfl_xor_sync(0xffffffffu, dn, 1);
      dn += __shfl_xor_sync(0xffffffffu, dn, 2);
      const int d = 16 * mt + 8 * hf + g;
      if (t == 0) {
        pmax[warp * DH + d] = mrun[mt][hf];
        pden[warp * DH + d] = dn;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float* dst = pctx + (size_t(warp) * DH + d) * DH + 8 * nt + 2 * t;
        dst[0] = acc[mt][nt][2 * hf];
        dst[1] = acc[mt][nt][2 * hf + 1];
      }
    }
  __syncthreads();
  // ---------------------------------------------------------------- merge -> ctx^T (bf16), q scale folded in
  for (int idx = threadIdx.x; idx < DH * DH; idx += 128) {
    const int d = idx >> 5, e = idx & 31;
    const float m0 = pmax[d], m1 = pmax[DH + d], m2 = pmax[2 * DH + d], m3 = pmax[3 * DH + d];
    const float mm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    const float f0 = __expf(m0 - mm), f1 = __expf(m1 - mm), f2 = __expf(m2 - mm), f3 = __expf(m3 - mm);
    const float num = pctx[(0 * DH + d) * DH + e] * f0 + pctx[(1 * DH + d) * DH + e] * f1 +
                      pctx[(2 * DH + d) * DH + e] * f2 + pctx[(3 * DH + d) * DH + e] * f3;
    const float dn = pden[d] * f0 + pden[DH + d] * f1 + pden[2 * DH + d] * f2 + pden[3 * DH + d] * f3;
    *reinterpret_cast<__nv_bfloat16*>(ctxT + swz(e, d >> 3) + (d & 7) * 2) = __float2bfloat16_rn(num / dn * q_scale);
  }
  __syncthreads();
  // ---------------------------------------------------------------- phase B
  uint32_t cb[4][4];                               // ctx as B fragments: [n-tile(e)][b0 ks0, b1 ks0, b0 ks1, b1 ks1]
  const uint32_t sc = ptx::smem_u32(ctxT);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
    ptx::ldmatrix_x4(sc + swz(8 * nt + (lane & 7), lane >> 3), cb[nt][0], cb[nt][1], cb[nt][2], cb[nt][3]);
  const int D = H * DH;
  auto load_q = [&](int c) {
    lin_load_chunk(stage0 + (c % LIN_QST) * (LIN_CH * 64), qg + size_t(c) * LIN_CH * DH, lane);
    ptx::cp_async_commit();
  };
  for (int c = 0; c < LIN_QST - 1 && c < nchunks; ++c) load_q(c);
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t cur = stage0 + (c % LIN_QST) * (LIN_CH * 64);
    if (c + LIN_QST - 1 < nchunks) {
      load_q(c + LIN_QST - 1);
      ptx::cp_async_wait<LIN_QST - 1>();
    } else {
      ptx::cp_async_wait<0>();                     // tail: the last LIN_QST - 1 chunks are all in flight already
    }
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {               // 16 query rows each
      uint32_t qa[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ptx::ldmatrix_x4(cur + swz(mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * ks + (lane >> 4)), qa[ks][0],
                         qa[ks][1], qa[ks][2], qa[ks][3]);
      // rows g (regs 0, 2) and g + 8 (regs 1, 3): softmax over the 32 features held by the quad
      if constexpr (QSOFT) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float2 x[4];
        x[0] = bf2_to_f2(qa[0][hf]); x[1] = bf2_to_f2(qa[0][hf + 2]);
        x[2] = bf2_to_f2(qa[1][hf]); x[3] = bf2_to_f2(qa[1][hf + 2]);
        float mx = fmaxf(fmaxf(fmaxf(x[0].x, x[0].y), fmaxf(x[1].x, x[1].y)),
                         fmaxf(fmaxf(x[2].x, x[2].y), fmaxf(x[3].x, x[3].y)));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float s = 0.f;
        const float mxl = mx * LOG2E;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[i].x = fast_ex2(fmaf(x[i].x, LOG2E, -mxl));
          x[i].y = fast_ex2(fmaf(x[i].y, LOG2E, -mxl));
          s += x[i].x + x[i].y;
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float inv = 1.f / s;
        qa[0][hf] = ptx::pack_bf16x2(x[0].x * inv, x[0].y * inv);
        qa[0][hf + 2] = ptx::pack_bf16x2(x[1].x * inv, x[1].y * inv);
        qa[1][hf] = ptx::pack_bf16x2(x[2].x * inv, x[2].y * inv);
        qa[1][hf + 2] = ptx::pack_bf16x2(x[3].x * inv, x[3].y * inv);
      }
      }
      float o[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
        ptx::mma_bf16_16816(o[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], cb[nt][0], cb[nt][1]);
        ptx::mma_bf16_16816(o[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], cb[nt][2], cb[nt][3]);
      }
      const size_t row0 = size_t(b) * L + size_t(warp) * rows_per_warp + size_t(c) * LIN_CH + mt * 16 + g;
      __nv_bfloat16* o0 = out + row0 * D + h * DH + 2 * t;
      __nv_bfloat16* o1 = o0 + size_t(8) * D;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<uint32_t*>(o0 + nt * 8) = ptx::pack_bf16x2(o[nt][0], o[nt][1]);
        *reinterpret_cast<uint32_t*>(o1 + nt * 8) = ptx::pack_bf16x2(o[nt][2], o[nt][3]);
      }
    }
    __syncwarp();
  }
}


// one 128 x 32 bf16 tile (64-byte rows) as TMA writes it with the 64-byte swizzle
constexpr int TC_TILE = WIN * 64;

// smem matrix descriptor, 64-byte rows with the 64-byte swizzle (SBO = 8 rows x 64 B); used both for the
// K-major Q / K tiles and for V as an MN-major B operand (SBO = stride between 8-key groups)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

// Timeline of CTA 0 of a tracing launch (test hook biom3_debug_trace(0, ...)): [stream][who: 0 issuers, 1 softmax warp 0]
// [block g < 128][event] = clock64()
//   issuers: 0 before the S issue (slot wait included), 1 S issued, 2 before the PV issue, 3 p_ready seen, 4 PV issued
//   softmax: 0 loop top, 1 s_full seen, 2 scores loaded, 3 exponentials done, 4 P stored + arrived, 5 after epilogue
__device__ long long g_ms_trace[2][2][128][6];

// ================================================================================================
// Windowed softmax attention on tcgen05 (heads [0, NL)): query window w attends to key windows w-1, w, w+1 (those that
// exist), softmax over the real keys only.  Persistent, one CTA per SM, two independent streams per CTA; every stream has a
// TMA producer warp, TWO MMA issuer warps (S and PV) with division-free control loops, and four softmax warps.
//
// Why (profiles/r02_attn_trace_ms_issuer_bound.log, profiles/r02_attn_ablations.log): a clock64 timeline of variant 3
// showed the softmax warps idle ~1250 of every ~1450 cycles per 64-key block, waiting for scores; removing ALL
// exponentials did not change the kernel time (85 us -> 85 us) and removing the softmax work altogether only took it
// to 70 us.  The one thread per stream that issued both S = Q K^T and O += P V spent ~365 cycles per S (2 MMAs),
// ~560 per PV (4 MMAs) and ~300 on cursor arithmetic per block: with 128 x 64 x 16 MMAs (16-32 tensor-pipe cycles each)
// the serial issue path, not the tensor pipe, the MUFU pipe or TMA, bounded the kernel.
//
//   warp 3s+0   TMA producer of stream s: Q (double buffered) and K/V tiles (ring of NST stages), 64-byte swizzle
//   warp 3s+1   S issuer:  waits for the slot (s_free) and the tile, issues S_g = Q K_g^T, commits s_full; runs ahead of
//               the PV issuer as far as the three S / P slots allow
//   warp 3s+2   PV issuer: waits for P_g (p_ready), issues O (+)= P_g V_g with P read from TMEM, commits s_free
//               (and kv_free / o_full at tile / item ends)
//   warps 6..13 softmax: four warps per stream, one thread per query row (unchanged arithmetic: block maximum, exp2
//               against a lazily rescaled reference maximum, bf16 P written over the slot's own S columns)
// Descriptors are built once per tile (low word = address >> 4, the high word is a constant); ring slots and barrier
// phases are carried incrementally.  tcgen05.commit tracks the MMAs of the issuing thread, so each issuer's commits
// cover exactly its own instructions.
// TMEM per stream: S / P slots at 0, 64, 128; O buffers at 192, 224.
// ================================================================================================
constexpr int LOCAL_THREADS = 14 * 32;
constexpr int LOCAL_NST = 5;           // K/V ring stages per stream (3: +4 %, 4: +1.5 %, 6: equal; profiles/r02_attn_variants_v5.log)
template <int NST>
struct LocalCfg {
  static constexpr int STREAM_TILES = 2 + 2 * NST;
  static constexpr int SMEM_BYTES = 2 * STREAM_TILES * TC_TILE + 1024;
};
template <int NST>
struct LocalBars {
  uint64_t q_full[2], q_free[2], kv_full[NST], kv_free[NST], s_full[3], p_ready[3], s_free[3], o_full[2], o_free[2];
};
struct LocalItem {                      // one (window, sample, head)
  int w, b, h, w_lo, nkt;
};
constexpr uint32_t LOCAL_DESC_HI = (512u >> 4) | (1u << 14) | (4u << 29);       // SBO = 512 B, version 1, SWIZZLE_64B
__device__ __forceinline__ uint32_t local_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t local_desc(uint32_t lo) { return (static_cast<uint64_t>(LOCAL_DESC_HI) << 32) | lo; }

template <int NST, int TRACE>
__global__ void __launch_bounds__(LOCAL_THREADS, 1)
local_attention_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int B, int H,
                          int L, int NL, float scale_log2e, int reverse) {
  const int nw = L / WIN;
  const int total = nw * B * NL, nlb = NL * B;
  // warp index through a shuffle: provably warp-uniform, so everything derived from it (stream, role, shared-memory and
  // tensor-memory addresses, descriptors) lives in uniform registers and the single-thread tcgen05 instructions need no
  // per-lane "waterfall" loop (ELECT / R2UR.BROADCAST / BRA.U.ANY around every UTCHMMA: ~15 dependent instructions each)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int stream = warp < 6 ? warp / 3 : (warp - 6) >> 2;
  const int role = warp < 6 ? warp % 3 : 3;              // 0 producer, 1 S issuer, 2 PV issuer, 3 softmax

  extern __shared__ uint8_t local_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(local_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + stream * LocalCfg<NST>::STREAM_TILES * TC_TILE;     // 2 tiles
  uint8_t* sKV = sQ + 2 * TC_TILE;                                      // NST x (K tile, V tile)
  __shared__ LocalBars<NST> bars[2];
  __shared__ uint32_t tmem_slot;
  LocalBars<NST>& bar = bars[stream];

  if (tid == 0) {
    ptx::tma_prefetch_desc(&tm_qkv);
    for (int s2 = 0; s2 < 2; ++s2) {
      LocalBars<NST>& x = bars[s2];
      for (int k = 0; k < 2; ++k) {
        ptx::mbar_init(&x.q_full[k], 1);
        ptx::mbar_init(&x.q_free[k], 1);
        ptx::mbar_init(&x.o_full[k], 1);
        ptx::mbar_init(&x.o_free[k], 4);
      }
      for (int k = 0; k < NST; ++k) {
        ptx::mbar_init(&x.kv_full[k], 1);
        ptx::mbar_init(&x.kv_free[k], 1);
      }
      for (int k = 0; k < 3; ++k) {
        ptx::mbar_init(&x.s_full[k], 1);
        ptx::mbar_init(&x.p_ready[k], 4);
        ptx::mbar_init(&x.s_free[k], 1);
      }
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0) + stream * 256;
  ptx::pdl_sync();

  // the n-th item of this stream: flat index blockIdx.x + (stream + 2 n) gridDim.x, walked window-major
  auto item = [&](int n, LocalItem& it) -> bool {
    const int f = int(blockIdx.x) + (stream + 2 * n) * int(gridDim.x);
    if (f >= total) return false;
    const int ff = reverse ? total - 1 - f : f;
    it.w = ff / nlb;
    const int r = ff - it.w * nlb;
    it.b = r / NL;
    it.h = r - it.b * NL;
    it.w_lo = max(it.w - 1, 0);
    it.nkt = min(it.w + 1, nw - 1) - it.w_lo + 1;
    return true;
  };

  if (role == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, one elected lane issues)
    {
      const int plane = B * H * L;
      LocalItem it;
      uint32_t st = 0, st_ph = 0;
      for (int n = 0; item(n, it); ++n) {
        const int rq = (it.b * H + it.h) * L;
        const int qb = n & 1;
        ptx::mbar_wait_parked(&bar.q_free[qb], ((n >> 1) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&bar.q_full[qb], TC_TILE);
          ptx::tma_load_2d(sQ + qb * TC_TILE, &tm_qkv, &bar.q_full[qb], 0, rq + it.w * WIN);
        }
        __syncwarp();
        for (int kt = 0; kt < it.nkt; ++kt) {
          ptx::mbar_wait_parked(&bar.kv_free[st], st_ph ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&bar.kv_full[st], 2 * TC_TILE);
            ptx::tma_load_2d(sKV + (2 * st) * TC_TILE, &tm_qkv, &bar.kv_full[st], 0, plane + rq + (it.w_lo + kt) * WIN);
            ptx::tma_load_2d(sKV + (2 * st + 1) * TC_TILE, &tm_qkv, &bar.kv_full[st], 0, 2 * plane + rq + (it.w_lo + kt) * WIN);
          }
          __syncwarp();
          if (++st == NST) { st = 0; st_ph ^= 1; }
        }
      }
    }
  } else if (role == 1) {
    // ------------------------------------------------------------------ S issuer (the whole warp walks the loop, one elected lane issues)
    {
      constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t q_lo0 = local_desc_lo(ptx::smem_u32(sQ)), kv_lo0 = local_desc_lo(ptx::smem_u32(sKV));
      LocalItem it;
      uint32_t st = 0, st_ph = 0, slot = 0, slot_ph = 0;
      int g = 0;
      for (int n = 0; item(n, it); ++n) {
        const int qb = n & 1;
        ptx::mbar_wait_parked(&bar.q_full[qb], (n >> 1) & 1);
        const uint32_t q_lo = q_lo0 + qb * (TC_TILE >> 4);
        for (int kt = 0; kt < it.nkt; ++kt) {
          ptx::mbar_wait_parked(&bar.kv_full[st], st_ph);
          const uint32_t k_lo = kv_lo0 + st * (2 * TC_TILE >> 4);
#pragma unroll
          for (int sb = 0; sb < 2; ++sb) {
            if (TRACE && blockIdx.x == 0 && g < 128 && lane == 0) g_ms_trace[stream][0][g][0] = clock64();
            ptx::mbar_wait_parked(&bar.s_free[slot], slot_ph ^ 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              ptx::umma_bf16(tmem + slot * 64, local_desc(q_lo), local_desc(k_lo + sb * 256), IDESC_S, 0);
              ptx::umma_bf16(tmem + slot * 64, local_desc(q_lo + 2), local_desc(k_lo + sb * 256 + 2), IDESC_S, 1);
              ptx::umma_commit(&bar.s_full[slot]);
            }
            __syncwarp();
            if (TRACE && blockIdx.x == 0 && g < 128 && lane == 0) g_ms_trace[stream][0][g][1] = clock64();
            if (++slot == 3) { slot = 0; slot_ph ^= 1; }
            ++g;
          }
          if (++st == NST) { st = 0; st_ph ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma_commit(&bar.q_free[qb]);
        __syncwarp();
      }
    }
  } else if (role == 2) {
    // ------------------------------------------------------------------ PV issuer (same pattern)
    {
      constexpr uint32_t IDESC_O = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t kv_lo0 = local_desc_lo(ptx::smem_u32(sKV));
      LocalItem it;
      uint32_t st = 0, st_ph = 0, slot = 0, slot_ph = 0;
      int g = 0;
      for (int n = 0; item(n, it); ++n) {
        const int ob = n & 1;
        ptx::mbar_wait_parked(&bar.o_free[ob], ((n >> 1) & 1) ^ 1);
        const uint32_t t_o = tmem + 192 + ob * 32;
        for (int kt = 0; kt < it.nkt; ++kt) {
          ptx::mbar_wait_parked(&bar.kv_full[st], st_ph);               // V of this tile has landed (visible to this thread)
          const uint32_t v_lo = kv_lo0 + st * (2 * TC_TILE >> 4) + (TC_TILE >> 4);
#pragma unroll
          for (int sb = 0; sb < 2; ++sb) {
            if (TRACE && blockIdx.x == 0 && g < 128 && lane == 0) g_ms_trace[stream][0][g][2] = clock64();
            ptx::mbar_wait_parked(&bar.p_ready[slot], slot_ph);
            if (TRACE && blockIdx.x == 0 && g < 128 && lane == 0) g_ms_trace[stream][0][g][3] = clock64();
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::umma_bf16_ts(t_o, tmem + slot * 64 + ks * 8, local_desc(v_lo + sb * 256 + ks * 64), IDESC_O,
                                  (kt | sb | ks) != 0);
              ptx::umma_commit(&bar.s_free[slot]);
              if (sb == 1) ptx::umma_commit(&bar.kv_free[st]);
              if (sb == 1 && kt == it.nkt - 1) ptx::umma_commit(&bar.o_full[ob]);
            }
            __syncwarp();
            if (TRACE && blockIdx.x == 0 && g < 128 && lane == 0) g_ms_trace[stream][0][g][4] = clock64();
            if (++slot == 3) { slot = 0; slot_ph ^= 1; }
            ++g;
          }
          if (++st == NST) { st = 0; st_ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output (4 warps per stream)
    // Online softmax per 64-key block against a lazily updated reference maximum: the first block of an item sets m_ref
    // to its row maximum; a later block rescales O and the row sum only when its maximum exceeds m_ref by more than 2^8
    // (P stays <= 256, exact in bf16's range), which real score distributions almost never do.  The output of item n is
    // read after the first block of item n + 1 (O is double buffered), when its last PV has long completed.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t(quarter) * 32u) << 16);
    const int D = H * DH;
    constexpr float LAZY_LOG2 = 8.0f;
    const bool tr = TRACE && blockIdx.x == 0 && quarter == 0 && lane == 0;
    float m_ref = 0.f, rs = 0.f;
    bool pend = false;
    int p_n = 0;
    float p_inv = 0.f;
    __nv_bfloat16* p_dst = nullptr;
    auto epilogue = [&]() {
      const int ob = p_n & 1;
      ptx::mbar_wait(&bar.o_full[ob], (p_n >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t ro[32];
      ptx::tmem_ld_32x32(lane_base + 192 + ob * 32, ro);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.o_free[ob]);
      uint4* dst = reinterpret_cast<uint4*>(p_dst);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        dst[k] = make_uint4(ptx::pack_bf16x2(__uint_as_float(ro[8 * k]) * p_inv, __uint_as_float(ro[8 * k + 1]) * p_inv),
                            ptx::pack_bf16x2(__uint_as_float(ro[8 * k + 2]) * p_inv, __uint_as_float(ro[8 * k + 3]) * p_inv),
                            ptx::pack_bf16x2(__uint_as_float(ro[8 * k + 4]) * p_inv, __uint_as_float(ro[8 * k + 5]) * p_inv),
                            ptx::pack_bf16x2(__uint_as_float(ro[8 * k + 6]) * p_inv, __uint_as_float(ro[8 * k + 7]) * p_inv));
    };
    LocalItem it;
    uint32_t slot = 0, slot_ph = 0;
    int g = 0;
    for (int n = 0; item(n, it); ++n) {
      const int nblk = 2 * it.nkt;
      for (int blk = 0; blk < nblk; ++blk) {
        const uint32_t t_s = lane_base + slot * 64;
        if (tr && g < 128) g_ms_trace[stream][1][g][0] = clock64();
        ptx::mbar_wait(&bar.s_full[slot], slot_ph);
        if (tr && g < 128) g_ms_trace[stream][1][g][1] = clock64();
        ptx::tc_fence_after();
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(t_s, r0);
        ptx::tmem_ld_32x32(t_s + 32, r1);
        ptx::tmem_ld_wait();
        if (tr && g < 128) g_ms_trace[stream][1][g][2] = clock64();
        float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          b0 = fmaxf(b0, __uint_as_float(r0[2 * k]));
          b1 = fmaxf(b1, __uint_as_float(r0[2 * k + 1]));
          b2 = fmaxf(b2, __uint_as_float(r1[2 * k]));
          b3 = fmaxf(b3, __uint_as_float(r1[2 * k + 1]));
        }
        const float bm = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
        if (blk == 0) {
          m_ref = bm;
          rs = 0.f;
        } else {
          const bool need = (bm - m_ref) * scale_log2e > LAZY_LOG2;
          if (__any_sync(0xffffffffu, need)) {
            // every PV issued so far must have landed in O: PV(g - 1)'s commit completes s_free of its slot
            const uint32_t pslot = slot == 0 ? 2 : slot - 1, pph = slot == 0 ? slot_ph ^ 1 : slot_ph;
            ptx::mbar_wait(&bar.s_free[pslot], pph);
            ptx::tc_fence_after();
            const float f = need ? fast_ex2((m_ref - bm) * scale_log2e) : 1.f;
            uint32_t ro[32];
            const uint32_t t_o = lane_base + 192 + (n & 1) * 32;
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
        const float ms = m_ref * scale_log2e;
        // With the issuers out of the way this loop is what bounds the kernel: the exponentials of a block keep the MUFU
        // pipe busy for 512 cycles (64 warp instructions at 4 lanes per clock per SMSP), the rest of a block (barrier,
        // tensor-memory load and store, maximum) is ~370 cycles, and the two warps sharing an SMSP drift into lockstep
        // (~1500 cycles per block pair, MUFU 2/3 busy; profiles/r02_attn_trace_v5_uniform_issue.log).  Measured and not
        // kept: taking turns on the pipe through a shared-memory hint (+5 %), a cubic exp2 on the FMA pipe for every 4th
        // or 8th element (+4 % / +8 %: with two warps per SMSP the issue slots, not MUFU, are what the extra
        // instructions cost), four streams of 32-key blocks (+8 %); profiles/r02_attn_variants_v5.log.  Strict alternation of
        // the two warps' exponential loops through a pair of named barriers per quarter (bar.sync / bar.arrive token, the
        // MUFU-free part of one warp under the other's loop): +5 % (88.8 -> 93.7 us isolated, 1.56 -> 1.70 ms per step;
        // profiles/r02_attn_pingpong_rejected.log): one warp alone does not keep the pipe full.  Dropping the block maximum
        // after an item's first block (exponentials against the first block's reference, floating point keeps the rest):
        // 77 us instead of 85 with no overflow check at all, but every correct form of the check (a vote on the block's row
        // sum after this loop; a score bound from q / k row norms written by the QKV epilogue) gave the gain back in the
        // step (profiles/r02_attn_nomax_variants.log).
        uint32_t pk[32];                                   // P (bf16 pairs) over the first 32 of the slot's 64 columns
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float p0 = fast_ex2(fmaf(__uint_as_float(r0[2 * k]), scale_log2e, -ms));
          const float p1 = fast_ex2(fmaf(__uint_as_float(r0[2 * k + 1]), scale_log2e, -ms));
          const float p2 = fast_ex2(fmaf(__uint_as_float(r1[2 * k]), scale_log2e, -ms));
          const float p3 = fast_ex2(fmaf(__uint_as_float(r1[2 * k + 1]), scale_log2e, -ms));
          s0 += p0; s1 += p1; s2 += p2; s3 += p3;
          pk[k] = ptx::pack_bf16x2(p0, p1);
          pk[16 + k] = ptx::pack_bf16x2(p2, p3);
        }
        rs += (s0 + s1) + (s2 + s3);
        if (tr && g < 128) g_ms_trace[stream][1][g][3] = clock64() + (__float_as_int(rs) & 0);
        ptx::tmem_st_32x32(t_s, pk);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.p_ready[slot]);
        if (tr && g < 128) g_ms_trace[stream][1][g][4] = clock64();
        if (blk == 0 && pend) {
          epilogue();
          pend = false;
        }
        if (tr && g < 128) g_ms_trace[stream][1][g][5] = clock64();
        if (++slot == 3) { slot = 0; slot_ph ^= 1; }
        ++g;
      }
      pend = true;
      p_n = n;
      p_inv = 1.f / rs;
      p_dst = out + (size_t(it.b) * L + size_t(it.w) * WIN + row) * D + it.h * DH;
    }
    if (pend) epilogue();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_slot, 512);
}

// ================================================================================================
// The same attention with THREE streams per CTA (round 2, the default; BIOM3_ATTN3=0 selects the kernel above, which also
// carries the clock64 trace hooks): three softmax warps per SM sub-partition instead of two, so that the MUFU pipe has a third warp's exponentials to run while the other two are in
// the MUFU-free part of their blocks (the two-stream kernel keeps it 2/3 busy).  Four warps per SMSP cap a thread at 128
// registers, which is what the softmax warps need, so the CTA has exactly 16 warps:
//   warp 0        TMA producer for all three streams (non-blocking: serves whichever stream has a free buffer)
//   warps 1..3    one issuer per stream: S one block ahead of P V (two S / P slots), both from the same thread
//   warps 4..15   softmax, four per stream (warp id % 4 = TMEM lane quarter)
// Per stream: TMEM 160 columns (S / P slots at 0 and 64, one O buffer at 128), shared memory 2 Q tiles + 3 K/V stages.
// With a single O buffer the output of an item is read right after its last block (the wait for the last P V is covered
// by the other two streams).  81.5 us against 87.7 us isolated at B = 64, 1.43 against 1.59 ms per step; the three warps of
// an SMSP still settle into lockstep (an initial stagger of the streams changes nothing), two K/V stages instead of three:
// +1 %.  A four-stream form (20 warps, 96 registers: the softmax in two halves of 32 keys with the scores read from TMEM
// twice, one S / P slot per stream, one controller warp per stream polling with mbarrier.test_wait) ran at 95 us.
// ================================================================================================
constexpr int L3_NST = 3;
constexpr int L3_THREADS = 16 * 32;
constexpr int L3_STREAM_TILES = 2 + 2 * L3_NST;
constexpr int L3_SMEM_BYTES = 3 * L3_STREAM_TILES * TC_TILE + 1024;
struct L3Bars {
  uint64_t q_full[2], q_free[2], kv_full[L3_NST], kv_free[L3_NST], s_full[2], p_ready[2], s_free[2], o_full, o_free;
};
struct L3Cursor {                       // position of an issuer in its stream's block sequence
  int n, kt, sb, g;
  uint32_t st, st_ph;
  bool valid;
  LocalItem it;
};

__global__ void __launch_bounds__(L3_THREADS, 1)
local_attention3_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int B, int H, int L,
                        int NL, float scale_log2e, int reverse) {
  const int nw = L / WIN;
  const int total = nw * B * NL, nlb = NL * B;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  extern __shared__ uint8_t local3_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(local3_raw) + 1023) & ~uintptr_t(1023));
  __shared__ L3Bars bars[3];
  __shared__ uint32_t tmem_slot;
  if (tid == 0) {
    ptx::tma_prefetch_desc(&tm_qkv);
    for (int s2 = 0; s2 < 3; ++s2) {
      L3Bars& x = bars[s2];
      for (int k = 0; k < 2; ++k) {
        ptx::mbar_init(&x.q_full[k], 1);
        ptx::mbar_init(&x.q_free[k], 1);
        ptx::mbar_init(&x.s_full[k], 1);
        ptx::mbar_init(&x.p_ready[k], 4);
        ptx::mbar_init(&x.s_free[k], 1);
      }
      for (int k = 0; k < L3_NST; ++k) {
        ptx::mbar_init(&x.kv_full[k], 1);
        ptx::mbar_init(&x.kv_free[k], 1);
      }
      ptx::mbar_init(&x.o_full, 1);
      ptx::mbar_init(&x.o_free, 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_slot, 0);
  ptx::pdl_sync();

  // the n-th item of stream s: flat index blockIdx.x + (s + 3 n) gridDim.x, walked window-major
  auto item = [&](int s, int n, LocalItem& it) -> bool {
    const int f = int(blockIdx.x) + (s + 3 * n) * int(gridDim.x);
    if (f >= total) return false;
    const int ff = reverse ? total - 1 - f : f;
    it.w = ff / nlb;
    const int r = ff - it.w * nlb;
    it.b = r / NL;
    it.h = r - it.b * NL;
    it.w_lo = max(it.w - 1, 0);
    it.nkt = min(it.w + 1, nw - 1) - it.w_lo + 1;
    return true;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer, three streams, non-blocking
    const int plane = B * H * L;
    int n[3] = {0, 0, 0}, kt[3] = {-1, -1, -1};
    uint32_t st[3] = {0, 0, 0}, st_ph[3] = {0, 0, 0};
    bool live[3];
    LocalItem it[3];
    int left = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      live[s] = item(s, 0, it[s]);
      left += live[s] ? 1 : 0;
    }
    while (left > 0) {
      bool progress = false;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (!live[s]) continue;
        L3Bars& bar = bars[s];
        uint8_t* sQ = smem + s * L3_STREAM_TILES * TC_TILE;
        uint8_t* sKV = sQ + 2 * TC_TILE;
        const int rq = (it[s].b * H + it[s].h) * L;
        if (kt[s] < 0) {
          const int qb = n[s] & 1;
          if (!ptx::mbar_test_wait(&bar.q_free[qb], ((n[s] >> 1) & 1) ^ 1)) continue;
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&bar.q_full[qb], TC_TILE);
            ptx::tma_load_2d(sQ + qb * TC_TILE, &tm_qkv, &bar.q_full[qb], 0, rq + it[s].w * WIN);
          }
          __syncwarp();
          kt[s] = 0;
          progress = true;
        } else {
          if (!ptx::mbar_test_wait(&bar.kv_free[st[s]], st_ph[s] ^ 1)) continue;
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&bar.kv_full[st[s]], 2 * TC_TILE);
            ptx::tma_load_2d(sKV + (2 * st[s]) * TC_TILE, &tm_qkv, &bar.kv_full[st[s]], 0, plane + rq + (it[s].w_lo + kt[s]) * WIN);
            ptx::tma_load_2d(sKV + (2 * st[s] + 1) * TC_TILE, &tm_qkv, &bar.kv_full[st[s]], 0, 2 * plane + rq + (it[s].w_lo + kt[s]) * WIN);
          }
          __syncwarp();
          if (++st[s] == L3_NST) { st[s] = 0; st_ph[s] ^= 1; }
          progress = true;
          if (++kt[s] == it[s].nkt) {
            kt[s] = -1;
            ++n[s];
            live[s] = item(s, n[s], it[s]);
            if (!live[s]) --left;
          }
        }
      }
      if (!progress) __nanosleep(64);
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ issuer of stream warp - 1: S one block ahead of P V
    const int s = warp - 1;
    L3Bars& bar = bars[s];
    const uint32_t tmem = tmem0 + s * 160;
    uint8_t* sQ = smem + s * L3_STREAM_TILES * TC_TILE;
    uint8_t* sKV = sQ + 2 * TC_TILE;
    constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t IDESC_O = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t q_lo0 = local_desc_lo(ptx::smem_u32(sQ)), kv_lo0 = local_desc_lo(ptx::smem_u32(sKV));
    auto start = [&](L3Cursor& c) {
      c.n = 0; c.kt = 0; c.sb = 0; c.g = 0; c.st = 0; c.st_ph = 0;
      c.valid = item(s, 0, c.it);
    };
    auto advance = [&](L3Cursor& c) {
      ++c.g;
      if (c.sb == 0) { c.sb = 1; return; }
      c.sb = 0;
      if (++c.st == L3_NST) { c.st = 0; c.st_ph ^= 1; }
      if (++c.kt == c.it.nkt) {
        c.kt = 0;
        ++c.n;
        c.valid = item(s, c.n, c.it);
      }
    };
    auto issue_s = [&](const L3Cursor& c) {
      const int qb = c.n & 1;
      const uint32_t slot = c.g & 1;
      if (c.kt == 0 && c.sb == 0) ptx::mbar_wait_parked(&bar.q_full[qb], (c.n >> 1) & 1);
      if (c.sb == 0) ptx::mbar_wait_parked(&bar.kv_full[c.st], c.st_ph);
      ptx::mbar_wait_parked(&bar.s_free[slot], ((c.g >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t q_lo = q_lo0 + qb * (TC_TILE >> 4);
      const uint32_t k_lo = kv_lo0 + c.st * (2 * TC_TILE >> 4);
      const bool last_s = c.kt == c.it.nkt - 1 && c.sb == 1;
      if (ptx::elect_one()) {
        ptx::umma_bf16(tmem + slot * 64, local_desc(q_lo), local_desc(k_lo + c.sb * 256), IDESC_S, 0);
        ptx::umma_bf16(tmem + slot * 64, local_desc(q_lo + 2), local_desc(k_lo + c.sb * 256 + 2), IDESC_S, 1);
        ptx::umma_commit(&bar.s_full[slot]);
        if (last_s) ptx::umma_commit(&bar.q_free[qb]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](const L3Cursor& c) {
      const uint32_t slot = c.g & 1;
      const bool first = c.kt == 0 && c.sb == 0, last = c.kt == c.it.nkt - 1 && c.sb == 1;
      if (first) ptx::mbar_wait_parked(&bar.o_free, (c.n & 1) ^ 1);       // the previous item's output has been read
      ptx::mbar_wait_parked(&bar.p_ready[slot], (c.g >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t v_lo = kv_lo0 + c.st * (2 * TC_TILE >> 4) + (TC_TILE >> 4);
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          ptx::umma_bf16_ts(tmem + 128, tmem + slot * 64 + ks * 8, local_desc(v_lo + c.sb * 256 + ks * 64), IDESC_O, (first ? ks : 1) != 0);
        ptx::umma_commit(&bar.s_free[slot]);
        if (c.sb == 1) ptx::umma_commit(&bar.kv_free[c.st]);
        if (last) ptx::umma_commit(&bar.o_full);
      }
      __syncwarp();
    };
    L3Cursor cs, cp;
    start(cs);
    start(cp);
    if (cs.valid) {
      issue_s(cs);
      advance(cs);
    }
    while (cp.valid) {
      if (cs.valid) {
        issue_s(cs);
        advance(cs);
      }
      issue_pv(cp);
      advance(cp);
    }
  } else {
    // ------------------------------------------------------------------ softmax + output: stream (warp - 4) / 4, quarter warp % 4
    const int s = (warp - 4) >> 2;
    L3Bars& bar = bars[s];
    const uint32_t tmem = tmem0 + s * 160;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t(quarter) * 32u) << 16);
    const int D = H * DH;
    constexpr float LAZY_LOG2 = 8.0f;
    float m_ref = 0.f, rs = 0.f;
    LocalItem it;
    int g = 0;
    for (int n = 0; item(s, n, it); ++n) {
      const int nblk = 2 * it.nkt;
      for (int blk = 0; blk < nblk; ++blk, ++g) {
        const uint32_t slot = g & 1, slot_ph = (g >> 1) & 1;
        const uint32_t t_s = lane_base + slot * 64;
        ptx::mbar_wait(&bar.s_full[slot], slot_ph);
        ptx::tc_fence_after();
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(t_s, r0);
        ptx::tmem_ld_32x32(t_s + 32, r1);
        ptx::tmem_ld_wait();
        float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          b0 = fmaxf(b0, __uint_as_float(r0[2 * k]));
          b1 = fmaxf(b1, __uint_as_float(r0[2 * k + 1]));
          b2 = fmaxf(b2, __uint_as_float(r1[2 * k]));
          b3 = fmaxf(b3, __uint_as_float(r1[2 * k + 1]));
        }
        const float bm = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
        if (blk == 0) {
          m_ref = bm;
          rs = 0.f;
        } else {
          const bool need = (bm - m_ref) * scale_log2e > LAZY_LOG2;
          if (__any_sync(0xffffffffu, need)) {
            // every P V issued so far must have landed in O: P V (g - 1)'s commit completes s_free of its slot
            ptx::mbar_wait(&bar.s_free[slot ^ 1], ((g - 1) >> 1) & 1);
            ptx::tc_fence_after();
            const float f = need ? fast_ex2((m_ref - bm) * scale_log2e) : 1.f;
            uint32_t ro[32];
            ptx::tmem_ld_32x32(lane_base + 128, ro);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) ro[k] = __float_as_uint(__uint_as_float(ro[k]) * f);
            ptx::tmem_st_32x32(lane_base + 128, ro);
            ptx::tmem_st_wait();
            rs *= f;
            if (need) m_ref = bm;
          }
        }
        const float ms = m_ref * scale_log2e;
        uint32_t pk[32];
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float p0 = fast_ex2(fmaf(__uint_as_float(r0[2 * k]), scale_log2e, -ms));
          const float p1 = fast_ex2(fmaf(__uint_as_float(r0[2 * k + 1]), scale_log2e, -ms));
          const float p2 = fast_ex2(fmaf(__uint_as_float(r1[2 * k]), scale_log2e, -ms));
          const float p3 = fast_ex2(fmaf(__uint_as_float(r1[2 * k + 1]), scale_log2e, -ms));
          s0 += p0; s1 += p1; s2 += p2; s3 += p3;
          pk[k] = ptx::pack_bf16x2(p0, p1);
          pk[16 + k] = ptx::pack_bf16x2(p2, p3);
        }
        rs += (s0 + s1) + (s2 + s3);
        ptx::tmem_st_32x32(t_s, pk);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.p_ready[slot]);
      }
      // output of this item (single O buffer): the last P V has been issued as soon as the fourth warp arrived above
      ptx::mbar_wait(&bar.o_full, n & 1);
      ptx::tc_fence_after();
      uint32_t ro[32];
      ptx::tmem_ld_32x32(lane_base + 128, ro);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.o_free);
      const float inv = 1.f / rs;
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t(it.b) * L + size_t(it.w) * WIN + row) * D + it.h * DH);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        dst[k] = make_uint4(ptx::pack_bf16x2(__uint_as_float(ro[8 * k]) * inv, __uint_as_float(ro[8 * k + 1]) * inv),
                            ptx::pack_bf16x2(__uint_as_float(ro[8 * k + 2]) * inv, __uint_as_float(ro[8 * k + 3]) * inv),
                            ptx::pack_bf16x2(__uint_as_float(ro[8 * k + 4]) * inv, __uint_as_float(ro[8 * k + 5]) * inv),
                            ptx::pack_bf16x2(__uint_as_float(ro[8 * k + 6]) * inv, __uint_as_float(ro[8 * k + 7]) * inv));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem0, 512);
}

}  // namespace attn
