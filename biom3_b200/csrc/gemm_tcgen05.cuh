// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] . W[N,K]^T  (+ fused epilogue)
//
//   warp 0        TMA producer: cp.async.bulk.tensor (128B swizzle) into a STAGES-deep smem ring
//   warp 1        tcgen05.mma issuer (one elected thread), fp32 accumulators in TMEM, double buffered
//   warps 2..     epilogue: tcgen05.ld TMEM -> registers -> fused bias / GELU / residual -> global
//
// Tiles are 128 x BN x 64; one CTA per SM walks tiles n-fastest so the A tile is re-read from L2.
// The reference runs these contractions as separate cuBLAS calls plus element-wise kernels
// (to_q/to_k/to_v, to_out, w1 + GELU, w2 in linear-attention-transformer, called from
// /root/reference/Stage3_source/cond_diff_transformer_layer.py:171).
#pragma once
#include "ptx.cuh"

namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;

enum Epi : int {
  EPI_STORE_BF16 = 0,      // C bf16 row-major [M, N]
  EPI_QKV_HEADMAJOR = 1,   // C bf16 as [3][B][H][L][32]  (column n = which*D + h*32 + d)
  EPI_BIAS_GELU_BF16 = 2,  // C bf16 row-major, gelu_erf(acc + bias[n])
  EPI_BIAS_RESID_F32 = 3,  // R fp32 row-major [M, N]: R += acc + bias[n] (+ cond[b(m)][n]), in place
  EPI_STORE_F32 = 4,       // C fp32 row-major (unit tests)
};

struct Params {
  int M, N, K;
  int b_row_offset;        // first row of this layer's weight inside the stacked weight tensor map
  void* out;               // bf16 or fp32, see Epi
  const float* bias;       // [N] or nullptr
  const float* cond;       // [B][cond_stride] fp32 or nullptr (added per sample, EPI_BIAS_RESID_F32)
  int cond_stride;
  int L;                   // tokens per sample (rows per batch entry)
  int H;                   // heads (EPI_QKV_HEADMAJOR)
  int Bsz;                 // batch (EPI_QKV_HEADMAJOR)
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024;   // + alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int BN, int STAGES, int EPI, int EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const Params p) {
  static_assert(BN == 128 || BN == 256, "BN");
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "epilogue warps");
  using SL = SmemLayout<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = 2 * BN;           // two accumulator stages
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(BM, BN);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int n_tiles = p.N / BN;
  const int num_tiles = (p.M / BM) * n_tiles;
  const int k_blocks = p.K / BK;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&tmap_a);
    ptx::tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN + p.b_row_offset;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SL::STAGE_BYTES;
          uint8_t* sb = sa + SL::A_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], SL::STAGE_BYTES);
          ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
          ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        ptx::mbar_wait(&acc_empty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint64_t da = ptx::umma_desc_sw128(sa);
          const uint64_t db = ptx::umma_desc_sw128(sa + SL::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
            ptx::umma_bf16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), IDESC, (kb | k) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);            // smem slot free once these MMAs retire
          if (kb == k_blocks - 1) ptx::umma_commit(&acc_full[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const uint32_t ew = warp - 2;
    const uint32_t quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter ..
    constexpr int COL_SPLIT = EPI_WARPS / 4;           // 1 or 2 warps share a lane quarter
    const uint32_t col_half = (COL_SPLIT == 2) ? (ew >> 2) : 0;
    constexpr int COLS_PER_WARP = BN / COL_SPLIT;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      ptx::mbar_wait(&acc_full[as], aphase);
      ptx::tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const int bidx = row / p.L;
      const uint32_t t_row = tmem_base + ((quarter * 32u) << 16) + as * BN + col_half * COLS_PER_WARP;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_row + c, r);
        ptx::tmem_ld_wait();
        const int n = n0 + col_half * COLS_PER_WARP + c;   // first of 32 consecutive output columns
        if constexpr (EPI == EPI_STORE_F32) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + size_t(row) * p.N + n);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                 __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + size_t(row) * p.N + n);
          const float4* bias = reinterpret_cast<const float4*>(p.bias + n);
          const float4* cond = p.cond ? reinterpret_cast<const float4*>(p.cond + size_t(bidx) * p.cond_stride + n)
                                      : nullptr;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 o = dst[i];
            const float4 b = __ldg(bias + i);
            o.x += __uint_as_float(r[4 * i]) + b.x;
            o.y += __uint_as_float(r[4 * i + 1]) + b.y;
            o.z += __uint_as_float(r[4 * i + 2]) + b.z;
            o.w += __uint_as_float(r[4 * i + 3]) + b.w;
            if (cond) {
              const float4 cv = __ldg(cond + i);
              o.x += cv.x; o.y += cv.y; o.z += cv.z; o.w += cv.w;
            }
            dst[i] = o;
          }
        } else {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          __nv_bfloat16* dst;
          if constexpr (EPI == EPI_QKV_HEADMAJOR) {
            const int D = p.N / 3;
            const int which = n / D, h = (n % D) >> 5, l = row % p.L;
            dst = reinterpret_cast<__nv_bfloat16*>(p.out) +
                  ((((size_t(which) * p.Bsz + bidx) * p.H + h) * p.L + l) << 5);
          } else {
            dst = reinterpret_cast<__nv_bfloat16*>(p.out) + size_t(row) * p.N + n;
          }
          if constexpr (EPI == EPI_BIAS_GELU_BF16) {
            const float4* bias = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(bias + i);
              v[4 * i] = gelu_erf(v[4 * i] + b.x);
              v[4 * i + 1] = gelu_erf(v[4 * i + 1] + b.y);
              v[4 * i + 2] = gelu_erf(v[4 * i + 2] + b.z);
              v[4 * i + 3] = gelu_erf(v[4 * i + 3] + b.w);
            }
          }
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            d4[i] = make_uint4(ptx::pack_bf16x2(v[8 * i], v[8 * i + 1]), ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                               ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        }
      }
      // accumulator stage drained: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace gemm
