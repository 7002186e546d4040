// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] . W[N,K]^T  (+ fused epilogue)
//
//   warp 0        TMA producer: cp.async.bulk.tensor (128B swizzle) into a STAGES-deep smem ring
//   warp 1        tcgen05.mma issuer (one elected thread), fp32 accumulators in TMEM, double buffered
//   warps 2..9    epilogue: tcgen05.ld TMEM -> registers (thread = row) -> fused math -> per-warp smem
//                 transpose (XOR swizzled, conflict free) -> fully coalesced global loads / stores
//
// Two tilings of the same kernel:
//   CG2 = false   one CTA per 128 x BN tile (cta_group::1).  Any M % 128 == 0.
//   CG2 = true    a CTA pair (cluster of 2, cta_group::2) per 256 x BN tile: each CTA stages its own 128 rows
//                 of A and HALF of the weight tile, the leader issues 256-row MMAs that read both CTAs' shared
//                 memory, each CTA keeps its 128 x BN accumulator in its own TMEM.  The mainloop of the 1-CTA
//                 version is bound by the SM's L2->smem ingest (~43 B/clk measured, 48 KB per 512 MMA cycles);
//                 the pair needs 32 KB per CTA for the same MMA work.
//
// Round-1 experiment, removed: an A-resident variant (the 256-row A block loaded once per run of column tiles, only W
// through the ring; halves the L2 -> SM operand traffic, bit-identical) measured no faster, before and after the
// single-thread issue path was fixed (profiles/r01_ab_ares.jsonl, profiles/r02_gemm_trace_uniform_issue.log).
//
// Fusions (the reference runs each as separate library / element-wise kernels; block structure from
// linear-attention-transformer, called at /root/reference/Stage3_source/cond_diff_transformer_layer.py:171):
//   * pre-norm LayerNorm is folded into the consumer GEMM:  LN(u) W^T = rstd (u (g*W)^T - mean * sum_k g_k W_nk)
//     + sum_k b_k W_nk.  The producer epilogue emits a raw bf16 copy of the residual stream plus per-row
//     partial (sum, sum of squares); the consumer epilogue applies mean / rstd per row.  No LayerNorm kernel.
//   * bias, erf-GELU, fp32 residual update and the next layer's additive conditioning vector.
#pragma once
#include "ptx.cuh"

namespace gemm {

constexpr int BM = 128;          // rows per CTA
constexpr int BK = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
// epilogue staging: one 32 x 32 block per warp (fp32: 4 KB, bf16: 2 KB), see epi_stg_bytes()
constexpr int CV_TOTAL = 16384;   // the two per-column epilogue vectors of every warp's column range, double buffered
                                  // (the next tile's vectors are fetched with cp.async while this tile is processed)

// Epilogue warps per CTA.  The fp32 residual epilogues use 8 (two per TMEM lane quarter, 128 columns each, 168
// registers for the residual prefetch).  The bf16 epilogues are latency-bound chains (TMEM load -> fused math -> smem
// transpose -> store): the GELU epilogue (FF1) uses 16 warps (four per quarter, 64 columns each) so the chains of different
// warps overlap; the QKV epilogue has almost no math per element, its cost is the ~244 instructions of per-tile set-up each
// warp executes, and it is 6 % faster with 8 warps of four chunks (A/B of compile-time variants, `tools/ab_variant.sh`:
// QKV 1.80 -> 1.69 ms per step with 8 warps, FF1 2.56 -> 2.69 ms, hence the mixed default).
#ifndef BIOM3_BF16_EPI_WARPS
#define BIOM3_BF16_EPI_WARPS 16      // GELU / plain bf16-store epilogues; compile-time A/B: python -m biom3_b200.build --variant ew8 BIOM3_BF16_EPI_WARPS=8
#endif
#ifndef BIOM3_QKV_EPI_WARPS
#define BIOM3_QKV_EPI_WARPS 8
#endif
__host__ __device__ constexpr int epi_warps(int epi) {
  return (epi >= 3 && epi <= 6) ? 8 : (epi == 1 ? BIOM3_QKV_EPI_WARPS : BIOM3_BF16_EPI_WARPS);   // 7: like 2
}
constexpr uint32_t EPI_WARP0 = 2;     // warp 0: TMA producer, warp 1: MMA issuer, then the epilogue warps

enum Epi : int {
  EPI_STORE_BF16 = 0,      // C bf16 row-major [M, N]
  EPI_QKV_HEADMAJOR = 1,   // C bf16 as [3][B][H][L][32]  (column n = which*D + h*32 + d)
  EPI_BIAS_GELU_BF16 = 2,  // C bf16 row-major, gelu_erf(x + bias[n])
  EPI_BIAS_RESID_F32 = 3,  // R fp32 row-major [M, N]: R += acc + bias[n] (+ cond[b(m)][n]), in place
  EPI_STORE_F32 = 4,       // C fp32 row-major (unit tests)
  EPI_BIAS_RESID_SPLIT = 5, // like 3, but R lives as two bf16 arrays, R = hi + lo (hi = bf16(R), lo = bf16(R - hi)):
                            // `out_bf16` is hi (the array the next GEMM reads as its A operand), `out` is lo.  Same
                            // bytes read, 2 instead of 6 bytes per element written; R keeps 16 significant bits.
  EPI_BIAS_RESID_SPLIT_TMA = 6, // same arithmetic and planes as 5, other data path (round 2): every epilogue warp owns a ring of
                            // RESID_SLOTS shared-memory slots of one 32 x 32 chunk (hi block + lo block, 64-byte swizzle) that TMA
                            // fills two chunks ahead; thread = row straight out of TMEM (no transposition), the slot is updated
                            // in place and handed back to TMA as the store source.  tmap_c / tmap_d = the hi / lo plane.
                            // Epilogue 5 keeps one chunk of residual in registers (16 x 16-byte loads per lane), which with the
                            // ~1.7 us loaded HBM latency of this kernel pinned it at 4.4 TB/s (ncu: 45 % of the stall samples on
                            // the first use of the prefetched registers).
  EPI_BIAS_GELU_SPLIT = 7,  // fp32-class mode FF1: gelu_erf(acc + bias[n]) with the exact erf, stored as the [hi | lo] bf16 split
                            // the next split3 GEMM reads: hi at column n, lo at column N + n of a [M][2N] matrix (tmap_c).
                            // Replaces an fp32 store + a separate bias / GELU / split pass (1 GB of HBM traffic per layer).
};
// (7, below the enum: fp32-class mode)
#ifndef BIOM3_RESID_SLOTS
#define BIOM3_RESID_SLOTS 3
#endif
constexpr int RESID_SLOTS = BIOM3_RESID_SLOTS;
#ifndef BIOM3_RESID_REFILL_LATE
#define BIOM3_RESID_REFILL_LATE 1
#endif
constexpr int RESID_SLOT_BYTES = 4096;   // 32 rows x 64 bytes, hi block then lo block

constexpr int QKV_MAX_CHUNKS = 256;   // 3 * dim / 32 <= 256

struct Params {
  int M, N, K;
  int b_row_offset;        // first row of this layer's weight inside the stacked weight tensor map
  int reverse;             // 1 = walk the tiles last-to-first (consume the previous kernel's freshest, still L2-resident, output first)
  int a_row_offset;        // first row of this launch's A rows inside the A tensor map (row-slab launches)
  void* out;               // bf16 or fp32, see Epi
  const float* bias;       // [N] or nullptr
  const float* cond;       // [B][cond_stride] fp32 or nullptr (added per sample, EPI_BIAS_RESID_F32)
  int cond_stride;
  int L;                   // tokens per sample (rows per batch entry)
  int H;                   // heads (EPI_QKV_HEADMAJOR)
  int Bsz;                 // batch (EPI_QKV_HEADMAJOR)
  // EPI_QKV_HEADMAJOR: what the i-th 32-column chunk of the output is (a chunk of a row IS one head of one token):
  // bits 0-11 head, 12-13 which (0 q, 1 k, 2 v), 14-15 kind.  Kind 1: stored as softmax over the head's 32 features, taken
  // from the fp32 accumulator (q of a linear-attention head).  The host permutes the weight rows so that the chunks with
  // extra work are spread evenly over the column tiles (biom3_model::qkv_chunk); the identity order with kind 0 is the
  // plain head-major store.
  unsigned short qkv_chunk[QKV_MAX_CHUNKS];
  // consumer side of the folded LayerNorm (bf16-output epilogues): x = rstd*(acc - mean*ln_s[n]) + ln_t[n]
  const float* ln_stats;   // [M][ln_parts][2] partial (sum, sumsq) over the K features, or nullptr
  const float* ln_s;       // [N]
  const float* ln_t;       // [N]; for EPI_BIAS_GELU_BF16 the caller stores t / 2 (that epilogue works on x / 2; halving is exact)
  int ln_parts;
  // producer side (EPI_BIAS_RESID_F32): raw bf16 copy of the updated rows + their partial statistics
  __nv_bfloat16* out_bf16; // [M][N] or nullptr
  float* stats_out;        // [M][(N/BN)*2][2] or nullptr
  int split3;              // fp32-class mode: A and W hold [hi | lo] bf16 halves (2K columns each); the K loop runs
                           // hi.hi, hi.lo, lo.hi (3K/BK blocks, fp32 accumulate) = the product to ~2^-16 relative
  int nt_shift, l_shift, d_shift;   // log2 of N / BN, of L and of the QKV head-block width N / 3 when those are powers of two, else -1
                           // (set by fill_shifts(); the per-tile index arithmetic then costs shifts instead of ~25-instruction
                           // integer divisions, which were 17 % of the QKV GEMM's executed instructions)
  int trace;               // test hook: CTA 0 records a clock64 timeline into g_gemm_trace
};

__host__ __device__ constexpr int log2_or_neg(int v) {
  if (v <= 0 || (v & (v - 1)) != 0) return -1;
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}
// call after M / N / K / L are set
inline void fill_shifts(Params& p, int bn) {
  p.nt_shift = log2_or_neg(p.N / bn);
  p.l_shift = log2_or_neg(p.L);
  p.d_shift = (p.N % 3 == 0) ? log2_or_neg(p.N / 3) : -1;
}

// Shared memory of one CTA: [operand ring] [epilogue staging] [per-column epilogue vectors].
// Staging and vector space follow the epilogue (EPI), so that the kernels that need less of them can afford a deeper ring.
__host__ __device__ constexpr bool epi_is_f32(int epi) { return epi >= 3 && epi <= 5; }
__host__ __device__ constexpr int epi_stg_bytes(int epi) {
  return epi == 6 ? epi_warps(epi) * RESID_SLOTS * RESID_SLOT_BYTES : epi_warps(epi) * ((epi_is_f32(epi) || epi == 7) ? 4096 : 2048);
}
// 1: the scale / shift vectors of the bf16 epilogues; 6: bias (+ conditioning) of the warp's 128 columns, double buffered
__host__ __device__ constexpr int epi_cv_bytes(int epi) { return (epi == 1 || epi == 2 || epi == 7) ? CV_TOTAL : (epi == 6 ? 8 * 1024 : 0); }

template <int BN, int STAGES, bool CG2, int EPI>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = CG2 ? BN / 2 : BN;         // weight rows staged by one CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES_TOTAL = epi_stg_bytes(EPI);
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int CV_OFFSET = STG_OFFSET + STG_BYTES_TOTAL;
  static constexpr int TOTAL = CV_OFFSET + epi_cv_bytes(EPI) + 1024;   // + alignment slack
  static_assert(TOTAL <= 232448 - 512, "more than the 227 KB a CTA can have (512 bytes left for the static barriers)");
};

// erf-GELU, x * Phi(x), evaluated as 0.5 x (1 + tanh(x (a + b x^2 + c x^4))): the three coefficients are a
// least-squares fit to the exact erf form (max abs error 3.0e-5 on [-8, 8], below the bf16 rounding of the
// stored result); one MUFU.TANH + 6 FP32 ops instead of erff()'s ~30, so the FF1 epilogue keeps up with the MMAs.
// The argument is h = x / 2 (the caller folds the 1/2 into its scale vectors): with g = h^2 clamped to 16 (|x| <= 8;
// beyond that the tanh saturates and the polynomial, whose x^4 coefficient is negative, must not change sign)
//   gelu = h + h * tanh(h * (2a + 8b g + 32c g^2)).
__device__ __forceinline__ float gelu_erf_half(float h) {
  const float g = fminf(h * h, 16.0f);
  const float poly = fmaf(fmaf(32.0f * -3.58732362e-4f, g, 8.0f * 3.70503451e-2f), g, 2.0f * 7.97458471e-1f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h * poly));
  return fmaf(h, th, h);
}

// 1 / sqrt(var + eps) of the folded LayerNorm, IEEE (the 2-ulp MUFU.RSQ measured no faster)
__device__ __forceinline__ float row_rstd(float v) { return 1.0f / sqrtf(v); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Timeline of CTA 0 (Params::trace != 0; test hook biom3_debug_trace(1, ...)): [who][tile < 64][event] = clock64()
//   who 0 MMA issuer:      0 before acc_empty wait, 1 stage free, 2 last k-block issued
//   who 1 epilogue warp 0: 0 before acc_full wait, 1 accumulator ready, 2 stage released (last TMEM load landed), 3 tile done
//   who 2 TMA producer:    0 first k-block of the tile requested, 1 last requested
__device__ long long g_gemm_trace[3][64][4];

template <int BN, int STAGES, int EPI, bool CG2>
__global__ void __launch_bounds__(32 * (EPI_WARP0 + epi_warps(EPI)), 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_d, const Params p) {
  static_assert(BN == 128 || BN == 256, "BN");
  using SL = SmemLayout<BN, STAGES, CG2, EPI>;
  constexpr int EPI_WARPS = epi_warps(EPI);
  constexpr int STG_BYTES = SL::STG_BYTES_TOTAL / EPI_WARPS;
  constexpr int CV_BYTES = CV_TOTAL / 2 / EPI_WARPS;   // one buffer: scale vector, then shift vector
  constexpr int CV_HALF = CV_BYTES / 2;             // bytes of one per-column vector of a warp
  constexpr uint32_t TMEM_COLS = 2 * BN;           // two accumulator stages
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(CG2 ? 2 * BM : BM, BN);
  constexpr int TILE_M = CG2 ? 2 * BM : BM;        // rows of one scheduled tile

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  constexpr bool TMA_RESID = (EPI == EPI_BIAS_RESID_SPLIT_TMA);
  __shared__ uint64_t res_full[TMA_RESID ? EPI_WARPS * RESID_SLOTS : 1];   // one barrier per (epilogue warp, residual slot)
  __shared__ uint32_t tmem_base_slot;

  // The warp index goes through a shuffle so that the compiler can prove it warp-uniform: everything derived from it (roles,
  // shared- and tensor-memory addresses, descriptors, TMA coordinates) then lives in uniform registers, and the
  // single-thread tcgen05 / TMA instructions (issued by one ELECTed lane of a converged warp) need no per-lane "waterfall"
  // loop.  With `threadIdx.x >> 5` and `if (lane == 0)` every UTCHMMA sat in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop of
  // ~17 dependent instructions (~100 cycles per MMA for the issuing thread: measured in the attention kernel's timeline,
  // profiles/r02_attn_trace_*.log).
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const uint32_t cta_rank = CG2 ? __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0) : 0u;
  const bool leader = cta_rank == 0;
  const int worker = CG2 ? int(blockIdx.x >> 1) : int(blockIdx.x);        // tile-walking unit (CTA or CTA pair)
  const int n_workers = CG2 ? int(gridDim.x >> 1) : int(gridDim.x);
  const int n_tiles = p.N / BN;
  const int num_tiles = (p.M / TILE_M) * n_tiles;
  const int nk = p.K / BK;
  const int k_blocks = p.split3 ? 3 * nk : nk;
  // tile -> (row block, column tile); rows -> (sample, position)
  auto tile_mb = [&](int tile) { int r; if (p.nt_shift >= 0) r = tile >> p.nt_shift; else r = tile / n_tiles; return r; };
  auto tile_nt = [&](int tile) { int r; if (p.nt_shift >= 0) r = tile & (n_tiles - 1); else r = tile % n_tiles; return r; };
  auto row_b = [&](int row) { int r; if (p.l_shift >= 0) r = row >> p.l_shift; else r = row / p.L; return r; };
  auto row_l = [&](int row) { int r; if (p.l_shift >= 0) r = row & (p.L - 1); else r = row % p.L; return r; };
  // Tile walk of this worker: t = t_begin, t_begin + t_step, ... < t_end (round robin over the workers)
  const int t_begin = worker, t_step = n_workers, t_end = num_tiles;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&tmap_a);
    ptx::tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], CG2 ? 2 * EPI_WARPS : EPI_WARPS);
    }
    if constexpr (TMA_RESID) {
      ptx::tma_prefetch_desc(&tmap_c);
      ptx::tma_prefetch_desc(&tmap_d);
      for (int s = 0; s < EPI_WARPS * RESID_SLOTS; ++s) ptx::mbar_init(&res_full[s], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG2) {
      ptx::tmem_alloc_cg2(&tmem_base_slot, TMEM_COLS);
      ptx::tmem_relinquish_cg2();
    } else {
      ptx::tmem_alloc(&tmem_base_slot, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (CG2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);
  ptx::pdl_sync();          // everything above touched only this CTA's smem / TMEM

  if (warp < EPI_WARP0) {
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA; converged warp, one elected lane issues)
    {
      uint32_t stage = 0, phase = 0;
      for (int t = t_begin; t < t_end; t += t_step) {
        const int tile = p.reverse ? num_tiles - 1 - t : t;
        const int m0 = tile_mb(tile) * TILE_M + int(cta_rank) * BM;
        const int n0 = tile_nt(tile) * BN + int(cta_rank) * SL::B_ROWS * (CG2 ? 1 : 0) + p.b_row_offset;
        // (Round 2: an L2 prefetch of the NEXT tile's A boxes from here, cp.async.bulk.prefetch.tensor, made the step 6 %
        // slower, FF2 +24 %: the operand feed of these GEMMs is not bound by first-touch HBM latency.)
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait_parked(&empty_bar[stage], phase ^ 1);
          if (p.trace && blockIdx.x == 0 && lane == 0 && (kb == 0 || kb == k_blocks - 1)) {
            const int ti = (t - t_begin) / t_step;
            if (ti < 64) g_gemm_trace[2][ti][kb == 0 ? 0 : 1] = clock64();
          }
          uint8_t* sa = smem + stage * SL::STAGE_BYTES;
          uint8_t* sb = sa + SL::A_BYTES;
          // split3: blocks [0, nk) A_hi.W_hi, [nk, 2nk) A_hi.W_lo, [2nk, 3nk) A_lo.W_hi
          const int ka = (kb < nk ? kb : kb - nk) * BK;
          const int kw = (kb < 2 * nk ? kb : kb - 2 * nk) * BK;
          if (ptx::elect_one()) {
            if constexpr (CG2) {
              // both CTAs' bytes are counted on the leader's barrier, which only the leader arms
              if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * SL::STAGE_BYTES);
              ptx::tma_load_2d_cg2(sa, &tmap_a, &full_bar[stage], ka, m0 + p.a_row_offset);
              ptx::tma_load_2d_cg2(sb, &tmap_b, &full_bar[stage], kw, n0);
            } else {
              ptx::mbar_arrive_expect_tx(&full_bar[stage], SL::STAGE_BYTES);
              ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], ka, m0 + p.a_row_offset);
              ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kw, n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only when paired; converged warp,
    // one elected lane issues the MMAs and commits of a k-block)
    if (leader) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int t = t_begin; t < t_end; t += t_step, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        const bool tr = p.trace && blockIdx.x == 0 && lane == 0 && it < 64;
        if (tr) g_gemm_trace[0][it][0] = clock64();
        ptx::mbar_wait_parked(&acc_empty[as], aphase ^ 1);
        if (tr) g_gemm_trace[0][it][1] = clock64();
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint64_t da = ptx::umma_desc_sw128(sa);
          const uint64_t db = ptx::umma_desc_sw128(sa + SL::A_BYTES);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
              if constexpr (CG2) ptx::umma_bf16_cg2(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), IDESC, (kb | k) != 0);
              else ptx::umma_bf16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), IDESC, (kb | k) != 0);
            }
            if constexpr (CG2) {
              ptx::umma_commit_cg2(&empty_bar[stage], 3);     // frees the slot in BOTH CTAs
              if (kb == k_blocks - 1) ptx::umma_commit_cg2(&acc_full[as], 3);
            } else {
              ptx::umma_commit(&empty_bar[stage]);            // smem slot free once these MMAs retire
              if (kb == k_blocks - 1) ptx::umma_commit(&acc_full[as]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (tr) g_gemm_trace[0][it][2] = clock64();
      }
    }
  }
  } else {
    // ------------------------------------------------------------ epilogue (8 or 16 warps, every CTA)
    const uint32_t ew = warp - EPI_WARP0;
    const uint32_t quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter ..
    const uint32_t col_half = ew >> 2;                 // 2 or 4 warps share a lane quarter and split the columns
    constexpr int COLS_PER_WARP = BN / (EPI_WARPS / 4);
    constexpr int NCH = COLS_PER_WARP / 32;
    const uint32_t stg = ptx::smem_u32(smem + SL::STG_OFFSET + ew * STG_BYTES);
    const uint32_t cvs0 = ptx::smem_u32(smem + SL::CV_OFFSET + ew * 2 * CV_BYTES);   // two buffers: scale vector, then shift vector
    const int rr = lane >> 3, ch = lane & 7;           // fp32 read-phase mapping: row 4j + rr, 16-byte group ch
    constexpr bool RESID = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_SPLIT);
    constexpr bool SPLIT = (EPI == EPI_BIAS_RESID_SPLIT);
    auto tile_goff = [&](int t) -> size_t {            // element offset of this lane's first element of a tile (row rr, group ch)
      const int tile = p.reverse ? num_tiles - 1 - t : t;
      const int m0 = tile_mb(tile) * TILE_M + int(cta_rank) * BM;
      return size_t(m0 + quarter * 32 + rr) * p.N + tile_nt(tile) * BN + col_half * COLS_PER_WARP + 4 * ch;
    };
    // 4 residual values at element offset `off`, as raw bits: fp32 x 4 or (hi bf16 x 4, lo bf16 x 4)
    auto load_res = [&](size_t off) -> uint4 {
      if constexpr (SPLIT) {
        const uint2 h = *reinterpret_cast<const uint2*>(p.out_bf16 + off);
        const uint2 l = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.out) + off);
        return make_uint4(h.x, h.y, l.x, l.y);
      } else {
        return *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.out) + off);
      }
    };
    // Hand an accumulator stage back to the MMA warp (leader CTA) as soon as this warp's last TMEM load of the tile
    // has landed in registers; the math and the stores of that last chunk then overlap the next-but-one mainloop.
    auto release_acc = [&](uint32_t as) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG2 && !leader) ptx::mbar_arrive_remote(&acc_empty[as], 0);
        else ptx::mbar_arrive(&acc_empty[as]);
      }
    };
    if constexpr (TMA_RESID) {
      // ---------------------------------------------------------- epilogue 6: TMA-fed residual ring, thread = row
      static_assert(!TMA_RESID || (BN == 256 && EPI_WARPS == 8), "epilogue 6: 128 columns per warp");
      constexpr int RS = RESID_SLOTS;
      uint8_t* const slot_p = smem + SL::STG_OFFSET + ew * (RS * RESID_SLOT_BYTES);
      const uint32_t slot0 = ptx::smem_u32(slot_p);
      const uint32_t addv0 = ptx::smem_u32(smem + SL::CV_OFFSET + ew * 1024);      // two 512-byte buffers
      uint64_t* const rbar = &res_full[ew * RS];
      const int my_tiles = t_begin < t_end ? (t_end - t_begin + t_step - 1) / t_step : 0;
      const int n_chunks = my_tiles * NCH;
      auto order4 = [](int i) { return ((i & 1) << 1) | (i >> 1); };               // 0, 2, 1, 3 (see below)
      // TMA coordinates (column, row) of the n-th chunk of this warp's walk
      auto chunk_coord = [&](int n, int& c0, int& c1) {
        const int t = t_begin + (n >> 2) * t_step;
        const int tile = p.reverse ? num_tiles - 1 - t : t;
        c1 = tile_mb(tile) * TILE_M + int(cta_rank) * BM + quarter * 32;
        c0 = tile_nt(tile) * BN + col_half * COLS_PER_WARP + order4(n & 3) * 32;
      };
      auto issue_load = [&](int n, uint32_t s) {                                   // one elected lane
        int c0, c1;
        chunk_coord(n, c0, c1);
#ifdef BIOM3_ABL_NO_RESID_LOAD
        ptx::mbar_arrive(&rbar[s]);
#else
        ptx::mbar_arrive_expect_tx(&rbar[s], RESID_SLOT_BYTES);
        ptx::tma_load_2d(slot_p + s * RESID_SLOT_BYTES, &tmap_c, &rbar[s], c0, c1);
        ptx::tma_load_2d(slot_p + s * RESID_SLOT_BYTES + 2048, &tmap_d, &rbar[s], c0, c1);
#endif
      };
      if (ptx::elect_one()) {
        for (int n = 0; n < RS && n < n_chunks; ++n) issue_load(n, uint32_t(n));
      }
      __syncwarp();
      int n = 0;                         // chunks consumed so far
      uint32_t slot = 0, sphase = 0;     // ring position of chunk n
      int pend_n = -1;                   // chunk whose slot waits for its refill (BIOM3_RESID_REFILL_LATE)
      uint32_t pend_slot = 0;
      uint32_t it = 0;
      for (int t = t_begin; t < t_end; t += t_step, ++it) {
        const int tile = p.reverse ? num_tiles - 1 - t : t;
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        const int m0 = tile_mb(tile) * TILE_M + int(cta_rank) * BM;
        const int n_tile = tile_nt(tile);
        const int rbase = m0 + quarter * 32;
        const int nbase = n_tile * BN + col_half * COLS_PER_WARP;
        const int bidx = row_b(rbase);
        const uint32_t t_row = tmem_base + ((quarter * 32u) << 16) + as * BN + col_half * COLS_PER_WARP;
        // bias (+ conditioning) of this warp's 128 columns -> warp-private smem, read back as broadcasts
        const uint32_t av = addv0 + (it & 1) * 512;
        {
          float4 a4 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase) + lane);
          if (p.cond) {
            const float4 cv = __ldg(reinterpret_cast<const float4*>(p.cond + size_t(bidx) * p.cond_stride + nbase) + lane);
            a4.x += cv.x; a4.y += cv.y; a4.z += cv.z; a4.w += cv.w;
          }
          st_shared_v4(av + lane * 16, __float_as_uint(a4.x), __float_as_uint(a4.y), __float_as_uint(a4.z), __float_as_uint(a4.w));
        }
        __syncwarp();
        float rs = 0.f, rq = 0.f;
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][0] = clock64();
        ptx::mbar_wait_parked(&acc_full[as], aphase);
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][1] = clock64();
        ptx::tc_fence_after();
        uint32_t r[2][32];                                   // TMEM loads run one chunk ahead
        ptx::tmem_ld_32x32(t_row + order4(0) * 32, r[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci, ++n) {
          // 32-column chunks in the order 0, 2, 1, 3: two chunks of a bf16 plane share a 128-byte line, and the chunk being
          // fetched must not hit the line the current chunk is storing to (round 1: walking 0, 1, 2, 3 was 2.5 x slower)
          const int c = order4(ci);
          ptx::tmem_ld_wait();
          if (ci + 1 < NCH) ptx::tmem_ld_32x32(t_row + order4(ci + 1) * 32, r[(ci + 1) & 1]);
          else {
            release_acc(as);
            if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][2] = clock64();
          }
          ptx::mbar_wait(&rbar[slot], sphase);
          const uint32_t sp = slot0 + slot * RESID_SLOT_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) {                      // 8 columns per step: one 16-byte piece of the hi and of the lo row
#if BIOM3_RESID_REFILL_LATE
            // the previous chunk's slot is refilled half a chunk later, when its TMA store has long finished reading it
            // (waiting right after the store stalled the warp for the store's issue-to-read latency every chunk)
            if (i == 2 && pend_n >= 0) {
              if (ptx::elect_one()) {
                ptx::tma_store_wait_read();
                issue_load(pend_n + RS, pend_slot);
              }
              pend_n = -1;
            }
#endif
            const uint32_t off = lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4);
            const uint4 h4 = ld_shared_v4(sp + off), l4 = ld_shared_v4(sp + 2048 + off);
            const uint4 a0 = ld_shared_v4(av + c * 128 + i * 32), a1 = ld_shared_v4(av + c * 128 + i * 32 + 16);
            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
            const float ad[8] = {__uint_as_float(a0.x), __uint_as_float(a0.y), __uint_as_float(a0.z), __uint_as_float(a0.w),
                                 __uint_as_float(a1.x), __uint_as_float(a1.y), __uint_as_float(a1.z), __uint_as_float(a1.w)};
            uint32_t nh[4], nl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {                    // same expression order as epilogue 5: acc + ((hi + lo) + add)
              const float r0 = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
              const float r1 = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
              const float o0 = __uint_as_float(r[ci & 1][8 * i + 2 * e]) + (r0 + ad[2 * e]);
              const float o1 = __uint_as_float(r[ci & 1][8 * i + 2 * e + 1]) + (r1 + ad[2 * e + 1]);
              nh[e] = ptx::pack_bf16x2(o0, o1);
              nl[e] = ptx::pack_bf16x2(o0 - __uint_as_float(nh[e] << 16), o1 - __uint_as_float(nh[e] & 0xffff0000u));
              rs += o0 + o1;
              rq = fmaf(o0, o0, fmaf(o1, o1, rq));
            }
            st_shared_v4(sp + off, nh[0], nh[1], nh[2], nh[3]);
            st_shared_v4(sp + 2048 + off, nl[0], nl[1], nl[2], nl[3]);
          }
          // the updated slot IS the TMA box image of both planes: publish it to the async proxy, one elected lane stores,
          // waits until the store has read the slot, and refills it with the chunk RS positions further on
          ptx::fence_proxy_async();
          __syncwarp();
          if (ptx::elect_one()) {
            int c0, c1;
            chunk_coord(n, c0, c1);
#ifndef BIOM3_ABL_NO_RESID_STORE
            ptx::tma_store_2d(&tmap_c, sp, c0, c1);
            ptx::tma_store_2d(&tmap_d, sp + 2048, c0, c1);
#endif
            ptx::tma_store_commit();
#if !BIOM3_RESID_REFILL_LATE
            if (n + RS < n_chunks) {
              ptx::tma_store_wait_read();
              issue_load(n + RS, slot);
            }
#endif
          }
#if BIOM3_RESID_REFILL_LATE
          if (n + RS < n_chunks) { pend_n = n; pend_slot = slot; }
#else
          __syncwarp();
#endif
          if (++slot == RS) { slot = 0; sphase ^= 1; }
        }
        if (p.stats_out) {
          const int parts = n_tiles * 2, part = n_tile * 2 + col_half;
          *reinterpret_cast<float2*>(p.stats_out + (size_t(rbase + lane) * parts + part) * 2) = make_float2(rs, rq);
        }
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][3] = clock64();
      }
    } else {
    // Residual rows of the chunk being processed; each slot is reloaded right after it is consumed with the value of the
    // next chunk of the walk (the next tile's first chunk at the end of a tile): 32 registers of prefetched data.  The walk
    // visits the 32-column chunks 0, 2, 1, 3 so that the prefetch of the next chunk never touches the 128-byte line the
    // current chunk is storing to (two chunks of a bf16 plane share a line; walking 0, 1, 2, 3 was 2.5 x slower).  A
    // second chunk in flight (setmaxnreg register reallocation), an 8-bit lo plane and a thread-per-row variant were
    // built in round 1 and measured slower (profiles/r01_ab_resid_depth.jsonl, r01_ab_lo8.jsonl, r01_ab_resid_direct.jsonl).
    auto order = [](int i) { return (SPLIT && NCH == 4) ? (((i & 1) << 1) | (i >> 1)) : i; };
    uint4 res[8];
    if constexpr (RESID) {
      if (t_begin < t_end) {
        const size_t g0 = tile_goff(t_begin);
#pragma unroll
        for (int j = 0; j < 8; ++j) res[j] = load_res(g0 + order(0) * 32 + size_t(4 * j) * p.N);
      }
    }
    // bf16 epilogues with a folded LayerNorm: the per-tile inputs (this warp's slices of ln_s / ln_t and its rows' partial
    // statistics) are fetched ONE TILE AHEAD — vectors with cp.async straight into the other smem buffer, statistics into
    // 8 registers — so their L2 latency no longer sits at the head of every tile (it was ~30 % of the epilogue warps'
    // stall samples in ncu).  Statistics with more than 4 partial slots per row fall back to loading at the point of use.
    constexpr bool kVecEpi = (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_QKV_HEADMAJOR);
    const bool pipe = kVecEpi && p.ln_stats != nullptr && (p.ln_parts == 2 || p.ln_parts == 4);
    float4 pst[2];                                        // prefetched partial statistics of this thread's row
    auto tile_rn = [&](int t, int& rbase_o, int& nbase_o) {
      const int tile = p.reverse ? num_tiles - 1 - t : t;
      rbase_o = tile_mb(tile) * TILE_M + int(cta_rank) * BM + quarter * 32;
      nbase_o = tile_nt(tile) * BN + col_half * COLS_PER_WARP;
    };
    auto prefetch_tile = [&](int t, uint32_t buf) {
      int rb, nb;
      tile_rn(t, rb, nb);
      if (lane < COLS_PER_WARP / 4) {
        ptx::cp_async_16(buf + lane * 16, p.ln_s + nb + lane * 4);
        ptx::cp_async_16(buf + CV_HALF + lane * 16, p.ln_t + nb + lane * 4);
      }
      ptx::cp_async_commit();
      const float4* sp4 = reinterpret_cast<const float4*>(p.ln_stats + size_t(rb + lane) * p.ln_parts * 2);
      pst[0] = sp4[0];
      if (p.ln_parts == 4) pst[1] = sp4[1];
    };
    if constexpr (kVecEpi) {
      if (pipe && t_begin < t_end) prefetch_tile(t_begin, cvs0);
    }
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; t += t_step, ++it) {
      const int tile = p.reverse ? num_tiles - 1 - t : t;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int m0 = tile_mb(tile) * TILE_M + int(cta_rank) * BM;
      const int n_tile = tile_nt(tile);
      const int rbase = m0 + quarter * 32;                 // first of this warp's 32 rows
      const int nbase = n_tile * BN + col_half * COLS_PER_WARP;
      const int bidx = row_b(rbase);                       // 32-row blocks never straddle samples (L % 128 == 0)
      const uint32_t t_row = tmem_base + ((quarter * 32u) << 16) + as * BN + col_half * COLS_PER_WARP;

      if constexpr (RESID || EPI == EPI_STORE_F32) {
        // fp32 path.  Read phase: iteration j covers rows 4j..4j+3, lane -> (row 4j + lane/8, 16-byte column
        // group lane%8): every global access is 4 full 128-byte lines per warp instruction (4 x 64 bytes per array
        // in the split form).  The residual of the NEXT chunk (or of the next tile's first chunk) is always in flight
        // while this one is processed.
        float rs[8], rq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rs[j] = rq[j] = 0.f;
        const size_t goff = tile_goff(t);
        const int next_t = t + t_step;
        const bool have_next = next_t < t_end;
        const size_t gnext = have_next ? tile_goff(next_t) : 0;
        float4 addv[NCH];                                  // bias (+ conditioning) of this lane's columns, per chunk
        if constexpr (RESID) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            addv[c] = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + c * 32 + 4 * ch));
            if (p.cond) {
              const float4 cv = __ldg(reinterpret_cast<const float4*>(p.cond + size_t(bidx) * p.cond_stride + nbase + c * 32 + 4 * ch));
              addv[c].x += cv.x; addv[c].y += cv.y; addv[c].z += cv.z; addv[c].w += cv.w;
            }
          }
        }
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][0] = clock64();
        ptx::mbar_wait_parked(&acc_full[as], aphase);
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][1] = clock64();
        ptx::tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = order(ci);
          // next position of the walk: the next chunk of this tile, or the next tile's first chunk
          const bool more = (ci + 1 < NCH) || have_next;
          const size_t gn = (ci + 1 < NCH) ? goff + order(ci + 1) * 32 : gnext + order(0) * 32;
          auto& resb = res;
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_row + c * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            st_shared_v4(stg + lane * 128 + ((i ^ (lane & 7)) << 4), r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          __syncwarp();
          float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (RESID) add = addv[c];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = 4 * j + rr;
            const uint4 a4 = ld_shared_v4(stg + row * 128 + ((ch ^ (row & 7)) << 4));
            float4 o = make_float4(__uint_as_float(a4.x), __uint_as_float(a4.y), __uint_as_float(a4.z), __uint_as_float(a4.w));
            if constexpr (RESID) {
              const uint4 rb = resb[j];
              if (more) resb[j] = load_res(gn + size_t(4 * j) * p.N);
              float4 rv;
              if constexpr (SPLIT) {
                rv = make_float4(__uint_as_float(rb.x << 16) + __uint_as_float(rb.z << 16),
                                 __uint_as_float(rb.x & 0xffff0000u) + __uint_as_float(rb.z & 0xffff0000u),
                                 __uint_as_float(rb.y << 16) + __uint_as_float(rb.w << 16),
                                 __uint_as_float(rb.y & 0xffff0000u) + __uint_as_float(rb.w & 0xffff0000u));
              } else {
                rv = make_float4(__uint_as_float(rb.x), __uint_as_float(rb.y), __uint_as_float(rb.z), __uint_as_float(rb.w));
              }
              o.x += rv.x + add.x; o.y += rv.y + add.y; o.z += rv.z + add.z; o.w += rv.w + add.w;
            }
            const size_t eoff = goff + size_t(4 * j) * p.N + c * 32;
            if constexpr (SPLIT) {
              const uint32_t h0 = ptx::pack_bf16x2(o.x, o.y), h1 = ptx::pack_bf16x2(o.z, o.w);
              *reinterpret_cast<uint2*>(p.out_bf16 + eoff) = make_uint2(h0, h1);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + eoff) =
                  make_uint2(ptx::pack_bf16x2(o.x - __uint_as_float(h0 << 16), o.y - __uint_as_float(h0 & 0xffff0000u)),
                             ptx::pack_bf16x2(o.z - __uint_as_float(h1 << 16), o.w - __uint_as_float(h1 & 0xffff0000u)));
            } else {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + eoff) = o;
            }
            if constexpr (RESID) {
              if constexpr (!SPLIT) {
                if (p.out_bf16)
                  *reinterpret_cast<uint2*>(p.out_bf16 + eoff) = make_uint2(ptx::pack_bf16x2(o.x, o.y), ptx::pack_bf16x2(o.z, o.w));
              }
              rs[j] += (o.x + o.y) + (o.z + o.w);
              rq[j] += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
            }
          }
          __syncwarp();
        }
        if constexpr (RESID) {
          if (p.stats_out) {
            const int parts = n_tiles * 2, part = n_tile * 2 + col_half;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float s = rs[j], q = rq[j];
#pragma unroll
              for (int o = 1; o < 8; o <<= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
              }
              if (ch == 0)
                *reinterpret_cast<float2*>(p.stats_out + (size_t(rbase + 4 * j + rr) * parts + part) * 2) = make_float2(s, q);
            }
          }
        }
        release_acc(as);           // (an earlier release costs registers this 168-register path does not have)
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][2] = clock64();
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][3] = clock64();
      } else {
        // bf16 path: thread = row while the fused math runs, then a 32 x 64-byte block is transposed
        // through smem so each warp store instruction writes 8 rows x 64 contiguous bytes.
        // Per-column epilogue vectors (x = acc * rstd + (-mean * rstd) * cs[n] + ct[n]; cs/ct = folded-LayerNorm
        // s/t, or 0/bias) are staged once per tile in warp-private smem BEFORE the accumulator wait: with ~225 KB of
        // the SM's 256 KB carved out as shared memory there is next to no L1 left for them.
        constexpr bool kHasVec = (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_QKV_HEADMAJOR || EPI == EPI_BIAS_GELU_SPLIT);
        float mean = 0.f, rstd = 1.f;
        const bool use_vec = kHasVec && (p.ln_stats != nullptr || p.bias != nullptr);
        const uint32_t cvs = cvs0 + (pipe ? (it & 1) * CV_BYTES : 0);
        if (pipe) {
          // this tile's vectors and statistics were requested one tile ago
          float sm = 0.f, q = 0.f;
          sm += pst[0].x; q += pst[0].y;
          sm += pst[0].z; q += pst[0].w;
          if (p.ln_parts == 4) {
            sm += pst[1].x; q += pst[1].y;
            sm += pst[1].z; q += pst[1].w;
          }
          mean = sm / float(p.K);
          rstd = row_rstd(fmaxf(q / float(p.K) - mean * mean, 0.f) + 1e-5f);
          ptx::cp_async_wait<0>();
          __syncwarp();
          if (t + t_step < t_end) prefetch_tile(t + t_step, cvs0 + ((it + 1) & 1) * CV_BYTES);
        } else if (use_vec) {
          const float* tsrc = p.ln_stats ? p.ln_t : p.bias;
          const int vl = lane < COLS_PER_WARP / 4 ? lane : 0;      // lanes that carry a float4 of the vectors
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(tsrc + nbase) + vl);
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.ln_stats) s4 = __ldg(reinterpret_cast<const float4*>(p.ln_s + nbase) + vl);
          __syncwarp();
          if (lane < COLS_PER_WARP / 4) {
            // the GELU epilogue works on x / 2; a folded-LayerNorm ln_t already holds t / 2 (see Params::ln_t)
            const float tsc = (EPI == EPI_BIAS_GELU_BF16 && !p.ln_stats) ? 0.5f : 1.0f;
            st_shared_v4(cvs + lane * 16, __float_as_uint(s4.x), __float_as_uint(s4.y), __float_as_uint(s4.z), __float_as_uint(s4.w));
            st_shared_v4(cvs + CV_HALF + lane * 16, __float_as_uint(t4.x * tsc), __float_as_uint(t4.y * tsc),
                         __float_as_uint(t4.z * tsc), __float_as_uint(t4.w * tsc));
          }
          if (p.ln_stats) {
            const float2* sp = reinterpret_cast<const float2*>(p.ln_stats) + size_t(rbase + lane) * p.ln_parts;
            float sm = 0.f, q = 0.f;
            for (int i = 0; i < p.ln_parts; ++i) {
              const float2 v = sp[i];
              sm += v.x;
              q += v.y;
            }
            mean = sm / float(p.K);
            rstd = row_rstd(fmaxf(q / float(p.K) - mean * mean, 0.f) + 1e-5f);
          }
          __syncwarp();
        }
        if constexpr (EPI == EPI_BIAS_GELU_BF16) rstd *= 0.5f;   // x / 2 = acc * (rstd / 2) + (-mean * rstd / 2) * s[n] + t[n] / 2
        const float nm = -mean * rstd;
        int tma_c0 = nbase, tma_c1 = rbase;                // TMA-store coordinates (column, row) of chunk 0
        if constexpr (EPI == EPI_QKV_HEADMAJOR) {
          tma_c0 = 0;
          tma_c1 = row_l(rbase);                           // row inside the (which, sample, head) plane; the plane is per chunk
        }
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][0] = clock64();
        ptx::mbar_wait_parked(&acc_full[as], aphase);
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][1] = clock64();
        ptx::tc_fence_after();
        uint32_t r[2][32];                                 // TMEM loads run one chunk ahead of the math
        ptx::tmem_ld_32x32(t_row, r[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          ptx::tmem_ld_wait();
          if (c + 1 < NCH) ptx::tmem_ld_32x32(t_row + (c + 1) * 32, r[(c + 1) & 1]);
          else { release_acc(as); if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][2] = clock64(); }
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[c & 1][i]);
          if (use_vec) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 su = ld_shared_v4(cvs + c * 128 + i * 16);
              const uint4 tu = ld_shared_v4(cvs + CV_HALF + c * 128 + i * 16);
              v[4 * i] = fmaf(v[4 * i], rstd, fmaf(nm, __uint_as_float(su.x), __uint_as_float(tu.x)));
              v[4 * i + 1] = fmaf(v[4 * i + 1], rstd, fmaf(nm, __uint_as_float(su.y), __uint_as_float(tu.y)));
              v[4 * i + 2] = fmaf(v[4 * i + 2], rstd, fmaf(nm, __uint_as_float(su.z), __uint_as_float(tu.z)));
              v[4 * i + 3] = fmaf(v[4 * i + 3], rstd, fmaf(nm, __uint_as_float(su.w), __uint_as_float(tu.w)));
            }
          }
          if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf_half(use_vec ? v[i] : 0.5f * v[i]);
          }
          [[maybe_unused]] uint32_t chunk_desc = 0;
          if constexpr (EPI == EPI_QKV_HEADMAJOR) {
            chunk_desc = p.qkv_chunk[(nbase >> 5) + c];
            if ((chunk_desc >> 14) == 1) {                 // q of a linear-attention head: softmax over its 32 features
              float m4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
              for (int i = 4; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], v[i]);
              const float ml = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * 1.4426950408889634f;
              float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[i]) : "f"(fmaf(v[i], 1.4426950408889634f, -ml)));
                s4[i & 3] += v[i];
              }
              const float inv = 1.f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= inv;
            }
          }
          if constexpr (EPI == EPI_BIAS_GELU_SPLIT) {
            // exact erf form (torch's default nn.GELU), then the second staging block takes the lo halves
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752440f));
            if (lane == 0) ptx::tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hw[e] = ptx::pack_bf16x2(v[8 * i + 2 * e], v[8 * i + 2 * e + 1]);
                lw[e] = ptx::pack_bf16x2(v[8 * i + 2 * e] - __uint_as_float(hw[e] << 16), v[8 * i + 2 * e + 1] - __uint_as_float(hw[e] & 0xffff0000u));
              }
              const uint32_t off = lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4);
              st_shared_v4(stg + off, hw[0], hw[1], hw[2], hw[3]);
              st_shared_v4(stg + 2048 + off, lw[0], lw[1], lw[2], lw[3]);
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (ptx::elect_one()) {
              ptx::tma_store_2d(&tmap_c, stg, nbase + c * 32, rbase);
              ptx::tma_store_2d(&tmap_c, stg + 2048, p.N + nbase + c * 32, rbase);
              ptx::tma_store_commit();
            }
            __syncwarp();
          } else {
          // the previous chunk's TMA store must have finished READING the staging block before it is rewritten
          if (lane == 0) ptx::tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_shared_v4(stg + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4), ptx::pack_bf16x2(v[8 * i], v[8 * i + 1]),
                         ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]), ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]),
                         ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          // the 64B-swizzled staging block IS the TMA box layout: publish it to the async proxy, one elected lane stores
          // (a coalesced st.global path and direct 256-bit stores were measured 5-7 % slower, profiles/r01_ab_store_mode.jsonl)
          ptx::fence_proxy_async();
          __syncwarp();
          if (ptx::elect_one()) {
            if constexpr (EPI == EPI_QKV_HEADMAJOR)
              ptx::tma_store_2d(&tmap_c, stg, 0, ((int((chunk_desc >> 12) & 3) * p.Bsz + bidx) * p.H + int(chunk_desc & 0xfff)) * p.L + tma_c1);
            else
              ptx::tma_store_2d(&tmap_c, stg, tma_c0 + c * 32, tma_c1);
            ptx::tma_store_commit();
          }
          __syncwarp();
          }
        }
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) g_gemm_trace[1][it][3] = clock64();
      }
    }
    }   // !TMA_RESID
  }

  if (warp >= EPI_WARP0 && lane == 0) ptx::tma_store_wait_read();   // smem must outlive the last TMA-store reads
  ptx::tc_fence_before();
  __syncwarp();
  if constexpr (CG2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if constexpr (CG2) ptx::tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace gemm
