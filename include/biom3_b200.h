/* biom3_b200 — C ABI of the B200-native ProteoScribe sampling path.
 *
 * Plain C types only: a host binds this with ctypes / cffi / cgo-style FFI.  Every pointer marked
 * "device" is a CUDA device pointer owned by the caller (e.g. torch.Tensor.data_ptr()); `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  All calls are asynchronous with
 * respect to the host unless stated; return 0 on success, a negative code on error, message through
 * biom3_last_error().  Nothing here falls back to the CPU: without a CUDA device every compute entry
 * point fails.
 *
 * Each entry point names the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef BIOM3_B200_H
#define BIOM3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct biom3_model biom3_model;

/* Subset of stage3_config.json the sampling path reads (stage3_config.json:16,21,28,37,43-45,56,58-60;
 * read at Stage3_source/cond_diff_transformer_layer.py:94,202-213). */
typedef struct biom3_config {
  int32_t seq_len;        /* diffusion_steps == sequence length L (cond_diff_transformer_layer.py:212) */
  int32_t dim;            /* transformer_dim */
  int32_t heads;          /* transformer_heads (dim / heads must be 32) */
  int32_t depth;          /* transformer_depth */
  int32_t n_blocks;       /* transformer_blocks (must be 1) */
  int32_t local_heads;    /* transformer_local_heads */
  int32_t local_window;   /* transformer_local_size (must be 128) */
  int32_t num_classes;    /* num_classes (<= 32) */
  int32_t text_emb_dim;   /* text_emb_dim */
  int32_t reversible;     /* transformer_reversible (must be 0) */
} biom3_config;

#define BIOM3_OK 0
#define BIOM3_ERR_INVALID -1
#define BIOM3_ERR_CUDA -2
#define BIOM3_ERR_STATE -3

/* Replaces get_model() (Stage3_source/cond_diff_transformer_layer.py:198-256): builds the graph for
 * `cfg` on CUDA device `device`, with workspace for up to `max_batch` sequences.  Unsupported
 * variants (reversible, n_blocks != 1, head dim != 32, window != 128, ...) are rejected here. */
int biom3_create(const biom3_config* cfg, int device, int max_batch, biom3_model** out);
void biom3_destroy(biom3_model* m);
const char* biom3_last_error(void);

/* Replaces model.load_state_dict() (run_ProteoScribe_sample.py:51).  `key` is the reference
 * state-dict key ("transformer.x_emb_NN.weight", ...); `data` points at the tensor's elements in HOST
 * or DEVICE memory (the library finds out which), `dtype` is one of BIOM3_DTYPE_*, `shape[ndim]` the
 * tensor's extents (row-major, contiguous).  biom3_finalize_weights() is strict: it fails if a key is
 * missing, unexpected or has the wrong number of elements.  It casts GEMM weights to bf16, stacks
 * to_q/to_k/to_v, folds the pre-norm LayerNorms, and precomputes the [L][depth][dim] time-conditioning
 * table (cond_diff_transformer_layer.py:152-154 depends on the step only). */
#define BIOM3_DTYPE_F32 0
#define BIOM3_DTYPE_BF16 1
#define BIOM3_DTYPE_F16 2
#define BIOM3_DTYPE_F64 3
int biom3_set_weight(biom3_model* m, const char* key, const void* data, int dtype, const int64_t* shape, int ndim);
int biom3_finalize_weights(biom3_model* m);

/* Arithmetic of the forward: 0 = bf16 operands, fp32 accumulate and fp32 residual stream (default; logits within
 * 1e-2 relative of the fp32 reference); 1 = fp32-class: fp32 activations, every contraction as a three-product
 * bf16 split (hi.hi + hi.lo + lo.hi, fp32 accumulate) on the same tcgen05 GEMM, fp32 attention / LayerNorm / erf
 * GELU (logits within 1e-4 relative).  The reference runs fp32 without autocast at inference
 * (Stage3_source/cond_diff_transformer_layer.py:149-176).  Must be called before biom3_finalize_weights(). */
int biom3_set_precision(biom3_model* m, int precision);

/* Replaces DiffTransformer.forward(x, t, y_c) (cond_diff_transformer_layer.py:149-176, 249-251).
 * x: device int64 [B][L]; t: device int64 [B]; y_c: device fp32 [B][text_emb_dim];
 * logits: device fp32 [B][num_classes][L] (the reference's permuted layout). */
int biom3_forward(biom3_model* m, const int64_t* x, const int64_t* t, const float* y_c, int B, float* logits,
                  void* stream);

/* Replaces batch_generate_denoised_sampled() (Stage3_source/sampling_analysis.py:204-265): the whole
 * step loop on the device, no host round trip per step.
 *   y_c      device fp32 [B][text_emb_dim]
 *   path     device int64 [B][L], a permutation per row (sampling_path)
 *   state0   device int64 [B][L] start state, or NULL for all-mask (zeros)
 *   start_step, num_steps: runs steps start_step .. start_step+num_steps-1 (<= L)
 *   group    samples per reference batch: the reference's unmask write
 *            (sampling_analysis.py:254-256) couples the samples of one call; B % group == 0
 *   noise    device fp32 [num_steps][B*L][num_classes] Exp(1) draws (what OneHotCategorical.sample()
 *            consumes, sampling_analysis.py:251), or NULL to draw on the device (Philox, `seed`)
 *   group_seeds  device uint64 [B / group] or NULL: one Philox seed per reference batch of a fused launch; batch g
 *            then draws exactly what a separate launch of its `group` samples with seed group_seeds[g] would draw
 *            (so fusing independent (prompt, replica-batch) units into one launch does not change their tokens)
 *   tokens   device int64 [B][L] final state (out)
 *   traj     device uint8 [num_steps][B][L] state after every step (out), or NULL */
int biom3_decode(biom3_model* m, const float* y_c, const int64_t* path, const int64_t* state0, int start_step,
                 int num_steps, int group, const float* noise, uint64_t seed, const uint64_t* group_seeds,
                 int64_t* tokens, uint8_t* traj, int B, void* stream);

/* Replaces `argmax(OneHotCategorical(probs=softmax(logits,1).permute(0,2,1)).sample(), -1)`
 * (transformer_training_helper.py:444-449 + sampling_analysis.py:251) at every position.
 * logits device fp32 [B][C][L]; noise device fp32 [B*L][C]; tok device int64 [B][L] (out). */
int biom3_sample_all(const float* logits, const float* noise, int64_t* tok, int B, int L, int C, void* stream);

/* Device-side replacement for the per-sample `torch.randperm(args.diffusion_steps)` sampling paths
 * (run_ProteoScribe_sample.py:103-105): one uniformly random permutation of 0..L-1 per row (sorted Philox keys).
 * path device int64 [B][L] (out), L <= 8192.  Deterministic in (seed, row); not torch's generator stream. */
int biom3_random_paths(uint64_t seed, int B, int L, int64_t* path, void* stream);

/* Replaces the unmask write `state[:, 0, loc] = tok[:, loc]` (sampling_analysis.py:254-256).
 * tok, state device int64 [B][L]; path device int64 [B][L]; step = current time index. */
int biom3_unmask(const int64_t* tok, const int64_t* path, int64_t* state, int B, int L, int group, int step,
                 void* stream);

/* Unit-test hook for the tcgen05 GEMM: C = A . W^T with one of the fused epilogues.
 * A device bf16 [M][K]; W device bf16 [N][K]; epi: 0 bf16 out, 2 bias+gelu bf16 out,
 * 3 fp32 in-place residual (+bias), 4 fp32 out, 5 in-place residual stored split (out = bf16 [2][M][N]: hi plane
 * bf16(R), lo plane bf16(R - hi)), 6 the same update through TMA-fed residual slots (pair tiling only; bit-identical
 * to 5), 7 gelu_erf(A W^T + bias) (exact erf) stored as the [hi | lo] bf16 split of the fp32-class mode (out = bf16
 * [M][2N]).  block_n: 256.  pair bit 0: CTA-pair (cta_group::2) tiling, 256 x 256 tiles
 * (needs M % 256 == 0; otherwise 128 x 256 tiles, one CTA each); pair bit 1: fp32-class K schedule, A and W are
 * [hi | lo] bf16 halves of width 2K and the result is hi.hi + hi.lo + lo.hi.  With BIOM3_GEMM_TRACE=1 in the
 * environment CTA 0 records its clock64 timeline (biom3_debug_trace(1, ...)). */
int biom3_gemm_test(const void* A, const void* W, const float* bias, void* out, int M, int N, int K, int epi,
                    int block_n, int pair, void* stream);

/* Unit-test hook for the attention kernels: qkv device bf16 [3][B][H][L][32] -> out device bf16 [B*L][H*32].
 * Heads [0, NL) windowed softmax (tcgen05 kernel; variant bit 0 = the same kernel recording CTA 0's clock64 timeline,
 * biom3_debug_trace(0, ...)), heads [NL, H) linear attention (variant bit 1: q of those heads is already softmax(q) over
 * the features -- the form the QKV GEMM epilogue writes in the decode). */
int biom3_attention_test(const void* qkv, void* out, int B, int H, int L, int NL, int variant, void* stream);

/* Per-kernel device timings (ms) of the last biom3_profile_step() call; for bench.py's roofline. */
typedef struct biom3_step_profile {
  float total_ms;
  float gemm_qkv_ms, gemm_out_ms, gemm_ff1_ms, gemm_ff2_ms;
  float local_attn_ms, linear_attn_ms, layernorm_ms, embed_ms, head_ms, other_ms;
  int32_t launches;
  int32_t compact_rows;   /* rows carried through the LAST layer's out-proj / MLP (selected tokens, padded to 256); 0 = all B*L */
} biom3_step_profile;
/* Runs one decode step un-graphed with CUDA events around every launch (synchronous). */
int biom3_profile_step(biom3_model* m, int B, int group, biom3_step_profile* out);

/* Test hook: synchronous device->host copy of an internal buffer after a forward/decode call.
 * name: "u" fp32 [B*L][D] | "a" bf16 [B*L][D] | "qkv" bf16 [3][B][H][L][32] | "att" bf16 [B*L][D] |
 * "hid" bf16 [B*L][4D] | "cvec" fp32 [B][depth][D] | "Y" fp32 [B][depth][D] | "Ttab" fp32 [L][depth][D] |
 * "state" u8 [B*L] | "stamps" u64 [L][2]: %globaltimer (ns) at the start of each step's first kernel and at the end of
 * its last one, of the last biom3_decode (consecutive CUDA-graph replays: no host-side gap between steps).
 * Copies min(nbytes, buffer size). */
int biom3_debug_copy(biom3_model* m, const char* name, void* host_dst, int64_t nbytes);

/* Replaces Facilitator.forward (Stage1_source/model.py:473-493; called at run_Facilitator_sample.py:79-83):
 * z_c = W1 . gelu_erf(W0 . z_t + b0) + b1 with weight_norm(dim=None) folded, W = g * V / ||V||_F.
 * create: weight_v / bias pointers are HOST fp32 (state-dict tensors main.0.weight_v [hid][in], main.0.weight_g
 * (scalar), main.0.bias, main.3.*); the folded weights are uploaded once and live on the handle.
 * forward: z_t device fp32 [P][in_dim] -> z_c device fp32 [P][out_dim], asynchronous on `stream`. */
typedef struct biom3_facilitator_t biom3_facilitator_t;
int biom3_facilitator_create(int in_dim, int hid_dim, int out_dim, const float* w0_v, float w0_g, const float* b0,
                             const float* w1_v, float w1_g, const float* b1, int device, biom3_facilitator_t** out);
int biom3_facilitator_forward(biom3_facilitator_t* f, const float* z_t, int P, float* z_c, void* stream);
void biom3_facilitator_destroy(biom3_facilitator_t* f);
/* One-shot form (create + forward + synchronize + destroy) on the current device. */
int biom3_facilitator(const float* z_t, int P, int in_dim, int hid_dim, int out_dim, const float* w0_v, float w0_g,
                      const float* b0, const float* w1_v, float w1_g, const float* b1, float* z_c, void* stream);

/* Test hook: the Exp(1) race noise biom3_decode() draws on the device when `noise` is NULL, for one step:
 * out device fp32 [B*L][C], out[pos][c] = the draw consumed for class c at flat position pos = b * L + l at time
 * index `step` under `seed` (Philox4x32-10 counter (pos, step, c / 4), inversion q = -log1p(-v); csrc/kernels.cuh).
 * Feeding these values as explicit `noise` to the CPU oracle reproduces the device decode bit for bit
 * (tests/test_gpu_parity.py); oracle/philox.py restates the stream independently. */
int biom3_debug_noise(uint64_t seed, int step, int B, int L, int C, float* out, void* stream);

/* Test hook: clock64() timelines of CTA 0.  which = 0: the last traced local-attention launch (biom3_attention_test with a
 * tracing variant): int64 [2 streams][2: issuers, softmax warp 0][128 blocks][6 events] (csrc/attention.cuh, g_ms_trace);
 * which = 1: the last biom3_gemm_test launch with BIOM3_GEMM_TRACE=1: int64 [3: MMA issuer, epilogue warp 0, TMA
 * producer][64 tiles][4 events] (csrc/gemm_tcgen05.cuh, g_gemm_trace).  Synchronous. */
int biom3_debug_trace(int which, void* host_dst, int64_t nbytes);

/* Number of kernel launches one decode step issues (for bench.py's gpu_launches): the count of the most recently
 * captured step graph, or the full-row estimate before the first decode. */
int biom3_launches_per_step(const biom3_model* m);

/* Range check of the device inputs of biom3_forward / biom3_decode, done where they enter the resident state.  The
 * reference raises for a token id >= num_classes (nn.Embedding, Stage3_source/cond_diff_transformer_layer.py:213) and
 * indexes with whatever time / path values it is given; here an out-of-range value is clamped (no out-of-bounds access)
 * and recorded.  Synchronises the device, writes the sticky bits to *flags and clears them:
 * 1 = token id outside [0, num_classes), 2 = time index outside [0, diffusion_steps), 4 = path entry outside
 * [0, diffusion_steps) (such an entry is skipped: a row that is not a permutation unmasks location 0 at the missing steps). */
int biom3_input_errors(biom3_model* m, int* flags);

#ifdef __cplusplus
}
#endif
#endif /* BIOM3_B200_H */
