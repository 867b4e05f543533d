/* protnote_b200 - C ABI of the B200-native ProtNote scoring path (sm_100a).
 *
 * The reference (microsoft/protnote) has no FFI layer: its boundary for this path is the Python nn.Module
 * contract (protnote/models/ProtNote.py:168-177, protnote/models/protein_encoders.py:109-123).  This header is
 * what a binding for that contract calls; every entry point cites the reference code whose arithmetic it replaces.
 * INTEGRATION.md shows the ctypes stub and the reference-side module that uses it.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host; the caller owns all memory,
 *     the library never allocates device memory and never synchronises the device;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - every function returns 0 on success, non-zero on error (pn_last_error() describes it); nothing throws;
 *   - fp32 in, fp32 out.  Internally a value travels as two fp16 planes (hi, lo) and every contraction runs on the
 *     tcgen05 tensor cores; mode PN_STRICT issues hi*hi + hi*lo + lo*hi (fp32-grade, |logit error| << 1e-4),
 *     PN_FAST issues hi*hi only (fp16 operands, comparable to the reference under torch.autocast);
 *   - a handle is not thread-safe: one per process / GPU.
 */
#ifndef PROTNOTE_B200_H
#define PROTNOTE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_STRICT 3 /* three tensor-core passes per k-step: fp32-grade result   */
#define PN_FAST 1   /* one pass: fp16 operands, fp32 accumulate                 */

#define PN_FUSION_CONCAT 0      /* [p; t]        configs/base_config.yaml:44 (default)      */
#define PN_FUSION_CONCAT_DIFF 1 /* [p; t; p - t] protnote/models/ProtNote.py:128-138        */
#define PN_FUSION_CONCAT_PROD 2 /* [p; t; p * t] protnote/models/ProtNote.py:139-150        */
#define PN_FUSION_SIMILARITY 3  /* cosine(p, t) / temperature, no output MLP (ProtNote.py:281-284) */

int pn_version(void);
/* Thread-local description of the last error on this thread ("" if none). */
const char* pn_last_error(void);
/* 0 if `device` is a compute-capability 10.x GPU this library can run on. */
int pn_device_check(int device);
/* Engine knobs.  STATE IS PROCESS-GLOBAL (one process drives one GPU in the intended deployment, bin/main.py: mp.spawn, one
 * rank per device): two models in one process cannot hold different options, and packed weights depend on some of them
 * (the truncation compensation follows the promotion periods and "bk") - re-pack after changing those.  pn_last_error is
 * thread-local; the optional timing list is mutex-guarded; everything else is plain state set before use.
 *   "bk" = 32 | 64             k-block / swizzle width (0 = auto: 32 strict, 64 fast)
 *   "promote_k_encoder" | "promote_k_pointwise" | "promote_k_heads" | "promote_k_scorer" | "promote_k_other" | "promote_k"
 *                              strict mode: K elements summed in TMEM between fp32 promotions, per stage of the path
 *                              (dilated convs + conv1 | 1x1 convs | W_p, W_l, layer-1 halves | output MLP | pn_linear,
 *                              pn_conv1d | all); 0 = never promote.  Defaults 64 | 64 | 32 | 256 | 64.
 *   "trunc_beta_ppt"           strict mode: expected shrink of the running sum per tensor-core accumulator add (the add
 *                              rounds toward zero), in 1e-12 units; folded into the packed weights per K position so that
 *                              the bias of every chunk cancels (csrc/pn_kernels.cuh, pack_weight_kernel).  Default 33000
 *                              (measured on B200, tools/trunc_comp_probe.py); 0 = off (round-1 arithmetic).
 *   "trunc_comp_c1" / "_c0"    uniform variant applied in the epilogue instead, (c1 * K_chunk + c0) * 1e-12; default off
 *   "cta2"                     CTA-pair kernels (tcgen05 cta_group::2, 256-row tiles, each CTA loads half the weight tile):
 *                              1 (default) wherever a problem has more than one 128-row tile, 0 never, -1 in PN_FAST mode
 *                              only, 2 always
 *   "f64_protein_head"         1 (default): PN_STRICT computes W_p and the protein half of output layer 1 in fp64
 *   "l2_hints"                 pair kernel: L2 eviction hints on the TMA loads (0 off - default, measured slower; 1; 2)
 *   "chunk_rows"               pairs per scorer chunk (0 = auto)
 *   "split_corr"               1: strict-mode encoder convolutions accumulate the hi*hi products and the lo corrections in
 *                              separate TMEM buffers (round-1 experiment; default 0)
 *   "fuse_features"            1: layer 1 of the pair scorer is built inside the first GEMM's operand producer (correct but
 *                              slower than the separate kernel, profiles/r01_fused_generator_probe.txt; default 0)
 *   "stats_tx"                 threads along the columns of a training-step reduction block */
int pn_set_option(const char* name, long long value);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long pn_launch_count(void);
/* Measurement hook: while enabled, every GEMM launch of the pair scorer is bracketed by CUDA events on its stream.
 * pn_gemm_timing(enable) clears earlier records; pn_gemm_timing_read waits for the recorded events and returns
 * the summed launch durations, the number of launches and their algorithmic FLOPs (2*M*N*K each). */
int pn_gemm_timing(int enable);
int pn_gemm_timing_read(double* total_ms, long long* launches, double* algorithmic_flops);

/* ------------------------------------------------------------------------------------------------------------
 * Sequence encoder: ProteInfer dilated ResNet, eval mode.
 * Replaces ProteInfer.get_embeddings (protnote/models/protein_encoders.py:109-118), i.e. MaskedConv1D (:8-17),
 * Residual (:23-67) and set_padding_to_sentinel (protnote/data/datasets.py:535-569).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pn_encoder_cfg {
  int input_channels;  /* 20   embed_sequences_params.INPUT_CHANNELS   */
  int channels;        /* 1100 OUTPUT_CHANNELS                         */
  int bottleneck;      /* 550  floor(channels * BOTTLENECK_FACTOR)     */
  int kernel_size;     /* 9                                            */
  int dilation_base;   /* 3                                            */
  int num_blocks;      /* 5                                            */
  float bn_eps;        /* 1e-3 (protein_encoders.py:36,48)             */
} pn_encoder_cfg;

/* Parameter pointers in reference state_dict order, num_batches_tracked omitted:
 *   conv1.weight (C,Cin,k), conv1.bias,
 *   then per block i: bn_activation_1.0.{weight,bias,running_mean,running_var},
 *                     masked_conv1.{weight (Cb,C,k), bias},
 *                     bn_activation_2.0.{weight,bias,running_mean,running_var},
 *                     masked_conv2.{weight (C,Cb,1), bias}
 * -> 2 + 12 * num_blocks fp32 device pointers (`params` itself is a HOST array). */
size_t pn_encoder_packed_bytes(const pn_encoder_cfg* cfg);
int pn_encoder_pack(const pn_encoder_cfg* cfg, const float* const* params, int num_params, void* packed,
                    size_t packed_bytes, void* stream);
size_t pn_encoder_workspace_bytes(const pn_encoder_cfg* cfg, int batch, int T);
/* x [batch][input_channels][T] fp32 (any float values, not only one-hot), lengths [batch] int64,
 * out [batch][channels] fp32 = masked mean over positions < length. */
int pn_encoder_forward(const pn_encoder_cfg* cfg, const void* packed, const float* x, const int64_t* lengths,
                       int batch, int T, float* out, void* workspace, size_t workspace_bytes, int mode,
                       void* stream);

/* Training-mode forward of the (frozen) encoder.  ProtNoteTrainer.train calls model.train(), which also flips the frozen
 * encoder's BatchNorm1d layers to BATCH statistics (protein_encoders.py:35-37,47-50; ProtNoteTrainer.py:844): every
 * BatchNorm normalises with the mean / biased variance over all batch x T positions of its input (padding positions are
 * zeros and count) and updates running_mean / running_var with `momentum` (0.01) and the unbiased variance.
 * packed_raw: pn_encoder_pack_raw (same parameter list as pn_encoder_pack; no BatchNorm is folded).
 * bn_params: HOST array of 8 * num_blocks fp32 device pointers, per block bn_activation_1.0.{weight, bias, running_mean,
 * running_var} then bn_activation_2.0.{...}; running statistics are updated in place when update_running != 0.
 * The whole batch is one unit (the statistics couple its sequences): workspace >= pn_encoder_train_workspace_bytes. */
int pn_encoder_pack_raw(const pn_encoder_cfg* cfg, const float* const* params, int num_params, void* packed,
                        size_t packed_bytes, void* stream);
size_t pn_encoder_train_workspace_bytes(const pn_encoder_cfg* cfg, int batch, int T);
int pn_encoder_forward_train(const pn_encoder_cfg* cfg, const void* packed_raw, const float* x, const int64_t* lengths,
                             int batch, int T, const float* const* bn_params, int num_bn_params, float momentum,
                             int update_running, float* out, void* workspace, size_t workspace_bytes, int mode,
                             void* stream);

/* The same forward with the SEQUENCES sharded over ranks (one process per GPU): this rank passes its `batch` sequences;
 * after the per-channel sums of every BatchNorm have been produced on `stream` the library calls
 *     reduce(stats, count, user, stream)          (host callback; must ENQUEUE a sum over ranks of the `count` doubles at
 *                                                   device pointer `stats` in stream order and return 0)
 * and normalises with the all-rank statistics over `total_positions` = T * (sequences of all ranks) positions.  The result
 * equals the unsharded pn_encoder_forward_train on the concatenated batch (running statistics included, on every rank).
 * `stats_buffer`: device buffer of >= 2 * round_up(channels, 64) doubles owned by the caller (so that the callback can
 * hand a tensor it already owns to its collective library); reduce == NULL -> plain single-rank behaviour. */
typedef int (*pn_reduce_fn)(double* stats, int count, void* user, void* stream);
int pn_encoder_forward_train_sharded(const pn_encoder_cfg* cfg, const void* packed_raw, const float* x,
                                     const int64_t* lengths, int batch, int T, const float* const* bn_params,
                                     int num_bn_params, float momentum, int update_running, float* out, void* workspace,
                                     size_t workspace_bytes, int mode, double total_positions, double* stats_buffer,
                                     pn_reduce_fn reduce, void* user, void* stream);

/* Same from token ids: tokens [batch][T] uint8 (index into the sorted amino-acid vocabulary, i.e. the argmax of the one-hot
 * the reference's collator builds, protnote/data/collators.py:123-133; ids >= input_channels give an all-zero column).
 * 1 byte per residue crosses PCIe instead of 80; the result is bit-identical to pn_encoder_forward on the one-hot. */
int pn_encoder_forward_tokens(const pn_encoder_cfg* cfg, const void* packed, const uint8_t* tokens,
                              const int64_t* lengths, int batch, int T, float* out, void* workspace,
                              size_t workspace_bytes, int mode, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Projection heads + pair scorer: ProtNote.forward from the projections on, eval mode
 * (protnote/models/ProtNote.py:270-322): W_p / W_l (torchvision MLP, :63-81), _get_joint_embeddings (:112-152),
 * output MLP get_mlp (:337-378), probability-space ensembling of k description rows (:308-322).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pn_scorer_cfg {
  int protein_dim;       /* 1100 PROTEIN_EMBEDDING_DIM                                   */
  int label_dim;         /* 1024 LABEL_EMBEDDING_DIM                                     */
  int latent_dim;        /* 1024 LATENT_EMBEDDING_DIM                                    */
  int proj_hidden;       /* 3072 latent_dim * PROJECTION_HEAD_HIDDEN_DIM_SCALE_FACTOR    */
  int proj_layers;       /* 4    PROJECTION_HEAD_NUM_LAYERS (>= 1)                       */
  int out_hidden;        /* 3072 int(round(OUTPUT_MLP_HIDDEN_DIM_SCALE_FACTOR * latent)) */
  int out_layers;        /* 3    OUTPUT_MLP_NUM_LAYERS hidden layers (>= 1)              */
  int out_batchnorm;     /* 1    OUTPUT_MLP_BATCHNORM (0: hidden Linear layers have a bias) */
  int fusion;            /* PN_FUSION_*                                                  */
  int descriptions_per_label; /* k: consecutive label rows ensembled per label (>= 1)   */
  float bn_eps;          /* 1e-5 torch.nn.BatchNorm1d default                            */
} pn_scorer_cfg;

/* Parameter pointers (HOST array of fp32 device pointers), reference module order:
 *   W_p: for i < proj_layers-1: {Linear.weight, BN.weight, BN.bias, BN.running_mean, BN.running_var}; last: Linear.weight
 *   W_l: same
 *   output_layer: for each hidden layer: Linear.weight, then (out_batchnorm ? BN x4 : Linear.bias);
 *                 final Linear.weight (1,H), final Linear.bias (1)
 *   (PN_FUSION_SIMILARITY: the module has no output_layer; only the W_p and W_l parameters are passed)
 */
int pn_scorer_num_params(const pn_scorer_cfg* cfg);
size_t pn_scorer_packed_bytes(const pn_scorer_cfg* cfg);
int pn_scorer_pack(const pn_scorer_cfg* cfg, const float* const* params, int num_params, void* packed,
                   size_t packed_bytes, void* stream);

/* W_p then the protein half of output layer 1:  P_f [n][protein_dim] -> P_e [n][latent_dim] (nullable) and
 * a [n][out_hidden] (BatchNorm-1 scale AND shift folded in).  workspace: pn_project_workspace_bytes(cfg, n).
 * PN_STRICT evaluates this head in fp64 on the CUDA cores (option "f64_protein_head", default 1): a[b] carries everything
 * protein b contributes to ALL of its logits, so its fp32-grade rounding error was the largest term of the logit error
 * (profiles/r02_scorer_error_by_stage.txt), and it costs 57 MFLOP per protein against 38 MFLOP per pair. */
size_t pn_project_workspace_bytes(const pn_scorer_cfg* cfg, long long rows);
int pn_project_sequences(const pn_scorer_cfg* cfg, const void* packed, const float* P_f, long long n, float* P_e,
                         float* a, void* workspace, size_t workspace_bytes, int mode, void* stream);
/* W_l then the label half of output layer 1: L_f [n][label_dim] -> L_e (nullable), c [n][out_hidden]
 * (BatchNorm-1 scale folded in).  Constant across batches in eval mode: compute once, cache (the reference
 * recomputes W_l(label_embeddings) every batch, ProtNote.py:271). */
int pn_project_labels(const pn_scorer_cfg* cfg, const void* packed, const float* L_f, long long n, float* L_e,
                      float* c, void* workspace, size_t workspace_bytes, int mode, void* stream);

/* logits[b][l / k] for b < B, label rows l < L (L multiple of k); logits row stride ld_logits floats.
 * a [B][out_hidden], c [L][out_hidden] from the two calls above; P_e / L_e only for PN_FUSION_CONCAT_PROD.
 * workspace: any size >= pn_scorer_min_workspace_bytes(cfg); more lets it run bigger chunks
 * (pn_scorer_workspace_bytes(cfg, B, L) = the size that scores everything in one chunk, capped). */
size_t pn_scorer_min_workspace_bytes(const pn_scorer_cfg* cfg);
size_t pn_scorer_workspace_bytes(const pn_scorer_cfg* cfg, long long B, long long L);
int pn_score_pairs(const pn_scorer_cfg* cfg, const void* packed, const float* a, const float* c, const float* P_e,
                   const float* L_e, long long B, long long L, float* logits, long long ld_logits, void* workspace,
                   size_t workspace_bytes, int mode, void* stream);

/* Same, and additionally stores the last hidden layer's activations relu(BN(z_n)) for every pair, fp32
 * [B*L][out_hidden] in pair order b*L + l (the "output_layer_embeddings" of ProtNote.forward(save_embeddings=True),
 * ProtNote.py:294-303).  hidden_out may be NULL. */
int pn_score_pairs_ex(const pn_scorer_cfg* cfg, const void* packed, const float* a, const float* c, const float* P_e,
                      const float* L_e, long long B, long long L, float* logits, long long ld_logits, float* hidden_out,
                      void* workspace, size_t workspace_bytes, int mode, void* stream);

/* PN_FUSION_SIMILARITY: logits[b][l/k] = <P_e[b]/|P_e[b]|, L_e[l]/|L_e[l]|> / temperature (F.normalize eps 1e-12,
 * ProtNote.py:281-284), then the same k-row ensembling.  P_e / L_e are the fp32 embeddings returned by the two
 * projection calls (their `a` / `c` outputs may be NULL for this fusion). */
size_t pn_similarity_workspace_bytes(const pn_scorer_cfg* cfg, long long B, long long L);
int pn_score_similarity(const pn_scorer_cfg* cfg, const float* P_e, const float* L_e, long long B, long long L,
                        float temperature, float* logits, long long ld_logits, void* workspace, size_t workspace_bytes,
                        int mode, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Evaluation post-processing on the device - what ProtNoteTrainer.evaluate does with every batch of logits before its
 * metrics (protnote/models/ProtNoteTrainer.py:522-537, calculate_tp_fn_fp :61-83): probs = sigmoid(logits) (nullable
 * output), predictions = probs >= threshold, per-label true positives / false negatives / false positives ADDED to
 * tp / fn / fp [L] (fp32, integer-valued).  labels: label_kind 0 none, 1 int64 multihots (collators.py), 2 float32;
 * row strides in elements.  topk > 0 (<= 64) also writes the k largest logits of every row (descending, ties by lower
 * index) and their label indices: topk_values [B][topk], topk_indices [B][topk].
 * ---------------------------------------------------------------------------------------------------------- */
int pn_postprocess(const float* logits, long long B, long long L, long long ld_logits, const void* labels, int label_kind,
                   long long ld_labels, float threshold, float* probs, long long ld_probs, float* tp, float* fn, float* fp,
                   int topk, float* topk_values, int* topk_indices, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Plain dense layer on the same engine: y[M][N] = x[M][K] * w[N][K]^T + bias   (ProteInfer.output_layer,
 * protnote/models/protein_encoders.py:120-123, and the unit tests of the engine).
 * workspace >= pn_linear_workspace_bytes(M, N, K).
 * ---------------------------------------------------------------------------------------------------------- */
size_t pn_linear_workspace_bytes(long long M, long long N, long long K);
int pn_linear(const float* x, long long M, long long K, long long ldx, const float* w, long long N,
              const float* bias, float* y, long long ldy, void* workspace, size_t workspace_bytes, int mode,
              void* stream);
/* 'same'-padded dilated Conv1d over channels-last activations on the same engine (unit test of the tap path):
 * x [batch][channels_in][T] fp32 (reference layout), w (channels_out, channels_in, taps), y [batch][T][channels_out]. */
size_t pn_conv1d_workspace_bytes(int batch, int T, int cin, int cout, int taps);
int pn_conv1d(const float* x, const int64_t* lengths, int batch, int cin, int T, const float* w, const float* bias,
              int cout, int taps, int dilation, float* y, void* workspace, size_t workspace_bytes, int mode,
              void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training step primitives (`pn_t_*`).  Replace, for the trainable part of ProtNote.forward in TRAINING mode
 * (protnote/models/ProtNote.py:168-334 with self.training; called from ProtNoteTrainer.train_one_epoch,
 * protnote/models/ProtNoteTrainer.py:721-738), the autograd graph of
 *   W_p / W_l  = torchvision MLP: [Linear(no bias), BatchNorm1d(batch statistics), ReLU] x (n-1), Linear   (ProtNote.py:63-81)
 *   output MLP = get_mlp: [Linear, BatchNorm1d(batch statistics over all B*L pairs), ReLU] x n, Linear(H -> 1) (:337-378)
 * protnote_b200/train.py sequences them (and all-reduces the per-column sums when the label axis is sharded).
 *
 * Activations travel as fp16 planes `hi`, `lo` ([rows][ld], ld a multiple of 64; lo NULL in PN_FAST) and, where a wgrad
 * will contract over rows, also as transposed planes `hiT`, `loT` in the K-blocked layout [blocksT][cols][64],
 * blocksT >= ceil(rows / 64): block i holds rows 64 i .. 64 i + 63 of every column (zero beyond `rows`).
 * A tensor may carry a power-of-two scale `sc` (device scalar; stored = true * sc; NULL = 1): gradients are tiny and
 * fp16 has 5 exponent bits.  All per-column sums are fp64.
 * ---------------------------------------------------------------------------------------------------------- */
/* sc[0] = power of two that puts max|x| into [2^5, 2^6); sc[1] is scratch (sc points to 2 floats). */
int pn_t_autoscale(const float* x, long long rows, long long cols, long long ldx, float* sc, void* stream);
/* planes (and transposed planes, nullable) of x * sc */
int pn_t_split(const float* x, long long rows, long long cols, long long ldx, const float* sc, void* hi, void* lo,
               long long ld, void* hiT, void* loT, long long blocksT, void* stream);
/* planes [N][ld] of w * ws, element (n, k) = w[n * stride_n + k * stride_k] (so a column block of a Linear weight, or
 * its transpose for dgrad, needs no copy); ws[0] = the power-of-two scale chosen, ws[1] scratch. */
int pn_t_pack_weight(const float* w, long long N, long long K, long long stride_n, long long stride_k, void* hi, void* lo,
                     long long ld, float* ws, void* stream);
/* D[M][N] = (A[M][K] * B[N][K]^T) / (s0 * s1 * s2) on the tensor-core engine (s* device scalars, nullable).
 *   forward Linear: A = activations, B = packed weight, s0 = ws
 *   dgrad:          A = g_z planes,  B = packed weight^T (output keeps g_z's scale: pass s0 = ws only)
 *   wgrad:          A = g_z^T, B = x^T (transposed planes, K-blocked: k_blocked = 1; lda / ldb unused), K = rows,
 *                   s0 = g scale, s1 = x scale
 * Output: fp32 `out_f32` (ld_out; accumulate != 0 adds to what is there) and/or planes out_hi/out_lo (ld_split).
 * scale_scratch: N floats.  promote_k: K elements summed in TMEM between fp32 promotions (0 = engine default).
 * split_k > 0 (multiple of 64, fp32 output only): K is cut into slices of split_k, one launch each, accumulated in
 * out_f32 - for wgrad, whose K (the rows of the batch) is so long that the CTAs of one launch drift out of L2 reach. */
int pn_t_gemm(const void* a_hi, const void* a_lo, long long M, long long K, long long lda, const void* b_hi,
              const void* b_lo, long long N, long long ldb, const float* s0, const float* s1, const float* s2,
              float* scale_scratch, float* out_f32, long long ld_out, int accumulate, void* out_hi, void* out_lo,
              long long ld_split, int mode, int promote_k, long long split_k, int k_blocked, void* stream);
/* out[0][c] = sum_r v[r][c], out[1][c] = sum_r v[r][c]^2 (fp64 [2][cols], overwritten) of planes (hi, lo) or, when x is
 * not NULL, of the fp32 matrix x; ld = row pitch of whichever is given. */
int pn_t_col_stats(const void* hi, const void* lo, const float* x, long long rows, int cols, long long ld, double* out,
                   void* stream);
/* BatchNorm1d training statistics -> state[4][cols] = scale (gamma*invstd), shift (beta - mean*scale), mean, invstd;
 * running_mean / running_var (nullable) are updated with `momentum` (unbiased variance), as torch.nn.BatchNorm1d does.
 * stats2 != NULL: the batch is the grid z[b][l] = a[b] + c[l] (layer 1 of the pair scorer): stats = sums of a over
 * `count` proteins, stats2 = sums of c over `count2` label rows; mean = mean_a + mean_c, var = var_a + var_c. */
int pn_t_bn_finalize(const double* stats, double count, const double* stats2, double count2, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var, int cols,
                     float* state, void* stream);
/* h = relu(z * scale + shift) as planes (+ transposed planes, nullable) */
int pn_t_bn_relu(const void* z_hi, const void* z_lo, long long rows, int cols, long long ld_z, const float* state,
                 void* h_hi, void* h_lo, long long ld_h, void* hT_hi, void* hT_lo, long long blocksT, void* stream);
/* Dropout INSIDE W_p / W_l / output_layer (OUTPUT_MLP_DROPOUT, base_config.yaml:39; ProtNote.py:70,80,101 -> torchvision MLP
 * and get_mlp :369-371): out = x * keep / (1 - p).  keep(row, col) is a pure function of (seed, row, col, cols) - a
 * counter-based generator, 16 random bits per element, keep <=> bits >= round(p * 65536) - so the backward applies the SAME
 * mask to the incoming gradient by calling the same function with the same seed; nothing is stored.  The planes variant
 * also writes the K-blocked transposed planes (nullable) a wgrad needs; it is not in-place.  p = 0 copies. */
int pn_t_dropout_planes(const void* x_hi, const void* x_lo, long long rows, int cols, long long ld_x,
                        unsigned long long seed, float p, void* hi, void* lo, long long ld, void* hiT, void* loT,
                        long long blocksT, void* stream);
/* the same mask on an fp32 matrix (the output of a projection head); out may alias x */
int pn_t_dropout_f32(const float* x, long long rows, int cols, long long ldx, unsigned long long seed, float p, float* out,
                     long long ldo, void* stream);
/* out[r] = relu(z[r] * scale + shift) . w + b  - the last hidden layer and Linear(H -> 1) (ProtNote.py:373-377) */
int pn_t_bn_relu_dot(const void* z_hi, const void* z_lo, long long rows, int cols, long long ld_z, const float* state,
                     const float* w, const float* b, float* out, void* stream);
/* The same pass with the loss fused in: additionally g_out[r] = grad_scale * d loss(x_r, targets[r]) / d x_r
 * and *loss_sum += sum_r loss(x_r, targets[r]) (fp64, caller zeroes it), so the [B, L] logits need not round-trip through
 * autograd and the loss module's elementwise passes disappear.  `out` (the logits) is optional.  Rows are pairs in
 * (protein, label) order, L = label rows per protein on this rank; targets [rows] fp32.
 *   PN_LOSS_BCE    torch.nn.BCEWithLogitsLoss(pos_weight) as built by protnote/utils/losses.py:270-272 ('BCE');
 *                  pos_weight [L] or NULL.
 *   PN_LOSS_FOCAL  protnote/utils/losses.py:171-213 FocalLoss(alpha, gamma, label_smoothing): BCE of the smoothed target,
 *                  (1 - exp(-BCE))^gamma modulation, alpha_t weighting when alpha >= 0.  pos_weight must be NULL.
 * For reduction 'mean' over a label-sharded batch pass grad_scale = 1 / (B * L_total) and divide loss_sum by B * L_total. */
#define PN_LOSS_BCE 1
#define PN_LOSS_FOCAL 2
int pn_t_bn_relu_dot_loss(const void* z_hi, const void* z_lo, long long rows, int cols, long long ld_z, const float* state,
                          const float* w, const float* b, float* out, const float* targets, long long L,
                          const float* pos_weight, int loss_kind, float gamma, float alpha, float label_smoothing,
                          float grad_scale, float* g_out, double* loss_sum, void* stream);
/* layer 1 of the pair scorer: h[(b, l)] = relu((a[b] + c[l]) * scale + shift), rows b * L + l; a [B][H], c [L][H] dense */
int pn_t_pair_hidden(const float* a, long long B, const float* c, long long L, int H, const float* state, void* hi,
                     void* lo, long long ld, void* hiT, void* loT, long long blocksT, void* stream);

/* Source of a BatchNorm+ReLU backward.  kind 0: g planes, z planes.  kind 1: g = g_logit[r] * w[n] (gradient of the final
 * Linear(H -> 1), generated on the fly), z planes.  kind 2: g planes, z[r] = a[r / L] + c[r % L] (layer 1). */
/* FEATURE_FUSION concatenation_prod in training (ProtNote.py:140-150): the product block of layer 1 is a real GEMM.
 *   pn_t_pair_product   q[b*L + l] = p[b] (.) t[l]   (p [B][d], t [L][d] fp32) as planes (+ K-blocked transposed planes)
 *   pn_t_pair_add       z1[b*L + l] = x[b*L + l] + a[b] + c[l]   (x planes [B*L][H], a [B][H], c [L][H] fp32) as planes
 *   pn_t_pair_marginals out_b[b] = (1/g_sc) sum_l G[b*L + l] (.) wl[l],  out_l[l] = (1/g_sc) sum_b G[b*L + l] (.) wb[b]
 *                       (G planes [B*L][cols] carrying the device scale g_sc, nullable; wl [L][cols] / wb [B][cols] fp32,
 *                       null = ones; either output nullable): the two marginals of a pair-grid gradient. */
int pn_t_pair_product(const float* p, long long B, const float* t, long long L, int d, void* hi, void* lo, long long ld,
                      void* hiT, void* loT, long long blocksT, void* stream);
int pn_t_pair_add(const void* x_hi, const void* x_lo, long long ld_x, const float* a, long long B, const float* c,
                  long long L, int H, void* hi, void* lo, long long ld, void* stream);
int pn_t_pair_marginals(const void* g_hi, const void* g_lo, long long ld_g, const float* g_sc, long long B, long long L,
                        int cols, const float* wb, const float* wl, float* out_b, float* out_l, void* stream);

/* FEATURE_FUSION similarity in training (ProtNote.py:281-284; logits = normalize(P_e) normalize(L_e)^T / temperature, the
 * matrix product itself is pn_t_gemm).  y = x * scale / max(|x|_2, 1e-12) per row of a contiguous fp32 [rows][cols] matrix
 * (torch.nn.functional.normalize's eps), inv_norm[r] = 1 / max(|x|_2, 1e-12); backward dx = scale * inv_norm *
 * (dy - u (u . dy)) with u = y / scale. */
int pn_t_normalize_rows(const float* x, long long rows, int cols, float scale, float* y, float* inv_norm, void* stream);
int pn_t_normalize_rows_bwd(const float* y, const float* inv_norm, const float* dy, long long rows, int cols, float scale,
                            float* dx, void* stream);

typedef struct pn_bwd_src {
  int kind;
  long long rows;
  int cols;
  const void* g_hi; const void* g_lo; long long ld_g; const float* g_sc;
  const float* g_logit; const float* w;
  const void* z_hi; const void* z_lo; long long ld_z;
  const float* a; const float* c; long long L;
  const float* state;
} pn_bwd_src;
/* sums[0][c] = sum_r g_y, sums[1][c] = sum_r g_y * xhat  with g_y = g * [z*scale+shift > 0], xhat = (z - mean) * invstd
 * (true scale, fp64 [2][cols], overwritten) = the gradients of BatchNorm's beta and gamma over these rows;
 * maxes[2] = max|g| (before the mask: a bound of max|g_y|), max|z| (pn_t_bwd_scale turns it into a bound of max|xhat|),
 * both read off the hi planes; kind 1 only: dw[c] = sum_r g_logit[r] * relu(z*scale+shift)[r][c], db = sum_r g_logit[r]
 * (gradients of the final Linear). */
/* kind 2 only, optional by-product: gyl[l][n] = sum_b g_y[b,l,n] (fp32 [L][cols]); with it pn_t_bwd_apply_pair derives dc
 * analytically instead of reading g a second time. */
int pn_t_bwd_stats(const pn_bwd_src* src, double* sums, float* maxes, double* dw, double* db, float* gyl, void* stream);
/* means[2][cols] (fp32) = sums / count, the two per-column means pass 2 subtracts; sc_out (nullable) = power-of-two scale
 * for the g_z tensor pn_t_bwd_apply will write (from an upper bound of |g_z|).  `sums` are the totals over ALL rows of
 * the batch (all-reduced over ranks when the rows are sharded), count = that number of rows. */
int pn_t_bwd_scale(const double* sums, const float* maxes, const float* state, double count, int cols, float* sc_out,
                   float* means, void* stream);
/* g_z = scale * (g_y - means[0] - xhat * means[1]) * sc_out as planes (+ transposed planes) */
int pn_t_bwd_apply(const pn_bwd_src* src, const float* means, const float* sc_out, void* hi, void* lo, long long ld,
                   void* hiT, void* loT, long long blocksT, void* stream);
/* kind 2 only: the same g_z reduced on the fly to da[b] = sum_l g_z[b, l] (fp64 scratch da64 [B][H] -> fp32 da) and
 * dc[l] = sum_b g_z[b, l] (fp32 [L][H]), true scale. */
/* gyl (from pn_t_bwd_stats) and a_stats (pn_t_col_stats of a: fp64 [2][cols]) are optional: when both are given
 * dc[l] = scale * (gyl[l] - B m1 - m2 * invstd * (sum_b a[b] + B c[l] - B mean)) without a pass over g. */
int pn_t_bwd_apply_pair(const pn_bwd_src* src, const float* means, long long B, const float* gyl, const double* a_stats,
                        double* da64, float* da, float* dc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROTNOTE_B200_H */
