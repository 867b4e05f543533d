// C ABI + host-side orchestration of the B200-native ProtNote scoring path (see include/protnote_b200.h).
// Everything here is launch logic: buffer carving, TMA descriptors, kernel sequencing.  No device allocation,
// no synchronisation, no CPU arithmetic on tensor data.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/protnote_b200.h"
#include "pn_kernels.cuh"
#include "pn_gemm2.cuh"
#include "pn_train.cuh"

namespace {

using namespace pn;

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
int g_bk_option = 0;             // 0 = auto (strict: 32, fast: 64)
long long g_chunk_rows = 0;      // 0 = auto
// 1: build layer 1 of the pair scorer inside the first GEMM's operand producer (no HBM round trip of h1).
// Measured on B200 (profiles/r01_fused_generator_probe.txt): correct, but 18 % SLOWER than the separate kernel, because
// the generator's extra shared-memory traffic (+32 KB per k-block on top of ~120 KB) competes with the tensor core's
// operand reads; it needs the 2-CTA operand sharing planned for the next round to pay off.  Off by default.
int g_fuse_features = 0;
// run the engine as CTA pairs (tcgen05 cta_group::2, 256-row tiles, each CTA loads half of the weight tile):
// 2 always (experiments), 1 (default) wherever the caller allows it and the problem has more than one 128-row tile,
// 0 never, -1 in fast mode only.  Round 1 measured strict-mode pairs 2 % SLOWER (profiles/r01_cta_pair_probe.txt) - but
// that engine was bound by its single MMA-issuing thread, not by the tensor core; with the issue loop fixed
// (profiles/r02_issue_thread_bound.txt) the SM's 64 B/clk operand ingest is what binds the 1-CTA kernel at tile widths
// below 256 and the pair, which halves the weight bytes per CTA, wins in strict mode too: encoder +8 %, scorer +2 %.
int g_cta2 = 1;
// threads along the columns of a reduction block (TX * TY == 256; TX * 8 consecutive columns per block row)
int g_stats_tx = 32;
// 1: strict-mode encoder convolutions keep the hi*hi products and the lo corrections in separate TMEM buffers
// (GemmParams::split_corr); the main chain then needs a promotion only every 64 K-elements and its drain is hidden
// behind the correction MMAs.  Measured on B200 (profiles/r01_split_corr_probe.txt): encoder 10 % faster (tensor pipe
// 50 % -> 57 % on the dilated conv), but the end-to-end logit error grows from 5.1e-5 to 7.2e-5 on base_small (bar:
// 1e-4).  The parity margin is worth more than 0.75 % of the headline step: off by default.
int g_split_corr = 0;
// strict mode: K elements accumulated in TMEM between fp32 promotions, per stage of the path
// (kStagePointwise: the 1x1 convolutions of the encoder's residual blocks, K = bottleneck width)
enum { kStageEncoder = 0, kStageHeads = 1, kStageScorer = 2, kStageOther = 3, kStagePointwise = 4, kNumStages = 5 };
int g_promote_k[kNumStages] = {64, 32, 256, 64, 64};
// strict mode: compensation of the accumulator's round-toward-zero bias (GemmParams::trunc_comp).  The factor applied to a
// launch is (c1 * K-elements per chunk + c0) * 1e-12; both coefficients are hardware properties measured with
// tools/trunc_comp_probe.py.  c1 == 0 and c0 == 0 switch the compensation off.
long long g_trunc_c1 = 0, g_trunc_c0 = 0;
// strict mode: per-K-position compensation folded into the packed weights (pack_weight_kernel, TruncComp), in units of
// 1e-12 per truncating add.  Weights must be re-packed after this or a promote_k / bk option changes.
long long g_trunc_beta_ppt = 33000;
// strict mode: W_p and the protein half of output layer 1 in fp64 on the CUDA cores (linear_f64_kernel); 0 = tensor-core
// path like every other layer.
int g_f64_protein_head = 1;
// pair kernel: L2 eviction-priority hints on the TMA loads (GemmParams::l2_hints).  Measured on B200 (16 x 32768 pairs,
// strict): no hints 43.4-45.2 ms, weights evict_last 45.0 ms, + activations evict_first 50.0-50.2 ms: the A tile that twelve
// clusters share is dropped before the last of them has read it.  Off by default (profiles/r02_ab_prefetch_pairfeatures.txt).
int g_l2_hints = 0;

// optional per-launch CUDA-event timing of the pair scorer's GEMM launches (bench.py's roofline numbers)
struct TimedLaunch {
  cudaEvent_t start, stop;
  double flops;
};
bool g_timing = false;
std::vector<TimedLaunch> g_timed;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define PN_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define PN_TRY(expr)          \
  do {                        \
    int r_ = (expr);          \
    if (r_ != 0) return r_;   \
  } while (0)

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point resolved at run time: the library links no libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int resolve_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return fail("cuTensorMapEncodeTiled not available");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// fp16 tensor, dims fastest-first.  strides_bytes[i] = stride of dim i+1.
int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box, int swizzle_bytes, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
  PN_TRY(resolve_encode());
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_32B;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail("TMA base %p not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (strides_bytes[i] % 16 != 0) return fail("TMA stride %llu not a multiple of 16", (unsigned long long)strides_bytes[i]);
  CUresult r = g_encode(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims,
                        strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu box %u,%u", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// engine launch
// ------------------------------------------------------------------------------------------------
struct Planes {             // an fp32 tensor carried as fp16 planes, row-major, K fastest
  const __half* hi = nullptr;
  const __half* lo = nullptr;
  long long rows = 0;       // plain: M or N.  conv: batch * T
  long long cols = 0;       // logical K extent (channels for conv)
  long long ld = 0;         // row pitch in elements (multiple of 8)
  bool kblocked = false;    // stored as [cols/64][rows][64] (K-blocked, see GemmParams::a_kblk); ld is unused
};

struct ConvView {           // A operand as [batch][T][C] with taps (taps == 0 -> plain matrix)
  int taps = 0, dil = 1, T = 0, batch = 0;
  int cpad = 0;             // K offset between taps in the packed weight rows
  const long long* lengths = nullptr;
};

struct Epilogue {
  const float* scale = nullptr; const float* shift = nullptr;
  const float* resid = nullptr; long long ld_resid = 0;
  float* out_f32 = nullptr; long long ld_out = 0;
  float* out_z = nullptr; long long ld_z = 0;
  const float* scale2 = nullptr; const float* shift2 = nullptr;
  int relu = 0;
  __half* out_hi = nullptr; __half* out_lo = nullptr; long long ld_split = 0;
  const float* dot_w = nullptr; float* dot_out = nullptr;
  int pair_nl = 0;
  const float* add_p = nullptr; long long ld_add_p = 0;
  const float* add_l = nullptr; long long ld_add_l = 0;
  // fused pair-feature A operand (A.hi may be null then): relu(gen_a[r / pair_nl] + gen_c[r % pair_nl])
  const float* gen_a = nullptr; long long ld_gen_a = 0;
  const float* gen_c = nullptr; long long ld_gen_c = 0;
};

int choose_bn(long long N) {
  if (N <= 256) return (int)round_up(N, 32);
  const int cand[] = {256, 224, 192};
  int best = 256;
  long long best_waste = 1LL << 60;
  for (int bn : cand) {
    const long long waste = round_up(N, bn) - N;
    if (waste < best_waste) {
      best_waste = waste;
      best = bn;
    }
  }
  return best;
}

int tiles_n_for(long long N) { return (int)((N + choose_bn(N) - 1) / choose_bn(N)); }

int pick_bk(int mode) {
  if (g_bk_option == 32 || g_bk_option == 64) return g_bk_option;
  return mode == PN_STRICT ? 32 : 64;
}

// k-blocks per accumulator chunk of a strict-mode launch of stage `stage_kind` (launch_gemm and pack_linear must agree)
int strict_chunk_kblocks(int stage_kind, int bk, int num_kblocks, int promote_override = -1) {
  const int promote_k = promote_override >= 0 ? promote_override : g_promote_k[stage_kind];
  int chunk = num_kblocks;
  if (promote_k > 0) {
    chunk = promote_k / bk > 0 ? promote_k / bk : 1;
    if (chunk > num_kblocks) chunk = num_kblocks;
  }
  return chunk;
}

// per-device caches (a process may drive more than one GPU: cuda:0 then cuda:1)
constexpr int kMaxDevices = 64;
int g_num_sms[kMaxDevices] = {0};
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
int num_sms() {
  const int dev = current_device();
  if (g_num_sms[dev] == 0) {
    cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms[dev] <= 0) g_num_sms[dev] = 148;
  }
  return g_num_sms[dev];
}
std::mutex g_timed_mutex;   // guards g_timed (the optional per-launch event list)

template <int BK, int NPASS, bool GEN = false>
int launch_gemm_t(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BK, NPASS, GEN>;
  // the attribute is per device: remember it per device, not per process
  static bool configured[kMaxDevices] = {false};
  if (!configured[current_device()]) {
    PN_CUDA(cudaFuncSetAttribute(gemm_kernel<BK, NPASS, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured[current_device()] = true;
  }
  const int total = p.tiles_m * p.tiles_n;
  const int grid = total < num_sms() ? total : num_sms();
  TimedLaunch tl;
  const bool timed = g_timing && p.timed_flops > 0;
  if (timed) {
    PN_CUDA(cudaEventCreate(&tl.start));
    PN_CUDA(cudaEventCreate(&tl.stop));
    tl.flops = p.timed_flops;
    PN_CUDA(cudaEventRecord(tl.start, stream));
  }
  gemm_kernel<BK, NPASS, GEN><<<grid, GEN ? kGenThreads : kGemmThreads, Cfg::kSmemBytes, stream>>>(p);
  PN_CUDA(cudaGetLastError());
  if (timed) {
    PN_CUDA(cudaEventRecord(tl.stop, stream));
    std::lock_guard<std::mutex> lock(g_timed_mutex);
    g_timed.push_back(tl);
  }
  g_launches++;
  return 0;
}

template <int BK, int NPASS>
int launch_gemm2_t(const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BK, NPASS>;
  static bool configured[kMaxDevices] = {false};
  if (!configured[current_device()]) {
    PN_CUDA(cudaFuncSetAttribute(gemm2_kernel<BK, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured[current_device()] = true;
  }
  const int total = p.tiles_m * p.tiles_n;
  int clusters = num_sms() / 2;
  if (total < clusters) clusters = total;
  TimedLaunch tl;
  const bool timed = g_timing && p.timed_flops > 0;
  if (timed) {
    PN_CUDA(cudaEventCreate(&tl.start));
    PN_CUDA(cudaEventCreate(&tl.stop));
    tl.flops = p.timed_flops;
    PN_CUDA(cudaEventRecord(tl.start, stream));
  }
  gemm2_kernel<BK, NPASS><<<2 * clusters, kGemmThreads, Cfg::kSmemBytes, stream>>>(p);
  PN_CUDA(cudaGetLastError());
  if (timed) {
    PN_CUDA(cudaEventRecord(tl.stop, stream));
    std::lock_guard<std::mutex> lock(g_timed_mutex);
    g_timed.push_back(tl);
  }
  g_launches++;
  return 0;
}

// D = A * B^T with the fused epilogue.  A: plain [M][K] or conv view; B: packed weights [N][K_total].
int launch_gemm(const Planes& A, const ConvView& cv, const Planes& B, long long N, const Epilogue& e, int mode,
                cudaStream_t stream, int stage_kind = kStageOther, int promote_override = -1, bool allow_pairs = true) {
  if (mode != PN_STRICT && mode != PN_FAST) return fail("mode must be PN_STRICT or PN_FAST");
  const bool gen = e.gen_a != nullptr;
  // CTA pairs pay off from two 128-row tiles up (a lone tile would run half empty)
  const long long rows_per_problem = cv.taps > 0 ? cv.T : A.rows;
  const bool cta2 = (g_cta2 == 2 || ((g_cta2 == 1 || (g_cta2 < 0 && mode == PN_FAST)) && allow_pairs && rows_per_problem > kBM)) &&
                    !gen && !A.kblocked && !B.kblocked;
  const int tile_rows = cta2 ? 2 * kBM : kBM;
  const int bk = gen ? 32 : pick_bk(mode);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.N = (int)N;
  p.bn = choose_bn(N);
  p.tiles_n = (int)((N + p.bn - 1) / p.bn);
  const bool need_lo = mode == PN_STRICT;
  if (need_lo && ((!gen && A.lo == nullptr) || B.lo == nullptr)) return fail("strict mode needs lo planes");
  if (cv.taps > 0) {
    const int cblocks = (int)((A.cols + bk - 1) / bk);
    p.conv_taps = cv.taps;
    p.conv_cblocks = cblocks;
    p.conv_cpad = cv.cpad;
    p.conv_dil = cv.dil;
    p.conv_T = cv.T;
    p.conv_tiles_per_seq = (cv.T + tile_rows - 1) / tile_rows;
    p.lengths = cv.lengths;
    p.M = cv.batch * cv.T;
    p.tiles_m = cv.batch * p.conv_tiles_per_seq;
    p.num_kblocks = cv.taps * cblocks;
    const cuuint64_t dims[3] = {(cuuint64_t)A.cols, (cuuint64_t)cv.T, (cuuint64_t)cv.batch};
    const cuuint64_t strides[2] = {(cuuint64_t)A.ld * 2, (cuuint64_t)A.ld * 2 * cv.T};
    const cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)kBM, 1};
    PN_TRY(make_map(&p.tm_a_hi, A.hi, 3, dims, strides, box, bk * 2));
    if (need_lo) PN_TRY(make_map(&p.tm_a_lo, A.lo, 3, dims, strides, box, bk * 2));
  } else {
    p.M = (int)A.rows;
    p.tiles_m = (int)((A.rows + tile_rows - 1) / tile_rows);
    p.num_kblocks = (int)((A.cols + bk - 1) / bk);
    if (!gen && A.kblocked) {
      const cuuint64_t dims[3] = {64, (cuuint64_t)A.rows, (cuuint64_t)((A.cols + 63) / 64)};
      const cuuint64_t strides[2] = {128, (cuuint64_t)A.rows * 128};
      const cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)kBM, 1};
      PN_TRY(make_map(&p.tm_a_hi, A.hi, 3, dims, strides, box, bk * 2));
      if (need_lo) PN_TRY(make_map(&p.tm_a_lo, A.lo, 3, dims, strides, box, bk * 2));
      p.a_kblk = 1;
    } else if (!gen) {
      const cuuint64_t dims[2] = {(cuuint64_t)A.cols, (cuuint64_t)A.rows};
      const cuuint64_t strides[1] = {(cuuint64_t)A.ld * 2};
      const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)kBM};
      PN_TRY(make_map(&p.tm_a_hi, A.hi, 2, dims, strides, box, bk * 2));
      if (need_lo) PN_TRY(make_map(&p.tm_a_lo, A.lo, 2, dims, strides, box, bk * 2));
    }
  }
  if (A.rows >= (1LL << 31) || p.tiles_m <= 0 || p.num_kblocks <= 0) return fail("bad GEMM shape");
  p.chunk_kblocks = p.num_kblocks;
  // algorithmic FLOPs of this launch (2*M*N*K), recorded only for the pair scorer's GEMMs
  p.timed_flops = stage_kind == kStageScorer ? 2.0 * (double)A.rows * (double)N * (double)A.cols : 0.0;
  // promote_override > 0 forces a promotion period in either mode (wgrad: K = rows of the batch, far too long a chain
  // for the truncating tensor-core accumulator even at fp16 operand precision)
  const int promote_k = promote_override >= 0 ? promote_override : g_promote_k[stage_kind];
  if (mode == PN_STRICT || promote_override > 0) p.chunk_kblocks = strict_chunk_kblocks(stage_kind, bk, p.num_kblocks, promote_override);
  if (g_split_corr && mode == PN_STRICT && stage_kind == kStageEncoder && !gen && !cta2 && bk == 32 && promote_override < 0) {
    // main chain: 2 truncating adds per k-block instead of 6 -> twice the K per promotion at fewer adds per chunk (4 vs 6)
    p.split_corr = 1;
    p.chunk_kblocks = 2 * promote_k / bk > 0 ? 2 * promote_k / bk : 1;
    if (p.chunk_kblocks > 3) p.chunk_kblocks = 3;      // the chunk's operand stages stay resident: < pipeline depth (4)
    if (p.chunk_kblocks > p.num_kblocks) p.chunk_kblocks = p.num_kblocks;
  }
  if (mode == PN_STRICT && !p.split_corr && (g_trunc_c1 != 0 || g_trunc_c0 != 0)) {
    const double kc = (double)p.chunk_kblocks * bk;
    const double f = ((double)g_trunc_c1 * kc + (double)g_trunc_c0) * 1e-12;
    p.trunc_comp = f > 0 ? (float)f : 0.f;
  }
  if (B.kblocked) {
    const cuuint64_t dims[3] = {64, (cuuint64_t)B.rows, (cuuint64_t)((B.cols + 63) / 64)};
    const cuuint64_t strides[2] = {128, (cuuint64_t)B.rows * 128};
    const cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)p.bn, 1};
    PN_TRY(make_map(&p.tm_b_hi, B.hi, 3, dims, strides, box, bk * 2));
    if (need_lo) PN_TRY(make_map(&p.tm_b_lo, B.lo, 3, dims, strides, box, bk * 2));
    p.b_kblk = 1;
  } else {
    const cuuint64_t dims[2] = {(cuuint64_t)B.cols, (cuuint64_t)B.rows};
    const cuuint64_t strides[1] = {(cuuint64_t)B.ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)(cta2 ? p.bn / 2 : p.bn)};
    PN_TRY(make_map(&p.tm_b_hi, B.hi, 2, dims, strides, box, bk * 2));
    if (need_lo) PN_TRY(make_map(&p.tm_b_lo, B.lo, 2, dims, strides, box, bk * 2));
  }
  p.l2_hints = g_l2_hints;
  p.scale = e.scale; p.shift = e.shift;
  p.resid = e.resid; p.ld_resid = e.ld_resid;
  p.out_f32 = e.out_f32; p.ld_out = e.ld_out;
  p.out_z = e.out_z; p.ld_z = e.ld_z;
  p.scale2 = e.scale2; p.shift2 = e.shift2;
  p.relu = e.relu;
  p.out_hi = e.out_hi; p.out_lo = need_lo ? e.out_lo : nullptr; p.ld_split = e.ld_split;
  p.dot_w = e.dot_w; p.dot_out = e.dot_out;
  p.pair_nl = e.pair_nl;
  p.add_p = e.add_p; p.ld_add_p = e.ld_add_p;
  p.add_l = e.add_l; p.ld_add_l = e.ld_add_l;
  p.gen_a = e.gen_a; p.ld_gen_a = e.ld_gen_a;
  if (gen) {
    if (e.pair_nl % kBM != 0 || A.cols % 32 != 0) return fail("generated A operand needs pair_nl %% 128 == 0 and K %% 32 == 0");
    const cuuint64_t dims[2] = {(cuuint64_t)A.cols, (cuuint64_t)e.pair_nl};
    const cuuint64_t strides[1] = {(cuuint64_t)e.ld_gen_c * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)kBM};
    PN_TRY(make_map(&p.tm_gen_c, e.gen_c, 2, dims, strides, box, 128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32));
  }
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vec_out = e.out_f32 && aligned16(e.out_f32) && e.ld_out % 4 == 0;
  p.vec_z = e.out_z && aligned16(e.out_z) && e.ld_z % 4 == 0;
  p.vec_resid = e.resid && aligned16(e.resid) && e.ld_resid % 4 == 0;
  p.vec_split = e.out_hi && aligned16(e.out_hi) && (!p.out_lo || aligned16(p.out_lo)) && e.ld_split % 8 == 0;
  if (gen) {
    if (cv.taps > 0 || e.pair_nl <= 0) return fail("generated A operand needs plain addressing and pair_nl");
    return mode == PN_STRICT ? launch_gemm_t<32, 3, true>(p, stream) : launch_gemm_t<32, 1, true>(p, stream);
  }
  if (cta2) {
    if (bk == 32) return mode == PN_STRICT ? launch_gemm2_t<32, 3>(p, stream) : launch_gemm2_t<32, 1>(p, stream);
    return mode == PN_STRICT ? launch_gemm2_t<64, 3>(p, stream) : launch_gemm2_t<64, 1>(p, stream);
  }
  if (bk == 32) return mode == PN_STRICT ? launch_gemm_t<32, 3>(p, stream) : launch_gemm_t<32, 1>(p, stream);
  return mode == PN_STRICT ? launch_gemm_t<64, 3>(p, stream) : launch_gemm_t<64, 1>(p, stream);
}

inline int ew_grid(long long work_items, int block = 256) {
  long long g = (work_items + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------
// buffer carving
// ------------------------------------------------------------------------------------------------
struct Arena {
  char* base;
  size_t size, off = 0;
  Arena(void* b, size_t s) : base(static_cast<char*>(b)), size(s) {}
  size_t take(size_t bytes) {   // returns offset (valid even when base == nullptr: used for sizing)
    const size_t o = off;
    off = (size_t)round_up((long long)(off + bytes), 256);
    return o;
  }
  template <typename T>
  T* at(size_t o) const { return reinterpret_cast<T*>(base + o); }
  bool ok() const { return off <= size; }
};

struct PackedLinear {         // offsets into the packed arena
  int N = 0, cin = 0, taps = 1, cpad = 0, ld = 0;
  size_t hi = 0, lo = 0, wscale = 0, absmax = 0, scale = 0, shift = 0;
};

PackedLinear carve_linear(Arena& ar, int N, int cin, int taps) {
  PackedLinear pl;
  pl.N = N;
  pl.cin = cin;
  pl.taps = taps;
  pl.cpad = (int)round_up(cin, 64);
  pl.ld = pl.cpad * taps;
  pl.hi = ar.take((size_t)N * pl.ld * 2);
  pl.lo = ar.take((size_t)N * pl.ld * 2);
  pl.wscale = ar.take(4);
  pl.absmax = ar.take(4);
  pl.scale = ar.take((size_t)N * 4);
  pl.shift = ar.take((size_t)N * 4);
  return pl;
}

Planes weight_planes(const Arena& ar, const PackedLinear& pl) {
  Planes B;
  B.hi = ar.at<__half>(pl.hi);
  B.lo = ar.at<__half>(pl.lo);
  B.rows = pl.N;
  B.cols = pl.ld;
  B.ld = pl.ld;
  return B;
}

// w laid out (N, cin, taps) with element strides (sn, sc, st); optional second matrix added with `sign2`
// is not needed: concatenation_diff is folded by two packs into separate buffers (see pack_scorer).
int pack_linear(const Arena& ar, const PackedLinear& pl, const float* w, long long sn, long long sc, long long st,
                long long span, const float* bias, const float* gamma, const float* beta, const float* mean,
                const float* var, float eps, cudaStream_t stream, int stage_kind) {
  // the K walk the strict-mode engine will use for this layer (see launch_gemm)
  TruncComp tc;
  tc.bk = pick_bk(PN_STRICT);
  tc.cblocks = (pl.cin + tc.bk - 1) / tc.bk;
  tc.num_kblocks = pl.taps * tc.cblocks;
  tc.chunk_kblocks = strict_chunk_kblocks(stage_kind, tc.bk, tc.num_kblocks);
  tc.beta = (float)((double)g_trunc_beta_ppt * 1e-12);
  unsigned* am = ar.at<unsigned>(pl.absmax);
  PN_CUDA(cudaMemsetAsync(am, 0, 4, stream));
  // absmax over the rows' used span (row n covers w[n*sn .. n*sn + span))
  absmax_kernel<<<ew_grid((long long)pl.N * span), 256, 0, stream>>>(w, pl.N, span, sn, am);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  pack_weight_kernel<<<ew_grid((long long)pl.N * pl.ld), 256, 0, stream>>>(
      w, pl.N, pl.cin, pl.taps, sn, sc, st, pl.cpad, pl.ld, am, ar.at<float>(pl.wscale), ar.at<__half>(pl.hi),
      ar.at<__half>(pl.lo), tc);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  fold_affine_kernel<<<(pl.N + 255) / 256, 256, 0, stream>>>(pl.N, bias, gamma, beta, mean, var, eps,
                                                             ar.at<float>(pl.wscale), ar.at<float>(pl.scale),
                                                             ar.at<float>(pl.shift));
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------------
struct EncoderLayout {
  PackedLinear conv1;
  struct Block {
    size_t bn1_scale, bn1_shift;
    PackedLinear conv_d, conv_p;
  };
  std::vector<Block> blocks;
  size_t bytes = 0;
};

EncoderLayout encoder_layout(const pn_encoder_cfg& c) {
  Arena ar(nullptr, 0);
  EncoderLayout L;
  L.conv1 = carve_linear(ar, c.channels, c.input_channels, c.kernel_size);
  for (int i = 0; i < c.num_blocks; ++i) {
    EncoderLayout::Block b;
    b.bn1_scale = ar.take((size_t)c.channels * 4);
    b.bn1_shift = ar.take((size_t)c.channels * 4);
    b.conv_d = carve_linear(ar, c.bottleneck, c.channels, c.kernel_size);
    b.conv_p = carve_linear(ar, c.channels, c.bottleneck, 1);
    L.blocks.push_back(b);
  }
  L.bytes = ar.off;
  return L;
}

int check_encoder_cfg(const pn_encoder_cfg* c) {
  if (!c) return fail("null encoder cfg");
  if (c->input_channels < 1 || c->channels < 1 || c->bottleneck < 1 || c->num_blocks < 0 || c->dilation_base < 1)
    return fail("bad encoder cfg");
  if (c->kernel_size < 1 || c->kernel_size % 2 == 0) return fail("kernel_size must be odd (got %d)", c->kernel_size);
  return 0;
}

struct EncoderWs {
  size_t in_hi, in_lo, x, act_hi, act_lo, hid_hi, hid_lo;
  int cin_pad, ldc, ldb;
  size_t bytes;
};

EncoderWs encoder_ws(const pn_encoder_cfg& c, long long batch, long long T) {
  Arena ar(nullptr, 0);
  EncoderWs w;
  w.cin_pad = (int)round_up(c.input_channels, 64);
  w.ldc = (int)round_up(c.channels, 64);
  w.ldb = (int)round_up(c.bottleneck, 64);
  const size_t pos = (size_t)batch * T;
  w.in_hi = ar.take(pos * w.cin_pad * 2);
  w.in_lo = ar.take(pos * w.cin_pad * 2);
  w.x = ar.take(pos * w.ldc * 4);
  w.act_hi = ar.take(pos * w.ldc * 2);
  w.act_lo = ar.take(pos * w.ldc * 2);
  w.hid_hi = ar.take(pos * w.ldb * 2);
  w.hid_lo = ar.take(pos * w.ldb * 2);
  w.bytes = ar.off;
  return w;
}

// ------------------------------------------------------------------------------------------------
// scorer
// ------------------------------------------------------------------------------------------------
struct ScorerLayout {
  std::vector<PackedLinear> wp, wl;      // projection heads
  PackedLinear l1_p, l1_l, l1_x;         // output layer 1 split over [p; t; (p*t)]
  size_t l1_shift = 0;                   // folded BN1 shift (+ bias) applied on the protein side
  std::vector<PackedLinear> hidden;      // output hidden layers 2..n
  size_t w_out = 0, b_out = 0;           // final Linear(H -> 1)
  size_t tmp = 0;                        // scratch for the concatenation_diff weight fold [H][latent]
  // fp64 protein-side head (strict mode): original fp32 weights of W_p and of the protein half of output layer 1, and
  // their BatchNorm folds in fp64 (linear_f64_kernel)
  std::vector<size_t> wp_raw, wp_scale64, wp_shift64;
  size_t l1p_raw = 0, l1p_scale64 = 0, l1p_shift64 = 0;
  size_t bytes = 0;
};

ScorerLayout scorer_layout(const pn_scorer_cfg& c) {
  Arena ar(nullptr, 0);
  ScorerLayout L;
  for (int head = 0; head < 2; ++head) {
    int in_dim = head == 0 ? c.protein_dim : c.label_dim;
    for (int i = 0; i < c.proj_layers; ++i) {
      const int out_dim = i == c.proj_layers - 1 ? c.latent_dim : c.proj_hidden;
      (head == 0 ? L.wp : L.wl).push_back(carve_linear(ar, out_dim, in_dim, 1));
      in_dim = out_dim;
    }
  }
  {
    int in_dim = c.protein_dim;
    for (int i = 0; i < c.proj_layers; ++i) {
      const int out_dim = i == c.proj_layers - 1 ? c.latent_dim : c.proj_hidden;
      L.wp_raw.push_back(ar.take((size_t)out_dim * in_dim * 4));
      L.wp_scale64.push_back(ar.take((size_t)out_dim * 8));
      L.wp_shift64.push_back(ar.take((size_t)out_dim * 8));
      in_dim = out_dim;
    }
  }
  if (c.fusion == PN_FUSION_SIMILARITY) {
    L.bytes = ar.off;
    return L;
  }
  L.l1p_raw = ar.take((size_t)c.out_hidden * c.latent_dim * 4);
  L.l1p_scale64 = ar.take((size_t)c.out_hidden * 8);
  L.l1p_shift64 = ar.take((size_t)c.out_hidden * 8);
  L.l1_p = carve_linear(ar, c.out_hidden, c.latent_dim, 1);
  L.l1_l = carve_linear(ar, c.out_hidden, c.latent_dim, 1);
  if (c.fusion == PN_FUSION_CONCAT_PROD) L.l1_x = carve_linear(ar, c.out_hidden, c.latent_dim, 1);
  L.l1_shift = ar.take((size_t)c.out_hidden * 4);
  for (int j = 1; j < c.out_layers; ++j) L.hidden.push_back(carve_linear(ar, c.out_hidden, c.out_hidden, 1));
  L.w_out = ar.take((size_t)c.out_hidden * 4);
  L.b_out = ar.take(4);
  if (c.fusion == PN_FUSION_CONCAT_DIFF) L.tmp = ar.take((size_t)c.out_hidden * c.latent_dim * 4 * 2);
  L.bytes = ar.off;
  return L;
}

int check_scorer_cfg(const pn_scorer_cfg* c) {
  if (!c) return fail("null scorer cfg");
  if (c->protein_dim < 1 || c->label_dim < 1 || c->latent_dim < 1 || c->proj_hidden < 1 ||
      (c->fusion != PN_FUSION_SIMILARITY && c->out_hidden < 1))
    return fail("bad scorer dims");
  if (c->proj_layers < 1) return fail("proj_layers must be >= 1");
  if (c->fusion < 0 || c->fusion > 3) return fail("unknown fusion %d", c->fusion);
  if (c->fusion != PN_FUSION_SIMILARITY && c->out_layers < 1) return fail("out_layers must be >= 1 (got %d)", c->out_layers);
  if (c->descriptions_per_label < 1) return fail("descriptions_per_label must be >= 1");
  return 0;
}

__global__ void combine_kernel(const float* __restrict__ a, const float* __restrict__ b, float sign, long long rows,
                               int cols, long long lda, long long ldb, float* __restrict__ out) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int cidx = (int)(i % cols);
    out[i] = a[r * lda + cidx] + sign * b[r * ldb + cidx];
  }
}

__global__ void copy_floats_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// fp64 layer of the protein-side head: the skinny kernel up to 64 rows, the tiled one above
void launch_linear_f64(const double* x, long long M, int K, long long ldx, const float* w, int N, long long ldw,
                       const double* scale, const double* shift, int relu, double* y64, long long ldy64, float* y32,
                       long long ldy32, cudaStream_t stream) {
  if (M <= 64) {
    const dim3 grid((unsigned)((N + 7) / 8), (unsigned)((M + 7) / 8));
    linear_f64_skinny_kernel<<<grid, 256, 0, stream>>>(x, M, K, ldx, w, N, ldw, scale, shift, relu, y64, ldy64, y32, ldy32);
  } else {
    const dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
    linear_f64_kernel<<<grid, 256, 0, stream>>>(x, M, K, ldx, w, N, ldw, scale, shift, relu, y64, ldy64, y32, ldy32);
  }
}

// One projection head (W_p or W_l) followed by its half of output layer 1.
int run_projection(const pn_scorer_cfg& c, const ScorerLayout& L, const Arena& pk, bool protein, const float* in,
                   long long n, float* emb_out, float* half_out, void* workspace, size_t workspace_bytes, int mode,
                   cudaStream_t stream) {
  const std::vector<PackedLinear>& head = protein ? L.wp : L.wl;
  const int in_dim = protein ? c.protein_dim : c.label_dim;
  if (protein && mode == PN_STRICT && g_f64_protein_head) {
    // fp64 path (see linear_f64_kernel): two ping-pong fp64 activation buffers per row chunk
    const long long ld64 = round_up(c.proj_hidden > c.latent_dim ? c.proj_hidden : c.latent_dim, 64);
    const long long ldi = round_up(in_dim, 64);
    const size_t per_row64 = (size_t)(ldi + 2 * ld64) * 8;
    long long chunk = (long long)(workspace_bytes / per_row64);
    if (chunk <= 0) return fail("projection workspace too small (%zu bytes)", workspace_bytes);
    for (long long r0 = 0; r0 < n; r0 += chunk) {
      const long long rows = (n - r0) < chunk ? (n - r0) : chunk;
      Arena ws(workspace, workspace_bytes);
      double* x64 = ws.at<double>(ws.take((size_t)rows * ldi * 8));
      double* buf[2] = {ws.at<double>(ws.take((size_t)rows * ld64 * 8)), ws.at<double>(ws.take((size_t)rows * ld64 * 8))};
      if (!ws.ok()) return fail("projection workspace accounting error");
      f32_to_f64_rows_kernel<<<ew_grid(rows * ldi), 256, 0, stream>>>(in + r0 * in_dim, rows, in_dim, in_dim, x64, ldi);
      g_launches++;
      PN_CUDA(cudaGetLastError());
      const double* cur = x64;
      long long ldc = ldi;
      int K = in_dim, q = 0;
      for (size_t i = 0; i < head.size(); ++i) {
        const bool last = i + 1 == head.size();
        const int N = head[i].N;
        launch_linear_f64(cur, rows, K, ldc, pk.at<float>(L.wp_raw[i]), N, K, last ? nullptr : pk.at<double>(L.wp_scale64[i]),
                          last ? nullptr : pk.at<double>(L.wp_shift64[i]), last ? 0 : 1, buf[q], ld64,
                          (last && emb_out) ? emb_out + r0 * c.latent_dim : nullptr, c.latent_dim, stream);
        g_launches++;
        PN_CUDA(cudaGetLastError());
        cur = buf[q];
        ldc = ld64;
        K = N;
        q ^= 1;
      }
      if (half_out) {
        launch_linear_f64(cur, rows, K, ldc, pk.at<float>(L.l1p_raw), c.out_hidden, K, pk.at<double>(L.l1p_scale64),
                          pk.at<double>(L.l1p_shift64), 0, nullptr, 0, half_out + r0 * c.out_hidden, c.out_hidden, stream);
        g_launches++;
        PN_CUDA(cudaGetLastError());
      }
    }
    return 0;
  }
  const int ld_in = (int)round_up(in_dim, 64);
  const int ld_h = (int)round_up(c.proj_hidden > c.latent_dim ? c.proj_hidden : c.latent_dim, 64);
  const size_t per_row = (size_t)ld_in * 4 + (size_t)ld_h * 4 * 2 + 1024;
  long long chunk = (long long)(workspace_bytes / per_row);
  chunk = chunk / kBM * kBM;
  if (chunk <= 0) return fail("projection workspace too small (%zu bytes)", workspace_bytes);
  for (long long r0 = 0; r0 < n; r0 += chunk) {
    const long long rows = (n - r0) < chunk ? (n - r0) : chunk;
    Arena ws(workspace, workspace_bytes);
    __half* in_hi = ws.at<__half>(ws.take((size_t)rows * ld_in * 2));
    __half* in_lo = ws.at<__half>(ws.take((size_t)rows * ld_in * 2));
    __half* buf_hi[2];
    __half* buf_lo[2];
    for (int q = 0; q < 2; ++q) {
      buf_hi[q] = ws.at<__half>(ws.take((size_t)rows * ld_h * 2));
      buf_lo[q] = ws.at<__half>(ws.take((size_t)rows * ld_h * 2));
    }
    if (!ws.ok()) return fail("projection workspace accounting error");
    split_rows_kernel<<<ew_grid(rows * (ld_in / 8)), 256, 0, stream>>>(in + r0 * in_dim, rows, in_dim, in_dim, in_hi,
                                                                       in_lo, ld_in);
    g_launches++;
    PN_CUDA(cudaGetLastError());
    Planes A;
    A.hi = in_hi; A.lo = in_lo; A.rows = rows; A.cols = in_dim; A.ld = ld_in;
    int cur = 0;
    for (size_t i = 0; i < head.size(); ++i) {
      const PackedLinear& pl = head[i];
      const bool last = i + 1 == head.size();
      Epilogue e;
      e.scale = pk.at<float>(pl.scale);
      e.shift = pk.at<float>(pl.shift);
      e.relu = last ? 0 : 1;
      e.out_hi = buf_hi[cur]; e.out_lo = buf_lo[cur]; e.ld_split = ld_h;
      if (last && emb_out) {
        e.out_f32 = emb_out + r0 * c.latent_dim;
        e.ld_out = c.latent_dim;
      }
      PN_TRY(launch_gemm(A, ConvView(), weight_planes(pk, pl), pl.N, e, mode, stream, kStageHeads));
      A.hi = buf_hi[cur]; A.lo = buf_lo[cur]; A.rows = rows; A.cols = pl.N; A.ld = ld_h;
      cur ^= 1;
    }
    if (half_out == nullptr) continue;
    // half of output layer 1: protein side carries the folded BN1 shift, both sides carry its scale
    const PackedLinear& pl = protein ? L.l1_p : L.l1_l;
    Epilogue e;
    e.scale = pk.at<float>(pl.scale);
    e.shift = protein ? pk.at<float>(L.l1_shift) : nullptr;
    e.out_f32 = half_out + r0 * c.out_hidden;
    e.ld_out = c.out_hidden;
    PN_TRY(launch_gemm(A, ConvView(), weight_planes(pk, pl), pl.N, e, mode, stream, kStageHeads));
  }
  return 0;
}

size_t scorer_row_bytes(const pn_scorer_cfg& c) {
  const size_t ld_h = (size_t)round_up(c.out_hidden, 64);
  return ld_h * 2 /*bytes*/ * 2 /*planes*/ * 2 /*ping-pong*/ + (size_t)tiles_n_for(c.out_hidden) * 4 * 4;
}

}  // namespace

// helpers of the training primitives
namespace {
template <class P>
int launch_emit(const P& prod, long long rows, int cols, void* hi, void* lo, long long ld, void* hiT, void* loT,
                long long blocksT, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return fail("empty tensor");
  if (hi == nullptr) return fail("hi plane is required");
  if (ld % 8 != 0 || ld < cols) return fail("plane pitch %lld must be a multiple of 8 and >= cols %d", ld, cols);
  if (hiT && blocksT * kTileDim < rows) return fail("transposed planes hold %lld blocks of 64 rows, %lld rows need more", blocksT, rows);
  if (hiT && ((lo != nullptr) != (loT != nullptr))) return fail("transposed planes must carry the same planes as the row-major ones");
  const long long row_tiles = (rows + kTileDim - 1) / kTileDim;
  const long long strips = (row_tiles + kStripTiles - 1) / kStripTiles;
  const long long col_tiles = (ld + kTileDim - 1) / kTileDim;
  if (col_tiles > 65535) return fail("too many columns");
  const dim3 grid((unsigned)strips, (unsigned)col_tiles);
  if (lo)
    emit_tile_kernel<P, true><<<grid, 256, 0, stream>>>(prod, rows, cols, static_cast<__half*>(hi), static_cast<__half*>(lo),
                                                        ld, static_cast<__half*>(hiT), static_cast<__half*>(loT));
  else
    emit_tile_kernel<P, false><<<grid, 256, 0, stream>>>(prod, rows, cols, static_cast<__half*>(hi), nullptr, ld,
                                                         static_cast<__half*>(hiT), nullptr);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

// Rows per block of a column-reduction kernel.  The grid is sized to MANY rounds of resident blocks (about 64 blocks per
// SM in total) so that the last, partial round costs a few percent, not a whole extra round: with exactly
// "8 blocks per SM" the first version ran 4.01 rounds of its 296 resident blocks, i.e. five.
long long slab_rows(long long rows, int col_blocks) {
  long long slabs = ((long long)num_sms() * 64 + col_blocks - 1) / col_blocks;
  if (slabs < 1) slabs = 1;
  long long per = (rows + slabs - 1) / slabs;
  per = (per + 15) / 16 * 16;
  return per < 256 ? 256 : per;
}

BwdSrc to_src(const pn_bwd_src& s) {
  BwdSrc d;
  d.kind = s.kind; d.rows = s.rows; d.cols = s.cols;
  d.g_hi = static_cast<const __half*>(s.g_hi); d.g_lo = static_cast<const __half*>(s.g_lo); d.ld_g = s.ld_g; d.g_sc = s.g_sc;
  d.g_logit = s.g_logit; d.w = s.w;
  d.z_hi = static_cast<const __half*>(s.z_hi); d.z_lo = static_cast<const __half*>(s.z_lo); d.ld_z = s.ld_z;
  d.a = s.a; d.c = s.c; d.L = s.L;
  d.state = s.state;
  return d;
}

// strict-mode sources carry lo planes on every plane operand they have
bool src_has_lo(const BwdSrc& s) { return s.kind == 1 ? s.z_lo != nullptr : s.g_lo != nullptr; }

int check_src(const pn_bwd_src* s) {
  if (!s) return fail("null backward source");
  if (s->kind < 0 || s->kind > 2) return fail("unknown backward source kind %d", s->kind);
  if (s->rows <= 0 || s->cols <= 0 || !s->state) return fail("empty backward source");
  if (s->kind == 1 ? (!s->g_logit || !s->w) : (!s->g_hi || s->ld_g % 8 != 0)) return fail("backward source: gradient operand missing");
  if (s->kind == 2 ? (!s->a || !s->c || s->L <= 0 || s->rows % s->L != 0) : (!s->z_hi || s->ld_z % 8 != 0))
    return fail("backward source: pre-activation operand missing");
  if (s->kind == 0 && ((s->g_lo != nullptr) != (s->z_lo != nullptr)))
    return fail("backward source: g and z must both carry lo planes or neither");
  return 0;
}
}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int pn_version(void) { return 1; }

const char* pn_last_error(void) { return g_err.c_str(); }

long long pn_launch_count(void) { return g_launches.load(); }

int pn_device_check(int device) {
  cudaDeviceProp prop;
  PN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  return 0;
}

int pn_gemm_timing(int enable) {
  std::lock_guard<std::mutex> lock(g_timed_mutex);
  for (TimedLaunch& t : g_timed) {
    cudaEventDestroy(t.start);
    cudaEventDestroy(t.stop);
  }
  g_timed.clear();
  g_timing = enable != 0;
  return 0;
}

int pn_gemm_timing_read(double* total_ms, long long* launches, double* algorithmic_flops) {
  double ms = 0.0, fl = 0.0;
  std::lock_guard<std::mutex> lock(g_timed_mutex);
  for (TimedLaunch& t : g_timed) {
    PN_CUDA(cudaEventSynchronize(t.stop));
    float e = 0.f;
    PN_CUDA(cudaEventElapsedTime(&e, t.start, t.stop));
    ms += e;
    fl += t.flops;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = (long long)g_timed.size();
  if (algorithmic_flops) *algorithmic_flops = fl;
  return 0;
}

int pn_set_option(const char* name, long long value) {
  if (!name) return fail("null option name");
  if (strcmp(name, "bk") == 0) {
    if (value != 0 && value != 32 && value != 64) return fail("bk must be 0, 32 or 64");
    g_bk_option = (int)value;
    return 0;
  }
  if (strncmp(name, "promote_k", 9) == 0) {
    if (value < 0) return fail("promote_k must be >= 0 (0 = never promote)");
    const char* which = name + 9;
    if (*which == 0) {
      for (int i = 0; i < kNumStages; ++i) g_promote_k[i] = (int)value;
    } else if (strcmp(which, "_encoder") == 0) {
      g_promote_k[kStageEncoder] = (int)value;
    } else if (strcmp(which, "_heads") == 0) {
      g_promote_k[kStageHeads] = (int)value;
    } else if (strcmp(which, "_scorer") == 0) {
      g_promote_k[kStageScorer] = (int)value;
    } else if (strcmp(which, "_other") == 0) {
      g_promote_k[kStageOther] = (int)value;
    } else if (strcmp(which, "_pointwise") == 0) {
      g_promote_k[kStagePointwise] = (int)value;
    } else {
      return fail("unknown option '%s'", name);
    }
    return 0;
  }
  if (strcmp(name, "l2_hints") == 0) {
    if (value < 0 || value > 2) return fail("l2_hints must be 0, 1 or 2");
    g_l2_hints = (int)value;
    return 0;
  }
  if (strcmp(name, "f64_protein_head") == 0) {
    g_f64_protein_head = value != 0;
    return 0;
  }
  if (strcmp(name, "trunc_beta_ppt") == 0) {
    if (value < 0) return fail("trunc_beta_ppt must be >= 0");
    g_trunc_beta_ppt = value;
    return 0;
  }
  if (strcmp(name, "trunc_comp_c1") == 0) {
    g_trunc_c1 = value;
    return 0;
  }
  if (strcmp(name, "trunc_comp_c0") == 0) {
    g_trunc_c0 = value;
    return 0;
  }
  if (strcmp(name, "cta2") == 0) {
    g_cta2 = value < 0 ? -1 : (value > 2 ? 2 : (int)value);
    return 0;
  }
  if (strcmp(name, "stats_tx") == 0) {
    if (value != 32 && value != 64 && value != 128 && value != 256) return fail("stats_tx must be 32, 64, 128 or 256");
    g_stats_tx = (int)value;
    return 0;
  }
  if (strcmp(name, "split_corr") == 0) {
    g_split_corr = value != 0;
    return 0;
  }
  if (strcmp(name, "fuse_features") == 0) {
    g_fuse_features = value != 0;
    return 0;
  }
  if (strcmp(name, "chunk_rows") == 0) {
    if (value < 0) return fail("chunk_rows must be >= 0");
    g_chunk_rows = value;
    return 0;
  }
  return fail("unknown option '%s'", name);
}

// ---------------------------------------------------------------------------------------- encoder
size_t pn_encoder_packed_bytes(const pn_encoder_cfg* cfg) {
  if (check_encoder_cfg(cfg)) return 0;
  return encoder_layout(*cfg).bytes;
}

static int encoder_pack_impl(const pn_encoder_cfg* cfg, const float* const* params, int num_params, void* packed,
                             size_t packed_bytes, bool fold_bn, void* stream_);

int pn_encoder_pack(const pn_encoder_cfg* cfg, const float* const* params, int num_params, void* packed,
                    size_t packed_bytes, void* stream_) {
  return encoder_pack_impl(cfg, params, num_params, packed, packed_bytes, true, stream_);
}

int pn_encoder_pack_raw(const pn_encoder_cfg* cfg, const float* const* params, int num_params, void* packed,
                        size_t packed_bytes, void* stream_) {
  return encoder_pack_impl(cfg, params, num_params, packed, packed_bytes, false, stream_);
}

static int encoder_pack_impl(const pn_encoder_cfg* cfg, const float* const* params, int num_params, void* packed,
                             size_t packed_bytes, bool fold_bn, void* stream_) {
  PN_TRY(check_encoder_cfg(cfg));
  const pn_encoder_cfg& c = *cfg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_params != 2 + 12 * c.num_blocks) return fail("encoder expects %d parameter pointers, got %d", 2 + 12 * c.num_blocks, num_params);
  const EncoderLayout L = encoder_layout(c);
  if (packed_bytes < L.bytes) return fail("packed buffer too small: %zu < %zu", packed_bytes, L.bytes);
  Arena pk(packed, packed_bytes);
  const int k = c.kernel_size;
  // conv1 (C, Cin, k) + bias; no BatchNorm behind it
  PN_TRY(pack_linear(pk, L.conv1, params[0], (long long)c.input_channels * k, k, 1, (long long)c.input_channels * k,
                     params[1], nullptr, nullptr, nullptr, nullptr, 0.f, stream, kStageEncoder));
  for (int i = 0; i < c.num_blocks; ++i) {
    const float* const* q = params + 2 + 12 * i;
    const EncoderLayout::Block& b = L.blocks[i];
    // bn_activation_1: applied by the epilogue that PRODUCES this block's input
    fold_affine_kernel<<<(c.channels + 255) / 256, 256, 0, stream>>>(c.channels, nullptr, q[0], q[1], q[2], q[3],
                                                                     c.bn_eps, nullptr, pk.at<float>(b.bn1_scale),
                                                                     pk.at<float>(b.bn1_shift));
    g_launches++;
    PN_CUDA(cudaGetLastError());
    // dilated conv (Cb, C, k) + bias, then bn_activation_2 folded into its epilogue
    // (training mode uses batch statistics: fold_bn == false keeps only the conv bias in the epilogue)
    PN_TRY(pack_linear(pk, b.conv_d, q[4], (long long)c.channels * k, k, 1, (long long)c.channels * k, q[5],
                       fold_bn ? q[6] : nullptr, fold_bn ? q[7] : nullptr, fold_bn ? q[8] : nullptr,
                       fold_bn ? q[9] : nullptr, c.bn_eps, stream, kStageEncoder));
    // pointwise conv (C, Cb, 1) + bias
    PN_TRY(pack_linear(pk, b.conv_p, q[10], c.bottleneck, 1, 0, c.bottleneck, q[11], nullptr, nullptr, nullptr, nullptr,
                       0.f, stream, kStagePointwise));
  }
  return 0;
}

size_t pn_encoder_workspace_bytes(const pn_encoder_cfg* cfg, int batch, int T) {
  if (check_encoder_cfg(cfg)) return 0;
  return encoder_ws(*cfg, batch, T).bytes;
}

static int encoder_forward_impl(const pn_encoder_cfg* cfg, const void* packed, const float* x, const uint8_t* tokens,
                                const int64_t* lengths, int batch, int T, float* out, void* workspace,
                                size_t workspace_bytes, int mode, void* stream_);

int pn_encoder_forward(const pn_encoder_cfg* cfg, const void* packed, const float* x, const int64_t* lengths,
                       int batch, int T, float* out, void* workspace, size_t workspace_bytes, int mode,
                       void* stream_) {
  if (!x) return fail("null encoder input");
  return encoder_forward_impl(cfg, packed, x, nullptr, lengths, batch, T, out, workspace, workspace_bytes, mode, stream_);
}

int pn_encoder_forward_tokens(const pn_encoder_cfg* cfg, const void* packed, const uint8_t* tokens,
                              const int64_t* lengths, int batch, int T, float* out, void* workspace,
                              size_t workspace_bytes, int mode, void* stream_) {
  if (!tokens) return fail("null token input");
  return encoder_forward_impl(cfg, packed, nullptr, tokens, lengths, batch, T, out, workspace, workspace_bytes, mode, stream_);
}

static int encoder_forward_impl(const pn_encoder_cfg* cfg, const void* packed, const float* x, const uint8_t* tokens,
                                const int64_t* lengths, int batch, int T, float* out, void* workspace,
                                size_t workspace_bytes, int mode, void* stream_) {
  PN_TRY(check_encoder_cfg(cfg));
  const pn_encoder_cfg& c = *cfg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (batch <= 0 || T <= 0) return fail("empty encoder input (batch %d, T %d)", batch, T);
  const EncoderLayout L = encoder_layout(c);
  Arena pk(const_cast<void*>(packed), L.bytes);
  const size_t per_seq = encoder_ws(c, 1, T).bytes + 4096;
  int sub = (int)(workspace_bytes / per_seq);
  if (sub <= 0) return fail("encoder workspace too small: %zu bytes < %zu for one sequence", workspace_bytes, per_seq);
  if (sub > batch) sub = batch;
  const long long* len64 = reinterpret_cast<const long long*>(lengths);
  for (int b0 = 0; b0 < batch; b0 += sub) {
    const int nb = batch - b0 < sub ? batch - b0 : sub;
    const EncoderWs W = encoder_ws(c, nb, T);
    if (W.bytes > workspace_bytes) return fail("encoder workspace accounting error");
    Arena ws(workspace, workspace_bytes);
    const long long* len = len64 + b0;
    const long long pos = (long long)nb * T;
    if (tokens)
      conv_input_tokens_kernel<<<ew_grid(pos * (W.cin_pad / 8)), 256, 0, stream>>>(
          tokens + (long long)b0 * T, len, nb, c.input_channels, T, W.cin_pad, ws.at<__half>(W.in_hi), ws.at<__half>(W.in_lo));
    else
      conv_input_kernel<<<ew_grid(pos), 256, 0, stream>>>(x + (long long)b0 * c.input_channels * T, len, nb,
                                                          c.input_channels, T, W.cin_pad, ws.at<__half>(W.in_hi),
                                                          ws.at<__half>(W.in_lo));
    g_launches++;
    PN_CUDA(cudaGetLastError());
    float* X = ws.at<float>(W.x);
    ConvView cv;
    cv.batch = nb; cv.T = T; cv.lengths = len;
    // conv1: x0 = mask(conv(x) + bias); act = mask(relu(bn1_0(x0)))
    {
      Planes A;
      A.hi = ws.at<__half>(W.in_hi); A.lo = ws.at<__half>(W.in_lo);
      A.rows = pos; A.cols = c.input_channels; A.ld = W.cin_pad;
      cv.taps = c.kernel_size; cv.dil = 1; cv.cpad = L.conv1.cpad;
      Epilogue e;
      e.scale = pk.at<float>(L.conv1.scale); e.shift = pk.at<float>(L.conv1.shift);
      e.out_f32 = X; e.ld_out = W.ldc;
      if (c.num_blocks > 0) {
        e.scale2 = pk.at<float>(L.blocks[0].bn1_scale); e.shift2 = pk.at<float>(L.blocks[0].bn1_shift);
        e.relu = 1;
        e.out_hi = ws.at<__half>(W.act_hi); e.out_lo = ws.at<__half>(W.act_lo); e.ld_split = W.ldc;
      }
      PN_TRY(launch_gemm(A, cv, weight_planes(pk, L.conv1), c.channels, e, mode, stream, kStageEncoder));
    }
    long long dil = 1;
    for (int i = 0; i < c.num_blocks; ++i) {
      const EncoderLayout::Block& b = L.blocks[i];
      {   // hid = mask(relu(bn2(dilated_conv(act) + bias)))
        Planes A;
        A.hi = ws.at<__half>(W.act_hi); A.lo = ws.at<__half>(W.act_lo);
        A.rows = pos; A.cols = c.channels; A.ld = W.ldc;
        cv.taps = c.kernel_size; cv.dil = (int)dil; cv.cpad = b.conv_d.cpad;
        Epilogue e;
        e.scale = pk.at<float>(b.conv_d.scale); e.shift = pk.at<float>(b.conv_d.shift);
        e.relu = 1;
        e.out_hi = ws.at<__half>(W.hid_hi); e.out_lo = ws.at<__half>(W.hid_lo); e.ld_split = W.ldb;
        PN_TRY(launch_gemm(A, cv, weight_planes(pk, b.conv_d), c.bottleneck, e, mode, stream, kStageEncoder));
      }
      {   // x = x + mask(conv1x1(hid) + bias); act = mask(relu(bn1_{i+1}(x)))
        Planes A;
        A.hi = ws.at<__half>(W.hid_hi); A.lo = ws.at<__half>(W.hid_lo);
        A.rows = pos; A.cols = c.bottleneck; A.ld = W.ldb;
        cv.taps = 1; cv.dil = 1; cv.cpad = b.conv_p.cpad;
        Epilogue e;
        e.scale = pk.at<float>(b.conv_p.scale); e.shift = pk.at<float>(b.conv_p.shift);
        e.resid = X; e.ld_resid = W.ldc;
        e.out_f32 = X; e.ld_out = W.ldc;
        if (i + 1 < c.num_blocks) {
          e.scale2 = pk.at<float>(L.blocks[i + 1].bn1_scale); e.shift2 = pk.at<float>(L.blocks[i + 1].bn1_shift);
          e.relu = 1;
          e.out_hi = ws.at<__half>(W.act_hi); e.out_lo = ws.at<__half>(W.act_lo); e.ld_split = W.ldc;
        }
        PN_TRY(launch_gemm(A, cv, weight_planes(pk, b.conv_p), c.channels, e, mode, stream, kStagePointwise));
      }
      dil *= c.dilation_base;
      if (dil > (1 << 24)) return fail("dilation overflow");
    }
    pool_mean_kernel<<<dim3((c.channels + 31) / 32, nb), dim3(32, 8), 0, stream>>>(X, W.ldc, len, T, c.channels,
                                                                                   out + (long long)b0 * c.channels,
                                                                                   c.channels);
    g_launches++;
    PN_CUDA(cudaGetLastError());
  }
  return 0;
}

// Training-mode encoder forward: every BatchNorm1d uses the statistics of THIS batch over all batch x T positions
// (padding positions are zeros and are counted, exactly as torch.nn.BatchNorm1d sees the [B, C, T] tensor in
// Residual.forward, protein_encoders.py:61-67) and updates its running statistics (momentum 0.01).  This is what the
// reference's frozen encoder does inside ProtNoteTrainer.train (model.train() reaches every submodule,
// ProtNoteTrainer.py:844).  The batch cannot be cut into sub-batches (the statistics couple all sequences).
struct EncoderTrainWs {
  size_t in_hi, in_lo, x, hraw, act_hi, act_lo, hid_hi, hid_lo, stats, state;
  int cin_pad, ldc, ldb;
  size_t bytes;
};

static EncoderTrainWs encoder_train_ws(const pn_encoder_cfg& c, long long batch, long long T) {
  Arena ar(nullptr, 0);
  EncoderTrainWs w;
  w.cin_pad = (int)round_up(c.input_channels, 64);
  w.ldc = (int)round_up(c.channels, 64);
  w.ldb = (int)round_up(c.bottleneck, 64);
  const size_t pos = (size_t)batch * T;
  w.in_hi = ar.take(pos * w.cin_pad * 2);
  w.in_lo = ar.take(pos * w.cin_pad * 2);
  w.x = ar.take(pos * w.ldc * 4);
  w.hraw = ar.take(pos * w.ldb * 4);
  w.act_hi = ar.take(pos * w.ldc * 2);
  w.act_lo = ar.take(pos * w.ldc * 2);
  w.hid_hi = ar.take(pos * w.ldb * 2);
  w.hid_lo = ar.take(pos * w.ldb * 2);
  w.stats = ar.take(sizeof(double) * 2 * (size_t)w.ldc);
  w.state = ar.take(sizeof(float) * 4 * (size_t)w.ldc);
  w.bytes = ar.off;
  return w;
}

size_t pn_encoder_train_workspace_bytes(const pn_encoder_cfg* cfg, int batch, int T) {
  if (check_encoder_cfg(cfg)) return 0;
  return encoder_train_ws(*cfg, batch, T).bytes;
}

int pn_encoder_forward_train(const pn_encoder_cfg* cfg, const void* packed_raw, const float* x, const int64_t* lengths,
                             int batch, int T, const float* const* bn_params, int num_bn_params, float momentum,
                             int update_running, float* out, void* workspace, size_t workspace_bytes, int mode,
                             void* stream_) {
  return pn_encoder_forward_train_sharded(cfg, packed_raw, x, lengths, batch, T, bn_params, num_bn_params, momentum,
                                          update_running, out, workspace, workspace_bytes, mode, (double)batch * T, nullptr,
                                          nullptr, nullptr, stream_);
}

int pn_encoder_forward_train_sharded(const pn_encoder_cfg* cfg, const void* packed_raw, const float* x,
                                     const int64_t* lengths, int batch, int T, const float* const* bn_params,
                                     int num_bn_params, float momentum, int update_running, float* out, void* workspace,
                                     size_t workspace_bytes, int mode, double total_positions, double* stats_buffer,
                                     pn_reduce_fn reduce, void* user, void* stream_) {
  PN_TRY(check_encoder_cfg(cfg));
  if (reduce && !stats_buffer) return fail("sharded training-mode encoder needs a caller-owned stats_buffer");
  if (!(total_positions >= (double)batch * T)) return fail("total_positions (%g) < this rank's positions", total_positions);
  const pn_encoder_cfg& c = *cfg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (batch <= 0 || T <= 0 || !x) return fail("empty encoder input (batch %d, T %d)", batch, T);
  if (num_bn_params != 8 * c.num_blocks) return fail("training-mode encoder expects %d BatchNorm pointers, got %d", 8 * c.num_blocks, num_bn_params);
  const EncoderLayout L = encoder_layout(c);
  Arena pk(const_cast<void*>(packed_raw), L.bytes);
  const EncoderTrainWs W = encoder_train_ws(c, batch, T);
  if (W.bytes > workspace_bytes) return fail("training-mode encoder workspace too small: %zu < %zu", workspace_bytes, W.bytes);
  Arena ws(workspace, workspace_bytes);
  const long long* len = reinterpret_cast<const long long*>(lengths);
  const long long pos = (long long)batch * T;
  const bool strict = mode == PN_STRICT;
  conv_input_kernel<<<ew_grid(pos), 256, 0, stream>>>(x, len, batch, c.input_channels, T, W.cin_pad, ws.at<__half>(W.in_hi),
                                                      ws.at<__half>(W.in_lo));
  g_launches++;
  PN_CUDA(cudaGetLastError());
  float* X = ws.at<float>(W.x);
  float* Hraw = ws.at<float>(W.hraw);
  double* stats = stats_buffer ? stats_buffer : ws.at<double>(W.stats);
  float* state = ws.at<float>(W.state);
  ConvView cv;
  cv.batch = batch; cv.T = T; cv.lengths = len;
  {   // conv1: x0 = mask(conv(x) + bias)
    Planes A;
    A.hi = ws.at<__half>(W.in_hi); A.lo = ws.at<__half>(W.in_lo);
    A.rows = pos; A.cols = c.input_channels; A.ld = W.cin_pad;
    cv.taps = c.kernel_size; cv.dil = 1; cv.cpad = L.conv1.cpad;
    Epilogue e;
    e.scale = pk.at<float>(L.conv1.scale); e.shift = pk.at<float>(L.conv1.shift);
    e.out_f32 = X; e.ld_out = W.ldc;
    PN_TRY(launch_gemm(A, cv, weight_planes(pk, L.conv1), c.channels, e, mode, stream, kStageEncoder));
  }
  // BatchNorm (batch statistics over all batch*T positions) + ReLU + input mask of the next conv, fp32 -> planes
  auto bn_relu = [&](const float* src, int cols, long long ld_src, const float* const* q, __half* hi, __half* lo,
                     long long ld_dst) -> int {
    PN_TRY(pn_t_col_stats(nullptr, nullptr, src, pos, cols, ld_src, stats, stream));
    if (reduce && reduce(stats, 2 * cols, user, stream_) != 0) return fail("the BatchNorm-sum reduction callback failed");
    PN_TRY(pn_t_bn_finalize(stats, total_positions, nullptr, 0.0, q[0], q[1], c.bn_eps, momentum,
                            update_running ? const_cast<float*>(q[2]) : nullptr,
                            update_running ? const_cast<float*>(q[3]) : nullptr, cols, state, stream));
    BnReluMaskF32Producer p{src, pos, cols, ld_src, state, len, T};
    return launch_emit(p, pos, cols, hi, strict ? lo : nullptr, ld_dst, nullptr, nullptr, 0, stream);
  };
  long long dil = 1;
  for (int i = 0; i < c.num_blocks; ++i) {
    const EncoderLayout::Block& b = L.blocks[i];
    const float* const* q = bn_params + 8 * i;
    PN_TRY(bn_relu(X, c.channels, W.ldc, q, ws.at<__half>(W.act_hi), ws.at<__half>(W.act_lo), W.ldc));
    {   // hraw = mask(dilated_conv(act) + bias)
      Planes A;
      A.hi = ws.at<__half>(W.act_hi); A.lo = ws.at<__half>(W.act_lo);
      A.rows = pos; A.cols = c.channels; A.ld = W.ldc;
      cv.taps = c.kernel_size; cv.dil = (int)dil; cv.cpad = b.conv_d.cpad;
      Epilogue e;
      e.scale = pk.at<float>(b.conv_d.scale); e.shift = pk.at<float>(b.conv_d.shift);
      e.out_f32 = Hraw; e.ld_out = W.ldb;
      PN_TRY(launch_gemm(A, cv, weight_planes(pk, b.conv_d), c.bottleneck, e, mode, stream, kStageEncoder));
    }
    PN_TRY(bn_relu(Hraw, c.bottleneck, W.ldb, q + 4, ws.at<__half>(W.hid_hi), ws.at<__half>(W.hid_lo), W.ldb));
    {   // x = x + mask(conv1x1(hid) + bias)
      Planes A;
      A.hi = ws.at<__half>(W.hid_hi); A.lo = ws.at<__half>(W.hid_lo);
      A.rows = pos; A.cols = c.bottleneck; A.ld = W.ldb;
      cv.taps = 1; cv.dil = 1; cv.cpad = b.conv_p.cpad;
      Epilogue e;
      e.scale = pk.at<float>(b.conv_p.scale); e.shift = pk.at<float>(b.conv_p.shift);
      e.resid = X; e.ld_resid = W.ldc;
      e.out_f32 = X; e.ld_out = W.ldc;
      PN_TRY(launch_gemm(A, cv, weight_planes(pk, b.conv_p), c.channels, e, mode, stream, kStagePointwise));
    }
    dil *= c.dilation_base;
    if (dil > (1 << 24)) return fail("dilation overflow");
  }
  pool_mean_kernel<<<dim3((c.channels + 31) / 32, batch), dim3(32, 8), 0, stream>>>(X, W.ldc, len, T, c.channels, out,
                                                                                    c.channels);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------- scorer
int pn_scorer_num_params(const pn_scorer_cfg* cfg) {
  if (check_scorer_cfg(cfg)) return -1;
  const int head = 5 * (cfg->proj_layers - 1) + 1;
  if (cfg->fusion == PN_FUSION_SIMILARITY) return 2 * head;
  const int out = cfg->out_layers * (cfg->out_batchnorm ? 5 : 2) + 2;
  return 2 * head + out;
}

size_t pn_scorer_packed_bytes(const pn_scorer_cfg* cfg) {
  if (check_scorer_cfg(cfg)) return 0;
  return scorer_layout(*cfg).bytes;
}

int pn_scorer_pack(const pn_scorer_cfg* cfg, const float* const* params, int num_params, void* packed,
                   size_t packed_bytes, void* stream_) {
  PN_TRY(check_scorer_cfg(cfg));
  const pn_scorer_cfg& c = *cfg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_params != pn_scorer_num_params(cfg)) return fail("scorer expects %d parameter pointers, got %d", pn_scorer_num_params(cfg), num_params);
  const ScorerLayout L = scorer_layout(c);
  if (packed_bytes < L.bytes) return fail("packed buffer too small: %zu < %zu", packed_bytes, L.bytes);
  Arena pk(packed, packed_bytes);
  const float* const* q = params;
  for (int head = 0; head < 2; ++head) {
    const std::vector<PackedLinear>& H = head == 0 ? L.wp : L.wl;
    for (size_t i = 0; i < H.size(); ++i) {
      const bool last = i + 1 == H.size();
      const float* w = *q++;
      const float *g = nullptr, *bt = nullptr, *mu = nullptr, *var = nullptr;
      if (!last) {
        g = *q++; bt = *q++; mu = *q++; var = *q++;
      }
      PN_TRY(pack_linear(pk, H[i], w, H[i].cin, 1, 0, H[i].cin, nullptr, g, bt, mu, var, c.bn_eps, stream, kStageHeads));
      if (head == 0) {   // fp64 protein-side head: the weights as they are + the BatchNorm fold in fp64
        PN_CUDA(cudaMemcpyAsync(pk.at<float>(L.wp_raw[i]), w, (size_t)H[i].N * H[i].cin * 4, cudaMemcpyDeviceToDevice, stream));
        fold_affine_f64_kernel<<<(H[i].N + 255) / 256, 256, 0, stream>>>(H[i].N, nullptr, g, bt, mu, var, (double)c.bn_eps,
                                                                         true, pk.at<double>(L.wp_scale64[i]),
                                                                         pk.at<double>(L.wp_shift64[i]));
        g_launches++;
        PN_CUDA(cudaGetLastError());
      }
    }
  }
  if (c.fusion == PN_FUSION_SIMILARITY) return 0;
  // output layer 1: weight (H, 2d or 3d), split by column block
  const int d = c.latent_dim;
  const int in1 = c.fusion == PN_FUSION_CONCAT ? 2 * d : 3 * d;
  {
    const float* w = *q++;
    const float *bias = nullptr, *g = nullptr, *bt = nullptr, *mu = nullptr, *var = nullptr;
    if (c.out_batchnorm) {
      g = *q++; bt = *q++; mu = *q++; var = *q++;
    } else {
      bias = *q++;
    }
    const float* wp_src = w;
    const float* wl_src = w + d;
    long long src_ld = in1;
    if (c.fusion == PN_FUSION_CONCAT_DIFF) {
      // W [p; t; p - t] = (Wp + Wd) p + (Wl - Wd) t   (exact in real arithmetic)
      float* tp = pk.at<float>(L.tmp);
      float* tl = tp + (size_t)c.out_hidden * d;
      combine_kernel<<<ew_grid((long long)c.out_hidden * d), 256, 0, stream>>>(w, w + 2 * d, 1.f, c.out_hidden, d, in1,
                                                                               in1, tp);
      combine_kernel<<<ew_grid((long long)c.out_hidden * d), 256, 0, stream>>>(w + d, w + 2 * d, -1.f, c.out_hidden, d,
                                                                               in1, in1, tl);
      g_launches += 2;
      PN_CUDA(cudaGetLastError());
      wp_src = tp;
      wl_src = tl;
      src_ld = d;
    }
    // fp64 protein-side head: W1p as it is (row pitch d) + BN1 folded in fp64
    PN_CUDA(cudaMemcpy2DAsync(pk.at<float>(L.l1p_raw), (size_t)d * 4, wp_src, (size_t)src_ld * 4, (size_t)d * 4, c.out_hidden,
                              cudaMemcpyDeviceToDevice, stream));
    fold_affine_f64_kernel<<<(c.out_hidden + 255) / 256, 256, 0, stream>>>(c.out_hidden, bias, g, bt, mu, var, (double)c.bn_eps,
                                                                           true, pk.at<double>(L.l1p_scale64),
                                                                           pk.at<double>(L.l1p_shift64));
    g_launches++;
    PN_CUDA(cudaGetLastError());
    // scale = BN1 scale / wscale on every part; the shift (bias, mean, beta) goes to the protein side only
    PN_TRY(pack_linear(pk, L.l1_p, wp_src, src_ld, 1, 0, d, bias, g, bt, mu, var, c.bn_eps, stream, kStageHeads));
    PN_TRY(pack_linear(pk, L.l1_l, wl_src, src_ld, 1, 0, d, nullptr, g, nullptr, nullptr, var, c.bn_eps, stream, kStageHeads));
    if (c.fusion == PN_FUSION_CONCAT_PROD)
      PN_TRY(pack_linear(pk, L.l1_x, w + 2 * d, src_ld, 1, 0, d, nullptr, g, nullptr, nullptr, var, c.bn_eps, stream,
                         kStageScorer));
    copy_floats_kernel<<<(c.out_hidden + 255) / 256, 256, 0, stream>>>(pk.at<float>(L.l1_p.shift), pk.at<float>(L.l1_shift),
                                                                       c.out_hidden);
    g_launches++;
    PN_CUDA(cudaGetLastError());
  }
  for (size_t j = 0; j < L.hidden.size(); ++j) {
    const float* w = *q++;
    const float *bias = nullptr, *g = nullptr, *bt = nullptr, *mu = nullptr, *var = nullptr;
    if (c.out_batchnorm) {
      g = *q++; bt = *q++; mu = *q++; var = *q++;
    } else {
      bias = *q++;
    }
    PN_TRY(pack_linear(pk, L.hidden[j], w, c.out_hidden, 1, 0, c.out_hidden, bias, g, bt, mu, var, c.bn_eps, stream,
                       kStageScorer));
  }
  copy_floats_kernel<<<(c.out_hidden + 255) / 256, 256, 0, stream>>>(*q++, pk.at<float>(L.w_out), c.out_hidden);
  copy_floats_kernel<<<1, 32, 0, stream>>>(*q++, pk.at<float>(L.b_out), 1);
  g_launches += 2;
  PN_CUDA(cudaGetLastError());
  return 0;
}

size_t pn_project_workspace_bytes(const pn_scorer_cfg* cfg, long long rows) {
  if (check_scorer_cfg(cfg)) return 0;
  const long long in_dim = cfg->protein_dim > cfg->label_dim ? cfg->protein_dim : cfg->label_dim;
  const long long ld_in = round_up(in_dim, 64);
  const long long ld_h = round_up(cfg->proj_hidden > cfg->latent_dim ? cfg->proj_hidden : cfg->latent_dim, 64);
  const long long r = round_up(rows < 1 ? 1 : rows, kBM);
  // the larger of the tensor-core path (fp16 planes) and the fp64 protein-side path (three fp64 row buffers)
  const long long planes = ld_in * 4 + ld_h * 4 * 2 + 1024, f64 = (ld_in + 2 * ld_h) * 8;
  return (size_t)(r * (planes > f64 ? planes : f64) + 8192);
}

int pn_project_sequences(const pn_scorer_cfg* cfg, const void* packed, const float* P_f, long long n, float* P_e,
                         float* a, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  PN_TRY(check_scorer_cfg(cfg));
  if (n <= 0) return fail("no sequences to project");
  if (a == nullptr && cfg->fusion != PN_FUSION_SIMILARITY) return fail("the layer-1 half output is required for this fusion");
  const ScorerLayout L = scorer_layout(*cfg);
  Arena pk(const_cast<void*>(packed), L.bytes);
  return run_projection(*cfg, L, pk, true, P_f, n, P_e, a, workspace, workspace_bytes, mode,
                        static_cast<cudaStream_t>(stream));
}

int pn_project_labels(const pn_scorer_cfg* cfg, const void* packed, const float* L_f, long long n, float* L_e,
                      float* c_out, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  PN_TRY(check_scorer_cfg(cfg));
  if (n <= 0) return fail("no label rows to project");
  if (c_out == nullptr && cfg->fusion != PN_FUSION_SIMILARITY) return fail("the layer-1 half output is required for this fusion");
  const ScorerLayout L = scorer_layout(*cfg);
  Arena pk(const_cast<void*>(packed), L.bytes);
  return run_projection(*cfg, L, pk, false, L_f, n, L_e, c_out, workspace, workspace_bytes, mode,
                        static_cast<cudaStream_t>(stream));
}

size_t pn_scorer_min_workspace_bytes(const pn_scorer_cfg* cfg) {
  if (check_scorer_cfg(cfg)) return 0;
  const long long rows = round_up(1024, (long long)kBM * cfg->descriptions_per_label);
  return scorer_row_bytes(*cfg) * (size_t)rows + 8192;
}

size_t pn_scorer_workspace_bytes(const pn_scorer_cfg* cfg, long long B, long long L) {
  if (check_scorer_cfg(cfg)) return 0;
  long long rows = B * L;
  const long long cap = 1LL << 19;   // 512 Ki pairs per chunk is already > 20 ms of tensor work
  if (rows > cap) rows = L <= cap ? cap / L * L : cap;
  const size_t need = scorer_row_bytes(*cfg) * (size_t)round_up(rows, kBM) + 8192;
  const size_t mn = pn_scorer_min_workspace_bytes(cfg);
  return need > mn ? need : mn;
}

int pn_score_pairs(const pn_scorer_cfg* cfg, const void* packed, const float* a, const float* c_in, const float* P_e,
                   const float* L_e, long long B, long long Lrows, float* logits, long long ld_logits, void* workspace,
                   size_t workspace_bytes, int mode, void* stream_) {
  return pn_score_pairs_ex(cfg, packed, a, c_in, P_e, L_e, B, Lrows, logits, ld_logits, nullptr, workspace, workspace_bytes,
                           mode, stream_);
}

int pn_score_pairs_ex(const pn_scorer_cfg* cfg, const void* packed, const float* a, const float* c_in, const float* P_e,
                      const float* L_e, long long B, long long Lrows, float* logits, long long ld_logits,
                      float* hidden_out, void* workspace, size_t workspace_bytes, int mode, void* stream_) {
  PN_TRY(check_scorer_cfg(cfg));
  const pn_scorer_cfg& c = *cfg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int k = c.descriptions_per_label;
  if (c.fusion == PN_FUSION_SIMILARITY) return fail("use pn_score_similarity for PN_FUSION_SIMILARITY");
  if (B <= 0 || Lrows <= 0) return fail("empty scorer input (B %lld, L %lld)", B, Lrows);
  if (Lrows % k != 0) return fail("label rows (%lld) not a multiple of descriptions_per_label (%d)", Lrows, k);
  if (c.fusion == PN_FUSION_CONCAT_PROD && (P_e == nullptr || L_e == nullptr))
    return fail("concatenation_prod needs P_e and L_e");
  const ScorerLayout L = scorer_layout(c);
  Arena pk(const_cast<void*>(packed), L.bytes);
  const int H = c.out_hidden;
  const int ld_h = (int)round_up(H, 64);
  // the fused generator reads a / c with 16-byte loads over whole 32-wide k-blocks
  const bool fuse_ok = g_fuse_features && !L.hidden.empty() && c.fusion != PN_FUSION_CONCAT_PROD && H % 32 == 0 &&
                       (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(c_in) & 15) == 0;
  const int parts = 4 * tiles_n_for(H);   // one partial dot (two floats: hi, lo of an fp64 sum) per (N tile, column half)
  const size_t per_row = scorer_row_bytes(c);
  long long max_rows = (long long)((workspace_bytes > 8192 ? workspace_bytes - 8192 : 0) / per_row);
  if (g_chunk_rows > 0 && g_chunk_rows < max_rows) max_rows = g_chunk_rows;
  max_rows = max_rows / ((long long)kBM * k) * ((long long)kBM * k);
  if (max_rows <= 0) return fail("scorer workspace too small: %zu bytes", workspace_bytes);
  // chunk = nb whole proteins x all label rows when that fits, else one protein x a label range
  long long nb_chunk = max_rows / Lrows, nl_chunk = Lrows;
  if (nb_chunk == 0) {
    nb_chunk = 1;
    nl_chunk = max_rows;
  }
  for (long long b0 = 0; b0 < B; b0 += nb_chunk) {
    const long long nb = B - b0 < nb_chunk ? B - b0 : nb_chunk;
    for (long long l0 = 0; l0 < Lrows; l0 += nl_chunk) {
      const long long nl = Lrows - l0 < nl_chunk ? Lrows - l0 : nl_chunk;
      const long long rows = nb * nl;
      Arena ws(workspace, workspace_bytes);
      __half* buf_hi[2];
      __half* buf_lo[2];
      for (int q = 0; q < 2; ++q) {
        buf_hi[q] = ws.at<__half>(ws.take((size_t)max_rows * ld_h * 2));
        buf_lo[q] = ws.at<__half>(ws.take((size_t)max_rows * ld_h * 2));
      }
      float* partial = ws.at<float>(ws.take((size_t)max_rows * parts * 4));
      if (!ws.ok()) return fail("scorer workspace accounting error (%zu > %zu)", ws.off, workspace_bytes);
      int cur = 0;
      if (c.fusion == PN_FUSION_CONCAT_PROD) {
        // x = p * t into buffer 1, then h1 = relu(x W1x^T * s + a[b] + c[l]) into buffer 0
        const int ld_d = (int)round_up(c.latent_dim, 64);
        pair_product_kernel<<<ew_grid(rows * (ld_d / 8)), 256, 0, stream>>>(P_e, c.latent_dim, L_e, c.latent_dim, (int)b0,
                                                                            (int)l0, (int)nl, rows, c.latent_dim,
                                                                            buf_hi[1], buf_lo[1], ld_d);
        g_launches++;
        PN_CUDA(cudaGetLastError());
        Planes A;
        A.hi = buf_hi[1]; A.lo = buf_lo[1]; A.rows = rows; A.cols = c.latent_dim; A.ld = ld_d;
        Epilogue e;
        e.scale = pk.at<float>(L.l1_x.scale);
        e.pair_nl = (int)nl;
        e.add_p = a + b0 * H; e.ld_add_p = H;
        e.add_l = c_in + l0 * H; e.ld_add_l = H;
        e.relu = 1;
        e.out_hi = buf_hi[0]; e.out_lo = buf_lo[0]; e.ld_split = ld_h;
        PN_TRY(launch_gemm(A, ConvView(), weight_planes(pk, L.l1_x), H, e, mode, stream, kStageScorer));
      }
      const bool fuse_features = fuse_ok && nl % kBM == 0;   // a tile = one protein x 128 consecutive label rows
      if (c.fusion != PN_FUSION_CONCAT_PROD && !fuse_features) {
        const int pf_threads = (int)(round_up(ld_h / 8, 32) < 1024 ? round_up(ld_h / 8, 32) : 1024);
        pair_features_kernel<<<(unsigned)((rows + kPairRows - 1) / kPairRows), pf_threads, 0, stream>>>(a, H, c_in, H, (int)b0, (int)l0, (int)nl, rows,
                                                                             H, buf_hi[0], mode == PN_STRICT ? buf_lo[0] : nullptr,
                                                                             ld_h);
        g_launches++;
        PN_CUDA(cudaGetLastError());
      }
      for (size_t j = 0; j < L.hidden.size(); ++j) {
        const PackedLinear& pl = L.hidden[j];
        const bool last = j + 1 == L.hidden.size();
        Planes A;
        A.hi = buf_hi[cur]; A.lo = buf_lo[cur]; A.rows = rows; A.cols = H; A.ld = ld_h;
        Epilogue e;
        e.scale = pk.at<float>(pl.scale); e.shift = pk.at<float>(pl.shift);
        e.relu = 1;
        if (j == 0 && fuse_features) {
          // layer 1 is built inside this GEMM's A-producer: h1 never touches HBM
          e.pair_nl = (int)nl;
          e.gen_a = a + b0 * H; e.ld_gen_a = H;
          e.gen_c = c_in + l0 * H; e.ld_gen_c = H;
        }
        if (last) {
          e.dot_w = pk.at<float>(L.w_out);
          e.dot_out = partial;
          if (hidden_out) {   // output_layer_embeddings of save_embeddings=True: rows in pair order (b*L + l)
            if (nl != Lrows) return fail("hidden_out needs a workspace that holds whole proteins (L rows) per chunk");
            e.out_z = hidden_out + b0 * Lrows * H;
            e.ld_z = H;
          }
        } else {
          e.out_hi = buf_hi[cur ^ 1]; e.out_lo = buf_lo[cur ^ 1]; e.ld_split = ld_h;
        }
        PN_TRY(launch_gemm(A, ConvView(), weight_planes(pk, pl), H, e, mode, stream, kStageScorer));
        cur ^= 1;
      }
      int nparts = parts;
      if (L.hidden.empty()) {
        // OUTPUT_MLP_NUM_LAYERS 1: the output neuron follows layer 1.  h1 = relu(...) is in buffer 0 as planes; one warp
        // per pair takes its dot product with w_out (the training step's last-layer kernel with the identity state)
        if (hidden_out) return fail("hidden_out is the activation in front of the last hidden layer: a one-layer output MLP has none");
        bn_relu_dot_kernel<<<ew_grid(rows * 32), 256, 0, stream>>>(buf_hi[0], mode == PN_STRICT ? buf_lo[0] : nullptr, rows, H,
                                                                   ld_h, nullptr, pk.at<float>(L.w_out), nullptr, partial);
        g_launches++;
        PN_CUDA(cudaGetLastError());
        nparts = 1;
      }
      finalize_logits_kernel<<<ew_grid(rows / k), 256, 0, stream>>>(partial, nparts, pk.at<float>(L.b_out), (int)b0, (int)l0,
                                                                    (int)nl, rows, k, logits, ld_logits);
      g_launches++;
      PN_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

size_t pn_similarity_workspace_bytes(const pn_scorer_cfg* cfg, long long B, long long L) {
  if (check_scorer_cfg(cfg)) return 0;
  const long long ld = round_up(cfg->latent_dim, 64);
  return (size_t)((round_up(B, kBM) + L) * ld * 4 + L * 4 + (cfg->descriptions_per_label > 1 ? B * L * 4 : 0) + 16384);
}

int pn_score_similarity(const pn_scorer_cfg* cfg, const float* P_e, const float* L_e, long long B, long long Lrows,
                        float temperature, float* logits, long long ld_logits, void* workspace, size_t workspace_bytes,
                        int mode, void* stream_) {
  PN_TRY(check_scorer_cfg(cfg));
  const pn_scorer_cfg& c = *cfg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int k = c.descriptions_per_label;
  if (B <= 0 || Lrows <= 0) return fail("empty scorer input (B %lld, L %lld)", B, Lrows);
  if (Lrows % k != 0) return fail("label rows (%lld) not a multiple of descriptions_per_label (%d)", Lrows, k);
  if (!(temperature > 0.f)) return fail("temperature must be positive");
  if (workspace_bytes < pn_similarity_workspace_bytes(cfg, B, Lrows)) return fail("similarity workspace too small");
  const int d = c.latent_dim;
  const int ld = (int)round_up(d, 64);
  Arena ws(workspace, workspace_bytes);
  __half* p_hi = ws.at<__half>(ws.take((size_t)B * ld * 2));
  __half* p_lo = ws.at<__half>(ws.take((size_t)B * ld * 2));
  __half* l_hi = ws.at<__half>(ws.take((size_t)Lrows * ld * 2));
  __half* l_lo = ws.at<__half>(ws.take((size_t)Lrows * ld * 2));
  float* scale = ws.at<float>(ws.take((size_t)Lrows * 4));
  float* raw = k > 1 ? ws.at<float>(ws.take((size_t)B * Lrows * 4)) : logits;
  if (!ws.ok()) return fail("similarity workspace accounting error");
  const float pow2 = 128.f;   // keeps the lo planes of unit vectors in fp16's normal range; undone in the epilogue
  normalize_split_kernel<<<(int)((B * 32 + 255) / 256), 256, 0, stream>>>(P_e, B, d, pow2, p_hi, p_lo, ld);
  normalize_split_kernel<<<(int)((Lrows * 32 + 255) / 256), 256, 0, stream>>>(L_e, Lrows, d, pow2, l_hi, l_lo, ld);
  fill_kernel<<<ew_grid(Lrows), 256, 0, stream>>>(scale, Lrows, 1.f / (pow2 * pow2) / temperature);
  g_launches += 3;
  PN_CUDA(cudaGetLastError());
  Planes A, Bm;
  A.hi = p_hi; A.lo = p_lo; A.rows = B; A.cols = d; A.ld = ld;
  Bm.hi = l_hi; Bm.lo = l_lo; Bm.rows = Lrows; Bm.cols = d; Bm.ld = ld;
  Epilogue e;
  e.scale = scale;
  e.out_f32 = raw;
  e.ld_out = k > 1 ? Lrows : ld_logits;
  PN_TRY(launch_gemm(A, ConvView(), Bm, Lrows, e, mode, stream, kStageHeads));
  if (k > 1) {
    finalize_logits_kernel<<<ew_grid(B * Lrows / k), 256, 0, stream>>>(raw, 1, nullptr, 0, 0, (int)Lrows, B * Lrows, k, logits,
                                                                     ld_logits);
    g_launches++;
    PN_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------------------- plain layers
size_t pn_linear_workspace_bytes(long long M, long long N, long long K) {
  const long long ld = round_up(K, 64);
  return (size_t)(round_up(M, kBM) * ld * 4 + N * ld * 4 + N * 8 + 8192);
}

int pn_linear(const float* x, long long M, long long K, long long ldx, const float* w, long long N, const float* bias,
              float* y, long long ldy, void* workspace, size_t workspace_bytes, int mode, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0) return fail("empty linear layer");
  if (workspace_bytes < pn_linear_workspace_bytes(M, N, K)) return fail("linear workspace too small");
  Arena ws(workspace, workspace_bytes);
  const int ld = (int)round_up(K, 64);
  __half* a_hi = ws.at<__half>(ws.take((size_t)M * ld * 2));
  __half* a_lo = ws.at<__half>(ws.take((size_t)M * ld * 2));
  PackedLinear pl = carve_linear(ws, (int)N, (int)K, 1);
  if (!ws.ok()) return fail("linear workspace accounting error");
  split_rows_kernel<<<ew_grid(M * (ld / 8)), 256, 0, stream>>>(x, M, (int)K, ldx, a_hi, a_lo, ld);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  PN_TRY(pack_linear(ws, pl, w, K, 1, 0, K, bias, nullptr, nullptr, nullptr, nullptr, 0.f, stream, kStageOther));
  Planes A;
  A.hi = a_hi; A.lo = a_lo; A.rows = M; A.cols = K; A.ld = ld;
  Epilogue e;
  e.scale = ws.at<float>(pl.scale); e.shift = ws.at<float>(pl.shift);
  e.out_f32 = y; e.ld_out = ldy;
  return launch_gemm(A, ConvView(), weight_planes(ws, pl), N, e, mode, stream);
}

size_t pn_conv1d_workspace_bytes(int batch, int T, int cin, int cout, int taps) {
  const long long cpad = round_up(cin, 64);
  return (size_t)((long long)batch * T * cpad * 4 + (long long)cout * cpad * taps * 4 + (long long)cout * 8 + 16384);
}

int pn_conv1d(const float* x, const int64_t* lengths, int batch, int cin, int T, const float* w, const float* bias,
              int cout, int taps, int dilation, float* y, void* workspace, size_t workspace_bytes, int mode,
              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (batch <= 0 || T <= 0 || cin <= 0 || cout <= 0) return fail("empty conv");
  if (taps < 1 || taps % 2 == 0) return fail("taps must be odd");
  if (workspace_bytes < pn_conv1d_workspace_bytes(batch, T, cin, cout, taps)) return fail("conv workspace too small");
  Arena ws(workspace, workspace_bytes);
  const int cpad = (int)round_up(cin, 64);
  const long long pos = (long long)batch * T;
  __half* a_hi = ws.at<__half>(ws.take((size_t)pos * cpad * 2));
  __half* a_lo = ws.at<__half>(ws.take((size_t)pos * cpad * 2));
  PackedLinear pl = carve_linear(ws, cout, cin, taps);
  if (!ws.ok()) return fail("conv workspace accounting error");
  const long long* len = reinterpret_cast<const long long*>(lengths);
  conv_input_kernel<<<ew_grid(pos), 256, 0, stream>>>(x, len, batch, cin, T, cpad, a_hi, a_lo);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  PN_TRY(pack_linear(ws, pl, w, (long long)cin * taps, taps, 1, (long long)cin * taps, bias, nullptr, nullptr, nullptr,
                     nullptr, 0.f, stream, kStageOther));
  Planes A;
  A.hi = a_hi; A.lo = a_lo; A.rows = pos; A.cols = cin; A.ld = cpad;
  ConvView cv;
  cv.taps = taps; cv.dil = dilation; cv.T = T; cv.batch = batch; cv.cpad = pl.cpad; cv.lengths = len;
  Epilogue e;
  e.scale = ws.at<float>(pl.scale); e.shift = ws.at<float>(pl.shift);
  e.out_f32 = y; e.ld_out = cout;
  return launch_gemm(A, cv, weight_planes(ws, pl), cout, e, mode, stream);
}

// ---------------------------------------------------------------------------------------- evaluation post-processing
int pn_postprocess(const float* logits, long long B, long long L, long long ld_logits, const void* labels, int label_kind,
                   long long ld_labels, float threshold, float* probs, long long ld_probs, float* tp, float* fn, float* fp,
                   int topk, float* topk_values, int* topk_indices, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B <= 0 || L <= 0 || !logits) return fail("empty logits");
  if (label_kind < 0 || label_kind > 2) return fail("label_kind must be 0 (none), 1 (int64) or 2 (float32)");
  if (label_kind != 0 && (!labels || !tp || !fn || !fp)) return fail("label counts need labels, tp, fn and fp");
  if (topk < 0 || topk > 64 || topk > L) return fail("topk must be in [0, min(64, L)]");
  if (topk > 0 && (!topk_values || !topk_indices)) return fail("topk needs output buffers");
  if (label_kind != 0 || probs) {
    const int col_blocks = (int)((L + 255) / 256);
    long long slabs = ((long long)num_sms() * 8 + col_blocks - 1) / col_blocks;
    long long per = (B + slabs - 1) / slabs;
    if (per < 8) per = 8;
    slabs = (B + per - 1) / per;
    postprocess_counts_kernel<<<dim3(col_blocks, (unsigned)slabs), 256, 0, stream>>>(
        logits, B, L, ld_logits, labels, label_kind, ld_labels, threshold, probs, ld_probs, per, tp, fn, fp);
    g_launches++;
    PN_CUDA(cudaGetLastError());
  }
  if (topk > 0) {
    if (B > 2147483647LL) return fail("too many rows");
    topk_rows_kernel<<<(unsigned)B, 256, 0, stream>>>(logits, L, ld_logits, topk, topk_values, topk_indices);
    g_launches++;
    PN_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------------------- training primitives
int pn_t_autoscale(const float* x, long long rows, long long cols, long long ldx, float* sc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || cols <= 0) return fail("empty tensor");
  unsigned* am = reinterpret_cast<unsigned*>(sc) + 1;
  PN_CUDA(cudaMemsetAsync(am, 0, 4, stream));
  absmax_kernel<<<ew_grid(rows * cols), 256, 0, stream>>>(x, rows, cols, ldx, am);
  autoscale_finish_kernel<<<1, 1, 0, stream>>>(sc);
  g_launches += 2;
  PN_CUDA(cudaGetLastError());
  return 0;
}


int pn_t_split(const float* x, long long rows, long long cols, long long ldx, const float* sc, void* hi, void* lo,
               long long ld, void* hiT, void* loT, long long blocksT, void* stream) {
  SplitProducer p{x, rows, (int)cols, ldx, sc};
  return launch_emit(p, rows, (int)cols, hi, lo, ld, hiT, loT, blocksT, static_cast<cudaStream_t>(stream));
}

int pn_t_pack_weight(const float* w, long long N, long long K, long long stride_n, long long stride_k, void* hi, void* lo,
                     long long ld, float* ws, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N <= 0 || K <= 0 || ld < K || ld % 64 != 0) return fail("bad weight shape (N %lld, K %lld, ld %lld)", N, K, ld);
  unsigned* am = reinterpret_cast<unsigned*>(ws) + 1;
  PN_CUDA(cudaMemsetAsync(am, 0, 4, stream));
  // max |w| over the N x K view: walk it as N*K single-element "rows" when it is not row-contiguous
  if (stride_k == 1) {
    absmax_kernel<<<ew_grid(N * K), 256, 0, stream>>>(w, N, K, stride_n, am);
  } else if (stride_n == 1) {
    absmax_kernel<<<ew_grid(N * K), 256, 0, stream>>>(w, K, N, stride_k, am);
  } else {
    return fail("weight view must be contiguous along one axis");
  }
  g_launches++;
  PN_CUDA(cudaGetLastError());
  pack_weight_kernel<<<ew_grid(N * ld), 256, 0, stream>>>(w, (int)N, (int)K, 1, stride_n, stride_k, 0, (int)ld, (int)ld, am,
                                                          ws, static_cast<__half*>(hi), static_cast<__half*>(lo),
                                                          TruncComp{16, 1, 1, 1, 0.f});
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_gemm(const void* a_hi, const void* a_lo, long long M, long long K, long long lda, const void* b_hi,
              const void* b_lo, long long N, long long ldb, const float* s0, const float* s1, const float* s2,
              float* scale_scratch, float* out_f32, long long ld_out, int accumulate, void* out_hi, void* out_lo,
              long long ld_split, int mode, int promote_k, long long split_k, int k_blocked, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0) return fail("empty GEMM");
  if (split_k < 0 || split_k % 64 != 0) return fail("split_k must be a non-negative multiple of 64");
  if (split_k > 0 && (!out_f32 || out_hi)) return fail("split_k needs an fp32 output only");
  if (!a_hi || !b_hi || !scale_scratch) return fail("GEMM operand missing");
  if (!k_blocked && (lda % 8 != 0 || ldb % 8 != 0)) return fail("operand pitches must be multiples of 8 elements");
  if (!out_f32 && !out_hi) return fail("GEMM has no output");
  scale_vector_kernel<<<(int)((N + 255) / 256), 256, 0, stream>>>(scale_scratch, (int)N, s0, s1, s2);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  Planes A, B;
  A.hi = static_cast<const __half*>(a_hi); A.lo = static_cast<const __half*>(a_lo); A.rows = M; A.cols = K; A.ld = lda;
  B.hi = static_cast<const __half*>(b_hi); B.lo = static_cast<const __half*>(b_lo); B.rows = N; B.cols = K; B.ld = ldb;
  A.kblocked = B.kblocked = k_blocked != 0;
  Epilogue e;
  e.scale = scale_scratch;
  if (out_f32) {
    e.out_f32 = out_f32; e.ld_out = ld_out;
    if (accumulate) { e.resid = out_f32; e.ld_resid = ld_out; }
  }
  if (out_hi) {
    e.out_hi = static_cast<__half*>(out_hi); e.out_lo = static_cast<__half*>(out_lo); e.ld_split = ld_split;
  }
  // (the training GEMMs measured 4 % slower as CTA pairs in fast mode: single-CTA kernels unless cta2 == 1 is forced)
  if (split_k == 0 || K <= split_k)
    return launch_gemm(A, ConvView(), B, N, e, mode, stream, kStageHeads, promote_k > 0 ? promote_k : -1, false);
  // K in slices, one launch each, accumulating in the fp32 output: CTAs of one launch stay within `split_k` of each other
  // along K, so the operand panels they share are still in L2 when the next CTA asks for them (a single launch over
  // K = millions of rows lets the CTAs drift apart and re-reads the panels from HBM ~8x, measured).
  for (long long k0 = 0; k0 < K; k0 += split_k) {
    Planes As = A, Bs = B;
    const long long kk = K - k0 < split_k ? K - k0 : split_k;
    // K-blocked operands [K/64][rows][64]: slice k0 starts at block k0/64
    const long long oa = k_blocked ? (k0 / 64) * M * 64 : k0, ob = k_blocked ? (k0 / 64) * N * 64 : k0;
    As.hi += oa; Bs.hi += ob;
    if (As.lo) As.lo += oa;
    if (Bs.lo) Bs.lo += ob;
    As.cols = kk; Bs.cols = kk;
    if (k0 > 0) { e.resid = out_f32; e.ld_resid = ld_out; }
    PN_TRY(launch_gemm(As, ConvView(), Bs, N, e, mode, stream, kStageHeads, promote_k > 0 ? promote_k : -1, false));
  }
  return 0;
}

int pn_t_col_stats(const void* hi, const void* lo, const float* x, long long rows, int cols, long long ld, double* out,
                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || cols <= 0) return fail("empty tensor");
  if (!x && (!hi || ld % 8 != 0)) return fail("column statistics: planes missing or pitch not a multiple of 8");
  PN_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 2 * cols, stream));
  const int col_blocks = (cols + g_stats_tx * 8 - 1) / (g_stats_tx * 8);
  const long long per = slab_rows(rows, col_blocks);
  const long long slabs = (rows + per - 1) / per;
  col_stats_kernel<<<dim3(col_blocks, (unsigned)slabs), dim3(g_stats_tx, 256 / g_stats_tx), 0, stream>>>(
      static_cast<const __half*>(hi), static_cast<const __half*>(lo), x, rows, cols, ld, per, out);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_bn_finalize(const double* stats, double count, const double* stats2, double count2, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var, int cols,
                     float* state, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (cols <= 0 || !(count > 0) || (stats2 && !(count2 > 0))) return fail("bad BatchNorm statistics");
  bn_finalize_kernel<<<(cols + 255) / 256, 256, 0, stream>>>(stats, count, stats2, count2, gamma, beta, eps, momentum,
                                                             running_mean, running_var, cols, state);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_bn_relu(const void* z_hi, const void* z_lo, long long rows, int cols, long long ld_z, const float* state,
                 void* h_hi, void* h_lo, long long ld_h, void* hT_hi, void* hT_lo, long long blocksT, void* stream) {
  if (!z_hi || ld_z % 8 != 0) return fail("pre-activation planes missing or pitch not a multiple of 8");
  BnReluProducer p{static_cast<const __half*>(z_hi), static_cast<const __half*>(z_lo), rows, cols, ld_z, state};
  return launch_emit(p, rows, cols, h_hi, h_lo, ld_h, hT_hi, hT_lo, blocksT, static_cast<cudaStream_t>(stream));
}

namespace {
// thr = round(p * 65536) in [0, 65535], kept values scaled by 65536 / (65536 - thr)
bool make_drop(unsigned long long seed, float p, int cols, Drop& d) {
  if (!(p >= 0.f) || !(p < 1.f)) return false;
  long t = lrintf(p * 65536.f);
  if (t > 65535) t = 65535;
  d.seed = seed;
  d.thr = (unsigned)t;
  d.scale = 65536.f / (float)(65536 - t);
  d.groups = (cols + 7) / 8;
  return true;
}
}  // namespace

int pn_t_dropout_planes(const void* x_hi, const void* x_lo, long long rows, int cols, long long ld_x,
                        unsigned long long seed, float p, void* hi, void* lo, long long ld, void* hiT, void* loT,
                        long long blocksT, void* stream) {
  if (!x_hi || ld_x % 8 != 0) return fail("dropout: input planes missing or pitch not a multiple of 8");
  if ((x_lo != nullptr) != (lo != nullptr)) return fail("dropout: input and output must carry the same planes");
  if (x_hi == hi) return fail("dropout: in-place operation is not supported (the transposed planes are written from the same tile)");
  Drop d;
  if (!make_drop(seed, p, cols, d)) return fail("dropout probability %f is not in [0, 1)", (double)p);
  DropPlanesProducer prod{static_cast<const __half*>(x_hi), static_cast<const __half*>(x_lo), rows, cols, ld_x, d};
  return launch_emit(prod, rows, cols, hi, lo, ld, hiT, loT, blocksT, static_cast<cudaStream_t>(stream));
}

int pn_t_dropout_f32(const float* x, long long rows, int cols, long long ldx, unsigned long long seed, float p, float* out,
                     long long ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || cols <= 0 || !x || !out || ldx < cols || ldo < cols) return fail("bad input to dropout_f32");
  Drop d;
  if (!make_drop(seed, p, cols, d)) return fail("dropout probability %f is not in [0, 1)", (double)p);
  dropout_f32_kernel<<<ew_grid(rows * d.groups), 256, 0, stream>>>(x, rows, cols, ldx, d, out, ldo);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_bn_relu_dot(const void* z_hi, const void* z_lo, long long rows, int cols, long long ld_z, const float* state,
                     const float* w, const float* b, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || cols <= 0 || !z_hi || ld_z % 8 != 0) return fail("bad input to bn_relu_dot");
  bn_relu_dot_kernel<<<ew_grid(rows * 32), 256, 0, stream>>>(static_cast<const __half*>(z_hi),
                                                             static_cast<const __half*>(z_lo), rows, cols, ld_z, state, w,
                                                             b, out);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_bn_relu_dot_loss(const void* z_hi, const void* z_lo, long long rows, int cols, long long ld_z, const float* state,
                          const float* w, const float* b, float* out, const float* targets, long long L,
                          const float* pos_weight, int loss_kind, float gamma, float alpha, float label_smoothing,
                          float grad_scale, float* g_out, double* loss_sum, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || cols <= 0 || !z_hi || ld_z % 8 != 0) return fail("bad input to bn_relu_dot_loss");
  if (loss_kind != PN_LOSS_BCE && loss_kind != PN_LOSS_FOCAL) return fail("loss_kind must be PN_LOSS_BCE or PN_LOSS_FOCAL");
  if (!targets || !g_out || !loss_sum || L <= 0 || rows % L != 0) return fail("bn_relu_dot_loss: targets / g_out / loss_sum / L");
  if (loss_kind == PN_LOSS_FOCAL && pos_weight) return fail("FocalLoss takes no pos_weight (protnote/utils/losses.py:171-213)");
  if (loss_kind == PN_LOSS_FOCAL && !(gamma >= 0.f)) return fail("focal gamma must be >= 0");
  LossSpec ls;
  ls.kind = loss_kind; ls.gamma = gamma; ls.alpha = alpha; ls.label_smoothing = label_smoothing; ls.grad_scale = grad_scale;
  ls.targets = targets; ls.pos_weight = pos_weight; ls.L = L; ls.g_out = g_out; ls.loss_sum = loss_sum;
  bn_relu_dot_loss_kernel<<<ew_grid(rows * 32), 256, 0, stream>>>(static_cast<const __half*>(z_hi),
                                                                  static_cast<const __half*>(z_lo), rows, cols, ld_z, state,
                                                                  w, b, out, ls);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_pair_hidden(const float* a, long long B, const float* c, long long L, int H, const float* state, void* hi,
                     void* lo, long long ld, void* hiT, void* loT, long long blocksT, void* stream) {
  if (B <= 0 || L <= 0) return fail("empty pair grid");
  PairHiddenProducer p{a, c, L, B * L, H, state};
  return launch_emit(p, B * L, H, hi, lo, ld, hiT, loT, blocksT, static_cast<cudaStream_t>(stream));
}

int pn_t_pair_product(const float* p, long long B, const float* t, long long L, int d, void* hi, void* lo, long long ld,
                      void* hiT, void* loT, long long blocksT, void* stream) {
  if (B <= 0 || L <= 0 || !p || !t) return fail("empty pair grid");
  PairProdProducer prod{p, t, L, B * L, d};
  return launch_emit(prod, B * L, d, hi, lo, ld, hiT, loT, blocksT, static_cast<cudaStream_t>(stream));
}

int pn_t_pair_add(const void* x_hi, const void* x_lo, long long ld_x, const float* a, long long B, const float* c,
                  long long L, int H, void* hi, void* lo, long long ld, void* stream) {
  if (B <= 0 || L <= 0 || !a || !c) return fail("empty pair grid");
  if (!x_hi || ld_x % 8 != 0 || ld_x < H) return fail("pair_add: planes missing or bad pitch");
  if ((x_lo != nullptr) != (lo != nullptr)) return fail("pair_add: input and output must carry the same planes");
  PairAddProducer prod{static_cast<const __half*>(x_hi), static_cast<const __half*>(x_lo), ld_x, a, c, L, B * L, H};
  return launch_emit(prod, B * L, H, hi, lo, ld, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream));
}

int pn_t_pair_marginals(const void* g_hi, const void* g_lo, long long ld_g, const float* g_sc, long long B, long long L,
                        int cols, const float* wb, const float* wl, float* out_b, float* out_l, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B <= 0 || L <= 0 || cols <= 0 || !g_hi || ld_g % 8 != 0 || ld_g < cols) return fail("pair_marginals: bad planes");
  if (!out_b && !out_l) return fail("pair_marginals: no output requested");
  const int col_blocks = (cols + 255) / 256;
  const long long l_blocks = (L + 7) / 8;
  if (B > 2147483647LL || l_blocks > 65535 || col_blocks > 65535) return fail("pair_marginals: grid too large (B %lld, L %lld)", B, L);
  const __half* gh = static_cast<const __half*>(g_hi);
  const __half* gl = static_cast<const __half*>(g_lo);
  if (out_l) {
    const dim3 grid(col_blocks, (unsigned)l_blocks);
    if (gl) pair_marginal_l_kernel<true><<<grid, dim3(32, 8), 0, stream>>>(gh, gl, ld_g, g_sc, B, L, cols, wb, out_l);
    else pair_marginal_l_kernel<false><<<grid, dim3(32, 8), 0, stream>>>(gh, gl, ld_g, g_sc, B, L, cols, wb, out_l);
    g_launches++;
  }
  if (out_b) {
    const dim3 grid((unsigned)B, col_blocks);
    if (gl) pair_marginal_b_kernel<true><<<grid, dim3(32, 8), 0, stream>>>(gh, gl, ld_g, g_sc, B, L, cols, wl, out_b);
    else pair_marginal_b_kernel<false><<<grid, dim3(32, 8), 0, stream>>>(gh, gl, ld_g, g_sc, B, L, cols, wl, out_b);
    g_launches++;
  }
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_normalize_rows(const float* x, long long rows, int cols, float scale, float* y, float* inv_norm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || cols <= 0 || !x || !y || !inv_norm || !(scale != 0.f)) return fail("bad input to normalize_rows");
  normalize_rows_kernel<<<ew_grid(rows * 32), 256, 0, stream>>>(x, rows, cols, scale, y, inv_norm);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_normalize_rows_bwd(const float* y, const float* inv_norm, const float* dy, long long rows, int cols, float scale,
                            float* dx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || cols <= 0 || !y || !inv_norm || !dy || !dx || !(scale != 0.f)) return fail("bad input to normalize_rows_bwd");
  normalize_rows_bwd_kernel<<<ew_grid(rows * 32), 256, 0, stream>>>(y, inv_norm, dy, rows, cols, scale, dx);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_bwd_stats(const pn_bwd_src* src, double* sums, float* maxes, double* dw, double* db, float* gyl, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PN_TRY(check_src(src));
  const BwdSrc s = to_src(*src);
  PN_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * s.cols, stream));
  PN_CUDA(cudaMemsetAsync(maxes, 0, 8, stream));
  if (s.kind == 1) {
    if (!dw || !db) return fail("kind 1 needs dw and db");
    PN_CUDA(cudaMemsetAsync(dw, 0, sizeof(double) * s.cols, stream));
    PN_CUDA(cudaMemsetAsync(db, 0, sizeof(double), stream));
  }
  const int col_blocks = (s.cols + g_stats_tx * 8 - 1) / (g_stats_tx * 8);
  // kind 2 walks a slab of LABELS for every protein (pn_train.cuh)
  const long long extent = s.kind == 2 ? s.L : s.rows;
  // kind 2: one step of labels per thread (TY row phases x kRif rows), see bwd_stats_kernel
  long long per = s.kind == 2 ? (256 / g_stats_tx) * kRif : slab_rows(s.rows, col_blocks);
  long long slabs = (extent + per - 1) / per;
  if (slabs > 65535) {
    if (s.kind == 2) return fail("too many label rows for one launch (%lld)", s.L);
    per = (extent + 65534) / 65535;
    per = (per + 31) / 32 * 32;
    slabs = (extent + per - 1) / per;
  }
  const dim3 grid(col_blocks, (unsigned)slabs), block(g_stats_tx, 256 / g_stats_tx);
  unsigned* mx = reinterpret_cast<unsigned*>(maxes);
  const bool lo = src_has_lo(s);
  if (s.kind == 0 && lo) bwd_stats_kernel<0, true><<<grid, block, 0, stream>>>(s, per, sums, mx, dw, db, gyl);
  else if (s.kind == 0) bwd_stats_kernel<0, false><<<grid, block, 0, stream>>>(s, per, sums, mx, dw, db, gyl);
  else if (s.kind == 1 && lo) bwd_stats_kernel<1, true><<<grid, block, 0, stream>>>(s, per, sums, mx, dw, db, gyl);
  else if (s.kind == 1) bwd_stats_kernel<1, false><<<grid, block, 0, stream>>>(s, per, sums, mx, dw, db, gyl);
  else if (lo) bwd_stats_kernel<2, true><<<grid, block, 0, stream>>>(s, per, sums, mx, dw, db, gyl);
  else bwd_stats_kernel<2, false><<<grid, block, 0, stream>>>(s, per, sums, mx, dw, db, gyl);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_bwd_scale(const double* sums, const float* maxes, const float* state, double count, int cols, float* sc_out,
                   float* means, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (cols <= 0 || !(count > 0) || !means) return fail("bad backward statistics");
  bwd_scale_kernel<<<1, 256, 0, stream>>>(sums, reinterpret_cast<const unsigned*>(maxes), state, count, cols, sc_out, means);
  g_launches++;
  PN_CUDA(cudaGetLastError());
  return 0;
}

int pn_t_bwd_apply(const pn_bwd_src* src, const float* means, const float* sc_out, void* hi, void* lo, long long ld,
                   void* hiT, void* loT, long long blocksT, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PN_TRY(check_src(src));
  if (!means) return fail("means missing");
  const BwdSrc s = to_src(*src);
  if (s.kind == 0) return launch_emit(BwdApplyProducer<0>{s, means, sc_out}, s.rows, s.cols, hi, lo, ld, hiT, loT, blocksT, stream);
  if (s.kind == 1) return launch_emit(BwdApplyProducer<1>{s, means, sc_out}, s.rows, s.cols, hi, lo, ld, hiT, loT, blocksT, stream);
  return launch_emit(BwdApplyProducer<2>{s, means, sc_out}, s.rows, s.cols, hi, lo, ld, hiT, loT, blocksT, stream);
}

int pn_t_bwd_apply_pair(const pn_bwd_src* src, const float* means, long long B, const float* gyl, const double* a_stats,
                        double* da64, float* da, float* dc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PN_TRY(check_src(src));
  if (src->kind != 2 || B <= 0 || B * src->L != src->rows) return fail("pair backward needs a kind-2 source with rows == B * L");
  if (!means) return fail("means missing");
  if (B > 65535) return fail("too many proteins for one launch (%lld)", B);
  const BwdSrc s = to_src(*src);
  PN_CUDA(cudaMemsetAsync(da64, 0, sizeof(double) * B * s.cols, stream));
  const int col_blocks = (s.cols + 255) / 256;
  const long long l_blocks = (s.L + 7) / 8;
  if (l_blocks > 65535) return fail("too many label rows for one launch (%lld)", s.L);
  const bool lo = src_has_lo(s);
  if (gyl && a_stats) {   // dc from the statistics pass's by-product: no second pass over g
    pair_dc_fixup_kernel<<<ew_grid(s.L * s.cols), 256, 0, stream>>>(gyl, s.c, a_stats, s.state, means, B, s.L, s.cols, dc);
  } else if (lo) {
    pair_dc_kernel<true><<<dim3(col_blocks, (unsigned)l_blocks), dim3(32, 8), 0, stream>>>(s, means, B, dc);
  } else {
    pair_dc_kernel<false><<<dim3(col_blocks, (unsigned)l_blocks), dim3(32, 8), 0, stream>>>(s, means, B, dc);
  }
  long long per = 1024;
  long long slabs = (s.L + per - 1) / per;
  if (slabs > 65535) return fail("too many label rows for one launch (%lld)", s.L);
  const dim3 da_grid((unsigned)B, col_blocks, (unsigned)slabs);
  if (lo) pair_da_kernel<true><<<da_grid, dim3(32, 8), 0, stream>>>(s, means, B, per, da64);
  else pair_da_kernel<false><<<da_grid, dim3(32, 8), 0, stream>>>(s, means, B, per, da64);
  f64_to_f32_kernel<<<ew_grid(B * s.cols), 256, 0, stream>>>(da64, da, B * s.cols);
  g_launches += 3;
  PN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
