// Element-wise / layout kernels around the tensor-core engine.  All of them are pure HBM streaming work:
// one thread handles 8 consecutive elements of the fastest axis (two 16-byte loads, one or two 16-byte stores),
// rows are 128-byte aligned, no shared memory is needed because nothing is re-read.
#pragma once

#include "pn_gemm.cuh"

namespace pn {

// ------------------------------------------------------------------------------------------------
// weight packing (once per weight version)
// ------------------------------------------------------------------------------------------------
// max |w| over `rows` rows of `span` consecutive elements, row pitch `pitch`
__global__ void absmax_kernel(const float* __restrict__ w, long long rows, long long span, long long pitch,
                              unsigned* __restrict__ out) {
  float m = 0.f;
  const long long n = rows * span;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = fabsf(w[(i / span) * pitch + (i % span)]);
    if (a == a) m = fmaxf(m, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like uints
}

// power-of-two scale that puts the largest |w| into [2^7, 2^8): hi AND lo fp16 planes stay normal for every
// weight within 2^-11 of the largest one, and nothing overflows.
__device__ __forceinline__ float weight_scale_from_absmax(unsigned bits) {
  const float m = __uint_as_float(bits);
  if (!(m > 0.f)) return 1.f;
  int e;
  frexpf(m, &e);   // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 8 - e);
}

// How the strict-mode engine will walk a packed weight row (the K loop of gemm_kernel): k-blocks of `bk` elements,
// `cblocks` per tap, accumulator chunks of `chunk_kblocks` k-blocks.  beta == 0 -> no compensation.
struct TruncComp {
  int bk, cblocks, chunk_kblocks, num_kblocks;
  float beta;
};

// dst[n][tap*cpad + c] = split(w[n*sn + c*sc + tap*st] * s * (1 + beta * r)) ; zero elsewhere in [0, ld)
// Linear (out,in):         taps=1, sn=in, sc=1, st=0, cin=in, cpad=ld
// Conv1d (out,in,k):       taps=k, sn=in*k, sc=k, st=1
//
// Truncation compensation.  Every tcgen05 accumulator add rounds the running sum toward zero, i.e. shrinks it by a factor
// (1 - beta) on average (beta = 3.3e-8 measured on B200, tools/trunc_comp_probe.py).  A product that enters the
// accumulator r adds before the chunk is promoted to fp32 registers therefore arrives scaled by (1 - beta)^r: the engine
// computes, in expectation, the dot product with weights w[k] * (1 - beta * r(k)), a deterministic sawtooth along K that
// is the same for every output row and does not average out downstream.  r(k) is known at pack time (strict-mode issue
// order inside a k-block: all lo*hi / hi*lo products first, then the hi*hi products, one MMA per 16 K-elements), so the
// weight is pre-multiplied by (1 + beta * r(k)).  What remains of the truncation is zero-mean noise of the size of
// round-to-nearest noise.
__global__ void pack_weight_kernel(const float* __restrict__ w, int N, int cin, int taps, long long sn, long long sc,
                                   long long st, int cpad, int ld, const unsigned* __restrict__ absmax,
                                   float* __restrict__ scale_out, __half* __restrict__ hi, __half* __restrict__ lo,
                                   TruncComp tc) {
  const float s = weight_scale_from_absmax(*absmax);
  if (blockIdx.x == 0 && threadIdx.x == 0) *scale_out = s;
  const long long total = (long long)N * ld;
  const int steps = tc.bk / 16;   // MMAs per pass per k-block
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ld);
    const int k = (int)(i % ld);
    const int tap = k / cpad, c = k % cpad;
    float v = 0.f;
    if (tap < taps && c < cin) {
      v = w[n * sn + c * sc + tap * st] * s;
      if (tc.beta != 0.f) {
        const int kb = tap * tc.cblocks + c / tc.bk;
        const int chunk0 = kb / tc.chunk_kblocks * tc.chunk_kblocks;
        const int nc = min(tc.chunk_kblocks, tc.num_kblocks - chunk0);
        const int q = kb - chunk0;
        const int ks = (c % tc.bk) / 16;
        const int r = 3 * steps * (nc - q) - 2 * steps - ks;   // adds from this product's own add to the end of the chunk
        v = fmaf(v, tc.beta * (float)r, v);
      }
    }
    __half h, l;
    split_f16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// Folds (conv/linear bias) + eval-mode BatchNorm + the weight scale into one per-column affine:
//   y = (acc / s + bias - mean) * g / sqrt(var + eps) + beta  =  acc * scale + shift
// Every pointer may be null (-> identity for that term).  BatchNorm eval: reference
// protnote/models/protein_encoders.py:35-37,47-50 (eps 1e-3) and protnote/models/ProtNote.py:364-365 (eps 1e-5).
__global__ void fold_affine_kernel(int n, const float* __restrict__ bias, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ mean,
                                   const float* __restrict__ var, float eps, const float* __restrict__ wscale,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // evaluated in fp64 and rounded once: these two numbers multiply / offset every output of the column, so their rounding
  // is a coherent error of the whole column (the reference's fp32 BatchNorm has several roundings here; fp64 is closer to
  // the exact value than either association order of fp32)
  double gs = 1.0;
  if (var) gs = 1.0 / sqrt((double)var[i] + (double)eps);
  if (gamma) gs *= (double)gamma[i];
  const double inv_s = wscale ? 1.0 / (double)*wscale : 1.0;
  scale[i] = (float)(gs * inv_s);
  shift[i] = (float)(((bias ? (double)bias[i] : 0.0) - (mean ? (double)mean[i] : 0.0)) * gs + (beta ? (double)beta[i] : 0.0));
}

// ------------------------------------------------------------------------------------------------
// fp64 protein-side head (strict mode).  Everything a protein contributes to its 32K logits goes through one vector,
// a[b] = BN1-folded protein half of output layer 1: an error in a[b] shifts ALL logits of protein b
// coherently, and fp32-grade arithmetic leaves ~7e-7 |a|max there (2e-5 at |a| = 30) - the largest single term of the
// logit error, in the reference's fp32 path as well (profiles/r02_scorer_error_by_stage.txt).  The work is negligible
// (57 MFLOP per protein against 38 MFLOP per PAIR), so strict mode evaluates W_p and that half in fp64 on the CUDA cores:
//   y[m][n] = act((sum_k x[m][k] * w[n][k]) * scale[n] + shift[n]),  x fp64, w fp32 (the original weights), scale/shift fp64.
// 64 x 64 output tile per block of 256 threads (4 x 4 outputs per thread), K in steps of 16 through shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void fold_affine_f64_kernel(int n, const float* __restrict__ bias, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                       const float* __restrict__ var, double eps, bool with_shift,
                                       double* __restrict__ scale, double* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double gs = 1.0;
  if (var) gs = 1.0 / sqrt((double)var[i] + eps);
  if (gamma) gs *= (double)gamma[i];
  scale[i] = gs;
  shift[i] = with_shift ? ((bias ? (double)bias[i] : 0.0) - (mean ? (double)mean[i] : 0.0)) * gs + (beta ? (double)beta[i] : 0.0) : 0.0;
}

__global__ void f32_to_f64_rows_kernel(const float* __restrict__ x, long long rows, int cols, long long ldx,
                                       double* __restrict__ y, long long ldy) {
  const long long total = rows * ldy;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ldy;
    const int c = (int)(i % ldy);
    y[i] = c < cols ? (double)x[r * ldx + c] : 0.0;
  }
}

__global__ void __launch_bounds__(256) linear_f64_kernel(const double* __restrict__ x, long long M, int K, long long ldx,
                                                         const float* __restrict__ w, int N, long long ldw,
                                                         const double* __restrict__ scale, const double* __restrict__ shift,
                                                         int relu, double* __restrict__ y64, long long ldy64,
                                                         float* __restrict__ y32, long long ldy32) {
  __shared__ double xs[16][64 + 1];
  __shared__ double ws[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // tx -> n, ty -> m
  const long long m0 = (long long)blockIdx.y * 64;
  const int n0 = blockIdx.x * 64;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    // 64 rows x 16 k of each operand: 1024 elements, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + e * 256;
      const int r = idx >> 4, kk = idx & 15;
      const long long m = m0 + r;
      const int n = n0 + r;
      xs[kk][r] = (m < M && k0 + kk < K) ? x[m * ldx + k0 + kk] : 0.0;
      ws[kk][r] = (n < N && k0 + kk < K) ? (double)w[(long long)n * ldw + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double xv[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(xv[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      double v = acc[i][j] * (scale ? scale[n] : 1.0) + (shift ? shift[n] : 0.0);
      if (relu) v = v > 0.0 ? v : 0.0;
      if (y64) y64[m * ldy64 + n] = v;
      if (y32) y32[m * ldy32 + n] = (float)v;
    }
  }
}

// The same layer for SKINNY batches (M <= 64 rows: the reference's evaluation batch is 8 proteins): the tiled kernel above
// would run 48 blocks for N = 3072 and take 0.45 ms per layer (5 % of a 32-protein step, ncu profiles/r02_launch_shares.txt).
// Here one warp owns one output column and 8 rows: the lanes stride over K (w[n][k] coalesced, x rows L1-resident), a fixed
// shuffle tree reduces them - deterministic - and lane 0 finishes the column.  grid (ceil(N / 8), ceil(M / 8)), 256 threads.
__global__ void __launch_bounds__(256) linear_f64_skinny_kernel(const double* __restrict__ x, long long M, int K, long long ldx,
                                                                const float* __restrict__ w, int N, long long ldw,
                                                                const double* __restrict__ scale, const double* __restrict__ shift,
                                                                int relu, double* __restrict__ y64, long long ldy64,
                                                                float* __restrict__ y32, long long ldy32) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long m0 = (long long)blockIdx.y * 8;
  if (n >= N) return;
  double acc[8] = {};
  const float* wrow = w + (long long)n * ldw;
  for (int k = lane; k < K; k += 32) {
    const double wv = (double)__ldg(wrow + k);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const long long m = m0 + r;
      acc[r] = fma(m < M ? x[m * ldx + k] : 0.0, wv, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const long long m = m0 + r;
      if (m >= M) break;
      double v = acc[r] * (scale ? scale[n] : 1.0) + (shift ? shift[n] : 0.0);
      if (relu) v = v > 0.0 ? v : 0.0;
      if (y64) y64[m * ldy64 + n] = v;
      if (y32) y32[m * ldy32 + n] = (float)v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// activations: fp32 -> fp16 hi/lo planes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split8_store(const float (&v)[8], __half* hi, __half* lo) {
  __align__(16) __half2 h2[4];
  __align__(16) __half2 l2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __half h0, l0, h1, l1;
    split_f16(fmaxf(fminf(v[2 * j], 65504.f), -65504.f), h0, l0);
    split_f16(fmaxf(fminf(v[2 * j + 1], 65504.f), -65504.f), h1, l1);
    h2[j] = __halves2half2(h0, h1);
    l2[j] = __halves2half2(l0, l1);
  }
  *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(h2);
  if (lo) *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(l2);
}

// x [M][ldx] fp32 -> hi/lo [M][ld] (ld multiple of 8, columns >= K zero-filled)
__global__ void split_rows_kernel(const float* __restrict__ x, long long M, int K, long long ldx,
                                  __half* __restrict__ hi, __half* __restrict__ lo, int ld) {
  const int chunks = ld / 8;
  const long long total = M * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / chunks;
    const int k0 = (int)(i % chunks) * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < K) ? x[m * ldx + k0 + j] : 0.f;
    split8_store(v, hi + m * ld + k0, lo ? lo + m * ld + k0 : nullptr);
  }
}

// Sequence input [B][Cin][T] fp32 (reference layout, collators.py:123-133) -> channels-last [B][T][cpad] planes,
// with positions >= length zeroed (MaskedConv1D masks its input, protein_encoders.py:14).
__global__ void conv_input_kernel(const float* __restrict__ x, const long long* __restrict__ lengths, int B, int cin,
                                  int T, int cpad, __half* __restrict__ hi, __half* __restrict__ lo) {
  const long long total = (long long)B * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / T);
    const int t = (int)(i % T);
    const bool valid = (long long)t < lengths[b];
    for (int c0 = 0; c0 < cpad; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        v[j] = (valid && c < cin) ? x[((long long)b * cin + c) * T + t] : 0.f;   // coalesced over t per channel
      }
      split8_store(v, hi + i * cpad + c0, lo ? lo + i * cpad + c0 : nullptr);
    }
  }
}

// Same from token ids (uint8, one per residue): the one-hot the reference's collator builds on the host
// (protnote/data/collators.py:123-133, 80 bytes per residue) is generated here from 1 byte per residue.
// One thread writes 8 channels of one position; ids >= cin or positions >= length give an all-zero column.
__global__ void conv_input_tokens_kernel(const uint8_t* __restrict__ tokens, const long long* __restrict__ lengths, int B,
                                         int cin, int T, int cpad, __half* __restrict__ hi, __half* __restrict__ lo) {
  const int chunks = cpad / 8;
  const long long total = (long long)B * T * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / chunks;
    const int c0 = (int)(i % chunks) * 8;
    const int b = (int)(pos / T), t = (int)(pos % T);
    const int tok = ((long long)t < lengths[b]) ? (int)tokens[pos] : -1;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (tok == c0 + j && tok < cin) ? 1.f : 0.f;
    split8_store(v, hi + pos * cpad + c0, lo ? lo + pos * cpad + c0 : nullptr);
  }
}

// Layer 1 of the pair scorer after the exact split of Linear(2d -> H) over [p; t]
// (reference protnote/models/ProtNote.py:112-126,293 materialises [B*L, 2d]; here it never exists):
//   h1[(b,l)][k] = relu(a[b][k] + c[l][k])      a = BN1-folded protein half, c = BN1-scaled label half
// rows of the chunk are pairs (b0 + r / nl, l0 + r % nl); output fp16 planes [nb*nl][ld].
// One block = kPairRows consecutive pair rows; thread t = 8-column chunk t of every one of them (row and protein
// indices are per-block scalars: the first version spent four 64-bit div/mod per 8 elements and was bound by
// instruction issue at the power-capped clock, not by HBM).  grid (ceil(rows / kPairRows)), block = ld / 8 threads
// rounded up to a warp (<= 1024).
// (A label-major walk that keeps c[l] in registers while the chunk's proteins stream past it reads 10x less - the c rows are
// re-read once per protein here, 4.0 GB per 2^19-pair chunk - but writes 16 rows that lie 200 MB apart per thread and ran
// 5.6 % SLOWER on B200: profiles/r02_ab_prefetch_pairfeatures.txt.  The kernel runs at the HBM peak as it is, 6.5 TB/s.)
constexpr int kPairRows = 16;
__global__ void pair_features_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ c,
                                     long long ldc, int b0, int l0, int nl, long long rows, int H,
                                     __half* __restrict__ hi, __half* __restrict__ lo, int ld) {
  const int chunks = ld / 8;
  const bool vec = ((lda | ldc) & 3) == 0;
  const long long r_begin = (long long)blockIdx.x * kPairRows;
  long long bb = r_begin / nl;                 // one division per block
  int l = (int)(r_begin - bb * nl);
  for (int rr = 0; rr < kPairRows; ++rr) {
    const long long r = r_begin + rr;
    if (r >= rows) break;
    const float* arow = a + (b0 + bb) * lda;
    const float* crow = c + (long long)(l0 + l) * ldc;
    for (int ch = threadIdx.x; ch < chunks; ch += blockDim.x) {
      const int k0 = ch * 8;
      float v[8];
      if (k0 + 8 <= H && vec) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(arow + k0)), a1 = __ldg(reinterpret_cast<const float4*>(arow + k0 + 4));
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(crow + k0)), c1 = __ldg(reinterpret_cast<const float4*>(crow + k0 + 4));
        v[0] = a0.x + c0.x; v[1] = a0.y + c0.y; v[2] = a0.z + c0.z; v[3] = a0.w + c0.w;
        v[4] = a1.x + c1.x; v[5] = a1.y + c1.y; v[6] = a1.z + c1.z; v[7] = a1.w + c1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (k0 + j < H) ? arow[k0 + j] + crow[k0 + j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
      split8_store(v, hi + r * ld + k0, lo ? lo + r * ld + k0 : nullptr);
    }
    if (++l == nl) {
      l = 0;
      ++bb;
    }
  }
}

// [p * t] block of the concatenation_prod fusion (ProtNote.py:139-150): x[(b,l)][k] = P_e[b][k] * L_e[l][k]
__global__ void pair_product_kernel(const float* __restrict__ p, long long ldp, const float* __restrict__ t,
                                    long long ldt, int b0, int l0, int nl, long long rows, int D,
                                    __half* __restrict__ hi, __half* __restrict__ lo, int ld) {
  const int chunks = ld / 8;
  const long long total = rows * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / chunks;
    const int k0 = (int)(i % chunks) * 8;
    const float* pp = p + (b0 + r / nl) * ldp + k0;
    const float* tp = t + (l0 + r % nl) * ldt + k0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < D) ? pp[j] * tp[j] : 0.f;
    split8_store(v, hi + r * ld + k0, lo ? lo + r * ld + k0 : nullptr);
  }
}

// F.normalize(x, dim=-1, p=2) (eps 1e-12) of every row, times a power of two, as fp16 planes (one warp per row).
// Used by the 'similarity' fusion (ProtNote.py:281-284).
__global__ void normalize_split_kernel(const float* __restrict__ x, long long n, int d, float pow2,
                                       __half* __restrict__ hi, __half* __restrict__ lo, int ld) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* xr = x + row * d;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) ss = fmaf(xr[k], xr[k], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int k = lane; k < ld; k += 32) {
    __half h, l;
    split_f16(k < d ? xr[k] / denom * pow2 : 0.f, h, l);
    hi[row * ld + k] = h;
    if (lo) lo[row * ld + k] = l;
  }
}

__global__ void fill_kernel(float* __restrict__ dst, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = v;
}

// ------------------------------------------------------------------------------------------------
// reductions at the two ends of the path
// ------------------------------------------------------------------------------------------------
// Masked mean pool (protein_encoders.py:114-117): out[b][c] = sum_{t < len[b]} x[b][t][c] / len[b].
// grid (ceil(C/32), B), block (32, 8): lanes span channels (coalesced), the 8 rows stride over t.
__global__ void pool_mean_kernel(const float* __restrict__ x, long long ldx, const long long* __restrict__ lengths,
                                 int T, int C, float* __restrict__ out, long long ldo) {
  __shared__ double part[8][33];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  long long len = lengths[b];
  const int n = (int)(len < T ? len : T);
  double s = 0.0;   // fp64 accumulation: the sum over up to T positions is then exact to fp32 rounding
  if (c < C)
    for (int t = threadIdx.y; t < n; t += 8) s += (double)x[((long long)b * T + t) * ldx + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += part[j][threadIdx.x];
    out[(long long)b * ldo + c] = (float)(tot / (double)len);
  }
}

// logits = sum of the per-N-tile partial dots + output bias; k consecutive description rows of one label are
// ensembled in probability space, logit(mean_k sigmoid(x), eps=1e-7)  (ProtNote.py:308-322).
// partial rows are the chunk's pairs (b0 + r / nl, l0 + r % nl); nl and l0 are multiples of k.
__global__ void finalize_logits_kernel(const float* __restrict__ partial, int parts, const float* __restrict__ bias,
                                       int b0, int l0, int nl, long long rows, int k, float* __restrict__ logits,
                                       long long ld_logits) {
  const long long groups = rows / k;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups;
       g += (long long)gridDim.x * blockDim.x) {
    const long long r0 = g * k;
    const long long b = b0 + r0 / nl;
    const long long l = (l0 + r0 % nl) / k;
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      const float* pp = partial + (r0 + j) * parts;
      double xs = 0.0;
      for (int q = 0; q < parts; ++q) xs += (double)pp[q];
      xs += bias ? (double)*bias : 0.0;
      const float x = (float)xs;
      if (k == 1) {
        acc = x;
      } else {
        acc += 1.f / (1.f + expf(-x));
      }
    }
    if (k > 1) {
      float pm = acc / (float)k;
      const float lo = 1e-7f, hi = 1.f - 1e-7f;
      pm = fminf(fmaxf(pm, lo), hi);
      acc = logf(pm / (1.f - pm));
    }
    logits[b * ld_logits + l] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// evaluation post-processing on the device (what ProtNoteTrainer.evaluate does with the logits of every batch,
// protnote/models/ProtNoteTrainer.py:522-537 and calculate_tp_fn_fp :61-83): probabilities = sigmoid(logits),
// predictions = probabilities >= threshold, per-label true positives / false negatives / false positives ADDED to
// tp / fn / fp (integer-valued floats: exact and order-independent below 2^24).
// grid (ceil(L/256), row slabs), 256 threads: thread = one label column, rows of the slab in turn (coalesced).
// label_kind 1: int64 multihots (collators.py), 2: float32.
// ------------------------------------------------------------------------------------------------
__global__ void postprocess_counts_kernel(const float* __restrict__ logits, long long B, long long L, long long ld_logits,
                                          const void* __restrict__ labels, int label_kind, long long ld_labels,
                                          float threshold, float* __restrict__ probs, long long ld_probs,
                                          long long rows_per_slab, float* __restrict__ tp, float* __restrict__ fn,
                                          float* __restrict__ fp) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= L) return;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  const long long r1 = r0 + rows_per_slab < B ? r0 + rows_per_slab : B;
  float ntp = 0.f, nfn = 0.f, nfp = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float p = 1.f / (1.f + expf(-logits[r * ld_logits + c]));
    if (probs) probs[r * ld_probs + c] = p;
    if (label_kind != 0) {
      const float y = label_kind == 1 ? (float)reinterpret_cast<const long long*>(labels)[r * ld_labels + c]
                                      : reinterpret_cast<const float*>(labels)[r * ld_labels + c];
      const float pred = p >= threshold ? 1.f : 0.f;
      ntp += pred * y;
      nfn += (1.f - pred) * y;
      nfp += pred * (1.f - y);
    }
  }
  if (label_kind != 0) {
    if (ntp != 0.f) atomicAdd(tp + c, ntp);
    if (nfn != 0.f) atomicAdd(fn + c, nfn);
    if (nfp != 0.f) atomicAdd(fp + c, nfp);
  }
}

// top-k logits of every row (descending; ties: lower index first), k <= 64: k selection sweeps over the row (L2-resident)
// by one block per row.  The "identical top-k label indices" check of the north star runs on the device with this.
__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ logits, long long L, long long ld,
                                                        int k, float* __restrict__ values, int* __restrict__ indices) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ float last_v;
  __shared__ int last_i;
  const float* row = logits + (long long)blockIdx.x * ld;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) {
    last_v = __int_as_float(0x7f800000);   // +inf
    last_i = -1;
  }
  __syncthreads();
  for (int sel = 0; sel < k; ++sel) {
    const float lv = last_v;
    const int li = last_i;
    float best = -__int_as_float(0x7f800000);
    int besti = 0x7fffffff;
    for (long long c = t; c < L; c += 256) {
      const float v = row[c];
      // candidates strictly after (lv, li) in (value descending, index ascending) order; NaNs are never selected
      const bool after = v < lv || (v == lv && (int)c > li);
      if (after && (v > best || (v == best && (int)c < besti))) {
        best = v;
        besti = (int)c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) {
        best = ov;
        besti = oi;
      }
    }
    if (lane == 0) {
      sv[warp] = best;
      si[warp] = besti;
    }
    __syncthreads();
    if (t == 0) {
      for (int w = 1; w < 8; ++w)
        if (sv[w] > best || (sv[w] == best && si[w] < besti)) {
          best = sv[w];
          besti = si[w];
        }
      values[(long long)blockIdx.x * k + sel] = best;
      indices[(long long)blockIdx.x * k + sel] = besti == 0x7fffffff ? -1 : besti;
      last_v = best;
      last_i = besti;
    }
    __syncthreads();
  }
}

}  // namespace pn
