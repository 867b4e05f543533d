// HBM-streaming kernels of the TRAINING step (reference: ProtNote.forward with self.training, protnote/models/
// ProtNote.py:168-334; BatchNorm1d batch statistics of torchvision MLP / get_mlp, ProtNote.py:63-81,337-378).
//
// The contractions of the training step (forward Linear, dgrad, wgrad) run on the tensor-core engine of pn_gemm.cuh.
// What is left is memory-bound work over [rows][cols] activation tensors stored as fp16 hi/lo planes:
//   column statistics (sum, sum of squares)            -> BatchNorm batch mean / variance
//   normalise + ReLU (+ transposed copy for wgrad)      -> next layer's operand
//   BatchNorm+ReLU backward: two column sums, then the gradient w.r.t. the Linear output (+ transposed copy)
//   layer 1 of the pair scorer, z1[b,l] = a[b] + c[l]: generated on the fly in forward, reduced to (da, dc) in backward
// Every kernel moves 16 bytes per thread per plane, rows are 128-byte aligned, column sums accumulate in fp64.
//
// Transposed copies.  wgrad contracts over ROWS (dW = g^T x), and the engine wants both operands K-major, so every
// tensor that feeds a wgrad is also written transposed, in the K-BLOCKED layout [ceil(rows/64)][cols][64]: the 64 rows
// of a block are the contiguous 128 bytes the tensor core's k-block wants, and the pieces of consecutive columns are
// adjacent, so an operand tile of a k-block is one contiguous 16-32 KB run (a plain [cols][rows] matrix would put
// every 128-byte piece on a different page: measured 7x slower wgrad from TLB / DRAM-page misses).  A 64x64 tile goes
// through shared memory as 32-bit words holding two vertically adjacent halves, so both the row-major and the
// transposed stores are 16-byte vectors.
#pragma once

#include "pn_kernels.cuh"

namespace pn {

constexpr int kTileDim = 64;
constexpr int kTilePitch = 33;
// rows a reduction thread keeps in flight per iteration (2 measured faster than 4 on B200: registers -> occupancy)
constexpr int kRif = 2;   // words per transposed tile row (odd pitch: conflict-free column reads)

__device__ __forceinline__ void load8(const __half* __restrict__ hi, const __half* __restrict__ lo, float (&v)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi);
  const __half2* h2 = reinterpret_cast<const __half2*>(&h);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
  if (lo) {
    const uint4 l = *reinterpret_cast<const uint4*>(lo);
    const __half2* l2 = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(l2[j]);
      v[2 * j] += f.x;
      v[2 * j + 1] += f.y;
    }
  }
}

__device__ __forceinline__ void load8_f32(const float* __restrict__ p, int n_valid, float (&v)[8]) {
  if (n_valid >= 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < n_valid ? __ldg(p + j) : 0.f;
  }
}

// BatchNorm state of one layer: float state[4][cols] = scale (gamma * invstd), shift (beta - mean * scale), mean, invstd
struct BnVec {
  float scale[8], shift[8], mean[8], invstd[8];
};
__device__ __forceinline__ void load_bn(const float* __restrict__ state, int cols, int c0, BnVec& b) {
  const int n = cols - c0;
  load8_f32(state + c0, n, b.scale);
  load8_f32(state + cols + c0, n, b.shift);
  load8_f32(state + 2 * (long long)cols + c0, n, b.mean);
  load8_f32(state + 3 * (long long)cols + c0, n, b.invstd);
}

// ------------------------------------------------------------------------------------------------
// producers of the tile emitter.  A thread owns 8 consecutive columns of two vertically adjacent rows.  The work is
// split in three so the emitter can software-pipeline a strip of tiles: init() loads the per-column constants once
// per strip, load() only issues the 16-byte plane loads of a tile (the NEXT tile's loads are in flight while the
// current one is converted and stored), eval() does the arithmetic.  LO = the tensors carry lo planes (strict mode).
// ------------------------------------------------------------------------------------------------
struct Raw8 {
  uint4 h, l;
};
template <bool LO>
__device__ __forceinline__ void raw_load(const __half* __restrict__ hi, const __half* __restrict__ lo, long long off,
                                         Raw8& r) {
  r.h = *reinterpret_cast<const uint4*>(hi + off);
  if (LO) r.l = *reinterpret_cast<const uint4*>(lo + off);
}
template <bool LO>
__device__ __forceinline__ void raw_to_f32(const Raw8& r, float (&v)[8]) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&r.h);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
  if (LO) {
    const __half2* l2 = reinterpret_cast<const __half2*>(&r.l);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g = __half22float2(l2[j]);
      v[2 * j] += g.x;
      v[2 * j + 1] += g.y;
    }
  }
}
__device__ __forceinline__ void zero8(float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
}

struct SplitProducer {          // x * sc
  const float* x; long long rows; int cols; long long ldx; const float* sc;
  struct Ctx { float s; };
  struct Raw { float a[8], b[8]; };
  __device__ __forceinline__ void init(int c0, Ctx& k) const { k.s = sc ? __ldg(sc) : 1.f; }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      q.a[j] = (r < rows && c0 + j < cols) ? x[r * ldx + c0 + j] : 0.f;
      q.b[j] = (r + 1 < rows && c0 + j < cols) ? x[(r + 1) * ldx + c0 + j] : 0.f;
    }
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      va[j] = q.a[j] * k.s;
      vb[j] = q.b[j] * k.s;
    }
  }
};

struct BnReluProducer {         // relu(z * scale + shift)
  const __half* zh; const __half* zl; long long rows; int cols; long long ld; const float* state;
  struct Ctx { float sc[8], sf[8]; };
  struct Raw { Raw8 a, b; };
  __device__ __forceinline__ void init(int c0, Ctx& k) const {
    zero8(k.sc);
    zero8(k.sf);
    if (c0 < cols) {
      load8_f32(state + c0, cols - c0, k.sc);
      load8_f32(state + cols + c0, cols - c0, k.sf);
    }
  }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
    if (c0 >= cols) return;
    if (r < rows) raw_load<LO>(zh, zl, r * ld + c0, q.a);
    if (r + 1 < rows) raw_load<LO>(zh, zl, (r + 1) * ld + c0, q.b);
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
    zero8(va);
    zero8(vb);
    if (c0 >= cols) return;
    if (r < rows) {
      raw_to_f32<LO>(q.a, va);
#pragma unroll
      for (int j = 0; j < 8; ++j) va[j] = c0 + j < cols ? fmaxf(fmaf(va[j], k.sc[j], k.sf[j]), 0.f) : 0.f;
    }
    if (r + 1 < rows) {
      raw_to_f32<LO>(q.b, vb);
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = c0 + j < cols ? fmaxf(fmaf(vb[j], k.sc[j], k.sf[j]), 0.f) : 0.f;
    }
  }
};

// relu(x * scale + shift) of an fp32 channels-last activation tensor [B*T][ld], zeroed at positions >= length:
// BatchNorm1d (batch statistics) + ReLU of the sequence encoder followed by the input mask of the next MaskedConv1D
// (protein_encoders.py:35-37,47-50 then :14)
struct BnReluMaskF32Producer {
  const float* x; long long rows; int cols; long long ldx; const float* state; const long long* lengths; int T;
  struct Ctx { float sc[8], sf[8]; };
  struct Raw { float a[8], b[8]; };
  __device__ __forceinline__ void init(int c0, Ctx& k) const {
    zero8(k.sc);
    zero8(k.sf);
    if (c0 < cols) {
      load8_f32(state + c0, cols - c0, k.sc);
      load8_f32(state + cols + c0, cols - c0, k.sf);
    }
  }
  __device__ __forceinline__ bool valid(long long r) const {
    return r < rows && (lengths == nullptr || (r % T) < lengths[r / T]);
  }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
    if (c0 >= cols) return;
    if (valid(r)) load8_f32(x + r * ldx + c0, cols - c0, q.a);
    if (valid(r + 1)) load8_f32(x + (r + 1) * ldx + c0, cols - c0, q.b);
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
    zero8(va);
    zero8(vb);
    if (c0 >= cols) return;
    if (valid(r)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) va[j] = c0 + j < cols ? fmaxf(fmaf(q.a[j], k.sc[j], k.sf[j]), 0.f) : 0.f;
    }
    if (valid(r + 1)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = c0 + j < cols ? fmaxf(fmaf(q.b[j], k.sc[j], k.sf[j]), 0.f) : 0.f;
    }
  }
};

struct PairHiddenProducer {     // relu((a[b] + c[l]) * scale + shift), row r = b * L + l   (ProtNote.py:112-126 + layer 1)
  const float* a; const float* c; long long L; long long rows; int cols; const float* state;
  struct Ctx { float sc[8], sf[8]; };
  struct Raw { float ca[8], cb[8]; int b0, b1; };      // label halves of the two rows + their protein indices
  __device__ __forceinline__ void init(int c0, Ctx& k) const {
    zero8(k.sc);
    zero8(k.sf);
    if (c0 < cols) {
      load8_f32(state + c0, cols - c0, k.sc);
      load8_f32(state + cols + c0, cols - c0, k.sf);
    }
  }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
    if (c0 >= cols || r >= rows) return;
    const long long b0 = r / L;                  // one division per thread and tile; the second row follows from it
    long long l0 = r - b0 * L, l1 = l0 + 1, b1 = b0;
    if (l1 == L) {
      l1 = 0;
      ++b1;
    }
    q.b0 = (int)b0;
    q.b1 = (int)b1;
    load8_f32(c + l0 * cols + c0, cols - c0, q.ca);
    if (r + 1 < rows) load8_f32(c + l1 * cols + c0, cols - c0, q.cb);
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
    zero8(va);
    zero8(vb);
    if (c0 >= cols || r >= rows) return;
    float av[8];
    load8_f32(a + (long long)q.b0 * cols + c0, cols - c0, av);     // B rows of a: L1 / L2 resident
#pragma unroll
    for (int j = 0; j < 8; ++j) va[j] = c0 + j < cols ? fmaxf(fmaf(av[j] + q.ca[j], k.sc[j], k.sf[j]), 0.f) : 0.f;
    if (r + 1 < rows) {
      if (q.b1 != q.b0) load8_f32(a + (long long)q.b1 * cols + c0, cols - c0, av);
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = c0 + j < cols ? fmaxf(fmaf(av[j] + q.cb[j], k.sc[j], k.sf[j]), 0.f) : 0.f;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Dropout inside W_p / W_l / output_layer (OUTPUT_MLP_DROPOUT, ProtNote.py:70,80,101 -> torchvision MLP / get_mlp :369-371).
// The keep mask is a pure function of (seed, row, column): a counter-based generator (splitmix64 finaliser over the
// counter of an 8-column group, 16 random bits per element), so forward and backward regenerate the same mask from 20 bytes
// of state and nothing is stored.  keep <=> bits >= thr, thr = round(p * 65536); kept values are scaled by
// 65536 / (65536 - thr) (the inverse of the quantised keep probability, so the expectation is exact).
// ------------------------------------------------------------------------------------------------
struct Drop {
  unsigned long long seed;
  unsigned thr;
  float scale;
  int groups;            // ceil(cols / 8): 8-column groups per row
};
__host__ __device__ __forceinline__ unsigned long long drop_mix64(unsigned long long x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
// bit j = keep element (r, c0 + j); c0 is a multiple of 8
__device__ __forceinline__ unsigned drop_keep8(const Drop& d, long long r, int c0) {
  const unsigned long long ctr = 2ULL * ((unsigned long long)r * (unsigned long long)d.groups + (unsigned long long)(c0 >> 3));
  const unsigned long long u0 = drop_mix64(d.seed + (ctr + 1ULL) * 0x9e3779b97f4a7c15ULL);
  const unsigned long long u1 = drop_mix64(d.seed + (ctr + 2ULL) * 0x9e3779b97f4a7c15ULL);
  unsigned m = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m |= (((unsigned)(u0 >> (16 * j)) & 0xffffu) >= d.thr ? 1u : 0u) << j;
    m |= (((unsigned)(u1 >> (16 * j)) & 0xffffu) >= d.thr ? 1u : 0u) << (4 + j);
  }
  return m;
}

struct DropPlanesProducer {     // x * keep * scale, x given as planes (true value x / sc: the scale stays with the tensor)
  const __half* xh; const __half* xl; long long rows; int cols; long long ld; Drop d;
  struct Ctx { int unused; };
  struct Raw { Raw8 a, b; };
  __device__ __forceinline__ void init(int c0, Ctx& k) const { k.unused = 0; }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
    if (c0 >= cols) return;
    if (r < rows) raw_load<LO>(xh, xl, r * ld + c0, q.a);
    if (r + 1 < rows) raw_load<LO>(xh, xl, (r + 1) * ld + c0, q.b);
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
    zero8(va);
    zero8(vb);
    if (c0 >= cols) return;
    if (r < rows) {
      raw_to_f32<LO>(q.a, va);
      const unsigned m = drop_keep8(d, r, c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) va[j] = (c0 + j < cols && ((m >> j) & 1u)) ? va[j] * d.scale : 0.f;
    }
    if (r + 1 < rows) {
      raw_to_f32<LO>(q.b, vb);
      const unsigned m = drop_keep8(d, r + 1, c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = (c0 + j < cols && ((m >> j) & 1u)) ? vb[j] * d.scale : 0.f;
    }
  }
};

// the same mask on a small fp32 matrix (the output of a projection head, [n][latent]): one thread per 8-column group
__global__ void dropout_f32_kernel(const float* __restrict__ x, long long rows, int cols, long long ldx, Drop d,
                                   float* __restrict__ out, long long ldo) {
  const long long total = rows * d.groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d.groups;
    const int c0 = (int)(i % d.groups) * 8;
    const unsigned m = drop_keep8(d, r, c0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < cols) out[r * ldo + c0 + j] = ((m >> j) & 1u) ? x[r * ldx + c0 + j] * d.scale : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// FEATURE_FUSION concatenation_prod in training (ProtNote.py:140-150): the third feature block p (.) t is not a sum of a
// protein and a label term, so its share of layer 1, x = (p (.) t) Wx^T, is a real [B*L, d] x [d, H] GEMM:
//   PairProdProducer   q[b*L + l] = P_e[b] (.) L_e[l] as planes (+ transposed planes for d Wx = g_z1^T q)
//   PairAddProducer    z1[b*L + l] = x[b*L + l] + a[b] + c[l] as planes - from there layer 1 is an ordinary layer
//   pair_marginals     out_b[b] = sum_l G[b,l] (.) wl[l],  out_l[l] = sum_b G[b,l] (.) wb[b]  (unit weights when null):
//                      da / dc from g_z1, and the product block's share of d P_e / d L_e from g_q
// ------------------------------------------------------------------------------------------------
struct PairProdProducer {
  const float* p; const float* t; long long L; long long rows; int cols;
  struct Ctx { int unused; };
  struct Raw { float ta[8], tb[8]; int b0, b1; };
  __device__ __forceinline__ void init(int c0, Ctx& k) const { k.unused = 0; }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
    if (c0 >= cols || r >= rows) return;
    const long long b0 = r / L;
    long long l0 = r - b0 * L, l1 = l0 + 1, b1 = b0;
    if (l1 == L) {
      l1 = 0;
      ++b1;
    }
    q.b0 = (int)b0;
    q.b1 = (int)b1;
    load8_f32(t + l0 * cols + c0, cols - c0, q.ta);
    if (r + 1 < rows) load8_f32(t + l1 * cols + c0, cols - c0, q.tb);
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
    zero8(va);
    zero8(vb);
    if (c0 >= cols || r >= rows) return;
    float pv[8];
    load8_f32(p + (long long)q.b0 * cols + c0, cols - c0, pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) va[j] = c0 + j < cols ? pv[j] * q.ta[j] : 0.f;
    if (r + 1 < rows) {
      if (q.b1 != q.b0) load8_f32(p + (long long)q.b1 * cols + c0, cols - c0, pv);
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = c0 + j < cols ? pv[j] * q.tb[j] : 0.f;
    }
  }
};

struct PairAddProducer {
  const __half* xh; const __half* xl; long long ld; const float* a; const float* c; long long L; long long rows; int cols;
  struct Ctx { int unused; };
  struct Raw { Raw8 xa, xb; float ca[8], cb[8]; int b0, b1; };
  __device__ __forceinline__ void init(int c0, Ctx& k) const { k.unused = 0; }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
    if (c0 >= cols || r >= rows) return;
    const long long b0 = r / L;
    long long l0 = r - b0 * L, l1 = l0 + 1, b1 = b0;
    if (l1 == L) {
      l1 = 0;
      ++b1;
    }
    q.b0 = (int)b0;
    q.b1 = (int)b1;
    raw_load<LO>(xh, xl, r * ld + c0, q.xa);
    load8_f32(c + l0 * cols + c0, cols - c0, q.ca);
    if (r + 1 < rows) {
      raw_load<LO>(xh, xl, (r + 1) * ld + c0, q.xb);
      load8_f32(c + l1 * cols + c0, cols - c0, q.cb);
    }
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
    zero8(va);
    zero8(vb);
    if (c0 >= cols || r >= rows) return;
    float av[8], xv[8];
    load8_f32(a + (long long)q.b0 * cols + c0, cols - c0, av);
    raw_to_f32<LO>(q.xa, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) va[j] = c0 + j < cols ? xv[j] + av[j] + q.ca[j] : 0.f;
    if (r + 1 < rows) {
      if (q.b1 != q.b0) load8_f32(a + (long long)q.b1 * cols + c0, cols - c0, av);
      raw_to_f32<LO>(q.xb, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = c0 + j < cols ? xv[j] + av[j] + q.cb[j] : 0.f;
    }
  }
};

// out_l[l][n] = (1/sc) * sum_b G[b*L + l][n] * (wb ? wb[b][n] : 1).  grid (ceil(cols/256), ceil(L/8)), block (32, 8):
// one thread per (label row, 8 columns) walks the B proteins (rows L apart, coalesced along the columns); fp64 sums.
template <bool LO>
__global__ void __launch_bounds__(256) pair_marginal_l_kernel(const __half* __restrict__ gh, const __half* __restrict__ gl,
                                                              long long ld, const float* __restrict__ sc, long long B,
                                                              long long L, int cols, const float* __restrict__ wb,
                                                              float* __restrict__ out) {
  const int c0 = blockIdx.x * 256 + threadIdx.x * 8;
  const long long l = (long long)blockIdx.y * 8 + threadIdx.y;
  if (c0 >= cols || l >= L) return;
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0;
  for (long long b = 0; b < B; ++b) {
    Raw8 q;
    float g[8], w[8];
    raw_load<LO>(gh, gl, (b * L + l) * ld + c0, q);
    raw_to_f32<LO>(q, g);
    if (wb) load8_f32(wb + b * cols + c0, cols - c0, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += (double)(wb ? g[j] * w[j] : g[j]);
  }
  const double inv = sc ? 1.0 / (double)__ldg(sc) : 1.0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (c0 + j < cols) out[l * cols + c0 + j] = (float)(acc[j] * inv);
}

// out_b[b][n] = (1/sc) * sum_l G[b*L + l][n] * (wl ? wl[l][n] : 1).  grid (B, ceil(cols/256)), block (32, 8): the 8 row
// phases of a block walk the L rows of protein b, fp64 partial sums, one shared-memory reduction (deterministic).
template <bool LO>
__global__ void __launch_bounds__(256) pair_marginal_b_kernel(const __half* __restrict__ gh, const __half* __restrict__ gl,
                                                              long long ld, const float* __restrict__ sc, long long B,
                                                              long long L, int cols, const float* __restrict__ wl,
                                                              float* __restrict__ out) {
  __shared__ double sh[8][8][32];        // [row phase][element][lane]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long b = blockIdx.x;
  const int c0 = blockIdx.y * 256 + tx * 8;
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0;
  if (c0 < cols) {
    for (long long l = ty; l < L; l += 8) {
      Raw8 q;
      float g[8], w[8];
      raw_load<LO>(gh, gl, (b * L + l) * ld + c0, q);
      raw_to_f32<LO>(q, g);
      if (wl) load8_f32(wl + l * cols + c0, cols - c0, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (double)(wl ? g[j] * w[j] : g[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[ty][j][tx] = acc[j];
  __syncthreads();
  const int i = ty * 32 + tx;            // column i of the block = lane i / 8, element i % 8
  const int c = blockIdx.y * 256 + i;
  if (c < cols) {
    double tot = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) tot += sh[y][i & 7][i >> 3];
    out[b * cols + c] = (float)(tot * (sc ? 1.0 / (double)__ldg(sc) : 1.0));
  }
}

// ------------------------------------------------------------------------------------------------
// FEATURE_FUSION similarity in training (ProtNote.py:281-284): logits = normalize(P_e) normalize(L_e)^T / temperature.
// One warp per row of a small fp32 matrix ([B][latent], [L][latent]); fp64 row sums.
//   forward   y = x * scale / max(|x|_2, 1e-12)  (F.normalize's eps), inv_norm[r] = 1 / max(|x|_2, 1e-12)
//   backward  dx = scale * inv_norm * (dy - u (u . dy)),  u = y / scale the unit row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, long long rows, int cols,
                                                             float scale, float* __restrict__ y,
                                                             float* __restrict__ inv_norm) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  double ss = 0.0;
  for (int c = lane; c < cols; c += 32) {
    const double v = (double)x[r * cols + c];
    ss += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const double n = sqrt(ss);
  const float inv = (float)(1.0 / (n > 1e-12 ? n : 1e-12));
  if (lane == 0) inv_norm[r] = inv;
  for (int c = lane; c < cols; c += 32) y[r * cols + c] = x[r * cols + c] * inv * scale;
}

__global__ void __launch_bounds__(256) normalize_rows_bwd_kernel(const float* __restrict__ y,
                                                                 const float* __restrict__ inv_norm,
                                                                 const float* __restrict__ dy, long long rows, int cols,
                                                                 float scale, float* __restrict__ dx) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float inv_scale = 1.f / scale;
  double dot = 0.0;
  for (int c = lane; c < cols; c += 32) dot += (double)(y[r * cols + c] * inv_scale) * (double)dy[r * cols + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float d = (float)dot, k = scale * inv_norm[r];
  for (int c = lane; c < cols; c += 32) dx[r * cols + c] = k * (dy[r * cols + c] - y[r * cols + c] * inv_scale * d);
}

// Sources of the BatchNorm+ReLU backward.  kind 0: g planes, z planes.  kind 1: g = g_logit[r] * w[n] (the gradient of
// the final Linear(H -> 1), never materialised), z planes.  kind 2: g planes, z = a[r / L] + c[r % L].
struct BwdSrc {
  int kind;
  long long rows; int cols;
  const __half* g_hi; const __half* g_lo; long long ld_g; const float* g_sc;
  const float* g_logit; const float* w;
  const __half* z_hi; const __half* z_lo; long long ld_z;
  const float* a; const float* c; long long L;
  const float* state;
};

// what one thread needs from memory for 8 columns of one row
struct BwdRaw {
  Raw8 g, z;
  float gl;
};
template <int KIND, bool LO>
__device__ __forceinline__ void bwd_raw_load(const BwdSrc& s, long long r, int c0, BwdRaw& q) {
  if (KIND == 1) q.gl = __ldg(s.g_logit + r);
  else raw_load<LO>(s.g_hi, s.g_lo, r * s.ld_g + c0, q.g);
  if (KIND != 2) raw_load<LO>(s.z_hi, s.z_lo, r * s.ld_z + c0, q.z);
}
// true-scale g_y = g * [relu active], xhat and the pre-activation for 8 columns of row r (columns >= cols give zeros)
template <int KIND, bool LO>
__device__ __forceinline__ void bwd_eval(const BwdSrc& s, const BwdRaw& q, long long r, int c0, const BnVec& b,
                                         const float (&wv)[8], float inv_gsc, float (&gy)[8], float (&xh)[8],
                                         float (&pre)[8]) {
  float g[8], z[8];
  if (KIND == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = q.gl * wv[j];
  } else {
    raw_to_f32<LO>(q.g, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv_gsc;
  }
  if (KIND == 2) {
    float av[8], cv[8];
    load8_f32(s.a + (r / s.L) * s.cols + c0, s.cols - c0, av);
    load8_f32(s.c + (r % s.L) * s.cols + c0, s.cols - c0, cv);
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = av[j] + cv[j];
  } else {
    raw_to_f32<LO>(q.z, z);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool ok = c0 + j < s.cols;
    pre[j] = ok ? fmaf(z[j], b.scale[j], b.shift[j]) : 0.f;
    gy[j] = (ok && pre[j] > 0.f) ? g[j] : 0.f;
    xh[j] = ok ? (z[j] - b.mean[j]) * b.invstd[j] : 0.f;
  }
}

// g_z = scale * (g_y - m1 - xhat * m2) * sc_out,  m = sums / count (fp32 [2][cols]).  Per column the thread keeps
//   A = scale * sc_out (* 1/g_sc, * w for kind 1),  D = scale * sc_out * invstd * m2,  E = scale * sc_out * m1,
// so g_z = A * g[masked] - E - D * (z - mean); the mask is (z * scale + shift > 0).
template <int KIND>
struct BwdApplyProducer {
  BwdSrc s; const float* means; const float* sc_out;
  struct Ctx { float sc[8], sf[8], mean[8], A[8], D[8], E[8]; };
  struct Raw { BwdRaw a, b; };
  __device__ __forceinline__ void init(int c0, Ctx& k) const {
    zero8(k.sc); zero8(k.sf); zero8(k.mean); zero8(k.A); zero8(k.D); zero8(k.E);
    if (c0 >= s.cols) return;
    float invstd[8], m1[8], m2[8], wv[8];
    load8_f32(s.state + c0, s.cols - c0, k.sc);
    load8_f32(s.state + s.cols + c0, s.cols - c0, k.sf);
    load8_f32(s.state + 2 * (long long)s.cols + c0, s.cols - c0, k.mean);
    load8_f32(s.state + 3 * (long long)s.cols + c0, s.cols - c0, invstd);
    load8_f32(means + c0, s.cols - c0, m1);
    load8_f32(means + s.cols + c0, s.cols - c0, m2);
    if (KIND == 1) load8_f32(s.w + c0, s.cols - c0, wv);
    const float inv_gsc = (KIND != 1 && s.g_sc) ? 1.f / __ldg(s.g_sc) : 1.f;
    const float so = sc_out ? __ldg(sc_out) : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = k.sc[j] * so;
      k.A[j] = t * (KIND == 1 ? wv[j] : inv_gsc);
      k.D[j] = t * invstd[j] * m2[j];
      k.E[j] = t * m1[j];
    }
  }
  template <bool LO>
  __device__ __forceinline__ void load(long long r, int c0, Raw& q) const {
    if (c0 >= s.cols) return;
    if (r < s.rows) bwd_raw_load<KIND, LO>(s, r, c0, q.a);
    if (r + 1 < s.rows) bwd_raw_load<KIND, LO>(s, r + 1, c0, q.b);
  }
  template <bool LO>
  __device__ __forceinline__ void one(const Ctx& k, const BwdRaw& q, long long r, int c0, float (&v)[8]) const {
    float g[8], z[8];
    if (KIND == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = q.gl;
    } else {
      raw_to_f32<LO>(q.g, g);
    }
    if (KIND == 2) {
      float av[8], cv[8];
      load8_f32(s.a + (r / s.L) * s.cols + c0, s.cols - c0, av);
      load8_f32(s.c + (r % s.L) * s.cols + c0, s.cols - c0, cv);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = av[j] + cv[j];
    } else {
      raw_to_f32<LO>(q.z, z);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gm = fmaf(z[j], k.sc[j], k.sf[j]) > 0.f ? g[j] : 0.f;
      v[j] = c0 + j < s.cols ? fmaf(k.A[j], gm, -k.E[j]) - k.D[j] * (z[j] - k.mean[j]) : 0.f;
    }
  }
  template <bool LO>
  __device__ __forceinline__ void eval(const Ctx& k, const Raw& q, long long r, int c0, float (&va)[8], float (&vb)[8]) const {
    zero8(va);
    zero8(vb);
    if (c0 >= s.cols) return;
    if (r < s.rows) one<LO>(k, q.a, r, c0, va);
    if (r + 1 < s.rows) one<LO>(k, q.b, r + 1, c0, vb);
  }
};

// ------------------------------------------------------------------------------------------------
// tile emitter: planes [rows][ld] and (optionally) K-blocked transposed planes [ceil(rows/64)][cols][64].
// grid (strips of kStripTiles row tiles, col tiles), 256 threads.  Thread (warp w, lane l) produces rows 2*rp, 2*rp+1
// (rp = 4w + l/8) x 8 columns (l%8) of every 64x64 tile of its strip; phase 2 thread t stores 16 rows of column t/4.
// ------------------------------------------------------------------------------------------------
constexpr int kStripTiles = 8;

template <class P, bool LO>
__global__ void __launch_bounds__(256) emit_tile_kernel(const P prod, long long rows, int cols, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, long long ld, __half* __restrict__ hiT,
                                                        __half* __restrict__ loT) {
  __shared__ uint32_t sh[2][kTileDim * kTilePitch];
  const long long row_tiles = (rows + kTileDim - 1) / kTileDim;
  const long long tile0 = (long long)blockIdx.x * kStripTiles;
  const long long tile_end = tile0 + kStripTiles < row_tiles ? tile0 + kStripTiles : row_tiles;
  const int c0 = blockIdx.y * kTileDim;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int cc = lane & 7, rp = warp * 4 + (lane >> 3);
  const int col = c0 + cc * 8;
  const int ct = t >> 2, part = t & 3;
  const int gc = c0 + ct;
  typename P::Ctx ctx;
  prod.init(col, ctx);
  typename P::Raw cur, nxt;
  prod.template load<LO>(tile0 * kTileDim + 2 * rp, col, cur);
  for (long long tile = tile0; tile < tile_end; ++tile) {
    const long long ra = tile * kTileDim + 2 * rp;
    if (tile + 1 < tile_end) prod.template load<LO>(ra + kTileDim, col, nxt);
    float va[8], vb[8];
    prod.template eval<LO>(ctx, cur, ra, col, va, vb);
    __align__(16) __half ha[8], la[8], hb[8], lb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      split_f16(fmaxf(fminf(va[j], 65504.f), -65504.f), ha[j], la[j]);
      split_f16(fmaxf(fminf(vb[j], 65504.f), -65504.f), hb[j], lb[j]);
    }
    if (col < ld) {
      if (ra < rows) {
        *reinterpret_cast<uint4*>(hi + ra * ld + col) = *reinterpret_cast<const uint4*>(ha);
        if (LO) *reinterpret_cast<uint4*>(lo + ra * ld + col) = *reinterpret_cast<const uint4*>(la);
      }
      if (ra + 1 < rows) {
        *reinterpret_cast<uint4*>(hi + (ra + 1) * ld + col) = *reinterpret_cast<const uint4*>(hb);
        if (LO) *reinterpret_cast<uint4*>(lo + (ra + 1) * ld + col) = *reinterpret_cast<const uint4*>(lb);
      }
    }
    if (hiT != nullptr) {   // uniform
      __syncthreads();      // the previous tile's column reads are done
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int w = (cc * 8 + j) * kTilePitch + rp;
        sh[0][w] = (uint32_t)__half_as_ushort(ha[j]) | ((uint32_t)__half_as_ushort(hb[j]) << 16);
        if (LO) sh[1][w] = (uint32_t)__half_as_ushort(la[j]) | ((uint32_t)__half_as_ushort(lb[j]) << 16);
      }
      __syncthreads();
      if (gc < cols) {
#pragma unroll
        for (int p = 0; p < (LO ? 2 : 1); ++p) {
          __half* dstT = p == 0 ? hiT : loT;
          uint32_t w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) w[k] = sh[p][ct * kTilePitch + part * 8 + k];
          // block `tile` of the K-blocked layout: [cols][64] halves, this thread owns rows part*16..+16 of column gc
          uint4* dst = reinterpret_cast<uint4*>(dstT + (tile * cols + gc) * kTileDim + part * 16);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
    }
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------
// column statistics: out[0][c] += sum_r v, out[1][c] += sum_r v^2   (fp64; BatchNorm1d batch mean / biased variance)
// grid (ceil(cols/256), row slabs), block (32, 8): x -> 8-column chunk, y -> row phase
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_stats_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                        const float* __restrict__ x, long long rows, int cols,
                                                        long long ld, long long rows_per_slab, double* __restrict__ out) {
  // block (TX, TY), TX * TY == 256: TX threads cover TX * 8 consecutive columns of a row (the wider, the longer the
  // contiguous run each row contributes to the DRAM stream), TY row phases
  __shared__ double sh[2][2048];
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int tid = ty * TX + tx;
  const int cpb = TX * 8;
  const int c0 = blockIdx.x * cpb + tx * 8;
  const long long r_begin = (long long)blockIdx.y * rows_per_slab;
  const long long r_end = r_begin + rows_per_slab < rows ? r_begin + rows_per_slab : rows;
  double s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.0;
  if (c0 < cols) {
    for (long long r = r_begin + ty; r < r_end; r += TY) {
      float v[8];
      if (x) load8_f32(x + r * ld + c0, cols - c0, v);
      else load8(hi + r * ld + c0, lo ? lo + r * ld + c0 : nullptr, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double d = (double)v[j];
        s[j] += d;
        ss[j] += d * d;
      }
    }
  }
  for (int i = tid; i < 2 * 2048; i += 256) (&sh[0][0])[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sh[0][tx * 8 + j], s[j]);
    atomicAdd(&sh[1][tx * 8 + j], ss[j]);
  }
  __syncthreads();
  for (int i = tid; i < cpb; i += 256) {
    const int c = blockIdx.x * cpb + i;
    if (c < cols) {
      atomicAdd(out + c, sh[0][i]);
      atomicAdd(out + cols + c, sh[1][i]);
    }
  }
}

// mean / biased variance -> BatchNorm state (+ running statistics update, momentum as torch.nn.BatchNorm1d).
// stats2 != null: statistics of the pair grid z[b,l] = a[b] + c[l]:  mean = mean_a + mean_c, var = var_a + var_c.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const double* __restrict__ stats2,
                                   double count2, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int cols, float* __restrict__ state) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double mean = stats[c] / count;
  double var = stats[cols + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  double n = count;
  if (stats2) {
    const double m2 = stats2[c] / count2;
    double v2 = stats2[cols + c] / count2 - m2 * m2;
    if (v2 < 0.0) v2 = 0.0;
    mean += m2;
    var += v2;
    n = count * count2;
  }
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const double sc = (gamma ? (double)gamma[c] : 1.0) * invstd;
  state[c] = (float)sc;
  state[cols + c] = (float)((beta ? (double)beta[c] : 0.0) - mean * sc);
  state[2 * (long long)cols + c] = (float)mean;
  state[3 * (long long)cols + c] = (float)invstd;
  if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
  if (running_var) {
    const double unbiased = var * (n / (n > 1.0 ? n - 1.0 : 1.0));
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// logits[r] = sum_n relu(z[r][n] * scale[n] + shift[n]) * w[n] + b     (last hidden layer + Linear(H -> 1), one warp per row)
__global__ void __launch_bounds__(256) bn_relu_dot_kernel(const __half* __restrict__ zh, const __half* __restrict__ zl,
                                                          long long rows, int cols, long long ld,
                                                          const float* __restrict__ state, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int chunks = (cols + 7) / 8;
  for (long long r = warp0; r < rows; r += nwarps) {
    float acc = 0.f;
    for (int ch = lane; ch < chunks; ch += 32) {
      const int c0 = ch * 8;
      float v[8], sc[8], sf[8], wv[8];
      load8(zh + r * ld + c0, zl ? zl + r * ld + c0 : nullptr, v);
      if (state) {                  // null state = identity (scale 1, shift 0)
        load8_f32(state + c0, cols - c0, sc);
        load8_f32(state + cols + c0, cols - c0, sf);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sc[j] = 1.f;
          sf[j] = 0.f;
        }
      }
      load8_f32(w + c0, cols - c0, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < cols) acc = fmaf(fmaxf(fmaf(v[j], sc[j], sf[j]), 0.f), wv[j], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = acc + (b ? __ldg(b) : 0.f);
  }
}

// Fused loss + backward seed: the same pass as bn_relu_dot_kernel; with the logit of pair r still in a
// register it evaluates the per-element loss against targets[r] and d loss / d logit, so the [B, L] logits never round-trip
// through autograd and the reference's 4 elementwise loss passes (protnote/utils/losses.py) disappear.
//   kind 1  BCE-with-logits, optional pos_weight per label column   (losses.py:270-272, torch.nn.BCEWithLogitsLoss):
//             l = (1 - t) x + (1 + (pw - 1) t) softplus(-x)
//   kind 2  FocalLoss(alpha, gamma, label_smoothing)                (losses.py:171-213):
//             t' = t (1 - s) + (1 - t) s;  b = BCE(x, t');  pt = exp(-b);  l = alpha_t (1 - pt)^gamma b,
//             alpha_t = alpha t' + (1 - alpha)(1 - t')  when alpha >= 0
// g[r] = grad_scale * dl/dx (grad_scale = 1 / (B * L_total) for reduction 'mean'), loss_sum += sum_r l  (fp64).
struct LossSpec {
  int kind;                  // 0 none, 1 BCE, 2 focal
  float gamma, alpha, label_smoothing, grad_scale;
  const float* targets;      // [rows], same order as the logits (row = b * L + l)
  const float* pos_weight;   // [L] or nullptr (BCE only)
  long long L;               // label rows per protein on this rank (indexes pos_weight)
  float* g_out;              // [rows]
  double* loss_sum;          // [1], accumulated
};

__device__ __forceinline__ void loss_and_seed(const LossSpec& ls, float x, long long r, float& loss, float& g) {
  const float t = ls.targets[r];
  const float sp = log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f);      // softplus(-x) = -log(sigmoid(x))
  const float sig = 1.f / (1.f + expf(-x));
  if (ls.kind == 1) {
    const float pw = ls.pos_weight ? ls.pos_weight[r % ls.L] : 1.f;
    const float cw = 1.f + (pw - 1.f) * t;
    loss = (1.f - t) * x + cw * sp;
    g = (1.f - t) - cw * (1.f - sig);
  } else {
    const float s = ls.label_smoothing;
    const float ts = s > 0.f ? t * (1.f - s) + (1.f - t) * s : t;
    const float b = (1.f - ts) * x + sp;
    const float pt = expf(-b);
    const float om = 1.f - pt;
    const float mod = powf(om, ls.gamma);
    const float at = ls.alpha >= 0.f ? ls.alpha * ts + (1.f - ls.alpha) * (1.f - ts) : 1.f;
    loss = at * mod * b;
    const float modm1 = ls.gamma == 1.f ? 1.f : (om > 0.f ? powf(om, ls.gamma - 1.f) : 0.f);
    g = at * (sig - ts) * (ls.gamma * modm1 * pt * b + mod);
  }
}

__global__ void __launch_bounds__(256) bn_relu_dot_loss_kernel(const __half* __restrict__ zh, const __half* __restrict__ zl,
                                                               long long rows, int cols, long long ld,
                                                               const float* __restrict__ state, const float* __restrict__ w,
                                                               const float* __restrict__ b, float* __restrict__ out,
                                                               const LossSpec ls) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int chunks = (cols + 7) / 8;
  double local = 0.0;
  for (long long r = warp0; r < rows; r += nwarps) {
    float acc = 0.f;
    for (int ch = lane; ch < chunks; ch += 32) {
      const int c0 = ch * 8;
      float v[8], sc[8], sf[8], wv[8];
      load8(zh + r * ld + c0, zl ? zl + r * ld + c0 : nullptr, v);
      load8_f32(state + c0, cols - c0, sc);
      load8_f32(state + cols + c0, cols - c0, sf);
      load8_f32(w + c0, cols - c0, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < cols) acc = fmaf(fmaxf(fmaf(v[j], sc[j], sf[j]), 0.f), wv[j], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float x = acc + (b ? __ldg(b) : 0.f);
      if (out) out[r] = x;
      float l, g;
      loss_and_seed(ls, x, r, l, g);
      ls.g_out[r] = g * ls.grad_scale;
      local += (double)l;
    }
  }
  if (lane == 0 && local != 0.0) atomicAdd(ls.loss_sum, local);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm + ReLU backward, pass 1: sums[0][c] = sum_r g_y, sums[1][c] = sum_r g_y * xhat (true scale, fp64),
// maxes = (max |g_y|, max |xhat|); kind 1 also dw[c] = sum_r g_logit[r] * relu(pre)[r][c] and db = sum_r g_logit[r].
// A thread walks its rows two at a time (both rows' loads in flight), keeps fp32 partial sums over 16 rows and flushes
// them into fp64 accumulators in shared memory, so few registers are live and many blocks fit an SM.
// ------------------------------------------------------------------------------------------------
// max |x| over the 8 halves of a raw 16-byte piece, as packed half2 operations (1 instruction per 2 elements)
__device__ __forceinline__ void absmax_raw(const uint4& h, __half2& m) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&h);
#pragma unroll
  for (int j = 0; j < 4; ++j) m = __hmax2(m, __habs2(h2[j]));
}

// The pass is bound by instruction issue, not by HBM (about 25 instructions per element in its first version), so the
// inner loop does the minimum per element: mask, sum g, sum g*z.  Everything that is per column or per tensor is
// applied once at the end:  s1 = inv_gsc * S_g,  s2 = inv_gsc * invstd * (S_gz - mean * S_g)  (fp64), and the two
// magnitudes the scale bound needs come from packed-half maxima of the raw planes:
//   maxes[0] = max |g| (before the ReLU mask, a bound of max |g_y|),  maxes[1] = max |z| (bwd_scale turns it into a
//   bound of max |xhat|).
template <int KIND, bool LO>
__global__ void __launch_bounds__(256) bwd_stats_kernel(const BwdSrc s, long long per_slab, double* __restrict__ sums,
                                                        unsigned* __restrict__ maxes, double* __restrict__ dw,
                                                        double* __restrict__ db, float* __restrict__ gyl) {
  // fp64 accumulators: one private slot per thread and column (no atomics while streaming), reduced over the 8 row
  // phases at the end.  [sum][ty][256 columns] doubles = 32 KB (48 KB for kind 1).
  __shared__ double sh[KIND == 1 ? 3 : 2][2048];
  // block (TX, TY), TX * TY == 256: TX threads cover TX * 8 consecutive columns of a row, TY row phases
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int cpb = TX * 8;
  // this thread's 8 private fp64 slots of every sum live at (ty * 8 + j) * TX + tx: consecutive lanes -> consecutive
  // doubles (the first layout, 8 consecutive doubles per thread, was a 16-way bank conflict: 1.1e9 conflicts per launch)
  const int slot0 = ty * 8 * TX + tx;
  const int c0 = blockIdx.x * cpb + tx * 8;
  // kind 0/1: a slab is a range of rows.  kind 2 (rows = b * L + l): a slab is a range of LABELS walked for every
  // protein in turn, so the slab's rows of c (per_slab x 256 columns, fp32) are re-read from L1/L2, not from HBM.
  const long long n_outer = KIND == 2 ? s.rows / s.L : 1;
  const long long extent = KIND == 2 ? s.L : s.rows;
  const long long i_begin = (long long)blockIdx.y * per_slab;
  const long long i_end = i_begin + per_slab < extent ? i_begin + per_slab : extent;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[0][slot0 + j * TX] = 0.0;
    sh[1][slot0 + j * TX] = 0.0;
    if (KIND == 1) sh[KIND == 1 ? 2 : 0][slot0 + j * TX] = 0.0;
  }
  float p1[8], p2[8], p3[8];
  zero8(p1);
  zero8(p2);
  zero8(p3);
  __half2 gm2 = __float2half2_rn(0.f), zm2 = __float2half2_rn(0.f);
  float glmax = 0.f, zmaxf = 0.f;
  double dbs = 0.0;
  auto flush = [&]() {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sh[0][slot0 + j * TX] += (double)p1[j];
      sh[1][slot0 + j * TX] += (double)p2[j];
      if (KIND == 1) sh[KIND == 1 ? 2 : 0][slot0 + j * TX] += (double)p3[j];
      p1[j] = p2[j] = p3[j] = 0.f;
    }
  };
  float wmax = 0.f;
  const bool full_chunk = c0 + 8 <= s.cols;
  if (c0 < s.cols && i_begin + ty < i_end) {
    float sc[8], sf[8], wv[8];
    load8_f32(s.state + c0, s.cols - c0, sc);
    load8_f32(s.state + s.cols + c0, s.cols - c0, sf);
    if (KIND == 1) {
      load8_f32(s.w + c0, s.cols - c0, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) wmax = fmaxf(wmax, fabsf(wv[j]));
    } else {
      zero8(wv);
    }
    // kind 2 by-product: gyl[l][n] = sum_b g_y[b,l,n] for the labels this thread owns - with it dc[l] = sum_b g_z[b,l]
    // follows analytically (pair_dc_fixup_kernel) and the separate pass over g_h1 for dc is not needed
    float creg[KIND == 2 ? kRif : 1][8], gacc[KIND == 2 ? kRif : 1][8];
    if (KIND == 2) {
#pragma unroll
      for (int k = 0; k < kRif; ++k) {
        zero8(creg[k]);
        zero8(gacc[k]);
        if (i_begin + ty + TY * k < i_end) load8_f32(s.c + (i_begin + ty + TY * k) * s.cols + c0, s.cols - c0, creg[k]);
      }
    }
    // columns beyond `cols` (only in the last chunk of a row): scale = shift = 0 -> mask false -> no contribution
    // software pipeline over (protein, row) steps of kRif rows: the next step's loads are issued before this step's math
    long long bb = 0, i = i_begin + ty;
    BwdRaw cur[kRif], nxt[kRif];
#pragma unroll
    for (int k = 0; k < kRif; ++k)
      if (i + TY * k < i_end) bwd_raw_load<KIND, LO>(s, i + TY * k, c0, cur[k]);
    int it = 0;
    bool have = true;
    while (have) {
      long long nbb = bb, ni = i + TY * kRif;
      if (ni >= i_end) {
        ni = i_begin + ty;
        ++nbb;
      }
      const bool have_next = nbb < n_outer;
      if (have_next) {
#pragma unroll
        for (int k = 0; k < kRif; ++k)
          if (ni + TY * k < i_end) bwd_raw_load<KIND, LO>(s, nbb * extent + ni + TY * k, c0, nxt[k]);
      }
#pragma unroll
      for (int k = 0; k < kRif; ++k) {
        if (i + TY * k >= i_end) continue;
        float g[8], z[8];
        if (KIND == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = cur[k].gl * wv[j];
          glmax = fmaxf(glmax, fabsf(cur[k].gl));
          if (blockIdx.x == 0 && tx == 0) dbs += (double)cur[k].gl;
        } else {
          raw_to_f32<LO>(cur[k].g, g);
          if (full_chunk) {
            absmax_raw(cur[k].g.h, gm2);
          } else {       // the columns beyond `cols` of a partial chunk hold whatever the allocation held
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + j < s.cols) glmax = fmaxf(glmax, fabsf(g[j]));
          }
        }
        if (KIND == 2) {
          // row r = bb * L + l.  The slab is exactly one step of labels per thread (host: per_slab = TY * kRif), so the
          // thread's label halves c[l] sit in registers for the whole walk over the proteins; a[bb] is an L1 hit
          float av[8];
          load8_f32(s.a + bb * s.cols + c0, s.cols - c0, av);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            z[j] = av[j] + creg[k][j];
            zmaxf = fmaxf(zmaxf, fabsf(z[j]));
          }
        } else {
          raw_to_f32<LO>(cur[k].z, z);
          if (full_chunk) {
            absmax_raw(cur[k].z.h, zm2);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + j < s.cols) zmaxf = fmaxf(zmaxf, fabsf(z[j]));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pre = fmaf(z[j], sc[j], sf[j]);
          const float gmk = pre > 0.f ? g[j] : 0.f;
          p1[j] += gmk;
          p2[j] = fmaf(gmk, z[j], p2[j]);
          if (KIND == 2) gacc[k][j] += gmk;
          if (KIND == 1) p3[j] = fmaf(cur[k].gl, fmaxf(pre, 0.f), p3[j]);
        }
      }
      if ((++it & 7) == 0) flush();   // fp32 partial sums cover at most 8 * kRif rows
#pragma unroll
      for (int k = 0; k < kRif; ++k) cur[k] = nxt[k];
      bb = nbb;
      i = ni;
      have = have_next;
    }
    flush();
    if (KIND == 2 && gyl) {
      const float inv = s.g_sc ? 1.f / __ldg(s.g_sc) : 1.f;
#pragma unroll
      for (int k = 0; k < kRif; ++k) {
        const long long l = i_begin + ty + TY * k;
        if (l >= i_end) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < s.cols) gyl[l * s.cols + c0 + j] = gacc[k][j] * inv;
      }
    }
  }
  __syncthreads();
  const float inv_gsc = (KIND != 1 && s.g_sc) ? 1.f / __ldg(s.g_sc) : 1.f;
  for (int i = ty * TX + tx; i < cpb; i += 256) {
    const int c = blockIdx.x * cpb + i;
    if (c >= s.cols) continue;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int y = 0; y < TY; ++y) {
      const int sl = (y * 8 + (i & 7)) * TX + (i >> 3);      // column i of the block = thread i / 8, element i % 8
      t0 += sh[0][sl];
      t1 += sh[1][sl];
      if (KIND == 1) t2 += sh[KIND == 1 ? 2 : 0][sl];
    }
    const double mean = (double)s.state[2 * (long long)s.cols + c], invstd = (double)s.state[3 * (long long)s.cols + c];
    atomicAdd(sums + c, t0 * (double)inv_gsc);
    atomicAdd(sums + s.cols + c, (t1 - mean * t0) * invstd * (double)inv_gsc);
    if (KIND == 1 && dw) atomicAdd(dw + c, t2);
  }
  float gmax, zmax;
  if (KIND == 1) gmax = glmax * wmax;
  else gmax = fmaxf(glmax, fmaxf(__low2float(gm2), __high2float(gm2))) * inv_gsc;
  zmax = fmaxf(zmaxf, fmaxf(__low2float(zm2), __high2float(zm2)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  if ((tx & 31) == 0) {   // lane 0 of every warp
    atomicMax(maxes, __float_as_uint(gmax));
    atomicMax(maxes + 1, __float_as_uint(zmax));
    if (KIND == 1 && blockIdx.x == 0 && tx == 0 && db) atomicAdd(db, dbs);
  }
}

// power of two that puts `bound` into [2^5, 2^6): fp16 planes of a gradient tensor keep ~22 bits for every element
// within 2^-9 of the largest one and cannot overflow through one more dgrad
__device__ __forceinline__ float grad_scale_from_bound(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
  int e;
  frexpf(bound, &e);
  return ldexpf(1.f, 6 - e);
}

// means[0][c] = sums[0][c] / n, means[1][c] = sums[1][c] / n (fp32, what pass 2 subtracts) and, when sc_out is given,
// the scale for the g_z tensor pass 2 writes, from an upper bound of |g_z|:
//   max|scale| * (max|g| + max|s1|/n + bound(|xhat|) * max|s2|/n)
__global__ void bwd_scale_kernel(const double* __restrict__ sums, const unsigned* __restrict__ maxes,
                                 const float* __restrict__ state, double count, int cols, float* __restrict__ sc_out,
                                 float* __restrict__ means) {
  __shared__ float red[5][32];
  float ms = 0.f, m1 = 0.f, m2 = 0.f, mi = 0.f, mm = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float a1 = (float)(sums[c] / count), a2 = (float)(sums[cols + c] / count);
    means[c] = a1;
    means[cols + c] = a2;
    ms = fmaxf(ms, fabsf(state[c]));
    m1 = fmaxf(m1, fabsf(a1));
    m2 = fmaxf(m2, fabsf(a2));
    mm = fmaxf(mm, fabsf(state[2 * (long long)cols + c]));
    mi = fmaxf(mi, fabsf(state[3 * (long long)cols + c]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, o));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = ms;
    red[1][threadIdx.x >> 5] = m1;
    red[2][threadIdx.x >> 5] = m2;
    red[3][threadIdx.x >> 5] = mi;
    red[4][threadIdx.x >> 5] = mm;
  }
  __syncthreads();
  if (threadIdx.x == 0 && sc_out) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      ms = fmaxf(ms, red[0][w]);
      m1 = fmaxf(m1, red[1][w]);
      m2 = fmaxf(m2, red[2][w]);
      mi = fmaxf(mi, red[3][w]);
      mm = fmaxf(mm, red[4][w]);
    }
    // maxes[0] = max |g| (bounds |g_y|), maxes[1] = max |z|  ->  |xhat| <= max invstd * (max |z| + max |mean|)
    const float gmax = __uint_as_float(maxes[0]), xmax = mi * (__uint_as_float(maxes[1]) + mm);
    *sc_out = grad_scale_from_bound(ms * (gmax + m1 + xmax * m2));
  }
}

// sc[0] = power-of-two scale from the absmax bits left in sc[1] by absmax_kernel
__global__ void autoscale_finish_kernel(float* __restrict__ sc) {
  sc[0] = grad_scale_from_bound(__uint_as_float(reinterpret_cast<const unsigned*>(sc)[1]));
}

// out[n] = 1 / (s0 * s1 * s2)   (null -> 1): the epilogue scale that undoes the power-of-two operand scales
__global__ void scale_vector_kernel(float* __restrict__ out, int n, const float* __restrict__ s0,
                                    const float* __restrict__ s1, const float* __restrict__ s2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = 1.f;
  if (s0) d *= *s0;
  if (s1) d *= *s1;
  if (s2) d *= *s2;
  out[i] = 1.f / d;
}

// ------------------------------------------------------------------------------------------------
// layer 1 backward (kind 2): the gradient w.r.t. z1[b,l] = a[b] + c[l] is never stored, only its two marginals:
//   dc[l][n] = sum_b g_z1[b,l,n]   one thread per (l, 8 columns) walks the B proteins (rows L apart), plain store
//   da[b][n] = sum_l g_z1[b,l,n]   a column sum over the L rows of protein b (same structure as pass 1), fp64 atomics
// Reading g_h1 twice at streaming speed is cheaper than synchronising a block once per protein.
// ------------------------------------------------------------------------------------------------
// g_z of row (protein bb, label l); no row-index division
template <bool LO>
__device__ __forceinline__ void pair_gz(const BwdSrc& s, const BwdRaw& q, long long bb, long long l, int c0, const BnVec& b,
                                        const float (&m1)[8], const float (&m2)[8], float inv_gsc, float (&gz)[8]) {
  float g[8], av[8], cv[8];
  raw_to_f32<LO>(q.g, g);
  load8_f32(s.a + bb * s.cols + c0, s.cols - c0, av);
  load8_f32(s.c + l * s.cols + c0, s.cols - c0, cv);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float z = av[j] + cv[j];
    const float gy = fmaf(z, b.scale[j], b.shift[j]) > 0.f ? g[j] * inv_gsc : 0.f;
    const float xh = (z - b.mean[j]) * b.invstd[j];
    gz[j] = c0 + j < s.cols ? b.scale[j] * (gy - m1[j] - xh * m2[j]) : 0.f;
  }
}

// dc[l][n] = sum_b g_z1[b,l,n] from the per-label masked sums gyl = sum_b g_y (by-product of the statistics pass):
//   sum_b g_z = scale * (gyl - B m1 - m2 * sum_b xhat),   sum_b xhat = invstd * (A + B c[l] - B mean),  A = sum_b a[b]
__global__ void pair_dc_fixup_kernel(const float* __restrict__ gyl, const float* __restrict__ c,
                                     const double* __restrict__ a_stats, const float* __restrict__ state,
                                     const float* __restrict__ means, long long B, long long L, int cols,
                                     float* __restrict__ dc) {
  const long long total = L * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % cols);
    const float scale = state[n], mean = state[2 * (long long)cols + n], invstd = state[3 * (long long)cols + n];
    const float m1 = means[n], m2 = means[cols + n];
    const float sumx = invstd * ((float)a_stats[n] + (float)B * (c[i] - mean));
    dc[i] = scale * (gyl[i] - (float)B * m1 - m2 * sumx);
  }
}

// grid (ceil(cols/256), ceil(L/8)), block (32, 8)
template <bool LO>
__global__ void __launch_bounds__(256) pair_dc_kernel(const BwdSrc s, const float* __restrict__ means, long long B,
                                                      float* __restrict__ dc) {
  const int c0 = blockIdx.x * 256 + threadIdx.x * 8;
  const long long l = (long long)blockIdx.y * 8 + threadIdx.y;
  if (c0 >= s.cols || l >= s.L) return;
  BnVec b;
  load_bn(s.state, s.cols, c0, b);
  float m1[8], m2[8], acc[8];
  load8_f32(means + c0, s.cols - c0, m1);
  load8_f32(means + s.cols + c0, s.cols - c0, m2);
  zero8(acc);
  const float inv_gsc = s.g_sc ? 1.f / __ldg(s.g_sc) : 1.f;
  BwdRaw cur[kRif], nxt[kRif];
#pragma unroll
  for (int k = 0; k < kRif; ++k)
    if (k < B) bwd_raw_load<2, LO>(s, k * s.L + l, c0, cur[k]);
  for (long long bb = 0; bb < B; bb += kRif) {
#pragma unroll
    for (int k = 0; k < kRif; ++k)
      if (bb + kRif + k < B) bwd_raw_load<2, LO>(s, (bb + kRif + k) * s.L + l, c0, nxt[k]);
#pragma unroll
    for (int k = 0; k < kRif; ++k) {
      if (bb + k >= B) continue;
      float gz[8];
      pair_gz<LO>(s, cur[k], bb + k, l, c0, b, m1, m2, inv_gsc, gz);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += gz[j];
    }
#pragma unroll
    for (int k = 0; k < kRif; ++k) cur[k] = nxt[k];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (c0 + j < s.cols) dc[l * s.cols + c0 + j] = acc[j];
}

// grid (B, ceil(cols/256), label slabs), block (32, 8): one protein and one slab of labels per block, walked two rows at
// a time with the next step's loads in flight; a single block reduction and 256 fp64 atomics at the end.  The protein
// index is the FASTEST grid dimension, so the B blocks that share a slab's rows of c run together and find them in L2.
template <bool LO>
__global__ void __launch_bounds__(256) pair_da_kernel(const BwdSrc s, const float* __restrict__ means, long long B,
                                                      long long labels_per_slab, double* __restrict__ da) {
  __shared__ float sh[8][8][32];   // [row phase][element][lane]: conflict-free
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long bb = blockIdx.x;
  const int c0 = blockIdx.y * 256 + tx * 8;
  const long long l_begin = (long long)blockIdx.z * labels_per_slab;
  const long long l_end = l_begin + labels_per_slab < s.L ? l_begin + labels_per_slab : s.L;
  float acc[8];
  zero8(acc);
  if (c0 < s.cols && l_begin + ty < l_end) {
    BnVec b;
    float m1[8], m2[8];
    load_bn(s.state, s.cols, c0, b);
    load8_f32(means + c0, s.cols - c0, m1);
    load8_f32(means + s.cols + c0, s.cols - c0, m2);
    const float inv_gsc = s.g_sc ? 1.f / __ldg(s.g_sc) : 1.f;
    BwdRaw cur[kRif], nxt[kRif];
    long long l = l_begin + ty;
#pragma unroll
    for (int k = 0; k < kRif; ++k)
      if (l + 8 * k < l_end) bwd_raw_load<2, LO>(s, bb * s.L + l + 8 * k, c0, cur[k]);
    for (; l < l_end; l += 8 * kRif) {
      const long long nl = l + 8 * kRif;
#pragma unroll
      for (int k = 0; k < kRif; ++k)
        if (nl + 8 * k < l_end) bwd_raw_load<2, LO>(s, bb * s.L + nl + 8 * k, c0, nxt[k]);
#pragma unroll
      for (int k = 0; k < kRif; ++k) {
        if (l + 8 * k >= l_end) continue;
        float gz[8];
        pair_gz<LO>(s, cur[k], bb, l + 8 * k, c0, b, m1, m2, inv_gsc, gz);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += gz[j];
      }
#pragma unroll
      for (int k = 0; k < kRif; ++k) cur[k] = nxt[k];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[ty][j][tx] = acc[j];
  __syncthreads();
  const int i = ty * 32 + tx;            // column i of the block = lane i / 8, element i % 8
  const int c = blockIdx.y * 256 + i;
  if (c < s.cols) {
    float tot = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) tot += sh[y][i & 7][i >> 3];
    atomicAdd(da + bb * s.cols + c, (double)tot);
  }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}

}  // namespace pn
