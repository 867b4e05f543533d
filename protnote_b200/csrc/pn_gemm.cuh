// The one tensor-core engine every dense contraction of the scoring path runs on (sm_100a only).
//
//   D[M,N] = sum over K of A[M,K] * B[N,K]        (both operands K-major, fp32 accumulate)
//
// Operands are fp32 values carried as TWO fp16 planes (hi = fp16(x), lo = fp16(x - hi)).  In `strict`
// mode (NPASS == 3) every k-step issues three tcgen05.mma into the same accumulator,
//   A_lo*B_hi + A_hi*B_lo + A_hi*B_hi,
// which reproduces an fp32 product sum to ~2^-22 relative (the dropped lo*lo term); `fast` mode (NPASS == 1)
// issues A_hi*B_hi only.  Weights are pre-scaled by a power of two at pack time so that their lo plane stays in
// fp16's normal range; the epilogue's per-column `scale` undoes it (and carries the folded BatchNorm).
//
// Accumulator promotion.  The tcgen05 accumulator add rounds toward zero, so a long accumulation chain picks up a
// bias proportional to its length (measured on B200: K=3072 strict, relative error 3e-5; K=9900, 9e-5 - see
// tools/accum_probe.py and profiles/).  The K loop is therefore cut into chunks of `chunk_kblocks` k-blocks
// (256 K-elements in strict mode); each chunk accumulates from zero into one of the two TMEM buffers and the
// epilogue warps add it into fp32 running sums held in registers (round-to-nearest).  With 256-element chunks the
// result is as accurate as an fp32 SIMT GEMM (rms 6e-7 vs 5e-7 for cuBLAS fp32 at K=3072).  The two TMEM buffers
// alternate, so the tensor core never waits for a drain; in fast mode a chunk is the whole K loop and the scheme
// degenerates to the usual "epilogue of tile i overlaps main loop of tile i+1".
//
// Structure: persistent CTAs (grid = #SMs), 12 warps, warp-specialised
//   warp 0     TMA producer   (cp.async.bulk.tensor 2D for plain matrices, 3D for dilated-conv taps)
//   warp 1     MMA issuer     (one elected lane, tcgen05.mma cta_group::1, M=128, N=bn<=256, K=16)
//   warp 2     TMEM allocator (512 columns = two accumulator buffers)
//   warps 4-11 epilogue       (tcgen05.ld 32x32b; warp w owns TMEM lanes 32*(w%4).. and column half (w-4)/4)
// Registers are re-balanced with setmaxnreg: the four control warps give theirs to the epilogue warps, whose
// running sums (128 fp32 per thread) live in registers.
//
// A-operand addressing modes
//   plain : A is a [M][K] matrix, tile rows = 128 consecutive rows.
//   conv  : A is a channels-last activation tensor [B][T][C]; an M tile is 128 consecutive positions of ONE
//           sequence, k-block kb reads tap = kb / cblocks at T-offset (tap - taps/2) * dilation.  TMA zero-fills
//           out-of-range positions, which is exactly Conv1d(padding="same") - reference
//           protnote/models/protein_encoders.py:8-17,39-46 - and positions >= length are kept at zero by the
//           producing epilogue (set_padding_to_sentinel, protnote/data/datasets.py:535-569).
#pragma once

#include "pn_ptx.cuh"

namespace pn {

constexpr int kBM = 128;          // accumulator rows per tile (UMMA M)
constexpr int kMaxBN = 256;       // accumulator columns per tile (UMMA N), runtime bn <= kMaxBN
constexpr int kGemmThreads = 384; // 4 control warps + 8 epilogue warps
constexpr int kSmemBudget = 200 * 1024;
constexpr int kCtrlRegs = 24;     // setmaxnreg: 128*24 + 256*240 = 64512 = the 384*168 registers the CTA owns at launch
constexpr int kEpiRegs = 240;
// fused pair-feature variant (GEN): 4 more warps generate the A operand; 512 threads x 128 registers at launch
constexpr int kGenThreads = 512;
constexpr int kGenWarpRegs = 56;   // 128*24 + 128*56 + 256*216 = 65536
constexpr int kGenEpiRegs = 216;

struct alignas(64) GemmParams {
  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  CUtensorMap tm_gen_c;  // GEN: fp32 [L][H] label halves, box 32 x 128, SWIZZLE_128B
  int M, N;              // logical extents of D (conv: M = B*T rows addressed as (b,t))
  int bn;                // tile width, multiple of 32, <= 256
  int tiles_m, tiles_n;
  int num_kblocks;       // K / BK (conv: taps * cblocks)
  int chunk_kblocks;     // k-blocks per accumulator chunk (promotion period), >= 1
  // conv addressing (conv_taps == 0 -> plain)
  int conv_taps, conv_cblocks, conv_cpad, conv_dil, conv_T, conv_tiles_per_seq;
  // K-blocked operands (wgrad): the operand is stored as [K/64][rows][64] so that the 128-byte row pieces of one k-block
  // are contiguous; its tensor map is 3-D (k within block, row, block) and the K coordinate is split accordingly
  int a_kblk, b_kblk;
  // strict mode, split accumulators: the large hi*hi products of a chunk go to TMEM buffer 0 and are promoted to fp32
  // registers chunk by chunk, the small lo*hi + hi*lo corrections of the WHOLE tile accumulate in TMEM buffer 1 (their
  // truncation error is 2^-11 of the main term's) and are added once at the end of the tile.  The promotion of buffer 0
  // overlaps the correction MMAs of the same chunk, so the tensor core never waits for a drain.  Needs
  // chunk_kblocks < pipeline stages (the chunk's operand stages stay resident until its corrections are issued).
  int split_corr;
  // strict mode: the tcgen05 accumulator add rounds toward zero, so a chunk's sum comes out short by a factor that is
  // proportional to the number of truncating adds of the chunk (measured: profiles/r02_trunc_comp_probe.txt).  The
  // promoted total is multiplied by (1 + trunc_comp) before the epilogue, which removes the mean of that bias; what is
  // left is zero-mean and uncorrelated between outputs, like round-to-nearest noise.  0 = off.
  float trunc_comp;
  int l2_hints;                  // pair kernel: 0 none, 1 weights evict_last, 2 weights evict_last + activations evict_first
  const long long* lengths;      // [B] valid positions per sequence (conv) or nullptr
  // pair-row addressing for the additive row terms: row r -> (r / pair_nl, r % pair_nl)
  int pair_nl;
  const float* add_p; long long ld_add_p;   // [B][ld] added per protein   (nullable)
  const float* add_l; long long ld_add_l;   // [L][ld] added per label row (nullable)
  // epilogue: y = acc*scale[n] + shift[n] (+ add rows) (+ resid); masked rows -> 0
  const float* scale; const float* shift;   // [N] (nullable -> 1 / 0)
  const float* resid; long long ld_resid;   // fp32 [M][ld] (nullable)
  float* out_f32; long long ld_out;         // y stored here (nullable)
  float* out_z; long long ld_z;             // z stored here as fp32 (nullable)
  // z = relu?(y*scale2 + shift2) ; masked rows -> 0
  const float* scale2; const float* shift2; // [N] (nullable)
  int relu;
  __half* out_hi; __half* out_lo; long long ld_split;   // z as fp16 planes (nullable; out_lo nullable)
  const float* dot_w; float* dot_out;       // float2 (hi, lo) at [(row*tiles_n + n_tile)*2 + half] = sum_n z*dot_w[n] (nullable)
  int vec_z;
  int vec_out, vec_resid, vec_split;        // 16-byte vector access is legal for that tensor (host-checked)
  // GEN kernels only: A[r][k] = relu(gen_a[r / pair_nl][k] + gen_c[r % pair_nl][k]) is built in shared memory by the
  // generator warps instead of being read through tm_a_* (layer 1 of the pair scorer, ProtNote.py:112-126,293)
  const float* gen_a; long long ld_gen_a;   // protein halves [B][ld]; the label halves come through tm_gen_c
  double timed_flops;                       // host-side bookkeeping only (algorithmic FLOPs of this launch)
};

// Per-tile copies of the per-column epilogue vectors in shared memory (all 128 rows of a tile use the same
// values: a broadcast LDS.128 fetches four columns at once instead of four global loads per thread).
struct EpiConsts {
  float scale[kMaxBN], shift[kMaxBN], scale2[kMaxBN], shift2[kMaxBN], dotw[kMaxBN];
};
constexpr int kStageRowBytes = 80;                       // 64 B of halves + 16 B pad (conflict-free 16 B accesses)
constexpr int kStageBytesPerWarp = 32 * kStageRowBytes;  // one 32x32 fp16 block per epilogue warp
constexpr int kEpiSmemBytes = (int)sizeof(EpiConsts) + 8 * kStageBytesPerWarp;

constexpr int kGenStages = 3;                       // GEN: depth of the fp32 label-half staging ring
constexpr int kGenTileBytes = kBM * 32 * 4;         // GEN: one 128-row x 32-float tile of c (SWIZZLE_128B)

template <int BK, int NPASS, bool GEN = false>
struct GemmCfg {
  static constexpr int kSwizzle = BK * 2;                    // bytes per smem row
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kATile = kBM * BK * 2;
  static constexpr int kBTile = kMaxBN * BK * 2;
  static constexpr int kStageBytes = kPlanes * (kATile + kBTile);
  static constexpr int kGenBytes = GEN ? kGenStages * kGenTileBytes : 0;
  static constexpr int kStages = (kSmemBudget - kGenBytes) / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kGenBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiSmemBytes;
  static_assert(kStages >= 2, "pipeline too shallow");
};

__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

template <int REGS>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}

// fp32 pair -> packed fp16 hi pair and lo pair (3 instructions per element)
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fmaxf(fminf(a, 65504.f), -65504.f);
  b = fmaxf(fminf(b, 65504.f), -65504.f);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// same for values known to be >= 0 (only the overflow clamp is needed)
__device__ __forceinline__ void split_pack_pos(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fminf(a, 65504.f);
  b = fminf(b, 65504.f);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// A warp's 32 rows x 32 halves: transpose through shared memory so that four consecutive lanes write 64
// contiguous bytes of one row (8 rows per store instruction instead of 32 scattered 16-byte pieces).
__device__ __forceinline__ void store_halves_coalesced(const uint32_t (&v)[16], uint8_t* stage, __half* row0_ptr,
                                                       long long ld, uint32_t row_mask, int lane) {
  uint4* mine = reinterpret_cast<uint4*>(stage + lane * kStageRowBytes);
#pragma unroll
  for (int j = 0; j < 4; ++j) mine[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + (lane >> 2);
    const int ch = lane & 3;
    const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * kStageRowBytes + ch * 16);
    if ((row_mask >> rr) & 1u) *reinterpret_cast<uint4*>(row0_ptr + rr * ld + ch * 8) = val;
  }
  __syncwarp();
}

// A warp's 32 rows x 32 fp32 columns <-> global memory, 16 columns (64 bytes of every row) at a time through the same
// staging block: four consecutive lanes move 64 contiguous bytes of one row, so one warp instruction touches 8 rows
// instead of 32 (a row-per-thread float4 access is 32 separate 16-byte pieces: 4x the L1 tag cycles, and on the store
// side 32 partial sectors).  `v` is this thread's row (lane = row); the load ADDS the block to it.
__device__ __forceinline__ void add_f32_coalesced(float (&v)[32], uint8_t* stage, const float* row0_ptr, long long ld,
                                                   uint32_t row_mask, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + (lane >> 2);
      const int ch = lane & 3;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if ((row_mask >> rr) & 1u) val = *reinterpret_cast<const uint4*>(row0_ptr + rr * ld + h * 16 + ch * 4);
      *reinterpret_cast<uint4*>(stage + rr * kStageRowBytes + ch * 16) = val;
    }
    __syncwarp();
    const uint4* mine = reinterpret_cast<const uint4*>(stage + lane * kStageRowBytes);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 q = mine[j];
      v[h * 16 + 4 * j] += __uint_as_float(q.x);
      v[h * 16 + 4 * j + 1] += __uint_as_float(q.y);
      v[h * 16 + 4 * j + 2] += __uint_as_float(q.z);
      v[h * 16 + 4 * j + 3] += __uint_as_float(q.w);
    }
    __syncwarp();
  }
}

// Residual blocks fetched ahead with cp.async (pair kernel): the 32 x 32 fp32 block of the NEXT group goes straight from
// global to a per-warp shared-memory buffer (rows of 128 + 16 pad bytes, so row-per-thread LDS.128 is conflict-free)
// while the current group is being finished; the first group's block is requested before the accumulator is even
// complete.  No registers are held across the wait (a register prefetch spilled: profiles/r02_ab_prefetch_pairfeatures.txt),
// and the DRAM latency of the fp32 residual stream - eight exposed round trips per tile for the encoder's 1x1
// convolutions - disappears behind the main loop.
constexpr int kResidRowBytes = 144;
constexpr int kResidBlockBytes = 32 * kResidRowBytes;     // per epilogue warp

__device__ __forceinline__ void resid_cp_async(uint8_t* buf, const float* row0_ptr, long long ld, uint32_t row_mask, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (lane >> 3);
    const int ch = lane & 7;
    if ((row_mask >> rr) & 1u)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + rr * kResidRowBytes + ch * 16)),
                   "l"(row0_ptr + rr * ld + ch * 4)
                   : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void add_resid_from_smem(float (&y)[32], const uint8_t* buf, int lane) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();                      // every lane's copies have landed and are visible to the warp
  const uint4* mine = reinterpret_cast<const uint4*>(buf + lane * kResidRowBytes);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 q = mine[j];
    y[4 * j] += __uint_as_float(q.x);
    y[4 * j + 1] += __uint_as_float(q.y);
    y[4 * j + 2] += __uint_as_float(q.z);
    y[4 * j + 3] += __uint_as_float(q.w);
  }
  __syncwarp();                      // the buffer may be refilled
}

__device__ __forceinline__ void store_f32_coalesced(const float (&v)[32], uint8_t* stage, float* row0_ptr, long long ld,
                                                    uint32_t row_mask, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4* mine = reinterpret_cast<uint4*>(stage + lane * kStageRowBytes);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      mine[j] = make_uint4(__float_as_uint(v[h * 16 + 4 * j]), __float_as_uint(v[h * 16 + 4 * j + 1]),
                           __float_as_uint(v[h * 16 + 4 * j + 2]), __float_as_uint(v[h * 16 + 4 * j + 3]));
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + (lane >> 2);
      const int ch = lane & 3;
      const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * kStageRowBytes + ch * 16);
      if ((row_mask >> rr) & 1u) *reinterpret_cast<uint4*>(row0_ptr + rr * ld + h * 16 + ch * 4) = val;
    }
    __syncwarp();
  }
}

// Epilogue math for one group of 32 consecutive columns of one row; `y` holds the raw accumulator sums on entry.
// c0 = column offset of the group inside the tile (index into EpiConsts), n0 = global column.
__device__ __forceinline__ void epilogue_group(const GemmParams& p, const EpiConsts& ec, uint8_t* stage, float (&y)[32],
                                               int c0, int n0, long long row, bool valid, uint32_t row_mask, int lane,
                                               const float* addp, const float* addl, float& dot,
                                               uint8_t* resid_smem = nullptr, const float* resid_next = nullptr) {
  const bool full = (n0 + 32 <= p.N);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 s4 = *reinterpret_cast<const float4*>(&ec.scale[c0 + j]);
    const float4 h4 = *reinterpret_cast<const float4*>(&ec.shift[c0 + j]);
    y[j] = fmaf(y[j], s4.x, h4.x);
    y[j + 1] = fmaf(y[j + 1], s4.y, h4.y);
    y[j + 2] = fmaf(y[j + 2], s4.z, h4.z);
    y[j + 3] = fmaf(y[j + 3], s4.w, h4.w);
  }
  if (addp || addl) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (full || n0 + j < p.N) {
        if (addp) y[j] += __ldg(addp + n0 + j);
        if (addl) y[j] += __ldg(addl + n0 + j);
      }
  }
  if (p.resid) {
    const float* rp = p.resid + row * p.ld_resid + n0;
    if (full && p.vec_resid) {   // warp-uniform; masked rows are zeroed below
      // (issuing these loads a group ahead - even before the accumulator is complete - was tried and LOST 2 % on the
      //  encoder and 7 % on the scorer to register spills: profiles/r02_ab_prefetch_pairfeatures.txt)
      if (resid_smem) {          // block already in shared memory (cp.async, requested a group ago)
        add_resid_from_smem(y, resid_smem, lane);
        if (resid_next) resid_cp_async(resid_smem, resid_next, p.ld_resid, row_mask, lane);   // ... and the next one leaves now
      } else {
        add_f32_coalesced(y, stage, rp - lane * p.ld_resid, p.ld_resid, row_mask, lane);
      }
    } else if (valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) y[j] += rp[j];
    }
  }
  if (!valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = 0.f;
  }
  if (p.out_f32) {
    float* op = p.out_f32 + row * p.ld_out + n0;
    if (full && p.vec_out) {     // warp-uniform
      store_f32_coalesced(y, stage, op - lane * p.ld_out, p.ld_out, row_mask, lane);
    } else if ((row_mask >> lane) & 1u) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) op[j] = y[j];
    }
  }
  if (p.out_hi || p.dot_w || p.out_z) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (p.scale2) {
        const float4 s4 = *reinterpret_cast<const float4*>(&ec.scale2[c0 + j]);
        const float4 h4 = *reinterpret_cast<const float4*>(&ec.shift2[c0 + j]);
        y[j] = fmaf(y[j], s4.x, h4.x);
        y[j + 1] = fmaf(y[j + 1], s4.y, h4.y);
        y[j + 2] = fmaf(y[j + 2], s4.z, h4.z);
        y[j + 3] = fmaf(y[j + 3], s4.w, h4.w);
      }
      if (p.relu) {
        y[j] = fmaxf(y[j], 0.f); y[j + 1] = fmaxf(y[j + 1], 0.f);
        y[j + 2] = fmaxf(y[j + 2], 0.f); y[j + 3] = fmaxf(y[j + 3], 0.f);
      }
      if (!valid) y[j] = y[j + 1] = y[j + 2] = y[j + 3] = 0.f;
      if (p.dot_w) {   // columns >= N carry dotw == 0
        const float4 w4 = *reinterpret_cast<const float4*>(&ec.dotw[c0 + j]);
        dot = fmaf(y[j], w4.x, dot);
        dot = fmaf(y[j + 1], w4.y, dot);
        dot = fmaf(y[j + 2], w4.z, dot);
        dot = fmaf(y[j + 3], w4.w, dot);
      }
    }
    if (p.out_z) {
      float* zp = p.out_z + row * p.ld_z + n0;
      if (full && p.vec_z) {     // warp-uniform
        store_f32_coalesced(y, stage, zp - lane * p.ld_z, p.ld_z, row_mask, lane);
      } else if ((row_mask >> lane) & 1u) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) zp[j] = y[j];
      }
    }
    if (p.out_hi) {
      uint32_t hi2[16], lo2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split_pack(y[2 * j], y[2 * j + 1], hi2[j], lo2[j]);
      __half* hp = p.out_hi + row * p.ld_split + n0;
      __half* lp = p.out_lo ? p.out_lo + row * p.ld_split + n0 : nullptr;
      if (full && p.vec_split) {   // warp-uniform
        store_halves_coalesced(hi2, stage, hp - lane * p.ld_split, p.ld_split, row_mask, lane);
        if (lp) store_halves_coalesced(lo2, stage, lp - lane * p.ld_split, p.ld_split, row_mask, lane);
      } else if ((row_mask >> lane) & 1u) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) {
            hp[j] = __ushort_as_half((unsigned short)((hi2[j >> 1] >> ((j & 1) * 16)) & 0xffffu));
            if (lp) lp[j] = __ushort_as_half((unsigned short)((lo2[j >> 1] >> ((j & 1) * 16)) & 0xffffu));
          }
      }
    }
  }
}

template <int BK, int NPASS, bool GEN>
__global__ void __launch_bounds__(GEN ? kGenThreads : kGemmThreads, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BK, NPASS, GEN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t gen_base = smem_base + Cfg::kStages * Cfg::kStageBytes;   // GEN: c staging ring (1024-aligned)
  const uint32_t bar_base = gen_base + Cfg::kGenBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  auto cfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 5 + s); };
  auto cempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 5 + kGenStages + s); };
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* epi_smem = smem_raw + (bar_base + 256u - smem_u32(smem_raw));   // EpiConsts, then 8 warp staging blocks

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool conv = p.conv_taps > 0;
  const int total_tiles = p.tiles_m * p.tiles_n;
  const int num_chunks = (p.num_kblocks + p.chunk_kblocks - 1) / p.chunk_kblocks;

  if (warp == 0 && lane == 0) {
    if (!GEN) tma_prefetch_desc(&p.tm_a_hi);
    tma_prefetch_desc(&p.tm_b_hi);
    if (GEN) tma_prefetch_desc(&p.tm_gen_c);
    if (NPASS == 3) {
      if (!GEN) tma_prefetch_desc(&p.tm_a_lo);
      tma_prefetch_desc(&p.tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), GEN ? 3 : 1);   // TMA producer (+ one arrive per warp of the generator team)
      mbar_init(empty_bar(s), 1);
    }
    if (GEN) {
      for (int s = 0; s < kGenStages; ++s) {
        mbar_init(cfull_bar(s), 1);
        mbar_init(cempty_bar(s), 2);   // one arrive per warp of the generator team
      }
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);   // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // A tile that lies entirely in the padding of its sequence needs no arithmetic: every role skips it the
  // same way (the epilogue still writes the zeros the next layer's shifted reads rely on).
  auto tile_is_padding = [&](int m_tile) -> bool {
    if (!conv || p.lengths == nullptr) return false;
    const int b = m_tile / p.conv_tiles_per_seq;
    const int t0 = (m_tile % p.conv_tiles_per_seq) * kBM;
    return (long long)t0 >= p.lengths[b];
  };

  if (warp < 4) {
    reg_dealloc<kCtrlRegs>();
    // The two control roles are entered through elect.sync, not `lane == 0`: the compiler then knows that exactly one
    // thread executes the warp-level (uniform datapath) UTMALDG / UTCHMMA / UTCBAR instructions and emits them back to
    // back; behind a lane test it wraps every one of them in an ELECT ... BRA.U.ANY loop (5 extra instructions per MMA),
    // and the single issuing thread - not the tensor core - becomes the limit for tiles narrower than 256 columns
    // (profiles/r02_issue_thread_bound.txt).
    if (warp == 0 && elect_one()) {
      // ---------------------------------------------------------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      int cs = 0;
      uint32_t cphase = 0;
      long long c_seq = 0, main_seq = 0;   // GEN: k-blocks whose label tile / operand stage has been issued
      int c_tile = blockIdx.x, c_kb = 0;
      const uint32_t tx_bytes = Cfg::kPlanes * ((GEN ? 0u : (uint32_t)Cfg::kATile) + (uint32_t)p.bn * BK * 2);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.tiles_n, n_tile = tile % p.tiles_n;
        if (tile_is_padding(m_tile)) continue;
        int seq = 0, t0 = 0;
        if (conv) {
          seq = m_tile / p.conv_tiles_per_seq;
          t0 = (m_tile % p.conv_tiles_per_seq) * kBM;
        }
        // conv: (tap, channel block) walk without a division per k-block
        int cb = 0, kc = 0, kbase = 0, t = t0 - (p.conv_taps / 2) * p.conv_dil;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
          mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
          int kcol = kb * BK;   // K coordinate in the weight rows
          if (conv) {
            kcol = kbase + kc;
            tma_load_3d(sa, &p.tm_a_hi, full_bar(stage), kc, t, seq);
            if (NPASS == 3) tma_load_3d(sa + Cfg::kATile, &p.tm_a_lo, full_bar(stage), kc, t, seq);
            kc += BK;
            if (++cb == p.conv_cblocks) {   // next tap
              cb = 0;
              kc = 0;
              kbase += p.conv_cpad;
              t += p.conv_dil;
            }
          } else if (!GEN) {
            if (p.a_kblk) {
              tma_load_3d(sa, &p.tm_a_hi, full_bar(stage), kcol & 63, m_tile * kBM, kcol >> 6);
              if (NPASS == 3) tma_load_3d(sa + Cfg::kATile, &p.tm_a_lo, full_bar(stage), kcol & 63, m_tile * kBM, kcol >> 6);
            } else {
              tma_load_2d(sa, &p.tm_a_hi, full_bar(stage), kcol, m_tile * kBM);
              if (NPASS == 3) tma_load_2d(sa + Cfg::kATile, &p.tm_a_lo, full_bar(stage), kcol, m_tile * kBM);
            }
          }
          if (p.b_kblk) {
            tma_load_3d(sb, &p.tm_b_hi, full_bar(stage), kcol & 63, n_tile * p.bn, kcol >> 6);
            if (NPASS == 3) tma_load_3d(sb + Cfg::kBTile, &p.tm_b_lo, full_bar(stage), kcol & 63, n_tile * p.bn, kcol >> 6);
          } else {
            tma_load_2d(sb, &p.tm_b_hi, full_bar(stage), kcol, n_tile * p.bn);
            if (NPASS == 3) tma_load_2d(sb + Cfg::kBTile, &p.tm_b_lo, full_bar(stage), kcol, n_tile * p.bn);
          }
          if (GEN) {
            // Label-half tiles run AHEAD of the operand stages (up to the depth of their own ring), so that the
            // generator can start on a k-block the moment its operand stage is released.
            while (c_seq < main_seq + kGenStages && c_tile < total_tiles) {
              mbar_wait(cempty_bar(cs), cphase ^ 1);
              mbar_arrive_expect_tx(cfull_bar(cs), kGenTileBytes);
              tma_load_2d(gen_base + cs * kGenTileBytes, &p.tm_gen_c, cfull_bar(cs), c_kb * BK,
                          (int)(((long long)(c_tile / p.tiles_n) * kBM) % p.pair_nl));
              if (++cs == kGenStages) {
                cs = 0;
                cphase ^= 1;
              }
              ++c_seq;
              if (++c_kb == p.num_kblocks) {
                c_kb = 0;
                c_tile += gridDim.x;
              }
            }
            ++main_seq;
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1 && elect_one()) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc = make_idesc_f16(kBM, p.bn, /*fp16*/ 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (NPASS == 3 && p.split_corr) {
        // ---- split accumulators: buffer 0 = hi*hi of the current chunk, buffer 1 = corrections of the whole tile
        uint32_t n_main = 0, n_corr = 0;   // uses of each buffer so far (mbarrier parities)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const int m_tile = tile / p.tiles_n;
          if (tile_is_padding(m_tile)) continue;
          const uint32_t d_main = tmem_base, d_corr = tmem_base + kMaxBN;
          mbar_wait(tempty_bar(1), (n_corr & 1u) ^ 1u);   // the previous tile's corrections have been read
          tc_fence_after();
          bool first_corr = true;
          int kb = 0;
          for (int chunk = 0; chunk < num_chunks; ++chunk) {
            const int kb_end = min(kb + p.chunk_kblocks, p.num_kblocks);
            mbar_wait(tempty_bar(0), (n_main & 1u) ^ 1u);   // the previous chunk has been promoted
            tc_fence_after();
            int st = stage;
            uint32_t ph = phase;
            bool first = true;
            for (int k2 = kb; k2 < kb_end; ++k2) {            // phase 1: the large products
              mbar_wait(full_bar(st), ph);
              tc_fence_after();
              const uint32_t sa = smem_base + st * Cfg::kStageBytes;
              const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks) {
                umma_f16(d_main, make_kmajor_desc<Cfg::kSwizzle>(sa + ks * 32), make_kmajor_desc<Cfg::kSwizzle>(sb + ks * 32),
                         idesc, first ? 0u : 1u);
                first = false;
              }
              if (++st == Cfg::kStages) {
                st = 0;
                ph ^= 1;
              }
            }
            umma_commit(tfull_bar(0));                        // epilogue warps promote buffer 0 ...
            ++n_main;
            for (int k2 = kb; k2 < kb_end; ++k2) {            // ... while phase 2 issues the chunk's corrections
              const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
              const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks) {
                const uint64_t a_hi = make_kmajor_desc<Cfg::kSwizzle>(sa + ks * 32);
                const uint64_t b_hi = make_kmajor_desc<Cfg::kSwizzle>(sb + ks * 32);
                const uint64_t a_lo = make_kmajor_desc<Cfg::kSwizzle>(sa + Cfg::kATile + ks * 32);
                const uint64_t b_lo = make_kmajor_desc<Cfg::kSwizzle>(sb + Cfg::kBTile + ks * 32);
                umma_f16(d_corr, a_lo, b_hi, idesc, first_corr ? 0u : 1u);
                umma_f16(d_corr, a_hi, b_lo, idesc, 1);
                first_corr = false;
              }
              umma_commit(empty_bar(stage));                  // frees the smem slot once every MMA that reads it is done
              if (++stage == Cfg::kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
            kb = kb_end;
          }
          umma_commit(tfull_bar(1));                          // corrections of the tile complete
          ++n_corr;
        }
      } else {
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.tiles_n;
        if (tile_is_padding(m_tile)) continue;
        int kb = 0;
        for (int chunk = 0; chunk < num_chunks; ++chunk) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // the epilogue has drained this buffer
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kMaxBN;
          const int kb_end = min(kb + p.chunk_kblocks, p.num_kblocks);
          bool first = true;
          for (; kb < kb_end; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
            // one descriptor per operand tile; the k-steps and the lo planes are reached by adding to its address field
            // (16-byte units; the field cannot carry out: shared memory is < 256 KB)
            const uint64_t a_hi0 = make_kmajor_desc<Cfg::kSwizzle>(sa);
            const uint64_t b_hi0 = make_kmajor_desc<Cfg::kSwizzle>(sb);
            // Within a k-block the small hi*lo / lo*hi products go first: the accumulator add truncates relative
            // to the accumulator's magnitude, so every tiny term added before the hi*hi terms is added exactly.
            if (NPASS == 3) {
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks) {
                umma_f16(d_tmem, a_hi0 + (uint64_t)((Cfg::kATile + ks * 32) >> 4), b_hi0 + (uint64_t)((ks * 32) >> 4), idesc,
                         first ? 0u : 1u);                                                    // A_lo * B_hi
                umma_f16(d_tmem, a_hi0 + (uint64_t)((ks * 32) >> 4), b_hi0 + (uint64_t)((Cfg::kBTile + ks * 32) >> 4), idesc,
                         1);                                                                  // A_hi * B_lo
                first = false;
              }
            }
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              umma_f16(d_tmem, a_hi0 + (uint64_t)((ks * 32) >> 4), b_hi0 + (uint64_t)((ks * 32) >> 4), idesc,
                       first ? 0u : 1u);                                                      // A_hi * B_hi
              first = false;
            }
            umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs have read it
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit(tfull_bar(acc));       // chunk complete -> epilogue warps promote it
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
      }
    }
  } else if (GEN && warp >= 12) {
    // ------------------------------------------------------------------ A-operand generator (4 warps, GEN only)
    // A[r][k] = relu(a[b][k] + c[l0 + r][k]) for the tile's protein b and its 128 consecutive label rows
    // (layer 1 of the pair scorer after the exact split of Linear(2d -> H), ProtNote.py:112-126,293).
    // The fp32 c tile arrives by TMA (SWIZZLE_128B) in a 3-deep staging ring; thread g owns tile row g: it reads its
    // 128-byte row (conflict-free: chunk ^= row & 7), adds the protein half, applies ReLU, splits into fp16 planes
    // and stores them in the SWIZZLE_64B layout the UMMA descriptors expect (chunk ^= (row >> 1) & 3).
    static_assert(!GEN || BK == 32, "the generator writes the 64-byte-swizzle layout");
    // Two teams of two warps take alternate k-blocks, so the load -> convert -> store -> publish latency of one
    // k-block overlaps the other team's; within a team thread t owns tile rows t and t + 64.
    reg_dealloc<kGenWarpRegs>();
    const int team = (warp - 12) >> 1;
    const int t = (threadIdx.x - 384) & 63;
    long long gk = 0;   // k-blocks this CTA has been through (both teams count all of them)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.tiles_n;
      const long long row0 = (long long)m_tile * kBM;
      const bool ok0 = row0 + t < p.M, ok1 = row0 + t + 64 < p.M;
      const float* arow = p.gen_a + (row0 / p.pair_nl) * p.ld_gen_a;   // one protein per tile (pair_nl % 128 == 0)
      for (int kb = 0; kb < p.num_kblocks; ++kb, ++gk) {
        if ((int)(gk & 1) != team) continue;
        const int stage = (int)(gk % Cfg::kStages);
        const uint32_t phase = (uint32_t)((gk / Cfg::kStages) & 1);
        const int cs = (int)(gk % kGenStages);
        const uint32_t cphase = (uint32_t)((gk / kGenStages) & 1);
        mbar_wait(cfull_bar(cs), cphase);
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sc = gen_base + cs * kGenTileBytes;
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const float* ak = arow + kb * BK;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {      // 8 k-values = one 16-byte chunk of each fp16 plane
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(ak + 8 * q4));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(ak + 8 * q4 + 4));
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = t + 64 * rr;
            float cv[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t chunk = (uint32_t)((2 * q4 + h) ^ (r & 7));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(cv[4 * h]), "=f"(cv[4 * h + 1]), "=f"(cv[4 * h + 2]), "=f"(cv[4 * h + 3])
                           : "r"(sc + r * 128 + (chunk << 4)));
            }
            uint32_t hi[4], lo[4];
            split_pack_pos(fmaxf(a0.x + cv[0], 0.f), fmaxf(a0.y + cv[1], 0.f), hi[0], lo[0]);
            split_pack_pos(fmaxf(a0.z + cv[2], 0.f), fmaxf(a0.w + cv[3], 0.f), hi[1], lo[1]);
            split_pack_pos(fmaxf(a1.x + cv[4], 0.f), fmaxf(a1.y + cv[5], 0.f), hi[2], lo[2]);
            split_pack_pos(fmaxf(a1.z + cv[6], 0.f), fmaxf(a1.w + cv[7], 0.f), hi[3], lo[3]);
            if (!(rr == 0 ? ok0 : ok1)) {
#pragma unroll
              for (int e = 0; e < 4; ++e) hi[e] = lo[e] = 0u;
            }
            const uint32_t dst = sa + (uint32_t)(r * 64 + ((q4 ^ ((r >> 1) & 3)) << 4));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                         "r"(hi[3])
                         : "memory");
            if (NPASS == 3)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + Cfg::kATile), "r"(lo[0]), "r"(lo[1]),
                           "r"(lo[2]), "r"(lo[3])
                           : "memory");
          }
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(full_bar(stage));
          mbar_arrive(cempty_bar(cs));
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    reg_alloc<GEN ? kGenEpiRegs : kEpiRegs>();
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;         // which half of the tile's column groups
    const int ngroups = p.bn >> 5;            // 32-column groups in a tile (<= 8)
    const int g_lo = half == 0 ? 0 : (ngroups + 1) >> 1;
    const int g_hi = half == 0 ? (ngroups + 1) >> 1 : ngroups;
    const int r_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t n_main = 0, n_corr = 0;   // split_corr: uses of TMEM buffer 0 / 1 so far (mbarrier parities)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.tiles_n, n_tile = tile % p.tiles_n;
      const bool padding_tile = tile_is_padding(m_tile);
      long long row;        // row of D / of every [M][ld] epilogue tensor
      bool in_range;        // row exists
      bool valid;           // row exists and is not masked
      if (conv) {
        const int seq = m_tile / p.conv_tiles_per_seq;
        const int t = (m_tile % p.conv_tiles_per_seq) * kBM + r_in_tile;
        row = (long long)seq * p.conv_T + t;
        in_range = t < p.conv_T;
        valid = in_range && (p.lengths == nullptr || (long long)t < p.lengths[seq]);
      } else {
        row = (long long)m_tile * kBM + r_in_tile;
        in_range = row < p.M;
        valid = in_range;
      }
      const float* addp = nullptr;
      const float* addl = nullptr;
      if (valid && p.pair_nl > 0) {
        if (p.add_p) addp = p.add_p + (row / p.pair_nl) * p.ld_add_p;
        if (p.add_l) addl = p.add_l + (row % p.pair_nl) * p.ld_add_l;
      }
      // per-tile epilogue vectors -> shared memory (one column per epilogue thread)
      EpiConsts& ec = *reinterpret_cast<EpiConsts*>(epi_smem);
      uint8_t* stage = epi_smem + sizeof(EpiConsts) + (warp - 4) * kStageBytesPerWarp;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // everyone is done with the previous tile's vectors
      {
        const int cc = threadIdx.x - 128;
        const int n = n_tile * p.bn + cc;
        const bool ok = cc < p.bn && n < p.N;
        ec.scale[cc] = (ok && p.scale) ? __ldg(p.scale + n) : 1.f;
        ec.shift[cc] = (ok && p.shift) ? __ldg(p.shift + n) : 0.f;
        ec.scale2[cc] = (ok && p.scale2) ? __ldg(p.scale2 + n) : 1.f;
        ec.shift2[cc] = (ok && p.shift2) ? __ldg(p.shift2 + n) : 0.f;
        ec.dotw[cc] = (ok && p.dot_w) ? __ldg(p.dot_w + n) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t row_mask = __ballot_sync(0xffffffffu, in_range);
      float sums[4][32];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 32; ++j) sums[g][j] = 0.f;
      // Every drain adds one TMEM buffer into the fp32 register sums: this warp's (up to four) 32-column groups two at a
      // time; the buffer is handed back to the MMA issuer as soon as the last load has landed in registers, BEFORE the
      // additions, so the round trip commit -> drain -> release that bounds short chunks stays as short as possible.
      // split_corr: num_chunks drains of buffer 0 (the hi*hi products of each chunk), then one of buffer 1 (the tile's
      // corrections); otherwise the chunks alternate between the two buffers.
      if (!padding_tile) {
        const bool split = NPASS == 3 && p.split_corr;
        const int n_drains = split ? num_chunks + 1 : num_chunks;
        for (int d = 0; d < n_drains; ++d) {
          int buf;
          uint32_t parity;
          if (split) {
            if (d < num_chunks) {
              buf = 0;
              parity = (n_main++) & 1u;
            } else {
              buf = 1;
              parity = (n_corr++) & 1u;
            }
          } else {
            buf = acc;
            parity = acc_phase;
            if (++acc == 2) {
              acc = 0;
              acc_phase ^= 1;
            }
          }
          mbar_wait(tfull_bar(buf), parity);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + buf * kMaxBN + g_lo * 32 + ((uint32_t)(q * 32) << 16);
          const int ng = g_hi - g_lo;   // warp-uniform
          uint32_t v0[32], v1[32];
          if (ng > 0) tmem_ld32(t_addr, v0);
          if (ng > 1) tmem_ld32(t_addr + 32, v1);
          tmem_ld_wait();
          if (ng <= 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
          }
          if (ng > 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sums[0][j] += __uint_as_float(v0[j]);
          }
          if (ng > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sums[1][j] += __uint_as_float(v1[j]);
          }
          if (ng > 2) {
            tmem_ld32(t_addr + 64, v0);
            if (ng > 3) tmem_ld32(t_addr + 96, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
#pragma unroll
            for (int j = 0; j < 32; ++j) sums[2][j] += __uint_as_float(v0[j]);
            if (ng > 3) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sums[3][j] += __uint_as_float(v1[j]);
            }
          }
        }
      }
      if (p.trunc_comp != 0.f) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int j = 0; j < 32; ++j) sums[g][j] = fmaf(sums[g][j], p.trunc_comp, sums[g][j]);
      }
      // Linear(H -> 1): the dot with w_out is summed in fp32 over 32 columns at a time and across groups in fp64 (the
      // partial sums of a calibrated / trained output neuron are far larger than the logit they cancel to); the fp64
      // partial leaves as two floats (hi, lo) and finalize_logits_kernel adds all of them in fp64
      double dot = 0.0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c0 = (g_lo + g) * 32;
        const int n0 = n_tile * p.bn + c0;
        if (g_lo + g < g_hi && n0 < p.N) {   // warp-uniform (rows out of range are masked inside)
          float dotg = 0.f;
          epilogue_group(p, ec, stage, sums[g], c0, n0, row, valid, row_mask, lane, addp, addl, dotg);
          dot += (double)dotg;
        }
      }
      if (p.dot_w && in_range) {
        const float dhi = (float)dot;
        float2* dst = reinterpret_cast<float2*>(p.dot_out) + (row * p.tiles_n + n_tile) * 2 + half;
        *dst = make_float2(dhi, (float)(dot - (double)dhi));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pn
