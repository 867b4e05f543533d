// The one tensor-core engine every dense contraction of the scoring path runs on (sm_100a only).
//
//   D[M,N] = sum over K of A[M,K] * B[N,K]        (both operands K-major, fp32 accumulate in TMEM)
//
// Operands are fp32 values carried as TWO fp16 planes (hi = fp16(x), lo = fp16(x - hi)).  In `strict`
// mode (NPASS == 3) every k-step issues three tcgen05.mma into the same accumulator,
//   A_lo*B_hi + A_hi*B_lo + A_hi*B_hi,
// which reproduces an fp32 GEMM to ~2^-22 relative (the dropped lo*lo term); `fast` mode (NPASS == 1) issues
// A_hi*B_hi only.  Weights are pre-scaled by a power of two at pack time so that their lo plane stays in
// fp16's normal range; the epilogue's per-column `scale` undoes it (and carries the folded BatchNorm).
//
// Structure: persistent CTAs (grid = #SMs), 8 warps, warp-specialised
//   warp 0   TMA producer   (cp.async.bulk.tensor 2D for plain matrices, 3D for dilated-conv taps)
//   warp 1   MMA issuer     (one elected lane, tcgen05.mma cta_group::1, M=128, N=bn<=256, K=16)
//   warp 2   TMEM allocator (512 columns = two accumulator buffers)
//   warps 4-7 epilogue      (tcgen05.ld 32x32b, one accumulator row per thread)
// with an smem ring (full/empty mbarriers) between 0 and 1 and a two-deep TMEM ring between 1 and 4-7,
// so the epilogue of tile i overlaps the main loop of tile i+1.
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
constexpr int kGemmThreads = 256;
constexpr int kSmemBudget = 200 * 1024;

struct alignas(64) GemmParams {
  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  int M, N;              // logical extents of D (conv: M = B*T rows addressed as (b,t))
  int bn;                // tile width, multiple of 32, <= 256
  int tiles_m, tiles_n;
  int num_kblocks;       // K / BK (conv: taps * cblocks)
  // conv addressing (conv_taps == 0 -> plain)
  int conv_taps, conv_cblocks, conv_cpad, conv_dil, conv_T, conv_tiles_per_seq;
  const long long* lengths;      // [B] valid positions per sequence (conv) or nullptr
  // pair-row addressing for the additive row terms: row r -> (r / pair_nl, r % pair_nl)
  int pair_nl;
  const float* add_p; long long ld_add_p;   // [B][ld] added per protein   (nullable)
  const float* add_l; long long ld_add_l;   // [L][ld] added per label row (nullable)
  // epilogue: y = acc*scale[n] + shift[n] (+ add rows) (+ resid); masked rows -> 0
  const float* scale; const float* shift;   // [N] (nullable -> 1 / 0)
  const float* resid; long long ld_resid;   // fp32 [M][ld] (nullable)
  float* out_f32; long long ld_out;         // y stored here (nullable)
  // z = relu?(y*scale2 + shift2) ; masked rows -> 0
  const float* scale2; const float* shift2; // [N] (nullable)
  int relu;
  __half* out_hi; __half* out_lo; long long ld_split;   // z as fp16 planes (nullable; out_lo nullable)
  const float* dot_w; float* dot_out;       // dot_out[row*tiles_n + n_tile] = sum_n z*dot_w[n] (nullable)
  int vec_out, vec_resid, vec_split;        // 16-byte vector access is legal for that tensor (host-checked)
};

template <int BK, int NPASS>
struct GemmCfg {
  static constexpr int kSwizzle = BK * 2;                    // bytes per smem row
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kATile = kBM * BK * 2;
  static constexpr int kBTile = kMaxBN * BK * 2;
  static constexpr int kStageBytes = kPlanes * (kATile + kBTile);
  static constexpr int kStages = kSmemBudget / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "pipeline too shallow");
};

__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

template <int BK, int NPASS>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BK, NPASS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool conv = p.conv_taps > 0;
  const int total_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a_hi);
    tma_prefetch_desc(&p.tm_b_hi);
    if (NPASS == 3) {
      tma_prefetch_desc(&p.tm_a_lo);
      tma_prefetch_desc(&p.tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = Cfg::kPlanes * (Cfg::kATile + (uint32_t)p.bn * BK * 2);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.tiles_n, n_tile = tile % p.tiles_n;
        if (tile_is_padding(m_tile)) continue;
        int seq = 0, t0 = 0;
        if (conv) {
          seq = m_tile / p.conv_tiles_per_seq;
          t0 = (m_tile % p.conv_tiles_per_seq) * kBM;
        }
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
          mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
          int kcol = kb * BK;   // K coordinate in the weight rows
          if (conv) {
            const int tap = kb / p.conv_cblocks;
            const int kc = (kb - tap * p.conv_cblocks) * BK;
            const int t = t0 + (tap - p.conv_taps / 2) * p.conv_dil;
            kcol = tap * p.conv_cpad + kc;
            tma_load_3d(sa, &p.tm_a_hi, full_bar(stage), kc, t, seq);
            if (NPASS == 3) tma_load_3d(sa + Cfg::kATile, &p.tm_a_lo, full_bar(stage), kc, t, seq);
          } else {
            tma_load_2d(sa, &p.tm_a_hi, full_bar(stage), kcol, m_tile * kBM);
            if (NPASS == 3) tma_load_2d(sa + Cfg::kATile, &p.tm_a_lo, full_bar(stage), kcol, m_tile * kBM);
          }
          tma_load_2d(sb, &p.tm_b_hi, full_bar(stage), kcol, n_tile * p.bn);
          if (NPASS == 3) tma_load_2d(sb + Cfg::kBTile, &p.tm_b_lo, full_bar(stage), kcol, n_tile * p.bn);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, p.bn, /*fp16*/ 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.tiles_n;
        if (tile_is_padding(m_tile)) continue;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMaxBN;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t a_hi = make_kmajor_desc<Cfg::kSwizzle>(sa + ks * 32);
            const uint64_t b_hi = make_kmajor_desc<Cfg::kSwizzle>(sb + ks * 32);
            const uint32_t first = (kb | ks) != 0;
            if (NPASS == 3) {
              const uint64_t a_lo = make_kmajor_desc<Cfg::kSwizzle>(sa + Cfg::kATile + ks * 32);
              const uint64_t b_lo = make_kmajor_desc<Cfg::kSwizzle>(sb + Cfg::kBTile + ks * 32);
              umma_f16(d_tmem, a_lo, b_hi, idesc, first);   // small terms first
              umma_f16(d_tmem, a_hi, b_lo, idesc, 1);
              umma_f16(d_tmem, a_hi, b_hi, idesc, 1);
            } else {
              umma_f16(d_tmem, a_hi, b_hi, idesc, first);
            }
          }
          umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull_bar(acc));       // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;                  // == warp % 4: the TMEM lane quarter this warp may read
    const int r_in_tile = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
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
      if (!padding_tile) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
      }
      float dot = 0.f;
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t v[32];
        if (!padding_tile) {
          tmem_ld32(tmem_base + acc * kMaxBN + c0 + ((uint32_t)(ew * 32) << 16), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int n0 = n_tile * p.bn + c0;
        if (!in_range || n0 >= p.N) continue;
        const bool full = (n0 + 32 <= p.N);
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + j;
          float a = __uint_as_float(v[j]);
          if (full || n < p.N) {
            const float s = p.scale ? __ldg(p.scale + n) : 1.f;
            const float h = p.shift ? __ldg(p.shift + n) : 0.f;
            a = fmaf(a, s, h);
            if (addp) a += __ldg(addp + n);
            if (addl) a += __ldg(addl + n);
          }
          y[j] = a;
        }
        if (p.resid && valid) {
          const float* rp = p.resid + row * p.ld_resid + n0;
          if (full && p.vec_resid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
              y[j] += r4.x; y[j + 1] += r4.y; y[j + 2] += r4.z; y[j + 3] += r4.w;
            }
          } else {
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
          if (full && p.vec_out) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(op + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) op[j] = y[j];
          }
        }
        if (p.out_hi || p.dot_w) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            float z = y[j];
            if (full || n < p.N) {
              if (p.scale2) z = fmaf(z, __ldg(p.scale2 + n), p.shift2 ? __ldg(p.shift2 + n) : 0.f);
              if (p.relu) z = fmaxf(z, 0.f);
              if (!valid) z = 0.f;
              if (p.dot_w) dot = fmaf(z, __ldg(p.dot_w + n), dot);
            } else {
              z = 0.f;
            }
            y[j] = z;
          }
          if (p.out_hi) {
            uint32_t hi2[16], lo2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __half h0, l0, h1, l1;
              split_f16(fmaxf(fminf(y[2 * j], 65504.f), -65504.f), h0, l0);
              split_f16(fmaxf(fminf(y[2 * j + 1], 65504.f), -65504.f), h1, l1);
              hi2[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
              lo2[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            __half* hp = p.out_hi + row * p.ld_split + n0;
            __half* lp = p.out_lo ? p.out_lo + row * p.ld_split + n0 : nullptr;
            if (full && p.vec_split) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                *reinterpret_cast<uint4*>(hp + 2 * j) = make_uint4(hi2[j], hi2[j + 1], hi2[j + 2], hi2[j + 3]);
                if (lp) *reinterpret_cast<uint4*>(lp + 2 * j) = make_uint4(lo2[j], lo2[j + 1], lo2[j + 2], lo2[j + 3]);
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) {
                  hp[j] = __ushort_as_half((unsigned short)((hi2[j >> 1] >> ((j & 1) * 16)) & 0xffffu));
                  if (lp) lp[j] = __ushort_as_half((unsigned short)((lo2[j >> 1] >> ((j & 1) * 16)) & 0xffffu));
                }
            }
          }
        }
      }
      if (p.dot_w && in_range) p.dot_out[row * p.tiles_n + n_tile] = dot;
      if (!padding_tile) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
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
