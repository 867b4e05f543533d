// CTA-pair (cta_group::2) variant of the tensor-core engine in pn_gemm.cuh.
//
// Two CTAs of one cluster (= the two SMs of a TPC) work on one 256-row x bn-column tile: CTA rank r owns rows
// [128 r, 128 r + 128) of the tile (its own A operand and its own half of the accumulator in its own TMEM) and
// loads only HALF of the B operand (weight rows [r bn/2, (r+1) bn/2)); tcgen05.mma.cta_group::2, issued by the leader
// CTA alone, reads both halves.  Per SM this halves the B traffic from L2 and from shared memory - the 1-CTA kernel
// moves ~120 KB of shared-memory traffic per 768-cycle k-block, which is where it saturates (profiles/, fused
// generator probe) - and the smaller stage (32 KB instead of 48 KB) makes the operand ring six deep.
//
// Protocol differences from the 1-CTA kernel (everything else - chunk promotion, epilogue - is shared code):
//   full[s]    lives in the LEADER; both CTAs' TMA loads complete_tx on it (cta_group::2 loads, barrier address with
//              the peer bit cleared); the leader's producer arms it with the bytes of both CTAs.
//   empty[s]   one per CTA; released by tcgen05.commit.cta_group::2 ... multicast::cluster to both.
//   tfull[a]   one per CTA, same multicast commit; tempty[a] lives in the leader and collects the 8 epilogue warps
//              of BOTH CTAs (the peer's arrive remotely through shared::cluster).
#pragma once

#include "pn_gemm.cuh"

namespace pn {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address (pair of 2)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// The same loads with an L2 eviction-priority hint (createpolicy-encoded 64-bit operand).  Motivation: ncu shows the
// scorer's layer-2 launch reading 16.8 GB from HBM where its A planes are 6.4 GB - the 38 MB of weights do not stay
// L2-resident while 12 KB of activations per row stream past them.  Experiment (option l2_hints): B evict_last made no
// difference, A evict_first cost 11-15 % (the A tile that twelve clusters share is dropped before the last one has read
// it).  DRAM runs at 21 % and is not the limiter; the hints stay off.
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma2_load_2d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int BK, int NPASS>
struct Gemm2Cfg {
  static constexpr int kSwizzle = BK * 2;
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kATile = kBM * BK * 2;              // this CTA's 128 rows
  static constexpr int kBTile = (kMaxBN / 2) * BK * 2;     // this CTA's half of the weight rows
  static constexpr int kStageBytes = kPlanes * (kATile + kBTile);
  static constexpr int kResidBytes = 8 * kResidBlockBytes;      // cp.async residual blocks, one per epilogue warp
  static constexpr int kStagesRaw = (kSmemBudget - kResidBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kEpiSmemBytes + kResidBytes;
};

template <int BK, int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = Gemm2Cfg<BK, NPASS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* epi_smem = smem_raw + (bar_base + 256u - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const bool conv = p.conv_taps > 0;
  const int total_tiles = p.tiles_m * p.tiles_n;     // tiles_m counts 256-row tiles here
  const int num_chunks = (p.num_kblocks + p.chunk_kblocks - 1) / p.chunk_kblocks;
  const int half_bn = p.bn >> 1;

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
      mbar_init(tempty_bar(a), 16);   // 8 epilogue warps of each CTA (used in the leader only)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem2_alloc(tmem_slot, 512);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto tile_is_padding = [&](int m_tile) -> bool {
    if (!conv || p.lengths == nullptr) return false;
    const int b = m_tile / p.conv_tiles_per_seq;
    const int t0 = (m_tile % p.conv_tiles_per_seq) * (2 * kBM);
    return (long long)t0 >= p.lengths[b];
  };

  if (warp < 4) {
    reg_dealloc<kCtrlRegs>();
    // control roles through elect.sync (see pn_gemm.cuh): no per-instruction election loops around UTMALDG / UTCHMMA
    if (warp == 0 && elect_one()) {
      // ---------------------------------------------------------------- TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t my_bytes = Cfg::kPlanes * (Cfg::kATile + (uint32_t)half_bn * BK * 2);
      const uint64_t pol_a = p.l2_hints == 2 ? kL2EvictFirst : 0x1000000000000000ull;   // 1: A evict_normal
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int m_tile = tile / p.tiles_n, n_tile = tile % p.tiles_n;
        if (tile_is_padding(m_tile)) continue;
        int seq = 0, t0 = 0;
        if (conv) {
          seq = m_tile / p.conv_tiles_per_seq;
          t0 = (m_tile % p.conv_tiles_per_seq) * (2 * kBM) + (int)rank * kBM;
        }
        const int row0 = m_tile * (2 * kBM) + (int)rank * kBM;
        const int brow0 = n_tile * p.bn + (int)rank * half_bn;
        int cb = 0, kc = 0, kbase = 0, t = t0 - (p.conv_taps / 2) * p.conv_dil;   // conv: (tap, channel block) walk
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
          const uint32_t fb = full_bar(stage) & kPeerBitMask;          // the LEADER's barrier
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * my_bytes);
          int kcol = kb * BK;
          if (conv) {
            kcol = kbase + kc;
            if (p.l2_hints) {
              tma2_load_3d_hint(sa, &p.tm_a_hi, fb, kc, t, seq, pol_a);
              if (NPASS == 3) tma2_load_3d_hint(sa + Cfg::kATile, &p.tm_a_lo, fb, kc, t, seq, pol_a);
            } else {
              tma2_load_3d(sa, &p.tm_a_hi, fb, kc, t, seq);
              if (NPASS == 3) tma2_load_3d(sa + Cfg::kATile, &p.tm_a_lo, fb, kc, t, seq);
            }
            kc += BK;
            if (++cb == p.conv_cblocks) {   // next tap
              cb = 0;
              kc = 0;
              kbase += p.conv_cpad;
              t += p.conv_dil;
            }
          } else {
            if (p.l2_hints) {
              tma2_load_2d_hint(sa, &p.tm_a_hi, fb, kcol, row0, pol_a);
              if (NPASS == 3) tma2_load_2d_hint(sa + Cfg::kATile, &p.tm_a_lo, fb, kcol, row0, pol_a);
            } else {
              tma2_load_2d(sa, &p.tm_a_hi, fb, kcol, row0);
              if (NPASS == 3) tma2_load_2d(sa + Cfg::kATile, &p.tm_a_lo, fb, kcol, row0);
            }
          }
          if (p.l2_hints) {
            tma2_load_2d_hint(sb, &p.tm_b_hi, fb, kcol, brow0, kL2EvictLast);
            if (NPASS == 3) tma2_load_2d_hint(sb + Cfg::kBTile, &p.tm_b_lo, fb, kcol, brow0, kL2EvictLast);
          } else {
            tma2_load_2d(sb, &p.tm_b_hi, fb, kcol, brow0);
            if (NPASS == 3) tma2_load_2d(sb + Cfg::kBTile, &p.tm_b_lo, fb, kcol, brow0);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1 && leader && elect_one()) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA only)
      const uint32_t idesc = make_idesc_f16(2 * kBM, p.bn, /*fp16*/ 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int m_tile = tile / p.tiles_n;
        if (tile_is_padding(m_tile)) continue;
        int kb = 0;
        for (int chunk = 0; chunk < num_chunks; ++chunk) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kMaxBN;
          const int kb_end = min(kb + p.chunk_kblocks, p.num_kblocks);
          bool first = true;
          for (; kb < kb_end; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            const uint32_t sb = sa + Cfg::kPlanes * Cfg::kATile;
            const uint64_t a_hi0 = make_kmajor_desc<Cfg::kSwizzle>(sa);
            const uint64_t b_hi0 = make_kmajor_desc<Cfg::kSwizzle>(sb);
            if (NPASS == 3) {
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks) {
                umma2_f16(d_tmem, a_hi0 + (uint64_t)((Cfg::kATile + ks * 32) >> 4), b_hi0 + (uint64_t)((ks * 32) >> 4), idesc,
                          first ? 0u : 1u);
                umma2_f16(d_tmem, a_hi0 + (uint64_t)((ks * 32) >> 4), b_hi0 + (uint64_t)((Cfg::kBTile + ks * 32) >> 4), idesc, 1);
                first = false;
              }
            }
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              umma2_f16(d_tmem, a_hi0 + (uint64_t)((ks * 32) >> 4), b_hi0 + (uint64_t)((ks * 32) >> 4), idesc, first ? 0u : 1u);
              first = false;
            }
            umma2_commit_both(empty_bar(stage));
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma2_commit_both(tfull_bar(acc));
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps per CTA)
    reg_alloc<kEpiRegs>();
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int ngroups = p.bn >> 5;
    const int g_lo = half == 0 ? 0 : (ngroups + 1) >> 1;
    const int g_hi = half == 0 ? (ngroups + 1) >> 1 : ngroups;
    const int r_in_tile = (int)rank * kBM + q * 32 + lane;     // row inside the 256-row tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int m_tile = tile / p.tiles_n, n_tile = tile % p.tiles_n;
      const bool padding_tile = tile_is_padding(m_tile);
      long long row;
      bool in_range, valid;
      if (conv) {
        const int seq = m_tile / p.conv_tiles_per_seq;
        const int t = (m_tile % p.conv_tiles_per_seq) * (2 * kBM) + r_in_tile;
        row = (long long)seq * p.conv_T + t;
        in_range = t < p.conv_T;
        valid = in_range && (p.lengths == nullptr || (long long)t < p.lengths[seq]);
      } else {
        row = (long long)m_tile * (2 * kBM) + r_in_tile;
        in_range = row < p.M;
        valid = in_range;
      }
      const float* addp = nullptr;
      const float* addl = nullptr;
      if (valid && p.pair_nl > 0) {
        if (p.add_p) addp = p.add_p + (row / p.pair_nl) * p.ld_add_p;
        if (p.add_l) addl = p.add_l + (row % p.pair_nl) * p.ld_add_l;
      }
      EpiConsts& ec = *reinterpret_cast<EpiConsts*>(epi_smem);
      uint8_t* stage = epi_smem + sizeof(EpiConsts) + (warp - 4) * kStageBytesPerWarp;
      asm volatile("bar.sync 1, 256;" ::: "memory");
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
      // residual stream: the block of group g + 1 is copied global -> shared (cp.async) while group g is finished, the
      // first block while the accumulator is still being produced
      uint8_t* rbuf = epi_smem + kEpiSmemBytes + (warp - 4) * kResidBlockBytes;
      const bool pre_ok = p.resid != nullptr && p.vec_resid;
      auto pre_group = [&](int g) { return pre_ok && g < 4 && g_lo + g < g_hi && n_tile * p.bn + (g_lo + g + 1) * 32 <= p.N; };
      auto pre_ptr = [&](int g) { return p.resid + (row - lane) * p.ld_resid + n_tile * p.bn + (g_lo + g) * 32; };
      if (pre_group(0)) resid_cp_async(rbuf, pre_ptr(0), p.ld_resid, row_mask, lane);
      float sums[4][32];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 32; ++j) sums[g][j] = 0.f;
      if (!padding_tile) {
        for (int chunk = 0; chunk < num_chunks; ++chunk) {
          mbar_wait(tfull_bar(acc), acc_phase);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + acc * kMaxBN + g_lo * 32 + ((uint32_t)(q * 32) << 16);
          const int ng = g_hi - g_lo;
          const uint32_t release = tempty_bar(acc) & kPeerBitMask;   // the leader's barrier, through shared::cluster
          uint32_t v0[32], v1[32];
          if (ng > 0) tmem_ld32(t_addr, v0);
          if (ng > 1) tmem_ld32(t_addr + 32, v1);
          tmem_ld_wait();
          if (ng <= 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(release);
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
            if (lane == 0) mbar_arrive_cluster(release);
#pragma unroll
            for (int j = 0; j < 32; ++j) sums[2][j] += __uint_as_float(v0[j]);
            if (ng > 3) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sums[3][j] += __uint_as_float(v1[j]);
            }
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
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
          if (pre_group(g))
            epilogue_group(p, ec, stage, sums[g], c0, n0, row, valid, row_mask, lane, addp, addl, dotg, rbuf,
                           pre_group(g + 1) ? pre_ptr(g + 1) : nullptr);
          else
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
  cluster_sync_all();   // the peer may still be signalling this CTA's barriers / reading its operand tiles
  if (warp == 2) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, 512);
  }
}

}  // namespace pn
