// LSTM recurrence, CTA-PAIR variant (tcgen05 cta_group::2) for batches of 129..256 utterances.
//
// Same arithmetic and data layout as lstm_tc.cuh (see there for the reference citations).  Why a second kernel:
// profiles/r01_decoder_v2_lstm_v1.md + the role-cycle instrumentation show the single-CTA kernel is INGEST bound --
// every step each CTA must pull all of h_{t-1} (256 x 2048 fp16 = 1 MB) plus its 256 KB weight slice through a
// pipeline whose depth is capped by shared memory (bytes in flight / L2 latency ~ 60-70 B/clk/SM), i.e. ~22k cycles
// per step against 8k cycles of tensor work.  With cta_group::2 a PAIR of CTAs computes one 256 x 128 tile:
// each CTA stages only ITS 128 rows of h (512 KB / step) and ITS half of the pair's weight slice (256 KB), so the
// bytes per CTA per step drop from 1280 KB to 768 KB for the same MMA work per SM.
//
//   pair p (CTAs 2p, 2p+1) owns cells [32p, 32p+32) = gate columns [128p, 128p+128) for all utterances
//   CTA rank r stages  A: rows r*128..r*128+127 of h_{t-1}      B: rows 128p + r*64 .. +63 of Wh (interleaved)
//   the leader (rank 0) issues tcgen05.mma.cta_group::2 (M=256, N=128, K=16); D row i lives in TMEM lane i%128 of
//   CTA i/128, columns 0..127, so each CTA's epilogue handles its own 128 utterances x 32 cells.
#pragma once
#include "lstm_tc.cuh"

namespace sttlstm {

constexpr int kPPCounters = 32;   // grid-barrier counters (per group), 128 B apart, one per lane of the producer warp
constexpr int kPairN = 128;        // gate columns per CTA pair
constexpr int kPairCells = 32;
constexpr int kPairStages = 8;

struct PairSmem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;       // this CTA's 128 rows
  static constexpr int kBBytes = (kPairN / 2) * BLOCK_K * 2;  // this CTA's 64 weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;       // 24 KB
  static constexpr int kBarrierOffset = kPairStages * kStageBytes;
  static constexpr int kTotal = kBarrierOffset + 256 + 1024;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
// local smem destination, completion signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  uint32_t a = acc ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(a)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on this offset in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0,
                                                 int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// tmap_h: box {64, 128} rows of h_all; tmap_wh: box {64, 64} rows of Wh.  Launch: cooperative, cluster (2,1,1).
__global__ void __launch_bounds__(kNumThreads, 1)
lstm_pair_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_wh,
                 const LstmParams p) {
  using L = PairSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);  // used in the leader only
  uint64_t* empty_bar = full_bar + kPairStages;
  uint64_t* tmem_full_bar = empty_bar + kPairStages;  // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int num_k_blocks = p.n_cell / BLOCK_K;
  const uint32_t crank = ptx::cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int n0 = pair * kPairN;                        // first gate column of the pair
  constexpr uint32_t kTmemCols = 128;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_h);
    ptx::prefetch_tmap(&tmap_wh);
    for (int i = 0; i < kPairStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_ptr_smem)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer (+ grid-barrier waiter), one per CTA =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int pre = num_k_blocks < kPairStages ? num_k_blocks : kPairStages;
      unsigned long long prof_wait = 0;
      for (int t = 0; t < p.T; ++t) {
        int st2 = stage;
        uint32_t ph2 = phase;
        for (int kb = 0; kb < pre; ++kb) {  // weight tiles first: they do not depend on h_{t-1}
          ptx::mbar_wait(&empty_bar[st2], ph2 ^ 1);
          const uint32_t leader_full = mapa_u32(ptx::smem_u32(&full_bar[st2]), 0);
          if (crank == 0) ptx::mbar_expect_tx(&full_bar[st2], 2 * L::kStageBytes);  // both CTAs' halves
          tma_load_2d_pair(smem + st2 * L::kStageBytes + L::kABytes, &tmap_wh, leader_full, kb * BLOCK_K,
                           n0 + (int)crank * (kPairN / 2));
          if (++st2 == kPairStages) { st2 = 0; ph2 ^= 1; }
        }
        if (t > 0) {
          const long long w0 = clock64();
          const unsigned int target = (unsigned int)t * gridDim.x;
          unsigned int seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.barrier) : "memory");
          } while (seen < target);
          ptx::fence_proxy_async();
          prof_wait += (unsigned long long)(clock64() - w0);
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          uint8_t* sa = smem + stage * L::kStageBytes;
          const uint32_t leader_full = mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
          if (kb >= pre) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (crank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            tma_load_2d_pair(sa + L::kABytes, &tmap_wh, leader_full, kb * BLOCK_K, n0 + (int)crank * (kPairN / 2));
          }
          tma_load_2d_pair(sa, &tmap_h, leader_full, kb * BLOCK_K, t * p.B + (int)crank * BLOCK_M);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.prof) p.prof[blockIdx.x * 4 + 0] = prof_wait;
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer: leader CTA only =====================
    if (crank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(256, kPairN);
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long prof_mma = 0;
      long long m0 = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          if (kb == 0) m0 = clock64();
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
            const uint64_t a_desc = ptx::make_smem_desc_k128(sa);
            const uint64_t b_desc = ptx::make_smem_desc_k128(sa + L::kABytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_f16_pair(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_k_blocks - 1) umma_commit_pair(tmem_full_bar);
          }
          __syncwarp();
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        prof_mma += (unsigned long long)(clock64() - m0);
      }
      if (lane == 0 && p.prof) p.prof[blockIdx.x * 4 + 1] = prof_mma;
    }
  } else {
    // ===================== epilogue warps: this CTA's 128 utterances x the pair's 32 cells =====================
    const int ew = warp_idx - 2;              // 0..7
    const int quarter = warp_idx % 4;
    const int chalf = ew / 4;                 // cells [16*chalf, 16*chalf+16) of the pair
    const int b = (int)crank * BLOCK_M + quarter * 32 + lane;   // utterance row
    const bool valid = b < p.B;
    const int cell0 = pair * kPairCells + chalf * 16;
    const int ncol0 = n0 + chalf * 64;
    float c_reg[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) c_reg[j] = valid ? p.c_state[(size_t)b * p.n_cell + cell0 + j] : 0.f;
    const size_t xw_row = (size_t)4 * p.n_cell;
    unsigned long long prof_epi = 0;
    for (int t = 0; t < p.T; ++t) {
      float4 xv[16];
      {
        const float4* xr = reinterpret_cast<const float4*>(p.xw + ((size_t)t * p.B + (valid ? b : 0)) * xw_row + ncol0);
#pragma unroll
        for (int q = 0; q < 16; ++q) xv[q] = valid ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && t + 1 < p.T) {
          const float* nx = p.xw + ((size_t)(t + 1) * p.B + b) * xw_row + ncol0;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32));
        }
      }
      ptx::mbar_wait(tmem_full_bar, t & 1);
      const long long e0 = clock64();
      ptx::tc_fence_after();
      float h_last[16];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + chalf * 64 + half * 32 + ((uint32_t)(quarter * 32) << 16), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int cell = half * 8 + q;
          const float gi = __uint_as_float(r[q * 4 + 0]) + xv[cell].x;
          const float gj = __uint_as_float(r[q * 4 + 1]) + xv[cell].y;
          const float gf = __uint_as_float(r[q * 4 + 2]) + xv[cell].z;
          const float go = __uint_as_float(r[q * 4 + 3]) + xv[cell].w;
          const float cn = sigmoid_fast(gf) * c_reg[cell] + sigmoid_fast(gi) * tanh_fast(gj);
          c_reg[cell] = cn;
          h_last[cell] = (p.exact_h ? sigmoid_fast(go) * tanh_fast(cn) : sigmoid_mufu(go) * tanh_mufu(cn));
        }
      }
      if (valid) {
        uint32_t hpk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const __half2 h2 = __floats2half2_rn(h_last[2 * e], h_last[2 * e + 1]);
          hpk[e] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        uint4* ho = reinterpret_cast<uint4*>(p.h_all + ((size_t)(t + 1) * p.B + b) * p.n_cell + cell0);
        ho[0] = make_uint4(hpk[0], hpk[1], hpk[2], hpk[3]);
        ho[1] = make_uint4(hpk[4], hpk[5], hpk[6], hpk[7]);
        if (t == p.T - 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            p.c_state[(size_t)b * p.n_cell + cell0 + j] = c_reg[j];
            p.h_state[(size_t)b * p.n_cell + cell0 + j] = h_last[j];
          }
        }
      }
      ptx::tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (threadIdx.x == 64) {
        __threadfence();
        ptx::fence_proxy_async();
        atomicAdd(p.barrier, 1u);
        prof_epi += (unsigned long long)(clock64() - e0);
      }
    }
    if (threadIdx.x == 64 && p.prof) p.prof[blockIdx.x * 4 + 2] = prof_epi;
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}


// ============================================================================================================
// PING-PONG variant: the batch is two independent groups of <= 128 utterances (an LSTM couples nothing across
// utterances), and each CTA pair alternates between them with M = 128 MMAs: while group g's epilogue and grid barrier
// run -- the serial tail that the kernel above cannot hide -- group g^1's h / weight tiles stream in and multiply.
//   * tcgen05.mma.cta_group::2 with M = 128: each CTA stages 64 rows of h and 64 of the pair's 128 weight rows, i.e.
//     16 KB per K block.  D (probed on the device, tests/native/probe_pair.py): CTA r holds rows 64r..64r+63;
//     TMEM lanes 0-63 carry gate columns 0-63 and lanes 64-127 carry gate columns 64-127 of the same rows, in 64 TMEM
//     columns.  Two accumulators (one per group) = 128 columns.
//   * bytes per CTA and step: 2 groups x (256 KB h + 256 KB weights) = 1 MB (the M = 256 kernel moves 768 KB) but
//     they arrive while the other group is busy, so a step costs ~ max(loads, epilogue + barrier) instead of their sum.
//   * one grid barrier counter per group.
//   * ONE TMA request brings KB consecutive K blocks of an operand (3-D tensor map: 64 columns x rows x K block).  The
//     producer thread needs ~260 cycles per request whatever its size (measured: 8 KB, 16 KB and 48 KB requests all
//     moved at that rate), so small requests, not L2, were what bounded the kernels above at 30-45 B/clk per SM.
constexpr int kPPEpiWarps = 16;
constexpr int kPPThreads = 64 + kPPEpiWarps * 32;
template <int KB, int STAGES>
struct PPSmem {
  static constexpr int kTile = 64 * BLOCK_K * 2;       // 64 rows of one K block: 8 KB
  static constexpr int kABytes = KB * kTile;           // this CTA's 64 utterance rows, KB K blocks
  static constexpr int kBBytes = KB * kTile;           // this CTA's 64 weight rows, KB K blocks
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarrierOffset + 512 + 1024;
};

// The same load delivered to every CTA of `cta_mask` (same shared-memory offset in each); with cta_group::2 the bytes
// are accounted on the mbarrier at this offset in the EVEN CTA of each destination's pair, which is why the barrier
// address has the peer bit cleared (the convention of CUTLASS's SM100_TMA_2SM_LOAD_MULTICAST).
__device__ __forceinline__ void tma_load_3d_pair_mcast(void* smem_dst, const CUtensorMap* m, uint32_t mbar_addr, int c0, int c1,
                                                       int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_addr), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_mask(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the bit that tells the two CTAs of a pair apart in a shared::cluster address

// tmap_h / tmap_wh: 3-D maps {64 columns, rows, K block} with box {64, 64, KB}.  Launch: cooperative, cluster (CS,1,1).
// p.barrier: 2 groups x kPPCounters counters, 128 B apart, zero-initialised.
//
// CS = 4 / 8 (needs KB = 4; tmap_h then has box {64, 64, KB * 2 / CS}): the kernel is bound by what L2 can deliver -- 128 CTAs x 1 MB
// per step = 128 MB / 21.5 k cycles = 5.9 KB/clk, the measured ceiling of the L2 slices -- and half of that is the SAME h
// rows fetched by all 64 pairs.  Four pairs share a cluster: each CTA fetches ONE of the four K blocks of its 64 rows
// and multicasts it to the four CTAs of its parity, so a step reads h 16 times instead of 64 (80 MB instead of 128).
// A stage is refilled when all four pairs have consumed it (every pair leader's commit arrives in all eight CTAs).
template <int KB, int STAGES, int CS = 2>
__global__ void __launch_bounds__(kPPThreads, 1)
lstm_pp_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_wh,
               const LstmParams p) {
  static_assert(CS == 2 || ((CS == 4 || CS == 8) && KB == 4), "multicast variant: KB * 2 / CS K blocks per CTA");
  constexpr int kSlice = KB * 2 / CS;   // K blocks of the h tile this CTA fetches for its parity (multicast variant)
  using L = PPSmem<KB, STAGES>;
  constexpr int kPPStages = STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);  // used in the leader only
  uint64_t* empty_bar = full_bar + kPPStages;
  uint64_t* tmem_full_bar = empty_bar + kPPStages;  // [2], one per group
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 2);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int num_k_blocks = (p.n_cell / BLOCK_K + KB - 1) / KB;   // in units of KB K blocks (a short tail is zero-filled by TMA)
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t crank = rank & 1u;                    // 0 = leader of its pair
  const uint32_t qpair = rank >> 1;                    // pair within the cluster
  const int pair = blockIdx.x >> 1;
  const int n0 = pair * kPairN;                        // first gate column of the pair
  const int n_groups = (p.B + 127) / 128;              // 1 or 2
  const unsigned int n_ctr = (gridDim.x % kPPCounters == 0) ? kPPCounters : 1;
  constexpr uint32_t kTmemCols = 128;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_h);
    ptx::prefetch_tmap(&tmap_wh);
    for (int i = 0; i < kPPStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], CS / 2);   // one commit per pair of the cluster
    }
    ptx::mbar_init(&tmem_full_bar[0], 1);
    ptx::mbar_init(&tmem_full_bar[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_ptr_smem)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer warp: lane 0 issues, all lanes watch the grid barrier =====================
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long prof_wait = 0, prof_pre = 0, prof_main = 0;
    for (int t = 0; t < p.T; ++t) {
      for (int g = 0; g < n_groups; ++g) {
        __syncwarp();
        if (t > 0) {
          // Arrivals are spread over n_ctr counters (same-address L2 atomics serialise at ~27 cycles each); lane c
          // polls counter c with an acquire load, so one poll of the whole barrier is a single round trip.
          const long long w0 = clock64();
          const unsigned int target = (unsigned int)t * (gridDim.x / n_ctr);
          const unsigned int* ctr = p.barrier + g * (kPPCounters * 32) + lane * 32;
          bool all_in;
          do {
            unsigned int seen = target;
            if ((unsigned int)lane < n_ctr) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
            all_in = __all_sync(0xffffffffu, seen >= target);
          } while (!all_in);
          prof_wait += (unsigned long long)(clock64() - w0);
        }
        if (lane == 0) {
          ptx::fence_proxy_async();
          const long long q1 = clock64();
          const int row0 = t * p.B + g * 128 + (int)crank * 64;
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            uint8_t* sa = smem + stage * L::kStageBytes;
            const uint32_t leader_full = ptx::smem_u32(&full_bar[stage]) & kPeerBitMask;   // the pair leader's barrier
            // No weight prefetch ahead of the barrier here: the other group's MMAs keep the ring busy until its last
            // stages drain, and by then this group's barrier has normally completed.
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (crank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);  // both CTAs' halves
            tma_load_3d_pair(sa + L::kABytes, &tmap_wh, leader_full, 0, n0 + (int)crank * 64, kb * KB);
            if (CS == 2) {
              tma_load_3d_pair(sa, &tmap_h, leader_full, 0, row0, kb * KB);
            } else {
              // this CTA's K block of the 64 rows, for the four CTAs of its parity in the cluster
              constexpr uint16_t kEven = (uint16_t)(0x55u & ((1u << CS) - 1u));
              tma_load_3d_pair_mcast(sa + qpair * kSlice * L::kTile, &tmap_h, leader_full, 0, row0, kb * KB + (int)qpair * kSlice,
                                     crank ? (uint16_t)(kEven << 1) : kEven);
            }
            if (++stage == kPPStages) { stage = 0; phase ^= 1; }
          }
          prof_main += (unsigned long long)(clock64() - q1);
        }
        __syncwarp();
      }
    }
    if (lane == 0 && p.prof) {
      p.prof[blockIdx.x * 4 + 0] = prof_wait;
      p.prof[2048 + blockIdx.x * 4 + 0] = prof_pre;
      p.prof[2048 + blockIdx.x * 4 + 1] = prof_main;
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer: leader CTA only =====================
    if (crank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, kPairN);
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long prof_mma = 0, prof_first = 0;
      long long m0 = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int g = 0; g < n_groups; ++g) {
          const uint32_t acc = tmem_base + (uint32_t)g * 64;
          const long long f0 = clock64();
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            if (kb == 0) { m0 = clock64(); prof_first += (unsigned long long)(m0 - f0); }
            ptx::tc_fence_after();
            if (lane == 0) {
              const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
#pragma unroll
              for (int i = 0; i < KB; ++i) {
                const uint64_t a_desc = ptx::make_smem_desc_k128(sa + i * L::kTile);
                const uint64_t b_desc = ptx::make_smem_desc_k128(sa + L::kABytes + i * L::kTile);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                  umma_f16_pair(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | i | k) != 0);
              }
              if (CS == 2) {
                umma_commit_pair(&empty_bar[stage]);
                if (kb == num_k_blocks - 1) umma_commit_pair(&tmem_full_bar[g]);
              } else {
                umma_commit_pair_mask(&empty_bar[stage], (uint16_t)((1u << CS) - 1u));   // every CTA's producer counts this pair
                if (kb == num_k_blocks - 1) umma_commit_pair_mask(&tmem_full_bar[g], (uint16_t)(3u << (2 * qpair)));
              }
            }
            __syncwarp();
            if (++stage == kPPStages) { stage = 0; phase ^= 1; }
          }
          prof_mma += (unsigned long long)(clock64() - m0);
        }
      }
      if (lane == 0 && p.prof) { p.prof[blockIdx.x * 4 + 1] = prof_mma; p.prof[2048 + blockIdx.x * 4 + 2] = prof_first; }
    }
  } else {
    // ===================== 16 epilogue warps: 64 utterances x the pair's 32 cells, per group =====================
    const int ew = warp_idx - 2;              // 0..15
    const int quarter = warp_idx % 4;         // TMEM lane quarter this warp may read
    const int sub = ew / 4;                   // TMEM columns [16*sub, 16*sub+16) of the accumulator = 4 cells
    const int row_in_cta = (quarter & 1) * 32 + lane;                  // lanes 64-127 repeat rows 0-63 ...
    const int cell_in_pair = (quarter >> 1) * 16 + sub * 4;            // ... for gate columns 64-127
    const int cell0 = pair * kPairCells + cell_in_pair;
    const int ncol0 = n0 + cell_in_pair * 4;
    const size_t xw_row = (size_t)4 * p.n_cell;
    float c_reg[2][4];
    int bidx[2];
    bool valid[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      bidx[g] = g * 128 + (int)crank * 64 + row_in_cta;
      valid[g] = g < n_groups && bidx[g] < p.B;
#pragma unroll
      for (int j = 0; j < 4; ++j) c_reg[g][j] = valid[g] ? p.c_state[(size_t)bidx[g] * p.n_cell + cell0 + j] : 0.f;
    }
    unsigned long long prof_epi = 0, prof_tw = 0;
    for (int t = 0; t < p.T; ++t) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g >= n_groups) break;
        const int b = bidx[g];
        const bool ok = valid[g];
        float4 xv[4];
        {
          const float4* xr = reinterpret_cast<const float4*>(p.xw + ((size_t)t * p.B + (ok ? b : 0)) * xw_row + ncol0);
#pragma unroll
          for (int q = 0; q < 4; ++q) xv[q] = ok ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && t + 1 < p.T && sub == 0) {
            const float* nx = p.xw + ((size_t)(t + 1) * p.B + b) * xw_row + ncol0;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32));
          }
        }
        const long long w8 = clock64();
        ptx::mbar_wait(&tmem_full_bar[g], t & 1);
        const long long e0 = clock64();
        prof_tw += (unsigned long long)(e0 - w8);
        ptx::tc_fence_after();
        float h_last[4];
        {
          uint32_t r[16];
          ptx::tmem_ld_32x16(tmem_base + (uint32_t)g * 64 + sub * 16 + ((uint32_t)(quarter * 32) << 16), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float gi = __uint_as_float(r[q * 4 + 0]) + xv[q].x;
            const float gj = __uint_as_float(r[q * 4 + 1]) + xv[q].y;
            const float gf = __uint_as_float(r[q * 4 + 2]) + xv[q].z;
            const float go = __uint_as_float(r[q * 4 + 3]) + xv[q].w;
            const float cn = sigmoid_fast(gf) * c_reg[g][q] + sigmoid_fast(gi) * tanh_fast(gj);
            c_reg[g][q] = cn;
            h_last[q] = (p.exact_h ? sigmoid_fast(go) * tanh_fast(cn) : sigmoid_mufu(go) * tanh_mufu(cn));
          }
        }
        if (ok) {
          const __half2 h01 = __floats2half2_rn(h_last[0], h_last[1]), h23 = __floats2half2_rn(h_last[2], h_last[3]);
          *reinterpret_cast<uint2*>(p.h_all + ((size_t)(t + 1) * p.B + b) * p.n_cell + cell0) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          if (t == p.T - 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              p.c_state[(size_t)b * p.n_cell + cell0 + j] = c_reg[g][j];
              p.h_state[(size_t)b * p.n_cell + cell0 + j] = h_last[j];
            }
          }
        }
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(kPPEpiWarps * 32) : "memory");
        if (threadIdx.x == 64) {
          __threadfence();
          ptx::fence_proxy_async();
          atomicAdd(p.barrier + g * (kPPCounters * 32) + (blockIdx.x % n_ctr) * 32, 1u);
          prof_epi += (unsigned long long)(clock64() - e0);
        }
      }
    }
    if (threadIdx.x == 64 && p.prof) { p.prof[blockIdx.x * 4 + 2] = prof_epi; p.prof[2048 + blockIdx.x * 4 + 3] = prof_tw; }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace sttlstm
