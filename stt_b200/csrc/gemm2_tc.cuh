// CTA-PAIR variant of the dense-layer GEMM (gemm_tc.cuh): one 256 x 256 output tile per pair of SMs with
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16).
//
// Why: the one-CTA kernel moves 48 KB (A 128 x 64 + W 256 x 64, fp16) through shared memory per 512 tensor cycles, i.e.
// ~94 B/clk per SM, and the chip's L2 delivers ~43 B/clk per SM when all 148 pull at once -- the 128 x 256 kernel sits
// at the L2 ceiling (0.80-0.88 of the cuBLAS rate), not at the tensor pipe's.  With a pair each CTA stages only ITS 128
// rows of A and ITS half (128 rows) of the weight tile: 32 KB per 512 tensor cycles for the same work per SM.
//
//   pair p (CTAs 2p, 2p+1), tile (m_blk, n_blk): CTA rank r stages A rows m_blk*256 + r*128 .. +127 and W rows
//   n_blk*256 + r*128 .. +127; the leader (rank 0) issues the MMAs; D row i lives in TMEM lane i % 128 of CTA i / 128,
//   so each CTA's four epilogue warps handle its own 128 output rows exactly as in the one-CTA kernel.
//   Accumulators are double buffered (2 x 256 TMEM columns): tile i+1's MMAs overlap tile i's epilogue.
#pragma once
#include "gemm_tc.cuh"

namespace sttgemm {

template <int STAGES>
struct Smem2Layout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;   // this CTA's 128 rows of A
  static constexpr int kBBytes = 128 * BLOCK_K * 2;       // this CTA's 128 weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;   // 32 KB
  static constexpr int kBarrierOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarrierOffset + 256 + 1024;
};

namespace pair {
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
// local smem destination, completion signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  uint32_t a = acc ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(a)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on this offset in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
}  // namespace pair

// tmap_a: box {64, 128} rows of A; tmap_b: box {64, 128} rows of W.  Launch with cluster (2,1,1), an even grid.
// Epilogues: kEpiClipReluF16, kEpiBiasF32 (the softmax layer stays on the one-CTA kernel).
template <int STAGES, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using L = Smem2Layout<STAGES>;
  constexpr int BLOCK_N = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;          // per CTA: its producer waits here (commit arrives in both CTAs)
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2] per CTA (commit arrives in both)
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2] used in the leader: 8 arrivals = 4 epilogue warps x 2 CTAs
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const uint32_t crank = ptx::cluster_ctarank();   // 0 = leader
  const int pair_idx = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles_n = p.N / BLOCK_N;
  const int n_tiles_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  constexpr uint32_t kTmemCols = 512;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], 8);
    }
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
    // ===================== TMA producer, one per CTA =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair_idx; tile < n_tiles; tile += n_pairs) {
        const int m_blk = tile / n_tiles_n, n_blk = tile % n_tiles_n;
        const int a_row = m_blk * 2 * BLOCK_M + (int)crank * BLOCK_M;
        const int b_row = n_blk * BLOCK_N + (int)crank * 128;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          const uint32_t leader_full = pair::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
          if (crank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);   // both CTAs' halves
          pair::tma_load_2d(sa, &tmap_a, leader_full, kb * BLOCK_K, a_row);
          pair::tma_load_2d(sa + L::kABytes, &tmap_b, leader_full, kb * BLOCK_K, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer: leader CTA only =====================
    if (crank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(256, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair_idx; tile < n_tiles; tile += n_pairs) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
            const uint64_t a_desc = ptx::make_smem_desc_k128(sa);
            const uint64_t b_desc = ptx::make_smem_desc_k128(sa + L::kABytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              pair::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            pair::umma_commit(&empty_bar[stage]);   // frees the smem slot in both CTAs once these MMAs retire
            if (kb == num_k_blocks - 1) pair::umma_commit(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps: this CTA's 128 rows of the tile =====================
    const int quarter = warp_idx % 4;          // TMEM lane quarter this warp may read
    const int row_in_cta = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair_idx; tile < n_tiles; tile += n_pairs) {
      const int m_blk = tile / n_tiles_n, n_blk = tile % n_tiles_n;
      const int out_row = m_blk * 2 * BLOCK_M + (int)crank * BLOCK_M + row_in_cta;
      const bool valid = out_row < p.M;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_addr + c * 32, r);
        ptx::tmem_ld_wait();
        if (valid) {
          const int col0 = n_blk * BLOCK_N + c * 32;
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + col0);
          if (EPI == kEpiClipReluF16) {
            __half* o = static_cast<__half*>(p.out) + (size_t)out_row * p.N + col0;
            uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float x[8];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 bv = __ldg(bias4 + q * 2 + h);
                x[h * 4 + 0] = __uint_as_float(r[q * 8 + h * 4 + 0]) + bv.x;
                x[h * 4 + 1] = __uint_as_float(r[q * 8 + h * 4 + 1]) + bv.y;
                x[h * 4 + 2] = __uint_as_float(r[q * 8 + h * 4 + 2]) + bv.z;
                x[h * 4 + 3] = __uint_as_float(r[q * 8 + h * 4 + 3]) + bv.w;
              }
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = fminf(fmaxf(x[2 * e], 0.f), p.relu_clip);
                const float b2 = fminf(fmaxf(x[2 * e + 1], 0.f), p.relu_clip);
                const __half2 h2 = __floats2half2_rn(a, b2);
                pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              o4[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else {  // kEpiBiasF32
            float* o = static_cast<float*>(p.out) + (size_t)out_row * p.N + col0;
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bv = __ldg(bias4 + q);
              o4[q] = make_float4(__uint_as_float(r[q * 4 + 0]) + bv.x, __uint_as_float(r[q * 4 + 1]) + bv.y,
                                  __uint_as_float(r[q * 4 + 2]) + bv.z, __uint_as_float(r[q * 4 + 3]) + bv.w);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) pair::mbar_arrive_cluster(pair::mapa_u32(ptx::smem_u32(&tmem_empty_bar[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace sttgemm
