// Bring-up probe: where does tcgen05.mma.cta_group::2 put D for a given M?  Each CTA of a pair stages its half of A
// (M/2 rows x 64) and of B (64 of the 128 N rows), the leader issues one K=64 block, and both CTAs dump their 128 TMEM
// lanes x 128 columns.  With D[i][n] = (i+1)*1024 + (n+1) the host reads the (row, column) of every TMEM cell.
#pragma once
#include "lstm2_tc.cuh"

namespace sttprobe {

__global__ void __launch_bounds__(128, 1)
probe_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;              // up to 128 rows x 128 B
  uint8_t* sb = smem + 16384;      // 64 rows x 128 B
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint64_t* done_bar = full_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t crank = ptx::cluster_ctarank();
  const int rows_a = M / 2;
  if (threadIdx.x == 0) {
    ptx::mbar_init(full_bar, 1);
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_ptr)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t leader_full = sttlstm::mapa_u32(ptx::smem_u32(full_bar), 0);
    if (crank == 0) ptx::mbar_expect_tx(full_bar, 2u * (uint32_t)(rows_a * 128 + 64 * 128));
    sttlstm::tma_load_2d_pair(sa, &tmap_a, leader_full, 0, (int)crank * rows_a);
    sttlstm::tma_load_2d_pair(sb, &tmap_b, leader_full, 0, (int)crank * 64);
    if (crank == 0) {
      ptx::mbar_wait(full_bar, 0);
      ptx::tc_fence_after();
      const uint32_t idesc = ptx::make_idesc_f16((uint32_t)M, 128);
      const uint64_t a_desc = ptx::make_smem_desc_k128(ptx::smem_u32(sa)), b_desc = ptx::make_smem_desc_k128(ptx::smem_u32(sb));
      for (int k = 0; k < 4; ++k) sttlstm::umma_f16_pair(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
      sttlstm::umma_commit_pair(done_bar);
    }
  }
  __syncwarp();
  ptx::mbar_wait(done_bar, 0);
  ptx::tc_fence_after();
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tmem_base + cb * 32 + ((uint32_t)(warp * 32) << 16), r);
    ptx::tmem_ld_wait();
    for (int q = 0; q < 32; ++q) out[((size_t)crank * 128 + warp * 32 + lane) * 128 + cb * 32 + q] = __uint_as_float(r[q]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

}  // namespace sttprobe
