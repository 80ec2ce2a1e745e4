// Persistent cooperative kernel for the LSTM recurrence (kernel K5 of SURVEY.md 2.2).
//
// Reference arithmetic: tf.nn.rnn_cell.LSTMCell.call, tensorflow/python/keras/layers/legacy_rnn/
// rnn_cell_impl.py:1054-1079, as unrolled by rnn_impl_static_rnn (deepspeech_model.py:144-168):
//     [i, j, f, o] = [x_t, h_{t-1}] @ K + bias ;  c = sigmoid(f) * c + sigmoid(i) * tanh(j) ;  h = sigmoid(o) * tanh(c)
// The x_t @ K[:n_hidden] + bias half has no time dependence and is hoisted into one big GEMM (gemm_tc.cuh,
// kEpiBiasF32) that writes `xw` [T, B, 4*n_cell] fp32.  This kernel runs the T-serial half:
//     gates_t = xw_t + h_{t-1} @ Wh          (Wh = K[n_hidden:], fp16, fp32 accumulate in TMEM)
// for all T steps in ONE launch.  Each CTA owns 16 cells = 64 gate columns for the whole batch; the weight
// matrix is stored gate-interleaved ([cell][i,j,f,o] rows of Wh^T) so one epilogue thread (= one utterance row)
// holds all four gates of its cells after a single TMEM read, and the cell state c never leaves registers for the
// whole utterance.  h_t is written as fp16 into h_all[(t+1)*B + b] which is both the layer-5 GEMM's A operand and,
// one grid barrier later, the next step's A operand via TMA.
//
// v2 (profiles/r01_decoder_v2_lstm_v1.md: v1 spent ~80 % of a step outside the tensor pipe, limited by every CTA
// re-streaming the same 1 MB of h per step from L2, 164 MB/step aggregate):
//   * CTAs form clusters of CS (8 when the grid allows).  The A operand (h_{t-1}, identical for every CTA) is loaded
//     ONCE per cluster: each CTA fetches a 128/CS-row slice of every A tile and TMA-MULTICASTS it into the same smem
//     offset of all CS CTAs; smem stages are released cluster-wide with a multicast tcgen05.commit.
//   * The weight tiles of the first STAGES k-blocks of step t+1 do not depend on h, so they are requested BEFORE the
//     grid barrier and arrive while the epilogue of step t is still running.
//   * 8 epilogue warps (two per TMEM lane quarter, 8 cells each) halve the gate-math latency on the serial path.
//
// Steps are separated by a device-wide arrive/wait on a global counter (cooperative launch guarantees
// co-residency); only the TMA producer thread waits, everybody else blocks on the mbarrier pipeline.
#pragma once
#include "ptx.cuh"

namespace sttlstm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 64;   // 16 cells x 4 gates
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kCellsPerCta = 16;
constexpr int kEpiWarps = 8;
constexpr int kNumThreads = 64 + kEpiWarps * 32;

struct LstmParams {
  int B, T;              // utterances (<= 128 * MT), timesteps
  int n_cell;            // padded cell dim (multiple of 64)
  const float* xw;       // [T, B, 4*n_cell] fp32, gate-interleaved columns, bias already added
  __half* h_all;         // [(T+1)*B, n_cell] fp16; block 0 = initial h, block t+1 = h_t
  float* c_state;        // [B, n_cell] fp32 in/out
  float* h_state;        // [B, n_cell] fp32 out (final h, full precision)
  unsigned int* barrier; // zero-initialised counter
  unsigned long long* prof; // [gridDim.x * 4] instrumentation (cycles): grid-barrier wait, load+MMA span, epilogue, -
  int exact_h;           // h = sigmoid(o) * tanh(c) through the ex2/rcp forms (~1e-7 abs) instead of tanh.approx (2^-11 rel)
};

template <int MT, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;  // per M tile
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = MT * kABytes + kBBytes;
  static constexpr int kBarrierOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarrierOffset + 256 + 1024;
};

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  // 1 - 2/(1+e^{2x}); saturates cleanly for |x| large (e^{2x} -> inf gives 1, -> 0 gives -1)
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}
// One MUFU op (rel. error ~2^-11): used only on the h = sigmoid(o) * tanh(c) path, whose result is rounded to fp16
// (2^-11) anyway; the cell-state update keeps the more accurate ex2/rcp forms so errors do not accumulate in c.
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_mufu(float x) { return fmaf(0.5f, tanh_mufu(0.5f * x), 0.5f); }

// tmap_h: box = {64, 128 / CS} rows of h_all; tmap_wh: box = {64, 64} rows of Wh.
template <int MT, int STAGES, int CS>
__global__ void __launch_bounds__(kNumThreads, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_wh,
               const LstmParams p) {
  using L = SmemLayout<MT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int num_k_blocks = p.n_cell / BLOCK_K;
  const int n0 = blockIdx.x * BLOCK_N;  // first gate column (interleaved order) owned by this CTA
  constexpr uint32_t kTmemCols = (MT * BLOCK_N <= 64) ? 64 : 128;
  constexpr int kSliceRows = BLOCK_M / CS;
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
  const uint32_t crank = (CS > 1) ? ptx::cluster_ctarank() : 0u;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_h);
    ptx::prefetch_tmap(&tmap_wh);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], CS);  // one (multicast) commit from every CTA of the cluster
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync();  // every CTA's barriers are initialised before any remote arrive / multicast
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer (+ grid-barrier waiter) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int pre = num_k_blocks < STAGES ? num_k_blocks : STAGES;
      unsigned long long prof_wait = 0;
      for (int t = 0; t < p.T; ++t) {
        // weight tiles of the first `pre` k-blocks: independent of h, requested before the grid barrier
        int st2 = stage;
        uint32_t ph2 = phase;
        for (int kb = 0; kb < pre; ++kb) {
          ptx::mbar_wait(&empty_bar[st2], ph2 ^ 1);
          ptx::mbar_expect_tx(&full_bar[st2], L::kStageBytes);
          ptx::tma_load_2d(smem + st2 * L::kStageBytes + MT * L::kABytes, &tmap_wh, &full_bar[st2], kb * BLOCK_K, n0);
          if (++st2 == STAGES) { st2 = 0; ph2 ^= 1; }
        }
        if (t > 0) {
          const long long w0 = clock64();
          const unsigned int target = (unsigned int)t * gridDim.x;
          unsigned int seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.barrier) : "memory");
          } while (seen < target);
          ptx::fence_proxy_async();  // other CTAs' generic-proxy stores of h_{t-1} -> async-proxy (TMA) reads
          prof_wait += (unsigned long long)(clock64() - w0);
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          uint8_t* sa = smem + stage * L::kStageBytes;
          if (kb >= pre) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
            ptx::tma_load_2d(sa + MT * L::kABytes, &tmap_wh, &full_bar[stage], kb * BLOCK_K, n0);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint8_t* dst = sa + mt * L::kABytes + crank * (kSliceRows * BLOCK_K * 2);
            const int row = t * p.B + mt * BLOCK_M + (int)crank * kSliceRows;
            if (CS > 1) ptx::tma_load_2d_mcast(dst, &tmap_h, &full_bar[stage], kb * BLOCK_K, row, kMask);
            else ptx::tma_load_2d(dst, &tmap_h, &full_bar[stage], kb * BLOCK_K, row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (p.prof) p.prof[blockIdx.x * 4 + 0] = prof_wait;
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_f16(BLOCK_M, BLOCK_N);
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
          const uint32_t sb = sa + MT * L::kABytes;
          const uint64_t b_desc = ptx::make_smem_desc_k128(sb);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t a_desc = ptx::make_smem_desc_k128(sa + mt * L::kABytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              ptx::umma_f16(tmem_base + mt * BLOCK_N, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          if (CS > 1) ptx::umma_commit_mcast(&empty_bar[stage], kMask);
          else ptx::umma_commit(&empty_bar[stage]);
          if (kb == num_k_blocks - 1) ptx::umma_commit(tmem_full_bar);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      prof_mma += (unsigned long long)(clock64() - m0);
    }
    if (lane == 0 && p.prof) p.prof[blockIdx.x * 4 + 1] = prof_mma;
  } else {
    // ===================== epilogue warps: gates -> (c, h) =====================
    const int ew = warp_idx - 2;              // 0..7
    const int quarter = warp_idx % 4;         // TMEM lane quarter this warp may read
    const int chalf = ew / 4;                 // which 8 of the CTA's 16 cells (32 of its 64 gate columns)
    const int row_in_tile = quarter * 32 + lane;
    const int cell0 = blockIdx.x * kCellsPerCta + chalf * 8;
    float c_reg[MT][8];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int b = mt * BLOCK_M + row_in_tile;
#pragma unroll
      for (int j = 0; j < 8; ++j) c_reg[mt][j] = (b < p.B) ? p.c_state[(size_t)b * p.n_cell + cell0 + j] : 0.f;
    }
    const size_t xw_row = (size_t)4 * p.n_cell;
    const int ncol0 = n0 + chalf * 32;
    unsigned long long prof_epi = 0;
    for (int t = 0; t < p.T; ++t) {
      // this step's xw values are loaded while the MMAs run; next step's rows are pulled towards L2
      float4 xv[MT][8];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int b = mt * BLOCK_M + row_in_tile;
        const bool valid = b < p.B;
        const float4* xr = reinterpret_cast<const float4*>(p.xw + ((size_t)t * p.B + (valid ? b : 0)) * xw_row + ncol0);
#pragma unroll
        for (int q = 0; q < 8; ++q) xv[mt][q] = valid ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && t + 1 < p.T)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.xw + ((size_t)(t + 1) * p.B + b) * xw_row + ncol0));
      }
      ptx::mbar_wait(tmem_full_bar, t & 1);
      const long long e0 = clock64();
      ptx::tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int b = mt * BLOCK_M + row_in_tile;
        const bool valid = b < p.B;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + mt * BLOCK_N + chalf * 32 + ((uint32_t)(quarter * 32) << 16), r);
        ptx::tmem_ld_wait();
        float h_last[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // one cell per float4: (i, j, f, o)
          const float gi = __uint_as_float(r[q * 4 + 0]) + xv[mt][q].x;
          const float gj = __uint_as_float(r[q * 4 + 1]) + xv[mt][q].y;
          const float gf = __uint_as_float(r[q * 4 + 2]) + xv[mt][q].z;
          const float go = __uint_as_float(r[q * 4 + 3]) + xv[mt][q].w;
          const float cn = sigmoid_fast(gf) * c_reg[mt][q] + sigmoid_fast(gi) * tanh_fast(gj);
          c_reg[mt][q] = cn;
          h_last[q] = (p.exact_h ? sigmoid_fast(go) * tanh_fast(cn) : sigmoid_mufu(go) * tanh_mufu(cn));
        }
        if (valid) {
          uint32_t hpk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __half2 h2 = __floats2half2_rn(h_last[2 * e], h_last[2 * e + 1]);
            hpk[e] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(p.h_all + ((size_t)(t + 1) * p.B + b) * p.n_cell + cell0) =
              make_uint4(hpk[0], hpk[1], hpk[2], hpk[3]);
          if (t == p.T - 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              p.c_state[(size_t)b * p.n_cell + cell0 + j] = c_reg[mt][j];
              p.h_state[(size_t)b * p.n_cell + cell0 + j] = h_last[j];
            }
          }
        }
      }
      // publish h_t: all epilogue threads' stores, then one release-arrive on the grid counter
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
  if (CS > 1) ptx::cluster_sync();  // nobody exits while a peer may still multicast into / arrive on its smem
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace sttlstm
