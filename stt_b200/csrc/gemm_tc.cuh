// Persistent warp-specialised tcgen05 GEMM for the acoustic model's dense layers (kernels K2-K4, K6 of
// SURVEY.md 2.2):   C[M, N] = epilogue( A[M, K] (fp16, K-major)  x  W[N, K]^T (fp16, K-major) + bias[N] )
// with fp32 accumulation in TMEM.  Replaces the TFLite FullyConnected calls of the reference graph
// (training/coqui_stt_training/deepspeech_model.py:66-89 `dense`, :204-263 layer stack).
//
// Roles (one CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A and W into a 4-stage 128B-swizzled smem ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16), accumulators in TMEM,
//               double-buffered so tile i+1's MMAs overlap tile i's epilogue; also owns TMEM alloc/dealloc
//   warps 2..5  epilogue: tcgen05.ld the accumulator (thread = one output row), fused bias + activation
//               (clipped ReLU / none / softmax), convert and store
//
// A-operand addressing modes:
//   kRows2D     A is a plain [M, K] matrix, output row = tile row.
//   kWindows3D  A is the stacked-context view of the MFCC stream: features are stored per utterance as
//               [T + 2*n_context, 32] fp16 (26 coefficients + 6 zero lanes) with n_context literal-zero frames
//               on both sides (native_client/stt.cc:256-261,533), and timestep t's 19-frame window is simply the
//               608 contiguous elements starting at frame t -- so the TMA descriptor uses OVERLAPPING rows
//               (row stride 64 B, row length 608) and no window matrix is ever materialised
//               (replaces StreamingState::pushMfccBuffer / processMfccWindow, stt.cc:272-309).
//               A tile covers t_box timesteps x b_box utterances; rows are emitted time-major (t*B + b).
#pragma once
#include "ptx.cuh"

namespace sttgemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 fp16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;

enum Epilogue { kEpiClipReluF16 = 0, kEpiBiasF32 = 1, kEpiSoftmaxF32 = 2 };
enum AMode { kRows2D = 0, kWindows3D = 1 };

struct GemmParams {
  int M, N, K;           // logical sizes; N is the padded weight row count (multiple of BLOCK_N)
  int n_valid;           // number of real output columns (softmax: classes)
  const float* bias;     // [N]
  void* out;             // fp16 [M, N] | fp32 [M, N] | fp32 probs
  float relu_clip;
  // kWindows3D / time-major bookkeeping
  int B, T;              // utterances, timesteps (rows = T*B, row = t*B + b)
  int b_box, t_box;      // tile = t_box timesteps x b_box utterances, t_box * b_box == BLOCK_M
  // softmax output layout: probs[(b * T_stride + t_offset + t) * n_valid + c]
  int out_T_stride, out_t_offset;
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarrierOffset + 256 + 1024;  // barriers + tmem ptr + 1024B alignment slack
};

template <int BLOCK_N, int STAGES, int EPI, int AMODE>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;

  const int n_tiles_n = p.N / BLOCK_N;
  int n_tiles_m, tiles_b = 1;
  if (AMODE == kWindows3D) {
    tiles_b = (p.B + p.b_box - 1) / p.b_box;
    n_tiles_m = ((p.T + p.t_box - 1) / p.t_box) * tiles_b;
  } else {
    n_tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  }
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                 : (2 * BLOCK_N <= 256) ? 256 : 512;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles_n, n_blk = tile % n_tiles_n;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          if (AMODE == kWindows3D) {
            const int t0 = (m_blk / tiles_b) * p.t_box, b0 = (m_blk % tiles_b) * p.b_box;
            ptx::tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, t0, b0);
          } else {
            ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          }
          ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_f16(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t a_desc = ptx::make_smem_desc_k128(sa);
          const uint64_t b_desc = ptx::make_smem_desc_k128(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes (16 fp16) along K inside the 128B swizzle row: +2 in the >>4 address field
            ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (kb == num_k_blocks - 1) ptx::umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp_idx % 4;          // TMEM lane quarter this warp may read
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles_n, n_blk = tile % n_tiles_n;
      int out_row;
      bool valid;
      if (AMODE == kWindows3D) {
        // TMA fills the box with t (dim 1) varying faster than b (dim 2): tile row r = b_local * t_box + t_local
        const int t = (m_blk / tiles_b) * p.t_box + row_in_tile % p.t_box;
        const int b = (m_blk % tiles_b) * p.b_box + row_in_tile / p.t_box;
        valid = (t < p.T) && (b < p.B);
        out_row = t * p.B + b;
      } else {
        out_row = m_blk * BLOCK_M + row_in_tile;
        valid = out_row < p.M;
      }
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
      if (EPI == kEpiSoftmaxF32 && BLOCK_N > 32) {
        // wide alphabets (up to 255 labels + blank, e.g. the UTF-8 bytes-output models): the row does not fit one
        // 32-column TMEM load, so it is read three times -- maximum, sum of exponentials, normalised write -- with the
        // same arithmetic, in the same column order, as the one-chunk case below
        const int t = out_row / p.B, b = out_row % p.B;
        float* o = static_cast<float*>(p.out) + ((size_t)b * p.out_T_stride + p.out_t_offset + t) * p.n_valid;
        float mx = -3.4e38f, sum = 0.f, inv = 0.f;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll 1
          for (int c = 0; c < BLOCK_N / 32; ++c) {
            if (c * 32 >= p.n_valid) break;
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_addr + c * 32, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = c * 32 + j;
              const float v = __uint_as_float(r[j]) + __ldg(p.bias + col);
              if (col < p.n_valid) {
                if (pass == 0) mx = fmaxf(mx, v);
                else if (pass == 1) sum += expf(v - mx);
                else if (valid) o[col] = expf(v - mx) * inv;
              }
            }
          }
          if (pass == 1) inv = 1.0f / sum;
        }
      } else if (EPI == kEpiSoftmaxF32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_addr, r);
        ptx::tmem_ld_wait();
        if (valid) {
          float v[32];
          float mx = -3.4e38f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r[j]) + __ldg(p.bias + j);
            if (j < p.n_valid) mx = fmaxf(mx, v[j]);
          }
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = (j < p.n_valid) ? expf(v[j] - mx) : 0.f;
            sum += v[j];
          }
          const float inv = 1.0f / sum;
          const int t = out_row / p.B, b = out_row % p.B;
          float* o = static_cast<float*>(p.out) + ((size_t)b * p.out_T_stride + p.out_t_offset + t) * p.n_valid;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < p.n_valid) o[j] = v[j] * inv;
        }
      } else {
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
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace sttgemm
