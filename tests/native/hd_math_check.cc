// Exhaustive pin of stt_b200/csrc/hd_math.h against the host libm (glibc logf/expf), the functions the
// reference decoder calls (decoder_utils.h:46-53, ctc_beam_search_decoder.cpp:355).
// usage: hd_math_check [stride]   (stride 1 = every float; tests use a larger stride for speed)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
#include "../../stt_b200/csrc/hd_math.h"

int main(int argc, char** argv) {
  uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
  unsigned nt = std::thread::hardware_concurrency();
  if (!nt) nt = 4;
  std::atomic<uint64_t> bad_log{0}, bad_exp{0}, n{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([&, t]() {
      uint64_t bl = 0, be = 0, cnt = 0;
      for (uint64_t u = t * stride; u < (1ull << 32); u += nt * stride) {
        float x = sttmath::as_f32((uint32_t)u);
        float a = logf(x), b = sttmath::glibc_logf(x);
        bool nan_a = a != a, nan_b = b != b;
        if (!(nan_a && nan_b) && sttmath::as_u32(a) != sttmath::as_u32(b)) {
          if (bl < 3) fprintf(stderr, "logf mismatch x=%a libm=%a ours=%a\n", x, a, b);
          ++bl;
        }
        float c = expf(x), d = sttmath::glibc_expf(x);
        bool nan_c = c != c, nan_d = d != d;
        if (!(nan_c && nan_d) && sttmath::as_u32(c) != sttmath::as_u32(d)) {
          if (be < 3) fprintf(stderr, "expf mismatch x=%a libm=%a ours=%a\n", x, c, d);
          ++be;
        }
        ++cnt;
      }
      bad_log += bl; bad_exp += be; n += cnt;
    });
  }
  for (auto& x : th) x.join();
  printf("{\"checked\": %llu, \"logf_mismatch\": %llu, \"expf_mismatch\": %llu}\n",
         (unsigned long long)n.load(), (unsigned long long)bad_log.load(), (unsigned long long)bad_exp.load());
  return (bad_log.load() || bad_exp.load()) ? 1 : 0;
}
