// Host harness for tests/test_loader_fuzz.py: the `stt` client's WAV reader (stt_b200/csrc/client.cc read_wav) on
// damaged RIFF headers, truncated files and odd-sized data chunks, under AddressSanitizer / UBSan.
#define main client_main
#include "../../stt_b200/csrc/client.cc"
#undef main
#include <random>

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const char* scratch = argv[2];
  const int iters = atoi(argv[3]);
  std::ifstream f(argv[1], std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  std::string s = ss.str();
  std::vector<short> pcm;
  if (!read_wav(argv[1], 16000, &pcm)) return 3;   // the intact file must load
  const size_t n_samples = pcm.size();
  s.resize(std::min<size_t>(s.size(), 4096));      // header + some samples is enough
  std::mt19937_64 rng(3);
  int ok = 0, bad = 0;
  for (int it = 0; it < iters; ++it) {
    std::string b = s;
    if (it % 4 == 0) b.resize(rng() % b.size());
    const int flips = 1 + (int)(rng() % 3);
    for (int k = 0; k < flips && !b.empty(); ++k) b[rng() % std::min<size_t>(64, b.size())] = (char)rng();
    if (it % 7 == 0 && b.size() > 44) b.resize(45 + rng() % 8);   // odd / tiny data sections
    {
      std::ofstream o(scratch, std::ios::binary);
      o.write(b.data(), (std::streamsize)b.size());
    }
    pcm.clear();
    (read_wav(scratch, 16000, &pcm) ? ok : bad)++;
  }
  printf("samples %zu accepted %d rejected %d\n", n_samples, ok, bad);
  return 0;
}
