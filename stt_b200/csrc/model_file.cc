#include "model_file.h"

#include <cstdio>
#include <cstring>

namespace sttmodel {

int deserialize_alphabet(const uint8_t* buf, size_t size, std::vector<std::string>* labels, uint32_t* space_label) {
  size_t off = 0;
  if (size - off < 2) return 1;
  uint16_t count;
  memcpy(&count, buf + off, 2);
  off += 2;
  labels->assign(count, std::string());
  *space_label = 0xfffffffeu;  // Alphabet::InitFromLabels uses -2 for "no space label"
  for (int i = 0; i < count; ++i) {
    uint16_t label, len;
    if (size - off < 2) return 1;
    memcpy(&label, buf + off, 2);
    off += 2;
    if (size - off < 2) return 1;
    memcpy(&len, buf + off, 2);
    off += 2;
    if (size - off < len) return 1;
    if (label >= count) return 1;  // the reference asserts contiguity (alphabet.cc:104)
    (*labels)[label].assign(reinterpret_cast<const char*>(buf + off), len);
    off += len;
    if ((*labels)[label] == " ") *space_label = label;
  }
  return 0;
}

namespace {
struct Reader {
  const uint8_t* p;
  size_t n, off = 0;
  bool ok = true;
  template <class T>
  T get() {
    T v{};
    if (off + sizeof(T) > n) {
      ok = false;
      return v;
    }
    memcpy(&v, p + off, sizeof(T));
    off += sizeof(T);
    return v;
  }
  bool floats(std::vector<float>* dst, size_t count) {
    if (!ok || off + count * 4 > n) {
      ok = false;
      return false;
    }
    dst->resize(count);
    memcpy(dst->data(), p + off, count * 4);
    off += count * 4;
    return true;
  }
};
}  // namespace

int load_from_buffer(const uint8_t* data, size_t size, HostModel* m) {
  if (!data || size < 12) return kNoModel;
  if (looks_like_tflite(data, size)) return load_tflite(data, size, m);
  if (memcmp(data, "STTB200W", 8) != 0) return kFailInitMmap;  // not a model file we can map
  Reader r{data, size};
  r.off = 8;
  const uint32_t version = r.get<uint32_t>();
  if (version != 1) return kIncompatible;
  m->sample_rate = r.get<uint32_t>();
  m->win_len = r.get<uint32_t>();
  m->win_step = r.get<uint32_t>();
  m->n_input = r.get<uint32_t>();
  m->n_context = r.get<uint32_t>();
  m->n_hidden = r.get<uint32_t>();
  m->n_cell = r.get<uint32_t>();
  m->n_classes = r.get<uint32_t>();
  m->n_steps = r.get<uint32_t>();
  m->beam_width = r.get<uint32_t>();
  m->relu_clip = r.get<float>();
  const uint32_t alpha_bytes = r.get<uint32_t>();
  if (!r.ok || r.off + alpha_bytes > size) return kInvalidShape;
  if (deserialize_alphabet(data + r.off, alpha_bytes, &m->labels, &m->space_label) != 0) return kInvalidAlphabet;
  r.off += alpha_bytes;
  // tflitemodelstate.cc:319-329: logits' last dimension must be alphabet size + 1
  if (m->n_classes != m->labels.size() + 1) return kInvalidAlphabet;
  if (!m->n_input || !m->n_hidden || !m->n_cell || !m->n_steps || !m->win_len || !m->win_step || !m->sample_rate)
    return kInvalidShape;
  // bound every dimension before sizes are multiplied (a crafted header must not wrap size_t or over-read)
  if (m->n_input > 4096 || m->n_context > 1024 || m->n_hidden > (1u << 16) || m->n_cell > (1u << 16) ||
      m->n_classes > (1u << 16) || m->n_steps > 4096 || m->win_len > (1u << 20) || m->win_step > (1u << 20))
    return kInvalidShape;
  const size_t in1 = (size_t)(2 * m->n_context + 1) * m->n_input, H = m->n_hidden, C = m->n_cell, K = m->n_classes;
  r.floats(&m->w1, in1 * H); r.floats(&m->b1, H);
  r.floats(&m->w2, H * H);   r.floats(&m->b2, H);
  r.floats(&m->w3, H * H);   r.floats(&m->b3, H);
  r.floats(&m->lstm_kernel, (H + C) * 4 * C); r.floats(&m->lstm_bias, 4 * C);
  r.floats(&m->w5, C * H);   r.floats(&m->b5, H);
  r.floats(&m->w6, H * K);   r.floats(&m->b6, K);
  if (!r.ok) return kInvalidShape;
  return kOk;
}

int load_from_file(const char* path, HostModel* out) {
  FILE* f = fopen(path, "rb");
  if (!f) return kFailInitMmap;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf(sz > 0 ? sz : 0);
  size_t got = sz > 0 ? fread(buf.data(), 1, sz, f) : 0;
  fclose(f);
  if ((long)got != sz) return kFailInitMmap;
  return load_from_buffer(buf.data(), buf.size(), out);
}

}  // namespace sttmodel
