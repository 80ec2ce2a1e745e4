// Reader of the reference's model file: a TFLite flatbuffer exported by training/coqui_stt_training/export.py.
//
// Replaces, for the data STT_CreateModel needs, TFLiteModelState::init (native_client/tflitemodelstate.cc:161-338):
//   * tensors by NAME (:211-225): input_node [1, n_steps, 2*n_context+1, n_input], previous_state_c/h [1, n_cell],
//     logits [n_steps, n_classes], metadata_version / _sample_rate / _feature_win_len / _feature_win_step / _beam_width
//     (int32 constants) and metadata_alphabet (string constant in Alphabet::Serialize format, alphabet.cc:101-169),
//     written by export.py:54-72;
//   * geometry from the tensor shapes (:312-335), the version gate (:264-278), window lengths in samples (:288-289);
//   * the weights: the graph of create_inference_graph(batch_size=1, n_steps=16, tflite=True)
//     (deepspeech_model.py:266-403) reaches `logits` through FULLY_CONNECTED operators -- layer 1, 2, 3, the LSTM
//     kernel (once per unrolled timestep, all reading the same [4*n_cell, n_hidden+n_cell] constant), layer 5, layer 6.
//     They are found by walking the operator list in execution order (tensor NAMES of constants are converter
//     dependent, shapes and order are not) and checked against the geometry.  Weights may be float32, float16 or int8
//     (the default "hybrid" export, export.py:145-146: f = scale * (q - zero_point), per tensor or per output row,
//     schema.fbs QuantizationParameters), directly or behind a DEQUANTIZE operator.
//   * relu_clip from the first MINIMUM operator with a scalar constant (deepspeech_model.py:79 `tf.minimum(relu, clip)`).
//
// The flatbuffer wire format is read directly (tensorflow/lite/schema/schema.fbs for the field numbers): a table is a
// signed offset to its vtable {u16 vtable bytes, u16 table bytes, u16 field offset...}; vectors are u32 length + items;
// strings are u32 length + bytes; every reference is an unsigned offset relative to where it is stored.  Every access
// is bounds checked against the caller's buffer, which is used IN PLACE (tflitemodelstate.cc:169-174: BuildFromBuffer
// does not copy either); nothing is retained after load.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "model_file.h"

namespace sttmodel {
namespace {

struct Fb {
  const uint8_t* p;
  size_t n;
  bool ok = true;

  template <class T>
  T rd(size_t off) {
    T v{};
    if (off > n || n - off < sizeof(T)) {
      ok = false;
      return v;
    }
    memcpy(&v, p + off, sizeof(T));
    return v;
  }
  // position of field `id` inside the table at `t`, or 0 if absent
  size_t field(size_t t, int id) {
    if (!t) return 0;
    const int32_t so = rd<int32_t>(t);
    const long long vt = (long long)t - so;
    if (!ok || vt < 0 || (size_t)vt + 4 > n) {
      ok = false;
      return 0;
    }
    const uint16_t vsize = rd<uint16_t>((size_t)vt);
    const size_t slot = 4 + 2 * (size_t)id;
    if (slot + 2 > vsize) return 0;
    const uint16_t fo = rd<uint16_t>((size_t)vt + slot);
    return fo ? t + fo : 0;
  }
  size_t indirect(size_t pos) {  // follow a uoffset stored at pos
    if (!pos) return 0;
    const uint32_t o = rd<uint32_t>(pos);
    const size_t tgt = pos + o;
    if (!ok || o == 0 || tgt >= n) {
      ok = false;
      return 0;
    }
    return tgt;
  }
  size_t child(size_t t, int id) { return indirect(field(t, id)); }
  template <class T>
  T scalar(size_t t, int id, T dflt) {
    const size_t f = field(t, id);
    return f ? rd<T>(f) : dflt;
  }
  uint32_t vec_len(size_t v) {
    if (!v) return 0;
    const uint32_t len = rd<uint32_t>(v);
    if (len > n) {  // even one byte per element would not fit: corrupt length
      ok = false;
      return 0;
    }
    return len;
  }
  // element position of a vector of scalars / inline structs
  size_t vec_at(size_t v, uint32_t i, size_t elem) {
    const size_t pos = v + 4 + (size_t)i * elem;
    if (pos + elem > n) {
      ok = false;
      return 0;
    }
    return pos;
  }
  size_t vec_table(size_t v, uint32_t i) { return indirect(vec_at(v, i, 4)); }
  std::string str(size_t s) {
    if (!s) return std::string();
    const uint32_t len = rd<uint32_t>(s);
    if (!ok || s + 4 + (size_t)len > n) {
      ok = false;
      return std::string();
    }
    return std::string(reinterpret_cast<const char*>(p + s + 4), len);
  }
};

// schema.fbs field numbers
enum { kModelVersion = 0, kModelOpCodes = 1, kModelSubgraphs = 2, kModelBuffers = 4 };
enum { kSgTensors = 0, kSgInputs = 1, kSgOutputs = 2, kSgOperators = 3 };
enum { kTShape = 0, kTType = 1, kTBuffer = 2, kTName = 3, kTQuant = 4 };
enum { kQScale = 2, kQZeroPoint = 3, kQDim = 6 };
enum { kOpIndex = 0, kOpInputs = 1, kOpOutputs = 2 };
enum { kOcDeprecated = 0, kOcBuiltin = 3 };
enum { kTypeF32 = 0, kTypeF16 = 1, kTypeI32 = 2, kTypeString = 5, kTypeI8 = 9 };
enum { kOpDequantize = 6, kOpFullyConnected = 9, kOpMinimum = 57 };

struct Tensor {
  std::vector<int32_t> shape;
  int type = 0;
  uint32_t buffer = 0;
  std::string name;
  std::vector<float> scale;
  std::vector<int64_t> zero_point;
  int qdim = 0;
};
struct Op {
  int code = -1;
  std::vector<int32_t> in, out;
};

struct Graph {
  Fb fb;
  size_t buffers = 0;
  std::vector<Tensor> tensors;
  std::vector<Op> ops;

  // constant data of a tensor: pointer + byte count into the caller's buffer (null if the tensor has none)
  const uint8_t* data(int ti, size_t* bytes) {
    *bytes = 0;
    if (ti < 0 || (size_t)ti >= tensors.size()) return nullptr;
    const uint32_t bi = tensors[ti].buffer;
    if (bi == 0 || bi >= fb.vec_len(buffers)) return nullptr;
    const size_t b = fb.vec_table(buffers, bi);
    const size_t d = fb.child(b, 0);
    if (!d) return nullptr;
    const uint32_t len = fb.vec_len(d);
    if (!fb.ok || d + 4 + (size_t)len > fb.n) return nullptr;
    *bytes = len;
    return fb.p + d + 4;
  }
  int producer(int ti) {
    for (size_t k = 0; k < ops.size(); ++k)
      for (int32_t o : ops[k].out)
        if (o == ti) return (int)k;
    return -1;
  }
  // constant behind `ti`: its own buffer, or the first constant input of the operator chain that produces it
  // (the reference runs those operators once: tflitemodelstate.cc:226-262)
  int visits = 0;   // bound on the walk below: a crafted graph must not make it exponential
  int constant_source(int ti, int depth = 0) {
    if (depth == 0) visits = 0;
    size_t nbytes;
    if (data(ti, &nbytes) && nbytes) return ti;
    if (depth > 4 || ++visits > 4096) return -1;
    const int k = producer(ti);
    if (k < 0) return -1;
    for (int32_t in : ops[k].in) {
      const int s = constant_source(in, depth + 1);
      if (s >= 0) return s;
    }
    return -1;
  }
  int find(const char* name) {
    for (size_t i = 0; i < tensors.size(); ++i)
      if (tensors[i].name == name) return (int)i;
    return -1;
  }
};

bool parse_graph(Graph* g) {
  Fb& fb = g->fb;
  if (fb.n < 16 || memcmp(fb.p + 4, "TFL3", 4) != 0) return false;
  const uint32_t root = fb.rd<uint32_t>(0);
  if (root < 8 || root >= fb.n) return false;
  const size_t model = root;
  g->buffers = fb.child(model, kModelBuffers);
  const size_t codes = fb.child(model, kModelOpCodes), sgs = fb.child(model, kModelSubgraphs);
  if (!fb.ok || !g->buffers || !sgs || fb.vec_len(sgs) < 1) return false;
  std::vector<int> code_of(fb.vec_len(codes));
  for (uint32_t i = 0; i < code_of.size(); ++i) {
    const size_t oc = fb.vec_table(codes, i);
    // builtin_code (int32) when the converter wrote it, else the deprecated byte field (schema.fbs:1114-1130)
    const int32_t b = fb.scalar<int32_t>(oc, kOcBuiltin, 0);
    const int8_t d = fb.scalar<int8_t>(oc, kOcDeprecated, 0);
    code_of[i] = b > d ? b : d;
  }
  const size_t sg = fb.vec_table(sgs, 0);
  const size_t tens = fb.child(sg, kSgTensors), opv = fb.child(sg, kSgOperators);
  if (!fb.ok || !tens) return false;
  const uint32_t nt = fb.vec_len(tens), no = fb.vec_len(opv);
  if (nt > (1u << 20) || no > (1u << 20)) return false;
  g->tensors.resize(nt);
  for (uint32_t i = 0; i < nt && fb.ok; ++i) {
    const size_t t = fb.vec_table(tens, i);
    Tensor& T = g->tensors[i];
    const size_t sh = fb.child(t, kTShape);
    const uint32_t rank = fb.vec_len(sh);
    if (rank > 8) return false;
    for (uint32_t d = 0; d < rank; ++d) T.shape.push_back(fb.rd<int32_t>(fb.vec_at(sh, d, 4)));
    T.type = fb.scalar<int8_t>(t, kTType, 0);
    T.buffer = fb.scalar<uint32_t>(t, kTBuffer, 0);
    T.name = fb.str(fb.child(t, kTName));
    const size_t q = fb.child(t, kTQuant);
    if (q) {
      const size_t sc = fb.child(q, kQScale), zp = fb.child(q, kQZeroPoint);
      const uint32_t nsc = fb.vec_len(sc), nzp = fb.vec_len(zp);
      if (nsc > (1u << 20) || nzp > (1u << 20)) return false;
      for (uint32_t k = 0; k < nsc; ++k) T.scale.push_back(fb.rd<float>(fb.vec_at(sc, k, 4)));
      for (uint32_t k = 0; k < nzp; ++k) T.zero_point.push_back(fb.rd<int64_t>(fb.vec_at(zp, k, 8)));
      T.qdim = fb.scalar<int32_t>(q, kQDim, 0);
    }
  }
  g->ops.resize(no);
  for (uint32_t i = 0; i < no && fb.ok; ++i) {
    const size_t o = fb.vec_table(opv, i);
    Op& O = g->ops[i];
    const uint32_t ci = fb.scalar<uint32_t>(o, kOpIndex, 0);
    O.code = ci < code_of.size() ? code_of[ci] : -1;
    const size_t in = fb.child(o, kOpInputs), out = fb.child(o, kOpOutputs);
    const uint32_t ni = fb.vec_len(in), nout = fb.vec_len(out);
    if (ni > 4096 || nout > 4096) return false;
    for (uint32_t k = 0; k < ni; ++k) O.in.push_back(fb.rd<int32_t>(fb.vec_at(in, k, 4)));
    for (uint32_t k = 0; k < nout; ++k) O.out.push_back(fb.rd<int32_t>(fb.vec_at(out, k, 4)));
  }
  return fb.ok;
}

float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 31u, man = h & 1023u;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {  // subnormal
      int e = -1;
      uint32_t m = man;
      do { ++e; m <<= 1; } while (!(m & 1024u));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 1023u) << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

// constant tensor -> float32 values (row-major, `count` elements)
bool read_floats(Graph* g, int ti, size_t count, std::vector<float>* out) {
  int src = ti;
  size_t nbytes = 0;
  const uint8_t* d = g->data(src, &nbytes);
  if (!d || !nbytes) {
    const int k = g->producer(ti);  // float16 / int8 constant behind a DEQUANTIZE operator
    if (k < 0 || g->ops[k].code != kOpDequantize || g->ops[k].in.empty()) return false;
    src = g->ops[k].in[0];
    d = g->data(src, &nbytes);
    if (!d || !nbytes) return false;
  }
  const Tensor& T = g->tensors[src];
  // the byte count is checked BEFORE anything is allocated: `count` comes from shape fields of the file
  const size_t elem = T.type == kTypeF32 ? 4 : T.type == kTypeF16 ? 2 : T.type == kTypeI8 ? 1 : 0;
  if (!elem || count > nbytes || nbytes != count * elem) return false;
  out->resize(count);
  if (T.type == kTypeF32) {
    if (nbytes != count * 4) return false;
    memcpy(out->data(), d, nbytes);
  } else if (T.type == kTypeF16) {
    if (nbytes != count * 2) return false;
    for (size_t i = 0; i < count; ++i) {
      uint16_t h;
      memcpy(&h, d + 2 * i, 2);
      (*out)[i] = half_to_float(h);
    }
  } else if (T.type == kTypeI8) {
    if (nbytes != count || T.scale.empty()) return false;
    // f = scale * (q - zero_point); per tensor, or per slice of dimension `qdim` (schema.fbs:71-95)
    const size_t nscale = T.scale.size();
    size_t inner = 1, dim = 1;
    if (nscale > 1) {
      if (T.qdim < 0 || (size_t)T.qdim >= T.shape.size() || (size_t)T.shape[T.qdim] != nscale) return false;
      dim = nscale;
      for (size_t k = (size_t)T.qdim + 1; k < T.shape.size(); ++k) inner *= (size_t)T.shape[k];
    }
    for (size_t i = 0; i < count; ++i) {
      const size_t c = nscale > 1 ? (i / inner) % dim : 0;
      const int64_t zp = T.zero_point.empty() ? 0 : T.zero_point[c < T.zero_point.size() ? c : 0];
      (*out)[i] = T.scale[c] * (float)((int)(int8_t)d[i] - (int)zp);
    }
  } else {
    return false;
  }
  return true;
}

bool read_i32(Graph* g, const char* name, int32_t* v) {
  const int ti = g->find(name);
  if (ti < 0) return false;
  const int src = g->constant_source(ti);
  if (src < 0 || g->tensors[src].type != kTypeI32) return false;
  size_t nbytes;
  const uint8_t* d = g->data(src, &nbytes);
  if (!d || nbytes < 4) return false;
  memcpy(v, d, 4);
  return true;
}

// [out, in] row-major -> TF's [in, out]
void transpose_into(const std::vector<float>& w, size_t n_out, size_t n_in, std::vector<float>* dst) {
  dst->resize(n_in * n_out);
  for (size_t o = 0; o < n_out; ++o)
    for (size_t i = 0; i < n_in; ++i) (*dst)[i * n_out + o] = w[o * n_in + i];
}

}  // namespace

bool looks_like_tflite(const uint8_t* data, size_t size) { return data && size >= 8 && memcmp(data + 4, "TFL3", 4) == 0; }

int load_tflite(const uint8_t* data, size_t size, HostModel* m) {
  Graph g;
  g.fb.p = data;
  g.fb.n = size;
  if (!parse_graph(&g)) {
    fprintf(stderr, "Error at reading model buffer: not a well-formed TFLite flatbuffer\n");
    return kFailInitMmap;
  }
  // ---- metadata (tflitemodelstate.cc:226-304)
  int32_t version = 0, sample_rate = 0, win_len_ms = 0, win_step_ms = 0, beam = 0;
  if (!read_i32(&g, "metadata_version", &version)) {
    fprintf(stderr, "Unable to read model file version.\n");
    return kIncompatible;
  }
  if (version < 6) {  // ds_graph_version(), training/coqui_stt_training/GRAPH_VERSION
    fprintf(stderr, "Specified model file version (%d) is incompatible with minimum version supported by this client (6).\n",
            version);
    return kIncompatible;
  }
  if (!read_i32(&g, "metadata_sample_rate", &sample_rate) || sample_rate <= 0) {
    fprintf(stderr, "Unable to read model sample rate.\n");
    return kIncompatible;
  }
  if (!read_i32(&g, "metadata_feature_win_len", &win_len_ms) || !read_i32(&g, "metadata_feature_win_step", &win_step_ms)) {
    fprintf(stderr, "Unable to read model feature window informations.\n");
    return kIncompatible;
  }
  if (!read_i32(&g, "metadata_beam_width", &beam) || beam <= 0) return kIncompatible;
  m->sample_rate = (uint32_t)sample_rate;
  m->win_len = (uint32_t)(sample_rate * (win_len_ms / 1000.0));    // :288-289
  m->win_step = (uint32_t)(sample_rate * (win_step_ms / 1000.0));
  m->beam_width = (uint32_t)beam;
  {
    const int ti = g.find("metadata_alphabet");
    const int src = ti < 0 ? -1 : g.constant_source(ti);
    size_t nbytes = 0;
    const uint8_t* d = src < 0 ? nullptr : g.data(src, &nbytes);
    // TFLite string tensor: int32 count, int32 offsets[count + 1], bytes (string_util.h); GetString(tensor, 0)
    if (!d || g.tensors[src].type != kTypeString || nbytes < 12) return kInvalidAlphabet;
    int32_t cnt, o0, o1;
    memcpy(&cnt, d, 4);
    memcpy(&o0, d + 4, 4);
    memcpy(&o1, d + 8, 4);
    if (cnt < 1 || o0 < 0 || o1 < o0 || (size_t)o1 > nbytes) return kInvalidAlphabet;
    if (deserialize_alphabet(d + o0, (size_t)(o1 - o0), &m->labels, &m->space_label) != 0) return kInvalidAlphabet;
  }
  // ---- geometry from tensor shapes (:312-335)
  const int t_in = g.find("input_node"), t_c = g.find("previous_state_c"), t_h = g.find("previous_state_h"),
            t_logits = g.find("logits");
  if (t_in < 0 || t_c < 0 || t_h < 0 || t_logits < 0) {
    fprintf(stderr, "Model file lacks input_node / previous_state_c / previous_state_h / logits tensors.\n");
    return kIncompatible;
  }
  const std::vector<int32_t>& s_in = g.tensors[t_in].shape;
  const std::vector<int32_t>& s_lg = g.tensors[t_logits].shape;
  const std::vector<int32_t>& s_c = g.tensors[t_c].shape;
  if (s_in.size() != 4 || s_lg.size() != 2 || s_c.size() != 2 || g.tensors[t_h].shape != s_c) return kInvalidShape;
  if (s_in[1] <= 0 || s_in[2] <= 0 || (s_in[2] & 1) == 0 || s_in[3] <= 0 || s_c[1] <= 0 || s_lg[1] <= 1) return kInvalidShape;
  m->n_steps = (uint32_t)s_in[1];
  m->n_context = (uint32_t)((s_in[2] - 1) / 2);
  m->n_input = (uint32_t)s_in[3];
  m->n_cell = (uint32_t)s_c[1];
  m->n_classes = (uint32_t)s_lg[1];
  if (m->n_classes - 1 != m->labels.size()) {
    fprintf(stderr, "Error: Alphabet size does not match loaded model: alphabet has size %zu, but model has %u classes in its "
                    "output. Make sure you're passing an alphabet file with the same size as the one used for training.\n",
            m->labels.size(), m->n_classes - 1);
    return kInvalidAlphabet;
  }
  if (m->n_steps > 4096 || m->n_input > 4096 || m->n_context > 1024 || m->n_cell > (1u << 16) || m->n_classes > (1u << 16))
    return kInvalidShape;
  // ---- the six distinct FULLY_CONNECTED weight matrices, in execution order
  struct Fc { int w, b; };
  std::vector<Fc> fcs;
  for (const Op& op : g.ops) {
    if (op.code != kOpFullyConnected || op.in.size() < 2) continue;
    const int w = op.in[1], b = op.in.size() > 2 ? op.in[2] : -1;
    if (w < 0 || (size_t)w >= g.tensors.size() || (b >= 0 && (size_t)b >= g.tensors.size())) {
      fprintf(stderr, "Model file: a FULLY_CONNECTED operator refers to a tensor that does not exist.\n");
      return kIncompatible;
    }
    bool seen = false;
    for (const Fc& f : fcs) {
      if (f.w == w) seen = true;
      // the unrolled LSTM may carry one copy of the kernel per timestep: same buffer or same bytes
      if (!seen && g.tensors[w].shape == g.tensors[f.w].shape) {
        size_t n1, n2;
        const uint8_t* d1 = g.data(g.constant_source(w), &n1);
        const uint8_t* d2 = g.data(g.constant_source(f.w), &n2);
        if (d1 && d2 && n1 == n2 && (d1 == d2 || memcmp(d1, d2, n1) == 0)) seen = true;
      }
    }
    if (!seen) fcs.push_back({w, b});
  }
  if (fcs.size() != 6) {
    fprintf(stderr, "Model file: expected 6 distinct FULLY_CONNECTED weight matrices (layers 1-3, LSTM, 5, 6), found %zu.\n",
            fcs.size());
    return kIncompatible;
  }
  auto dims = [&](int ti, size_t* n_out, size_t* n_in) {
    if (ti < 0 || (size_t)ti >= g.tensors.size() || g.tensors[ti].shape.size() != 2) return false;
    if (g.tensors[ti].shape[0] <= 0 || g.tensors[ti].shape[1] <= 0) return false;
    *n_out = (size_t)g.tensors[ti].shape[0];
    *n_in = (size_t)g.tensors[ti].shape[1];
    return true;
  };
  size_t o[6], in[6];
  for (int k = 0; k < 6; ++k)
    if (!dims(fcs[k].w, &o[k], &in[k])) return kInvalidShape;
  const size_t in1 = (size_t)(2 * m->n_context + 1) * m->n_input, H = o[0], C = m->n_cell, K = m->n_classes;
  if (H == 0 || H > (1u << 16)) return kInvalidShape;
  // layer 1 [H, in1]; 2, 3 [H, H]; LSTM [4C, H + C]; 5 [H5, C]; 6 [K, H5] with H5 == H in every released geometry
  if (in[0] != in1 || o[1] != H || in[1] != H || o[2] != H || in[2] != H || o[3] != 4 * C || in[3] != H + C ||
      in[4] != C || o[4] != H || o[5] != K || in[5] != H) {
    fprintf(stderr, "Model file: FULLY_CONNECTED shapes do not form the expected layer stack.\n");
    return kInvalidShape;
  }
  m->n_hidden = (uint32_t)H;
  std::vector<float>* wdst[6] = {&m->w1, &m->w2, &m->w3, &m->lstm_kernel, &m->w5, &m->w6};
  std::vector<float>* bdst[6] = {&m->b1, &m->b2, &m->b3, &m->lstm_bias, &m->b5, &m->b6};
  for (int k = 0; k < 6; ++k) {
    std::vector<float> w;
    if (!read_floats(&g, fcs[k].w, o[k] * in[k], &w)) {
      fprintf(stderr, "Model file: cannot read the weights of FULLY_CONNECTED #%d.\n", k + 1);
      return kIncompatible;
    }
    transpose_into(w, o[k], in[k], wdst[k]);
    if (fcs[k].b >= 0) {
      if (!read_floats(&g, fcs[k].b, o[k], bdst[k])) return kIncompatible;
    } else {
      bdst[k]->assign(o[k], 0.f);
    }
  }
  // ---- relu_clip
  m->relu_clip = 20.f;
  for (const Op& op : g.ops) {
    if (op.code != kOpMinimum) continue;
    bool found = false;
    for (int32_t ti : op.in) {
      size_t nbytes;
      const uint8_t* d = g.data(ti, &nbytes);
      if (d && nbytes == 4 && g.tensors[ti].type == kTypeF32) {
        memcpy(&m->relu_clip, d, 4);
        found = true;
      }
    }
    if (found) break;
  }
  if (!g.fb.ok) return kFailInitMmap;
  if (!(m->relu_clip > 0.f) || !m->win_len || !m->win_step) return kInvalidShape;
  return kOk;
}

}  // namespace sttmodel
