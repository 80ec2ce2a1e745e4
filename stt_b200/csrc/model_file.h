// Native acoustic-model file (".sttw").  The reference loads a TFLite flatbuffer (tflitemodelstate.cc:161-338);
// no model file ships in the reference tree and TFLite cannot be built offline, so this build defines a
// trivial container carrying exactly the data STT_CreateModel extracts from the flatbuffer: the metadata_* values
// (sample rate, window length/step, beam width, alphabet), the geometry the reference derives from tensor shapes
// (n_steps, n_context, n_input, n_hidden / n_cell, n_classes) and the fp32 weights of
// training/coqui_stt_training/deepspeech_model.py:204-263 in TF's own [in, out] layout.
// The reference's own container, a .tflite flatbuffer, is read by tflite_reader.cc into the same `HostModel`.
//
// Layout (little endian):
//   char[8] "STTB200W" | u32 version (=1)
//   u32 sample_rate | u32 win_len_samples | u32 win_step_samples | u32 n_input | u32 n_context
//   u32 n_hidden | u32 n_cell | u32 n_classes | u32 n_steps | u32 beam_width | f32 relu_clip
//   u32 alphabet_bytes | alphabet (Alphabet::Serialize format, alphabet.cc:101-169: u16 count, {u16 label, u16 len, bytes})
//   f32 tensors: w1[(2c+1)*n_input, H] b1[H] w2[H,H] b2[H] w3[H,H] b3[H] lstm_kernel[H+C, 4C] lstm_bias[4C]
//                w5[C,H] b5[H] w6[H,K] b6[K]
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace sttmodel {

struct HostModel {
  uint32_t sample_rate = 0, win_len = 0, win_step = 0, n_input = 0, n_context = 0;
  uint32_t n_hidden = 0, n_cell = 0, n_classes = 0, n_steps = 0, beam_width = 0;
  float relu_clip = 20.f;
  std::vector<std::string> labels;  // label index -> UTF-8 string
  uint32_t space_label = 0;
  std::vector<float> w1, b1, w2, b2, w3, b3, lstm_kernel, lstm_bias, w5, b5, w6, b6;
};

enum LoadError {
  kOk = 0,
  kNoModel = 0x1000,          // STT_ERR_NO_MODEL
  kInvalidAlphabet = 0x2000,  // STT_ERR_INVALID_ALPHABET
  kInvalidShape = 0x2001,     // STT_ERR_INVALID_SHAPE
  kIncompatible = 0x2003,     // STT_ERR_MODEL_INCOMPATIBLE
  kFailInitMmap = 0x3000,     // STT_ERR_FAIL_INIT_MMAP
};

// Either container: a TFLite flatbuffer as exported by the reference (tflite_reader.cc) or the native .sttw layout above.
int load_from_buffer(const uint8_t* data, size_t size, HostModel* out);
bool looks_like_tflite(const uint8_t* data, size_t size);
int load_tflite(const uint8_t* data, size_t size, HostModel* out);
int load_from_file(const char* path, HostModel* out);
// Alphabet::Deserialize (alphabet.cc:127-169) incl. the space-label detection.
int deserialize_alphabet(const uint8_t* buf, size_t size, std::vector<std::string>* labels, uint32_t* space_label);

}  // namespace sttmodel
