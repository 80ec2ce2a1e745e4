// Host-side loader for `.scorer` packages; see scorer_image.cc.
#pragma once
#include <stddef.h>

#include <string>
#include <vector>

#include "scorer_view.h"

namespace sttscorer {

// Same numeric values as the STT_ERR_SCORER_* codes (native_client/coqui-stt.h:100-106).
enum ScorerLoadError {
  SCORER_OK = 0,
  SCORER_UNREADABLE = 0x2005,
  SCORER_INVALID_LM = 0x2006,
  SCORER_NO_TRIE = 0x2007,
  SCORER_INVALID_TRIE = 0x2008,
  SCORER_VERSION_MISMATCH = 0x2009,
};

struct AlphabetBytes {
  std::vector<std::string> labels;  // label index -> UTF-8 string (Alphabet::DecodeSingle, alphabet.cc:204-211)
  uint32_t space_label = 0;
};

// Fills every offset of `view` from the file bytes; view->blob is left NULL for the caller to point
// at the host copy (tests) or the device copy (kernels).
int parse_scorer(const uint8_t* file, size_t size, const AlphabetBytes& alphabet, ScorerView* view);

}  // namespace sttscorer
