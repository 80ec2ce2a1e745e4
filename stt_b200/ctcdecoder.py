"""Decoder-only Python surface mirroring the reference's `coqui_stt_ctcdecoder` package
(native_client/ctcdecode/__init__.py: Alphabet :17-80, Scorer :82-120, DecodeResult :117-120,
ctc_beam_search_decoder :122-178, ctc_beam_search_decoder_batch :244-312), executed by the GPU beam search.

`cutoff_prob` / `cutoff_top_n` (vocabulary pruning, get_pruned_emissions), alphabets of up to 255 labels and bytes-output
(UTF-8) scorers with `UTF8Alphabet` run the general kernel (stt_b200/csrc/decoder_general.cuh); the 28-letter alphabet with
the defaults runs the shared-memory kernel.  `num_processes` is accepted and ignored (utterances are decoded one CTA each).
"""
from collections import namedtuple

import numpy as np

from . import api, synth

DecodeResult = namedtuple("DecodeResult", ["confidence", "transcript", "tokens", "timesteps"])


class Alphabet(object):
    """native_client/alphabet.cc: one label per line, lines starting with '#' are comments ('\\#' escapes)."""

    def __init__(self, config_path=None):
        self._labels = []
        if config_path:
            with open(config_path, "r", encoding="utf-8") as f:
                for line in f.read().split("\n"):
                    if line.startswith("\\#"):
                        line = line[1:]
                    elif line.startswith("#"):
                        continue
                    if line == "":
                        continue
                    self._labels.append(line)
        self._index = {l: i for i, l in enumerate(self._labels)}

    def InitFromLabels(self, data):
        self._labels = list(data)
        self._index = {l: i for i, l in enumerate(self._labels)}

    def GetSize(self):
        return len(self._labels)

    def GetLabels(self):
        return list(self._labels)

    def CanEncodeSingle(self, input):
        return input in self._index

    def CanEncode(self, input):
        return all(ch in self._index for ch in input)

    def EncodeSingle(self, input):
        return self._index[input]

    def Encode(self, input):
        return [self._index[ch] for ch in input]

    def DecodeSingle(self, input):
        return self._labels[input]

    def Decode(self, input):
        return "".join(self._labels[int(i)] for i in input)


class UTF8Alphabet(Alphabet):
    """Bytes-output mode (native_client/alphabet.h:80-100, ctcdecode/__init__.py:575-627): 255 labels, label n is the
    single byte n + 1; text is encoded to / decoded from its UTF-8 bytes."""

    def __init__(self):
        super(UTF8Alphabet, self).__init__()
        self._labels = [bytes([i + 1]) for i in range(255)]
        self._index = {l: i for i, l in enumerate(self._labels)}

    def CanEncodeSingle(self, input):
        return True

    def CanEncode(self, input):
        return True

    def EncodeSingle(self, input):
        return self._index[input.encode("utf-8")]

    def Encode(self, input):
        return [b - 1 for b in input.encode("utf-8")]

    def DecodeSingle(self, input):
        return self._labels[input].decode("utf-8")

    def Decode(self, input):
        return b"".join(self._labels[int(i)] for i in input).decode("utf-8", errors="replace")


class Scorer(object):
    def __init__(self, alpha=None, beta=None, scorer_path=None, alphabet=None):
        self.alpha, self.beta, self.scorer_path, self.alphabet = alpha, beta, scorer_path, alphabet
        if alphabet:
            assert alpha is not None, "alpha parameter is required"
            assert beta is not None, "beta parameter is required"
            assert scorer_path, "scorer_path parameter is required"

    def reset_params(self, alpha, beta):
        self.alpha, self.beta = alpha, beta


_hosts = {}


def _host(alphabet, scorer):
    """A model object that carries only what the decoder needs (alphabet + scorer); cached per (labels, scorer)."""
    key = (tuple(alphabet.GetLabels()), scorer.scorer_path if scorer else None)
    if key not in _hosts:
        w = synth.make_weights(n_hidden=16, n_classes=alphabet.GetSize() + 1, seed=0)
        m = api.Model(synth.model_bytes(w, labels=alphabet.GetLabels()))
        if scorer:
            m.enableExternalScorer(scorer.scorer_path)
        _hosts[key] = m
    m = _hosts[key]
    if scorer:
        m.setScorerAlphaBeta(scorer.alpha, scorer.beta)
    return m


def ctc_beam_search_decoder_batch(probs_seq, seq_lengths, alphabet, beam_size, num_processes=1, cutoff_prob=1.0,
                                  cutoff_top_n=40, scorer=None, hot_words=dict(), num_results=1):
    probs = np.ascontiguousarray(probs_seq, dtype=np.float64)
    if probs.ndim != 3:
        raise ValueError("probs_seq must be [batch, time, classes]")
    B, T, C = probs.shape
    if C != alphabet.GetSize() + 1:
        raise ValueError("class dimension must be alphabet size + 1")
    m = _host(alphabet, scorer)
    m.setBeamWidth(beam_size)
    lib = api.lib()
    lib.STT_ClearHotWords(m._impl) if scorer else None
    for w, b in (hot_words or {}).items():
        m.addHotWord(w, b)
    out = []
    for base in range(0, B, 256):
        n = min(256, B - base)
        bt = m.createBatch(n, max(T, 1) * 320 + 512)
        bt.set_probs64(probs[base:base + n], np.asarray(seq_lengths[base:base + n], np.int32))
        bt.set_cutoff(cutoff_prob, cutoff_top_n)
        bt.decode(num_results)
        bt.fetch()
        for u in range(n):
            out.append([DecodeResult(c, alphabet.Decode(tok), [int(t) for t in tok], [int(t) for t in ts])
                        for c, tok, ts in bt.results(u, max_tokens=max(T, 1))])
    return out


def ctc_beam_search_decoder(probs_seq, alphabet, beam_size, cutoff_prob=1.0, cutoff_top_n=40, scorer=None,
                            hot_words=dict(), num_results=1):
    p = np.asarray(probs_seq, dtype=np.float64)
    return ctc_beam_search_decoder_batch(p[None], [p.shape[0]], alphabet, beam_size, 1, cutoff_prob, cutoff_top_n,
                                         scorer, hot_words, num_results)[0]
