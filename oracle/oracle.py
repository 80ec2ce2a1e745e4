"""TEST INFRASTRUCTURE ONLY.  ctypes loaders for the oracle libraries built by oracle/Makefile:

  _ref/libstt_oracle.so   CPU restatement of MFCC + acoustic model + stream buffering (oracle/stt_oracle.c)
  _ref/libref_decoder.so  the GENUINE reference decoder / scorer / KenLM / OpenFst behind oracle/ref_shim.cc

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
Nothing in stt_b200/ does, and the product has no CPU path.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_uint, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.realpath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def _load(name):
    p = os.path.join(REF_DIR, name)
    if not os.path.exists(p):
        raise FileNotFoundError("%s missing: run `make -C oracle` (and `make -C oracle ref` where /root/reference exists)" % p)
    return ctypes.CDLL(p)


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_decoder.so"))


def have_port():
    return os.path.exists(os.path.join(REF_DIR, "libstt_oracle.so"))


# ------------------------------------------------------------------------------------------------ port (MFCC + AM)
class _OrcAm(ctypes.Structure):
    _fields_ = [("n_input", c_int), ("n_context", c_int), ("n_hidden", c_int), ("n_cell", c_int), ("n_classes", c_int),
                ("relu_clip", c_float)] + [(n, c_void_p) for n in
                                           ("w1", "b1", "w2", "b2", "w3", "b3", "lstm_kernel", "lstm_bias", "w5", "b5",
                                            "w6", "b6")]


_port = None


def port():
    global _port
    if _port is None:
        L = _load("libstt_oracle.so")
        L.orc_mfcc_new.restype = c_void_p
        L.orc_mfcc_new.argtypes = [c_int, c_double, c_double, c_double, c_int, c_int]
        L.orc_mfcc_free.argtypes = [c_void_p]
        L.orc_mfcc_compute.argtypes = [c_void_p, c_void_p, c_void_p]
        L.orc_spectrogram_frame.argtypes = [c_void_p, c_int, c_int, c_void_p]
        L.orc_am_infer.argtypes = [POINTER(_OrcAm), c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        L.orc_am_quantize.restype = c_void_p
        L.orc_am_quantize.argtypes = [POINTER(_OrcAm)]
        L.orc_am_q_free.argtypes = [c_void_p]
        L.orc_am_infer_hybrid.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        L.orc_stream_new.restype = c_void_p
        L.orc_stream_new.argtypes = [POINTER(_OrcAm), c_int, c_int, c_int, c_int]
        L.orc_stream_free.argtypes = [c_void_p]
        L.orc_stream_feed.argtypes = [c_void_p, c_void_p, c_uint]
        L.orc_stream_flush.argtypes = [c_void_p, c_int]
        L.orc_stream_timesteps.argtypes = [c_void_p]
        L.orc_stream_emitted.argtypes = [c_void_p]
        L.orc_stream_frames.argtypes = [c_void_p]
        L.orc_stream_probs.argtypes = [c_void_p]
        L.orc_stream_probs.restype = POINTER(c_float)
        L.orc_stream_mfcc.argtypes = [c_void_p]
        L.orc_stream_mfcc.restype = POINTER(c_float)
        _port = L
    return _port


def mfcc_from_spectrum(spectrum, sample_rate, lower, upper, n_channels, n_dct):
    L = port()
    spec = np.ascontiguousarray(spectrum, np.float32)
    m = L.orc_mfcc_new(spec.size, sample_rate, lower, upper, n_channels, n_dct)
    out = np.zeros(n_dct, np.float32)
    L.orc_mfcc_compute(m, spec.ctypes.data, out.ctypes.data)
    L.orc_mfcc_free(m)
    return out


def spectrogram_frame(samples, window):
    L = port()
    s = np.ascontiguousarray(samples, np.float32)
    fft = 1
    while fft < window:
        fft *= 2
    out = np.zeros(fft // 2 + 1, np.float32)
    L.orc_spectrogram_frame(s.ctypes.data, s.size, window, out.ctypes.data)
    return out


class PortAM(object):
    """Holds the weight arrays alive and exposes orc_am / orc_stream."""

    def __init__(self, weights, n_input=26, n_context=9, relu_clip=20.0):
        self.w = {k: np.ascontiguousarray(v, np.float32) for k, v in weights.items()}
        H = self.w["b1"].size
        C = self.w["lstm_bias"].size // 4
        K = self.w["b6"].size
        self.am = _OrcAm(n_input, n_context, H, C, K, relu_clip,
                         *[self.w[k].ctypes.data for k in ("w1", "b1", "w2", "b2", "w3", "b3", "lstm_kernel",
                                                          "lstm_bias", "w5", "b5", "w6", "b6")])
        self.n_classes, self.n_input = K, n_input

    def forward_features(self, mfcc, mode="fp32", n_steps=16):
        """mfcc [F, n_input] -> probs [F, K], `n_steps` timesteps per infer call with carried LSTM state, zero-padded last
        call (tflitemodelstate.cc:369-405).  mode "hybrid8" = TFLite hybrid int8 arithmetic (orc_am_infer_hybrid)."""
        L = port()
        F = mfcc.shape[0]
        nc, ni = self.am.n_context, self.n_input
        pad = np.zeros((nc, ni), np.float32)
        seq = np.concatenate([pad, np.asarray(mfcc, np.float32), pad])
        win = np.stack([seq[t:t + 2 * nc + 1].reshape(-1) for t in range(F)]) if F else np.zeros((0, (2 * nc + 1) * ni), np.float32)
        C = self.am.n_cell
        c = np.zeros(C, np.float32)
        h = np.zeros(C, np.float32)
        out = np.zeros((F, self.n_classes), np.float32)
        q = L.orc_am_quantize(byref(self.am)) if mode == "hybrid8" else None
        for t0 in range(0, F, n_steps):
            x = np.zeros((n_steps, win.shape[1]), np.float32)
            n = min(n_steps, F - t0)
            x[:n] = win[t0:t0 + n]
            p = np.zeros((n_steps, self.n_classes), np.float32)
            if q:
                L.orc_am_infer_hybrid(q, x.ctypes.data, n_steps, c.ctypes.data, h.ctypes.data, p.ctypes.data)
            else:
                L.orc_am_infer(byref(self.am), x.ctypes.data, n_steps, c.ctypes.data, h.ctypes.data, p.ctypes.data)
            out[t0:t0 + n] = p[:n]
        if q:
            L.orc_am_q_free(q)
        return out

    def stream(self, pcm, sample_rate=16000, win_len=512, win_step=320, n_steps=16, chunks=None, flush_at=()):
        """Run the restated streaming runtime.  chunks: list of chunk sizes (default: everything at once);
        flush_at: chunk indices after which flushBuffers(false) is called (IntermediateDecodeFlushBuffers)."""
        L = port()
        pcm = np.ascontiguousarray(pcm, np.int16)
        s = L.orc_stream_new(byref(self.am), sample_rate, win_len, win_step, n_steps)
        pos = 0
        sizes = chunks if chunks is not None else [pcm.size]
        for i, n in enumerate(sizes):
            n = min(n, pcm.size - pos)
            if n > 0:
                L.orc_stream_feed(s, pcm[pos:pos + n].ctypes.data, n)
            pos += n
            if i in flush_at:
                L.orc_stream_flush(s, 0)
        if pos < pcm.size:
            L.orc_stream_feed(s, pcm[pos:].ctypes.data, pcm.size - pos)
        L.orc_stream_flush(s, 1)
        T = L.orc_stream_emitted(s)
        F = L.orc_stream_frames(s)
        probs = np.ctypeslib.as_array(L.orc_stream_probs(s), shape=(T, self.n_classes)).copy()
        mfcc = np.ctypeslib.as_array(L.orc_stream_mfcc(s), shape=(F, self.n_input)).copy()
        L.orc_stream_free(s)
        return probs, mfcc


def features_only(pcm, sample_rate=16000, win_len=512, win_step=320, n_steps=16):
    L = port()
    pcm = np.ascontiguousarray(pcm, np.int16)
    s = L.orc_stream_new(None, sample_rate, win_len, win_step, n_steps)
    if pcm.size:
        L.orc_stream_feed(s, pcm.ctypes.data, pcm.size)
    L.orc_stream_flush(s, 1)
    T, F = L.orc_stream_timesteps(s), L.orc_stream_frames(s)
    mfcc = np.ctypeslib.as_array(L.orc_stream_mfcc(s), shape=(F, 26)).copy()
    L.orc_stream_free(s)
    return T, mfcc


# ------------------------------------------------------------------------------------------------ reference decoder
_ref = None


def ref():
    global _ref
    if _ref is None:
        L = _load("libref_decoder.so")
        vp = c_void_p
        L.ref_alphabet_from_file.restype = vp
        L.ref_alphabet_from_file.argtypes = [c_char_p]
        L.ref_alphabet_from_labels.restype = vp
        L.ref_alphabet_from_labels.argtypes = [c_char_p, c_int]
        L.ref_alphabet_size.argtypes = [vp]
        L.ref_alphabet_free.argtypes = [vp]
        L.ref_alphabet_decode.argtypes = [vp, c_void_p, c_int, c_char_p, c_int]
        L.ref_scorer_load.argtypes = [c_char_p, vp, POINTER(vp)]
        L.ref_scorer_free.argtypes = [vp]
        L.ref_scorer_reset_params.argtypes = [vp, c_float, c_float]
        L.ref_scorer_alpha.restype = c_double
        L.ref_scorer_alpha.argtypes = [vp]
        L.ref_scorer_beta.restype = c_double
        L.ref_scorer_beta.argtypes = [vp]
        L.ref_scorer_max_order.argtypes = [vp]
        L.ref_scorer_log_cond_prob.restype = c_double
        L.ref_scorer_log_cond_prob.argtypes = [vp, c_char_p, c_int, c_int, c_int]
        L.ref_make_scorer_package.argtypes = [c_char_p, c_char_p, c_int, vp, c_char_p, c_float, c_float]
        L.ref_make_scorer_package_utf8.argtypes = [c_char_p, c_char_p, c_int, vp, c_char_p, c_float, c_float]
        L.ref_decoder_new.restype = vp
        L.ref_decoder_new.argtypes = [vp, c_int, c_double, c_int, vp, c_char_p, c_void_p, c_int]
        L.ref_decoder_next.argtypes = [vp, c_void_p, c_int, c_int]
        L.ref_decoder_decode.argtypes = [vp, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        L.ref_decoder_beam.argtypes = [vp, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.ref_decoder_free.argtypes = [vp]
        L.ref_decode_batch.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, vp, c_int, c_int, c_double, c_int, vp,
                                       c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.ref_log_sum_exp.restype = c_float
        L.ref_log_sum_exp.argtypes = [c_float, c_float]
        L.ref_class_logprob.restype = c_float
        L.ref_class_logprob.argtypes = [c_double]
        _ref = L
    return _ref


def _nul_join(items):
    return b"".join(s.encode("utf-8") + b"\0" for s in items)


class RefAlphabet(object):
    def __init__(self, labels):
        self.labels = list(labels)
        self.h = ref().ref_alphabet_from_labels(_nul_join(labels), len(labels))

    def decode(self, tokens):
        return "".join(self.labels[t] for t in tokens)


class RefByteAlphabet(RefAlphabet):
    """The 255 single-byte labels of UTF8Alphabet (alphabet.h:80-100): label n is the byte n + 1."""

    def __init__(self):
        self.labels = [bytes([i + 1]) for i in range(255)]
        self.h = ref().ref_alphabet_from_labels(b"".join(l + b"\0" for l in self.labels), 255)

    def decode(self, tokens):
        return b"".join(self.labels[t] for t in tokens)


class RefScorer(object):
    def __init__(self, path, alphabet):
        self.h = c_void_p()
        err = ref().ref_scorer_load(str(path).encode(), alphabet.h, byref(self.h))
        if err != 0:
            raise RuntimeError("reference scorer load failed: 0x%X" % err)

    def set_alpha_beta(self, a, b):
        ref().ref_scorer_reset_params(self.h, a, b)

    def log_cond_prob(self, words, bos):
        return ref().ref_scorer_log_cond_prob(self.h, _nul_join(words), len(words), int(bos), 0)


def ref_decode(probs, alphabet, beam, scorer=None, num_results=1, cutoff_prob=1.0, cutoff_top_n=40, chunk=None):
    """Genuine DecoderState init/next/decode on float32/64 probs [T, C].  Returns [(confidence, tokens, timesteps)]."""
    L = ref()
    p = np.ascontiguousarray(probs, np.float64)  # stt.cc:327 feeds vector<double>(float logits)
    T, C = p.shape
    d = L.ref_decoder_new(alphabet.h, beam, cutoff_prob, cutoff_top_n, scorer.h if scorer else None, b"", None, 0)
    step = chunk or T
    for t0 in range(0, T, step):
        n = min(step, T - t0)
        L.ref_decoder_next(d, p[t0:t0 + n].ctypes.data, n, C)
    max_tok = max(T, 1)
    conf = np.zeros(num_results, np.float64)
    nt = np.zeros(num_results, np.int32)
    tok = np.zeros((num_results, max_tok), np.uint32)
    ts = np.zeros((num_results, max_tok), np.uint32)
    n = L.ref_decoder_decode(d, num_results, max_tok, conf.ctypes.data, nt.ctypes.data, tok.ctypes.data, ts.ctypes.data)
    L.ref_decoder_free(d)
    return [(conf[r], tok[r, :nt[r]].copy(), ts[r, :nt[r]].copy()) for r in range(n)]


def ref_decode_batch(probs, lengths, alphabet, beam, scorer=None, num_processes=1, num_results=1, cutoff_prob=1.0,
                     cutoff_top_n=40):
    """ctc_beam_search_decoder_batch (:608-652) on float64 [B, T, C]."""
    L = ref()
    p = np.ascontiguousarray(probs, np.float64)
    B, T, C = p.shape
    lens = np.ascontiguousarray(lengths, np.int32)
    nres = np.zeros(B, np.int32)
    conf = np.zeros((B, num_results), np.float64)
    nt = np.zeros((B, num_results), np.int32)
    tok = np.zeros((B, num_results, T), np.uint32)
    ts = np.zeros((B, num_results, T), np.uint32)
    L.ref_decode_batch(p.ctypes.data, B, T, C, lens.ctypes.data, alphabet.h, beam, num_processes, cutoff_prob,
                       cutoff_top_n, scorer.h if scorer else None, num_results, T, nres.ctypes.data, conf.ctypes.data,
                       nt.ctypes.data, tok.ctypes.data, ts.ctypes.data)
    return [[(conf[b, r], tok[b, r, :nt[b, r]].copy(), ts[b, r, :nt[b, r]].copy()) for r in range(nres[b])]
            for b in range(B)]
