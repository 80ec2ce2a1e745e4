"""ctypes binding of libstt_b200.so with the method names of the reference's Python package
(native_client/python/__init__.py: Model :26-221, Stream :223-383, metadata wrappers :386-430)."""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_longlong, c_short, c_uint,
                    c_void_p)

import numpy as np

_HERE = os.path.dirname(os.path.realpath(__file__))


def lib_path():
    # STT_B200_LIB: an A/B build of the same library (Makefile `variant`), for measurements only
    return os.environ.get("STT_B200_LIB") or os.path.join(_HERE, "libstt_b200.so")


class STTError(RuntimeError):
    pass


class _TokenMetadata(Structure):
    _fields_ = [("text", c_char_p), ("timestep", c_uint), ("start_time", c_float)]


class _CandidateTranscript(Structure):
    _fields_ = [("tokens", POINTER(_TokenMetadata)), ("num_tokens", c_uint), ("confidence", c_double)]


class _Emissions(Structure):
    _fields_ = [("num_symbols", c_int), ("symbols", POINTER(c_char_p)), ("num_timesteps", c_int),
                ("emissions", POINTER(c_double))]


class _Metadata(Structure):
    _fields_ = [("transcripts", POINTER(_CandidateTranscript)), ("num_transcripts", c_uint),
                ("emissions", POINTER(_Emissions))]


class _Timings(Structure):
    _fields_ = [(n, c_float) for n in ("h2d", "mfcc", "dense123", "lstm_in", "lstm", "dense56", "decode", "d2h", "total")]


_lib = None

# every symbol include/stt_capi.h declares (tests/test_capi_symbols.py checks the .so exports all of them)
DECLARED_SYMBOLS = [
    "STT_CreateModel", "STT_CreateModelFromBuffer", "STT_GetModelBeamWidth", "STT_SetModelBeamWidth",
    "STT_GetModelSampleRate", "STT_FreeModel", "STT_EnableExternalScorer", "STT_EnableExternalScorerFromBuffer",
    "STT_AddHotWord", "STT_EraseHotWord", "STT_ClearHotWords", "STT_DisableExternalScorer", "STT_SetScorerAlphaBeta",
    "STT_SpeechToText", "STT_SpeechToTextWithMetadata", "STT_SpeechToTextWithEmissions", "STT_CreateStream",
    "STT_FeedAudioContent", "STT_IntermediateDecode", "STT_IntermediateDecodeWithMetadata",
    "STT_IntermediateDecodeFlushBuffers", "STT_IntermediateDecodeWithMetadataFlushBuffers", "STT_FinishStream",
    "STT_FinishStreamWithMetadata", "STT_FreeStream", "STT_FreeMetadata", "STT_FreeString", "STT_Version",
    "STT_ErrorCodeToErrorMessage",
    "STTX_SpeechToTextBatch", "STTX_BatchCreate", "STTX_BatchFree", "STTX_BatchUpload", "STTX_BatchForward",
    "STTX_BatchDecode", "STTX_BatchNumResults", "STTX_BatchTranscript", "STTX_BatchTokens", "STTX_BatchFetch",
    "STTX_BatchGetTimings", "STTX_BatchKernelLaunches", "STTX_BatchSetInstrumented", "STTX_BatchTimesteps", "STTX_BatchCopyFeatures",
    "STTX_BatchCopyProbs", "STTX_BatchSetProbs", "STTX_ModelInfo", "STTX_BatchLmStats", "STTX_BatchDecoderScalars", "STTX_BatchSetCutoff", "STTX_BatchPhaseCycles", "STTX_BatchHostBuffer", "STTX_BatchLstmProfile", "STTX_BatchSetProbs64", "STTX_InspectModel", "STTX_InspectModelTensor", "STTX_StreamArenaCompactions",
]


_dev = None


def dev_lib():
    """The unit-test build of the same sources (libstt_b200_dev.so = + STTX_DebugGemm / STTX_DebugPairLayout).  Test
    infrastructure: nothing in this package calls it."""
    global _dev
    if _dev is None:
        p = os.path.join(_HERE, "libstt_b200_dev.so")
        if not os.path.exists(p):
            raise STTError("%s is missing: run `make`" % p)
        L = ctypes.CDLL(p)
        L.STTX_DebugPairLayout.argtypes = [c_int, c_void_p]
        L.STTX_DebugGemm.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                     POINTER(c_float)]
        _dev = L
    return _dev


def lib():
    """Load libstt_b200.so (built in-tree by `make` / __graft_entry__.build()).  Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise STTError("%s is missing: run `make` (or __graft_entry__.build()) first; there is no CPU fallback" % p)
    L = ctypes.CDLL(p)
    vp = c_void_p
    L.STT_CreateModel.argtypes = [c_char_p, POINTER(vp)]
    L.STT_CreateModelFromBuffer.argtypes = [c_char_p, c_uint, POINTER(vp)]
    L.STT_GetModelBeamWidth.argtypes = [vp]
    L.STT_GetModelBeamWidth.restype = c_uint
    L.STT_SetModelBeamWidth.argtypes = [vp, c_uint]
    L.STT_GetModelSampleRate.argtypes = [vp]
    L.STT_FreeModel.argtypes = [vp]
    L.STT_FreeModel.restype = None
    L.STT_EnableExternalScorer.argtypes = [vp, c_char_p]
    L.STT_EnableExternalScorerFromBuffer.argtypes = [vp, c_char_p, c_uint]
    L.STT_AddHotWord.argtypes = [vp, c_char_p, c_float]
    L.STT_EraseHotWord.argtypes = [vp, c_char_p]
    L.STT_ClearHotWords.argtypes = [vp]
    L.STT_DisableExternalScorer.argtypes = [vp]
    L.STT_SetScorerAlphaBeta.argtypes = [vp, c_float, c_float]
    L.STT_SpeechToText.argtypes = [vp, c_void_p, c_uint]
    L.STT_SpeechToText.restype = c_void_p
    for name in ("STT_SpeechToTextWithMetadata", "STT_SpeechToTextWithEmissions"):
        getattr(L, name).argtypes = [vp, c_void_p, c_uint, c_uint]
        getattr(L, name).restype = POINTER(_Metadata)
    L.STT_CreateStream.argtypes = [vp, POINTER(vp)]
    L.STT_FeedAudioContent.argtypes = [vp, c_void_p, c_uint]
    L.STT_FeedAudioContent.restype = None
    for name in ("STT_IntermediateDecode", "STT_IntermediateDecodeFlushBuffers", "STT_FinishStream"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = c_void_p
    for name in ("STT_IntermediateDecodeWithMetadata", "STT_IntermediateDecodeWithMetadataFlushBuffers",
                 "STT_FinishStreamWithMetadata"):
        getattr(L, name).argtypes = [vp, c_uint]
        getattr(L, name).restype = POINTER(_Metadata)
    L.STT_FreeStream.argtypes = [vp]
    L.STT_FreeStream.restype = None
    L.STT_FreeMetadata.argtypes = [POINTER(_Metadata)]
    L.STT_FreeMetadata.restype = None
    L.STT_FreeString.argtypes = [c_void_p]
    L.STT_FreeString.restype = None
    L.STT_Version.restype = c_void_p
    L.STT_ErrorCodeToErrorMessage.argtypes = [c_int]
    L.STT_ErrorCodeToErrorMessage.restype = c_void_p
    # extension
    L.STTX_SpeechToTextBatch.argtypes = [vp, POINTER(c_void_p), POINTER(c_uint), c_uint, POINTER(c_void_p)]
    L.STTX_BatchCreate.argtypes = [vp, c_uint, c_uint, POINTER(vp)]
    L.STTX_BatchFree.argtypes = [vp]
    L.STTX_BatchFree.restype = None
    L.STTX_BatchUpload.argtypes = [vp, POINTER(c_void_p), POINTER(c_uint), c_uint]
    L.STTX_BatchHostBuffer.argtypes = [vp, c_uint]
    L.STTX_BatchHostBuffer.restype = c_void_p
    L.STTX_BatchForward.argtypes = [vp]
    L.STTX_BatchDecode.argtypes = [vp, c_uint]
    L.STTX_BatchFetch.argtypes = [vp]
    L.STTX_BatchNumResults.argtypes = [vp, c_uint]
    L.STTX_BatchTranscript.argtypes = [vp, c_uint, c_uint]
    L.STTX_BatchTranscript.restype = c_void_p
    L.STTX_BatchTokens.argtypes = [vp, c_uint, c_uint, c_void_p, c_void_p, c_uint, POINTER(c_double)]
    L.STTX_BatchGetTimings.argtypes = [vp, POINTER(_Timings)]
    L.STTX_BatchKernelLaunches.argtypes = [vp]
    L.STTX_BatchSetInstrumented.argtypes = [vp, c_int]
    L.STTX_StreamArenaCompactions.argtypes = [vp]
    L.STTX_StreamArenaCompactions.restype = c_longlong
    L.STTX_InspectModel.argtypes = [c_char_p, c_uint, POINTER(c_uint), POINTER(c_float)]
    L.STTX_InspectModelTensor.argtypes = [c_char_p, c_uint, c_char_p, c_void_p, ctypes.c_ulonglong]
    L.STTX_InspectModelTensor.restype = c_longlong
    L.STTX_BatchKernelLaunches.restype = c_longlong
    L.STTX_BatchTimesteps.argtypes = [vp, c_uint]
    L.STTX_BatchPhaseCycles.argtypes = [vp, POINTER(ctypes.c_ulonglong)]
    L.STTX_BatchLstmProfile.argtypes = [vp, POINTER(ctypes.c_ulonglong)]
    L.STTX_BatchLmStats.argtypes = [vp, POINTER(ctypes.c_ulonglong), POINTER(ctypes.c_ulonglong)]
    L.STTX_BatchDecoderScalars.argtypes = [vp, POINTER(ctypes.c_ulonglong)]
    L.STTX_BatchSetCutoff.argtypes = [vp, ctypes.c_double, c_uint]
    L.STTX_BatchCopyFeatures.argtypes = [vp, c_uint, c_void_p]
    L.STTX_BatchCopyProbs.argtypes = [vp, c_uint, c_void_p]
    L.STTX_BatchSetProbs.argtypes = [vp, c_void_p, c_void_p, c_uint, c_uint]
    L.STTX_BatchSetProbs64.argtypes = [vp, c_void_p, c_void_p, c_uint, c_uint]
    L.STTX_ModelInfo.argtypes = [vp] + [POINTER(c_uint)] * 5
    _lib = L
    return L


def _take_string(ptr):
    if not ptr:
        return None
    s = ctypes.cast(ptr, c_char_p).value.decode("utf-8", errors="replace")
    lib().STT_FreeString(ptr)
    return s


def _err(code):
    return _take_string(lib().STT_ErrorCodeToErrorMessage(code))


def version():
    return _take_string(lib().STT_Version())


def _as_i16(audio):
    a = np.ascontiguousarray(audio, dtype=np.int16)
    return a, a.ctypes.data_as(c_void_p), a.size


class TokenMetadata(object):
    def __init__(self, text, timestep, start_time):
        self.text, self.timestep, self.start_time = text, timestep, start_time


class CandidateTranscript(object):
    def __init__(self, tokens, confidence):
        self.tokens, self.confidence = tokens, confidence


class Metadata(object):
    def __init__(self, transcripts, emissions=None, symbols=None):
        self.transcripts, self.emissions, self.symbols = transcripts, emissions, symbols


def _take_metadata(ptr):
    if not ptr:
        return None
    m = ptr.contents
    transcripts = []
    for i in range(m.num_transcripts):
        ct = m.transcripts[i]
        toks = [TokenMetadata(ct.tokens[j].text.decode("utf-8", errors="surrogateescape"), ct.tokens[j].timestep, ct.tokens[j].start_time)
                for j in range(ct.num_tokens)]
        transcripts.append(CandidateTranscript(toks, ct.confidence))
    emissions = symbols = None
    if m.emissions:
        e = m.emissions.contents
        n = e.num_symbols + 1
        emissions = np.ctypeslib.as_array(e.emissions, shape=(e.num_timesteps, n)).copy() if e.num_timesteps else \
            np.zeros((0, n))
        symbols = [e.symbols[i].decode("utf-8", errors="surrogateescape") for i in range(n)]
    lib().STT_FreeMetadata(ptr)
    return Metadata(transcripts, emissions, symbols)


class Model(object):
    """native_client/python/__init__.py:26-221"""

    def __init__(self, model_path):
        self._impl = None
        impl = c_void_p()
        if isinstance(model_path, (bytes, bytearray)) and not os.path.exists(model_path):
            status = lib().STT_CreateModelFromBuffer(bytes(model_path), len(model_path), byref(impl))
        else:
            p = model_path if isinstance(model_path, bytes) else str(model_path).encode()
            status = lib().STT_CreateModel(p, byref(impl))
        if status != 0:
            raise STTError("CreateModel failed with '{}' (0x{:X})".format(_err(status), status))
        self._impl = impl

    def __del__(self):
        if getattr(self, "_impl", None):
            lib().STT_FreeModel(self._impl)
            self._impl = None

    def beamWidth(self):
        return lib().STT_GetModelBeamWidth(self._impl)

    def setBeamWidth(self, beam_width):
        return lib().STT_SetModelBeamWidth(self._impl, beam_width)

    def sampleRate(self):
        return lib().STT_GetModelSampleRate(self._impl)

    def enableExternalScorer(self, scorer_path):
        status = lib().STT_EnableExternalScorer(self._impl, str(scorer_path).encode())
        if status != 0:
            raise STTError("EnableExternalScorer failed with '{}' (0x{:X})".format(_err(status), status))

    def disableExternalScorer(self):
        return lib().STT_DisableExternalScorer(self._impl)

    def addHotWord(self, word, boost):
        status = lib().STT_AddHotWord(self._impl, word.encode(), boost)
        if status != 0:
            raise STTError("AddHotWord failed with '{}' (0x{:X})".format(_err(status), status))

    def eraseHotWord(self, word):
        status = lib().STT_EraseHotWord(self._impl, word.encode())
        if status != 0:
            raise STTError("EraseHotWord failed with '{}' (0x{:X})".format(_err(status), status))

    def clearHotWords(self):
        status = lib().STT_ClearHotWords(self._impl)
        if status != 0:
            raise STTError("ClearHotWords failed with '{}' (0x{:X})".format(_err(status), status))

    def setScorerAlphaBeta(self, alpha, beta):
        return lib().STT_SetScorerAlphaBeta(self._impl, alpha, beta)

    def stt(self, audio_buffer):
        a, p, n = _as_i16(audio_buffer)
        return _take_string(lib().STT_SpeechToText(self._impl, p, n))

    def sttWithMetadata(self, audio_buffer, num_results=1):
        a, p, n = _as_i16(audio_buffer)
        return _take_metadata(lib().STT_SpeechToTextWithMetadata(self._impl, p, n, num_results))

    def sttWithEmissions(self, audio_buffer, num_results=1):
        a, p, n = _as_i16(audio_buffer)
        return _take_metadata(lib().STT_SpeechToTextWithEmissions(self._impl, p, n, num_results))

    def createStream(self):
        ctx = c_void_p()
        status = lib().STT_CreateStream(self._impl, byref(ctx))
        if status != 0:
            raise STTError("CreateStream failed with '{}' (0x{:X})".format(_err(status), status))
        return Stream(ctx, self)

    # ---- additive (STTX_*)
    def sttBatch(self, audio_buffers):
        """STTX_SpeechToTextBatch: list of int16 arrays -> list of transcripts (host buffers in, strings out)."""
        arrs = [np.ascontiguousarray(a, dtype=np.int16) for a in audio_buffers]
        n = len(arrs)
        ptrs = (c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (c_uint * n)(*[a.size for a in arrs])
        outs = (c_void_p * n)()
        status = lib().STTX_SpeechToTextBatch(self._impl, ptrs, lens, n, outs)
        if status != 0:
            raise STTError("SpeechToTextBatch failed with '{}' (0x{:X})".format(_err(status), status))
        return [_take_string(outs[i]) for i in range(n)]

    def createBatch(self, max_utterances, max_samples):
        return Batch(self, max_utterances, max_samples)

    def info(self):
        v = [c_uint() for _ in range(5)]
        lib().STTX_ModelInfo(self._impl, *[byref(x) for x in v])
        return dict(zip(("n_classes", "n_input", "n_hidden", "n_steps", "n_sms"), [x.value for x in v]))


def inspect_model(data):
    """Parse model-file bytes (.tflite flatbuffer or .sttw) without a device; returns (info dict, tensors dict)."""
    buf = bytes(data)
    info = (c_uint * 12)()
    clip = c_float()
    status = lib().STTX_InspectModel(buf, len(buf), info, byref(clip))
    if status != 0:
        raise STTError("model file rejected: {} (0x{:X})".format(_err(status), status))
    names = ("sample_rate", "win_len", "win_step", "n_input", "n_context", "n_hidden", "n_cell", "n_classes", "n_steps",
             "beam_width", "space_label", "n_labels")
    out = dict(zip(names, [int(x) for x in info]))
    out["relu_clip"] = clip.value
    tensors = {}
    for n in ("w1", "b1", "w2", "b2", "w3", "b3", "lstm_kernel", "lstm_bias", "w5", "b5", "w6", "b6"):
        cnt = lib().STTX_InspectModelTensor(buf, len(buf), n.encode(), None, 0)
        arr = np.zeros(cnt, np.float32)
        lib().STTX_InspectModelTensor(buf, len(buf), n.encode(), arr.ctypes.data, cnt)
        tensors[n] = arr
    return out, tensors


class Stream(object):
    """native_client/python/__init__.py:223-383"""

    def __init__(self, native_stream, model=None):
        self._impl = native_stream
        self._model = model   # the stream's device context belongs to the model: keep it alive as long as the stream

    def arenaCompactions(self):
        """How often this stream's decoder arena has been garbage-collected so far (STTX_StreamArenaCompactions)."""
        self._check()
        return lib().STTX_StreamArenaCompactions(self._impl)

    def __del__(self):
        if getattr(self, "_impl", None):
            self.freeStream()

    def _check(self):
        if not self._impl:
            raise RuntimeError("Stream object is not valid. Trying to feed an already finished stream?")

    def feedAudioContent(self, audio_buffer):
        self._check()
        a, p, n = _as_i16(audio_buffer)
        lib().STT_FeedAudioContent(self._impl, p, n)

    def intermediateDecode(self):
        self._check()
        return _take_string(lib().STT_IntermediateDecode(self._impl))

    def intermediateDecodeWithMetadata(self, num_results=1):
        self._check()
        return _take_metadata(lib().STT_IntermediateDecodeWithMetadata(self._impl, num_results))

    def intermediateDecodeFlushBuffers(self):
        self._check()
        return _take_string(lib().STT_IntermediateDecodeFlushBuffers(self._impl))

    def intermediateDecodeWithMetadataFlushBuffers(self, num_results=1):
        self._check()
        return _take_metadata(lib().STT_IntermediateDecodeWithMetadataFlushBuffers(self._impl, num_results))

    def finishStream(self):
        self._check()
        result = _take_string(lib().STT_FinishStream(self._impl))
        self._impl = None
        return result

    def finishStreamWithMetadata(self, num_results=1):
        self._check()
        result = _take_metadata(lib().STT_FinishStreamWithMetadata(self._impl, num_results))
        self._impl = None
        return result

    def freeStream(self):
        self._check()
        lib().STT_FreeStream(self._impl)
        self._impl = None


class Batch(object):
    """Staged batch context (STTX_Batch*): upload -> forward -> decode -> fetch, with device timings."""

    def __init__(self, model, max_utterances, max_samples):
        self._model = model
        self._impl = c_void_p()
        status = lib().STTX_BatchCreate(model._impl, max_utterances, max_samples, byref(self._impl))
        if status != 0:
            self._impl = None
            raise STTError("BatchCreate failed with '{}' (0x{:X})".format(_err(status), status))
        self._keep = None
        self.n = 0

    def __del__(self):
        if getattr(self, "_impl", None):
            lib().STTX_BatchFree(self._impl)
            self._impl = None

    def _ok(self, status, what):
        if status != 0:
            raise STTError("{} failed with '{}' (0x{:X})".format(what, _err(status), status))

    def upload(self, audio_buffers):
        arrs = [np.ascontiguousarray(a, dtype=np.int16) for a in audio_buffers]
        n = len(arrs)
        ptrs = (c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (c_uint * n)(*[a.size for a in arrs])
        self._keep = arrs
        self.n = n
        self._ok(lib().STTX_BatchUpload(self._impl, ptrs, lens, n), "BatchUpload")

    def host_buffer(self, u, n_samples):
        """numpy int16 view of utterance u's PINNED staging row; pass it to upload() for a copy-free H2D."""
        ptr = lib().STTX_BatchHostBuffer(self._impl, u)
        return np.ctypeslib.as_array(ctypes.cast(ptr, POINTER(c_short)), shape=(n_samples,))

    def forward(self):
        self._ok(lib().STTX_BatchForward(self._impl), "BatchForward")

    def decode(self, num_results=1):
        self._ok(lib().STTX_BatchDecode(self._impl, num_results), "BatchDecode")

    def fetch(self):
        self._ok(lib().STTX_BatchFetch(self._impl), "BatchFetch")

    def transcripts(self):
        return [_take_string(lib().STTX_BatchTranscript(self._impl, u, 0)) for u in range(self.n)]

    def results(self, u, max_tokens=4096):
        out = []
        for r in range(lib().STTX_BatchNumResults(self._impl, u)):
            tok = np.zeros(max_tokens, np.uint32)
            ts = np.zeros(max_tokens, np.uint32)
            conf = c_double()
            n = lib().STTX_BatchTokens(self._impl, u, r, tok.ctypes.data, ts.ctypes.data, max_tokens, byref(conf))
            out.append((conf.value, tok[:n].copy(), ts[:n].copy()))
        return out

    def timings(self):
        t = _Timings()
        lib().STTX_BatchGetTimings(self._impl, byref(t))
        return {n: getattr(t, n) for n, _ in _Timings._fields_}

    def phase_cycles(self):
        arr = (ctypes.c_ulonglong * 8)()
        lib().STTX_BatchPhaseCycles(self._impl, arr)
        names = ("cutoff", "parent_lookup", "lm", "live_update", "children", "select", "commit_new_nodes", "compact")
        return dict(zip(names, [int(x) for x in arr]))

    def lstm_profile(self):
        arr = (ctypes.c_ulonglong * 3)()
        lib().STTX_BatchLstmProfile(self._impl, arr)
        return dict(zip(("grid_barrier_wait", "load_mma_span", "epilogue"), [int(x) for x in arr]))

    def lm_stats(self):
        w, c = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        lib().STTX_BatchLmStats(self._impl, byref(w), byref(c))
        return {"words_scored": w.value, "lm_calls": c.value}

    def set_cutoff(self, cutoff_prob=1.0, cutoff_top_n=40):
        """Vocabulary pruning of the following decodes (the decoder-only surface's cutoff_prob / cutoff_top_n)."""
        self._ok(lib().STTX_BatchSetCutoff(self._impl, float(cutoff_prob), int(cutoff_top_n)), "BatchSetCutoff")

    def decoder_scalars(self):
        arr = (ctypes.c_ulonglong * 16)()
        lib().STTX_BatchDecoderScalars(self._impl, arr)
        return [int(x) for x in arr]

    def kernel_launches(self):
        return lib().STTX_BatchKernelLaunches(self._impl)

    def set_instrumented(self, on=True):
        """Following decode() calls run the statistics build of the decoder kernel (phase_cycles / lm_stats)."""
        lib().STTX_BatchSetInstrumented(self._impl, 1 if on else 0)

    def timesteps(self, u):
        return lib().STTX_BatchTimesteps(self._impl, u)

    def features(self, u):
        info = self._model.info()
        T = self.timesteps(u)
        out = np.zeros((T, info["n_input"]), np.float32)
        lib().STTX_BatchCopyFeatures(self._impl, u, out.ctypes.data)
        return out

    def probs(self, u):
        info = self._model.info()
        T = self.timesteps(u)
        out = np.zeros((T, info["n_classes"]), np.float32)
        lib().STTX_BatchCopyProbs(self._impl, u, out.ctypes.data)
        return out

    def set_probs64(self, probs, lengths):
        """probs: float64 [B, T_stride, C] (the Python decoder API's input type)."""
        p = np.ascontiguousarray(probs, dtype=np.float64)
        t = np.ascontiguousarray(lengths, dtype=np.int32)
        self.n = p.shape[0]
        self._ok(lib().STTX_BatchSetProbs64(self._impl, p.ctypes.data, t.ctypes.data, p.shape[0], p.shape[1]),
                 "BatchSetProbs64")

    def set_probs(self, probs, lengths):
        """probs: float32 [B, T_stride, C]; lengths: per-utterance T.  Decoder-only parity tests."""
        p = np.ascontiguousarray(probs, dtype=np.float32)
        t = np.ascontiguousarray(lengths, dtype=np.int32)
        self.n = p.shape[0]
        self._ok(lib().STTX_BatchSetProbs(self._impl, p.ctypes.data, t.ctypes.data, p.shape[0], p.shape[1]),
                 "BatchSetProbs")


class BatchPipeline(object):
    """`depth` staged batch contexts driven by one host thread each, so that batch i+1's pinned staging and H2D copy
    (and the host-side result parsing of batch i-1) overlap batch i's kernels.  Every batch still pays its own
    upload -> forward -> decode -> fetch; only the waiting is shared.  Results come back in submission order."""

    def __init__(self, model, max_utterances, max_samples, depth=2):
        self.batches = [Batch(model, max_utterances, max_samples) for _ in range(depth)]

    def host_buffers(self, slot, n_utt, n_samples):
        return [self.batches[slot].host_buffer(u, n_samples) for u in range(n_utt)]

    @staticmethod
    def _one(bt, audio_buffers, num_results):
        bt.upload(audio_buffers)
        bt.forward()
        bt.decode(num_results)
        bt.fetch()
        return bt.transcripts()

    def map(self, batches_of_audio, num_results=1):
        """batches_of_audio: sequence of lists of int16 arrays; returns one transcript list per batch, in order.
        Batch i runs on context i % depth, so callers that pre-fill pinned rows use host_buffers(i % depth, ...)."""
        import threading
        items = list(batches_of_audio)
        out = [None] * len(items)
        errs = []
        depth = len(self.batches)

        def worker(k):
            try:
                for i in range(k, len(items), depth):
                    out[i] = self._one(self.batches[k], items[i], num_results)
            except Exception as ex:  # surfaced to the caller below
                errs.append(ex)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(min(depth, len(items)))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        return out
