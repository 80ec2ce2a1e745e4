"""The reference's model container: `.tflite` flatbuffers (SURVEY 8b "model-file contract", 8f rank 1).

  * stt_b200/tflite_export.py writes files with the structure of the reference's exporter; an INDEPENDENT walk of the
    flatbuffer wire format below (second implementation, test side) checks that they are well formed: every scalar
    aligned to its size, vtables consistent, strings terminated, buffers 16-byte aligned, offsets inside the file;
  * stt_b200/csrc/tflite_reader.cc (through STTX_InspectModel, no device needed) recovers geometry, metadata, alphabet
    and weights: float32 exactly, float16 to half precision, int8 = scale * q exactly as the hybrid kernels dequantise;
  * error behaviour of TFLiteModelState::init (tflitemodelstate.cc:264-330): old graph version -> MODEL_INCOMPATIBLE,
    alphabet/logits mismatch -> INVALID_ALPHABET, garbage -> FAIL_INIT_MMAP; mutated and truncated files never crash.
"""
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def weights():
    from stt_b200 import synth
    return synth.make_weights(n_hidden=64, seed=11)


# ------------------------------------------------------------------------------------------------ independent verifier
class _Walk(object):
    """Minimal flatbuffer walker written from the format description (FlatBuffers internals), not from the writer."""

    def __init__(self, b):
        self.b = b

    def u(self, fmt, pos):
        size = struct.calcsize(fmt)
        assert 0 <= pos and pos + size <= len(self.b), "read outside the file"
        assert pos % size == 0, "scalar at %d is not aligned to %d" % (pos, size)
        return struct.unpack_from(fmt, self.b, pos)[0]

    def table(self, pos):
        soff = self.u("<i", pos)
        vt = pos - soff
        vsize, tsize = self.u("<H", vt), self.u("<H", vt + 2)
        assert vsize >= 4 and vsize % 2 == 0 and tsize >= 4
        offs = [self.u("<H", vt + 4 + 2 * i) for i in range((vsize - 4) // 2)]
        assert all(o == 0 or 4 <= o < tsize for o in offs), "field offset outside its table"
        return offs

    def field(self, pos, offs, i):
        return pos + offs[i] if i < len(offs) and offs[i] else None

    def ref(self, fpos):
        o = self.u("<I", fpos)
        assert o > 0
        return fpos + o

    def vec(self, pos):
        return self.u("<I", pos), pos + 4

    def string(self, pos):
        n, d = self.vec(pos)
        assert self.b[d + n] == 0, "string is not NUL terminated"
        return bytes(self.b[d:d + n]).decode("utf-8")


def _verify(b):
    w = _Walk(b)
    assert b[4:8] == b"TFL3"
    model = w.u("<I", 0)
    mo = w.table(model)
    assert w.u("<I", w.field(model, mo, 0)) == 3                       # Model.version
    n_buf, bufs = w.vec(w.ref(w.field(model, mo, 4)))                  # Model.buffers
    sizes = []
    for i in range(n_buf):
        bt = w.ref(bufs + 4 * i)
        bo = w.table(bt)
        f = w.field(bt, bo, 0)
        if f is None:
            sizes.append(0)
            continue
        n, d = w.vec(w.ref(f))
        assert d % 16 == 0, "buffer %d data is not 16-byte aligned" % i  # schema.fbs:1197-1199 force_align
        assert d + n <= len(b)
        sizes.append(n)
    assert sizes[0] == 0, "buffer 0 must be the empty sentinel"
    n_sg, sgs = w.vec(w.ref(w.field(model, mo, 2)))
    assert n_sg == 1
    sg = w.ref(sgs)
    so = w.table(sg)
    n_t, tens = w.vec(w.ref(w.field(sg, so, 0)))
    names, elem = {}, {0: 4, 1: 2, 2: 4, 9: 1}
    for i in range(n_t):
        t = w.ref(tens + 4 * i)
        to = w.table(t)
        n_d, dims = w.vec(w.ref(w.field(t, to, 0)))
        shape = [w.u("<i", dims + 4 * k) for k in range(n_d)]
        ttype = w.u("<b", w.field(t, to, 1)) if w.field(t, to, 1) else 0
        buf = w.u("<I", w.field(t, to, 2)) if w.field(t, to, 2) else 0
        name = w.string(w.ref(w.field(t, to, 3)))
        assert buf < n_buf
        if buf and ttype in elem:
            assert sizes[buf] == int(np.prod(shape)) * elem[ttype] if shape else elem[ttype], name
        q = w.field(t, to, 4)
        if q is not None:
            qt = w.ref(q)
            qo = w.table(qt)
            n_s, sc = w.vec(w.ref(w.field(qt, qo, 2)))
            n_z, zp = w.vec(w.ref(w.field(qt, qo, 3)))
            assert n_s == n_z == 1 and w.u("<f", sc) > 0 and w.u("<q", zp) == 0
        names[name] = (i, shape, ttype)
    n_op, ops = w.vec(w.ref(w.field(sg, so, 3)))
    n_codes, _ = w.vec(w.ref(w.field(model, mo, 1)))
    for i in range(n_op):
        o = w.ref(ops + 4 * i)
        oo = w.table(o)
        code = w.u("<I", w.field(o, oo, 0)) if w.field(o, oo, 0) else 0
        assert code < n_codes
        for fid in (1, 2):
            n, d = w.vec(w.ref(w.field(o, oo, fid)))
            for k in range(n):
                assert -1 <= w.u("<i", d + 4 * k) < n_t
    return names


@pytest.mark.parametrize("weight_type", ["float32", "int8", "float16"])
@pytest.mark.parametrize("metadata_via_op", [True, False])
def test_written_files_are_wellformed_and_read_back(weights, weight_type, metadata_via_op):
    from stt_b200 import api, tflite_export
    b = tflite_export.model_bytes(weights, weight_type=weight_type, metadata_via_op=metadata_via_op, beam_width=321,
                                  relu_clip=17.5)
    names = _verify(b)
    # the names and shapes native_client/tflitemodelstate.cc:211-335 relies on
    assert names["input_node"][1] == [1, 16, 19, 26]
    assert names["previous_state_c"][1] == names["previous_state_h"][1] == [1, 64]
    assert names["logits"][1] == [16, 29] and names["input_samples"][1] == [512]
    for n in ("new_state_c", "new_state_h", "mfccs", "metadata_version", "metadata_sample_rate", "metadata_feature_win_len",
              "metadata_feature_win_step", "metadata_beam_width", "metadata_alphabet"):
        assert n in names
    info, tens = api.inspect_model(b)
    assert info == {"sample_rate": 16000, "win_len": 512, "win_step": 320, "n_input": 26, "n_context": 9, "n_hidden": 64,
                    "n_cell": 64, "n_classes": 29, "n_steps": 16, "beam_width": 321, "space_label": 0, "n_labels": 28,
                    "relu_clip": 17.5}
    for k, got in tens.items():
        want = np.asarray(weights[k], np.float32).reshape(-1)
        if weight_type == "float32" or k.startswith("b") or k == "lstm_bias":
            np.testing.assert_array_equal(got, want)                  # biases stay float32 in every export
        elif weight_type == "float16":
            np.testing.assert_array_equal(got, want.astype(np.float16).astype(np.float32))
        else:
            q, s = tflite_export.quantize_int8(np.asarray(weights[k], np.float32).T)
            np.testing.assert_array_equal(got, (q.astype(np.float32) * s).T.reshape(-1))   # f = scale * q
            assert np.abs(got - want).max() <= 0.5 * float(s) * 1.0001


def test_same_host_model_as_the_native_container(weights):
    """A float32 .tflite and a .sttw of the same weights parse to the same HostModel."""
    from stt_b200 import api, synth, tflite_export
    i1, t1 = api.inspect_model(tflite_export.model_bytes(weights))
    i2, t2 = api.inspect_model(synth.model_bytes(weights))
    assert i1 == i2
    for k in t1:
        np.testing.assert_array_equal(t1[k], t2[k])


def test_duplicated_lstm_kernel_and_other_alphabets(weights):
    from stt_b200 import api, synth, tflite_export
    b = tflite_export.model_bytes(weights, duplicate_lstm_kernel=True)
    _verify(b)
    _, t = api.inspect_model(b)
    np.testing.assert_array_equal(t["lstm_kernel"], weights["lstm_kernel"].reshape(-1))
    labels = ["a", "b", "é", " ", "中"]                     # space not first, multi-byte labels
    w = synth.make_weights(n_hidden=64, n_classes=len(labels) + 1, seed=3)
    info, _ = api.inspect_model(tflite_export.model_bytes(w, labels=labels, sample_rate=8000))
    assert (info["n_classes"], info["space_label"], info["n_labels"]) == (6, 3, 5)
    assert (info["sample_rate"], info["win_len"], info["win_step"]) == (8000, 256, 160)


def test_error_codes(weights):
    from stt_b200 import api, tflite_export
    L = api.lib()

    def status(b):
        return L.STTX_InspectModel(bytes(b), len(b), None, None)

    assert status(tflite_export.model_bytes(weights, graph_version=5)) == 0x2003          # STT_ERR_MODEL_INCOMPATIBLE
    assert status(tflite_export.model_bytes(weights, labels=["a", "b", " "])) == 0x2000   # STT_ERR_INVALID_ALPHABET
    assert status(b"\x00" * 64) == 0x3000                                                  # STT_ERR_FAIL_INIT_MMAP
    junk = bytearray(64)
    junk[4:8] = b"TFL3"
    assert status(junk) == 0x3000


def test_mutated_files_never_crash(weights):
    """Truncations and byte flips: the reader may accept or reject, but must stay inside the buffer (every access is
    bounds checked; the run itself is the assertion -- a wild read would take the interpreter down)."""
    from stt_b200 import api, tflite_export
    L = api.lib()
    good = tflite_export.model_bytes(weights, weight_type="int8")
    rng = np.random.default_rng(123)
    outcomes = set()
    for cut in list(range(0, 4096, 97)) + [len(good) // 2, len(good) - 1]:
        outcomes.add(L.STTX_InspectModel(good[:cut], cut, None, None))
    head = 6000   # tables and vtables sit in front of the big buffers
    for _ in range(400):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, head))] = int(rng.integers(0, 256))
        outcomes.add(L.STTX_InspectModel(bytes(b), len(b), None, None))
    assert 0x3000 in outcomes
