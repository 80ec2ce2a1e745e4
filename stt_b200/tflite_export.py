"""Writer of `.tflite` model files with the structure the reference's exporter produces
(training/coqui_stt_training/export.py:40-150 over create_inference_graph(batch_size=1, n_steps=16, tflite=True),
deepspeech_model.py:266-403), so that acoustic-model weights held as arrays -- synthetic ones in the tests, or a
checkpoint converted offline -- can be handed to anything that opens Coqui STT model files, this library included.

TensorFlow is not available here, so the flatbuffer is emitted directly against tensorflow/lite/schema/schema.fbs
(field numbers cited below).  What is written:

  tensors   input_node [1, n_steps, 2c+1, n_input], previous_state_c/h [1, n_cell], input_samples [win_len],
            logits [n_steps, K], new_state_c/h, mfccs, metadata_version/_sample_rate/_feature_win_len/_feature_win_step/
            _beam_width (int32 [1]) and metadata_alphabet (string, Alphabet::Serialize bytes) -- the names and shapes
            native_client/tflitemodelstate.cc:211-335 looks up;
  operators AUDIO_SPECTROGRAM + MFCC custom ops (feeding.py:51-72), RESHAPE, three FULLY_CONNECTED(fused RELU) + MINIMUM(clip),
            the LSTM unrolled n_steps times (CONCATENATION -> FULLY_CONNECTED on the shared [4C, H+C] kernel -> SPLIT ->
            LOGISTIC/TANH/MUL/ADD, rnn_cell_impl.py:1054-1079), PACK, FULLY_CONNECTED + MINIMUM, FULLY_CONNECTED, SOFTMAX;
  weights   "float32", "int8" (what `export_quantize` = Optimize.DEFAULT produces: per-tensor symmetric int8 with
            QuantizationParameters.scale, biases float32) or "float16" (constants behind DEQUANTIZE operators).

`metadata_via_op=True` routes each metadata constant through an operator (the reference looks for the metadata tensors'
parent nodes, tflitemodelstate.cc:234-240); False leaves them as plain constant outputs.  Both are valid TFLite.
"""
import struct

import numpy as np

from .synth import ENGLISH_LABELS, serialize_alphabet

# ---------------------------------------------------------------------------------------------- flatbuffer objects
_SCALAR = {"bool": ("<?", 1), "i8": ("<b", 1), "u8": ("<B", 1), "i32": ("<i", 4), "u32": ("<I", 4), "f32": ("<f", 4),
           "i64": ("<q", 8)}


class _Str(object):
    def __init__(self, s):
        self.b = s if isinstance(s, bytes) else s.encode("utf-8")


class _Vec(object):
    """kind: a scalar kind, "bytes" (ubyte payload given as bytes, optional alignment) or "off" (offsets to objects)."""

    def __init__(self, kind, items, align=None):
        self.kind, self.items, self.align = kind, items, align


class _Table(object):
    def __init__(self, fields):
        """fields: {field id: (kind, value)}; kind "off" -> value is an object (or None = absent)."""
        self.fields = {k: v for k, v in fields.items() if not (v[0] == "off" and v[1] is None)}


def _align_up(x, a):
    return (x + a - 1) // a * a


def _finish(root, identifier=b"TFL3"):
    """Lay the object graph out front to back (parents before children, so every uoffset points forward) and emit it."""
    pos = [8]           # u32 root offset + 4-byte file identifier
    placed = []         # (object, position, extra)

    def place(obj):
        if isinstance(obj, _Str):
            p = _align_up(pos[0], 4)
            pos[0] = p + 4 + len(obj.b) + 1
            placed.append((obj, p, None))
            return p
        if isinstance(obj, _Vec):
            if obj.kind == "bytes":
                a, size = (obj.align or 1), len(obj.items)
            elif obj.kind == "off":
                a, size = 4, 4 * len(obj.items)
            else:
                a = _SCALAR[obj.kind][1]
                size = a * len(obj.items)
            a = max(a, 4)
            p = _align_up(pos[0] + 4, a) - 4        # the data (after the u32 length) is aligned to `a`
            pos[0] = p + 4 + size
            entry = [obj, p, None]
            placed.append(entry)
            if obj.kind == "off":
                entry[2] = [place(c) for c in obj.items]
            return p
        # table: vtable immediately before the inline data
        ids = sorted(obj.fields)
        n_slots = (ids[-1] + 1) if ids else 0
        vt_size = 4 + 2 * n_slots
        sizes = {i: (4 if obj.fields[i][0] == "off" else _SCALAR[obj.fields[i][0]][1]) for i in ids}
        order = sorted(ids, key=lambda i: -sizes[i])
        off, layout = 4, {}
        for i in order:
            off = _align_up(off, sizes[i])
            layout[i] = off
            off += sizes[i]
        t_align = max([4] + list(sizes.values()))
        t_size = _align_up(off, 4)
        vt = _align_up(pos[0], 2)
        while (vt + vt_size) % t_align:
            vt += 2
        t = vt + vt_size
        pos[0] = t + t_size
        entry = [obj, t, {"vt": vt, "vt_size": vt_size, "t_size": t_size, "layout": layout, "children": {}}]
        placed.append(entry)
        for i in ids:
            if obj.fields[i][0] == "off":
                entry[2]["children"][i] = place(obj.fields[i][1])
        return t

    root_pos = place(root)
    buf = bytearray(_align_up(pos[0], 8))
    struct.pack_into("<I", buf, 0, root_pos)
    buf[4:8] = identifier
    for obj, p, extra in placed:
        if isinstance(obj, _Str):
            struct.pack_into("<I", buf, p, len(obj.b))
            buf[p + 4:p + 4 + len(obj.b)] = obj.b
        elif isinstance(obj, _Vec):
            struct.pack_into("<I", buf, p, len(obj.items))
            if obj.kind == "bytes":
                buf[p + 4:p + 4 + len(obj.items)] = obj.items
            elif obj.kind == "off":
                for k, cp in enumerate(extra):
                    fp = p + 4 + 4 * k
                    struct.pack_into("<I", buf, fp, cp - fp)
            else:
                fmt, sz = _SCALAR[obj.kind]
                for k, v in enumerate(obj.items):
                    struct.pack_into(fmt, buf, p + 4 + sz * k, v)
        else:
            vt, layout = extra["vt"], extra["layout"]
            struct.pack_into("<HH", buf, vt, extra["vt_size"], extra["t_size"])
            for i, off in layout.items():
                struct.pack_into("<H", buf, vt + 4 + 2 * i, off)
            struct.pack_into("<i", buf, p, p - vt)
            for i, off in layout.items():
                kind, val = obj.fields[i]
                if kind == "off":
                    struct.pack_into("<I", buf, p + off, extra["children"][i] - (p + off))
                else:
                    struct.pack_into(_SCALAR[kind][0], buf, p + off, val)
    return bytes(buf)


# ---------------------------------------------------------------------------------------------- TFLite graph
# BuiltinOperator (schema.fbs:230-364), TensorType (:36-56), BuiltinOptions union index of FullyConnectedOptions (:393-401)
ADD, CONCATENATION, DEQUANTIZE, FULLY_CONNECTED, LOGISTIC, MUL, RESHAPE, SOFTMAX, TANH, CUSTOM = 0, 2, 6, 9, 14, 18, 22, 25, 28, 32
SPLIT, MINIMUM, PACK, UNPACK = 49, 57, 83, 88
F32, F16, I32, STRING, I8 = 0, 1, 2, 5, 9
_OPT_FULLY_CONNECTED = 8
_ACT_NONE, _ACT_RELU = 0, 1


def quantize_int8(w):
    """Per-tensor symmetric int8 like the converter's weight quantiser: scale = max|w|/127, round half away from zero."""
    w = np.asarray(w, np.float32)
    rng = float(np.abs(w).max())
    if rng == 0.0:
        return np.zeros(w.shape, np.int8), np.float32(1.0)
    scale = np.float32(rng / 127.0)
    q = np.clip(np.sign(w) * np.floor(np.abs(w * np.float32(127.0 / rng)) + 0.5), -127, 127).astype(np.int8)
    return q, scale


class _GraphBuilder(object):
    def __init__(self):
        self.buffers = [b""]        # buffer 0 is the empty sentinel (schema.fbs:1253-1257)
        self.tensors = []
        self.ops = []
        self.codes = []

    def buffer(self, data):
        self.buffers.append(bytes(data))
        return len(self.buffers) - 1

    def tensor(self, name, shape, ttype=F32, data=None, scale=None):
        t = {"name": name, "shape": [int(x) for x in shape], "type": ttype, "buffer": self.buffer(data) if data is not None else 0,
             "scale": scale}
        self.tensors.append(t)
        return len(self.tensors) - 1

    def code(self, builtin, custom=None):
        key = (builtin, custom)
        if key not in self.codes:
            self.codes.append(key)
        return self.codes.index(key)

    def op(self, builtin, inputs, outputs, custom=None, fc_act=None):
        self.ops.append({"code": self.code(builtin, custom), "in": list(inputs), "out": list(outputs), "fc_act": fc_act})

    def finish(self, inputs, outputs, description):
        tensors = []
        for t in self.tensors:
            f = {0: ("off", _Vec("i32", t["shape"])), 1: ("i8", t["type"]), 2: ("u32", t["buffer"]), 3: ("off", _Str(t["name"]))}
            if t["scale"] is not None:   # QuantizationParameters: scale = 2, zero_point = 3 (schema.fbs:71-95)
                f[4] = ("off", _Table({2: ("off", _Vec("f32", [float(t["scale"])])), 3: ("off", _Vec("i64", [0]))}))
            tensors.append(_Table(f))
        ops = []
        for o in self.ops:
            f = {0: ("u32", o["code"]), 1: ("off", _Vec("i32", o["in"])), 2: ("off", _Vec("i32", o["out"]))}
            if o["fc_act"] is not None:  # builtin_options_type = 3, builtin_options = 4; FullyConnectedOptions.fused_activation_function = 0
                f[3] = ("u8", _OPT_FULLY_CONNECTED)
                f[4] = ("off", _Table({0: ("i8", o["fc_act"])}))
            ops.append(_Table(f))
        codes = []
        for builtin, custom in self.codes:   # OperatorCode: deprecated_builtin_code = 0, custom_code = 1, version = 2, builtin_code = 3
            f = {0: ("i8", min(builtin, 127)), 2: ("i32", 1), 3: ("i32", builtin)}
            if custom:
                f[1] = ("off", _Str(custom))
            codes.append(_Table(f))
        subgraph = _Table({0: ("off", _Vec("off", tensors)), 1: ("off", _Vec("i32", inputs)), 2: ("off", _Vec("i32", outputs)),
                           3: ("off", _Vec("off", ops)), 4: ("off", _Str("main"))})
        buffers = [_Table({0: ("off", _Vec("bytes", b, align=16))} if b else {}) for b in self.buffers]
        model = _Table({0: ("u32", 3), 1: ("off", _Vec("off", codes)), 2: ("off", _Vec("off", [subgraph])),
                        3: ("off", _Str(description)), 4: ("off", _Vec("off", buffers))})
        return _finish(model)


def _string_tensor_bytes(values):
    """TFLite string tensor buffer (tensorflow/lite/string_util.h): int32 count, int32 offsets[count+1], chars."""
    n = len(values)
    head = 4 * (n + 2)
    offs, blob = [head], b""
    for v in values:
        blob += v
        offs.append(head + len(blob))
    return struct.pack("<i", n) + b"".join(struct.pack("<i", o) for o in offs) + blob


def model_bytes(weights, labels=ENGLISH_LABELS, sample_rate=16000, win_len_ms=32, win_step_ms=20, n_input=26, n_context=9,
                n_steps=16, beam_width=500, relu_clip=20.0, weight_type="float32", graph_version=6, metadata_via_op=True,
                duplicate_lstm_kernel=False):
    """weights: dict in TF layout ([in, out]) as stt_b200.synth.make_weights returns.  Returns the .tflite bytes."""
    assert weight_type in ("float32", "int8", "float16")
    g = _GraphBuilder()
    H = weights["b1"].shape[0]
    C = weights["lstm_bias"].shape[0] // 4
    K = weights["b6"].shape[0]
    nf = 2 * n_context + 1
    win_len = int(sample_rate * (win_len_ms / 1000.0))

    def const_weight(name, w_in_out):
        """FULLY_CONNECTED wants [out, in]; returns the tensor index the operator reads."""
        wt = np.ascontiguousarray(np.asarray(w_in_out, np.float32).T)
        if weight_type == "float32":
            return g.tensor(name, wt.shape, F32, wt.astype("<f4").tobytes())
        if weight_type == "int8":
            q, s = quantize_int8(wt)
            return g.tensor(name, wt.shape, I8, q.tobytes(), scale=s)
        src = g.tensor(name + "_f16", wt.shape, F16, wt.astype("<f2").tobytes())
        dst = g.tensor(name + "_dequantized", wt.shape, F32)
        g.op(DEQUANTIZE, [src], [dst])
        return dst

    def const_f32(name, v):
        v = np.ascontiguousarray(np.asarray(v, np.float32))
        return g.tensor(name, v.shape, F32, v.astype("<f4").tobytes())

    # ---- placeholders (deepspeech_model.py:271-330)
    t_samples = g.tensor("input_samples", [win_len])
    t_input = g.tensor("input_node", [1, n_steps, nf, n_input])
    t_pc = g.tensor("previous_state_c", [1, C])
    t_ph = g.tensor("previous_state_h", [1, C])
    # ---- feature sub-graph: AudioSpectrogram + Mfcc custom operators (feeding.py:51-72) -> "mfccs"
    t_spec = g.tensor("AudioSpectrogram", [1, 1, win_len // 2 + 1])
    t_rate = g.tensor("Mfcc/sample_rate", [], I32, struct.pack("<i", sample_rate))
    t_mfcc = g.tensor("mfccs", [1, n_input])
    g.op(CUSTOM, [t_samples], [t_spec], custom="AudioSpectrogram")
    g.op(CUSTOM, [t_spec, t_rate], [t_mfcc], custom="Mfcc")
    # ---- layers 1-3 (deepspeech_model.py:66-89, 204-231)
    t_shape = g.tensor("Reshape/shape", [2], I32, struct.pack("<ii", n_steps, nf * n_input))
    t_x = g.tensor("Reshape", [n_steps, nf * n_input])
    g.op(RESHAPE, [t_input, t_shape], [t_x])
    t_clip = g.tensor("Minimum/y", [], F32, struct.pack("<f", relu_clip))
    cur = t_x
    for li, (wn, bn, n_out) in enumerate((("w1", "b1", H), ("w2", "b2", H), ("w3", "b3", H)), start=1):
        w = const_weight("layer_%d/weights/transpose" % li, weights[wn])
        b = const_f32("layer_%d/bias" % li, weights[bn])
        fc = g.tensor("layer_%d/Relu" % li, [n_steps, n_out])
        g.op(FULLY_CONNECTED, [cur, w, b], [fc], fc_act=_ACT_RELU)
        mn = g.tensor("layer_%d/Minimum" % li, [n_steps, n_out])
        g.op(MINIMUM, [fc, t_clip], [mn])
        cur = mn
    # ---- LSTM, unrolled (deepspeech_model.py:144-168; rnn_cell_impl.py:1054-1079)
    rows = [g.tensor("unstack:%d" % t, [1, H]) for t in range(n_steps)]
    g.op(UNPACK, [cur], rows)
    kern_name = "cudnn_lstm/rnn/multi_rnn_cell/cell_0/cudnn_compatible_lstm_cell/kernel/transpose"
    kern = const_weight(kern_name, weights["lstm_kernel"])
    kbias = const_f32("cudnn_lstm/rnn/multi_rnn_cell/cell_0/cudnn_compatible_lstm_cell/bias", weights["lstm_bias"])
    t_axis = g.tensor("split/split_dim", [], I32, struct.pack("<i", 1))
    c_prev, h_prev = t_pc, t_ph
    hs = []
    for t in range(n_steps):
        p = "cell_%d/" % t
        if duplicate_lstm_kernel and t > 0:   # some converters materialise one copy of a shared constant per consumer
            kern = const_weight(kern_name + "_%d" % t, weights["lstm_kernel"])
        xh = g.tensor(p + "concat", [1, H + C])
        g.op(CONCATENATION, [rows[t], h_prev], [xh])
        gates = g.tensor(p + "BiasAdd", [1, 4 * C])
        g.op(FULLY_CONNECTED, [xh, kern, kbias], [gates], fc_act=_ACT_NONE)
        gi, gj, gf, go = [g.tensor(p + "split:%d" % k, [1, C]) for k in range(4)]
        g.op(SPLIT, [t_axis, gates], [gi, gj, gf, go])
        si, tj, sf, so = [g.tensor(p + n, [1, C]) for n in ("Sigmoid", "Tanh", "Sigmoid_1", "Sigmoid_2")]
        g.op(LOGISTIC, [gi], [si])
        g.op(TANH, [gj], [tj])
        g.op(LOGISTIC, [gf], [sf])
        g.op(LOGISTIC, [go], [so])
        fc_, ij, c_new = [g.tensor(p + n, [1, C]) for n in ("mul", "mul_1", "add_1")]
        g.op(MUL, [sf, c_prev], [fc_])
        g.op(MUL, [si, tj], [ij])
        g.op(ADD, [fc_, ij], [c_new])
        tc, h_new = g.tensor(p + "Tanh_1", [1, C]), g.tensor(p + "mul_2", [1, C])
        g.op(TANH, [c_new], [tc])
        g.op(MUL, [so, tc], [h_new])
        hs.append(h_new)
        c_prev, h_prev = c_new, h_new
    t_newc, t_newh = g.tensor("new_state_c", [1, C]), g.tensor("new_state_h", [1, C])
    t_one = g.tensor("identity/shape", [2], I32, struct.pack("<ii", 1, C))
    g.op(RESHAPE, [c_prev, t_one], [t_newc])
    g.op(RESHAPE, [h_prev, t_one], [t_newh])
    t_stack = g.tensor("stack", [n_steps, C])
    g.op(PACK, hs, [t_stack])
    # ---- layers 5, 6 and the softmax (deepspeech_model.py:241-258, 353-357)
    w5 = const_weight("layer_5/weights/transpose", weights["w5"])
    b5 = const_f32("layer_5/bias", weights["b5"])
    t5 = g.tensor("layer_5/Relu", [n_steps, H])
    g.op(FULLY_CONNECTED, [t_stack, w5, b5], [t5], fc_act=_ACT_RELU)
    t5m = g.tensor("layer_5/Minimum", [n_steps, H])
    g.op(MINIMUM, [t5, t_clip], [t5m])
    w6 = const_weight("layer_6/weights/transpose", weights["w6"])
    b6 = const_f32("layer_6/bias", weights["b6"])
    t6 = g.tensor("raw_logits", [n_steps, K])
    g.op(FULLY_CONNECTED, [t5m, w6, b6], [t6], fc_act=_ACT_NONE)
    t_logits = g.tensor("logits", [n_steps, K])
    g.op(SOFTMAX, [t6], [t_logits])
    # ---- metadata (export.py:54-72)
    meta = []
    vals = (("metadata_version", graph_version), ("metadata_sample_rate", sample_rate), ("metadata_feature_win_len", win_len_ms),
            ("metadata_feature_win_step", win_step_ms), ("metadata_beam_width", beam_width))
    t_mshape = g.tensor("metadata/shape", [1], I32, struct.pack("<i", 1)) if metadata_via_op else None
    for name, v in vals:
        if metadata_via_op:
            src = g.tensor(name + "/value", [1], I32, struct.pack("<i", int(v)))
            dst = g.tensor(name, [1], I32)
            g.op(RESHAPE, [src, t_mshape], [dst])
        else:
            dst = g.tensor(name, [1], I32, struct.pack("<i", int(v)))
        meta.append(dst)
    alpha = _string_tensor_bytes([serialize_alphabet(labels)])
    if metadata_via_op:
        src = g.tensor("metadata_alphabet/value", [1], STRING, alpha)
        dst = g.tensor("metadata_alphabet", [1], STRING)
        g.op(RESHAPE, [src, t_mshape], [dst])
    else:
        dst = g.tensor("metadata_alphabet", [1], STRING, alpha)
    meta.append(dst)
    inputs = [t_input, t_pc, t_ph, t_samples]
    outputs = [t_logits, t_newc, t_newh, t_mfcc] + meta
    return g.finish(inputs, outputs, "Coqui STT acoustic model (written by stt_b200.tflite_export)")


def write_model(path, weights, **kw):
    with open(path, "wb") as f:
        f.write(model_bytes(weights, **kw))
