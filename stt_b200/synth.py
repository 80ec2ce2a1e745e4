"""Synthetic inputs for tests / bench (SURVEY.md 8d "Synthetic inputs"): no model file, dataset or checkpoint ships
with the reference and there is no network, so weights are random-initialised like
training/coqui_stt_training/deepspeech_model.py:66-75 (VarianceScaling fan_avg uniform, zero biases), PCM is
band-limited noise + amplitude-modulated sinusoids, and decoder-only runs use CTC-like emissions built from the
scorer vocabulary.  Also the writer of the native `.sttw` model container (stt_b200/csrc/model_file.h)."""
import struct

import numpy as np

ENGLISH_LABELS = [" "] + [chr(ord("a") + i) for i in range(26)] + ["'"]  # data/alphabet.txt


def serialize_alphabet(labels):
    """Alphabet::Serialize format (native_client/alphabet.cc:101-125)."""
    out = struct.pack("<H", len(labels))
    for i, l in enumerate(labels):
        b = l.encode("utf-8")
        out += struct.pack("<HH", i, len(b)) + b
    return out


def make_weights(n_hidden=2048, n_input=26, n_context=9, n_classes=29, seed=1234, n_cell=None, lstm_scale=1.0):
    """Random weights in TF layout ([in, out]); dict of float32 arrays."""
    rng = np.random.default_rng(seed)
    n_cell = n_cell or n_hidden
    H, C, K = n_hidden, n_cell, n_classes
    in1 = (2 * n_context + 1) * n_input

    def vs(n_in, n_out):  # VarianceScaling(scale=1, mode="fan_avg", distribution="uniform")
        limit = np.sqrt(3.0 * 2.0 / (n_in + n_out))
        return rng.uniform(-limit, limit, size=(n_in, n_out)).astype(np.float32)

    w = {
        "w1": vs(in1, H), "b1": np.zeros(H, np.float32),
        "w2": vs(H, H), "b2": np.zeros(H, np.float32),
        "w3": vs(H, H), "b3": np.zeros(H, np.float32),
        "lstm_kernel": (rng.uniform(-1, 1, size=(H + C, 4 * C)) * lstm_scale / np.sqrt(H + C)).astype(np.float32),
        "lstm_bias": np.zeros(4 * C, np.float32),
        "w5": vs(C, H), "b5": np.zeros(H, np.float32),
        "w6": vs(H, K), "b6": np.zeros(K, np.float32),
    }
    # small non-zero biases so the bias path is exercised
    for k in ("b1", "b2", "b3", "lstm_bias", "b5", "b6"):
        w[k] = (rng.standard_normal(w[k].shape) * 0.05).astype(np.float32)
    return w


def model_bytes(weights, labels=ENGLISH_LABELS, sample_rate=16000, win_len=512, win_step=320, n_input=26, n_context=9,
                n_steps=16, beam_width=500, relu_clip=20.0):
    H = weights["b1"].shape[0]
    C = weights["lstm_bias"].shape[0] // 4
    K = weights["b6"].shape[0]
    alpha = serialize_alphabet(labels)
    out = b"STTB200W" + struct.pack("<I", 1)
    out += struct.pack("<10If", sample_rate, win_len, win_step, n_input, n_context, H, C, K, n_steps, beam_width,
                       relu_clip)
    out += struct.pack("<I", len(alpha)) + alpha
    for k in ("w1", "b1", "w2", "b2", "w3", "b3", "lstm_kernel", "lstm_bias", "w5", "b5", "w6", "b6"):
        out += np.ascontiguousarray(weights[k], dtype="<f4").tobytes()
    return out


def write_model(path, weights, **kw):
    with open(path, "wb") as f:
        f.write(model_bytes(weights, **kw))


def make_pcm(n_samples, utt=0, sample_rate=16000, base_seed=20260922):
    """Band-limited noise + 3-5 random sinusoid "formants" amplitude-modulated at 4 Hz, int16, peak 0.3 FS."""
    rng = np.random.default_rng(base_seed + utt)
    if n_samples <= 0:
        return np.zeros(0, np.int16)
    t = np.arange(n_samples) / sample_rate
    x = rng.standard_normal(n_samples)
    # crude band-limit: moving average of 4 samples
    x = np.convolve(x, np.ones(4) / 4.0, mode="same") * 0.2
    for _ in range(rng.integers(3, 6)):
        f = rng.uniform(200, 3500)
        x += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28)) * \
            (0.5 + 0.5 * np.sin(2 * np.pi * 4 * t + rng.uniform(0, 6.28)))
    x = x / (np.abs(x).max() + 1e-9) * 0.3 * 32767
    return x.astype(np.int16)


def make_ctc_probs(words, T, n_classes=29, labels=ENGLISH_LABELS, utt=0, seed=7, noise=0.02, peak=(0.5, 0.95)):
    """CTC-like emissions: text drawn from `words`, one emitting frame (p in `peak`) plus 1-3 blank frames
    (p ~ 0.9) per character, Dirichlet-ish noise elsewhere.  Returns float32 [T, n_classes] (rows sum to 1)."""
    rng = np.random.default_rng(seed + utt)
    lab = {l: i for i, l in enumerate(labels)}
    blank = n_classes - 1
    seq = []  # (class, is_emit)
    while len(seq) < T:
        w = words[rng.integers(len(words))]
        for ch in w + " ":
            if ch not in lab:
                continue
            seq.append((lab[ch], True))
            for _ in range(rng.integers(1, 4)):
                seq.append((blank, False))
    seq = seq[:T]
    probs = rng.gamma(0.3, 1.0, size=(T, n_classes)) * noise
    for t, (c, emit) in enumerate(seq):
        p = rng.uniform(*peak) if emit else rng.uniform(0.85, 0.95)
        probs[t] *= (1 - p) / probs[t].sum()
        probs[t, c] += p
        # some confusable mass on a neighbouring letter for emitting frames
        if emit and rng.random() < 0.5:
            alt = int(rng.integers(0, n_classes - 1))
            shift = probs[t, c] * rng.uniform(0.05, 0.4)
            probs[t, c] -= shift
            probs[t, alt] += shift
    probs = probs / probs.sum(axis=1, keepdims=True)
    return probs.astype(np.float32)


def n_timesteps(n_samples, win_len=512, win_step=320):
    """Frames (= timesteps) the streaming runtime emits for n_samples (stt.cc:105-128,236-254)."""
    return (max(0, (n_samples - win_len) // win_step + 1) if n_samples >= win_len else 0) + 1
