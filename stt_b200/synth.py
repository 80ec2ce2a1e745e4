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
        b = l if isinstance(l, bytes) else l.encode("utf-8")   # bytes-output alphabets hold raw single bytes
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


def _np_forward_h5(w, mfcc, n_context=9, relu_clip=20.0):
    """numpy restatement of the layer stack up to layer 5 (deepspeech_model.py:204-251); used only to calibrate the
    synthetic output layer below."""
    F, ni = mfcc.shape
    pad = np.zeros((n_context, ni), np.float32)
    seq = np.concatenate([pad, mfcc.astype(np.float32), pad])
    x = np.stack([seq[t:t + 2 * n_context + 1].reshape(-1) for t in range(F)])
    relu = lambda v: np.clip(v, 0.0, relu_clip)
    a = relu(relu(relu(x @ w["w1"] + w["b1"]) @ w["w2"] + w["b2"]) @ w["w3"] + w["b3"])
    C = w["lstm_bias"].size // 4
    c = np.zeros(C, np.float32)
    h = np.zeros(C, np.float32)
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    hs = []
    for t in range(F):
        g = np.concatenate([a[t], h]) @ w["lstm_kernel"] + w["lstm_bias"]
        c = sig(g[2 * C:3 * C]) * c + sig(g[:C]) * np.tanh(g[C:2 * C])
        h = sig(g[3 * C:]) * np.tanh(c)
        hs.append(h)
    return relu(np.stack(hs).astype(np.float32) @ w["w5"] + w["b5"])


def np_mfcc(pcm, sample_rate=16000, win_len=512, win_step=320, n_dct=26, n_channels=40, lower=20.0):
    """numpy MFCC of every analysis window incl. the zero-padded flush window (same recipe as SURVEY appendix B);
    only used to calibrate synthetic weights -- the tests use the oracle restatement, not this."""
    pcm = np.asarray(pcm, np.int16)
    F = n_timesteps(pcm.size, win_len, win_step)
    x = np.zeros((F - 1) * win_step + win_len, np.float64)
    x[:pcm.size] = pcm.astype(np.float32) * np.float32(1.0 / 32768.0)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_len) / win_len)
    frames = np.stack([x[f * win_step:f * win_step + win_len] * hann for f in range(F)])
    fft = 1
    while fft < win_len:
        fft *= 2
    power = (np.abs(np.fft.rfft(frames, fft, axis=1)) ** 2).astype(np.float32).astype(np.float64)
    n_bins = fft // 2 + 1
    upper = sample_rate / 2.0
    mel = lambda f: 1127.0 * np.log1p(f / 700.0)
    centers = mel(lower) + (mel(upper) - mel(lower)) / (n_channels + 1) * (np.arange(n_channels + 1) + 1)
    hz = 0.5 * sample_rate / (n_bins - 1)
    start, end = int(1.5 + lower / hz), int(upper / hz)
    out = np.zeros((F, n_channels))
    spec = np.sqrt(power)
    ch = 0
    for i in range(start, min(end, n_bins - 1) + 1):
        m = mel(i * hz)
        while ch < n_channels and centers[ch] < m:
            ch += 1
        c = ch - 1
        wgt = (centers[c + 1] - m) / (centers[c + 1] - centers[c]) if c >= 0 else (centers[0] - m) / (centers[0] - mel(lower))
        if c >= 0:
            out[:, c] += spec[:, i] * wgt
        if c + 1 < n_channels:
            out[:, c + 1] += spec[:, i] * (1.0 - wgt)
    logmel = np.log(np.maximum(out, 1e-12))
    dct = np.sqrt(2.0 / n_channels) * np.cos(np.arange(n_dct)[:, None] * (np.pi / n_channels) * (np.arange(n_channels)[None, :] + 0.5))
    return (logmel @ dct.T).astype(np.float32)


def bench_weights(n_hidden=2048, seed=1234, scale=200.0, blank_bias=18.0):
    """The benchmark / smoke-test acoustic model: random init (make_weights) + CTC-like output layer calibrated on a
    fixed 4 s synthetic clip (make_ctc_like).  Deterministic in (n_hidden, seed)."""
    w = make_weights(n_hidden=n_hidden, seed=seed)
    return make_ctc_like(w, np_mfcc(make_pcm(64000, utt=777)), scale=scale, blank_bias=blank_bias)


def make_ctc_like(weights, mfcc_calib, scale=200.0, blank_bias=18.0):
    """Random-initialised networks emit a near-uniform, almost time-invariant softmax (max p ~ 0.04), which no CTC
    decoder ever sees in practice.  Keep every weight random but re-initialise the OUTPUT layer so the softmax is
    peaky, blank-favouring and follows the input: w6 <- scale * w6, b6 <- -mean_t(h5) @ w6 (centres the logits on a
    calibration clip) with +blank_bias on the CTC blank.  The reference CPU arm runs the same weights."""
    w = dict(weights)
    h5 = _np_forward_h5(w, mfcc_calib)
    w6 = (w["w6"] * scale).astype(np.float32)
    b6 = -(h5.mean(0) @ w6)
    b6[-1] += blank_bias
    w["w6"], w["b6"] = w6, b6.astype(np.float32)
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


_PHONES = None


def _phone_inventory(sample_rate):
    """40 fixed "phones": three sinusoid formants each (frequencies / amplitudes drawn once, seed 4242)."""
    global _PHONES
    if _PHONES is None:
        rng = np.random.default_rng(4242)
        _PHONES = [(rng.uniform(200, 3800, size=3), rng.uniform(0.3, 1.0, size=3)) for _ in range(40)]
    return _PHONES


def make_pcm(n_samples, utt=0, sample_rate=16000, base_seed=20260922):
    """Speech-like synthetic audio: a random sequence of "phones" (50-150 ms each, three sinusoid formants from a
    fixed 40-phone inventory, raised-cosine edges) with ~20 % low-noise pauses, plus a noise floor; int16, peak
    0.3 FS.  Every utterance shares the same long-term statistics, only the phone sequence differs."""
    rng = np.random.default_rng(base_seed + utt)
    if n_samples <= 0:
        return np.zeros(0, np.int16)
    phones = _phone_inventory(sample_rate)
    x = rng.standard_normal(n_samples) * 0.01
    pos = 0
    while pos < n_samples:
        dur = int(rng.uniform(0.05, 0.15) * sample_rate)
        seg = min(dur, n_samples - pos)
        if rng.random() >= 0.2:
            f, a = phones[int(rng.integers(len(phones)))]
            t = np.arange(seg) / sample_rate
            env = 0.5 - 0.5 * np.cos(2 * np.pi * np.minimum(np.arange(seg), seg - 1 - np.arange(seg)).clip(0, 160) / 320.0)
            sig = sum(ai * np.sin(2 * np.pi * fi * t + rng.uniform(0, 6.28)) for fi, ai in zip(f, a))
            x[pos:pos + seg] += sig * env
        pos += seg
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
