"""The acoustic-model oracle in its three arithmetic modes (oracle/am_modes.py, oracle/stt_oracle.c):
fp32, f16-operand (the B200 path's precision mode) and TFLite hybrid int8 (the reference's default export arithmetic,
tensorflow/lite/kernels/fully_connected.cc:435-503 + internal/reference/portable_tensor_utils.cc:51-70,138-161).
The hybrid mode exists twice (C loops and torch) and both are checked against hand arithmetic."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def setup(oracle):
    from stt_b200 import synth
    w = synth.bench_weights(n_hidden=128)
    pcm = synth.make_pcm(12000, utt=3)
    _, mf = oracle.features_only(pcm)
    return w, pcm, mf


def test_fp32_mode_equals_c_restatement(oracle, setup):
    from oracle.am_modes import ModeAM
    w, pcm, mf = setup
    am = oracle.PortAM(w)
    p_stream, _ = am.stream(pcm)
    np.testing.assert_array_equal(am.forward_features(mf), p_stream)      # same C code, two drivers
    assert np.abs(ModeAM(w, "fp32").forward_features(mf) - p_stream).max() <= 1e-4
    # all f16 rounding sources switched off == fp32
    off = ModeAM(w, "f16", knobs={}).forward_features(mf)
    assert np.abs(off - ModeAM(w, "fp32").forward_features(mf)).max() <= 1e-5


def test_weight_quantisation_known_answers(oracle):
    import ctypes
    L = oracle.port()
    L.orc_quantize_weights.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    w = np.array([0.5, -1.0, 0.25, 0.0, 0.0039, 0.996, -0.0117], np.float32)
    q = np.zeros(w.size, np.int8)
    s = np.zeros(1, np.float32)
    L.orc_quantize_weights(w.ctypes.data, w.size, q.ctypes.data, s.ctypes.data)
    assert s[0] == np.float32(1.0) / np.float32(127.0)
    # 0.5*127 = 63.5 -> 64 (TfLiteRound: half away from zero); 0.25*127 = 31.75 -> 32; 0.0039*127 = 0.4953 -> 0;
    # 0.996*127 = 126.49 -> 126; -0.0117*127 = -1.486 -> -1
    assert q.tolist() == [64, -127, 32, 0, 0, 126, -1]
    from oracle.am_modes import quantize_weights_int8
    qt, st = quantize_weights_int8(w)
    assert qt.numpy().astype(np.int8).tolist() == q.tolist() and np.float32(st) == s[0]


def test_hybrid_modes_agree_and_match_hand_arithmetic(oracle, setup):
    from oracle.am_modes import ModeAM, _hybrid_matmul, quantize_weights_int8
    import torch
    w, pcm, mf = setup
    hc = oracle.PortAM(w).forward_features(mf, mode="hybrid8")
    ht = ModeAM(w, "hybrid8").forward_features(mf)
    assert np.abs(hc - ht).max() <= 1e-5
    # one FULLY_CONNECTED by hand, integer arithmetic in numpy
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 40)).astype(np.float32) * np.float32(3.0)
    x[1] = 0.0                                                       # all-zero row: output = bias
    W = rng.standard_normal((40, 7)).astype(np.float32)
    b = rng.standard_normal(7).astype(np.float32)
    wmax = np.abs(W).max()
    Wq = np.clip(np.sign(W) * np.floor(np.abs(W * np.float32(127.0 / wmax)) + 0.5), -127, 127).astype(np.int64)
    exp = np.tile(b, (3, 1))
    for r in (0, 2):
        m = np.abs(x[r]).max()
        xq = np.clip(np.sign(x[r]) * np.floor(np.abs(x[r] * np.float32(127.0) / m) + 0.5), -127, 127).astype(np.int64)
        dot = xq @ Wq
        exp[r] = b + dot.astype(np.float32) * (np.float32(m / np.float32(127.0)) * np.float32(wmax / 127.0))
    qw, sw = quantize_weights_int8(W)
    got = _hybrid_matmul(torch.from_numpy(x), qw, sw, torch.from_numpy(b)).numpy()
    np.testing.assert_allclose(got, exp, rtol=2e-6, atol=1e-6)


def test_distance_of_each_mode_from_fp32(setup):
    """The ordering every parity statement rests on: f16-operand arithmetic stays an order of magnitude closer to fp32
    than the reference's own default (hybrid int8) does -- on the calibrated (x200 output layer) benchmark model."""
    from oracle.am_modes import ModeAM
    w, pcm, mf = setup
    p32 = ModeAM(w, "fp32").forward_features(mf)
    d16 = np.abs(ModeAM(w, "f16").forward_features(mf) - p32).max()
    d8 = np.abs(ModeAM(w, "hybrid8").forward_features(mf) - p32).max()
    d8a = np.abs(ModeAM(w, "hybrid8", asymmetric=True).forward_features(mf) - p32).max()
    print("max|dp| vs fp32: f16 %.3e, hybrid8 %.3e, hybrid8 asymmetric %.3e" % (d16, d8, d8a))
    assert d16 < 0.1 * min(d8, d8a)
