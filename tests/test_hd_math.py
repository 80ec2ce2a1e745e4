"""stt_b200/csrc/hd_math.h (the glibc logf/expf restatement the GPU decoder uses) vs the host libm -- the functions
the reference decoder calls (decoder_utils.h:46-53, ctc_beam_search_decoder.cpp:355).  EXHAUSTIVE over all 2^32 floats."""
import json
import os
import subprocess

from conftest import ROOT


def test_logf_expf_bit_exact_for_every_float(tmp_path):
    exe = str(tmp_path / "hd_math_check")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "hd_math_check.cc"), "-lpthread", "-lm"])
    out = subprocess.run([exe, "1"], capture_output=True, text=True)
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["checked"] == 2 ** 32
    assert res["logf_mismatch"] == 0 and res["expf_mismatch"] == 0, out.stderr


def test_log_sum_exp_matches_reference(ref_decoder, tmp_path):
    """log_sum_exp<float> of the compiled reference vs ours on random + edge pairs (via a tiny host harness)."""
    import ctypes
    import numpy as np
    src = tmp_path / "lse.cc"
    src.write_text('#include "%s/stt_b200/csrc/hd_math.h"\nextern "C" float our_lse(float x, float y) { return sttmath::log_sum_exp(x, y); }\n'
                   'extern "C" float our_logp(float p) { return sttmath::glibc_logf(p + 1.175494351e-38f); }\n' % ROOT)
    so = str(tmp_path / "liblse.so")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, str(src)])
    L = ctypes.CDLL(so)
    L.our_lse.restype = ctypes.c_float
    L.our_lse.argtypes = [ctypes.c_float, ctypes.c_float]
    L.our_logp.restype = ctypes.c_float
    L.our_logp.argtypes = [ctypes.c_float]
    R = ref_decoder.ref()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-90, 0, 20000), [-3.4028234663852886e38, 0.0, -1e-30, -87.0, -104.0]]).astype(np.float32)
    ys = np.concatenate([rng.uniform(-90, 0, 20000), [-5.0, -3.4028234663852886e38, 0.0, -87.5, -200.0]]).astype(np.float32)
    for x, y in zip(xs, ys):
        a, b = R.ref_log_sum_exp(float(x), float(y)), L.our_lse(float(x), float(y))
        assert np.float32(a).tobytes() == np.float32(b).tobytes(), (x, y, a, b)
    for p in np.concatenate([rng.uniform(0, 1, 5000), [0.0, 1.0, 1e-38, 1e-45, 0.999]]).astype(np.float32):
        a, b = R.ref_class_logprob(float(p)), L.our_logp(float(p))
        assert np.float32(a).tobytes() == np.float32(b).tobytes(), (p, a, b)
