"""The `stt` CLI (stt_b200/csrc/client.cc) over the public C ABI: same flags / outputs as native_client/client.cc,
checked like ci_scripts/asserts.sh does (plain, --extended, --json, --stream, --init_from_bytes must agree)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SCORER

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "stt_b200", "stt")


def _write_wav(path, pcm, rate=16000):
    data = np.asarray(pcm, np.int16).tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def _run(*args):
    r = subprocess.run([CLI] + list(args), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout, r.stderr


def test_cli_modes_agree(small_model, tmp_path):
    from stt_b200 import Model, synth
    path, _ = small_model
    pcm = synth.make_pcm(40000, utt=12)
    wav = str(tmp_path / "a.wav")
    _write_wav(wav, pcm)
    m = Model(path)
    m.setBeamWidth(64)
    m.enableExternalScorer(SCORER)
    expect = m.stt(pcm)
    base = ["--model", path, "--scorer", SCORER, "--audio", wav, "--beam_width", "64"]
    out, err = _run(*base)
    assert out.strip().split("\n")[-1] == expect
    assert " Coqui STT:" in err and "TensorFlow:" in err           # asserts.sh:284-321
    assert _run(*base, "--extended")[0].strip() == expect
    assert _run(*base, "--init_from_bytes")[0].strip() == expect
    assert _run(*base, "--stream", "1280")[0].strip().split("\n")[-1] == expect  # asserts.sh:591-604
    assert _run(*base, "--extended_stream", "1280")[0].strip().split("\n")[-1] == expect
    js = json.loads(_run(*base, "--json", "--candidate_transcripts", "3")[0])
    assert " ".join(w["word"] for w in js["words"]).strip() == expect.strip()
    assert len(js.get("alternatives", [])) <= 2 and "confidence" in js["metadata"]
    t_out = _run(*base, "-t")[0]
    assert "cpu_time_overall=" in t_out
    hot = _run(*base, "--hot_words", "the:2.0,and:-1.5")[0]
    assert isinstance(hot, str)


def test_cli_errors(small_model, tmp_path):
    path, _ = small_model
    r = subprocess.run([CLI, "--model", "/nonexistent.sttw", "--audio", "x.wav"], capture_output=True, text=True)
    assert r.returncode != 0 and "Could not create model" in r.stderr
    r = subprocess.run([CLI, "--model", path, "--audio", "x.wav", "--stream", "100"], capture_output=True, text=True)
    assert "multiples of 160" in r.stdout


def test_cli_on_the_reference_smoke_test_wav(small_model):
    """asserts.sh runs `stt --model ... --audio LDC93S1_pcms16le_1_16000.wav`: the CLI's WAV reader on the reference's
    own file (a 44-byte canonical header here, but parsed chunk by chunk) must feed exactly the samples the library sees."""
    from conftest import LDC93S1_WAV
    from stt_b200 import Model
    import wave
    path, _ = small_model
    with wave.open(LDC93S1_WAV, "rb") as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    m = Model(path)
    m.setBeamWidth(64)
    m.enableExternalScorer(SCORER)
    base = ["--model", path, "--scorer", SCORER, "--audio", LDC93S1_WAV, "--beam_width", "64"]
    assert _run(*base)[0].strip().split("\n")[-1] == m.stt(pcm)


REF_CLI = os.path.join(ROOT, "oracle", "_ref", "ref_stt_cli")


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/ref_stt_cli not built (make -C oracle ref)")
def test_the_reference_cli_binary_transcribes_through_our_library(small_model):
    """Drop-in, literally: native_client/client.cc compiled UNMODIFIED (-DNO_SOX) and linked against libstt_b200.so
    transcribes the reference's smoke-test WAV; its output equals the library's own answer."""
    from conftest import LDC93S1_WAV
    from stt_b200 import Model
    import wave
    path, _ = small_model
    with wave.open(LDC93S1_WAV, "rb") as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    m = Model(path)
    m.setBeamWidth(64)
    m.enableExternalScorer(SCORER)
    r = subprocess.run([REF_CLI, "--model", path, "--scorer", SCORER, "--audio", LDC93S1_WAV, "--beam_width", "64"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().split("\n")[-1] == m.stt(pcm)
