"""The product library's file parsers on damaged files: `.scorer` packages (stt_b200/csrc/scorer_image.cc: KenLM binary
header, section sizes, probing tables, OpenFst header) and model files (model_file.cc: STTB200W; tflite_reader.cc: the
hand-written flatbuffer reader, which works on the caller's buffer in place: STT_CreateModelFromBuffer).  Thousands of
truncated and byte-flipped copies of valid files go through the parsers compiled with AddressSanitizer +
UndefinedBehaviorSanitizer: a damaged file must be refused (or loaded) without reading outside the buffer, overflowing
or hanging.  The reference leaves these checks to KenLM / OpenFst / flatbuffers::Verifier; here they are ours."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

CSRC = os.path.join(ROOT, "stt_b200", "csrc")


@pytest.fixture(scope="module")
def fuzz_bin(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("fz") / "loader_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", exe,
           os.path.join(ROOT, "tests", "native", "loader_fuzz.cc"), os.path.join(CSRC, "scorer_image.cc"),
           os.path.join(CSRC, "model_file.cc"), os.path.join(CSRC, "tflite_reader.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("this g++ has no sanitizer runtime: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _run(exe, kind, path, utf8, n_trunc, n_flip, seed):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=1")
    r = subprocess.run([exe, kind, path, str(int(utf8)), str(n_trunc), str(n_flip), str(seed)], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, "parser fault on a damaged copy of %s:\n%s" % (path, r.stderr[-3000:])
    acc, rej = [int(x) for x in r.stdout.split()[1::2]]
    assert acc + rej == n_trunc + n_flip and rej >= n_trunc // 2   # truncations inside the headers are always refused
    return acc, rej


SCORERS = [("pruned_lm.scorer", 0), ("lm_variants/trie.scorer", 0), ("lm_variants/quant_trie.scorer", 0),
           ("lm_variants/array_trie.scorer", 0), ("lm_variants/quant_array_trie.scorer", 0),
           ("lm_variants/probing.scorer", 0), ("bytes/multilingual.bytes.scorer", 1)]


@pytest.mark.parametrize("name,utf8", SCORERS)
def test_damaged_scorer_files(fuzz_bin, name, utf8):
    _run(fuzz_bin, "scorer", os.path.join(GOLDEN, name), utf8, 800, 4000, seed=len(name))


@pytest.mark.parametrize("weight_type", ["float32", "float16", "int8"])
@pytest.mark.parametrize("metadata_via_op", [True, False])
def test_damaged_tflite_models(fuzz_bin, tmp_path, weight_type, metadata_via_op):
    from stt_b200 import synth, tflite_export
    w = synth.make_weights(n_hidden=16, seed=0)
    p = str(tmp_path / "m.tflite")
    open(p, "wb").write(tflite_export.model_bytes(w, weight_type=weight_type, metadata_via_op=metadata_via_op))
    _run(fuzz_bin, "model", p, 0, 1500, 6000, seed=3)


def test_damaged_native_model(fuzz_bin, tmp_path):
    from stt_b200 import synth
    p = str(tmp_path / "m.sttw")
    open(p, "wb").write(synth.model_bytes(synth.make_weights(n_hidden=16, seed=0)))
    _run(fuzz_bin, "model", p, 0, 1500, 6000, seed=4)


def test_damaged_wav_files_in_the_client(tmp_path):
    """The `stt` client's WAV reader (client.cc read_wav; the reference's NO_SOX path, client.cc:390-426): short `fmt `
    chunks, odd-sized or over-announced `data` chunks (0xffffffff from a pipe), truncated files."""
    from conftest import LDC93S1_WAV
    lib = os.path.join(ROOT, "stt_b200", "libstt_b200.so")
    if not os.path.exists(lib):
        pytest.skip("libstt_b200.so not built")
    exe = str(tmp_path / "wav_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I" + os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "native", "wav_fuzz.cc"),
           "-L" + os.path.join(ROOT, "stt_b200"), "-lstt_b200", "-Wl,-rpath," + os.path.join(ROOT, "stt_b200"),
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([exe, LDC93S1_WAV, str(tmp_path / "f.wav"), "6000"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.startswith("samples 46797 ")
