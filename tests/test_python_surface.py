"""The ctypes mirror offers every public method of the reference's Python binding (native_client/python/__init__.py
Model / Stream) and of the decoder package's Alphabet (native_client/ctcdecode/__init__.py), under the same names."""
import ast
import os

import pytest

from conftest import ROOT

REF = "/root/reference/native_client"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _surface(path):
    out = {}
    for n in ast.parse(open(path).read()).body:
        if isinstance(n, ast.ClassDef):
            out[n.name] = {m.name for m in n.body if isinstance(m, ast.FunctionDef) and not m.name.startswith("_")}
        elif isinstance(n, ast.FunctionDef) and not n.name.startswith("_"):
            out.setdefault("<functions>", set()).add(n.name)
    return out


def test_model_and_stream_methods():
    ref = _surface(os.path.join(REF, "python", "__init__.py"))
    ours = _surface(os.path.join(ROOT, "stt_b200", "api.py"))
    for cls in ("Model", "Stream"):
        assert ref[cls] <= ours[cls], (cls, sorted(ref[cls] - ours[cls]))
    assert "version" in ours["<functions>"]


def test_decoder_package_surface():
    ref = _surface(os.path.join(REF, "ctcdecode", "__init__.py"))
    ours = _surface(os.path.join(ROOT, "stt_b200", "ctcdecoder.py"))
    assert ref["Alphabet"] <= ours["Alphabet"], sorted(ref["Alphabet"] - ours["Alphabet"])
    assert "Scorer" in ours
    # the CTC beam-search entry points of the hot path; the wav2vec2 / flashlight decoders are other paths (out)
    assert {"ctc_beam_search_decoder", "ctc_beam_search_decoder_batch"} <= ours["<functions>"]
