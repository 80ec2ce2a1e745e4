"""The C-ABI library loads on a CPU-only box and exports every symbol include/stt_capi.h declares; the product
fails loudly (no CPU fallback) when no CUDA device is present; nothing under stt_b200/ touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT, _have_gpu


def _declared():
    hdr = open(os.path.join(ROOT, "include", "stt_capi.h")).read()
    return sorted(set(re.findall(r"STT_EXPORT[^;(]*?\b(STTX?_[A-Za-z0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from stt_b200 import api
    if not os.path.exists(api.lib_path()):
        subprocess.check_call(["make", "-C", ROOT, "-j4"])
    L = ctypes.CDLL(api.lib_path())
    names = _declared()
    assert len([n for n in names if n.startswith("STT_")]) == 29  # coqui-stt.h exports 29 STT_* functions
    for n in names:
        assert hasattr(L, n), n
    assert sorted(api.DECLARED_SYMBOLS) == names
    exported = subprocess.run(["nm", "-D", "--defined-only", api.lib_path()], capture_output=True, text=True).stdout
    other = [l.split()[-1] for l in exported.splitlines() if " T " in l and not l.split()[-1].startswith(("STT_", "STTX_"))]
    assert other == [], "only the C ABI may be exported: %s" % other[:5]


def test_struct_layout_matches_reference_abi():
    from stt_b200 import api
    assert ctypes.sizeof(api._TokenMetadata) == 16 and api._TokenMetadata.timestep.offset == 8
    assert ctypes.sizeof(api._CandidateTranscript) == 24 and api._CandidateTranscript.confidence.offset == 16
    assert ctypes.sizeof(api._Metadata) == 24 and api._Metadata.emissions.offset == 16
    assert ctypes.sizeof(api._Emissions) == 32


def test_error_messages_without_gpu():
    from stt_b200 import api
    assert api._err(0) == "No error."
    assert api._err(0x2002) == "Invalid scorer file."
    assert api._err(0x3010) == "Could not erase hot-word."
    assert api._err(777).startswith("Unknown error")
    assert api.version().startswith("1.4.0")


@pytest.mark.skipif(_have_gpu(), reason="CPU-only behaviour")
def test_fails_loudly_without_cuda(small_model):
    from stt_b200 import Model, STTError
    path, _ = small_model
    with pytest.raises(STTError):
        Model(path)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "stt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle restatement", "").replace("use the oracle", "") or \
                    f in ("synth.py",), "%s mentions oracle/" % f
    assert "import oracle" not in open(os.path.join(pkg, "synth.py")).read()
    assert "from oracle" not in open(os.path.join(pkg, "synth.py")).read()
