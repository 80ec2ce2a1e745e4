"""The C-ABI library loads on a CPU-only box and exports every symbol include/stt_capi.h declares; the product
fails loudly (no CPU fallback) when no CUDA device is present; nothing under stt_b200/ touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT, _have_gpu


def _declared(dev_hooks=False):
    hdr = open(os.path.join(ROOT, "include", "stt_capi.h")).read()
    hooks = re.findall(r"#ifdef STT_B200_DEV_HOOKS(.*?)#endif", hdr, flags=re.S)
    if dev_hooks:
        hdr = "\n".join(hooks)
    else:
        for h in hooks:   # declared for the unit-test build only (libstt_b200_dev.so)
            hdr = hdr.replace(h, "")
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
    # the unit-test hooks live in the dev build only
    hooks = _declared(dev_hooks=True)
    assert hooks and not any(hasattr(L, n) for n in hooks), "test hooks must not be exported by the product library"
    dev = ctypes.CDLL(os.path.join(os.path.dirname(api.lib_path()), "libstt_b200_dev.so"))
    assert all(hasattr(dev, n) for n in hooks + names)


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


REF_HEADER_DIR = "/root/reference/native_client"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_HEADER_DIR, "coqui-stt.h")), reason="reference tree not mounted")
def test_client_built_against_the_reference_header_links_and_runs(tmp_path):
    """Drop-in at the source level: a C client that includes the REFERENCE's coqui-stt.h compiles, links against
    libstt_b200.so with every one of the 29 entry points resolved, and sees the reference's struct layouts."""
    from stt_b200 import api
    exe = str(tmp_path / "ref_client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", REF_HEADER_DIR,
                           os.path.join(ROOT, "tests", "native", "ref_header_client.c"), "-o", exe,
                           "-L", os.path.dirname(api.lib_path()), "-lstt_b200",
                           "-Wl,-rpath," + os.path.dirname(api.lib_path())])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    assert lines["TokenMetadata"] == "16 8 12" and lines["CandidateTranscript"] == "24 8 16"
    assert lines["AcousticModelEmissions"] == "32 8 16 24" and lines["Metadata"] == "24 8 16"
    assert lines["version"].startswith("1.4.0") and lines["err"] == "Invalid scorer file."
    assert lines["create_empty"] == "0x1000 1"          # STT_ERR_NO_MODEL, no model created
    assert lines["resolved"] == "29"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_HEADER_DIR, "coqui-stt.h")), reason="reference tree not mounted")
def test_our_header_declares_the_reference_prototypes_verbatim():
    """Every STT_* prototype in include/stt_capi.h must be token-for-token the one in the reference header."""
    def protos(path):
        txt = re.sub(r"/\*.*?\*/", " ", open(path).read(), flags=re.S)
        txt = re.sub(r"//[^\n]*", " ", txt)
        out = {}
        for m in re.finditer(r"STT_EXPORT\s+([^;{]*?\b(STT_[A-Za-z]+)\s*\([^;]*?\))\s*;", txt, flags=re.S):
            out[m.group(2)] = re.sub(r"\s+", " ", m.group(1)).replace("( ", "(").replace(" )", ")").strip()
        return out
    ref = protos(os.path.join(REF_HEADER_DIR, "coqui-stt.h"))
    ours = protos(os.path.join(ROOT, "include", "stt_capi.h"))
    assert len(ref) == 29
    for name, proto in ref.items():
        assert name in ours, name
        norm = lambda t: t.replace(" ", "").replace("(void)", "()")   # `f(void)` and `f()` declare the same C++ function
        assert norm(ours[name]) == norm(proto), (name, ours[name], proto)


REF_CLI = os.path.join(ROOT, "oracle", "_ref", "ref_stt_cli")


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/ref_stt_cli not built (make -C oracle ref)")
def test_reference_cli_runs_on_our_library(small_model):
    """The reference's OWN client.cc (compiled unmodified, -DNO_SOX) linked against libstt_b200.so: `--version` goes
    through our STT_Version, and without a GPU model creation fails the way client.cc reports it."""
    from stt_b200 import api
    out = subprocess.run([REF_CLI, "--version"], capture_output=True, text=True, timeout=60)
    assert out.stdout.strip() == "Coqui " + api.version()   # (client.cc exits 1 after printing the version)
    if not _have_gpu():
        path, _ = small_model
        r = subprocess.run([REF_CLI, "--model", path, "--audio", os.path.join(ROOT, "tests", "golden", "LDC93S1_pcms16le_1_16000.wav")],
                           capture_output=True, text=True, timeout=60)
        assert r.returncode != 0 and "Could not create model" in r.stderr


def test_library_runs_on_tcgen05_tma_not_on_legacy_mma():
    """The built product library's SASS (sm_100a): tcgen05.mma incl. cta_group::2 pairs, TMEM loads, TMA loads -- and no
    legacy mma.sync (HMMA), no cuBLAS / cuDNN dependency.  B200_PROFILING.md names the mnemonics."""
    import shutil
    from stt_b200 import api
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", api.lib_path()], capture_output=True, text=True).stdout
    n_mma = len(re.findall(r"\bUTC[A-Z]*MMA", sass))
    n_pair = len(re.findall(r"\bUTC[A-Z]*MMA\.2CTA", sass))
    assert n_mma >= 100 and n_pair >= 50, (n_mma, n_pair)
    assert len(re.findall(r"\bLDTM", sass)) >= 10 and len(re.findall(r"\bUTMALDG", sass)) >= 100
    assert not re.search(r"\bHMMA", sass)
    needed = subprocess.run(["readelf", "-d", api.lib_path()], capture_output=True, text=True).stdout
    assert "cublas" not in needed.lower() and "cudnn" not in needed.lower() and "torch" not in needed.lower()
