"""STT_CreateModel / STT_CreateModelFromBuffer on the reference's container (a TFLite flatbuffer) -- SURVEY 8b model-file
contract, 8f rank 1: the model loaded from a .tflite computes exactly what the same weights do from the native .sttw."""
import numpy as np
import pytest

from conftest import SCORER

pytestmark = pytest.mark.gpu


def _run(model, pcm):
    b = model.createBatch(1, pcm.size)
    b.upload([pcm])
    b.forward()
    b.decode(1)
    b.fetch()
    return b.probs(0), b.transcripts()[0]


def test_tflite_float_model_equals_native_container(tmp_path, small_model):
    from stt_b200 import Model, synth, tflite_export
    path, w = small_model
    pcm = synth.make_pcm(24000, utt=12)
    native = Model(path)
    native.enableExternalScorer(SCORER)
    p0, t0 = _run(native, pcm)
    tfl = str(tmp_path / "model.tflite")
    tflite_export.write_model(tfl, w)
    from_file = Model(tfl)
    from_file.enableExternalScorer(SCORER)
    p1, t1 = _run(from_file, pcm)
    np.testing.assert_array_equal(p0, p1)
    assert t0 == t1
    assert from_file.beamWidth() == 500 and from_file.sampleRate() == 16000
    # from a caller-owned buffer (tflitemodelstate.cc:169-174); nothing of the buffer is needed after the call returns
    data = bytearray(open(tfl, "rb").read())
    from_buf = Model(bytes(data))
    for i in range(0, len(data), 4096):
        data[i] = 0xFF
    from_buf.enableExternalScorer(SCORER)
    p2, t2 = _run(from_buf, pcm)
    np.testing.assert_array_equal(p0, p2)
    assert t0 == t2
    assert from_buf.stt(pcm) == native.stt(pcm)       # the reference's own entry point (stream API) on the .tflite model


@pytest.mark.parametrize("weight_type", ["int8", "float16"])
def test_tflite_quantised_exports(tmp_path, oracle, small_model, weight_type):
    """The default export (hybrid int8) and the float16 export: the device computes with the dequantised weights, i.e.
    exactly what a .sttw holding those dequantised weights gives; the distance to the reference's hybrid arithmetic
    (activations quantised to int8 per row as well) is reported."""
    from stt_b200 import Model, api, synth, tflite_export
    _, w = small_model
    pcm = synth.make_pcm(24000, utt=13)
    data = tflite_export.model_bytes(w, weight_type=weight_type)
    m = Model(data)
    p_q, _ = _run(m, pcm)
    _, tens = api.inspect_model(data)
    wd = {k: tens[k].reshape(np.asarray(w[k]).shape) for k in tens}
    native = str(tmp_path / "dequantised.sttw")
    synth.write_model(native, wd)
    p_n, _ = _run(Model(native), pcm)
    np.testing.assert_array_equal(p_q, p_n)
    am = oracle.PortAM(wd)
    _, mf = oracle.features_only(pcm)
    d32 = float(np.abs(p_q - am.forward_features(mf)).max())
    print("%s export: max|dp| vs fp32 arithmetic on the dequantised weights %.3e" % (weight_type, d32))
    assert d32 <= 2e-3
    if weight_type == "int8":
        dh = float(np.abs(p_q - oracle.PortAM(w).forward_features(mf, mode="hybrid8")).max())
        d0 = float(np.abs(oracle.PortAM(w).forward_features(mf) - oracle.PortAM(w).forward_features(mf, mode="hybrid8")).max())
        print("int8 export: max|dp| device vs the reference's hybrid arithmetic %.3e (hybrid vs fp32 of the original weights: %.3e)"
              % (dh, d0))


def test_tflite_rejections(tmp_path, small_model):
    from stt_b200 import Model, tflite_export
    from stt_b200.api import STTError
    _, w = small_model
    with pytest.raises(STTError):
        Model(tflite_export.model_bytes(w, graph_version=5))
    with pytest.raises(STTError):
        Model(tflite_export.model_bytes(w, labels=["a", " "]))
