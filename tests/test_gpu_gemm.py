"""tcgen05 GEMM kernels (stt_b200/csrc/gemm2_tc.cuh: CTA-pair 256 x 256 tiles, the dense layers' kernel; gemm_tc.cuh: one-CTA
128 x 256 tiles, layer 1 / softmax / fallback) vs a plain torch fp32 reference of the same op, through the unit-test build of
the library (libstt_b200_dev.so: the product library does not export these hooks)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(M, N, K, epi, seed=0, clip=20.0, single=False):
    import torch
    from stt_b200 import api
    L = api.dev_lib()
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(M, K, generator=g) * 0.5).half()
    w = (torch.randn(N, K, generator=g) / (K ** 0.5)).half()
    bias = torch.randn(N, generator=g) * 0.1
    if epi == 0:
        out = np.zeros((M, N), np.float16)
    elif epi == 1:
        out = np.zeros((M, N), np.float32)
    else:
        out = np.zeros((M, 200 if N == 256 else min(N, 29)), np.float32)
    ms = ctypes.c_float()
    an, wn, bn = a.numpy(), w.numpy(), bias.numpy().astype(np.float32)
    rc = L.STTX_DebugGemm(M, N, K, an.ctypes.data, wn.ctypes.data, bn.ctypes.data, epi + (16 if single else 0), clip, out.ctypes.data,
                          ctypes.byref(ms))
    assert rc == 0
    ref = a.float() @ w.float().t() + bias
    return out, ref, ms.value


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 128), (1000, 512, 640), (4096, 2048, 2048), (77, 256, 2048),
                                   (300, 256, 64), (513, 768, 192)])
def test_gemm_bias_f32(M, N, K, single):
    out, ref, _ = _run(M, N, K, 1, single=single)
    # fp16 operands, fp32 accumulate: exact products, only summation-order differences
    np.testing.assert_allclose(out, ref.numpy(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (3000, 2048, 2048)])
def test_gemm_clipped_relu_f16(M, N, K, single):
    import torch
    out, ref, _ = _run(M, N, K, 0, clip=1.0, single=single)
    exp = torch.clamp(ref, 0.0, 1.0).half().float().numpy()
    np.testing.assert_allclose(out.astype(np.float32), exp, rtol=2e-3, atol=1e-3)
    assert (out > 0).any() and (out == 1.0).any()


def test_gemm_softmax():
    import torch
    out, ref, _ = _run(500, 32, 2048, 2)
    exp = torch.softmax(ref[:, :29], dim=1).numpy()
    np.testing.assert_allclose(out, exp, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out.sum(1), 1.0, rtol=1e-5)


def test_gemm_wide_softmax():
    """Alphabets beyond 31 labels (bytes-output models: 256 classes): the three-pass softmax epilogue over 256 columns."""
    import torch
    out, ref, _ = _run(700, 256, 1024, 2)
    exp = torch.softmax(ref[:, :200], dim=1).numpy()
    np.testing.assert_allclose(out, exp, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out.sum(1), 1.0, rtol=1e-5)


def test_pair_and_single_kernels_agree_bit_for_bit():
    """Same fp16 operands, same accumulation order along K: the tile shape must not show in the result."""
    a, _, _ = _run(1000, 512, 2048, 1)
    b, _, _ = _run(1000, 512, 2048, 1, single=True)
    np.testing.assert_array_equal(a, b)


def test_gemm_throughput_report():
    """Not a pass/fail perf gate: records that the big dense shape runs and prints achieved TFLOP/s."""
    M, N, K = 128000, 2048, 2048
    out, ref, ms = _run(M, N, K, 0)
    tflops = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    print("gemm 128000x2048x2048: %.3f ms, %.1f TFLOP/s" % (ms, tflops))
    assert ms > 0
