"""tcgen05 GEMM kernel (stt_b200/csrc/gemm_tc.cuh) vs a plain torch fp32 reference of the same op, through the C ABI."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(M, N, K, epi, seed=0, clip=20.0):
    import torch
    from stt_b200 import api
    L = api.lib()
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(M, K, generator=g) * 0.5).half()
    w = (torch.randn(N, K, generator=g) / (K ** 0.5)).half()
    bias = torch.randn(N, generator=g) * 0.1
    if epi == 0:
        out = np.zeros((M, N), np.float16)
    elif epi == 1:
        out = np.zeros((M, N), np.float32)
    else:
        out = np.zeros((M, min(N, 29)), np.float32)
    ms = ctypes.c_float()
    an, wn, bn = a.numpy(), w.numpy(), bias.numpy().astype(np.float32)
    rc = L.STTX_DebugGemm(M, N, K, an.ctypes.data, wn.ctypes.data, bn.ctypes.data, epi, clip, out.ctypes.data,
                          ctypes.byref(ms))
    assert rc == 0
    ref = a.float() @ w.float().t() + bias
    return out, ref, ms.value


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 128), (1000, 512, 640), (4096, 2048, 2048), (77, 256, 2048)])
def test_gemm_bias_f32(M, N, K):
    out, ref, _ = _run(M, N, K, 1)
    # fp16 operands, fp32 accumulate: exact products, only summation-order differences
    np.testing.assert_allclose(out, ref.numpy(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (3000, 2048, 2048)])
def test_gemm_clipped_relu_f16(M, N, K):
    import torch
    out, ref, _ = _run(M, N, K, 0, clip=1.0)
    exp = torch.clamp(ref, 0.0, 1.0).half().float().numpy()
    np.testing.assert_allclose(out.astype(np.float32), exp, rtol=2e-3, atol=1e-3)
    assert (out > 0).any() and (out == 1.0).any()


def test_gemm_softmax():
    import torch
    out, ref, _ = _run(500, 32, 2048, 2)
    exp = torch.softmax(ref[:, :29], dim=1).numpy()
    np.testing.assert_allclose(out, exp, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out.sum(1), 1.0, rtol=1e-5)


def test_gemm_throughput_report():
    """Not a pass/fail perf gate: records that the big dense shape runs and prints achieved TFLOP/s."""
    M, N, K = 128000, 2048, 2048
    out, ref, ms = _run(M, N, K, 0)
    tflops = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    print("gemm 128000x2048x2048: %.3f ms, %.1f TFLOP/s" % (ms, tflops))
    assert ms > 0
