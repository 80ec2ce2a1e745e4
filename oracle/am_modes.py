"""TEST INFRASTRUCTURE ONLY.  The reference acoustic model restated in the three arithmetic modes that matter for
parity statements (SURVEY.md 8c "Acoustic model", 8d parity gate 2 "GPU vs restated AM in the same precision mode"):

  mode "fp32"     float32 everywhere -- what an unquantised export computes (deepspeech_model.py:66-89,144-168,204-263;
                  rnn_cell_impl.py:1054-1079).  Same arithmetic as oracle/am_torch.py and oracle/stt_oracle.c.
  mode "f16"      the operand rounding of the B200 path, nothing else: weights, MFCC features, every activation that
                  feeds a matmul and the recurrent h are rounded to IEEE fp16 (round-to-nearest-even); products are
                  accumulated in fp32; biases, the hoisted x*Wx term, the cell state c, gate math and softmax stay fp32.
                  This is "the same precision mode" as stt_b200/csrc/{gemm_tc,lstm2_tc}.cuh; what is left between this
                  mode and the GPU is accumulation order and the MUFU approximations of sigmoid/tanh.
  mode "hybrid8"  what a DEFAULT Coqui export computes (training/coqui_stt_training/util/config.py:616-622
                  `export_quantize` default true -> export.py:145-146 `converter.optimizations = [Optimize.DEFAULT]`):
                  TFLite "hybrid" FULLY_CONNECTED, tensorflow/lite/kernels/fully_connected.cc:435-503 (EvalHybridDense):
                  int8 per-tensor symmetric weights (scale = max|w|/127), activations quantised per batch row at run
                  time (tensor_utils::BatchQuantizeFloats -> PortableSymmetricQuantizeFloats,
                  internal/reference/portable_tensor_utils.cc:51-70: scale = max|x|/127, TfLiteRound = round half away
                  from zero, clamp to +-127), int32 dot product, `result += dotprod * (act_scale * weight_scale)` in
                  float (:138-161).  The LSTM's matmul sees concat([x_t, h_{t-1}]) as ONE row (rnn_cell_impl.py:1060),
                  so x (range 0..20) and h (range -1..1) share a scale.  `asymmetric=True` selects
                  PortableAsymmetricQuantizeFloats (:72-117), which newer converters request
                  (fully_connected.cc:473 `params->asymmetric_quantize_inputs`).

Every mode processes audio like the reference: batch 1, n_steps = 16 timesteps per `infer`
(native_client/tflitemodelstate.cc:369-405), LSTM state carried between calls (native_client/stt.cc:311-334).
`knobs` (mode "f16" only) switch the individual rounding sources on/off for the per-source error study
(tools/precision_study.py).
"""
import numpy as np
import torch

F16_KNOBS = ("weights", "features", "activations", "h_feedback", "h_output", "gate_fn")


def _r16(x):
    return x.to(torch.float16).to(torch.float32)


def _round_half_away(x):
    return torch.sign(x) * torch.floor(torch.abs(x) + 0.5)


def quantize_weights_int8(w):
    """Per-tensor symmetric int8 (tensorflow/lite/tools/optimize/quantize_weights + SymmetricQuantizeTensor):
    scale = max|w| / 127, q = round-half-away(w / scale) clamped to +-127.  w is [in, out] (TF layout)."""
    w = torch.as_tensor(w, dtype=torch.float32)
    rng = float(w.abs().max())
    if rng == 0.0:
        return torch.zeros_like(w), 1.0
    scale = np.float32(rng / 127.0)
    inv = np.float32(127.0 / rng)
    q = torch.clamp(_round_half_away(w * inv), -127, 127)
    return q, float(scale)


def _hybrid_matmul(x, wq, wscale, bias, asymmetric=False):
    """EvalHybridDense on rows of x [n, K] with int8 weights wq [K, N] (held as float32 integers)."""
    n, K = x.shape
    if asymmetric:
        rmin = torch.clamp(x.min(dim=1).values, max=0.0).double()
        rmax = torch.clamp(x.max(dim=1).values, min=0.0).double()
        scale = (rmax - rmin) / 255.0
        deg = rmin == rmax
        scale = torch.where(deg, torch.ones_like(scale), scale)
        zp_min = -128.0 - rmin / scale
        zp_max = 127.0 - rmax / scale
        err_min = 128.0 + torch.abs(rmin / scale)
        err_max = 127.0 + torch.abs(rmax / scale)
        zp = torch.where(err_min < err_max, zp_min, zp_max)
        zp = torch.where(zp <= -128.0, torch.full_like(zp, -128.0), torch.where(zp >= 127.0, torch.full_like(zp, 127.0),
                                                                                   _round_half_away(zp)))
        zp = torch.where(deg, torch.zeros_like(zp), zp)
        sf = scale.float()
        inv = (1.0 / sf)
        q = torch.clamp(_round_half_away(zp.float()[:, None] + x * inv[:, None]), -128, 127)
        q = torch.where(deg[:, None], torch.zeros_like(q), q)
        offs = zp
    else:
        rng = x.abs().max(dim=1).values
        zero = rng == 0
        safe = torch.where(zero, torch.ones_like(rng), rng)
        sf = torch.where(zero, torch.ones_like(rng), safe / 127.0)
        inv = 127.0 / safe
        q = torch.clamp(_round_half_away(x * inv[:, None]), -127, 127)
        q = torch.where(zero[:, None], torch.zeros_like(q), q)
        offs = None
    # exact int32 dot products: |q*w| <= 16256, so 1024-long partial sums stay below 2^24 and are exact in fp32
    acc = torch.zeros((n, wq.shape[1]), dtype=torch.float64)
    for k0 in range(0, K, 1024):
        acc += (q[:, k0:k0 + 1024] @ wq[k0:k0 + 1024]).double()
    if offs is not None:
        acc -= offs[:, None] * wq.sum(dim=0).double()[None, :]
    out = acc.float() * (sf * np.float32(wscale))[:, None]      # int32 -> float, times the float scaling factor
    return out + bias


class ModeAM(object):
    def __init__(self, weights, mode="fp32", n_input=26, n_context=9, relu_clip=20.0, n_steps=16, knobs=None,
                 asymmetric=False, threads=None):
        assert mode in ("fp32", "f16", "hybrid8")
        if threads:
            torch.set_num_threads(threads)
        self.mode, self.asym = mode, asymmetric
        self.n_input, self.n_context, self.clip, self.n_steps = n_input, n_context, relu_clip, n_steps
        w = {k: torch.from_numpy(np.ascontiguousarray(v, np.float32)) for k, v in weights.items()}
        self.H = w["b1"].numel()
        self.C = w["lstm_bias"].numel() // 4
        self.k = {n: (mode == "f16") for n in F16_KNOBS}
        if knobs is not None:
            assert mode == "f16"
            self.k = {n: bool(knobs.get(n, False)) for n in F16_KNOBS}
        mats = ("w1", "w2", "w3", "lstm_kernel", "w5", "w6")
        if mode == "f16" and self.k["weights"]:
            for m in mats:
                w[m] = _r16(w[m])
        self.ws = {}
        if mode == "hybrid8":
            for m in mats:
                w[m], self.ws[m] = quantize_weights_int8(w[m])
        self.w = w

    # one FULLY_CONNECTED
    def _fc(self, x, wn, bn, relu, round_in):
        if self.mode == "hybrid8":
            y = _hybrid_matmul(x, self.w[wn], self.ws[wn], self.w[bn], self.asym)
        else:
            if round_in:
                x = _r16(x)
            y = x @ self.w[wn] + self.w[bn]
        return torch.clamp(y, 0.0, self.clip) if relu else y

    def _sig(self, x, out_path):
        y = torch.sigmoid(x)
        return _r16(y) if (out_path and self.k["gate_fn"]) else y

    def _tanh(self, x, out_path):
        y = torch.tanh(x)
        return _r16(y) if (out_path and self.k["gate_fn"]) else y

    def infer(self, x, c, h):
        """x [n, (2c+1)*n_input] fp32 windows; returns (logits [n, K], probs, c, h)."""
        k = self.k
        a = self._fc(x, "w1", "b1", True, k["features"])
        a = self._fc(a, "w2", "b2", True, k["activations"])
        a = self._fc(a, "w3", "b3", True, k["activations"])
        C, H = self.C, self.H
        K = self.w["lstm_kernel"]
        outs = []
        if self.mode == "hybrid8":
            for t in range(a.shape[0]):
                g = _hybrid_matmul(torch.cat([a[t], h])[None, :], K, self.ws["lstm_kernel"], self.w["lstm_bias"],
                                   self.asym)[0]
                i, j, f, o = g[:C], g[C:2 * C], g[2 * C:3 * C], g[3 * C:]
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(j)
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
        else:
            ax = _r16(a) if k["activations"] else a
            xw = ax @ K[:H] + self.w["lstm_bias"]          # hoisted input half (fp32 result, as on the GPU)
            Wh = K[H:]
            for t in range(a.shape[0]):
                hh = _r16(h) if k["h_feedback"] else h
                g = xw[t] + hh @ Wh
                i, j, f, o = g[:C], g[C:2 * C], g[2 * C:3 * C], g[3 * C:]
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(j)
                h = self._sig(o, True) * self._tanh(c, True)
                outs.append(h)
        hs = torch.stack(outs)
        a5 = self._fc(hs, "w5", "b5", True, k["h_output"])
        logits = self._fc(a5, "w6", "b6", False, k["activations"])
        return logits, torch.softmax(logits, dim=1), c, h

    def forward_features(self, mfcc, return_logits=False):
        """mfcc [F, n_input] (all frames incl. the flush frame) -> probs [F, K], 16 timesteps per call."""
        F = mfcc.shape[0]
        nc, ni = self.n_context, self.n_input
        pad = np.zeros((nc, ni), np.float32)
        seq = torch.from_numpy(np.concatenate([pad, np.asarray(mfcc, np.float32), pad]))
        windows = torch.stack([seq[t:t + 2 * nc + 1].reshape(-1) for t in range(F)])
        c = torch.zeros(self.C)
        h = torch.zeros(self.C)
        out, lg = [], []
        with torch.no_grad():
            for t0 in range(0, F, self.n_steps):
                l, p, c, h = self.infer(windows[t0:t0 + self.n_steps], c, h)
                out.append(p)
                lg.append(l)
        probs = torch.cat(out).numpy()
        return (probs, torch.cat(lg).numpy()) if return_logits else probs
