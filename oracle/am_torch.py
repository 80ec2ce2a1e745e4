"""TEST INFRASTRUCTURE ONLY.  Independent torch-CPU fp32 restatement of the reference acoustic model, used
(a) to cross-check oracle/stt_oracle.c and (b) as the acoustic half of the CPU baseline in bench.py
(TFLite cannot be built offline: SURVEY F6b).

Processes audio the way the reference does: batch 1, `n_steps`=16 timesteps per `infer`
(native_client/tflitemodelstate.cc:369-405), LSTM state carried between calls (native_client/stt.cc:311-334).
Layer definitions: training/coqui_stt_training/deepspeech_model.py:66-89 (dense + clipped ReLU), :144-168 and
tensorflow/python/keras/layers/legacy_rnn/rnn_cell_impl.py:1054-1079 (LSTMCell, gates i,j,f,o, forget_bias 0),
:241-258,353-357 (layer 5, layer 6, softmax).
"""
import numpy as np
import torch


class TorchAM(object):
    def __init__(self, weights, n_input=26, n_context=9, relu_clip=20.0, n_steps=16, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v, np.float32)) for k, v in weights.items()}
        self.n_input, self.n_context, self.clip, self.n_steps = n_input, n_context, relu_clip, n_steps
        self.H = self.w["b1"].numel()
        self.C = self.w["lstm_bias"].numel() // 4

    def _dense(self, x, w, b, relu=True):
        y = x @ self.w[w] + self.w[b]
        return torch.clamp(y, 0.0, self.clip) if relu else y

    def infer(self, x, c, h):
        """x [n, (2c+1)*n_input]; returns probs [n, K], c, h."""
        a = self._dense(self._dense(self._dense(x, "w1", "b1"), "w2", "b2"), "w3", "b3")
        C = self.C
        outs = []
        for t in range(a.shape[0]):
            g = torch.cat([a[t], h]) @ self.w["lstm_kernel"] + self.w["lstm_bias"]
            i, j, f, o = g[:C], g[C:2 * C], g[2 * C:3 * C], g[3 * C:]
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(j)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        hs = torch.stack(outs)
        logits = self._dense(self._dense(hs, "w5", "b5"), "w6", "b6", relu=False)
        return torch.softmax(logits, dim=1), c, h

    def forward_features(self, mfcc):
        """mfcc [F, n_input] (all frames incl. the flush frame) -> probs [F, K], 16 timesteps per call."""
        F = mfcc.shape[0]
        nc, ni = self.n_context, self.n_input
        pad = np.zeros((nc, ni), np.float32)
        seq = torch.from_numpy(np.concatenate([pad, np.asarray(mfcc, np.float32), pad]))
        windows = torch.stack([seq[t:t + 2 * nc + 1].reshape(-1) for t in range(F)])
        c = torch.zeros(self.C)
        h = torch.zeros(self.C)
        out = []
        with torch.no_grad():
            for t0 in range(0, F, self.n_steps):
                p, c, h = self.infer(windows[t0:t0 + self.n_steps], c, h)
                out.append(p)
        return torch.cat(out).numpy()
