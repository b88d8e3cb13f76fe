"""Recurrent classifiers of the v1 model zoo: a tanh RNN and an LSTM, unrolled over the time steps of the graph (every step is a
pair of GEMMs and fused elementwise gates); `model(x [B, T, F], labels)` -> (loss, logits) from the last hidden state.
(ref: hetu/v1/examples/cnn/models/{RNN,LSTM}.py)"""
from __future__ import annotations

from .. import ops
from ..nn import Linear, Module


class _Recurrent(Module):
    def _steps(self, x):
        b, t, f = x.shape
        return [ops.reshape(ops.slice(x, [0, i, 0], [b, 1, f]), [b, f]) for i in range(t)], b

    def _out(self, h, labels):
        logits = self.head(h)
        return logits if labels is None else (ops.softmax_cross_entropy_sparse(logits, labels, reduction="mean"), logits)


class RNN(_Recurrent):
    """h_t = tanh(W_x x_t + W_h h_{t-1} + b)"""

    def __init__(self, input_size=28, hidden_size=128, num_classes=10):
        super().__init__()
        self.hidden_size = hidden_size
        self.wx = Linear(input_size, hidden_size, name="rnn_wx")
        self.wh = Linear(hidden_size, hidden_size, bias=False, name="rnn_wh")
        self.head = Linear(hidden_size, num_classes, name="rnn_head")

    def forward(self, x, labels=None):
        xs, b = self._steps(x)
        h = None
        for xt in xs:
            pre = self.wx(xt) if h is None else self.wx(xt) + self.wh(h)
            h = ops.tanh(pre)
        return self._out(h, labels)


class LSTM(_Recurrent):
    """gates (i, f, g, o) from one fused projection of x_t and one of h_{t-1} (PyTorch gate order, so weights are interchangeable)"""

    def __init__(self, input_size=28, hidden_size=128, num_classes=10):
        super().__init__()
        self.hidden_size = hidden_size
        self.wx = Linear(input_size, 4 * hidden_size, name="lstm_wx")
        self.wh = Linear(hidden_size, 4 * hidden_size, name="lstm_wh")
        self.head = Linear(hidden_size, num_classes, name="lstm_head")

    def cell(self, xt, h, c, b):
        n = self.hidden_size
        z = self.wx(xt) + self.wh.bias if h is None else self.wx(xt) + self.wh(h)        # h_0 = 0: only the recurrent bias remains
        i, f, g, o = (ops.slice(z, [0, k * n], [b, n]) for k in range(4))
        i, f, g, o = ops.sigmoid(i), ops.sigmoid(f), ops.tanh(g), ops.sigmoid(o)
        c = i * g if c is None else f * c + i * g
        return o * ops.tanh(c), c

    def forward(self, x, labels=None, return_sequence=False):
        xs, b = self._steps(x)
        h = c = None
        hs = []
        for xt in xs:
            h, c = self.cell(xt, h, c, b)
            hs.append(h)
        if return_sequence:
            return hs, (h, c)
        return self._out(h, labels)
