"""LSTM recognition network (drop-in for model/video_prediction/encoder.py:7-57).

Every contraction runs on the tensor cores as a hand-written tcgen05 kernel over 3xTF32 operand planes
(csrc/lstm_tc.cu: fp32-level accuracy), forward and backward; the LSTM cell is the epilogue of the gate
GEMM and the fc head is one fused kernel (csrc/enc_head.cu).  One structural change is free: the
reference feeds the *same* flattened frame for every one of the `num_obj` LSTM steps (encoder.py:50), so
`W_ih x + b` is computed once per frame instead of once per step (1/3 of the input GEMM at O = 3) and
the recurrence runs on the small hidden-to-hidden GEMM only.  Parameter names are those of `nn.LSTM`
(`rnn.weight_ih_l0`, ...) so reference checkpoints load.
"""
import torch
import torch.nn as nn

from .. import ops


class RnnStates(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.c = config
        self.z_size = 4
        self.lstm_size = 256
        img = self.c.channels * self.c.width * self.c.height
        self.rnn = nn.LSTM(img, self.lstm_size)
        self.fc1 = nn.Linear(self.lstm_size, 50)
        self.fc2 = nn.Linear(50, 2 * self.z_size)
        nn.init.xavier_uniform_(self.fc1.weight)
        nn.init.xavier_uniform_(self.fc2.weight)
        nn.init.constant_(self.fc1.bias, 0.1)
        nn.init.constant_(self.fc2.bias, 0.1)

    def prepare(self):
        """Issue the frame-independent part of the next forward pass (operand splits of the LSTM weights) now,
        on a side stream; `forward` picks the result up.  Called by Stove.forward before the frames are
        transformed, so that work is off the chain of the step."""
        if not self.rnn.weight_ih_l0.is_cuda:
            return
        rnn = self.rnn
        ws = (rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0)
        # remembered with the parameter versions: a result that has gone stale (in-place update in between)
        # is dropped by `forward`
        self._prepared = (tuple(w._version for w in ws), ops.LstmEncoder.prepare(*ws))

    def forward(self, frames, planes=None):
        """frames (N, c, w, h) -> (N, O, 8): means and raw stds of (sx, sy/sx, x, y).  `planes` (optional): the
        (hi, lo) TF32 operand planes (2, N, c*w*h) of the flattened frames when a producer already wrote them
        (ops.bw_transform)."""
        x = frames.flatten(start_dim=1)
        rnn = self.rnn
        # one fused autograd node (ops.LstmEncoder) for LSTM + head
        prepared, self._prepared = getattr(self, '_prepared', None), None
        if prepared is not None:
            versions, prepared = prepared
            if versions != tuple(w._version for w in (rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0,
                                                      rnn.bias_hh_l0)):
                prepared = None
        return ops.LstmEncoder.apply(x, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0,
                                     self.c.num_obj, self.fc1.weight, self.fc1.bias, self.fc2.weight,
                                     self.fc2.bias, prepared, planes)
