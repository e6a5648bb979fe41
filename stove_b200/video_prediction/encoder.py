"""LSTM recognition network (drop-in for model/video_prediction/encoder.py:7-57).

The GEMMs stay library GEMMs (SURVEY.md section 8f ranks the encoder first under "next") but run
as 3xTF32 on the tensor cores (ops.Linear3: fp32-level accuracy, ~3x faster than the SIMT-fp32
kernels cuBLAS picks when TF32 is off); the gate/state update is a fused kernel.  One structural
change is free: the reference feeds the *same* flattened frame for every one
of the `num_obj` LSTM steps (encoder.py:50), so `W_ih x + b` is computed once per frame
instead of once per step (1/3 of the input GEMM at O = 3) and the recurrence runs on the
small hidden-to-hidden GEMM only.  Parameter names are those of `nn.LSTM`
(`rnn.weight_ih_l0`, ...) so reference checkpoints load.
"""
import os

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
        if os.environ.get('STOVE_ENCODER_FP32') or not self.rnn.weight_ih_l0.is_cuda:
            return
        rnn = self.rnn
        ws = (rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0)
        # remembered with the parameter versions: a result that has gone stale (in-place update in between)
        # is dropped by `forward`
        self._prepared = (tuple(w._version for w in ws), ops.LstmEncoder.prepare(*ws))

    def forward(self, frames):
        """frames (N, c, w, h) -> (N, O, 8): means and raw stds of (sx, sy/sx, x, y)."""
        x = frames.flatten(start_dim=1)
        H = self.lstm_size
        rnn = self.rnn
        h = cell = None                              # zero initial state: first step needs no W_hh GEMM
        outs = []
        if os.environ.get('STOVE_ENCODER_FP32'):     # plain SIMT-fp32 library GEMMs (A/B reference)
            gates_x = torch.addmm(rnn.bias_ih_l0 + rnn.bias_hh_l0, x, rnn.weight_ih_l0.t())
            for _ in range(self.c.num_obj):
                gates_h = torch.mm(h, rnn.weight_hh_l0.t()) if h is not None else None
                h, cell = ops.LstmCell.apply(gates_x, gates_h, cell)
                outs.append(h)
        else:
            # one fused autograd node (ops.LstmEncoder): 3xTF32 GEMMs (every operand split once into a
            # TF32-exact part and a remainder; three tensor-core GEMMs reproduce the fp32 product to
            # ~4e-6 relative) and cell kernels that fold in the bias, the stacking and the splits
            prepared, self._prepared = getattr(self, '_prepared', None), None
            if prepared is not None:
                versions, prepared = prepared
                if versions != tuple(w._version for w in (rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0,
                                                          rnn.bias_hh_l0)):
                    prepared = None
            return ops.LstmEncoder.apply(x, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0,
                                         self.c.num_obj, self.fc1.weight, self.fc1.bias, self.fc2.weight,
                                         self.fc2.bias, prepared)
        zps = torch.stack(outs, 1)
        zps = torch.sigmoid(self.fc1(zps))
        return self.fc2(zps)
