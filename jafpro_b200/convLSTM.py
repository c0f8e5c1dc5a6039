"""Drop-in for ``src/convLSTM.py``: ``ConvLSTMCell`` (:7-63) and ``ConvLSTM`` (:66-165) with the
whole cell step (cat, conv, split, gates, state update) in one fused kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class ConvLSTMCell(nn.Module):
    """Same constructor, parameters (``conv.weight`` [4Ch,Cin+Ch,kh,kw], ``conv.bias``) and forward
    signature as the reference cell, so its state_dict loads unchanged."""

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, bias):
        super().__init__()
        self.height, self.width = input_size
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.kernel_size = kernel_size
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.bias = bias
        # parameter holder only (same init as the reference); the convolution runs in the fused kernel
        self.conv = nn.Conv2d(in_channels=input_dim + hidden_dim, out_channels=4 * hidden_dim,
                              kernel_size=kernel_size, padding=self.padding, bias=bias)

    # 3x3 cells with Ch % 4 == 0 run on the tensor cores with split-bf16 arithmetic (abs. error ~3e-5; 3-6x faster at
    # the reference's cell sizes).  That is far closer to fp32 than what the reference itself executes on this GPU:
    # torch's cuDNN convolutions default to TF32 (~1e-3).  Set to False for the exact-fp32 CUDA-core kernel (~1e-6).
    tensor_cores = True

    def invalidate_packed_weights(self):
        """Drop the packed copy of ``conv.weight`` (call after an in-place update through ``.data``, which does not
        bump the parameter's version counter: EMA, ``w.data.copy_``, ``.data.normal_``)."""
        self._wpack_key = None

    def _packed(self):
        w = self.conv.weight
        # repacked whenever the parameter object, its version or its device changes, and always in training mode
        # (an optimizer step bumps the version; a `.data` write does not — see invalidate_packed_weights)
        key = (w.data_ptr(), w._version, w.device)
        if self.training or getattr(self, "_wpack_key", None) != key:
            self._wpack = ops.convlstm_gpack_weight(w.detach()[None].contiguous(), self.input_dim, self.hidden_dim)
            self._wpack_key = key
        return self._wpack

    def _tensor_core_ok(self, B, H, W):
        return (self.tensor_cores and tuple(self.kernel_size) == (3, 3)
                and ops.convlstm_grouped_supported(1, B, self.input_dim, self.hidden_dim, H, W))

    def forward(self, input, prev_state):
        h_prev, c_prev = prev_state
        # forward-only kernels: a training script must not silently lose its gradients (the reference trains these
        # cells, train/4.convLSTM_flowpro_interval.py:278)
        ops.forbid_grad(input, h_prev, c_prev, self.conv.weight, self.conv.bias, name="ConvLSTMCell input / parameter")
        B, _, H, W = input.shape
        if self._tensor_core_ok(B, H, W):
            b = None if self.conv.bias is None else self.conv.bias.detach()[None].contiguous()
            h, c = ops.convlstm_step_grouped(input.contiguous()[None], h_prev.contiguous()[None], c_prev.contiguous()[None],
                                             self._packed(), b, self.input_dim, self.hidden_dim)
            return h[0], c[0]
        # every other cell the reference constructor accepts (5x5 kernels, odd channel counts, cells too wide for
        # the shared-memory plan of the tensor-core kernel): the exact-fp32 CUDA-core kernel
        return ops.convlstm_step(input.contiguous(), h_prev.contiguous(), c_prev.contiguous(),
                                 self.conv.weight.contiguous(), self.conv.bias)

    def init_hidden(self, batch_size, cuda=True):
        dev = self.conv.weight.device  # the reference hard-codes .cuda() (:58-63)
        z = torch.zeros(batch_size, self.hidden_dim, self.height, self.width, device=dev)
        return z, z.clone()


class ConvLSTM(nn.Module):
    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, num_layers, batch_first=False, bias=True,
                 return_all_layers=False):
        super().__init__()
        self._check_kernel_size_consistency(kernel_size)
        kernel_size = self._extend_for_multilayer(kernel_size, num_layers)
        hidden_dim = self._extend_for_multilayer(hidden_dim, num_layers)
        if not len(kernel_size) == len(hidden_dim) == num_layers:
            raise ValueError('Inconsistent list length.')
        self.height, self.width = input_size
        self.input_dim, self.hidden_dim, self.kernel_size = input_dim, hidden_dim, kernel_size
        self.num_layers, self.batch_first, self.bias = num_layers, batch_first, bias
        self.return_all_layers = return_all_layers
        self.cell_list = nn.ModuleList([
            ConvLSTMCell((self.height, self.width), input_dim if i == 0 else hidden_dim[i - 1], hidden_dim[i],
                         kernel_size[i], bias) for i in range(num_layers)])

    def forward(self, input, hidden_state=None):
        """(t,b,c,h,w) or, with batch_first, (b,t,c,h,w) -> (layer_output of the last layer, [(h, c)] per layer),
        src/convLSTM.py:102-147."""
        if not self.batch_first:
            input = input.permute(1, 0, 2, 3, 4)
        if hidden_state is None:
            hidden_state = self.get_init_states(batch_size=input.size(0))
        layer_output_list, last_state_list = [], []
        cur = input
        for layer_idx in range(self.num_layers):
            h, c = hidden_state[layer_idx]
            outs = []
            for t in range(cur.size(1)):
                h, c = self.cell_list[layer_idx](input=cur[:, t], prev_state=[h, c])
                outs.append(h)
            cur = torch.stack(outs, dim=1)
            layer_output_list.append(cur)
            last_state_list.append((h, c))
        layer_output = layer_output_list[-1]
        if not self.batch_first:
            layer_output = layer_output.permute(1, 0, 2, 3, 4)
        return layer_output, last_state_list

    def get_init_states(self, batch_size, cuda=True):
        return [cell.init_hidden(batch_size, cuda) for cell in self.cell_list]

    @staticmethod
    def _check_kernel_size_consistency(kernel_size):
        if not (isinstance(kernel_size, tuple) or
                (isinstance(kernel_size, list) and all(isinstance(e, tuple) for e in kernel_size))):
            raise ValueError('`kernel_size` must be tuple or list of tuples')

    @staticmethod
    def _extend_for_multilayer(param, num_layers):
        if not isinstance(param, list):
            param = [param] * num_layers
        return param


class ConvLSTMCellTC(nn.Module):
    """The wide cell (BASELINE config 4: Cin = Ch = 256 @ 64x64) on tcgen05 tensor cores.
    Activations are channels-last bf16 ([B,H,W,C]); the cell state stays fp32.  Build it from a
    (reference or drop-in) ConvLSTMCell with ``from_cell``; weights are repacked once."""

    def __init__(self, input_dim, hidden_dim, weight, bias):
        super().__init__()
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.register_buffer('wpack', ops.convlstm_pack_weight(weight.detach().float().contiguous(), input_dim,
                                                               hidden_dim))
        self.register_buffer('bias', None if bias is None else bias.detach().float().contiguous())

    @classmethod
    def from_cell(cls, cell):
        return cls(cell.input_dim, cell.hidden_dim, cell.conv.weight, cell.conv.bias)

    def forward(self, x_nhwc, prev_state):
        h, c = prev_state
        ops.forbid_grad(x_nhwc, h, c, name="ConvLSTMCellTC input")
        return ops.convlstm_step_tc(x_nhwc, h, c, self.wpack, self.bias, self.input_dim, self.hidden_dim)


class ConvLSTMGrouped(nn.Module):
    """G independent reference-sized ConvLSTM layers advanced together, one launch per recurrent step: the 24
    part-specific ``ConvLSTM`` modules of one pyramid level of ``Downsampler_convLSTM`` inside
    ``Accumulate_LSTM_no_loss`` (src/networks.py:1304-1313,:1346-1355,:1641-1662).  fp32 in / out in the reference
    layout; tensor cores with split-bf16 arithmetic (abs. error ~3e-5, within the 1e-4 fp32 bound).
    Build it from the G single-layer (reference or drop-in) ``ConvLSTM`` modules with ``from_lstms``; weights are
    repacked once."""

    def __init__(self, input_dim, hidden_dim, weight, bias):
        super().__init__()
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.groups = weight.shape[0]
        self.register_buffer('wpack', ops.convlstm_gpack_weight(weight.detach().float().contiguous(), input_dim,
                                                                hidden_dim))
        self.register_buffer('bias', None if bias is None else bias.detach().float().contiguous())

    @classmethod
    def from_lstms(cls, lstms):
        cells = [m.cell_list[0] if hasattr(m, 'cell_list') else m for m in lstms]
        c0 = cells[0]
        if any(len(getattr(m, 'cell_list', [None])) != 1 for m in lstms):
            raise ValueError('ConvLSTMGrouped groups single-layer ConvLSTM modules')
        if any((c.input_dim, c.hidden_dim) != (c0.input_dim, c0.hidden_dim) or tuple(c.kernel_size) != (3, 3)
               for c in cells):
            raise ValueError('all grouped cells need the same channel counts and a 3x3 kernel')
        weight = torch.stack([c.conv.weight for c in cells])
        bias = None if c0.conv.bias is None else torch.stack([c.conv.bias for c in cells])
        return cls(c0.input_dim, c0.hidden_dim, weight, bias)

    def step(self, x, prev_state):
        """x [G,B,Cin,H,W], (h, c) [G,B,Ch,H,W] -> (h_next, c_next)."""
        h, c = prev_state
        return ops.convlstm_step_grouped(x.contiguous(), h.contiguous(), c.contiguous(), self.wpack, self.bias,
                                         self.input_dim, self.hidden_dim)

    def forward(self, input, hidden_state=None):
        """input [G,B,T,Cin,H,W] (the batch_first stacks x*_con of networks.py:1340-1344, one per part)
        -> (outputs [G,B,T,Ch,H,W], (h_last, c_last)); zero initial state like ConvLSTM.forward."""
        G, B, T, _, H, W = input.shape
        ops.forbid_grad(input, name="ConvLSTMGrouped input")
        if hidden_state is None:
            z = torch.zeros(G, B, self.hidden_dim, H, W, device=input.device)
            hidden_state = (z, z)
        h, c = hidden_state
        # the whole recurrence in one C call: T launches back to back, steps read / write the sequence tensors in place
        out, c_last = ops.convlstm_sequence_grouped(input.contiguous(), h.contiguous(), c.contiguous(), self.wpack, self.bias,
                                                    self.input_dim, self.hidden_dim)
        return out, (out[:, :, -1], c_last)
