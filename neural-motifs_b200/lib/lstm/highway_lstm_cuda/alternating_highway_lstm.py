"""Alternating highway LSTM over the persistent sm_100a kernels (csrc/lstm.cu).

Same surface as the reference's lib/lstm/highway_lstm_cuda/alternating_highway_lstm.py:
`AlternatingHighwayLSTM(input_size, hidden_size, num_layers=1, recurrent_dropout_probability=0)`
applied to a PackedSequence returns `(PackedSequence, None)` (:259-303); parameters are the flat
`weight` [sum_l 6H*In_l + 5H*H] and `bias` [5H*L] (:206-230), so state dicts interchange.
Layer l even runs forward in time, odd backward; highway gate r and linear carry come from the
input projection's 5th/6th chunk; recurrent (variational) dropout mask per (layer, batch, unit).
"""
import itertools

import torch
from torch.autograd import Function
from torch.nn import Parameter
from torch.nn.utils.rnn import PackedSequence, pad_packed_sequence, pack_padded_sequence

import motifs_cabi as _c


def block_orthogonal(tensor, split_sizes, gain=1.0):
    """Orthogonal init per block (alternating_highway_lstm.py:12-59): each
    split_sizes-shaped block of `tensor` gets its own orthogonal matrix."""
    sizes = list(tensor.size())
    if any(a % b != 0 for a, b in zip(sizes, split_sizes)):
        raise ValueError("tensor dimensions must be divisible by their respective split_sizes. "
                         "Found size: {} and split_sizes: {}".format(sizes, split_sizes))
    starts = [range(0, m, s) for m, s in zip(sizes, split_sizes)]
    with torch.no_grad():
        for origin in itertools.product(*starts):
            assert len(origin) == 2
            r, c = split_sizes
            side = max(r, c)
            block = tensor.new_empty(side, side)
            torch.nn.init.orthogonal_(block, gain=gain)
            tensor[origin[0]:origin[0] + r, origin[1]:origin[1] + c] = block[:r, :c]
    return tensor


class _AlternatingHighwayLSTMFunction(Function):
    @staticmethod
    def forward(ctx, inputs, weight, bias, dropout_mask, lengths_dev, hidden_size, num_layers, save_gates):
        _c.require_cuda(inputs, weight, bias, dropout_mask, lengths_dev)
        inputs = inputs.contiguous()
        T, B, In = inputs.shape
        H, L = hidden_size, num_layers
        dev = inputs.device
        state_acc = torch.zeros(L, T + 1, B, H, device=dev, dtype=torch.float32)
        memory_acc = torch.zeros(L, T + 1, B, H, device=dev, dtype=torch.float32)
        gates = torch.empty(L, T, B, 6 * H, device=dev, dtype=torch.float32) if save_gates else None
        scratch = None if save_gates else torch.empty(T, B, 6 * H, device=dev, dtype=torch.float32)
        lib = _c.load()
        with torch.cuda.device(dev):
            rc = lib.mb200_highway_lstm_forward(In, H, B, L, T, _c.ptr(inputs), _c.ptr(lengths_dev),
                                                _c.ptr(state_acc), _c.ptr(memory_acc), _c.ptr(weight), _c.ptr(bias),
                                                _c.ptr(dropout_mask), _c.ptr(gates), _c.ptr(scratch), _c.cur_stream())
        _c.check(rc, "mb200_highway_lstm_forward")
        ctx.dims = (T, B, In, H, L)
        ctx.have_gates = save_gates
        if save_gates:
            ctx.save_for_backward(inputs, lengths_dev, weight, bias, state_acc, memory_acc, dropout_mask, gates)
        # output = last layer, all slots but the initial state (:104-108)
        return state_acc[-1, 1:, :, :]

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.have_gates:
            raise _c.MotifsB200Error("AlternatingHighwayLSTM backward needs the gates saved in forward")
        inputs, lengths_dev, weight, bias, state_acc, memory_acc, dropout_mask, gates = ctx.saved_tensors
        T, B, In, H, L = ctx.dims
        dev = inputs.device
        grad_output = grad_output.contiguous()
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        grad_input = torch.empty_like(inputs)
        grad_state = torch.zeros_like(state_acc)
        grad_memory = torch.zeros_like(memory_acc)
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros_like(bias)
        h_out_grad = torch.empty(L, T, B, H, device=dev, dtype=torch.float32)
        dG = torch.empty(T, B, 6 * H, device=dev, dtype=torch.float32)
        lib = _c.load()
        with torch.cuda.device(dev):
            rc = lib.mb200_highway_lstm_backward(In, H, B, L, T, _c.ptr(grad_output), _c.ptr(lengths_dev),
                                                 _c.ptr(grad_state), _c.ptr(grad_memory), _c.ptr(inputs),
                                                 _c.ptr(state_acc), _c.ptr(memory_acc), _c.ptr(weight), _c.ptr(gates),
                                                 _c.ptr(dropout_mask), _c.ptr(h_out_grad), _c.ptr(grad_input),
                                                 _c.ptr(grad_weight), _c.ptr(grad_bias), 1 if need_w else 0,
                                                 _c.ptr(dG), _c.cur_stream())
        _c.check(rc, "mb200_highway_lstm_backward")
        return (grad_input, grad_weight if need_w else None, grad_bias if need_w else None,
                None, None, None, None, None)


class AlternatingHighwayLSTM(torch.nn.Module):
    def __init__(self, input_size, hidden_size, num_layers=1, recurrent_dropout_probability=0):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.recurrent_dropout_probability = recurrent_dropout_probability
        self.training = True
        total_weight, total_bias = 0, 0
        for layer in range(num_layers):
            layer_in = input_size if layer == 0 else hidden_size
            total_weight += 6 * hidden_size * layer_in + 5 * hidden_size * hidden_size
            total_bias += 5 * hidden_size
        self.weight = Parameter(torch.empty(total_weight, dtype=torch.float32))
        self.bias = Parameter(torch.empty(total_bias, dtype=torch.float32))
        self.reset_parameters()

    def reset_parameters(self):
        """Block-orthogonal weights, zero bias, forget-gate bias 1 (:233-257)."""
        H = self.hidden_size
        with torch.no_grad():
            self.bias.zero_()
            w, b = 0, 0
            for i in range(self.num_layers):
                insz = self.input_size if i == 0 else H
                init = block_orthogonal(self.weight.new_zeros(insz, 6 * H), [insz, H])
                self.weight[w:w + init.numel()].view_as(init).copy_(init)
                w += init.numel()
                init = block_orthogonal(self.weight.new_zeros(H, 5 * H), [H, H])
                self.weight[w:w + init.numel()].view_as(init).copy_(init)
                w += init.numel()
                self.bias[b + H:b + 2 * H].fill_(1)
                b += 5 * H

    def forward(self, inputs, initial_state=None, dropout_weights=None):
        """inputs: PackedSequence (time-major packing, lengths descending). `dropout_weights`
        [L,B,H] may be injected for reproducible parity runs; otherwise Bernoulli(1-p)/(1-p)
        in training and ones in eval (:283-288)."""
        if not isinstance(inputs, PackedSequence):
            raise ValueError('inputs must be PackedSequence but got %s' % (type(inputs)))
        padded, lengths = pad_packed_sequence(inputs, batch_first=False)
        T, B, _ = padded.shape
        dev = padded.device
        if dropout_weights is None:
            dropout_weights = torch.ones(self.num_layers, B, self.hidden_size, device=dev, dtype=torch.float32)
            if self.training and self.recurrent_dropout_probability > 0:
                keep = 1 - self.recurrent_dropout_probability
                dropout_weights.bernoulli_(keep).div_(keep)
        dropout_weights = dropout_weights.to(dev).contiguous()
        lengths_dev = lengths.to(device=dev, dtype=torch.int32)
        save_gates = torch.is_grad_enabled() and (padded.requires_grad or self.weight.requires_grad
                                                  or self.bias.requires_grad)
        output = _AlternatingHighwayLSTMFunction.apply(padded, self.weight, self.bias, dropout_weights, lengths_dev,
                                                       self.hidden_size, self.num_layers, save_gates)
        output = pack_padded_sequence(output, lengths, batch_first=False)
        return output, None
