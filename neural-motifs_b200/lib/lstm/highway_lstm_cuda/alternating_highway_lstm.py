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


TC_MIN_BATCH = 48          # below this the recurrence is latency-bound and the SIMT kernel (csrc/lstm.cu) is faster
TC_RECURRENCE = __import__("os").environ.get("MOTIFS_LSTM_TC", "1") == "1"


def use_tensor_core_recurrence(H, B):
    return TC_RECURRENCE and B >= TC_MIN_BATCH and bool(_c.load().mb200_highway_lstm_tc_supported(H, B))


def _recurrent_weight_slices(wh):
    """W_h [H,5H] fp32 -> the K-major bf16 pair the tensor-core recurrence keeps in shared memory: [H/16 * 80, H] with row
    (s*80 + g*16 + u) = column (g*H + 16 s + u) — slice s holds the five gates of hidden units 16 s .. 16 s + 15."""
    from lib import tc_ops
    H = wh.size(0)
    wt = wh.t().reshape(5, H // 16, 16, H).permute(1, 0, 2, 3).reshape(5 * H, H).contiguous()
    return tc_ops.split_rows(wt)


class _HighwayLayerFunction(Function):
    """One layer of the recurrence on the persistent kernel (csrc/lstm.cu). P [T,B,6H] is the hoisted
    input projection (a tcgen05 GEMM, lib/tc_ops.py); the kernel overwrites it in place with the six
    gate activations, which is what backward needs (elementWise_fp/bp, highway_lstm_kernel.cu:46-160)."""

    @staticmethod
    def forward(ctx, P, Wh, bias, dropout, lengths_dev, direction, save_gates, base=None, tc_tag=None):
        _c.require_cuda(P, Wh, bias, dropout, lengths_dev)
        P = P.contiguous()
        Wh = Wh.contiguous()
        bias = bias.contiguous()
        T, B, H6 = P.shape
        H = H6 // 6
        dev = P.device
        h = torch.zeros(T + 1, B, H, device=dev, dtype=torch.float32)
        c = torch.zeros(T + 1, B, H, device=dev, dtype=torch.float32)
        lib = _c.load()
        if use_tensor_core_recurrence(H, B):
            # large batches: the per-step [B,H] x [H,5H] product on tcgen05 (csrc/lstm_tc.cu); h / c / gates come out in
            # the same layout, so backward is unchanged
            from lib import tc_ops
            wt = tc_ops._cached_view(base, (tc_tag, "wh_tc"), Wh, _recurrent_weight_slices) if base is not None and tc_tag is not None \
                else _recurrent_weight_slices(Wh.detach())
            hb_hi = torch.zeros(T + 1, B, H, device=dev, dtype=torch.bfloat16)
            hb_lo = torch.zeros(T + 1, B, H, device=dev, dtype=torch.bfloat16)
            with torch.cuda.device(dev):
                rc = lib.mb200_highway_lstm_layer_forward_tc(H, B, T, direction, _c.ptr(P), _c.ptr(wt.hi), _c.ptr(wt.lo),
                                                             _c.ptr(bias), _c.ptr(dropout), _c.ptr(h), _c.ptr(c),
                                                             _c.ptr(hb_hi), _c.ptr(hb_lo),
                                                             _c.ptr(P) if save_gates else None, _c.ptr(lengths_dev),
                                                             _c.cur_stream())
            _c.check(rc, "mb200_highway_lstm_layer_forward_tc")
        else:
            with torch.cuda.device(dev):
                rc = lib.mb200_highway_lstm_layer_forward(H, B, T, direction, _c.ptr(P), _c.ptr(Wh), _c.ptr(bias),
                                                          _c.ptr(dropout), _c.ptr(h), _c.ptr(c),
                                                          _c.ptr(P) if save_gates else None, _c.ptr(lengths_dev),
                                                          _c.cur_stream())
            _c.check(rc, "mb200_highway_lstm_layer_forward")
        ctx.direction = direction
        ctx.have_gates = save_gates
        ctx.base = base            # the flat parameter Wh is a view of (direct gradient writes, tc_ops.direct_grad_target)
        if save_gates:
            ctx.save_for_backward(P, Wh, h, c, dropout, lengths_dev)
        return h[1:]

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.have_gates:
            raise _c.MotifsB200Error("highway LSTM backward needs the gates saved in forward")
        from lib import tc_ops
        gates, Wh, h, c, dropout, lengths_dev = ctx.saved_tensors
        T, B, H6 = gates.shape
        H = H6 // 6
        dev = gates.device
        grad_out = grad_out.contiguous()
        h_grad = torch.zeros_like(h)
        c_grad = torch.zeros_like(c)
        dG = torch.empty_like(gates)
        lib = _c.load()
        with torch.cuda.device(dev):
            rc = lib.mb200_highway_lstm_layer_backward(H, B, T, ctx.direction, _c.ptr(grad_out), _c.ptr(Wh), _c.ptr(h),
                                                       _c.ptr(c), _c.ptr(gates), _c.ptr(dropout), _c.ptr(h_grad),
                                                       _c.ptr(c_grad), _c.ptr(dG), _c.ptr(lengths_dev), _c.cur_stream())
        _c.check(rc, "mb200_highway_lstm_layer_backward")
        dWh = dbias = None
        dG2 = dG.view(T * B, 6 * H)
        if ctx.needs_input_grad[1]:
            # dW_h = Hprev^T dG[:, :5H] (:329-340): even layers read slot t, odd layers slot t+2 (t <= T-2)
            if ctx.direction == 0:
                hp, g5 = h[:T].reshape(T * B, H), dG2[:, :5 * H]
            else:
                hp, g5 = h[2:].reshape((T - 1) * B, H), dG2[:(T - 1) * B, :5 * H]
            if hp.size(0) > 0:
                tgt = tc_ops.direct_grad_target(ctx.base, Wh) if ctx.base is not None else None
                if tc_ops.GEMM_MN:      # Hprev^T dG straight from the row-major buffers (csrc/gemm_mn.cu)
                    dWh = tc_ops.gemm_mn(tc_ops.split_rows(hp), tc_ops.split_rows(g5), out=tgt)
                else:
                    dWh = tc_ops.gemm(tc_ops.split_transposed(hp), tc_ops.split_transposed(g5), out=tgt)
                if tgt is not None:
                    dWh = None
            else:
                dWh = torch.zeros_like(Wh)
        if ctx.needs_input_grad[2]:
            dbias = dG2[:, :5 * H].sum(0)
        return dG, dWh, dbias, None, None, None, None, None, None


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
        from lib.pytorch_misc import to_device_async
        lengths_dev = to_device_async(lengths, dev, torch.int32)       # (pageable H2D copies stall the host)
        save_gates = torch.is_grad_enabled() and (padded.requires_grad or self.weight.requires_grad
                                                  or self.bias.requires_grad)
        from lib import tc_ops
        H = self.hidden_size
        x = padded.contiguous()
        off = 0
        for layer in range(self.num_layers):
            insz = self.input_size if layer == 0 else H
            wi = self.weight[off:off + insz * 6 * H].view(insz, 6 * H)
            off += insz * 6 * H
            wh = self.weight[off:off + H * 5 * H].view(H, 5 * H)
            off += H * 5 * H
            b = self.bias[layer * 5 * H:(layer + 1) * 5 * H]
            # hoisted input projection for every timestep at once (the reference does one small
            # cublasSgemm per step, highway_lstm_kernel.cu:441-452)
            P = tc_ops.matmul_tc(x.view(T * B, insz), wi, self.weight, ("wi", layer)).view(T, B, 6 * H)
            x = _HighwayLayerFunction.apply(P, wh, b, dropout_weights[layer], lengths_dev, layer % 2, save_gates,
                                            self.weight, ("wh", layer))
        output = x
        output = pack_padded_sequence(output, lengths, batch_first=False)
        return output, None
