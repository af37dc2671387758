"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement (differentiable torch fp32) of the
alternating highway LSTM: highway_lstm_forward_ongpu (highway_lstm_kernel.cu:377-496) with
elementWise_fp (:108-160), under the wrapper conventions of alternating_highway_lstm.py:259-303.
Backward (elementWise_bp :46-104, highway_lstm_backward_ongpu :162-375) is obtained by autograd
through this forward, which is an independent check of the hand-written backward kernels."""
import torch


def weight_views(weight, input_size, H, L):
    """Per-layer (W_i [In_l,6H], W_h [H,5H]) views of the flat weight
    (alternating_highway_lstm.py:212-221, highway_lstm_kernel.cu:435)."""
    views, off = [], 0
    for layer in range(L):
        insz = input_size if layer == 0 else H
        wi = weight[off:off + insz * 6 * H].view(insz, 6 * H)
        off += insz * 6 * H
        wh = weight[off:off + H * 5 * H].view(H, 5 * H)
        off += H * 5 * H
        views.append((wi, wh))
    return views


def highway_lstm_forward(x, lengths, weight, bias, dropout, H, L, return_all=False):
    """x [T,B,In] padded; lengths: python list, descending; dropout [L,B,H].
    Returns the last layer's outputs [T,B,H] (zero where t >= length), optionally with the
    per-layer state / memory accumulators [L,T+1,B,H] and gate activations [L,T,B,6H]."""
    T, B, In = x.shape
    views = weight_views(weight, In, H, L)
    zeros = x.new_zeros(B, H)
    all_h, all_c, all_g = [], [], []
    layer_in = [x[t] for t in range(T)]
    for layer in range(L):
        wi, wh = views[layer]
        b = bias[layer * 5 * H:(layer + 1) * 5 * H]
        hs = [zeros] * (T + 1)
        cs = [zeros] * (T + 1)
        gs = [x.new_zeros(B, 6 * H)] * T
        fwd = layer % 2 == 0
        order = range(T) if fwd else range(T - 1, -1, -1)
        for t in order:
            cov = sum(1 for l in lengths if l > t)   # :414-421 (lengths sorted descending)
            prev = t if fwd else (t + 2) % (T + 1)
            inp = layer_in[t][:cov]
            tmp_i = inp @ wi                           # [cov,6H]
            tmp_h = hs[prev][:cov] @ wh                # [cov,5H]
            g = tmp_i[:, :5 * H] + tmp_h + b
            i_g = torch.sigmoid(g[:, 0 * H:1 * H])
            f_g = torch.sigmoid(g[:, 1 * H:2 * H])
            a_g = torch.tanh(g[:, 2 * H:3 * H])
            o_g = torch.sigmoid(g[:, 3 * H:4 * H])
            r_g = torch.sigmoid(g[:, 4 * H:5 * H])
            lin = tmp_i[:, 5 * H:6 * H]
            c_new = f_g * cs[prev][:cov] + i_g * a_g
            h_new = o_g * torch.tanh(c_new)
            h_new = h_new * r_g + (1.0 - r_g) * lin
            h_new = h_new * dropout[layer][:cov]
            pad = x.new_zeros(B - cov, H)
            hs[t + 1] = torch.cat((h_new, pad), 0)
            cs[t + 1] = torch.cat((c_new, pad), 0)
            gs[t] = torch.cat((torch.cat((i_g, f_g, a_g, o_g, r_g, lin), 1), x.new_zeros(B - cov, 6 * H)), 0)
        layer_in = [hs[t + 1] for t in range(T)]
        all_h.append(torch.stack(hs, 0))
        all_c.append(torch.stack(cs, 0))
        all_g.append(torch.stack(gs, 0))
    out = torch.stack(layer_in, 0)
    if return_all:
        return out, torch.stack(all_h, 0), torch.stack(all_c, 0), torch.stack(all_g, 0)
    return out
