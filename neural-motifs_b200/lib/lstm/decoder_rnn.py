"""Object-label decoder — same surface as the reference's lib/lstm/decoder_rnn.py:40-251:
`DecoderRNN(classes, embed_dim, inputs_dim, hidden_dim, recurrent_dropout_probability)` applied to
a PackedSequence with `labels=` / `boxes_for_nms=` returns `(dists [N,num_classes], commitments [N])`.

The reference unrolls a highway-LSTM cell in a Python loop, three small GEMMs per timestep (:186-227).
Here, when the fed-back label is known up front (training with no background label — always the
case with GT boxes, lib/object_detector.py:200-224) the whole thing is teacher forced: ONE
input-projection GEMM over every timestep, ONE persistent recurrent kernel (the highway-LSTM
layer kernel of csrc/lstm.cu, direction forward) and ONE output GEMM. Greedy decoding (eval, or
background labels in training) keeps the step loop, on the tensor-core GEMM."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import PackedSequence

from lib.fpn.box_utils import nms_overlaps
from lib.word_vectors import obj_edge_vectors
from lib.lstm.highway_lstm_cuda.alternating_highway_lstm import block_orthogonal, _HighwayLayerFunction
from lib import tc_ops
from lib.pytorch_misc import to_device_async


def get_dropout_mask(dropout_probability, tensor_for_masking):
    """Bernoulli keep-mask scaled by 1/(1-p), shaped like the given tensor (decoder_rnn.py:13-37)."""
    binary_mask = (torch.rand(tensor_for_masking.size(), device=tensor_for_masking.device) > dropout_probability)
    return binary_mask.float().div(1.0 - dropout_probability)


class DecoderRNN(torch.nn.Module):
    def __init__(self, classes, embed_dim, inputs_dim, hidden_dim, recurrent_dropout_probability=0.2,
                 use_highway=True, use_input_projection_bias=True):
        super().__init__()
        self.classes = classes
        # decoder_rnn.py:56-58: the table really is ['start'] + classes rows of width 100, whatever
        # embed_dim says (the declared nn.Embedding is overwritten by the word vectors)
        embed_vecs = obj_edge_vectors(['start'] + self.classes, wv_dim=100)
        self.obj_embed = nn.Embedding(len(self.classes) + 1, 100)
        self.obj_embed.weight.data = embed_vecs
        self.hidden_size = hidden_dim
        self.inputs_dim = inputs_dim
        self.nms_thresh = 0.3
        self.recurrent_dropout_probability = recurrent_dropout_probability
        self.use_highway = use_highway
        if not use_highway:
            raise NotImplementedError("only the highway cell (the reference's default) is built")
        self.input_linearity = torch.nn.Linear(self.input_size, 6 * self.hidden_size, bias=use_input_projection_bias)
        self.state_linearity = torch.nn.Linear(self.hidden_size, 5 * self.hidden_size, bias=True)
        self.out = nn.Linear(self.hidden_size, len(self.classes))
        self.reset_parameters()

    @property
    def input_size(self):
        return self.inputs_dim + self.obj_embed.weight.size(1)

    def reset_parameters(self):
        """decoder_rnn.py:85-94."""
        block_orthogonal(self.input_linearity.weight.data, [self.hidden_size, self.input_size])
        block_orthogonal(self.state_linearity.weight.data, [self.hidden_size, self.hidden_size])
        self.state_linearity.bias.data.fill_(0.0)
        self.state_linearity.bias.data[self.hidden_size:2 * self.hidden_size].fill_(1.0)

    def lstm_equations(self, timestep_input, previous_state, previous_memory, dropout_mask=None):
        """One cell step (decoder_rnn.py:96-131), GEMMs on the tcgen05 path."""
        H = self.hidden_size
        pi = tc_ops.linear_tc(timestep_input, self.input_linearity.weight, self.input_linearity.bias)
        ps = tc_ops.linear_tc(previous_state, self.state_linearity.weight, self.state_linearity.bias)
        input_gate = torch.sigmoid(pi[:, 0 * H:1 * H] + ps[:, 0 * H:1 * H])
        forget_gate = torch.sigmoid(pi[:, 1 * H:2 * H] + ps[:, 1 * H:2 * H])
        memory_init = torch.tanh(pi[:, 2 * H:3 * H] + ps[:, 2 * H:3 * H])
        output_gate = torch.sigmoid(pi[:, 3 * H:4 * H] + ps[:, 3 * H:4 * H])
        memory = input_gate * memory_init + forget_gate * previous_memory
        timestep_output = output_gate * torch.tanh(memory)
        highway_gate = torch.sigmoid(pi[:, 4 * H:5 * H] + ps[:, 4 * H:5 * H])
        timestep_output = highway_gate * timestep_output + (1 - highway_gate) * pi[:, 5 * H:6 * H]
        if dropout_mask is not None and self.training:
            timestep_output = timestep_output * dropout_mask
        return timestep_output, memory

    # ------------------------------------------------------------------ teacher-forced fast path
    def _forward_teacher_forced(self, sequence_tensor, batch_lengths, labels, dropout_mask):
        H = self.hidden_size
        dev = sequence_tensor.device
        bl = np.asarray(batch_lengths, dtype=np.int64)
        T, B = len(bl), int(bl[0])
        off = np.concatenate(([0], np.cumsum(bl)[:-1]))
        N = int(bl.sum())
        t_of = np.repeat(np.arange(T), bl)
        b_of = np.arange(N) - off[t_of]
        src = np.where(t_of > 0, off[np.maximum(t_of - 1, 0)] + b_of, -1)       # packed position of (t-1, b)
        src_d = to_device_async(src, dev)
        prev_idx = torch.where(src_d >= 0, labels[src_d.clamp_min(0)] + 1, torch.zeros_like(src_d))
        x = torch.cat((sequence_tensor, self.obj_embed(prev_idx)), 1)
        P = tc_ops.linear_tc(x, self.input_linearity.weight, self.input_linearity.bias)          # [N,6H]
        flat = to_device_async(t_of * B + b_of, dev)
        P_pad = torch.zeros(T * B, 6 * H, device=dev, dtype=torch.float32).index_copy(0, flat, P).view(T, B, 6 * H)
        lengths = (bl[None, :] > np.arange(B)[:, None]).sum(1)                                     # per sequence
        lengths_dev = to_device_async(lengths, dev, torch.int32)
        if dropout_mask is None or not self.training:
            dropout_mask = torch.ones(B, H, device=dev, dtype=torch.float32)
        wh = self.state_linearity.weight.t()          # [H,5H] view; the Function makes it contiguous
        save = torch.is_grad_enabled()
        h = _HighwayLayerFunction.apply(P_pad, wh, self.state_linearity.bias, dropout_mask.contiguous(), lengths_dev, 0, save)
        h_packed = h.reshape(T * B, H)[flat]
        dists = tc_ops.linear_tc(h_packed, self.out.weight, self.out.bias)
        return dists, labels.clone()

    def _commit_with_overlaps(self, boxes, probs):
        import motifs_cabi as _c
        N, C = probs.shape
        dev = probs.device
        if N * C * 4 <= 200 * 1024:
            boxes = boxes.contiguous().float(); probs = probs.contiguous().float()
            commit = torch.empty(N, dtype=torch.long, device=dev)
            with torch.cuda.device(dev):
                rc = _c.load().mb200_decoder_commit(_c.ptr(boxes), _c.ptr(probs), N, C, float(self.nms_thresh), _c.ptr(commit),
                                                    _c.cur_stream())
            _c.check(rc, "mb200_decoder_commit")
            return commit
        is_overlap = nms_overlaps(boxes).view(N, N, C).cpu().numpy() >= self.nms_thresh
        sampled = probs.cpu().numpy()
        sampled[:, 0] = 0
        commit = np.zeros(N, dtype=np.int64)
        for i in range(N):
            box_ind, cls_ind = np.unravel_index(sampled.argmax(), sampled.shape)
            commit[int(box_ind)] = int(cls_ind)
            sampled[is_overlap[box_ind, :, cls_ind], cls_ind] = 0.0
            sampled[box_ind] = -1.0
        return to_device_async(commit, dev)

    def forward(self, inputs, initial_state=None, labels=None, boxes_for_nms=None, dropout_mask=None, labels_all_fg=None):
        """`labels_all_fg` (superset of the reference signature): whether every label is foreground, when the caller
        already holds the labels on the host; None = read it back here (one D2H)."""
        if not isinstance(inputs, PackedSequence):
            raise ValueError('inputs must be PackedSequence but got %s' % (type(inputs)))
        sequence_tensor, batch_lengths = inputs[0], inputs[1]
        batch_lengths = [int(b) for b in batch_lengths]
        batch_size = batch_lengths[0]
        dev = sequence_tensor.device
        H = self.hidden_size
        if dropout_mask is None and self.recurrent_dropout_probability > 0.0:
            dropout_mask = get_dropout_mask(self.recurrent_dropout_probability,
                                            torch.empty(batch_size, H, device=dev))

        if self.training and initial_state is None and labels is not None and \
                (labels_all_fg if labels_all_fg is not None else bool((labels > 0).all())):
            return self._forward_teacher_forced(sequence_tensor, batch_lengths, labels, dropout_mask)

        # ---------------------------------------------------------------- step loop (decoder_rnn.py:160-227)
        if initial_state is None:
            previous_memory = sequence_tensor.new_zeros(batch_size, H)
            previous_state = sequence_tensor.new_zeros(batch_size, H)
        else:
            previous_state = initial_state[0].squeeze(0)
            previous_memory = initial_state[1].squeeze(0)
        previous_embed = self.obj_embed.weight[0, None].expand(batch_size, 100)
        out_dists, out_commitments = [], []
        end_ind = 0
        for i, l_batch in enumerate(batch_lengths):
            start_ind, end_ind = end_ind, end_ind + l_batch
            if previous_memory.size(0) != l_batch:
                previous_memory = previous_memory[:l_batch]
                previous_state = previous_state[:l_batch]
                previous_embed = previous_embed[:l_batch]
                if dropout_mask is not None:
                    dropout_mask = dropout_mask[:l_batch]
            timestep_input = torch.cat((sequence_tensor[start_ind:end_ind], previous_embed), 1)
            previous_state, previous_memory = self.lstm_equations(timestep_input, previous_state, previous_memory,
                                                                  dropout_mask=dropout_mask)
            pred_dist = tc_ops.linear_tc(previous_state, self.out.weight, self.out.bias)
            out_dists.append(pred_dist)
            if self.training:
                labels_to_embed = labels[start_ind:end_ind].clone()
                nonzero_pred = pred_dist[:, 1:].max(1)[1] + 1
                is_bg = labels_to_embed == 0
                labels_to_embed = torch.where(is_bg, nonzero_pred, labels_to_embed)
                out_commitments.append(labels_to_embed)
                previous_embed = self.obj_embed(labels_to_embed + 1)
            else:
                assert l_batch == 1
                best_ind = F.softmax(pred_dist, dim=1)[:, 1:].max(1)[1] + 1
                out_commitments.append(best_ind)
                previous_embed = self.obj_embed(best_ind + 1)

        if boxes_for_nms is not None and not self.training:
            # overlap-aware greedy commitment (decoder_rnn.py:230-247): one single-CTA kernel (csrc/boxes.cu) instead of the
            # reference's D2H of the [N,N,C] overlaps + host loop; the host loop remains for > ~330 detections
            out_commitments = self._commit_with_overlaps(boxes_for_nms.detach(), F.softmax(torch.cat(out_dists, 0), 1).detach())
        else:
            out_commitments = torch.cat(out_commitments, 0)
        return torch.cat(out_dists, 0), out_commitments
