"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement (plain torch fp32, differentiable) of the
MotifNet forward of the reference: lib/object_detector.py (`ObjectDetector` :50-423, `filter_det`
:425-485, `RPNHead` :488-597, `filter_roi_proposals` :600-612, `load_vgg` :623-633),
lib/rel_model.py (`_sort_by_score` :31-61, `LinearizedContext` :66-296, `RelModel` :299-560),
lib/lstm/decoder_rnn.py (:40-251), lib/get_union_boxes.py (:15-93), lib/sparse_targets.py (:32-37),
lib/surgery.py (:21-59), lib/fpn/proposal_assignments/proposal_assignments_gtbox.py (:9-87).

PINNING. The reference cannot run as shipped here (PyTorch 0.3 API, CUDA-only operators, needs VG data — SURVEY.md
§8c), but its own model code CAN be executed on the CPU once its three CUDA extensions are replaced by the operator
restatements of oracle/ops.py / oracle/highway_lstm.py (themselves pinned on the GPU against the reference's .cu
files) and PyTorch-0.3 semantics are shimmed: tests/golden/make_golden_model.py does that, and
tests/test_reference_model_pin.py holds this file to the outputs of the reference's RelModel — PredCls / SGCls /
SGDet eval tuples and an SGCls training forward + backward (logits 1e-4, loss, gradients of every trainable
parameter 1e-3) — with both sides loading the same synthetic state dict (same keys and shapes).
Module and parameter names equal the reference's, so one state dict drives the reference, the oracle and the
product. All randomness (dropout masks, sampling RNG) is injected."""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F
from torchvision.models.vgg import vgg16

from . import host, ops
from .highway_lstm import highway_lstm_forward

BATCHNORM_MOMENTUM = 0.01
RELS_PER_IMG, REL_FG_FRACTION = 256, 0.25


def t2n(t):
    return t.detach().cpu().numpy()


def roi_align(features, rois, size=7, scale=1 / 16):
    """functions/roi_align.py:17-48 (no gradient: every caller on this path detaches the map)."""
    f = t2n(features)
    r = ops.normalize_rois(t2n(rois), f.shape[2], f.shape[3], scale)
    return torch.from_numpy(ops.roi_align_forward(f, r, size, size))


def load_vgg(use_dropout=True, use_relu=True, use_linear=True):
    model = vgg16(weights=None)
    del model.features._modules['30']
    del model.classifier._modules['6']
    if not use_dropout:
        del model.classifier._modules['5']
        if not use_relu:
            del model.classifier._modules['4']
            if not use_linear:
                del model.classifier._modules['3']
    return model


def run_classifier(classifier, x, masks, prefix):
    for name, m in classifier._modules.items():
        if isinstance(m, nn.Dropout):
            key = prefix + name
            if not classifier.training:
                continue
            if masks is not None and key in masks:
                x = x * masks[key]
            else:
                raise RuntimeError("oracle needs the dropout mask %s injected" % key)
        else:
            x = m(x)
    return x


class Result(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


class AlternatingHighwayLSTM(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers=1, recurrent_dropout_probability=0):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        n = sum(6 * hidden_size * (input_size if l == 0 else hidden_size) + 5 * hidden_size * hidden_size
                for l in range(num_layers))
        self.weight = nn.Parameter(torch.zeros(n))
        self.bias = nn.Parameter(torch.zeros(5 * hidden_size * num_layers))

    def forward(self, x_packed, batch_sizes, dropout):
        """x_packed [N,In] time-major packed; returns packed outputs [N,H]."""
        bs = np.asarray(batch_sizes)
        T, B = len(bs), int(bs[0])
        lengths = [int((bs > b).sum()) for b in range(B)]
        off = np.concatenate(([0], np.cumsum(bs)[:-1]))
        padded = x_packed.new_zeros(T, B, x_packed.size(1))
        idx_t = np.repeat(np.arange(T), bs)
        idx_b = np.arange(int(bs.sum())) - off[idx_t]
        flat = torch.as_tensor(idx_t * B + idx_b)
        padded = padded.view(T * B, -1).index_copy(0, flat, x_packed).view(T, B, -1)
        if dropout is None:
            dropout = torch.ones(self.num_layers, B, self.hidden_size)
        out = highway_lstm_forward(padded, lengths, self.weight, self.bias, dropout, self.hidden_size, self.num_layers)
        return out.reshape(T * B, -1)[flat]


class DecoderRNN(nn.Module):
    """lib/lstm/decoder_rnn.py:40-251, step loop as written there."""

    def __init__(self, classes, inputs_dim, hidden_dim):
        super().__init__()
        self.classes = classes
        self.obj_embed = nn.Embedding(len(classes) + 1, 100)
        self.hidden_size = hidden_dim
        self.input_linearity = nn.Linear(inputs_dim + 100, 6 * hidden_dim)
        self.state_linearity = nn.Linear(hidden_dim, 5 * hidden_dim)
        self.out = nn.Linear(hidden_dim, len(classes))
        self.nms_thresh = 0.3

    def lstm_equations(self, x, h, c, mask):
        H = self.hidden_size
        pi, ps = self.input_linearity(x), self.state_linearity(h)
        i = torch.sigmoid(pi[:, :H] + ps[:, :H])
        f = torch.sigmoid(pi[:, H:2 * H] + ps[:, H:2 * H])
        g = torch.tanh(pi[:, 2 * H:3 * H] + ps[:, 2 * H:3 * H])
        o = torch.sigmoid(pi[:, 3 * H:4 * H] + ps[:, 3 * H:4 * H])
        mem = i * g + f * c
        out = o * torch.tanh(mem)
        r = torch.sigmoid(pi[:, 4 * H:5 * H] + ps[:, 4 * H:5 * H])
        out = r * out + (1 - r) * pi[:, 5 * H:6 * H]
        if mask is not None and self.training:
            out = out * mask
        return out, mem

    def forward(self, seq, batch_lengths, labels=None, boxes_for_nms=None, dropout_mask=None):
        bsz = int(batch_lengths[0])
        c = seq.new_zeros(bsz, self.hidden_size)
        h = seq.new_zeros(bsz, self.hidden_size)
        prev = self.obj_embed.weight[0, None].expand(bsz, 100)
        dists, commits, end = [], [], 0
        for l in [int(b) for b in batch_lengths]:
            start, end = end, end + l
            if c.size(0) != l:
                c, h, prev = c[:l], h[:l], prev[:l]
                if dropout_mask is not None:
                    dropout_mask = dropout_mask[:l]
            h, c = self.lstm_equations(torch.cat((seq[start:end], prev), 1), h, c, dropout_mask)
            pd = self.out(h)
            dists.append(pd)
            if self.training:
                lab = labels[start:end].clone()
                nz = pd[:, 1:].max(1)[1] + 1
                lab = torch.where(lab == 0, nz, lab)
                commits.append(lab)
                prev = self.obj_embed(lab + 1)
            else:
                assert l == 1
                best = F.softmax(pd, 1)[:, 1:].max(1)[1] + 1
                commits.append(best)
                prev = self.obj_embed(best + 1)
        if boxes_for_nms is not None and not self.training:
            n = boxes_for_nms.size(0)
            is_overlap = ops.nms_overlaps(t2n(boxes_for_nms)) >= self.nms_thresh
            sampled = t2n(F.softmax(torch.cat(dists, 0), 1)).copy()
            sampled[:, 0] = 0
            out = np.zeros(len(commits), dtype=np.int64)
            for _ in range(out.shape[0]):
                bi, ci = np.unravel_index(sampled.argmax(), sampled.shape)
                out[int(bi)] = int(ci)
                sampled[is_overlap[bi, :, ci], ci] = 0.0
                sampled[bi] = -1.0
            commits = torch.from_numpy(out)
        else:
            commits = torch.cat(commits, 0)
        return torch.cat(dists, 0), commits


def sort_by_score(im_inds, scores):
    """rel_model.py:31-61."""
    im = t2n(im_inds)
    num_im = int(im[-1]) + 1
    rpi = torch.zeros(num_im)
    lengths = []
    for i, s, e in host.enumerate_by_image(im):
        rpi[i] = 2 * (s - e) * num_im + i
        lengths.append(e - s)
    lengths = sorted(lengths, reverse=True)
    inds, ls = host.transpose_packed_sequence_inds(lengths)
    roi_order = scores - 2 * rpi[im_inds]
    _, perm = torch.sort(roi_order, dim=0, descending=True, stable=True)
    perm = perm[torch.as_tensor(inds)]
    _, inv = torch.sort(perm)
    return perm, inv, ls


def center_size(b):
    wh = b[:, 2:] - b[:, :2] + 1.0
    return torch.cat((b[:, :2] + 0.5 * wh, wh), 1)


class LinearizedContext(nn.Module):
    def __init__(self, classes, rel_classes, mode, embed_dim, hidden_dim, obj_dim, nl_obj, nl_edge, order,
                 pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False):
        super().__init__()
        self.classes, self.rel_classes, self.mode, self.order = classes, rel_classes, mode, order
        # rel_model.py:115-117 sizes the decoder input as hidden + obj_dim + embed_dim while :213 feeds it
        # obj_dim + embed_dim + 128 + hidden: with pass_in_obj_feats_to_decoder the reference only runs in predcls
        # (no decoder call); the scripts never set the flag. Restated for predcls only.
        assert not pass_in_obj_feats_to_decoder or mode == 'predcls', "the reference itself fails on this combination"
        self.to_decoder, self.to_edge = pass_in_obj_feats_to_decoder, pass_in_obj_feats_to_edge
        self.nl_obj, self.nl_edge = nl_obj, nl_edge
        nc = len(classes)
        self.obj_embed = nn.Embedding(nc, embed_dim)
        self.obj_embed2 = nn.Embedding(nc, embed_dim)
        self.pos_embed = nn.Sequential(nn.BatchNorm1d(4, momentum=BATCHNORM_MOMENTUM / 10.0), nn.Linear(4, 128),
                                       nn.ReLU(inplace=True), nn.Dropout(0.1))
        if nl_obj > 0:
            self.obj_ctx_rnn = AlternatingHighwayLSTM(obj_dim + embed_dim + 128, hidden_dim, nl_obj)
            self.decoder_rnn = DecoderRNN(classes, hidden_dim + (obj_dim + embed_dim if self.to_decoder else 0), hidden_dim)
        else:                           # the scripts' "baseline" (-nl_obj 0 -nl_edge 0): a linear object classifier, :125-126
            self.decoder_lin = nn.Linear(obj_dim + embed_dim + 128, nc)
        if nl_edge > 0:                 # :128-137
            self.edge_ctx_rnn = AlternatingHighwayLSTM(embed_dim + (hidden_dim if nl_obj > 0 else 0)
                                                       + (obj_dim if self.to_edge else 0), hidden_dim, nl_edge)
        self.masks = None

    def sort_rois(self, batch_idx, confidence, box_priors):
        cx = center_size(box_priors)
        if self.order == 'size':
            sizes = cx[:, 2] * cx[:, 3]
            scores = sizes / (sizes.max() + 1)
        elif self.order == 'confidence':
            scores = confidence
        elif self.order == 'leftright':
            scores = cx[:, 0] / (cx[:, 0].max() + 1)
        else:
            raise ValueError(self.order)
        return sort_by_score(batch_idx, scores)

    def forward(self, obj_fmaps, obj_logits, im_inds, obj_labels, box_priors, boxes_per_cls):
        m = (self.masks or {}) if self.training else {}
        nc = len(self.classes)
        obj_embed = F.softmax(obj_logits, 1) @ self.obj_embed.weight
        pe = self.pos_embed
        pos = pe[2](pe[1](pe[0](center_size(box_priors))))
        if self.training:
            pos = pos * m["pos_embed.3"]
        obj_pre_rep = torch.cat((obj_fmaps, obj_embed, pos), 1)
        if self.nl_obj > 0:
            # obj_ctx (:197-234)
            confidence = F.softmax(obj_logits, 1).detach()[:, 1:].max(1)[0]
            perm, inv, ls = self.sort_rois(im_inds, confidence, box_priors)
            inp = obj_pre_rep[perm].contiguous()
            enc = self.obj_ctx_rnn(inp, ls, m.get("obj_ctx_rnn"))
            if self.mode != 'predcls':
                d, p = self.decoder_rnn(enc, ls, labels=obj_labels[perm] if obj_labels is not None else None,
                                        boxes_for_nms=boxes_per_cls[perm] if boxes_per_cls is not None else None,
                                        dropout_mask=m.get("decoder_rnn"))
                obj_preds, obj_dists2 = p[inv], d[inv]
            else:
                obj_preds = obj_labels
                obj_dists2 = torch.full((obj_labels.size(0), nc), -1000.0)
                obj_dists2[torch.arange(obj_labels.size(0)), obj_labels] = 1000.0
            obj_ctx = enc[inv]
        else:
            # no object context (:259-283): linear classifier; SGDet eval picks labels through a per-class NMS
            if self.mode == 'predcls':
                obj_dists2 = torch.full((obj_labels.size(0), nc), -1000.0)
                obj_dists2[torch.arange(obj_labels.size(0)), obj_labels] = 1000.0
            else:
                obj_dists2 = self.decoder_lin(obj_pre_rep)
            if self.mode == 'sgdet' and not self.training:
                probs = F.softmax(obj_dists2, 1)
                nms_mask = torch.zeros_like(probs)
                for c in range(1, nc):
                    keep = ops.apply_nms(t2n(probs[:, c]), t2n(boxes_per_cls[:, c]), pre_nms_topn=probs.size(0),
                                         post_nms_topn=probs.size(0), nms_thresh=0.3)
                    nms_mask[:, c][torch.from_numpy(np.asarray(keep, dtype=np.int64))] = 1
                obj_preds = (nms_mask * probs)[:, 1:].max(1)[1] + 1
            else:
                obj_preds = obj_labels if obj_labels is not None else obj_dists2[:, 1:].max(1)[1] + 1
            obj_ctx = obj_pre_rep
        if self.nl_edge == 0:
            return obj_dists2, obj_preds, None
        # edge_ctx (:171-195)
        edge_in = torch.cat((obj_fmaps, obj_ctx), 1) if self.to_edge else obj_ctx            # :287
        inp_feats = torch.cat((self.obj_embed2(obj_preds), edge_in), 1)
        conf = F.softmax(obj_dists2.detach(), 1).view(-1)[obj_preds + torch.arange(obj_preds.size(0)) * nc]
        perm, inv, ls = self.sort_rois(im_inds, conf, box_priors)
        edge = self.edge_ctx_rnn(inp_feats[perm], ls, m.get("edge_ctx_rnn"))[inv]
        return obj_dists2, obj_preds, edge


class UnionBoxesAndFeats(nn.Module):
    def __init__(self, pooling_size=7, stride=16, dim=512):
        super().__init__()
        self.pooling_size, self.stride = pooling_size, stride
        self.conv = nn.Sequential(
            nn.Conv2d(2, dim // 2, kernel_size=7, stride=2, padding=3, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(dim // 2, momentum=BATCHNORM_MOMENTUM), nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(dim // 2, dim, kernel_size=3, stride=1, padding=1, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(dim, momentum=BATCHNORM_MOMENTUM))

    def forward(self, fmap, rois, union_inds):
        r = t2n(rois)
        ui = t2n(union_inds)
        u = ops.union_rois(r, ui)
        pools = roi_align(fmap, torch.from_numpy(u), self.pooling_size, 1 / self.stride)
        pairs = np.concatenate((r[:, 1:][ui[:, 0]], r[:, 1:][ui[:, 1]]), 1)
        rects = torch.from_numpy(ops.draw_union_boxes(pairs, self.pooling_size * 4 - 1) - np.float32(0.5))
        return pools + self.conv(rects)


class RPNHead(nn.Module):
    def __init__(self, dim=512, input_dim=512):
        super().__init__()
        self.stride, self.A = 16, len(host.ANCHOR_RATIOS) * len(host.ANCHOR_SCALES)
        self.conv = nn.Sequential(nn.Conv2d(input_dim, dim, 3, padding=1), nn.ReLU6(inplace=True),
                                  nn.Conv2d(dim, 6 * self.A, 1))
        self.register_buffer('anchors', torch.FloatTensor(host.generate_anchors()))

    def forward(self, fmap):
        x = self.conv(fmap)
        B, nc, h, w = x.shape
        return x.view(B, nc, -1).transpose(1, 2).contiguous().view(B, h, w, self.A, 6)

    def roi_proposals(self, fmap, im_sizes, nms_thresh=0.7, pre_nms_topn=6000, post_nms_topn=1000):
        cls = F.softmax(fmap[..., :2], 4)[..., 1].detach().contiguous()
        box_fmap = fmap[..., 2:].detach().contiguous()
        B = fmap.size(0)
        anchors = torch.cat([self.anchors[None]] * B, 0)
        bp = torch.from_numpy(ops.bbox_preds(t2n(anchors.view(-1, 4)), t2n(box_fmap.view(-1, 4)))).view(*box_fmap.shape)
        for i, (h, w, scale) in enumerate(np.asarray(im_sizes)):
            h_end, w_end = int(h) // self.stride, int(w) // self.stride
            if h_end < cls.size(1):
                cls[i, h_end:] = -0.01
            if w_end < cls.size(2):
                cls[i, :, w_end:] = -0.01
            bp[i, ..., 0].clamp_(min=0, max=w - 1); bp[i, ..., 1].clamp_(min=0, max=h - 1)
            bp[i, ..., 2].clamp_(min=0, max=w - 1); bp[i, ..., 3].clamp_(min=0, max=h - 1)
        sizes = center_size(bp.view(-1, 4))
        cls.view(-1)[(sizes[:, 2] < 4) | (sizes[:, 3] < 4)] = -0.01
        per = int(np.prod(bp.shape[1:-1]))
        inds, im_per = ops.apply_nms(t2n(cls.view(-1)), t2n(bp.view(-1, 4)), pre_nms_topn, post_nms_topn,
                                     boxes_per_im=[per] * B, nms_thresh=nms_thresh)
        img = np.concatenate([np.full(n, v, np.float32) for v, n in enumerate(im_per)])
        return torch.from_numpy(np.concatenate((img[:, None], t2n(bp.view(-1, 4))[inds]), 1))


def filter_det(scores, boxes, start_ind=0, max_per_img=100, thresh=0.001, pre_nms_topn=6000, post_nms_topn=300,
               nms_thresh=0.3):
    """object_detector.py:425-485 (nms_filter_duplicates=True branch)."""
    s, b = t2n(scores), t2n(boxes)
    valid = np.where(s[:, 1:].max(0) > thresh)[0] + 1
    if valid.size == 0:
        return None
    mask = np.zeros_like(s)
    for c in valid:
        keep = ops.apply_nms(s[:, c], b[:, c], pre_nms_topn, post_nms_topn, nms_thresh=nms_thresh)
        mask[keep, c] = 1
    d = mask * s
    sp, lp = d.max(1), d.argmax(1)
    inds = np.nonzero(sp)[0]
    la, sa = lp[inds], sp[inds]
    idx = np.argsort(-sa, kind="stable")
    idx = idx[sa[idx] > thresh][:max_per_img]
    return torch.from_numpy(inds[idx] + start_ind), torch.from_numpy(sa[idx]), torch.from_numpy(la[idx])


def load_resnet():
    """lib/object_detector.py:615-620: torchvision resnet101 minus layer4 / avgpool / fc (third-party arithmetic
    boundary, SURVEY.md section 8c: the architecture comes from torchvision here as it does in the reference)."""
    from torchvision.models.resnet import resnet101
    model = resnet101(weights=None)
    del model.layer4
    del model.avgpool
    del model.fc
    return model


class ObjectDetector(nn.Module):
    def __init__(self, classes, mode='gtbox', max_per_img=64, thresh=0.05, use_resnet=False):
        super().__init__()
        self.classes, self.mode, self.max_per_img, self.thresh = classes, mode, max_per_img, thresh
        self.use_resnet = use_resnet
        if not use_resnet:
            vgg = load_vgg()
            self.features, self.roi_fmap = vgg.features, vgg.classifier
            rpn_input_dim, output_dim = 512, 4096
        else:                               # lib/object_detector.py:84-101 ("Deprecated" there, BASELINE config 3)
            self.features = load_resnet()
            self.compress = nn.Sequential(nn.Conv2d(1024, 256, kernel_size=1), nn.ReLU(inplace=True), nn.BatchNorm2d(256))
            self.roi_fmap = nn.Sequential(nn.Linear(256 * 7 * 7, 2048), nn.SELU(inplace=True), nn.AlphaDropout(p=0.05),
                                          nn.Linear(2048, 2048), nn.SELU(inplace=True), nn.AlphaDropout(p=0.05))
            rpn_input_dim, output_dim = 1024, 2048
        self.score_fc = nn.Linear(output_dim, len(classes))
        self.bbox_fc = nn.Linear(output_dim, len(classes) * 4)
        self.rpn_head = RPNHead(512, rpn_input_dim)
        self.masks = None
        self.rng = np.random

    def feature_map(self, x):
        """lib/object_detector.py:110-127."""
        if not self.use_resnet:
            return self.features(x)
        f = self.features
        x = f.maxpool(f.relu(f.bn1(f.conv1(x))))
        return f.layer3(f.layer2(f.layer1(x)))

    def forward(self, x, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None, train_anchor_inds=None,
                proposals=None):
        with torch.no_grad():
            fmap = self.feature_map(x)
            rel_labels = obj_labels = None
            if self.mode == 'gtbox':
                im_inds = gt_classes[:, 0] - image_offset
                rois = torch.cat((im_inds.float()[:, None], gt_boxes), 1)
                if gt_rels is not None and self.training:
                    rel_labels = proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, image_offset, self.rng)
                obj_labels = gt_classes[:, 1]
            elif self.training and self.mode == 'rpntrain':
                # detector training (models/train_detector.py; object_detector.py:140-191): proposals with the training
                # limits, RPN outputs at the sampled anchors, proposal -> GT assignment with the injected RNG
                rpn_feats = self.rpn_head(fmap)
                rois = self.rpn_head.roi_proposals(rpn_feats, im_sizes, pre_nms_topn=12000, post_nms_topn=2000)
                tai = train_anchor_inds.clone()
                tai[:, 0] -= image_offset
                picked = rpn_feats[tai[:, 0], tai[:, 1], tai[:, 2], tai[:, 3]]                  # gather_nd, :533-545
                rpn_scores, rpn_box_deltas = picked[:, :2], picked[:, 2:]
                r, l, tg = host.proposal_assignments_det(t2n(rois), t2n(gt_boxes), t2n(gt_classes), image_offset, self.rng)
                rois = torch.from_numpy(r)
                pool = roi_align(fmap, rois)
                obj_fmap = run_classifier(self.roi_fmap, pool.view(rois.size(0), -1), self.masks, "roi_fmap.")
                return Result(od_obj_dists=self.score_fc(obj_fmap),
                              od_box_deltas=self.bbox_fc(obj_fmap).view(-1, len(self.classes), 4),
                              od_obj_labels=torch.from_numpy(l), od_box_targets=torch.from_numpy(tg),
                              od_box_priors=rois[:, 1:], rpn_scores=rpn_scores, rpn_box_deltas=rpn_box_deltas,
                              rois=rois, fmap=fmap)
            elif self.mode == 'proposals':
                # pre-computed proposals [n, 6] = (image, score, x1, y1, x2, y2), 2000 per image (object_detector.py:216-258,
                # filter_roi_proposals :600-612)
                assert proposals is not None and not self.training
                B = len(im_sizes)
                inds, im_per = ops.apply_nms(t2n(proposals[:, 1]), t2n(proposals[:, 2:]), 6000, 1000,
                                             boxes_per_im=[2000] * B, nms_thresh=0.7)
                img = np.concatenate([np.full(n, v, np.float32) for v, n in enumerate(im_per)])
                rois = torch.from_numpy(np.concatenate((img[:, None], t2n(proposals[:, 2:])[inds]), 1).astype(np.float32))
            else:
                rois = self.rpn_head.roi_proposals(self.rpn_head(fmap), im_sizes)
            pool = roi_align(self.compress(fmap) if self.use_resnet else fmap, rois)        # :136-137
            obj_fmap = run_classifier(self.roi_fmap, pool.view(rois.size(0), -1), self.masks, "roi_fmap.")
            od_obj_dists = self.score_fc(obj_fmap)
            if self.mode == 'gtbox':
                return Result(rm_obj_dists=od_obj_dists, rm_obj_labels=obj_labels, rm_box_priors=rois[:, 1:],
                              boxes_all=None, rel_labels=rel_labels, im_inds=rois[:, 0].long() + image_offset, fmap=fmap,
                              od_obj_dists=od_obj_dists, obj_fmap=obj_fmap, rois=rois)
            deltas = self.bbox_fc(obj_fmap).view(-1, len(self.classes), 4)
            N, K = deltas.shape[:2]
            boxes = torch.from_numpy(ops.bbox_preds(t2n(rois[:, None, 1:].expand(N, K, 4).reshape(-1, 4)),
                                                    t2n(deltas.reshape(-1, 4)))).view(N, K, 4)
            inds = rois[:, 0].long()
            dets = []
            for i, s, e in host.enumerate_by_image(t2n(inds)):
                h, w = np.asarray(im_sizes)[i, :2]
                boxes[s:e, :, 0].clamp_(min=0, max=w - 1); boxes[s:e, :, 1].clamp_(min=0, max=h - 1)
                boxes[s:e, :, 2].clamp_(min=0, max=w - 1); boxes[s:e, :, 3].clamp_(min=0, max=h - 1)
                d = filter_det(F.softmax(od_obj_dists[s:e], 1), boxes[s:e], start_ind=s, max_per_img=self.max_per_img,
                               thresh=self.thresh)
                if d is not None:
                    dets.append(d)
            if not dets:
                return None
            nms_inds, nms_scores, nms_labels = [torch.cat(z, 0) for z in zip(*dets)]
            nms_boxes = torch.cat((rois[:, 1:][nms_inds][:, None], boxes[nms_inds][:, 1:]), 1)
            rm_obj_labels = None
            if self.training:           # SGDet training (:316-326): label the detections by IoU >= 0.5 with a GT box of the image
                ov = torch.from_numpy(ops.bbox_overlaps_f32(t2n(nms_boxes[:, 0]), t2n(gt_boxes)))
                ov[(inds[nms_inds] + image_offset)[:, None] != gt_classes[None, :, 0]] = 0.0
                mx, am = ov.max(1)
                rm_obj_labels = gt_classes[:, 1][am].clone()
                rm_obj_labels[mx < 0.5] = 0
            return Result(rm_obj_dists=od_obj_dists[nms_inds], rm_obj_labels=rm_obj_labels, rm_box_priors=nms_boxes[:, 0],
                          boxes_all=nms_boxes, rel_labels=None, im_inds=inds[nms_inds] + image_offset, fmap=fmap,
                          od_obj_dists=od_obj_dists, obj_fmap=obj_fmap[nms_inds], rois=rois, obj_scores=nms_scores,
                          obj_preds=nms_labels)


def proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, image_offset, rng):
    """proposal_assignments_gtbox.py:26-87, numpy; rng.choice calls in the reference's order."""
    im_inds = t2n(rois[:, 0]).astype(np.int64)
    num_im = int(im_inds[-1]) + 1
    n = im_inds.shape[0]
    fg = t2n(gt_rels).copy()
    fg[:, 0] -= image_offset
    offset = {i: s for i, s, e in host.enumerate_by_image(im_inds)}
    for i, s, e in host.enumerate_by_image(fg[:, 0]):
        fg[s:e, 1:3] += offset[i]
    is_cand = im_inds[:, None] == im_inds[None]
    is_cand[np.arange(n), np.arange(n)] = False
    is_cand.reshape(-1)[fg[:, 1] * n + fg[:, 2]] = False
    bgc = np.column_stack(np.nonzero(is_cand))
    num_fg = min(fg.shape[0], int(RELS_PER_IMG * REL_FG_FRACTION * num_im))
    if num_fg < fg.shape[0]:
        fg = fg[rng.choice(fg.shape[0], size=num_fg, replace=False)]
    num_bg = min(bgc.shape[0], int(RELS_PER_IMG * num_im) - num_fg)
    if num_bg > 0:
        bg = np.column_stack((im_inds[bgc[:, 0]], bgc, np.zeros(bgc.shape[0], np.int64)))
        if num_bg < bgc.shape[0]:
            bg = bg[rng.choice(bg.shape[0], size=num_bg, replace=False)]
        rel = np.concatenate((fg, bg), 0)
    else:
        rel = fg
    G = gt_boxes.size(0)
    perm = np.argsort(rel[:, 0] * (G ** 2) + rel[:, 1] * G + rel[:, 2], kind="stable")
    return torch.from_numpy(rel[perm])


class FrequencyBias(nn.Module):
    def __init__(self, num_objs, num_rels):
        super().__init__()
        self.num_objs = num_objs
        self.obj_baseline = nn.Embedding(num_objs * num_objs, num_rels)

    def index_with_labels(self, labels):
        return self.obj_baseline(labels[:, 0] * self.num_objs + labels[:, 1])


class Flattener(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class RelModel(nn.Module):
    def __init__(self, classes, rel_classes, mode='sgcls', embed_dim=200, hidden_dim=512, pooling_dim=4096,
                 nl_obj=2, nl_edge=4, order='leftright', thresh=0.01, use_bias=True, use_tanh=False,
                 limit_vision=False, require_overlap_det=True, pass_in_obj_feats_to_decoder=False,
                 pass_in_obj_feats_to_edge=False, use_proposals=False):
        super().__init__()
        self.classes, self.rel_classes, self.mode = classes, rel_classes, mode
        self.pooling_dim, self.use_bias, self.use_tanh, self.limit_vision = pooling_dim, use_bias, use_tanh, limit_vision
        self.require_overlap = require_overlap_det and mode == 'sgdet'
        self.detector = ObjectDetector(classes, mode=('proposals' if use_proposals else 'refinerels') if mode == 'sgdet'
                                       else 'gtbox', thresh=thresh)                  # rel_model.py:340-346
        self.context = LinearizedContext(classes, rel_classes, mode, embed_dim, hidden_dim, 4096, nl_obj, nl_edge, order,
                                         pass_in_obj_feats_to_decoder, pass_in_obj_feats_to_edge)
        self.union_boxes = UnionBoxesAndFeats(7, 16, 512)
        roi_fmap = [Flattener(), load_vgg(use_dropout=False, use_relu=False, use_linear=pooling_dim == 4096).classifier]
        if pooling_dim != 4096:                                 # rel_model.py:371-372
            roi_fmap.append(nn.Linear(4096, pooling_dim))
        self.roi_fmap = nn.Sequential(*roi_fmap)
        self.roi_fmap_obj = load_vgg().classifier
        self.post_lstm = nn.Linear(hidden_dim, pooling_dim * 2)
        if nl_edge == 0:                                        # rel_model.py:386-388
            self.post_emb = nn.Embedding(len(classes), pooling_dim * 2)
        self.rel_compress = nn.Linear(pooling_dim, len(rel_classes), bias=True)
        if use_bias:
            self.freq_bias = FrequencyBias(len(classes), len(rel_classes))
        self.masks = None

    def forward(self, x, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None, proposals=None):
        result = self.detector(x, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, proposals=proposals)
        im_inds = result.im_inds - image_offset
        boxes = result.rm_box_priors
        if self.training and result.rel_labels is None:
            # SGDet training (rel_model.py:479-487): relation labels for the DETECTED boxes
            assert self.mode == 'sgdet'
            result.rel_labels = torch.from_numpy(host.rel_assignments(
                t2n(im_inds), t2n(boxes), t2n(result.rm_obj_labels), t2n(gt_boxes), t2n(gt_classes), t2n(gt_rels),
                image_offset, self.detector.rng, filter_non_overlap=True, num_sample_per_gt=1))
        if self.training:
            rel_inds = result.rel_labels[:, :3].clone()
        else:
            cands = im_inds[:, None] == im_inds[None]
            cands.fill_diagonal_(False)
            if self.require_overlap:
                cands = cands & (torch.from_numpy(ops.bbox_overlaps_f32(t2n(boxes), t2n(boxes))) > 0)
            cands = cands.nonzero()
            rel_inds = torch.cat((im_inds[cands[:, 0]][:, None], cands), 1)
        rois = torch.cat((im_inds[:, None].float(), boxes), 1)
        pool = roi_align(result.fmap, rois)
        result.obj_fmap = run_classifier(self.roi_fmap_obj, pool.view(rois.size(0), -1), self.masks, "roi_fmap_obj.")
        result.rm_obj_dists, result.obj_preds, edge_ctx = self.context(
            result.obj_fmap, result.rm_obj_dists.detach(), im_inds,
            result.rm_obj_labels if self.training or self.mode == 'predcls' else None, boxes.detach(), result.boxes_all)
        edge_rep = (self.post_emb(result.obj_preds) if edge_ctx is None else self.post_lstm(edge_ctx)).view(-1, 2, self.pooling_dim)
        prod_rep = edge_rep[:, 0][rel_inds[:, 1]] * edge_rep[:, 1][rel_inds[:, 2]]
        ub = self.union_boxes(result.fmap, rois, rel_inds[:, 1:])
        vr = run_classifier(self.roi_fmap[1], ub.view(ub.size(0), -1), self.masks, "roi_fmap.1.")
        if len(self.roi_fmap) > 2:                              # pooling_dim != 4096: the extra projection, rel_model.py:371-372
            vr = self.roi_fmap[2](vr)
        if self.limit_vision:
            prod_rep = torch.cat((prod_rep[:, :2048] * vr[:, :2048], prod_rep[:, 2048:]), 1)
        else:
            prod_rep = prod_rep * vr
        if self.use_tanh:
            prod_rep = torch.tanh(prod_rep)
        result.rel_dists = self.rel_compress(prod_rep)
        if self.use_bias:
            result.rel_dists = result.rel_dists + self.freq_bias.index_with_labels(torch.stack((
                result.obj_preds[rel_inds[:, 1]], result.obj_preds[rel_inds[:, 2]]), 1))
        result.rel_inds = rel_inds
        self.last_result = result
        if self.training:
            return result
        nc = len(self.classes)
        twod = torch.arange(result.obj_preds.size(0)) * nc + result.obj_preds
        obj_scores = F.softmax(result.rm_obj_dists, 1).view(-1)[twod]
        if self.mode == 'sgdet':
            bboxes = result.boxes_all.view(-1, 4)[twod].view(result.boxes_all.size(0), 4)
        else:
            bboxes = result.rm_box_priors
        rel_rep = F.softmax(result.rel_dists, 1)
        # filter_dets, surgery.py:21-59
        s0, s1 = obj_scores[rel_inds[:, 1]], obj_scores[rel_inds[:, 2]]
        pm = rel_rep[:, 1:].max(1)[0]
        _, idx = torch.sort((pm * s0 * s1).view(-1), dim=0, descending=True, stable=True)
        return t2n(bboxes), t2n(result.obj_preds), t2n(obj_scores), t2n(rel_inds[:, 1:][idx]), t2n(rel_rep[idx])
