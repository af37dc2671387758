"""MotifNet relation model — same surface as the reference's lib/rel_model.py: `RelModel(classes,
rel_classes, mode, num_gpus, use_vision, require_overlap_det, embed_dim, hidden_dim, pooling_dim,
nl_obj, nl_edge, use_resnet, order, thresh, use_proposals, pass_in_obj_feats_to_decoder,
pass_in_obj_feats_to_edge, rec_dropout, use_bias, use_tanh, limit_vision)` (:303-308), `forward` with
the `Blob.__getitem__` tuple (:450-452) returning a `Result` in training and the `filter_dets`
5-tuple in eval (:533-547); module and parameter names match the reference's state dict
(detector.*, context.{obj_embed,obj_embed2,pos_embed,obj_ctx_rnn,decoder_rnn,edge_ctx_rnn},
union_boxes.conv.*, roi_fmap.1.{0,3}, roi_fmap_obj.{0,3}, post_lstm, rel_compress, freq_bias).

Every GEMM goes through the tcgen05 path (lib/tc_ops.py), the LSTMs through the persistent kernels
(csrc/lstm.cu), RoIAlign / union boxes / masks through csrc/roi_align.cu and csrc/boxes.cu."""
import math

import os

import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F
from torch.nn.utils.rnn import PackedSequence

from config import BATCHNORM_MOMENTUM
from lib import fused_optim, tc_ops
from lib.fpn.box_utils import bbox_overlaps, center_size
from lib.fpn.nms.functions.nms import apply_nms
from lib.fpn.roi_align.functions.roi_align import RoIAlignFunction, roi_align_from_nhwc
from lib.get_union_boxes import UnionBoxesAndFeats
from lib.lstm.decoder_rnn import DecoderRNN
from lib.lstm.highway_lstm_cuda.alternating_highway_lstm import AlternatingHighwayLSTM
from lib.object_detector import ObjectDetector, gather_res, load_vgg, run_classifier
from lib.pytorch_misc import transpose_packed_sequence_inds, to_onehot, arange, image_segments, Flattener, to_device_async
from lib.sparse_targets import FrequencyBias
from lib.surgery import filter_dets
from lib.word_vectors import obj_edge_vectors

MODES = ('sgdet', 'sgcls', 'predcls')

# With GT boxes (sgcls / predcls) the image index of every object is an INPUT (gt_classes[:, 0]): read it back once,
# before the backbone is queued, and hand the host copy to the code that builds the packed-sequence order, instead of
# D2H reads in the middle of forward (each one drains the stream, after which the GPU idles while the host queues the small
# kernels of the context code one by one): the same read-back also tells whether every label is foreground, which is what
# the decoder's teacher-forcing test (decoder_rnn.py:206-213) would otherwise read back after the object-context LSTM.
# With the early read the host queues the whole forward while the backbone runs. "0" keeps the reference's read-back points.
EARLY_HOST_INDS = os.environ.get("MOTIFS_EARLY_HOST_INDS", "1") == "1"


def _sort_by_score(im_inds, scores, host=None):
    """Permutation that keeps each image's objects together, longest image first, ordered by
    descending score inside an image; its inverse; per-timestep batch sizes (rel_model.py:31-61).
    The fp32 key `score - 2*(2*(s-e)*num_im + i)` is the reference's, bit for bit."""
    segs = image_segments(im_inds, host)
    num_im = (int(host[-1]) if host is not None else int(im_inds[-1])) + 1
    rois_per_image = np.zeros(num_im, dtype=np.float32)
    lengths = []
    for i, s, e in segs:
        rois_per_image[i] = 2 * (s - e) * num_im + i
        lengths.append(e - s)
    lengths = sorted(lengths, reverse=True)
    inds, ls_transposed = transpose_packed_sequence_inds(lengths)
    inds = to_device_async(inds, im_inds.device, torch.long)
    rpi = to_device_async(rois_per_image, im_inds.device)
    roi_order = scores - 2 * rpi[im_inds]
    # stable: objects whose keys tie exactly (PredCls ordered by confidence: every confidence is 1.0) keep their input
    # order, as in the oracle; an unstable sort leaves the order to the backend (CPU and CUDA differ)
    _, perm = torch.sort(roi_order, dim=0, descending=True, stable=True)
    perm = perm[inds]
    _, inv_perm = torch.sort(perm)
    return perm, inv_perm, ls_transposed


class LinearizedContext(nn.Module):
    """Object context + decoder + edge context (rel_model.py:66-296)."""

    def __init__(self, classes, rel_classes, mode='sgdet', embed_dim=200, hidden_dim=256, obj_dim=2048,
                 nl_obj=2, nl_edge=2, dropout_rate=0.2, order='confidence', pass_in_obj_feats_to_decoder=True,
                 pass_in_obj_feats_to_edge=True):
        super().__init__()
        self.classes = classes
        self.rel_classes = rel_classes
        assert mode in MODES
        self.mode = mode
        self.nl_obj = nl_obj
        self.nl_edge = nl_edge
        self.embed_dim = embed_dim
        self.hidden_dim = hidden_dim
        self.obj_dim = obj_dim
        self.dropout_rate = dropout_rate
        self.pass_in_obj_feats_to_decoder = pass_in_obj_feats_to_decoder
        self.pass_in_obj_feats_to_edge = pass_in_obj_feats_to_edge
        assert order in ('size', 'confidence', 'random', 'leftright')
        self.order = order

        embed_vecs = obj_edge_vectors(self.classes, wv_dim=self.embed_dim)
        self.obj_embed = nn.Embedding(self.num_classes, self.embed_dim)
        self.obj_embed.weight.data = embed_vecs.clone()
        self.obj_embed2 = nn.Embedding(self.num_classes, self.embed_dim)
        self.obj_embed2.weight.data = embed_vecs.clone()
        self.pos_embed = nn.Sequential(*[
            nn.BatchNorm1d(4, momentum=BATCHNORM_MOMENTUM / 10.0),
            nn.Linear(4, 128),
            nn.ReLU(inplace=True),
            nn.Dropout(0.1),
        ])
        if self.nl_obj > 0:
            self.obj_ctx_rnn = AlternatingHighwayLSTM(input_size=self.obj_dim + self.embed_dim + 128,
                                                      hidden_size=self.hidden_dim, num_layers=self.nl_obj,
                                                      recurrent_dropout_probability=dropout_rate)
            decoder_inputs_dim = self.hidden_dim
            if self.pass_in_obj_feats_to_decoder:
                decoder_inputs_dim += self.obj_dim + self.embed_dim
            self.decoder_rnn = DecoderRNN(self.classes, embed_dim=self.embed_dim, inputs_dim=decoder_inputs_dim,
                                          hidden_dim=self.hidden_dim, recurrent_dropout_probability=dropout_rate)
        else:
            self.decoder_lin = nn.Linear(self.obj_dim + self.embed_dim + 128, self.num_classes)
        if self.nl_edge > 0:
            input_dim = self.embed_dim
            if self.nl_obj > 0:
                input_dim += self.hidden_dim
            if self.pass_in_obj_feats_to_edge:
                input_dim += self.obj_dim
            self.edge_ctx_rnn = AlternatingHighwayLSTM(input_size=input_dim, hidden_size=self.hidden_dim,
                                                       num_layers=self.nl_edge,
                                                       recurrent_dropout_probability=dropout_rate)
        # injected randomness for parity runs: {"pos_embed.3", "obj_ctx_rnn", "decoder_rnn", "edge_ctx_rnn"}
        self.dropout_masks = None

    @property
    def num_classes(self):
        return len(self.classes)

    @property
    def num_rels(self):
        return len(self.rel_classes)

    def _mask(self, name):
        if self.dropout_masks is None or not self.training:
            return None
        return self.dropout_masks.get(name)

    def sort_rois(self, batch_idx, confidence, box_priors, host=None):
        """rel_model.py:139-161."""
        cxcywh = center_size(box_priors)
        if self.order == 'size':
            sizes = cxcywh[:, 2] * cxcywh[:, 3]
            scores = sizes / (sizes.max() + 1)
        elif self.order == 'confidence':
            scores = confidence
        elif self.order == 'random':
            scores = to_device_async(np.random.rand(batch_idx.size(0)), batch_idx.device, torch.float32)
        elif self.order == 'leftright':
            centers = cxcywh[:, 0]
            scores = centers / (centers.max() + 1)
        else:
            raise ValueError("invalid mode {}".format(self.order))
        return _sort_by_score(batch_idx, scores, host)

    def edge_ctx(self, obj_feats, obj_dists, im_inds, obj_preds, box_priors=None, im_inds_host=None):
        """rel_model.py:171-195."""
        obj_embed2 = self.obj_embed2(obj_preds)
        inp_feats = torch.cat((obj_embed2, obj_feats), 1)
        confidence = F.softmax(obj_dists, dim=1).detach().view(-1)[
            obj_preds.detach() + arange(obj_preds) * self.num_classes]
        perm, inv_perm, ls_transposed = self.sort_rois(im_inds.detach(), confidence, box_priors, im_inds_host)
        edge_input_packed = PackedSequence(inp_feats[perm], torch.as_tensor(ls_transposed))
        edge_reps = self.edge_ctx_rnn(edge_input_packed, dropout_weights=self._mask("edge_ctx_rnn"))[0][0]
        return edge_reps[inv_perm]

    def obj_ctx(self, obj_feats, obj_dists, im_inds, obj_labels=None, box_priors=None, boxes_per_cls=None,
                im_inds_host=None, labels_all_fg=None):
        """rel_model.py:197-234."""
        confidence = F.softmax(obj_dists, dim=1).detach()[:, 1:].max(1)[0]
        perm, inv_perm, ls_transposed = self.sort_rois(im_inds.detach(), confidence, box_priors, im_inds_host)
        obj_inp_rep = obj_feats[perm].contiguous()
        bs = torch.as_tensor(ls_transposed)
        encoder_rep = self.obj_ctx_rnn(PackedSequence(obj_inp_rep, bs), dropout_weights=self._mask("obj_ctx_rnn"))[0][0]
        if self.mode != 'predcls':
            decoder_inp = PackedSequence(torch.cat((obj_inp_rep, encoder_rep), 1)
                                         if self.pass_in_obj_feats_to_decoder else encoder_rep, bs)
            obj_dists, obj_preds = self.decoder_rnn(
                decoder_inp, labels=obj_labels[perm] if obj_labels is not None else None,
                boxes_for_nms=boxes_per_cls[perm] if boxes_per_cls is not None else None,
                dropout_mask=self._mask("decoder_rnn"), labels_all_fg=labels_all_fg)
            obj_preds = obj_preds[inv_perm]
            obj_dists = obj_dists[inv_perm]
        else:
            assert obj_labels is not None
            obj_preds = obj_labels
            obj_dists = to_onehot(obj_preds.detach(), self.num_classes)
        encoder_rep = encoder_rep[inv_perm]
        return obj_dists, obj_preds, encoder_rep

    def forward(self, obj_fmaps, obj_logits, im_inds, obj_labels=None, box_priors=None, boxes_per_cls=None,
                im_inds_host=None, labels_all_fg=None):
        """rel_model.py:236-296. `im_inds_host`: optional numpy copy of im_inds the caller already holds;
        `labels_all_fg`: whether every entry of `obj_labels` is > 0, when the caller already knows (else read back)."""
        obj_embed = tc_ops.matmul_tc(F.softmax(obj_logits, dim=1), self.obj_embed.weight, self.obj_embed.weight, "E")
        pe = self.pos_embed
        pos = pe[0](center_size(box_priors))
        pos = torch.relu(tc_ops.linear_tc(pos, pe[1].weight, pe[1].bias))
        m = self._mask("pos_embed.3")
        pos_embed = pos * m if (m is not None and self.training) else F.dropout(pos, pe[3].p, self.training)
        obj_pre_rep = torch.cat((obj_fmaps, obj_embed, pos_embed), 1)

        if self.nl_obj > 0:
            obj_dists2, obj_preds, obj_ctx = self.obj_ctx(obj_pre_rep, obj_logits, im_inds, obj_labels, box_priors,
                                                          boxes_per_cls, im_inds_host, labels_all_fg)
        else:
            if self.mode == 'predcls':
                obj_dists2 = to_onehot(obj_labels.detach(), self.num_classes)
            else:
                obj_dists2 = tc_ops.linear_tc(obj_pre_rep, self.decoder_lin.weight, self.decoder_lin.bias)
            if self.mode == 'sgdet' and not self.training:
                probs = F.softmax(obj_dists2, 1)
                nms_mask = torch.zeros_like(obj_dists2)
                for c_i in range(1, obj_dists2.size(1)):
                    scores_ci = probs.detach()[:, c_i]
                    boxes_ci = boxes_per_cls.detach()[:, c_i]
                    keep = apply_nms(scores_ci, boxes_ci, pre_nms_topn=scores_ci.size(0),
                                     post_nms_topn=scores_ci.size(0), nms_thresh=0.3)
                    nms_mask[:, c_i][keep] = 1
                obj_preds = (nms_mask * probs.detach())[:, 1:].max(1)[1] + 1
            else:
                obj_preds = obj_labels if obj_labels is not None else obj_dists2[:, 1:].max(1)[1] + 1
            obj_ctx = obj_pre_rep

        edge_ctx = None
        if self.nl_edge > 0:
            edge_ctx = self.edge_ctx(torch.cat((obj_fmaps, obj_ctx), 1) if self.pass_in_obj_feats_to_edge else obj_ctx,
                                     obj_dists=obj_dists2.detach(), im_inds=im_inds, obj_preds=obj_preds,
                                     box_priors=box_priors, im_inds_host=im_inds_host)
        return obj_dists2, obj_preds, edge_ctx


class RelModel(nn.Module):
    def __init__(self, classes, rel_classes, mode='sgdet', num_gpus=1, use_vision=True, require_overlap_det=True,
                 embed_dim=200, hidden_dim=256, pooling_dim=2048, nl_obj=1, nl_edge=2, use_resnet=False,
                 order='confidence', thresh=0.01, use_proposals=False, pass_in_obj_feats_to_decoder=True,
                 pass_in_obj_feats_to_edge=True, rec_dropout=0.0, use_bias=True, use_tanh=True, limit_vision=True):
        super().__init__()
        self.classes = classes
        self.rel_classes = rel_classes
        self.num_gpus = num_gpus
        assert mode in MODES
        self.mode = mode
        self.pooling_size = 7
        self.embed_dim = embed_dim
        self.hidden_dim = hidden_dim
        self.obj_dim = 2048 if use_resnet else 4096
        self.pooling_dim = pooling_dim
        self.use_bias = use_bias
        self.use_vision = use_vision
        self.use_tanh = use_tanh
        self.limit_vision = limit_vision
        self.require_overlap = require_overlap_det and self.mode == 'sgdet'

        self.detector = ObjectDetector(
            classes=classes,
            mode=('proposals' if use_proposals else 'refinerels') if mode == 'sgdet' else 'gtbox',
            use_resnet=use_resnet, thresh=thresh, max_per_img=64)
        self.context = LinearizedContext(self.classes, self.rel_classes, mode=self.mode, embed_dim=self.embed_dim,
                                         hidden_dim=self.hidden_dim, obj_dim=self.obj_dim, nl_obj=nl_obj,
                                         nl_edge=nl_edge, dropout_rate=rec_dropout, order=order,
                                         pass_in_obj_feats_to_decoder=pass_in_obj_feats_to_decoder,
                                         pass_in_obj_feats_to_edge=pass_in_obj_feats_to_edge)
        self.union_boxes = UnionBoxesAndFeats(pooling_size=self.pooling_size, stride=16,
                                              dim=1024 if use_resnet else 512)
        roi_fmap = [Flattener(),
                    load_vgg(use_dropout=False, use_relu=False, use_linear=pooling_dim == 4096, pretrained=False).classifier]
        if pooling_dim != 4096:
            roi_fmap.append(nn.Linear(4096, pooling_dim))
        self.roi_fmap = nn.Sequential(*roi_fmap)
        self.roi_fmap_obj = load_vgg(pretrained=False).classifier

        self.post_lstm = nn.Linear(self.hidden_dim, self.pooling_dim * 2)
        self.post_lstm.weight.data.normal_(0, 10.0 * math.sqrt(1.0 / self.hidden_dim))   # rel_model.py:383
        self.post_lstm.bias.data.zero_()
        if nl_edge == 0:
            self.post_emb = nn.Embedding(self.num_classes, self.pooling_dim * 2)
            self.post_emb.weight.data.normal_(0, math.sqrt(1.0))
        self.rel_compress = nn.Linear(self.pooling_dim, self.num_rels, bias=True)
        torch.nn.init.xavier_normal_(self.rel_compress.weight, gain=1.0)
        if self.use_bias:
            self.freq_bias = FrequencyBias(num_objs=self.num_classes, num_rels=self.num_rels)
        self.dropout_masks = None   # {"roi_fmap_obj.2", "roi_fmap_obj.5", "roi_fmap.1.2"} for parity runs
        tc_ops.install_load_hook(self)   # load_state_dict copies in place: cached bf16 splits are stale afterwards

    @property
    def num_classes(self):
        return len(self.classes)

    @property
    def num_rels(self):
        return len(self.rel_classes)

    def _run_roi_fmap(self, x):
        for m in self.roi_fmap:
            if isinstance(m, nn.Sequential):
                x = run_classifier(m, x, self.dropout_masks, "roi_fmap.1.")
            elif isinstance(m, nn.Linear):
                x = tc_ops.linear_tc(x, m.weight, m.bias)
            else:
                x = m(x)
        return x

    def _nhwc_of(self, features):
        """The detector's NHWC copy of `features` iff `features` IS the map it produced last (same memory) and carries no
        gradient; anything else (another map, a micro-batch, a trainable backbone) takes the NCHW autograd path."""
        nh = self.detector._fmap_nhwc
        if nh is None or features.requires_grad or features.data_ptr() != nh.data_ptr() or \
                tuple(features.shape) != (nh.size(0), nh.size(3), nh.size(1), nh.size(2)):
            return None
        return nh

    def visual_rep(self, features, rois, pair_inds):
        """Union-box visual features -> fc6/fc7 (no final ReLU) (rel_model.py:403-414)."""
        assert pair_inds.size(1) == 2
        uboxes = self.union_boxes(features, rois, pair_inds, fmap_nhwc=self._nhwc_of(features))
        return self._run_roi_fmap(uboxes)

    def get_rel_inds(self, rel_labels, im_inds, box_priors):
        """rel_model.py:416-437."""
        if self.training:
            return rel_labels[:, :3].detach().clone()
        rel_cands = im_inds.detach()[:, None] == im_inds.detach()[None]
        rel_cands.fill_diagonal_(False)
        if self.require_overlap:
            rel_cands = rel_cands & (bbox_overlaps(box_priors.detach(), box_priors.detach()) > 0)
        rel_cands = rel_cands.nonzero()
        if rel_cands.numel() == 0:
            rel_cands = im_inds.new_zeros(1, 2)
        return torch.cat((im_inds.detach()[rel_cands[:, 0]][:, None], rel_cands), 1)

    def obj_feature_map(self, features, rois):
        """RoIAlign + the trainable fc6/fc7 copy (rel_model.py:439-448)."""
        nh = self._nhwc_of(features)
        if nh is not None:
            pool = roi_align_from_nhwc(nh, rois, self.pooling_size, self.pooling_size, 1 / 16)
        else:
            pool = RoIAlignFunction(self.pooling_size, self.pooling_size, spatial_scale=1 / 16)(features, rois)
        return run_classifier(self.roi_fmap_obj, pool.view(rois.size(0), -1), self.dropout_masks, "roi_fmap_obj.")

    def forward(self, x, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None, proposals=None,
                train_anchor_inds=None, return_fmap=False):
        im_inds_host = labels_all_fg = None
        if EARLY_HOST_INDS and self.detector.mode == 'gtbox' and gt_classes is not None:
            gtc = getattr(gt_classes, "_mb200_host", None)                      # the loader's host copy (dataloaders/synthetic.py) ...
            if gtc is None or gtc.shape != tuple(gt_classes.shape):
                gtc = gt_classes.cpu().numpy()                                  # ... else the one early read-back
            im_inds_host = gtc[:, 0] - image_offset
            labels_all_fg = bool((gtc[:, 1] > 0).all())                         # rm_obj_labels = gt_classes[:, 1] in gtbox mode
        # A deferred optimizer update (lib/fused_optim.FlatSGD(defer_step=True): gradient all-reduce + fused SGD on a side
        # stream) may still be in flight. A frozen detector (models/train_rels.py:51-52) reads none of the parameters being
        # updated, so it runs underneath; everything after it waits here.
        det_frozen = not any(p.requires_grad for p in self.detector.parameters())
        if not det_frozen:
            fused_optim.wait_pending_updates()
        result = self.detector(x, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, proposals,
                               train_anchor_inds, return_fmap=True, im_inds_host=im_inds_host)
        fused_optim.wait_pending_updates()
        if result.is_none():
            return ValueError("heck")   # rel_model.py:474-475 returns (does not raise) this

        im_inds = result.im_inds - image_offset
        boxes = result.rm_box_priors
        if self.training and result.rel_labels is None:
            assert self.mode == 'sgdet'
            from lib.fpn.proposal_assignments.rel_assignments import rel_assignments
            result.rel_labels = rel_assignments(im_inds.detach(), boxes.detach(), result.rm_obj_labels.detach(),
                                                gt_boxes.detach(), gt_classes.detach(), gt_rels.detach(), image_offset,
                                                filter_non_overlap=True, num_sample_per_gt=1,
                                                rng=getattr(self.detector, "rng", np.random))
        rel_inds = self.get_rel_inds(result.rel_labels, im_inds, boxes)
        rois = torch.cat((im_inds[:, None].float(), boxes), 1)
        result.obj_fmap = self.obj_feature_map(result.fmap.detach(), rois)

        result.rm_obj_dists, result.obj_preds, edge_ctx = self.context(
            result.obj_fmap, result.rm_obj_dists.detach(), im_inds,
            result.rm_obj_labels if self.training or self.mode == 'predcls' else None,
            boxes.detach(), result.boxes_all, im_inds_host=im_inds_host, labels_all_fg=labels_all_fg)

        if edge_ctx is None:
            edge_rep = self.post_emb(result.obj_preds)
        else:
            edge_rep = tc_ops.linear_tc(edge_ctx, self.post_lstm.weight, self.post_lstm.bias)
        edge_rep = edge_rep.view(edge_rep.size(0), 2, self.pooling_dim)
        subj_rep = edge_rep[:, 0]
        obj_rep = edge_rep[:, 1]
        prod_rep = subj_rep[rel_inds[:, 1]] * obj_rep[rel_inds[:, 2]]

        if self.use_vision:
            vr = self.visual_rep(result.fmap.detach(), rois, rel_inds[:, 1:])
            if self.limit_vision:
                prod_rep = torch.cat((prod_rep[:, :2048] * vr[:, :2048], prod_rep[:, 2048:]), 1)
            else:
                prod_rep = prod_rep * vr
        if self.use_tanh:
            prod_rep = torch.tanh(prod_rep)
        result.rel_dists = tc_ops.linear_tc(prod_rep, self.rel_compress.weight, self.rel_compress.bias)
        if self.use_bias:
            result.rel_dists = result.rel_dists + self.freq_bias.index_with_labels(torch.stack((
                result.obj_preds[rel_inds[:, 1]], result.obj_preds[rel_inds[:, 2]]), 1))
        if getattr(self, "keep_last_result", False):
            self.last_result = result          # debugging / parity tests: logits before filter_dets
        self.detector._fmap_nhwc = self.detector._fmap_split = None     # do not pin this batch's feature map until the next one
        if self.training:
            return result

        twod_inds = arange(result.obj_preds) * self.num_classes + result.obj_preds.detach()
        result.obj_scores = F.softmax(result.rm_obj_dists, dim=1).view(-1)[twod_inds]
        if self.mode == 'sgdet':
            bboxes = result.boxes_all.view(-1, 4)[twod_inds].view(result.boxes_all.size(0), 4)
        else:
            bboxes = result.rm_box_priors
        rel_rep = F.softmax(result.rel_dists, dim=1)
        return filter_dets(bboxes, result.obj_scores, result.obj_preds, rel_inds[:, 1:], rel_rep)

    def __getitem__(self, batch):
        """`detector[blob]` as models/train_rels.py:137 calls it. One process per GPU (torch.distributed,
        NCCL): no per-step replicate/broadcast as rel_model.py:549-560; this rank runs its own shard."""
        batch.scatter()
        return self(*batch[0])
