"""Faster-R-CNN detector of the MotifNet hot path — same surface as the reference's
lib/object_detector.py: `ObjectDetector(classes, mode, num_gpus, nms_filter_duplicates, max_per_img,
use_resnet, thresh).forward(x, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, proposals,
train_anchor_inds, return_fmap)` -> `Result` (:22-37, :274-361); module / parameter names match so
state dicts interchange (features.N, roi_fmap.{0,3}, score_fc, bbox_fc, rpn_head.conv.{0,2}).

What runs where (B200 path): VGG16 conv1_1..conv5_3 on the tcgen05 implicit-GEMM convolution
(lib/tc_ops.py, csrc/gemm_tc.cu), feature map kept NHWC; RoIAlign on the channel-vectorised NHWC
kernel; fc6/fc7/score_fc/bbox_fc and the RPN 1x1 conv on the tcgen05 GEMM; box decode fused
(csrc/boxes.cu); RPN proposal NMS and the per-class detection NMS each as ONE segmented
on-device launch pair (csrc/nms.cu) instead of <=150 host round trips per image (:445-452)."""
import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F
from torchvision.models.vgg import vgg16

import motifs_cabi as _c
from config import ANCHOR_SIZE, ANCHOR_RATIOS, ANCHOR_SCALES
from lib import tc_ops
from lib.fpn.box_utils import bbox_preds_fused, center_size, bbox_overlaps
from lib.fpn.generate_anchors import generate_anchors
from lib.fpn.nms.functions.nms import apply_nms, nms_segments
from lib.fpn.proposal_assignments.proposal_assignments_gtbox import proposal_assignments_gtbox
from lib.fpn.proposal_assignments.proposal_assignments_det import proposal_assignments_det
from lib.fpn.roi_align.functions.roi_align import RoIAlignFunction, roi_align_from_nhwc
from lib.pytorch_misc import image_segments, gather_nd, to_device_async


class Result(object):
    """Container for detector / relation-model outputs (object_detector.py:22-37); od = object
    detector, rm = relation model."""

    FIELDS = ('od_obj_dists', 'rm_obj_dists', 'obj_scores', 'obj_preds', 'obj_fmap', 'od_box_deltas',
              'rm_box_deltas', 'od_box_targets', 'rm_box_targets', 'od_box_priors', 'rm_box_priors',
              'boxes_assigned', 'boxes_all', 'od_obj_labels', 'rm_obj_labels', 'rpn_scores', 'rpn_box_deltas',
              'rel_labels', 'im_inds', 'fmap', 'rel_dists', 'rel_inds', 'rel_rep')

    def __init__(self, **kwargs):
        for f in self.FIELDS:
            setattr(self, f, kwargs.pop(f, None))
        if kwargs:
            raise TypeError("unknown Result fields: %s" % sorted(kwargs))

    def is_none(self):
        return all(v is None for v in self.__dict__.values())


def gather_res(outputs, target_device, dim=0):
    """Concatenate the non-None fields of several Results (object_detector.py:40-47). With one
    process per GPU this only merges results of the same rank (e.g. micro-batches)."""
    out = outputs[0]
    args = {f: torch.cat([getattr(o, f).to(target_device) for o in outputs], dim)
            for f, v in out.__dict__.items() if v is not None}
    return type(out)(**args)


def load_vgg(use_dropout=True, use_relu=True, use_linear=True, pretrained=False):
    """VGG16 minus the last max-pool and the class layer (object_detector.py:623-633). Weights come
    from the state dict; there is no network here, so `pretrained` must stay False."""
    if pretrained:
        raise ValueError("pretrained ImageNet weights are not available offline; load a state dict instead")
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


def load_resnet(pretrained=False):
    """torchvision resnet101 minus layer4 / avgpool / fc (object_detector.py:615-620); weights come from the state
    dict (no network here)."""
    if pretrained:
        raise ValueError("pretrained ImageNet weights are not available offline; load a state dict instead")
    from torchvision.models.resnet import resnet101
    model = resnet101(weights=None)
    del model.layer4
    del model.avgpool
    del model.fc
    return model


def run_classifier(classifier, x, dropout_masks=None, prefix=""):
    """Apply a (possibly trimmed) VGG classifier Sequential — Linear / ReLU / Dropout — with the
    Linear layers on the tcgen05 GEMM. `dropout_masks` ({name: mask}) injects masks for parity runs."""
    for name, m in classifier._modules.items():
        if isinstance(m, nn.Linear):
            x = tc_ops.linear_tc(x, m.weight, m.bias)
        elif isinstance(m, nn.ReLU):
            x = torch.relu(x)
        elif isinstance(m, nn.Dropout):
            key = prefix + name
            if classifier.training and dropout_masks is not None and key in dropout_masks:
                x = x * dropout_masks[key]
            else:
                x = F.dropout(x, m.p, classifier.training)
        else:
            x = m(x)
    return x


class ObjectDetector(nn.Module):
    MODES = ('rpntrain', 'gtbox', 'refinerels', 'proposals')

    def __init__(self, classes, mode='rpntrain', num_gpus=1, nms_filter_duplicates=True,
                 max_per_img=64, use_resnet=False, thresh=0.05):
        super().__init__()
        if mode not in self.MODES:
            raise ValueError("invalid mode")
        self.mode = mode
        self.classes = classes
        self.num_gpus = num_gpus
        self.pooling_size = 7
        self.nms_filter_duplicates = nms_filter_duplicates
        self.max_per_img = max_per_img
        self.use_resnet = use_resnet
        self.thresh = thresh
        if not use_resnet:
            vgg_model = load_vgg()
            self.features = vgg_model.features
            self.roi_fmap = vgg_model.classifier
            rpn_input_dim, output_dim = 512, 4096
        else:                           # object_detector.py:84-101 (same module tree -> same state-dict keys)
            self.features = load_resnet()
            self.compress = nn.Sequential(nn.Conv2d(1024, 256, kernel_size=1), nn.ReLU(inplace=True),
                                          nn.BatchNorm2d(256))
            self.roi_fmap = nn.Sequential(nn.Linear(256 * 7 * 7, 2048), nn.SELU(inplace=True), nn.AlphaDropout(p=0.05),
                                          nn.Linear(2048, 2048), nn.SELU(inplace=True), nn.AlphaDropout(p=0.05))
            rpn_input_dim, output_dim = 1024, 2048
        self.score_fc = nn.Linear(output_dim, self.num_classes)
        self.bbox_fc = nn.Linear(output_dim, self.num_classes * 4)
        self.rpn_head = RPNHead(dim=512, input_dim=rpn_input_dim)
        self.dropout_masks = None     # {"roi_fmap.2": mask, "roi_fmap.5": mask} for parity runs
        self._fmap_nhwc = None
        self._fmap_split = None
        tc_ops.install_load_hook(self)

    @property
    def num_classes(self):
        return len(self.classes)

    def _convs(self):
        return [m for m in self.features if isinstance(m, nn.Conv2d)]

    def feature_map(self, x):
        """[B,3,S,S] -> stride-16 map [B,512,S/16,S/16] (object_detector.py:110-127). The map lives in
        NHWC; the returned tensor is its NCHW view (same memory)."""
        if any(p.requires_grad for p in self.features.parameters()) and torch.is_grad_enabled():
            if self.use_resnet:
                raise NotImplementedError("training the ResNet-101 backbone: only the VGG gradient path exists "
                                          "(lib/conv_tc.py); freeze `features` (models/train_rels.py:51-52 does)")
            from lib import conv_tc
            nhwc = conv_tc.vgg_features_train(x.contiguous().float(), self._convs(), tc_ops.VGG16_CFG)
            self._fmap_nhwc, self._fmap_split = None, None       # consumers take the autograd paths
            return nhwc.permute(0, 3, 1, 2)
        with torch.no_grad():
            need_split = self.mode in ('rpntrain', 'refinerels')
            if self.use_resnet:         # conv1 .. layer3 (:119-127) walked on the kernels, NHWC [B,37,37,1024]
                from lib.resnet_tc import resnet_c4_forward, KernelOps
                nhwc = resnet_c4_forward(self.features, x.contiguous().float(), KernelOps()).contiguous()
                split = None
                if need_split:
                    B, H, W, C = nhwc.shape
                    sp = tc_ops.split_rows(nhwc.view(-1, C))
                    split = (sp.hi.view(B, H, W, C), sp.lo.view(B, H, W, C))
            else:
                nhwc, split = tc_ops.vgg_features_forward(x, self._convs(), want_last_split=need_split)
        self._fmap_nhwc, self._fmap_split = nhwc, split
        return nhwc.permute(0, 3, 1, 2)

    def obj_feature_map(self, features, rois):
        """RoIAlign 7x7 + fc6/fc7 (object_detector.py:129-138)."""
        if self.use_resnet:             # :136-137: RoIAlign over compress(features) = 1x1 conv 1024->256 + ReLU + BN
            c, bn = self.compress[0], self.compress[2]
            B, C, H, W = features.shape
            rows = features.permute(0, 2, 3, 1).reshape(-1, C)
            y = torch.relu(tc_ops.linear_tc(rows, c.weight.view(c.out_channels, -1), c.bias))
            y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, bn.momentum, bn.eps)
            if not y.requires_grad:
                pool = roi_align_from_nhwc(y.view(B, H, W, -1), rois, self.pooling_size, self.pooling_size, 1 / 16)
            else:
                pool = RoIAlignFunction(self.pooling_size, self.pooling_size, spatial_scale=1 / 16)(
                    y.view(B, H, W, -1).permute(0, 3, 1, 2), rois)
            return run_classifier(self.roi_fmap, pool.reshape(rois.size(0), -1), self.dropout_masks, "roi_fmap.")
        if self._fmap_nhwc is not None and features.data_ptr() == self._fmap_nhwc.data_ptr() and not features.requires_grad:
            pool = roi_align_from_nhwc(self._fmap_nhwc, rois, self.pooling_size, self.pooling_size, 1 / 16)
        else:
            pool = RoIAlignFunction(self.pooling_size, self.pooling_size, spatial_scale=1 / 16)(features, rois)
        return run_classifier(self.roi_fmap, pool.view(rois.size(0), -1), self.dropout_masks, "roi_fmap.")

    # ------------------------------------------------------------------ box sources
    def rpn_boxes(self, fmap, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None,
                  train_anchor_inds=None, proposals=None):
        rpn_feats = self.rpn_head(fmap, fmap_split=self._fmap_split)
        rois = self.rpn_head.roi_proposals(
            rpn_feats, im_sizes, nms_thresh=0.7,
            pre_nms_topn=12000 if self.training and self.mode == 'rpntrain' else 6000,
            post_nms_topn=2000 if self.training and self.mode == 'rpntrain' else 1000)
        if self.training:
            if gt_boxes is None or gt_classes is None or train_anchor_inds is None:
                raise ValueError("Must supply GT boxes, GT classes, trainanchors when in train mode")
            rpn_scores, rpn_box_deltas = self.rpn_head.anchor_preds(rpn_feats, train_anchor_inds, image_offset)
            if gt_rels is not None and self.mode == 'rpntrain':
                raise ValueError("Training the object detector and the relationship model with detection"
                                 "at the same time isn't supported")
            if self.mode == 'refinerels':
                return rois, None, None, rpn_scores, rpn_box_deltas, None
            all_rois, labels, bbox_targets = proposal_assignments_det(
                rois, gt_boxes.detach(), gt_classes.detach(), image_offset, fg_thresh=0.5, rng=getattr(self, "rng", np.random))
            return all_rois, labels, bbox_targets, rpn_scores, rpn_box_deltas, None
        return rois, None, None, None, None, None

    def gt_boxes(self, fmap, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None,
                 train_anchor_inds=None, proposals=None, im_inds_host=None):
        assert gt_boxes is not None
        im_inds = gt_classes[:, 0] - image_offset
        rois = torch.cat((im_inds.float()[:, None], gt_boxes), 1)
        if gt_rels is not None and self.training:
            rois, labels, rel_labels = proposal_assignments_gtbox(
                rois.detach(), gt_boxes.detach(), gt_classes.detach(), gt_rels.detach(), image_offset, fg_thresh=0.5,
                rng=getattr(self, "rng", np.random),
                num_im=(int(im_inds_host[-1]) + 1) if im_inds_host is not None and len(im_inds_host) else None)
        else:
            labels = gt_classes[:, 1]
            rel_labels = None
        return rois, labels, None, None, None, rel_labels

    def proposal_boxes(self, fmap, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None,
                       train_anchor_inds=None, proposals=None):
        assert proposals is not None
        rois = filter_roi_proposals(proposals[:, 2:].detach().contiguous(), proposals[:, 1].detach().contiguous(),
                                    np.array([2000] * len(im_sizes)), nms_thresh=0.7,
                                    pre_nms_topn=12000 if self.training and self.mode == 'rpntrain' else 6000,
                                    post_nms_topn=2000 if self.training and self.mode == 'rpntrain' else 1000)
        if self.training:
            all_rois, labels, bbox_targets = proposal_assignments_det(
                rois, gt_boxes.detach(), gt_classes.detach(), image_offset, fg_thresh=0.5, rng=getattr(self, "rng", np.random))
            all_rois = torch.cat((all_rois, rois), 0)       # object_detector.py:254-255
            return all_rois, labels, bbox_targets, None, None, None
        return rois, None, None, None, None, None

    def get_boxes(self, *args, **kwargs):
        if self.mode == 'gtbox':
            fn = self.gt_boxes
        elif self.mode == 'proposals':
            assert kwargs['proposals'] is not None
            fn = self.proposal_boxes
        else:
            fn = self.rpn_boxes
        return fn(*args, **kwargs)

    def forward(self, x, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None, proposals=None,
                train_anchor_inds=None, return_fmap=False, im_inds_host=None):
        """`im_inds_host` (superset of the reference signature): host copy of the GT image indices the caller has
        already read back; with it the GT-box assignment — whose `nonzero` needs the host anyway — runs BEFORE the
        backbone is queued, so the host never waits on the backbone."""
        frozen = not any(p.requires_grad for p in self.parameters())
        if not frozen:          # a deferred optimizer update of these parameters may be in flight (lib/fused_optim.py)
            from lib import fused_optim
            fused_optim.wait_pending_updates()
        with torch.set_grad_enabled(torch.is_grad_enabled() and not frozen):
            return self._forward(x, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, proposals,
                                 train_anchor_inds, return_fmap, im_inds_host)

    def _forward(self, x, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, proposals, train_anchor_inds,
                 return_fmap, im_inds_host=None):
        # GT-box mode: the relation sampling (proposal_assignments_gtbox) depends on the inputs only, and it reads candidate
        # counts back to the host (`nonzero`). Queue the backbone FIRST and run the assignment on a side stream: its host
        # waits then cover a few tiny kernels instead of the 4.5 ms backbone, and the GPU is busy meanwhile (round 2 trace:
        # 0.49 ms of idle at the start of every step when the assignment ran first on the compute stream).
        side = self.mode == 'gtbox' and x.is_cuda and gt_boxes is not None
        if side:
            main = torch.cuda.current_stream(x.device)
            inputs_ready = torch.cuda.Event()
            inputs_ready.record(main)
        fmap = self.feature_map(x)
        if side:
            if getattr(self, "_assign_stream", None) is None or self._assign_stream.device != x.device:
                self._assign_stream = torch.cuda.Stream(x.device)
            st = self._assign_stream
            st.wait_event(inputs_ready)
            with torch.cuda.stream(st):
                got = self.gt_boxes(None, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, train_anchor_inds,
                                    proposals=proposals, im_inds_host=im_inds_host)
            main.wait_stream(st)
            for t in got:
                if torch.is_tensor(t):
                    t.record_stream(main)
        else:
            got = self.get_boxes(fmap, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, train_anchor_inds,
                                 proposals=proposals)
        rois, obj_labels, bbox_targets, rpn_scores, rpn_box_deltas, rel_labels = got
        obj_fmap = self.obj_feature_map(fmap, rois)
        od_obj_dists = tc_ops.linear_tc(obj_fmap, self.score_fc.weight, self.score_fc.bias)
        od_box_deltas = tc_ops.linear_tc(obj_fmap, self.bbox_fc.weight, self.bbox_fc.bias).view(
            -1, len(self.classes), 4) if self.mode != 'gtbox' else None
        od_box_priors = rois[:, 1:]

        if (not self.training and not self.mode == 'gtbox') or self.mode in ('proposals', 'refinerels'):
            nms = self.nms_boxes(od_obj_dists, rois, od_box_deltas, im_sizes)
            if nms is None:
                return Result()
            nms_inds, nms_scores, nms_preds, nms_boxes_assign, nms_boxes, nms_imgs = nms
            im_inds = nms_imgs + image_offset
            obj_dists = od_obj_dists[nms_inds]
            obj_fmap = obj_fmap[nms_inds]
            box_deltas = od_box_deltas[nms_inds]
            box_priors = nms_boxes[:, 0]
            if self.training and not self.mode == 'gtbox':
                # label the surviving detections by IoU >= 0.5 with a GT box of the same image (:319-326)
                pred_to_gtbox = bbox_overlaps(box_priors, gt_boxes)
                pred_to_gtbox[im_inds[:, None] != gt_classes[None, :, 0]] = 0.0
                max_overlaps, argmax_overlaps = pred_to_gtbox.max(1)
                rm_obj_labels = gt_classes[:, 1][argmax_overlaps].clone()
                rm_obj_labels[max_overlaps < 0.5] = 0
            else:
                rm_obj_labels = None
        else:
            im_inds = rois[:, 0].long().contiguous() + image_offset
            nms_scores = nms_preds = nms_boxes_assign = nms_boxes = None
            box_priors = rois[:, 1:]
            rm_obj_labels = obj_labels
            box_deltas = od_box_deltas
            obj_dists = od_obj_dists

        return Result(
            od_obj_dists=od_obj_dists, rm_obj_dists=obj_dists, obj_scores=nms_scores, obj_preds=nms_preds,
            obj_fmap=obj_fmap, od_box_deltas=od_box_deltas, rm_box_deltas=box_deltas,
            od_box_targets=bbox_targets, rm_box_targets=bbox_targets, od_box_priors=od_box_priors,
            rm_box_priors=box_priors, boxes_assigned=nms_boxes_assign, boxes_all=nms_boxes,
            od_obj_labels=obj_labels, rm_obj_labels=rm_obj_labels, rpn_scores=rpn_scores,
            rpn_box_deltas=rpn_box_deltas, rel_labels=rel_labels, im_inds=im_inds,
            fmap=fmap if return_fmap else None)

    def nms_boxes(self, obj_dists, rois, box_deltas, im_sizes):
        """Decode all class boxes, clamp, per-class NMS, top max_per_img per image (:363-408)."""
        N, K = box_deltas.size(0), box_deltas.size(1)
        inds = rois[:, 0].long().contiguous()
        dev = rois.device
        im_hw = to_device_async(np.ascontiguousarray(np.asarray(im_sizes)[:, :2].astype(np.float32)), dev)
        boxes = bbox_preds_fused(rois[:, 1:].contiguous(), box_deltas.reshape(-1, 4), K, im_hw,
                                 inds.to(torch.int32)).view(N, K, 4)
        probs = F.softmax(obj_dists, 1)
        dets = []
        for i, s, e in image_segments(inds):
            d = filter_det(probs[s:e], boxes[s:e], start_ind=s, nms_filter_duplicates=self.nms_filter_duplicates,
                           max_per_img=self.max_per_img, thresh=self.thresh)
            if d is not None:
                dets.append(d)
        if len(dets) == 0:
            print("nothing was detected", flush=True)
            return None
        nms_inds, nms_scores, nms_labels = [torch.cat(x, 0) for x in zip(*dets)]
        twod_inds = nms_inds * boxes.size(1) + nms_labels
        nms_boxes_assign = boxes.view(-1, 4)[twod_inds]
        nms_boxes = torch.cat((rois[:, 1:][nms_inds][:, None], boxes[nms_inds][:, 1:]), 1)
        return nms_inds, nms_scores, nms_labels, nms_boxes_assign, nms_boxes, inds[nms_inds]

    def __getitem__(self, batch):
        """`detector[blob]` (object_detector.py:410-423): one process per GPU, so no replicate /
        parallel_apply — the batch's tuple for this rank is run directly."""
        batch.scatter()
        return self(*batch[0])


def filter_det(scores, boxes, start_ind=0, max_per_img=100, thresh=0.001, pre_nms_topn=6000,
               post_nms_topn=300, nms_thresh=0.3, nms_filter_duplicates=True):
    """Detections of one image (object_detector.py:425-485). scores [N,C] probabilities, boxes
    [N,C,4] clamped. All classes go through ONE segmented NMS launch; classes whose best score is
    below `thresh` are masked afterwards, which is equivalent to skipping them (:439,:445)."""
    scores = scores.detach()
    boxes = boxes.detach()
    N, C = scores.shape
    dev = scores.device
    valid = scores[:, 1:].max(0)[0] > thresh                         # [C-1]
    vs, order = torch.sort(scores[:, 1:], dim=0, descending=True)    # per class
    n_use = min(N, pre_nms_topn)
    order = order[:n_use]                                            # [n_use, C-1]
    cls = torch.arange(1, C, device=dev)
    boxes_sorted = boxes[order, cls[None, :]].permute(1, 0, 2).contiguous().view(-1, 4)   # class-major segments
    keep, num_keep, _ = nms_segments(boxes_sorted, [n_use] * (C - 1), nms_thresh)
    keep = keep.view(C - 1, n_use).long()
    rank = torch.arange(n_use, device=dev)[None, :]
    limit = torch.clamp(num_keep.long(), max=post_nms_topn)[:, None]
    sel = (rank < limit) & valid[:, None]                            # kept entries of valid classes
    keep_c = keep.clamp_(0, n_use - 1)
    rows = order.t().gather(1, keep_c)                               # original roi index of each kept entry
    nms_mask = torch.zeros(N, C, device=dev, dtype=scores.dtype)
    cc = cls[:, None].expand_as(rows)
    nms_mask[rows[sel], cc[sel]] = 1
    if not bool(valid.any()):
        return None
    dists_all = nms_mask * scores
    if nms_filter_duplicates:
        scores_pre, labels_pre = dists_all.max(1)
        inds_all = scores_pre.nonzero().squeeze(1)
        labels_all = labels_pre[inds_all]
        scores_all = scores_pre[inds_all]
    else:
        nz = nms_mask.nonzero()
        inds_all, labels_all = nz[:, 0], nz[:, 1]
        scores_all = scores.reshape(-1)[inds_all * C + labels_all]
    vs, idx = torch.sort(scores_all, dim=0, descending=True)
    idx = idx[vs > thresh]
    if max_per_img < idx.size(0):
        idx = idx[:max_per_img]
    return inds_all[idx] + start_ind, scores_all[idx], labels_all[idx]


class RPNHead(nn.Module):
    """Class / box outputs over the 37x37x20 anchor grid (object_detector.py:488-597)."""

    def __init__(self, dim=512, input_dim=1024):
        super().__init__()
        self.anchor_target_dim = 6
        self.stride = 16
        self.conv = nn.Sequential(
            nn.Conv2d(input_dim, dim, kernel_size=3, padding=1),
            nn.ReLU6(inplace=True),
            nn.Conv2d(dim, self.anchor_target_dim * self._A, kernel_size=1))
        ans_np = generate_anchors(base_size=ANCHOR_SIZE, feat_stride=self.stride, anchor_scales=ANCHOR_SCALES,
                                  anchor_ratios=ANCHOR_RATIOS)
        self.register_buffer('anchors', torch.FloatTensor(ans_np))

    @property
    def _A(self):
        return len(ANCHOR_RATIOS) * len(ANCHOR_SCALES)

    def forward(self, fmap, fmap_split=None):
        """[B,C,h,w] -> [B,h,w,A,6]. With the backbone's NHWC bf16 pair at hand the 3x3 conv runs on
        the tcgen05 implicit GEMM and the 1x1 conv is a plain GEMM whose output is already NHWC
        (the reference transposes NCHW->NHWC here, :551-558)."""
        B, C, h, w = fmap.shape
        c0 = self.conv[0]
        if torch.is_grad_enabled() and (fmap.requires_grad or c0.weight.requires_grad):
            # training the head (models/train_detector.py): the forward-only kernel call below would silently drop
            # the gradients of conv[0] and of the feature map
            from lib import conv_tc
            y = conv_tc.conv3x3(fmap.permute(0, 2, 3, 1).contiguous(), c0.weight, c0.bias, relu=True).clamp(max=6.0)
            c1 = self.conv[2]
            rez = tc_ops.linear_tc(y.view(B * h * w, -1), c1.weight.view(c1.weight.size(0), -1), c1.bias)
            return rez.view(B, h, w, self._A, self.anchor_target_dim)
        if fmap_split is None:
            xs = tc_ops.split_rows(fmap.detach().permute(0, 2, 3, 1).reshape(-1, C))
            fmap_split = (xs.hi.view(B, h, w, C), xs.lo.view(B, h, w, C))
        y, _ = tc_ops.conv3x3_relu(fmap_split, B, h, w, C, self.conv[0], want_f32=True, want_split=False)
        y = y.clamp_(max=6.0)                                        # ReLU6
        c1 = self.conv[2]
        rez = tc_ops.linear_tc(y.view(B * h * w, -1), c1.weight.view(c1.weight.size(0), -1), c1.bias)
        return rez.view(B, h, w, self._A, self.anchor_target_dim)

    def anchor_preds(self, preds, train_anchor_inds, image_offset):
        assert train_anchor_inds.size(1) == 4
        tai = train_anchor_inds.detach().clone()
        tai[:, 0] -= image_offset
        train_regions = gather_nd(preds, tai)
        return train_regions[:, :2], train_regions[:, 2:]

    def roi_proposals(self, fmap, im_sizes, nms_thresh=0.7, pre_nms_topn=12000, post_nms_topn=2000):
        """:560-597 — fg probability, delta decode vs anchors (fused, clamped per image), mask the
        padded region and <4 px boxes with score -0.01, NMS per image."""
        fmap = fmap.detach()
        B = fmap.size(0)
        dev = fmap.device
        class_preds = F.softmax(fmap[..., :2], 4)[..., 1].contiguous()          # [B,h,w,A]
        box_fmap = fmap[..., 2:].contiguous()
        per_im = int(np.prod(box_fmap.shape[1:-1]))
        im_sizes = np.asarray(im_sizes)
        im_hw = to_device_async(np.ascontiguousarray(im_sizes[:, :2].astype(np.float32)), dev)
        im_idx = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(per_im)
        anchors = self.anchors.view(-1, 4).repeat(B, 1)
        box_preds = bbox_preds_fused(anchors, box_fmap.view(-1, 4), 1, im_hw, im_idx)
        hh = torch.arange(class_preds.size(1), device=dev)[None, :, None, None]
        wwi = torch.arange(class_preds.size(2), device=dev)[None, None, :, None]
        h_end = to_device_async(im_sizes[:, 0].astype(np.int64) // self.stride, dev)[:, None, None, None]
        w_end = to_device_async(im_sizes[:, 1].astype(np.int64) // self.stride, dev)[:, None, None, None]
        class_preds = class_preds.masked_fill((hh >= h_end) | (wwi >= w_end), -0.01)
        sizes = center_size(box_preds)
        class_preds = class_preds.reshape(-1)
        class_preds = class_preds.masked_fill((sizes[:, 2] < 4) | (sizes[:, 3] < 4), -0.01)
        return filter_roi_proposals(box_preds, class_preds, boxes_per_im=np.array([per_im] * B),
                                    nms_thresh=nms_thresh, pre_nms_topn=pre_nms_topn, post_nms_topn=post_nms_topn)


def filter_roi_proposals(box_preds, class_preds, boxes_per_im, nms_thresh=0.7, pre_nms_topn=12000, post_nms_topn=2000):
    """:600-612."""
    inds, im_per = apply_nms(class_preds, box_preds, pre_nms_topn=pre_nms_topn, post_nms_topn=post_nms_topn,
                             boxes_per_im=boxes_per_im, nms_thresh=nms_thresh)
    img_inds = torch.repeat_interleave(torch.arange(len(im_per), device=box_preds.device),
                                       to_device_async(np.asarray(im_per, dtype=np.int64), box_preds.device)).float()
    return torch.cat((img_inds[:, None], box_preds[inds]), 1)
