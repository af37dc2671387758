"""GPU checks of the detector-training gradient path (lib/conv_tc.py): conv3x3 with autograd on the
tcgen05 kernels vs fp64, the trainable VGG stack vs the forward-only path and fp64 gradients, and one rpntrain step of
ObjectDetector (models/train_detector.py:78-155). The formulas are pinned on the CPU by tests/test_conv_tc_walk.py."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def l2err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,H,W,Ci,Co", [(2, 16, 24, 64, 128), (1, 37, 37, 512, 512), (3, 9, 70, 128, 64)])
def test_conv3x3_kernel_backend_forward_and_gradients_vs_fp64(cuda, B, H, W, Ci, Co):
    from lib import conv_tc
    torch.manual_seed(B + H)
    x = torch.randn(B, H, W, Ci, device=cuda, requires_grad=True)
    w = (torch.randn(Co, Ci, 3, 3, device=cuda) / (9 * Ci) ** 0.5).requires_grad_(True)
    b = torch.randn(Co, device=cuda, requires_grad=True)
    y = conv_tc.conv3x3(x, w, b, relu=True)
    x2, w2, b2 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = torch.relu(F.conv2d(x2.permute(0, 3, 1, 2), w2, b2, 1, 1)).permute(0, 2, 3, 1)
    assert relerr(y, ref) < 3e-5
    g = torch.randn_like(y) * (y > 0).float()            # keep the comparison on the shared ReLU piece
    y.backward(g)
    ref.backward(g.double())
    assert l2err(x.grad, x2.grad) < 1e-4 and l2err(w.grad, w2.grad) < 1e-4 and relerr(b.grad, b2.grad) < 1e-4


def test_vgg_features_train_matches_forward_only_path_and_fp64_gradients(cuda):
    from torchvision.models.vgg import vgg16
    from lib import conv_tc, tc_ops
    torch.manual_seed(0)
    feats = vgg16(weights=None).features
    del feats._modules['30']
    feats = feats.to(cuda)
    with torch.no_grad():
        for m in feats:
            if isinstance(m, torch.nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight, nonlinearity='relu'); m.bias.normal_(0, 0.01)
    convs = [m for m in feats if isinstance(m, torch.nn.Conv2d)]
    x = torch.randn(1, 3, 96, 128, device=cuda)
    y = conv_tc.vgg_features_train(x, convs, tc_ops.VGG16_CFG)
    with torch.no_grad():
        fwd, _ = tc_ops.vgg_features_forward(x, convs)
    assert relerr(y, fwd) < 1e-4
    import copy
    ref_feats = copy.deepcopy(torch.nn.Sequential(*[m for m in feats])).double()     # (.double() converts in place)
    ref = ref_feats(x.double())
    assert relerr(y.permute(0, 3, 1, 2), ref) < 3e-4
    g = torch.randn_like(y)
    y.backward(g)
    ref.backward(g.permute(0, 3, 1, 2).double())
    for (n, p), (_, q) in zip(feats.named_parameters(), ref_feats.named_parameters()):
        assert p.grad is not None and l2err(p.grad, q.grad) < 2e-2, (n, l2err(p.grad, q.grad))   # ReLU / pool pieces differ at ties


def test_detector_rpntrain_step_runs(cuda):
    """models/train_detector.py:78-155 shape of a step: RPN + detection losses, gradients reach conv1_1."""
    import numpy as np
    from lib.object_detector import ObjectDetector
    from lib.fpn.anchor_targets import anchor_target_layer
    from dataloaders.synthetic import make_numpy_batch, to_tuple
    classes = ['__background__'] + ['c%d' % i for i in range(150)]
    torch.manual_seed(0)
    det = ObjectDetector(classes, mode='rpntrain').to(cuda).train()
    det.rng = np.random.RandomState(0)
    B = 1
    nb = make_numpy_batch(B, seed=3, boxes_per_img=8, rels_per_img=4)
    gb = nb["gt_boxes"]
    _, inds, _, labels = anchor_target_layer(gb, (592, 592), rng=np.random.RandomState(1))
    tai = torch.from_numpy(np.column_stack((np.zeros(inds.shape[0]), inds)).astype(np.int64)).to(cuda)
    tup = list(to_tuple(nb, cuda))
    res = det(tup[0], tup[1], tup[2], tup[3], tup[4], None, None, tai)
    loss = F.cross_entropy(res.od_obj_dists, res.od_obj_labels) + \
        F.cross_entropy(res.rpn_scores, torch.from_numpy(labels.astype(np.int64)).to(cuda))    # labels: [num_used] in {0, 1}
    loss.backward()
    g0 = det.features[0].weight.grad
    assert torch.isfinite(loss) and g0 is not None and torch.isfinite(g0).all() and float(g0.abs().max()) > 0
    assert det.rpn_head.conv[0].weight.grad is not None
