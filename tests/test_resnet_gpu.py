"""GPU checks of the ResNet-101 C4 path (SURVEY.md §8a row a1', lib/resnet_tc.py): the kernel backend of the layer walk
(pinned on the CPU against torchvision by tests/test_resnet_walk.py) against fp64 and against the oracle's detector.
First run on a B200 in round 2 (all green); the round-1 environment gate is gone."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_kernel_ops_match_fp64(cuda):
    from lib.resnet_tc import KernelOps
    ops = KernelOps()
    torch.manual_seed(0)
    x = torch.randn(2, 3, 64, 96, device=cuda)
    stem = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(cuda)
    ref = F.conv2d(x.double(), stem.weight.double(), None, 2, 3).permute(0, 2, 3, 1)
    assert relerr(ops.stem(x, stem), ref) < 3e-5
    y = torch.randn(2, 18, 26, 128, device=cuda)
    for stride in (1, 2):
        c1 = torch.nn.Conv2d(128, 256, 1, stride, bias=False).to(cuda)
        ref = F.conv2d(y.permute(0, 3, 1, 2).double(), c1.weight.double(), None, stride).permute(0, 2, 3, 1)
        assert relerr(ops.conv1x1(y, c1), ref) < 3e-5
        c3 = torch.nn.Conv2d(128, 64, 3, stride, 1, bias=False).to(cuda)
        ref = F.conv2d(y.permute(0, 3, 1, 2).double(), c3.weight.double(), None, stride, 1).permute(0, 2, 3, 1)
        assert relerr(ops.conv3x3(y, c3), ref) < 3e-5
    ref = F.max_pool2d(y.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(ops.maxpool(y), ref)
    for training in (False, True):
        bn = torch.nn.BatchNorm2d(128).to(cuda).train(training)
        bn2 = torch.nn.BatchNorm2d(128).to(cuda).double().train(training)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
        bn2.load_state_dict(bn.state_dict())
        res = torch.randn_like(y)
        with torch.no_grad():
            ref = torch.relu(bn2(y.permute(0, 3, 1, 2).double()).permute(0, 2, 3, 1) + res.double())
            got = ops.bn(y, bn, True, residual=res)
        assert relerr(got, ref) < 1e-5
        assert relerr(bn.running_var, bn2.running_var) < 1e-6 and relerr(bn.running_mean, bn2.running_mean) < 1e-6


@pytest.mark.parametrize("training", [False, True])
def test_small_resnet_walk_on_kernels_vs_torchvision_fp64(cuda, training):
    import copy
    from torchvision.models.resnet import ResNet, Bottleneck
    from lib.resnet_tc import resnet_c4_forward, KernelOps
    torch.manual_seed(1)
    m = ResNet(Bottleneck, [2, 2, 3, 1]).to(cuda).train(training)
    ref_m = copy.deepcopy(m).double()
    x = torch.randn(2, 3, 160, 224, device=cuda)
    with torch.no_grad():
        r = ref_m.maxpool(ref_m.relu(ref_m.bn1(ref_m.conv1(x.double()))))
        ref = ref_m.layer3(ref_m.layer2(ref_m.layer1(r)))
        got = resnet_c4_forward(m, x, KernelOps()).permute(0, 3, 1, 2)
    assert relerr(got, ref) < 3e-4, relerr(got, ref)          # ~30 stacked convs; the 13-conv VGG stack sits at 1.1e-4


def test_resnet_detector_eval_matches_oracle(cuda):
    from lib.object_detector import ObjectDetector
    from oracle import model as OM
    classes = ['__background__'] + ['c%d' % i for i in range(150)]
    torch.manual_seed(2)
    prod = ObjectDetector(classes, mode='gtbox', use_resnet=True)
    with torch.no_grad():
        for mod in prod.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
    orc = OM.ObjectDetector(classes, mode='gtbox', use_resnet=True)
    orc.load_state_dict(prod.state_dict())
    prod = prod.to(cuda).eval(); orc.eval()
    for p in prod.parameters():
        p.requires_grad = False
    x = torch.randn(1, 3, 592, 592)
    rng = np.random.RandomState(0)
    xy = rng.uniform(0, 400, (12, 2)); wh = rng.uniform(32, 190, (12, 2))
    gt_boxes = torch.from_numpy(np.concatenate([xy, np.minimum(xy + wh, 591)], 1).astype(np.float32))
    gt_classes = torch.from_numpy(np.stack([np.zeros(12), rng.randint(1, 151, 12)], 1).astype(np.int64))
    im_sizes = np.array([[592, 592, 1.0]])
    with torch.no_grad():
        ro = orc(x, im_sizes, 0, gt_boxes, gt_classes)
        rp = prod(x.to(cuda), im_sizes, 0, gt_boxes.to(cuda), gt_classes.to(cuda), return_fmap=True)
    assert relerr(rp.fmap.cpu(), ro.fmap) < 1e-3
    assert relerr(rp.od_obj_dists.cpu(), ro.od_obj_dists) < 1e-3
