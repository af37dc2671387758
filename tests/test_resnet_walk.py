"""CPU pin of the ResNet C4 graph walk (lib/resnet_tc.py, SURVEY.md §8a row a1'): the layer-by-layer walk over
torchvision's modules — NHWC activations, stride-2 1x1 convs as subsample + GEMM, stride-2 3x3 convs as the
stride-1 result subsampled, BatchNorm in the module's mode with the residual add fused — against torchvision's
own forward (the reference's `feature_map`, lib/object_detector.py:119-127), through a torch backend that
implements the five backend operations with the same contracts as the kernel backend."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))


class TorchOps(object):
    def stem(self, x, conv):
        return F.conv2d(x, conv.weight, None, 2, 3).permute(0, 2, 3, 1).contiguous()

    def conv1x1(self, x, conv):
        if conv.stride == (2, 2):
            x = x[:, ::2, ::2, :]
        return F.linear(x, conv.weight.view(conv.out_channels, -1), conv.bias)

    def conv3x3(self, x, conv):
        y = F.conv2d(x.permute(0, 3, 1, 2), conv.weight, conv.bias, 1, 1).permute(0, 2, 3, 1)
        return y[:, ::2, ::2, :].contiguous() if conv.stride == (2, 2) else y

    def bn(self, x, bn, relu, residual=None):
        C = x.size(-1)
        x2 = x.reshape(-1, C)
        if bn.training:
            mean = x2.mean(0); var = x2.var(0, unbiased=False)
            with torch.no_grad():
                n = x2.size(0)
                bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * mean)
                bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * var * n / (n - 1))
                bn.num_batches_tracked += 1
            invstd = torch.rsqrt(var + bn.eps)
        else:
            mean, invstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
        y = (x - mean) * (invstd * bn.weight) + bn.bias
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y

    def maxpool(self, x):
        return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()


def _small_resnet(seed):
    from torchvision.models.resnet import ResNet, Bottleneck
    torch.manual_seed(seed)
    m = ResNet(Bottleneck, [2, 2, 3, 1])
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
    return m


def _tv_c4(m, x):
    x = m.maxpool(m.relu(m.bn1(m.conv1(x))))
    return m.layer3(m.layer2(m.layer1(x)))


def test_resnet_c4_walk_matches_torchvision_eval_and_train():
    import copy
    from lib.resnet_tc import resnet_c4_forward
    m = _small_resnet(0).double()
    x = torch.randn(2, 3, 80, 112, dtype=torch.float64)
    with torch.no_grad():
        for training in (False, True):
            a, b = copy.deepcopy(m).train(training), copy.deepcopy(m).train(training)
            ref = _tv_c4(a, x)
            got = resnet_c4_forward(b, x, TorchOps()).permute(0, 3, 1, 2)
            assert got.shape == ref.shape == (2, 1024, 5, 7)
            assert float((got - ref).abs().max()) < 1e-9 * float(ref.abs().max())
            for (n, p), (_, q) in zip(a.named_buffers(), b.named_buffers()):      # running statistics moved alike
                assert torch.allclose(p.double(), q.double(), rtol=1e-9, atol=1e-12), n


def test_resnet101_shapes_for_the_detector_config():
    """592x592 -> 37x37x1024 through 3 + 4 + 23 bottlenecks; only the structure is checked here."""
    from torchvision.models.resnet import resnet101
    from lib.resnet_tc import resnet_c4_forward
    m = resnet101(weights=None).eval()
    calls = {"conv1x1": 0, "conv3x3": 0, "s2": 0}

    class Shapes(TorchOps):
        def stem(self, x, conv):
            return x.new_zeros(x.size(0), (x.size(2) + 1) // 2, (x.size(3) + 1) // 2, 64)

        def conv1x1(self, x, conv):
            calls["conv1x1"] += 1
            if conv.stride == (2, 2):
                calls["s2"] += 1; x = x[:, ::2, ::2]
            assert x.size(-1) == conv.in_channels
            return x.new_zeros(*x.shape[:3], conv.out_channels)

        def conv3x3(self, x, conv):
            calls["conv3x3"] += 1
            assert x.size(-1) == conv.in_channels and conv.in_channels % 64 == 0
            if conv.stride == (2, 2):
                calls["s2"] += 1; x = x[:, ::2, ::2]
            return x.new_zeros(*x.shape[:3], conv.out_channels)

        def bn(self, x, bn, relu, residual=None):
            assert x.size(-1) == bn.num_features and (residual is None or residual.shape == x.shape)
            return x

    y = resnet_c4_forward(m, torch.zeros(1, 3, 592, 592), Shapes())
    assert y.shape == (1, 37, 37, 1024)
    assert calls == {"conv1x1": 2 * 30 + 3, "conv3x3": 30, "s2": 4}


def test_resnet_detector_state_dict_keys_match_between_product_and_oracle(monkeypatch):
    """Construction only (no CUDA): the product's ResNet detector and the oracle's carry the same state dict
    (the reference's module tree, lib/object_detector.py:84-103), and the oracle one runs on the CPU."""
    import numpy as np
    import pytest
    from lib.object_detector import ObjectDetector
    from oracle import model as OM
    classes = ['__background__'] + ['c%d' % i for i in range(10)]
    prod = ObjectDetector(classes, mode='gtbox', use_resnet=True)
    orc = OM.ObjectDetector(classes, mode='gtbox', use_resnet=True)
    assert set(prod.state_dict().keys()) == set(orc.state_dict().keys())
    orc.load_state_dict(prod.state_dict())
    orc.eval()
    x = torch.randn(1, 3, 128, 160)
    gt_boxes = torch.tensor([[10., 12., 90., 100.], [30., 40., 150., 120.]])
    gt_classes = torch.tensor([[0, 3], [0, 7]])
    res = orc(x, np.array([[128, 160, 1.0]]), 0, gt_boxes, gt_classes)
    assert res.fmap.shape == (1, 1024, 8, 10) and res.od_obj_dists.shape == (2, 11)
    assert torch.isfinite(res.od_obj_dists).all()
