"""Deterministic synthetic model state and inputs shared by tests/golden/make_golden_model.py (which loads them into
the REFERENCE's RelModel) and tests/test_reference_model_pin.py (which loads them into the oracle's): every tensor of
a state dict is regenerated from (name, shape, seed), so the 1.7 GB of weights never has to be stored."""
import zlib

import numpy as np
import torch

CLASSES = ['__background__'] + ['obj%d' % i for i in range(150)]
RELS = ['__background__'] + ['rel%d' % i for i in range(50)]
KW = dict(hidden_dim=512, pooling_dim=4096, nl_obj=2, nl_edge=4, order='leftright', use_bias=True,
          use_tanh=False, limit_vision=False)           # scripts/train_models_sgcls.sh:19-21


def synthetic_state(named, seed=0):
    """named: iterable of (key, shape, dtype). Scales keep activations O(1) through VGG16, fc6/fc7 and the LSTMs."""
    sd = {}
    for key, shape, dtype in named:
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
        if not dtype.is_floating_point:
            sd[key] = torch.zeros(shape, dtype=dtype)                      # num_batches_tracked
            continue
        n = int(np.prod(shape)) if len(shape) else 1
        if "pos_embed.0.running_" in key:
            # BatchNorm1d over (cx, cy, w, h) in pixels of the 592-scale image (rel_model.py:97-102): statistics of the
            # size real boxes have. With the generic N(0, 0.1) / U(0.5, 1.5) statistics below the 4 -> 128 position
            # embedding comes out ~1e3 large, the highway LSTMs saturate and the recurrence turns chaotic: measured on the
            # CPU, the reference recurrence in fp32 and in fp64 then differ by 1.3 % and a 1e-5 input perturbation moves
            # the output by 10-40 % - no two fp32 implementations (cuBLAS included) could agree to 1e-3 on that state.
            v = (250.0 + 30.0 * torch.randn(shape, generator=g)) if key.endswith("running_mean") else \
                (120.0 * (torch.rand(shape, generator=g) + 0.5)) ** 2
        elif key.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif key.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 4:                                              # conv: Kaiming (ReLU)
            v = torch.randn(shape, generator=g) * float(np.sqrt(2.0 / (shape[1] * shape[2] * shape[3])))
        elif len(shape) == 2 and "embed" not in key and "obj_baseline" not in key:
            v = torch.randn(shape, generator=g) * float(np.sqrt(1.0 / shape[1]))
        elif len(shape) == 2:                                              # embeddings, frequency-bias table
            v = torch.randn(shape, generator=g)
        elif "rnn.weight" in key:                                          # flat highway-LSTM weights
            v = torch.randn(shape, generator=g) * 0.03
        elif key.endswith(".bn3.weight"):                                  # last BN of a ResNet bottleneck: small, so that
            v = torch.rand(shape, generator=g) * 0.2 + 0.1                 # 33 residual blocks keep activations O(1)
        elif key.endswith(".weight") and len(shape) == 1:                  # BatchNorm scale
            v = torch.rand(shape, generator=g) + 0.5
        else:                                                              # biases
            v = torch.randn(shape, generator=g) * 0.05
        assert v.numel() == n
        sd[key] = v.to(dtype)
    return sd


def make_inputs(seed=0, boxes=20, rels=15):
    """One 592x592 image, `boxes` GT boxes, `rels` GT relations (BASELINE config 1)."""
    rng = np.random.RandomState(seed)
    imgs = rng.randn(1, 3, 592, 592).astype(np.float32)
    x1 = rng.uniform(0, 400, boxes); y1 = rng.uniform(0, 400, boxes)
    w = rng.uniform(32, 190, boxes); h = rng.uniform(32, 190, boxes)
    gt_boxes = np.stack([x1, y1, np.minimum(x1 + w, 591), np.minimum(y1 + h, 591)], 1).astype(np.float32)
    gt_classes = np.stack([np.zeros(boxes), rng.randint(1, 151, boxes)], 1).astype(np.int64)
    pairs = [(a, b) for a in range(boxes) for b in range(boxes) if a != b]
    sel = np.sort(rng.choice(len(pairs), rels, replace=False))
    gt_rels = np.array([[0, pairs[k][0], pairs[k][1], rng.randint(1, 51)] for k in sel], dtype=np.int64)
    return dict(imgs=imgs, im_sizes=np.array([[592, 592, 0.578]], dtype=np.float32), gt_boxes=gt_boxes,
                gt_classes=gt_classes, gt_rels=gt_rels)
