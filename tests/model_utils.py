"""Shared helpers for the model-level parity tests (oracle on CPU vs product on cuda:0)."""
import numpy as np
import torch

from dataloaders.synthetic import make_numpy_batch, synthetic_model_state

CLASSES = ['__background__'] + ['obj%d' % i for i in range(150)]
RELS = ['__background__'] + ['rel%d' % i for i in range(50)]
KW = dict(hidden_dim=512, pooling_dim=4096, nl_obj=2, nl_edge=4, order='leftright', use_bias=True,
          use_tanh=False, limit_vision=False)


def build_pair(mode, seed=0, rec_dropout=0.1):
    """Product RelModel (CPU-constructed, MotifNet script config scripts/train_models_sgcls.sh:19-21)
    and an oracle RelModel carrying the SAME state dict."""
    from lib.rel_model import RelModel
    from oracle import model as OM
    torch.manual_seed(seed)
    prod = RelModel(CLASSES, RELS, mode=mode, num_gpus=1, require_overlap_det=True, use_resnet=False,
                    use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                    rec_dropout=rec_dropout, **KW)
    synthetic_model_state(prod, seed)
    with torch.no_grad():   # random-init VGG is badly scaled for deep stacks; keep activations O(1)
        for m in prod.modules():
            if isinstance(m, torch.nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight, nonlinearity='relu')
                m.bias.normal_(0, 0.01)
    orc = OM.RelModel(CLASSES, RELS, mode=mode, **KW)
    orc.load_state_dict(prod.state_dict())
    for p in prod.detector.parameters():       # models/train_rels.py:51-52
        p.requires_grad = False
    for p in orc.detector.parameters():
        p.requires_grad = False
    return prod, orc


def make_masks(n_obj, n_rel, n_img, seed=0, H=512, nl_obj=2, nl_edge=4):
    g = torch.Generator().manual_seed(seed)

    def bern(shape, p):
        return (torch.rand(shape, generator=g) > p).float() / (1 - p)

    det = {"roi_fmap.2": bern((n_obj, 4096), 0.5), "roi_fmap.5": bern((n_obj, 4096), 0.5)}
    top = {"roi_fmap_obj.2": bern((n_obj, 4096), 0.5), "roi_fmap_obj.5": bern((n_obj, 4096), 0.5),
           "roi_fmap.1.2": bern((n_rel, 4096), 0.5)}
    ctx = {"pos_embed.3": bern((n_obj, 128), 0.1), "obj_ctx_rnn": bern((nl_obj, n_img, H), 0.1),
           "decoder_rnn": bern((n_img, H), 0.1), "edge_ctx_rnn": bern((nl_edge, n_img, H), 0.1)}
    return det, top, ctx


def to_dev(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


def relerr(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2err(a, b):
    """Relative L2 error — used for gradients, where a handful of ReLU units whose pre-activation
    sits within fp noise of zero legitimately flip between two fp32 implementations."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
