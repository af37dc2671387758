"""Synthetic Visual-Genome-shaped batches (no dataset / network in this environment): the per-GPU
tuple `Blob.__getitem__` hands to `RelModel.forward` in the reference (dataloaders/blob.py:182-229):
(imgs, im_sizes[B,3]=(h,w,scale), image_offset, gt_boxes[G,4] @592 scale, gt_classes[G,2]=(global img,
class), gt_rels[R,4]=(global img, subj, obj, predicate; box indices local to the image), proposals,
train_anchor_inds). Shapes and distributions are those of SURVEY.md §8d configs 1-2."""
import numpy as np
import torch

IM_SCALE = 592


def make_numpy_batch(batch_size, seed=0, boxes_per_img=20, rels_per_img=15, num_classes=151, num_rels=51,
                     image_offset=0, vg_shaped=False):
    rng = np.random.RandomState(seed)
    imgs = rng.randn(batch_size, 3, IM_SCALE, IM_SCALE).astype(np.float32)
    im_sizes = np.tile(np.array([[IM_SCALE, IM_SCALE, 0.578]], np.float32), (batch_size, 1))
    gt_boxes, gt_classes, gt_rels = [], [], []
    for i in range(batch_size):
        n = int(np.clip(rng.poisson(12), 3, 40)) if vg_shaped else boxes_per_img
        x1 = rng.uniform(0, 400, n); y1 = rng.uniform(0, 400, n)
        w = rng.uniform(32, 190, n); h = rng.uniform(32, 190, n)
        gt_boxes.append(np.stack([x1, y1, np.minimum(x1 + w, IM_SCALE - 1), np.minimum(y1 + h, IM_SCALE - 1)], 1))
        gt_classes.append(np.stack([np.full(n, i + image_offset), rng.randint(1, num_classes, n)], 1))
        pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
        sel = rng.choice(len(pairs), size=min(rels_per_img, len(pairs)), replace=False)
        sel.sort()
        rel = np.array([[i + image_offset, pairs[k][0], pairs[k][1], rng.randint(1, num_rels)] for k in sel])
        gt_rels.append(rel)
    return dict(imgs=imgs, im_sizes=im_sizes, image_offset=image_offset,
                gt_boxes=np.concatenate(gt_boxes).astype(np.float32),
                gt_classes=np.concatenate(gt_classes).astype(np.int64),
                gt_rels=np.concatenate(gt_rels).astype(np.int64))


def to_tuple(nb, device, is_train=True):
    """numpy batch -> the positional tuple of `forward` (proposals None, train_anchor_inds None)."""
    t = lambda a: torch.from_numpy(a).to(device)
    return (t(nb["imgs"]), nb["im_sizes"], nb["image_offset"], t(nb["gt_boxes"]), t(nb["gt_classes"]),
            t(nb["gt_rels"]), None, None)


class SyntheticBlob(object):
    """Stands in for dataloaders/blob.py:Blob in `detector[blob]` (models/train_rels.py:137): holds
    PINNED host tensors; `scatter()` issues the async H2D copies (blob.py:155-180); `blob[0]` is the
    forward tuple of this rank."""

    def __init__(self, nb, device, is_train=True):
        self.device = torch.device(device)
        self.is_train = is_train
        self.im_sizes = nb["im_sizes"]
        self.image_offset = nb["image_offset"]
        pin = lambda a: torch.from_numpy(a).pin_memory() if torch.cuda.is_available() else torch.from_numpy(a)
        self.host = {k: pin(nb[k]) for k in ("imgs", "gt_boxes", "gt_classes", "gt_rels")}
        self.dev = None

    def h2d_bytes(self):
        return int(sum(v.numel() * v.element_size() for v in self.host.values()))

    def scatter(self):
        self.dev = {k: v.to(self.device, non_blocking=True) for k, v in self.host.items()}
        # the labels exist on the host already: RelModel.forward takes the image index / foreground test from here instead
        # of reading gt_classes back from the device (lib/rel_model.py EARLY_HOST_INDS)
        self.dev["gt_classes"]._mb200_host = self.host["gt_classes"].numpy()

    def __getitem__(self, index):
        if index != 0:
            raise ValueError("one process per GPU: only index 0 exists")
        d = self.dev
        return (d["imgs"], self.im_sizes, self.image_offset, d["gt_boxes"], d["gt_classes"], d["gt_rels"], None, None)


def synthetic_model_state(model, seed=0):
    """Fill the data-dependent tables the reference builds from VG / GloVe (frequency bias
    `lib/sparse_targets.py:16-30`, word vectors) with seeded synthetic values, in place."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        fb = getattr(model, "freq_bias", None)
        if fb is not None:
            p = torch.rand(fb.obj_baseline.weight.shape, generator=g)
            p = p / p.sum(1, keepdim=True)
            fb.obj_baseline.weight.copy_(torch.log(p + 1e-3))
    return model
