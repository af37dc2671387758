"""Generates tests/golden/reference_blob.npz by RUNNING THE REFERENCE's dataloaders/blob.py:Blob (the collation that
produces `RelModel.forward`'s positional tuple, blob.py:62-229) on the CPU over dataset entries built from this
repo's synthetic generator. blob.py does not parse on Python >= 3.7 (`x.cuda(dev, async=True)`): as in
make_golden_relassign.py the keyword is renamed `non_blocking=` in memory before exec; `Tensor.cuda` is the identity.
The fixture holds the tuple the reference hands to forward; tests/test_reference_host_pins.py checks that
dataloaders/synthetic.py's `to_tuple` / `SyntheticBlob` produce the same tuple from the same numpy batch.

    python tests/golden/make_golden_blob.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import make_golden_host2 as H2  # noqa: E402


def main():
    MG.import_reference()
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    src = open(os.path.join(MG.REF, "dataloaders", "blob.py")).read()
    ns = {"__name__": "ref_blob"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(src.replace("async=True", "non_blocking=True"), "blob.py", "exec"), ns)
    Blob = ns["Blob"]
    syn = H2.load_synthetic()
    g = {}
    B = 3
    nb = syn.make_numpy_batch(B, seed=9, boxes_per_img=7, rels_per_img=5, image_offset=0)
    imgs = nb["imgs"][:, :, :64, :80].copy()                  # small images keep the fixture small; layout is what matters
    for mode, train in (("rel", True), ("rel", False)):
        blob = Blob(mode=mode, is_train=train, num_gpus=1, batch_size_per_gpu=B)
        np.random.seed(5)
        for i in range(B):
            sel = nb["gt_classes"][:, 0] == i
            rel = nb["gt_rels"][nb["gt_rels"][:, 0] == i, 1:]
            blob.append(dict(img=torch.from_numpy(imgs[i]), img_size=tuple(nb["im_sizes"][i]),
                             gt_boxes=nb["gt_boxes"][sel], gt_classes=nb["gt_classes"][sel, 1], gt_relations=rel,
                             scale=1.0, index=i, flipped=False))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            blob.reduce()
            blob.scatter()
            tup = blob[0]
        tag = "train" if train else "eval"
        assert len(tup) == (8 if train else 7)
        g[tag + "_imgs"] = tup[0].detach().numpy()
        g[tag + "_im_sizes"] = np.asarray(tup[1])
        g[tag + "_image_offset"] = np.array(tup[2])
        g[tag + "_gt_boxes"] = tup[3].detach().numpy()
        g[tag + "_gt_classes"] = tup[4].detach().numpy()
        g[tag + "_gt_rels"] = tup[5].detach().numpy()
        assert tup[6] is None
        if train:
            g["train_anchor_inds"] = tup[7].detach().numpy()
    for k in ("gt_boxes", "gt_classes", "gt_rels", "im_sizes"):
        g["nb_" + k] = nb[k]
    g["nb_imgs"] = imgs
    np.savez_compressed(os.path.join(HERE, "reference_blob.npz"), **g)
    print("wrote reference_blob.npz:", {k: v.shape for k, v in g.items() if k.startswith("train_")})


if __name__ == "__main__":
    main()
