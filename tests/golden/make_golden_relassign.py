"""Generates tests/golden/reference_rel_assignments.npz by RUNNING THE REFERENCE's
lib/fpn/proposal_assignments/rel_assignments.py (SGDet training: relation labels for detected boxes) on the CPU.
That file does not parse on Python >= 3.7 (`.cuda(device, async=True)`, rel_assignments.py:143-144): it is read as
text, the keyword `async=` is renamed `non_blocking=` IN MEMORY (the modernisation any caller needs),
and exec'd; nothing else is changed and nothing is written back. `torch.Tensor.cuda` is shimmed to the identity.

    python tests/golden/make_golden_relassign.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import make_golden_host2 as H2  # noqa: E402


def main():
    MG.import_reference()
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    src = open(os.path.join(MG.REF, "lib", "fpn", "proposal_assignments", "rel_assignments.py")).read()
    assert src.count("async=True") == 1
    ns = {"__name__": "ref_rel_assignments"}
    exec(compile(src.replace("async=True", "non_blocking=True"), "rel_assignments.py", "exec"), ns)
    rel_assignments = ns["rel_assignments"]
    syn = H2.load_synthetic()
    g = {}
    for tag, (nper, nsg, fno, seed) in {"a": (1, 1, True, 2), "b": (4, 4, True, 3), "c": (2, 1, False, 4)}.items():
        rng = np.random.RandomState(40 + seed)
        off = 3
        nb = syn.make_numpy_batch(2, seed=50 + seed, boxes_per_img=10, rels_per_img=8, image_offset=off)
        gt_boxes, gt_classes, gt_rels = nb["gt_boxes"], nb["gt_classes"], nb["gt_rels"]
        boxes, labels, ims = [], [], []
        for im in range(2):
            sel = gt_classes[:, 0] - off == im
            gb, gc = gt_boxes[sel], gt_classes[sel, 1]
            for rep in range(nper):                                   # detections that match a GT box (IoU >= 0.5) ...
                boxes.append(np.clip(gb + rng.uniform(-4, 4, gb.shape), 0, 591)); labels.append(gc)
            boxes.append(np.clip(gb + rng.uniform(-60, 60, gb.shape), 0, 591)); labels.append(np.zeros_like(gc))   # ... and misses
            ims.append(np.full((nper + 1) * gb.shape[0], im))
        boxes = np.concatenate(boxes, 0).astype(np.float32); labels = np.concatenate(labels).astype(np.int64)
        ims = np.concatenate(ims).astype(np.int64)
        np.random.seed(seed)
        out = rel_assignments(torch.from_numpy(ims), torch.from_numpy(boxes), torch.from_numpy(labels),
                              torch.from_numpy(gt_boxes), torch.from_numpy(gt_classes.copy()), torch.from_numpy(gt_rels.copy()),
                              off, filter_non_overlap=fno, num_sample_per_gt=nsg)
        for k, v in dict(ims=ims, boxes=boxes, labels=labels, gt_boxes=gt_boxes, gt_classes=gt_classes, gt_rels=gt_rels,
                         out=out.numpy()).items():
            g["ra_%s_%s" % (tag, k)] = v
        g["ra_%s_meta" % tag] = np.array([off, seed, nsg, int(fno)])
    np.savez_compressed(os.path.join(HERE, "reference_rel_assignments.npz"), **g)
    print("wrote reference_rel_assignments.npz:", {t: (g["ra_%s_out" % t].shape, int((g["ra_%s_out" % t][:, 3] > 0).sum()))
                                                    for t in "abc"})


if __name__ == "__main__":
    main()
