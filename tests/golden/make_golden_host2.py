"""Generates tests/golden/reference_host_ops2.npz by RUNNING more of the REFERENCE's own host code in this container:
  * lib/fpn/proposal_assignments/proposal_assignments_gtbox.py and proposal_assignments_det.py (train-time sampling,
    numpy RNG consumed in the reference's own call order),
  * lib/surgery.py:filter_dets,
  * lib/rel_model.py:_sort_by_score — extracted from the file by its AST span and exec'd unmodified (the module
    itself cannot be imported: it pulls in the torch.utils.ffi extensions).
The reference hard-codes `.cuda(...)`; in this CPU process `torch.Tensor.cuda` is shimmed to the identity (an
environment shim like the h5py stub / np.float alias of make_golden.py — the sources are not edited).

    python tests/golden/make_golden_host2.py
"""
import ast
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def load_synthetic():
    spec = importlib.util.spec_from_file_location("syn", os.path.join(ROOT, "neural-motifs_b200", "dataloaders", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    return syn


def extract_function(path, name):
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
    raise KeyError(name)


def main():
    MG.import_reference()
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self              # CPU shim for the reference's hard .cuda() calls
    from lib.fpn.proposal_assignments.proposal_assignments_gtbox import proposal_assignments_gtbox
    from lib.fpn.proposal_assignments.proposal_assignments_det import proposal_assignments_det
    from lib.surgery import filter_dets
    from lib.pytorch_misc import enumerate_by_image, transpose_packed_sequence_inds
    syn = load_synthetic()
    g = {}

    # ---- proposal_assignments_gtbox: (a) bg subsampled, (b) fg AND bg subsampled, (c) nothing subsampled
    for tag, (nim, boxes, rels, off, seed) in {"a": (3, 20, 15, 6, 3), "b": (2, 20, 100, 0, 4), "c": (1, 5, 4, 2, 5)}.items():
        nb = syn.make_numpy_batch(nim, seed=10 + seed, boxes_per_img=boxes, rels_per_img=rels, image_offset=off)
        gt_boxes, gt_classes, gt_rels = (torch.from_numpy(nb[k]) for k in ("gt_boxes", "gt_classes", "gt_rels"))
        rois = torch.cat(((gt_classes[:, 0] - off).float()[:, None], gt_boxes), 1)
        np.random.seed(seed)
        r, labels, rel_labels = proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, off, fg_thresh=0.5)
        for k, v in dict(gt_boxes=gt_boxes, gt_classes=gt_classes, gt_rels=gt_rels, rois=rois, labels=labels,
                         rel_labels=rel_labels).items():
            g["gtbox_%s_%s" % (tag, k)] = v.detach().numpy()
        g["gtbox_%s_meta" % tag] = np.array([off, seed])

    # ---- proposal_assignments_det: proposals = jittered GT boxes + random boxes, 2 images
    rng = np.random.RandomState(21)
    nb = syn.make_numpy_batch(2, seed=31, boxes_per_img=9, rels_per_img=4, image_offset=4)
    gt_boxes, gt_classes = torch.from_numpy(nb["gt_boxes"]), torch.from_numpy(nb["gt_classes"])
    props = []
    for im in range(2):
        gb = nb["gt_boxes"][nb["gt_classes"][:, 0] - 4 == im]
        jit = np.concatenate([gb + rng.uniform(-s, s, gb.shape) for s in (3, 10, 30, 60)], 0)
        rnd = MG.rand_boxes(rng, 400, lo=20.0)
        b = np.clip(np.concatenate((jit, rnd), 0), 0, 591).astype(np.float32)
        props.append(np.column_stack((np.full(b.shape[0], im, np.float32), b)))
    rois = torch.from_numpy(np.concatenate(props, 0))
    np.random.seed(8)
    out_rois, out_labels, out_targets = proposal_assignments_det(rois, gt_boxes, gt_classes, 4, fg_thresh=0.5)
    for k, v in dict(gt_boxes=gt_boxes, gt_classes=gt_classes, rois=rois, out_rois=out_rois, out_labels=out_labels,
                     out_targets=out_targets).items():
        g["det_" + k] = v.detach().numpy()
    g["det_meta"] = np.array([4, 8])
    # the candidate order the reference saw: its torch.sort on the image index (:33), unspecified among equal keys
    ims = torch.cat([rois[:, 0].long(), gt_classes[:, 0] - 4], 0)
    g["det_sort_idx"] = torch.sort(ims, 0)[1].numpy()

    # ---- filter_dets
    rng = np.random.RandomState(5)
    n, nr = 14, 60
    boxes = torch.from_numpy(MG.rand_boxes(rng, n))
    obj_scores = torch.from_numpy(rng.uniform(0.05, 1, n).astype(np.float32))
    obj_classes = torch.from_numpy(rng.randint(1, 151, n).astype(np.int64))
    pairs = np.array([(i, j) for i in range(n) for j in range(n) if i != j])
    rel_inds = torch.from_numpy(pairs[rng.choice(len(pairs), nr, replace=False)])
    e = np.exp(rng.randn(nr, 51)); pred_scores = torch.from_numpy((e / e.sum(1, keepdims=True)).astype(np.float32))
    fb, fo, fs, fr, fp = filter_dets(boxes, obj_scores, obj_classes, rel_inds, pred_scores)
    for k, v in dict(boxes=boxes, obj_scores=obj_scores, obj_classes=obj_classes, rel_inds=rel_inds, pred_scores=pred_scores).items():
        g["fd_in_" + k] = v.numpy()
    for k, v in dict(boxes=fb, objs=fo, obj_scores=fs, rels=fr, pred_scores=fp).items():
        g["fd_out_" + k] = np.asarray(v)

    # ---- _sort_by_score, exec'd from lib/rel_model.py:31-61
    ns = dict(torch=torch, enumerate_by_image=enumerate_by_image, transpose_packed_sequence_inds=transpose_packed_sequence_inds)
    exec(extract_function(os.path.join(MG.REF, "lib", "rel_model.py"), "_sort_by_score"), ns)
    class T03(torch.Tensor):
        """PyTorch-0.3 indexing semantics the function was written for: a 0-dim result is a Python number."""
        def __getitem__(self, idx):
            r = torch.Tensor.__getitem__(self.as_subclass(torch.Tensor), idx)
            return r.item() if r.dim() == 0 else r

    rng = np.random.RandomState(9)
    for tag, counts in {"a": [7, 3, 9, 1, 4], "b": [20] * 6, "c": [1], "d": [2, 2, 5, 5, 3]}.items():
        im = torch.from_numpy(np.repeat(np.arange(len(counts)), counts))
        scores = torch.from_numpy(rng.rand(int(sum(counts))).astype(np.float32))
        if tag == "d":
            scores[3] = scores[2]                                # a tie inside an image
        perm, inv, ls = ns["_sort_by_score"](im.as_subclass(T03), scores)
        g["sort_%s_im" % tag], g["sort_%s_scores" % tag] = im.numpy(), scores.numpy()
        g["sort_%s_perm" % tag], g["sort_%s_inv" % tag] = perm.numpy(), inv.numpy()
        g["sort_%s_ls" % tag] = np.asarray(ls)
    # ---- clip_grad_norm (lib/pytorch_misc.py:416-459), clipping active (max_norm 5) and inactive (1e6)
    from lib.pytorch_misc import clip_grad_norm
    rng = np.random.RandomState(13)
    shapes = [(33, 7), (5,), (100, 29), (3, 3, 3)]
    grads = [rng.randn(*sh).astype(np.float32) * 2.0 for sh in shapes]
    for tag, mx in (("clip", 5.0), ("noclip", 1e6)):
        ps = []
        for i, gr in enumerate(grads):
            p_ = torch.nn.Parameter(torch.zeros(gr.shape)); p_.grad = torch.from_numpy(gr.copy()); ps.append(("p%d" % i, p_))
        ps.append(("nograd", torch.nn.Parameter(torch.zeros(4))))
        total = clip_grad_norm(ps, max_norm=mx, clip=True, verbose=False)
        g["cgn_%s_total" % tag] = np.array(float(total))
        for i in range(len(grads)):
            g["cgn_%s_after%d" % (tag, i)] = ps[i][1].grad.numpy().copy()
    for i, gr in enumerate(grads):
        g["cgn_grad%d" % i] = gr
    np.savez_compressed(os.path.join(HERE, "reference_host_ops2.npz"), **g)
    print("wrote reference_host_ops2.npz with", len(g), "arrays;",
          {k: g[k].shape for k in ("gtbox_a_rel_labels", "gtbox_b_rel_labels", "gtbox_c_rel_labels", "det_out_rois", "fd_out_rels")})


if __name__ == "__main__":
    main()
