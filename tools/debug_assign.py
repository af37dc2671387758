import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location('t', os.path.join(ROOT, 'tests/test_ops_gpu.py')); t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
from lib.fpn.proposal_assignments.proposal_assignments_det import proposal_assignments_det
from oracle import host
cuda = torch.device("cuda:0")
rng = np.random.RandomState(22)
im_inds, rois, labels, gt_boxes, gt_classes, gt_rels = t._synthetic_detections(rng, num_im=2, per_im=400, gt_per_im=10)
rois5 = np.concatenate([im_inds[:, None].astype(np.float32), rois], 1)
r, l, tt = proposal_assignments_det(torch.from_numpy(rois5).to(cuda), torch.from_numpy(gt_boxes).to(cuda), torch.from_numpy(gt_classes).to(cuda), 0, rng=np.random.RandomState(9))
er, el, et = host.proposal_assignments_det(rois5, gt_boxes, gt_classes, 0, np.random.RandomState(9))
r, l, tt = r.cpu().numpy(), l.cpu().numpy(), tt.cpu().numpy()
print("shapes", r.shape, er.shape, l.shape, el.shape)
if r.shape == er.shape:
    print("rois equal", np.array_equal(r, er), "labels equal", np.array_equal(l, el), "targets equal", np.array_equal(tt, et))
    bad = np.where((r != er).any(1))[0]; print("bad roi rows", bad[:10], len(bad))
    bad = np.where(l != el)[0]; print("bad labels", bad[:10], len(bad))
    bad = np.where((tt != et).any(1))[0]; print("bad targets", bad[:10], len(bad))
    if len(bad): print(tt[bad[0]], et[bad[0]], l[bad[0]], el[bad[0]])
