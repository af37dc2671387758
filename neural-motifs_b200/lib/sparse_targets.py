"""FrequencyBias lookup — lib/sparse_targets.py:11-37 of the reference. The table
log P(pred | subj, obj) is built there by iterating the Visual Genome training set at
construction (get_dataset_counts.py:12); here the [num_objs^2, num_rels] table is supplied by the
caller (state dict / synthetic) — only the lookup is on the hot path."""
import torch
import torch.nn as nn


class FrequencyBias(nn.Module):
    def __init__(self, num_objs=151, num_rels=51, pred_dist=None, eps=1e-3):
        super().__init__()
        self.num_objs = num_objs
        self.obj_baseline = nn.Embedding(num_objs * num_objs, num_rels)
        if pred_dist is not None:
            self.obj_baseline.weight.data = torch.as_tensor(pred_dist, dtype=torch.float32).view(-1, num_rels)

    def index_with_labels(self, labels):
        """labels [N,2] (subject class, object class) -> [N,num_rels] (:32-37)."""
        return self.obj_baseline(labels[:, 0] * self.num_objs + labels[:, 1])

    def forward(self, obj_cands0, obj_cands1):
        joint = obj_cands0[:, :, None] * obj_cands1[:, None]
        return joint.view(joint.size(0), -1) @ self.obj_baseline.weight
