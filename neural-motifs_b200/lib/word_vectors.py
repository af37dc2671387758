"""Embedding initialisation. The reference's lib/word_vectors.py:17-45 loads GloVe (network /
disk cache) and falls back to N(0,1) for unknown tokens (:20-21). Loading GloVe is init-time and
out of scope here: every row takes the reference's own fallback distribution; trained values come
through the state dict."""
import torch


def obj_edge_vectors(names, wv_type='glove.6B', wv_dir=None, wv_dim=300):
    vectors = torch.empty(len(names), wv_dim)
    vectors.normal_(0, 1)
    return vectors
