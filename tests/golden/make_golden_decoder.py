"""Generates tests/golden/reference_decoder.npz by RUNNING THE REFERENCE's lib/lstm/decoder_rnn.py:DecoderRNN on the
CPU in this container: teacher-forced training forward (labels with background entries, 3 images packed time-major)
and greedy eval with the overlap-aware commitment loop (1 image, boxes_for_nms [T,151,4]).
Environment shims only (sources unedited): `lib.word_vectors.obj_edge_vectors` returns seeded N(0,1) vectors (GloVe is
not downloadable here), `block_orthogonal` is exec'd from the AST span of alternating_highway_lstm.py (the module
itself imports the torch.utils.ffi extension), `torch.Tensor.cuda` is the identity.

    python tests/golden/make_golden_decoder.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import make_golden_host2 as H2  # noqa: E402


def main():
    MG.import_reference()
    import torch
    from torch.nn.utils.rnn import PackedSequence
    torch.Tensor.cuda = lambda self, *a, **k: self
    wv = types.ModuleType("lib.word_vectors")
    wv.obj_edge_vectors = lambda names, wv_dim=100, **kw: torch.randn(len(names), wv_dim)
    sys.modules["lib.word_vectors"] = wv
    import itertools
    # PyTorch-0.3 semantics: `.data` tensors are not Variables (on >= 0.4 every Tensor is one -> infinite recursion)
    ns = dict(torch=torch, itertools=itertools, Variable=type("Variable03", (), {}))
    if not hasattr(torch.nn.init, "orthogonal"):
        torch.nn.init.orthogonal = torch.nn.init.orthogonal_
    exec(H2.extract_function(os.path.join(MG.REF, "lib", "lstm", "highway_lstm_cuda", "alternating_highway_lstm.py"),
                             "block_orthogonal"), ns)
    ah = types.ModuleType("lib.lstm.highway_lstm_cuda.alternating_highway_lstm")
    ah.block_orthogonal = ns["block_orthogonal"]
    import lib.lstm  # noqa  (reference package)
    pkg = types.ModuleType("lib.lstm.highway_lstm_cuda"); pkg.__path__ = []
    sys.modules["lib.lstm.highway_lstm_cuda"] = pkg
    sys.modules["lib.lstm.highway_lstm_cuda.alternating_highway_lstm"] = ah
    from lib.lstm.decoder_rnn import DecoderRNN

    class PS03(PackedSequence):
        """PyTorch-0.3 PackedSequence: a (data, batch_sizes) pair with batch_sizes a Python list."""
        def __new__(cls, data, batch_sizes):
            self = tuple.__new__(cls, (data, list(batch_sizes), None, None))
            return self

        def __iter__(self):
            return iter((tuple.__getitem__(self, 0), tuple.__getitem__(self, 1)))

    torch.manual_seed(0)
    classes = ['__background__'] + ['c%d' % i for i in range(150)]
    H, D = 64, 48
    dec = DecoderRNN(classes, embed_dim=100, inputs_dim=D, hidden_dim=H, recurrent_dropout_probability=0.0)
    with torch.no_grad():
        dec.out.weight.normal_(0, 0.3); dec.out.bias.normal_(0, 0.1)
        dec.input_linearity.bias.normal_(0, 0.1)
    g = {"sd_" + k: v.detach().numpy() for k, v in dec.state_dict().items()}
    g["dims"] = np.array([H, D])

    rng = np.random.RandomState(0)
    # ---- training, teacher forcing: batch sizes per step for lengths (4, 3, 1)
    bl = [3, 2, 2, 1]
    N = sum(bl)
    x = torch.from_numpy(rng.randn(N, D).astype(np.float32))
    labels = torch.from_numpy(rng.randint(0, 151, N).astype(np.int64))
    labels[1] = 0; labels[5] = 0                        # background GT: the model's own argmax is embedded instead
    dec.train()
    dists, commits = dec(PS03(x, bl), labels=labels)
    g["tr_x"], g["tr_bl"], g["tr_labels"] = x.numpy(), np.array(bl), labels.numpy()
    g["tr_dists"], g["tr_commits"] = dists.detach().numpy(), commits.detach().numpy()

    # ---- eval, one image of T objects, greedy + overlap-aware commitments
    T = 40
    xe = rng.randn(T, D).astype(np.float32)
    xe[3] = xe[2] + 0.01 * rng.randn(D); xe[7] = xe[2] + 0.01 * rng.randn(D)     # three near-identical objects ...
    x = torch.from_numpy(xe * 3.0)
    base = MG.rand_boxes(rng, T, lo=40.0)
    base[::2] = base[0] + rng.uniform(-2, 2, (len(base[::2]), 4))   # half of the boxes overlap: equal labels among them get suppressed
    boxes = np.repeat(base[:, None, :], 151, 1) + rng.uniform(-1, 1, (T, 151, 4)).astype(np.float32)
    boxes = torch.from_numpy(np.clip(boxes, 0, 591).astype(np.float32))
    dec.eval()
    with torch.no_grad():
        dists, commits = dec(PS03(x, [1] * T), boxes_for_nms=boxes)
        dists2, commits2 = dec(PS03(x, [1] * T))
    g["ev_x"], g["ev_boxes"] = x.numpy(), boxes.numpy()
    g["ev_dists"], g["ev_commits"] = dists.numpy(), np.asarray(commits.numpy())
    g["ev_commits_greedy"] = commits2.numpy()
    assert not np.array_equal(g["ev_commits"], g["ev_commits_greedy"]), "the fixture must exercise the suppression"
    np.savez_compressed(os.path.join(HERE, "reference_decoder.npz"), **g)
    print("wrote reference_decoder.npz; train commits", g["tr_commits"].tolist(), "eval commits", g["ev_commits"].tolist(),
          "greedy", g["ev_commits_greedy"].tolist())


if __name__ == "__main__":
    main()
