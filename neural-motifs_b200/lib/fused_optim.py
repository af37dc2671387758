"""Flat-buffer SGD for the relation model's trainable parameters: same update rule as the caller's
`clip_grad_norm(..., max_norm=conf.clip)` + `optim.SGD(params, lr, momentum=0.9, weight_decay=conf.l2)`
(models/train_rels.py:57-70,145-150) — including the lr/10 group for the VGG fc layers — but every
parameter and gradient lives in ONE contiguous buffer per group, so that
  * the global gradient norm is one reduction,
  * the data-parallel gradient all-reduce is one NCCL call on the flat gradient (no bucket copies),
  * clip + weight decay + momentum + update + gradient zeroing is one fused kernel (csrc/optim.cu).
Parameters stay ordinary nn.Parameters (names / state dict unchanged); their storage is re-pointed."""
import torch
import torch.distributed as dist

import motifs_cabi as _c
from lib import tc_ops


class FlatGroup(object):
    def __init__(self, params, lr, chunk_bytes=128 << 20):
        self.params = params
        self.lr = lr
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4            # keep every view 16-byte aligned
        self.n = n
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[o:o + p.numel()].view_as(p)
        # communication chunks: contiguous runs of whole parameters, ~chunk_bytes each. A chunk is
        # all-reduced as soon as autograd has produced the gradient of every parameter in it.
        self.chunks, cur, start, size = [], [], 0, 0
        for p, o in zip(params, offs):
            n = (p.numel() + 3) // 4 * 4
            if cur and size + n * 4 > chunk_bytes:
                self.chunks.append((start, o, cur))
                cur, start, size = [], o, 0
            cur.append(p)
            size += n * 4
        if cur:
            self.chunks.append((start, self.n, cur))


class FlatSGD(object):
    """groups: list of (params, lr). momentum / weight_decay / max_norm shared (train_rels.py:66,145)."""

    def __init__(self, groups, momentum=0.9, weight_decay=1e-4, max_norm=5.0, overlap_comm=True, chunk_bytes=128 << 20):
        self.groups = [FlatGroup([p for p in ps if p.requires_grad], lr, chunk_bytes) for ps, lr in groups if len(ps)]
        self.momentum, self.weight_decay, self.max_norm = momentum, weight_decay, max_norm
        self.steps = 0
        tc_ops.bump_weight_epoch()       # storages moved
        self._works = []
        self._pending = {}
        self._distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self._overlap = self._distributed and overlap_comm
        self._seen = set()
        # One hook per parameter. (1) It marks the parameter's flat gradient as touched by autograd, which ends
        # the window in which weight-gradient GEMMs may write straight into it (tc_ops.direct_grad_target).
        # (2) With data-parallel overlap it counts down the parameter's communication chunk; the last one
        # launches the chunk's all-reduce (NCCL stream) while autograd keeps going.
        for g in self.groups:
            for ci, (a, b, ps) in enumerate(g.chunks):
                for p in ps:
                    hook = self._make_hook(g, ci)
                    p._mb200_direct = tc_ops.DirectGradState()
                    p.register_post_accumulate_grad_hook(self._make_autograd_hook(hook))

    def _make_autograd_hook(self, hook):
        def on_accumulate(param):
            param._mb200_direct.dirty = True
            if self._overlap:
                hook(param)
        return on_accumulate

    def _make_hook(self, group, ci):
        def hook(param):
            if id(param) in self._seen:      # count every parameter once per step
                return
            self._seen.add(id(param))
            key = (id(group), ci)
            left = self._pending.get(key, len(group.chunks[ci][2])) - 1
            self._pending[key] = left
            if left == 0:
                a, b, _ = group.chunks[ci]
                self._works.append((dist.all_reduce(group.flat_g[a:b], op=dist.ReduceOp.SUM, async_op=True), group, a, b))
        return hook

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed by the fused step itself; kept for API symmetry (never set to None:
        autograd accumulates into the flat views)."""
        if self.steps == 0:
            for g in self.groups:
                g.flat_g.zero_()
                for p in g.params:
                    p._mb200_direct.reset()

    def all_reduce_grads(self):
        """Data-parallel average of the flat gradient buffers. With overlap the chunk all-reduces were
        launched from the autograd hooks during backward; here they are only waited for (and any chunk
        whose parameters received no gradient this step is reduced now)."""
        if not self._distributed:
            return
        inv = 1.0 / dist.get_world_size()
        if self._overlap:
            done = set()
            for w, g, a, b in self._works:
                w.wait()
                done.add((id(g), a))
            for g in self.groups:
                for (a, b, _) in g.chunks:
                    if (id(g), a) not in done:
                        dist.all_reduce(g.flat_g[a:b], op=dist.ReduceOp.SUM)
            self._works, self._pending = [], {}
            self._seen.clear()
            for g in self.groups:
                g.flat_g.mul_(inv)
            return
        works = [dist.all_reduce(g.flat_g, op=dist.ReduceOp.SUM, async_op=True) for g in self.groups]
        for w, g in zip(works, self.groups):
            w.wait()
            g.flat_g.mul_(inv)

    def step(self):
        lib = _c.load()
        acc = torch.zeros(1, dtype=torch.float64, device=self.groups[0].flat_g.device)
        for g in self.groups:                # global gradient norm: one streaming pass per flat buffer
            with torch.cuda.device(g.flat_g.device):
                _c.check(lib.mb200_sumsq_accum(_c.ptr(g.flat_g), g.n, _c.ptr(acc), _c.cur_stream()), "mb200_sumsq_accum")
        total = acc.sqrt().float()
        first = 1 if self.steps == 0 else 0
        for g in self.groups:
            with torch.cuda.device(g.flat_p.device):
                rc = lib.mb200_sgd_momentum_clip(_c.ptr(g.flat_p), _c.ptr(g.flat_g), _c.ptr(g.flat_m), g.n, float(g.lr),
                                                 float(self.momentum), float(self.weight_decay), _c.ptr(total),
                                                 float(self.max_norm), first, 1, _c.cur_stream())
            _c.check(rc, "mb200_sgd_momentum_clip")
        self.steps += 1
        for g in self.groups:                # gradients are zero again: re-open the direct-write window
            for p in g.params:
                p._mb200_direct.reset()
        tc_ops.bump_weight_epoch()       # raw-pointer update: invalidate the bf16 split caches
        return total
