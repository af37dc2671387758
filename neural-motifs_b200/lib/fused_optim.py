"""Flat-buffer SGD for the relation model's trainable parameters: same update rule as the caller's
`clip_grad_norm(..., max_norm=conf.clip)` + `optim.SGD(params, lr, momentum=0.9, weight_decay=conf.l2)`
(models/train_rels.py:57-70,145-150) — including the lr/10 group for the VGG fc layers — but every
parameter and gradient lives in ONE contiguous buffer per group, so that
  * the global gradient norm is one reduction,
  * the data-parallel gradient all-reduce runs on chunks of the flat gradient (no bucket copies),
  * clip + 1/world + weight decay + momentum + update + gradient zeroing is one fused kernel (csrc/optim.cu).
Parameters stay ordinary nn.Parameters (names / state dict unchanged); their storage is re-pointed.

`FlatSGD` IS a `torch.optim.Optimizer`: `param_groups[i]['lr']` is read on every step, so the caller's
`ReduceLROnPlateau(optimizer, ...)` and its early stop on `optimizer.param_groups[...]['lr']`
(models/train_rels.py:70,204-205) work; `state_dict()` / `load_state_dict()` carry the momentum buffers.

Parameters that have NEVER received a gradient are skipped (no weight decay, no momentum), as torch's SGD skips
`p.grad is None` — e.g. `context.decoder_rnn` in predcls. Once touched, a parameter is updated every step (with a
zero gradient if a step does not reach it): PyTorch 0.3's `zero_grad()` zeroes instead of dropping gradients, so that
is what the reference recipe does.

Data parallel (one process per GPU): chunk all-reduces are launched from autograd hooks during backward in a FIXED
chunk order on every rank (chunk k goes out only after every chunk before it in `self._order`), the fallback in
`all_reduce_grads()` uses the same order — ranks whose autograd graphs differ (a parameter unused on one rank, a rank
that skipped backward) still issue identical collective sequences. One backward per `all_reduce_grads()`; a second
backward after a chunk has gone out raises (it would add local gradients to an already averaged chunk).

`defer_step=True` (opt-in): `step()` enqueues [wait for the all-reduce, norm, fused update] on a side stream and
returns at once; the next forward's frozen backbone (9 of 21 ms) runs underneath it and `wait_pending_updates()`
— called by RelModel / ObjectDetector.forward before the first trainable parameter is read — joins the two
streams. The gradient all-reduce is then off the critical path entirely at any world size."""
import torch
import torch.distributed as dist

import motifs_cabi as _c
from lib import tc_ops

_PENDING = []        # optimizers with an update still in flight on their side stream


def wait_pending_updates():
    """Make the current stream wait for every deferred optimizer update (cheap no-op when there is none)."""
    while _PENDING:
        opt = _PENDING.pop()
        ev, opt._pending_ev = opt._pending_ev, None
        if ev is not None:
            torch.cuda.current_stream(opt._device).wait_event(ev)
        opt._reserve_sms(False)          # kernels queued from here on run after the collectives: full width again


class FlatGroup(object):
    def __init__(self, params, chunk_bytes=128 << 20):
        self.params = params
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4            # keep every view 16-byte aligned
        self.n = n
        self.offs = offs
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_hi = self.flat_lo = None          # bf16 pairs of the parameters, written by the fused update (presplit)
        self.touched = [False] * len(params)
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

    def touched_ranges(self):
        """Contiguous [a, b) runs of the flat buffer covering the parameters that have ever had a gradient."""
        runs, a = [], None
        for i, t in enumerate(self.touched):
            if t and a is None:
                a = self.offs[i]
            if not t and a is not None:
                runs.append((a, self.offs[i])); a = None
        if a is not None:
            runs.append((a, self.n))
        return runs


class FlatSGD(torch.optim.Optimizer):
    """groups: list of (params, lr) tuples, or torch-style dicts {'params': [...], 'lr': ...} (then `lr` is the
    default). momentum / weight_decay / max_norm shared (train_rels.py:66,145)."""

    def __init__(self, groups, lr=None, momentum=0.9, weight_decay=1e-4, max_norm=5.0, overlap_comm=True,
                 chunk_bytes=128 << 20, defer_step=False, presplit=True):
        pgs = []
        for g in groups:
            if isinstance(g, dict):
                ps, glr = list(g['params']), g.get('lr', lr)
            else:
                ps, glr = list(g[0]), g[1]
            ps = [p for p in ps if p.requires_grad]
            if not ps:
                continue
            if glr is None:
                raise ValueError("FlatSGD: a group has no learning rate")
            pgs.append({'params': ps, 'lr': float(glr)})
        super().__init__(pgs, dict(lr=0.0, momentum=momentum, weight_decay=weight_decay, max_norm=max_norm))
        self.groups = [FlatGroup(pg['params'], chunk_bytes) for pg in self.param_groups]
        self.momentum, self.weight_decay, self.max_norm = momentum, weight_decay, max_norm
        self.steps = 0
        self._device = self.groups[0].flat_p.device
        tc_ops.bump_weight_epoch()       # storages moved
        self._distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self._overlap = self._distributed and overlap_comm
        # fixed collective order: the groups last in `groups` first, inside a group the last chunk first —
        # roughly the order backward produces them (late layers first), identical on every rank by construction
        self._order = [(gi, ci) for gi in reversed(range(len(self.groups)))
                       for ci in reversed(range(len(self.groups[gi].chunks)))]
        self._pos = {k: i for i, k in enumerate(self._order)}
        self._reset_comm()
        self._defer = bool(defer_step) and self._device.type == "cuda"
        self._stream = torch.cuda.Stream(self._device) if self._defer else None
        self._pending_ev = None
        self._reduced = False
        self._reserved = False
        import os as _os
        self._sm_reserve = int(_os.environ.get("MOTIFS_NCCL_SM_RESERVE", "16")) if self._distributed else 0
        # presplit: the update kernel also writes the bf16 (hi, lo) pair of every updated parameter (+2 x 2 B per parameter);
        # weight matrices whose rows are a multiple of 64 long hand those views to lib/tc_ops as their GEMM operand
        self._presplit = bool(presplit) and self._device.type == "cuda"
        if self._presplit:
            for g in self.groups:
                g.flat_hi = torch.zeros(g.n, device=self._device, dtype=torch.bfloat16)
                g.flat_lo = torch.zeros(g.n, device=self._device, dtype=torch.bfloat16)
        self._acc = torch.zeros(1, dtype=torch.float64, device=self._device)
        self._total = torch.zeros(1, dtype=torch.float32, device=self._device)
        # One hook per parameter. (1) It marks the parameter's flat gradient as touched by autograd, which ends
        # the window in which weight-gradient GEMMs may write straight into it (tc_ops.direct_grad_target).
        # (2) With data-parallel overlap it counts down the parameter's communication chunk; a complete chunk is
        # launched (NCCL stream) as soon as every chunk before it in the fixed order has been.
        for gi, g in enumerate(self.groups):
            for ci, (a, b, ps) in enumerate(g.chunks):
                for p in ps:
                    p._mb200_direct = tc_ops.DirectGradState()
                    p.register_post_accumulate_grad_hook(self._make_autograd_hook(gi, ci))

    # ------------------------------------------------------------------ communication
    def _reset_comm(self):
        self._works = []
        self._left = {k: len(self.groups[k[0]].chunks[k[1]][2]) for k in self._order}
        self._seen = set()
        self._next = 0               # index into self._order of the next chunk to launch

    def _reserve_sms(self, on):
        """While chunk all-reduces are in flight NCCL's channel CTAs hold SMs: the persistent tcgen05 kernels queued in that
        window (rest of backward, the next step's backbone) get a grid that leaves `nccl_sm_reserve` SMs free, else their
        last CTAs would run in a second wave behind NCCL's (measured at 2 GPUs: +1.6 ms per step)."""
        if self._device.type != "cuda" or self._sm_reserve <= 0 or on == self._reserved:
            return
        _c.load().mb200_set_sm_budget(148 - self._sm_reserve if on else 148)
        self._reserved = on

    def _launch_ready(self, force=False):
        if self._next < len(self._order) and (force or self._left[self._order[self._next]] <= 0):
            self._reserve_sms(True)
        while self._next < len(self._order):
            gi, ci = self._order[self._next]
            if not force and self._left[(gi, ci)] > 0:
                break
            a, b, _ = self.groups[gi].chunks[ci]
            self._works.append(dist.all_reduce(self.groups[gi].flat_g[a:b], op=dist.ReduceOp.SUM, async_op=True))
            self._next += 1

    def _make_autograd_hook(self, gi, ci):
        def on_accumulate(param):
            param._mb200_direct.dirty = True
            if not self._overlap:
                return
            if id(param) in self._seen:
                if self._pos[(gi, ci)] < self._next:
                    raise RuntimeError("FlatSGD: a second backward reached a gradient chunk whose all-reduce is already "
                                       "in flight; call all_reduce_grads()/step() after every backward, or construct "
                                       "with overlap_comm=False for gradient accumulation")
                return
            self._seen.add(id(param))
            self._left[(gi, ci)] -= 1
            self._launch_ready()
        return on_accumulate

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed by the fused step itself; kept for API symmetry (never set to None:
        autograd accumulates into the flat views)."""
        if self.steps == 0:
            for g in self.groups:
                g.flat_g.zero_()
                for p in g.params:
                    p._mb200_direct.reset()

    def all_reduce_grads(self):
        """Data-parallel SUM of the flat gradient buffers (the 1/world factor is folded into the fused update).
        With overlap the chunk all-reduces were launched from the autograd hooks during backward; the rest go out
        here in the same fixed order. Without `defer_step` this also waits for them."""
        if not self._distributed:
            return
        if self._overlap:
            self._launch_ready(force=True)
        else:
            self._works = [dist.all_reduce(g.flat_g, op=dist.ReduceOp.SUM, async_op=True) for g in self.groups]
        self._reduced = True
        if not self._defer:
            self._wait_works()
            self._reserve_sms(False)

    def _wait_works(self):
        for w in self._works:
            w.wait()                 # NCCL: the current stream waits for the collective's stream; gloo: host wait
        self._reset_comm()

    # ------------------------------------------------------------------ update
    def _mark_touched(self):
        for g in self.groups:
            for i, p in enumerate(g.params):
                st = p._mb200_direct
                if st.dirty or st.written:
                    g.touched[i] = True

    def _update(self):
        lib = _c.load()
        world = dist.get_world_size() if (self._distributed and getattr(self, "_reduced", False)) else 1
        inv = 1.0 / world
        self._acc.zero_()
        for g in self.groups:                # global gradient norm: one streaming pass per flat buffer
            with torch.cuda.device(g.flat_g.device):
                _c.check(lib.mb200_sumsq_accum(_c.ptr(g.flat_g), g.n, _c.ptr(self._acc), _c.cur_stream()), "mb200_sumsq_accum")
        torch.mul(self._acc.sqrt(), inv, out=self._acc)
        self._total.copy_(self._acc)         # norm of the AVERAGED gradient
        first = 1 if self.steps == 0 else 0
        for g, pg in zip(self.groups, self.param_groups):
            lr = float(pg['lr'])
            for a, b in g.touched_ranges():
                with torch.cuda.device(g.flat_p.device):
                    if self._presplit:
                        rc = lib.mb200_sgd_momentum_clip_split(
                            _c.ptr(g.flat_p[a:b]), _c.ptr(g.flat_g[a:b]), _c.ptr(g.flat_m[a:b]), _c.ptr(g.flat_hi[a:b]),
                            _c.ptr(g.flat_lo[a:b]), b - a, lr, float(pg.get('momentum', self.momentum)),
                            float(pg.get('weight_decay', self.weight_decay)), _c.ptr(self._total), float(self.max_norm),
                            float(inv), first, 1, _c.cur_stream())
                    else:
                        rc = lib.mb200_sgd_momentum_clip_scaled(
                            _c.ptr(g.flat_p[a:b]), _c.ptr(g.flat_g[a:b]), _c.ptr(g.flat_m[a:b]), b - a, lr,
                            float(pg.get('momentum', self.momentum)), float(pg.get('weight_decay', self.weight_decay)),
                            _c.ptr(self._total), float(self.max_norm), float(inv), first, 1, _c.cur_stream())
                _c.check(rc, "mb200_sgd_momentum_clip")
        self._reduced = False

    @torch.no_grad()
    def step(self, closure=None):
        """One fused update. Returns the (pre-clip) global gradient norm as a device tensor; with `defer_step` the
        update is only enqueued and None is returned — `total_norm()` joins and returns it."""
        if closure is not None:
            raise ValueError("FlatSGD.step: closures are not supported")
        self._mark_touched()
        if self._defer:
            wait_pending_updates()                      # at most one update in flight
            self._stream.wait_stream(torch.cuda.current_stream(self._device))
            with torch.cuda.stream(self._stream):
                if self._distributed:
                    self._wait_works()
                self._update()
                self._pending_ev = torch.cuda.Event()
                self._pending_ev.record(self._stream)
            _PENDING.append(self)
        else:
            if self._distributed and self._works:
                self._wait_works()
            self._update()
        self.steps += 1
        for g in self.groups:                # gradients are zero again: re-open the direct-write window
            for p in g.params:
                p._mb200_direct.reset()
        tc_ops.bump_weight_epoch()       # raw-pointer update: invalidate the bf16 split caches
        if self._presplit:               # ... and hand over the splits the update kernel has just written
            for g in self.groups:
                for i, (p, o) in enumerate(zip(g.params, g.offs)):
                    if g.touched[i] and p.dim() == 2 and p.size(1) % 64 == 0:
                        n = p.numel()
                        tc_ops.preset_rows_split(p, g.flat_hi[o:o + n].view_as(p), g.flat_lo[o:o + n].view_as(p))
        return None if self._defer else self._total

    def total_norm(self):
        wait_pending_updates()
        return self._total

    # ------------------------------------------------------------------ checkpointing
    def state_dict(self):
        wait_pending_updates()
        if self._device.type == "cuda":
            torch.cuda.current_stream(self._device).synchronize()
        return {"steps": self.steps,
                "param_groups": [{k: v for k, v in pg.items() if k != 'params'} for pg in self.param_groups],
                "momentum_buffers": [g.flat_m.detach().clone() for g in self.groups],
                "touched": [list(g.touched) for g in self.groups]}

    def load_state_dict(self, state):
        wait_pending_updates()
        if len(state["momentum_buffers"]) != len(self.groups):
            raise ValueError("FlatSGD.load_state_dict: group count differs")
        for g, pg, m, t, spg in zip(self.groups, self.param_groups, state["momentum_buffers"], state["touched"],
                                    state["param_groups"]):
            if m.numel() != g.n or len(t) != len(g.params):
                raise ValueError("FlatSGD.load_state_dict: flat layout differs")
            g.flat_m.copy_(m)
            g.touched = list(t)
            pg.update(spg)
        self.steps = int(state["steps"])
