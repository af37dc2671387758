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

`defer_step=True` (opt-in): `step()` enqueues [exchange, norm, fused update] on a side stream and returns at once; the
next forward's frozen backbone (5 of 17 ms) is queued meanwhile and `wait_pending_updates()` — called by RelModel /
ObjectDetector.forward before the first trainable parameter is read — joins the two streams. The data movement of the
exchange hides underneath the backbone; the update KERNELS do not (a tcgen05 GEMM CTA needs an SM to itself: the GEMMs wait
for them, ~1 ms at one GPU, 1/W of that sharded), see DESIGN.md section 4.

Data parallel over NVSwitch: a SHARDED update in peer memory instead of an all-reduce. Gradients, parameters, their bf16
operand pairs and a landing area live in ONE symmetric allocation per optimizer (torch's symmetric memory does the plumbing:
cuMem allocation, peer + multicast mapping, device-side barrier); each rank owns a contiguous shard of every flat buffer.
`comm="ce"` (default on CUDA when the rendezvous succeeds) — transport by COPY ENGINE, the SMs only see local, shard-sized work:
  1  my copy of rank q's gradient shard -> my slot of q's landing area (cudaMemcpyAsync into the peer mapping)   | barrier
  2  own shard += the W-1 landed copies, and its squared norm (one kernel); the W partial norms are exchanged     | barrier
  3  fused clip + SGD on the own shard only (1/W of the update's HBM traffic; momentum exists for the own shard only)
  4  the updated shard (parameters + operand pairs) -> the same place in every other arena, copy engines          | barrier
`comm="nvls"` — the same with the transport inside our kernels (csrc/optim.cu): `multimem.ld_reduce.add` of the own shard (the
switch returns the sum over all ranks) in step 2, `multimem.st` (one store lands in every rank's copy) in step 3.
Why not NCCL (`comm="nccl"`, kept as the fallback and for the gloo CPU tests): its 32 channel CTAs cannot share an SM with a
200 KB / 54 K-register tcgen05 GEMM CTA, so an all-reduce underneath the backbone takes SMs away from the persistent GEMMs
(2 GPUs: 18.3 ms/step against 17.8 with comm="ce"; profiles/r02_bench_n2_*.json), and fewer channels cannot carry 1.1 GB in
time (profiles/r02_nccl_channels.json). Parameters are bit-identical on all ranks by construction (one owner per element).
tools/check_dp_sharded.py checks both modes against clip + torch SGD on NCCL-averaged gradients on 2 and 4 GPUs."""
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


def flat_size(params):
    return sum((p.numel() + 3) // 4 * 4 for p in params)


class SymmArena(object):
    """One symmetric (peer-mapped + multicast-mapped) allocation, carved into typed flat buffers. torch's symmetric
    memory does the plumbing: cuMem allocation, handle exchange, multicast binding, device-side barrier."""

    def __init__(self, nbytes, device):
        import torch.distributed._symmetric_memory as symm
        self.buf = symm.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.buf.zero_()
        self.base = self.buf.data_ptr()
        self.mc_base = int(self.hdl.multicast_ptr) if getattr(self.hdl, "has_multicast_support", False) else 0
        rank, world = dist.get_rank(), dist.get_world_size()
        # the other ranks' arenas, mapped into this process (same layout: a view here has the same offset there)
        self.peers = {q: self.hdl.get_buffer(q, (int(nbytes),), torch.uint8) for q in range(world) if q != rank}
        self.used = 0

    def peer_view(self, q, t):
        """The tensor occupying, in rank q's arena, the bytes `t` occupies in this one."""
        off = t.data_ptr() - self.base
        return self.peers[q][off:off + t.numel() * t.element_size()].view(t.dtype)

    def take(self, n, dtype):
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        t = self.buf[self.used:self.used + nbytes].view(dtype)
        self.used += (nbytes + 255) // 256 * 256
        return t

    def mc(self, t):
        """Multicast address of (the start of) a view of the arena."""
        import ctypes
        if not self.mc_base:
            raise RuntimeError("symmetric memory without multicast support")
        return ctypes.c_void_p(self.mc_base + (t.data_ptr() - self.base))

    def barrier(self):
        self.hdl.barrier(channel=0)


class FlatGroup(object):
    def __init__(self, params, chunk_bytes=128 << 20, arena=None):
        self.params = params
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4            # keep every view 16-byte aligned
        self.n = n
        self.offs = offs
        if arena is None:
            self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
            self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        else:
            self.flat_p = arena.take(n, torch.float32)
            self.flat_g = arena.take(n, torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.shard = (0, n)                             # [lo, hi) of the flat buffers this rank updates
        self.per_rank = n
        self.flat_hi = self.flat_lo = None          # bf16 pairs of the parameters, written by the fused update (presplit)
        self.touched = [False] * len(params)            # parameters the update covers (agreed across ranks)
        self.local_touched = [False] * len(params)      # ... that THIS rank's autograd has ever reached
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

    def shard_of(self, rank):
        return (min(self.n, rank * self.per_rank), min(self.n, (rank + 1) * self.per_rank))

    def touched_ranges(self, want=True):
        """Contiguous [a, b) runs of the flat buffer covering the parameters that have ever had a gradient
        (want=False: the complement)."""
        runs, a = [], None
        for i, t in enumerate(self.touched):
            t = (t == want)
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
                 chunk_bytes=128 << 20, defer_step=False, presplit=True, comm="auto"):
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
        import os as _os
        self._device = pgs[0]['params'][0].device
        self._distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self._presplit = bool(presplit) and self._device.type == "cuda"
        # ---- communication mode (module docstring): "ce" / "nvls" = sharded update over NVSwitch peer memory, "nccl" = all-reduce
        comm = _os.environ.get("MOTIFS_DP_COMM", comm)
        if comm not in ("auto", "ce", "nvls", "nccl"):
            raise ValueError("FlatSGD: comm must be 'auto', 'ce', 'nvls' or 'nccl'")
        self._arena = None
        if self._distributed and self._device.type == "cuda" and comm in ("auto", "ce", "nvls"):
            world = dist.get_world_size()
            sizes = [flat_size(pg['params']) for pg in pgs]
            per = 8 + (4 if self._presplit else 0) + 4                   # p + g (+ hi + lo) + staging, bytes per element
            try:
                self._arena = SymmArena(sum(n * per + 8 * 256 + 128 * world for n in sizes) + 8 * world + 256, self._device)
                if comm == "nvls" and not self._arena.mc_base:
                    raise RuntimeError("symmetric memory without multicast support")
            except Exception as e:                                       # noqa: no symmetric memory / multicast here
                if comm != "auto":
                    raise
                import warnings
                warnings.warn("FlatSGD: symmetric memory unavailable (%r); using the NCCL all-reduce" % (e,))
                self._arena = None
        self.comm = ("nvls" if comm == "nvls" else "ce") if self._arena is not None else ("nccl" if self._distributed else "none")
        self.groups = [FlatGroup(pg['params'], chunk_bytes, self._arena) for pg in self.param_groups]
        self.momentum, self.weight_decay, self.max_norm = momentum, weight_decay, max_norm
        self.steps = 0
        tc_ops.bump_weight_epoch()       # storages moved
        self._overlap = self._distributed and overlap_comm and self.comm == "nccl"
        self._sharded = self.comm in ("ce", "nvls")
        if self.comm in ("ce", "nvls"):
            world, rank = dist.get_world_size(), dist.get_rank()
            for g in self.groups:                                        # shard boundaries: multiples of 32 elements
                g.per_rank = ((g.n + world - 1) // world + 31) // 32 * 32
                g.shard = g.shard_of(rank)
                # landing area for the other ranks' copies of my shard (copy-engine reduce-scatter)
                g.stage = self._arena.take((world - 1) * g.per_rank, torch.float32) if self.comm == "ce" else None
            self._slots = self._arena.take(world, torch.float64)         # the ranks' partial squared norms
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
        self._sm_reserve = int(_os.environ.get("MOTIFS_NCCL_SM_RESERVE", "0")) if self._distributed else 0
        # presplit: the update kernel also writes the bf16 (hi, lo) pair of every updated parameter (+2 x 2 B per parameter);
        # weight matrices whose rows are a multiple of 64 long hand those views to lib/tc_ops as their GEMM operand
        if self._presplit:
            for g in self.groups:
                if self._arena is not None:
                    g.flat_hi = self._arena.take(g.n, torch.bfloat16)
                    g.flat_lo = self._arena.take(g.n, torch.bfloat16)
                else:
                    g.flat_hi = torch.zeros(g.n, device=self._device, dtype=torch.bfloat16)
                    g.flat_lo = torch.zeros(g.n, device=self._device, dtype=torch.bfloat16)
        if self._defer and _os.environ.get("MOTIFS_OPTIM_BACKGROUND", "0") == "1":
            # opt-in: launch the deferred kernels in the shape that fits BESIDE a tcgen05 GEMM CTA (csrc/optim.cu). Measured
            # (profiles/r02_trace_gaps_bg.log): the 1 ms hole in the compute stream closes, but the co-resident GEMMs slow
            # down by more than that (14.7 vs 12.8 ms busy) and the update itself takes 4x longer — 18.1 vs 17.1 ms/step.
            _c.load().mb200_optim_set_background(1)
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
        """comm="nccl", opt-in (MOTIFS_NCCL_SM_RESERVE=n, default 0): while chunk all-reduces are in flight NCCL's channel CTAs
        hold SMs, so the persistent tcgen05 kernels queued in that window get a grid that leaves n SMs free instead of running
        their last CTAs in a second wave. Measured at 2 GPUs it does not pay (18.85 ms/step without, 19.25 with n = 32): the
        host cannot know when the collective ends, so the narrower grid covers the whole backbone."""
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
        self._reduced = True
        if self._sharded:                # the exchange is part of the update (`_update_ce` / `_update_nvls`)
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
                    g.local_touched[i] = True
        if not self._distributed:
            for g in self.groups:
                g.touched = list(g.local_touched)
            return
        # Data parallel: WHICH parameters the update covers must be the same on every rank, or the replicas drift apart
        # (a parameter one rank's graph never reaches would be updated on the others only). The flags are agreed by a MAX
        # all-reduce — on steps 0-2, when the set is being discovered, and every 64th step after that (it costs a host
        # read-back, so not every step). A parameter first touched in between joins the update at the next agreement; its
        # gradient up to then is dropped on every rank alike (`_update` zeroes the ranges outside the agreed set).
        if self.steps < 3 or self.steps % 64 == 0:
            flags = torch.tensor([1 if t else 0 for g in self.groups for t in g.local_touched], dtype=torch.int32)
            flags = flags.to(self._device)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX)
            flags = flags.cpu().tolist()
            k = 0
            for g in self.groups:
                g.touched = [bool(f) for f in flags[k:k + len(g.params)]]
                k += len(g.params)

    def _shard_ranges(self, g, want=True):
        lo, hi = g.shard
        return [(max(a, lo), min(b, hi)) for a, b in g.touched_ranges(want) if min(b, hi) > max(a, lo)]

    def _update_nvls(self):
        """The sharded update over NVSwitch multicast (module docstring): barrier, reduce own shard + norm, barrier,
        clip + SGD on the own shard with multicast stores, barrier. All on the current stream."""
        lib, ar = _c.load(), self._arena
        world, rank = dist.get_world_size(), dist.get_rank()
        inv = 1.0 / world
        with torch.cuda.device(self._device):
            ar.barrier()                                   # A: every rank's gradients are final
            self._acc.zero_()
            for g in self.groups:
                lo, hi = g.shard
                if hi > lo:
                    sh = g.flat_g[lo:hi]
                    _c.check(lib.mb200_dp_reduce_shard_sumsq(ar.mc(sh), _c.ptr(sh), hi - lo, _c.ptr(self._acc), _c.cur_stream()),
                             "mb200_dp_reduce_shard_sumsq")
            _c.check(lib.mb200_dp_bcast_slot(_c.ptr(self._acc), ar.mc(self._slots), rank, _c.cur_stream()), "mb200_dp_bcast_slot")
            ar.barrier()                                   # B: all partial norms are in; nobody reads my gradients any more
            torch.sum(self._slots, dim=0, keepdim=True, out=self._acc)     # same order on every rank: identical clip factor
            torch.mul(self._acc.sqrt(), inv, out=self._acc)
            self._total.copy_(self._acc)                   # norm of the AVERAGED gradient
            first = 1 if self.steps == 0 else 0
            for g, pg in zip(self.groups, self.param_groups):
                lo, hi = g.shard
                for a, b in ((0, lo), (hi, g.n)):           # my copies of the other ranks' shards: zero for the next backward
                    if b > a:
                        _c.check(lib.mb200_zero_async(_c.ptr(g.flat_g[a:b]), (b - a) * 4, _c.cur_stream()), "mb200_zero_async")
                for a, b in self._shard_ranges(g):
                    rc = lib.mb200_sgd_momentum_clip_mc(
                        _c.ptr(g.flat_p[a:b]), _c.ptr(g.flat_g[a:b]), _c.ptr(g.flat_m[a:b]), ar.mc(g.flat_p[a:b]),
                        ar.mc(g.flat_hi[a:b]) if self._presplit else None, ar.mc(g.flat_lo[a:b]) if self._presplit else None,
                        b - a, float(pg['lr']), float(pg.get('momentum', self.momentum)),
                        float(pg.get('weight_decay', self.weight_decay)), _c.ptr(self._total), float(self.max_norm), float(inv),
                        first, 1, _c.cur_stream())
                    _c.check(rc, "mb200_sgd_momentum_clip_mc")
                for a, b in self._shard_ranges(g, want=False):     # outside the agreed set: nothing applied, nothing piles up
                    _c.check(lib.mb200_zero_async(_c.ptr(g.flat_g[a:b]), (b - a) * 4, _c.cur_stream()), "mb200_zero_async")
            ar.barrier()                                   # C: every rank's parameter stores have landed everywhere
        self._reduced = False

    def _update_ce(self):
        """Sharded update with COPY-ENGINE transport over NVLink (module docstring). The only kernels are local and
        shard-sized; every byte that crosses NVSwitch is moved by a DMA engine into peer-mapped symmetric memory."""
        lib, ar = _c.load(), self._arena
        world, rank = dist.get_world_size(), dist.get_rank()
        inv = 1.0 / world
        order = [(rank + k) % world for k in range(1, world)]       # staggered: at any moment every rank has one sender
        with torch.cuda.device(self._device):
            # 1. reduce-scatter: my copy of rank q's shard -> my slot of q's landing area
            for q in order:
                slot = (rank - q - 1) % world
                for g in self.groups:
                    lo, hi = g.shard_of(q)
                    if hi > lo:
                        dst = ar.peer_view(q, g.stage)[slot * g.per_rank: slot * g.per_rank + (hi - lo)]
                        dst.copy_(g.flat_g[lo:hi], non_blocking=True)
            ar.barrier()                                   # A: every copy of my shard has landed here
            self._acc.zero_()
            for g in self.groups:
                lo, hi = g.shard
                if hi > lo:
                    _c.check(lib.mb200_dp_reduce_staged_sumsq(_c.ptr(g.flat_g[lo:hi]), _c.ptr(g.stage), g.per_rank, world - 1,
                                                              hi - lo, _c.ptr(self._acc), _c.cur_stream()),
                             "mb200_dp_reduce_staged_sumsq")
            self._slots[rank:rank + 1].copy_(self._acc, non_blocking=True)
            for q in order:                                # the W partial squared norms, 8 bytes each
                ar.peer_view(q, self._slots)[rank:rank + 1].copy_(self._acc, non_blocking=True)
            ar.barrier()                                   # B: all partial norms are in
            torch.sum(self._slots, dim=0, keepdim=True, out=self._acc)     # same order on every rank: identical clip factor
            torch.mul(self._acc.sqrt(), inv, out=self._acc)
            self._total.copy_(self._acc)                   # norm of the AVERAGED gradient
            first = 1 if self.steps == 0 else 0
            for g, pg in zip(self.groups, self.param_groups):
                lo, hi = g.shard
                for a, b in ((0, lo), (hi, g.n)):           # my copies of the other ranks' shards: zero for the next backward
                    if b > a:
                        _c.check(lib.mb200_zero_async(_c.ptr(g.flat_g[a:b]), (b - a) * 4, _c.cur_stream()), "mb200_zero_async")
                for a, b in self._shard_ranges(g):
                    args = (b - a, float(pg['lr']), float(pg.get('momentum', self.momentum)),
                            float(pg.get('weight_decay', self.weight_decay)), _c.ptr(self._total), float(self.max_norm),
                            float(inv), first, 1, _c.cur_stream())
                    if self._presplit:
                        rc = lib.mb200_sgd_momentum_clip_split(_c.ptr(g.flat_p[a:b]), _c.ptr(g.flat_g[a:b]), _c.ptr(g.flat_m[a:b]),
                                                               _c.ptr(g.flat_hi[a:b]), _c.ptr(g.flat_lo[a:b]), *args)
                    else:
                        rc = lib.mb200_sgd_momentum_clip_scaled(_c.ptr(g.flat_p[a:b]), _c.ptr(g.flat_g[a:b]), _c.ptr(g.flat_m[a:b]), *args)
                    _c.check(rc, "mb200_sgd_momentum_clip")
                for a, b in self._shard_ranges(g, want=False):     # outside the agreed set: nothing applied, nothing piles up
                    _c.check(lib.mb200_zero_async(_c.ptr(g.flat_g[a:b]), (b - a) * 4, _c.cur_stream()), "mb200_zero_async")
            # 3. all-gather: my updated shard (parameters + operand pairs) -> the same place in every other arena
            for q in order:
                for g in self.groups:
                    lo, hi = g.shard
                    if hi > lo:
                        for buf in ((g.flat_p, g.flat_hi, g.flat_lo) if self._presplit else (g.flat_p,)):
                            ar.peer_view(q, buf)[lo:hi].copy_(buf[lo:hi], non_blocking=True)
            ar.barrier()                                   # C: every rank's shard has landed everywhere
        self._reduced = False

    def _update(self):
        if self.comm == "nvls":
            return self._update_nvls()
        if self.comm == "ce":
            return self._update_ce()
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
            if self._distributed:        # outside the agreed set: nothing is applied, nothing may pile up
                for a, b in g.touched_ranges(want=False):
                    g.flat_g[a:b].zero_()
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
        moms = [g.flat_m.detach().clone() for g in self.groups]
        if self._sharded:                # momentum exists for the own shard only: COLLECTIVE — every rank must call state_dict()
            for g, m in zip(self.groups, moms):
                m[:g.shard[0]].zero_(); m[g.shard[1]:].zero_()
                dist.all_reduce(m, op=dist.ReduceOp.SUM)
        return {"steps": self.steps,
                "param_groups": [{k: v for k, v in pg.items() if k != 'params'} for pg in self.param_groups],
                "momentum_buffers": moms,
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
            if self._sharded:            # keep the invariant state_dict() relies on: momentum is zero outside the own shard
                g.flat_m[:g.shard[0]].zero_(); g.flat_m[g.shard[1]:].zero_()
            g.touched = list(t)
            g.local_touched = list(t)
            pg.update(spg)
        self.steps = int(state["steps"])
