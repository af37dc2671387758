"""CPU (gloo, world_size 2): the gradient all-reduce used for the N>1 path averages bucketed gradients
and leaves every rank with identical values — host logic of lib/data_parallel.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    from lib.data_parallel import init_from_env, GradAllReducer
    r, w, _ = init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 300, 7, 1024)]
    frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    params[2].grad = None if rank == 1 else params[2].grad          # a rank without a grad contributes zeros
    red = GradAllReducer(params + [frozen], bucket_bytes=2048)     # forces several buckets
    assert len(red.buckets) >= 2
    red.all_reduce()
    out[rank] = [p.grad.clone() for p in params]
    dist.destroy_process_group()


def test_grad_all_reduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g0, g1 = out[0], out[1]
    for i, (a, b) in enumerate(zip(g0, g1)):
        assert torch.equal(a, b)
        expect = (1 + 2) / 2 * (i + 1) if i != 2 else 1 * (i + 1) / 2
        assert torch.allclose(a, torch.full_like(a, expect))


def _worker_flat(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    from lib.data_parallel import init_from_env
    from lib.fused_optim import FlatSGD
    init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 32), torch.nn.ReLU(),
                              torch.nn.Linear(32, 4))
    unused = torch.nn.Parameter(torch.zeros(10))            # never receives a gradient
    params = list(net.parameters()) + [unused]
    opt = FlatSGD([(params[:2], 0.1), (params[2:], 0.01)], overlap_comm=True, chunk_bytes=4096)
    assert sum(len(g.chunks) for g in opt.groups) >= 3
    opt.zero_grad()
    torch.manual_seed(100 + rank)                           # different data per rank
    x = torch.randn(8, 16)
    net(x).pow(2).mean().backward()                         # hooks launch the chunk all-reduces
    local = [p.grad.clone() for p in params]                # (already being reduced in place)
    opt.all_reduce_grads()
    out[rank] = [p.grad.clone() for p in params]
    # reference: recompute this rank's own gradient without hooks
    ref = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 32), torch.nn.ReLU(),
                              torch.nn.Linear(32, 4))
    ref.load_state_dict(net.state_dict())
    ref(x).pow(2).mean().backward()
    out[10 + rank] = [p.grad.clone() for p in ref.parameters()]
    dist.destroy_process_group()


def test_flat_sgd_overlapped_all_reduce_gloo_world2():
    """lib/fused_optim.FlatSGD: per-chunk all-reduce launched from autograd hooks == plain SUM over ranks (the
    1/world factor is folded into the fused update kernel)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_flat, args=(2, _free_port(), out), nprocs=2, join=True)
    for i in range(6):
        avg = out[10][i] + out[11][i]
        assert torch.allclose(out[0][i], avg, atol=1e-7) and torch.allclose(out[1][i], avg, atol=1e-7)
    assert float(out[0][6].abs().max()) == 0.0


def _worker_flat_direct(rank, world, port, out):
    """FlatSGD + direct gradient writes (tc_ops.direct_grad_target): weight gradients written straight into the
    flat buffer by the producing op — whole parameters and slices of a flat parameter — while the remaining
    gradients arrive through autograd. The stand-in ops below do on the CPU what _LinearTC / _MatmulTC /
    _HighwayLayerFunction do with the GEMM's `out=`; the bookkeeping under test is the host logic."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    from lib.data_parallel import init_from_env
    from lib.fused_optim import FlatSGD
    from lib import tc_ops
    init_from_env("gloo")

    class DirectLinear(torch.autograd.Function):           # y = x @ w^T, w a whole parameter
        @staticmethod
        def forward(ctx, x, w):
            ctx.save_for_backward(x, w)
            return x @ w.t()

        @staticmethod
        def backward(ctx, gy):
            x, w = ctx.saved_tensors
            tgt = tc_ops.direct_grad_target(w)
            gw = gy.t() @ x
            if tgt is not None:
                tgt.copy_(gw); gw = None
            return gy @ w, gw

    class DirectSlice(torch.autograd.Function):            # y = x @ v, v a contiguous slice of the flat `base`
        @staticmethod
        def forward(ctx, x, v, base):
            ctx.save_for_backward(x, v); ctx.base = base
            return x @ v

        @staticmethod
        def backward(ctx, gy):
            x, v = ctx.saved_tensors
            tgt = tc_ops.direct_grad_target(ctx.base, v)
            gv = x.t() @ gy
            if tgt is not None:
                tgt.copy_(gv); gv = None
            return gy @ v.t(), gv, None

    def build():
        torch.manual_seed(0)
        return [torch.nn.Parameter(torch.randn(24, 16)), torch.nn.Parameter(torch.randn(16 * 8 + 8 * 8 + 8)),
                torch.nn.Parameter(torch.randn(24, 24))]

    def run(ps, x, direct):
        w1, flat, w2 = ps
        a, b, bias = flat[:128].view(16, 8), flat[128:192].view(8, 8), flat[192:]
        lin = DirectLinear.apply if direct else (lambda t, w: t @ w.t())
        sl = (lambda t, v: DirectSlice.apply(t, v, flat)) if direct else (lambda t, v: t @ v)
        h = torch.tanh(lin(x, w1))                                          # [8,24]
        h = lin(h, w2)                                                      # w2: its only use -> whole-param direct write
        h2 = torch.tanh(sl(h[:, :16], a))                                   # slice 1 direct
        h3 = sl(h2, b) + bias                                               # slice 2 direct, bias through autograd
        y = lin(h3 @ torch.ones(8, 16), w1)                                 # w1 used twice: second use falls back to autograd
        return y.pow(2).mean() + h.pow(2).mean()

    ps = build()
    opt = FlatSGD([(ps[:1], 0.1), (ps[1:], 0.01)], overlap_comm=True, chunk_bytes=1024)
    opt.zero_grad()
    torch.manual_seed(100 + rank)
    x = torch.randn(8, 16)
    run(ps, x, True).backward()
    assert ps[2]._mb200_direct.written == {"all"}           # (torch still runs its AccumulateGrad hooks, with no gradient)
    assert len(ps[1]._mb200_direct.written) == 2 and ps[1]._mb200_direct.dirty            # bias came through autograd
    opt.all_reduce_grads()
    out[rank] = [p.grad.clone() for p in ps]
    ref = build()
    run(ref, x, False).backward()
    out[10 + rank] = [p.grad.clone() for p in ref]
    # a second backward before any step must accumulate (the direct-write window is closed)
    g_before = [p.grad.clone() for p in ps]
    opt._overlap = False                                  # no asynchronous all-reduce while the deltas are read
    run(ps, x, True).backward()
    out[20 + rank] = [(p.grad - g).clone() for p, g in zip(ps, g_before)]
    dist.barrier()
    dist.destroy_process_group()


def test_flat_sgd_direct_gradient_writes_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_flat_direct, args=(2, _free_port(), out), nprocs=2, join=True)
    for i in range(3):
        avg = out[10][i] + out[11][i]
        assert torch.allclose(out[0][i], avg, atol=1e-6), i
        assert torch.allclose(out[1][i], avg, atol=1e-6), i
        for r in (0, 1):        # second backward added this rank's own gradient on top (no overwrite)
            assert torch.allclose(out[20 + r][i], out[10 + r][i], rtol=1e-4, atol=1e-5), (r, i)   # (sum + own) - sum in fp32


def _worker_flat_uneven(rank, world, port, out):
    """Ranks whose autograd graphs differ: rank 1 never uses the middle layer (its chunk gets no gradient there), and
    in a second step rank 0 skips backward altogether. The chunk all-reduces must still pair up (fixed order)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    from lib.data_parallel import init_from_env
    from lib.fused_optim import FlatSGD
    init_from_env("gloo")
    torch.manual_seed(0)
    l1, l2, l3 = torch.nn.Linear(16, 16), torch.nn.Linear(16, 16), torch.nn.Linear(16, 4)
    params = list(l1.parameters()) + list(l2.parameters()) + list(l3.parameters())
    opt = FlatSGD([(params, 0.1)], overlap_comm=True, chunk_bytes=1100)     # one chunk per layer
    assert len(opt.groups[0].chunks) == 3
    opt.zero_grad()
    torch.manual_seed(100 + rank)
    x = torch.randn(8, 16)
    h = l1(x)
    h = l2(h) if rank == 0 else h
    l3(h).pow(2).mean().backward()
    opt.all_reduce_grads()
    out[rank] = [p.grad.clone() for p in params]
    for g in opt.groups:
        g.flat_g.zero_()
    for p in params:
        p._mb200_direct.reset()
    if rank == 1:                                   # step 2: rank 0 has nothing to differentiate
        l3(l2(l1(x))).pow(2).mean().backward()
    out[30 + rank] = [p.grad.clone() for p in params]
    opt.all_reduce_grads()
    out[20 + rank] = [p.grad.clone() for p in params]
    # a second backward once chunks are in flight is an error, not silent corruption (both ranks do the same, so
    # every collective still pairs up)
    l3(l2(l1(x))).pow(2).mean().backward()
    try:
        l3(l2(l1(x))).pow(2).mean().backward()
        raised = False
    except RuntimeError:
        raised = True
    opt.all_reduce_grads()
    out["raised%d" % rank] = raised
    dist.barrier()
    dist.destroy_process_group()


def test_flat_sgd_ranks_with_different_graphs_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_flat_uneven, args=(2, _free_port(), out), nprocs=2, join=True)
    for i in range(6):
        assert torch.equal(out[0][i], out[1][i]) and torch.equal(out[20][i], out[21][i])
        assert torch.allclose(out[20][i], out[31][i], atol=1e-7)      # only rank 1 contributed in step 2
    assert float(out[0][2].abs().max()) > 0                          # layer 2: rank 0's gradient alone
    assert out["raised0"] is True and out["raised1"] is True


def _worker_weighted(rank, world, port, out):
    """Ranks holding different numbers of objects / relations: the count-weighted loss + SUM all-reduce + 1/world must
    reproduce the gradient of ONE mean over the concatenated elements (the reference gathers to GPU 0 and averages)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    from lib.data_parallel import init_from_env, count_weighted_loss
    import torch.nn.functional as F
    init_from_env("gloo")
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 7))
    n_obj, n_rel = (3, 11) if rank == 0 else (9, 2)
    g = torch.Generator().manual_seed(10 + rank)
    xo, yo = torch.randn(n_obj, 7, generator=g), torch.randint(0, 5, (n_obj,), generator=g)
    xr, yr = torch.randn(n_rel, 7, generator=g), torch.randint(0, 5, (n_rel,), generator=g)
    loss = count_weighted_loss([(F.cross_entropy(xo @ w.t(), yo, reduction='sum'), n_obj),
                                (F.cross_entropy(xr @ w.t(), yr, reduction='sum'), n_rel)])
    loss.backward()
    grad = w.grad.clone()
    dist.all_reduce(grad)
    out[rank] = grad / world
    out[10 + rank] = (xo, yo, xr, yr)
    dist.barrier()
    dist.destroy_process_group()


def test_count_weighted_loss_matches_gathered_mean_gloo_world2():
    import torch.nn.functional as F
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_weighted, args=(2, _free_port(), out), nprocs=2, join=True)
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 7))
    xo = torch.cat((out[10][0], out[11][0])); yo = torch.cat((out[10][1], out[11][1]))
    xr = torch.cat((out[10][2], out[11][2])); yr = torch.cat((out[10][3], out[11][3]))
    (F.cross_entropy(xo @ w.t(), yo) + F.cross_entropy(xr @ w.t(), yr)).backward()
    assert torch.allclose(out[0], w.grad, atol=1e-6) and torch.allclose(out[1], w.grad, atol=1e-6)


def _worker_settle(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    import bench
    from lib.data_parallel import init_from_env
    init_from_env("gloo")
    import random
    rnd = random.Random(rank)
    steps = []

    def step(i):                                  # a step with a collective inside and rank-dependent timing noise
        time.sleep(0.004 + (0.004 * rnd.random() if (rank == 1 and i < 20) else 0.0))
        t = torch.ones(4)
        dist.all_reduce(t)
        steps.append(i)

    n = bench.settle_warmup(step, lambda: None, world, torch.device("cpu"), min_steps=6, max_steps=40, tol=1.5)
    out[rank] = (n, len(steps))
    dist.barrier()
    dist.destroy_process_group()


def test_bench_settle_warmup_leaves_collectively_gloo_world2():
    """bench.py's adaptive warm-up: ranks with different timing noise must run the SAME number of steps (each step
    holds the gradient all-reduce; a rank leaving early would pair its next collective with the others' all-reduce)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_settle, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] == out[1] and 6 <= out[0][0] <= 40


def _worker_touched(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    from lib.data_parallel import init_from_env
    from lib.fused_optim import FlatSGD
    init_from_env("gloo")
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(n)) for n in (40, 7, 300, 12)]
    opt = FlatSGD([(ps[:2], 0.01), (ps[2:], 0.1)], comm="nccl")
    # rank 0's graph reaches parameters 0 and 2, rank 1's reaches 0 and 1; nobody reaches 3
    used = [0, 2] if rank == 0 else [0, 1]
    sum((ps[i] * 2.0).sum() for i in used).backward()
    opt._mark_touched()                                    # what step() does first; the fused kernels need a GPU
    out[rank] = ([list(g.local_touched) for g in opt.groups], [list(g.touched) for g in opt.groups],
                 [g.touched_ranges() for g in opt.groups], [g.touched_ranges(want=False) for g in opt.groups])
    dist.destroy_process_group()


def test_touched_parameter_set_is_agreed_across_ranks():
    """Which parameters the fused update covers must not depend on the rank (lib/fused_optim.FlatSGD._mark_touched): the
    union of what any rank's autograd has reached, the same ranges everywhere; what nobody reached stays out (torch's
    `grad is None` skip)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_touched, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0][0] == [[True, False], [True, False]] and out[1][0] == [[True, True], [False, False]]     # local
    assert out[0][1] == out[1][1] == [[True, True], [True, False]]                                          # agreed
    assert out[0][2] == out[1][2] == [[(0, 48)], [(0, 300)]]          # 40 + 7 padded to multiples of 4; 300
    assert out[0][3] == out[1][3] == [[], [(300, 312)]]


def test_shard_layout_partitions_every_flat_buffer():
    """The sharded update (comm="ce" / "nvls"): rank r owns [r * per_rank, (r + 1) * per_rank) of each flat buffer, in
    multiples of 32 elements; the shards are disjoint, cover the buffer, and a parameter range is split across owners
    exactly at those boundaries (FlatGroup.shard_of, FlatSGD._shard_ranges)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-motifs_b200"))
    from lib.fused_optim import FlatGroup, FlatSGD
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (1000, 37, 4096, 5, 130)]
    for world in (2, 3, 4, 8):
        g = FlatGroup(ps)
        g.per_rank = ((g.n + world - 1) // world + 31) // 32 * 32
        shards = [g.shard_of(r) for r in range(world)]
        assert shards[0][0] == 0 and shards[-1][1] == g.n
        assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))                        # contiguous, disjoint
        assert all(lo % 32 == 0 and (hi % 4 == 0) for lo, hi in shards)
        g.touched = [True, False, True, True, False]
        covered = []
        for r in range(world):
            g.shard = shards[r]
            covered += FlatSGD._shard_ranges(None, g)
            assert all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in FlatSGD._shard_ranges(None, g, want=False))
        covered.sort()
        merged = [list(covered[0])]
        for a, b in covered[1:]:
            if a == merged[-1][1]:
                merged[-1][1] = b
            else:
                merged.append([a, b])
        assert [tuple(m) for m in merged] == g.touched_ranges()                             # the union is the touched set
