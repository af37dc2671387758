"""Multi-GPU check of FlatSGD(comm="ce" | "nvls") (run under torchrun, one rank per GPU; CHECK_COMM selects the mode): the
sharded update over NVSwitch against clip_grad_norm_ + torch.optim.SGD on NCCL-all-reduced gradients.
  * parameters after every step == the torch reference (rtol 1e-5) and BIT-IDENTICAL on all ranks
  * the bf16 operand pairs the update multicasts == split_rows(param) bit for bit, on every rank
  * a parameter no rank's graph reaches is untouched; one that only rank 1 reaches is updated on both
  * the momentum gathered by state_dict() == the reference's momentum buffers
  * deferred (side stream) == immediate, bit for bit
Prints PASS / FAIL lines from rank 0; exit code 1 on any failure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import torch
import torch.distributed as dist
from lib.data_parallel import init_from_env
from lib import fused_optim, tc_ops
from lib.fused_optim import FlatSGD

MODE = os.environ.get("CHECK_COMM", "ce")        # "ce" (copy-engine transport) or "nvls" (multimem kernels)
rank, world, local = init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
fails = []

def check(name, ok, detail=""):
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(("PASS " if int(flag) else "FAIL ") + name + (" " + str(detail) if detail else ""), flush=True)
    if not int(flag):
        fails.append(name)

SHAPES = [(96, 128), (1000,), (64, 64), (33, 7), (300001,), (17,), (256, 192)]

def build():
    torch.manual_seed(0)                                    # identical on every rank
    return [torch.nn.Parameter(torch.randn(s, device=dev)) for s in SHAPES]

def losses(ps, step):
    g = torch.Generator(device=dev).manual_seed(1000 * step + rank)       # per-rank gradients
    scale = 3.0 if step == 1 else 0.02                                     # step 1 clips
    loss = 0
    for i, p in enumerate(ps):
        if i == 5:
            continue                                                       # never used by anyone
        if i == 3 and rank != 1 % world:
            continue                                                       # only rank 1's graph reaches it
        loss = loss + (p * torch.randn(p.shape, device=dev, generator=g) * scale).sum()
    return loss

def run(defer, steps=4):
    ps = build()
    opt = FlatSGD([{'params': ps[:2], 'lr': 0.01}, {'params': ps[2:]}], lr=0.1, momentum=0.9, weight_decay=1e-2, max_norm=5.0,
                  defer_step=defer, comm=MODE)
    for step in range(steps):
        fused_optim.wait_pending_updates()
        opt.zero_grad()
        losses(ps, step).backward()
        opt.all_reduce_grads()
        opt.step()
    fused_optim.wait_pending_updates()
    torch.cuda.synchronize()
    return ps, opt

ps, opt = run(False)
check("comm mode is " + MODE, opt.comm == MODE, opt.comm)

# ---- torch reference: all-reduced (averaged) gradients, clip, SGD
qs = build()
ref = torch.optim.SGD([{'params': qs[:2], 'lr': 0.01}, {'params': qs[2:]}], lr=0.1, momentum=0.9, weight_decay=1e-2)
for step in range(4):
    for q in qs:
        q.grad = None
    losses(qs, step).backward()
    for i, q in enumerate(qs):
        if i == 5:
            continue
        g = q.grad if q.grad is not None else torch.zeros_like(q)
        dist.all_reduce(g); q.grad = g / world
    # FlatSGD joins parameter 3 (reached by rank 1 only) at the first agreement of the touched set = step 0: same as here
    torch.nn.utils.clip_grad_norm_([q for q in qs if q.grad is not None], 5.0)
    ref.step()
err = max(float((p - q).abs().max() / (q.abs().max() + 1e-12)) for p, q in zip(ps, qs))
check("parameters == clip + torch SGD on averaged gradients", all(torch.allclose(p, q, rtol=1e-5, atol=1e-6) for p, q in zip(ps, qs)), "max rel %.2e" % err)
check("never-used parameter untouched", torch.equal(ps[5], build()[5]))
check("parameter reached by one rank only is updated everywhere", not torch.equal(ps[3], build()[3]))

# ---- bit-identical across ranks
flat = torch.cat([p.detach().reshape(-1) for p in ps])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
check("parameters bit-identical on all ranks", all(torch.equal(gathered[0], t) for t in gathered))

# ---- multicast operand pairs
ok = True
for i in (0, 2, 6):
    got = tc_ops.weight_split(ps[i]); want = tc_ops.split_rows(ps[i].detach())
    ok = ok and hasattr(ps[i], "_mb200_presplit") and got is ps[i]._mb200_presplit[1] and torch.equal(got.hi, want.hi) and torch.equal(got.lo, want.lo)
check("multicast bf16 operand pairs == split_rows(param)", ok)

# ---- gradients zero again, momentum gather
check("flat gradients are zero after the step", all(float(g.flat_g.abs().max()) == 0.0 for g in opt.groups))
sd = opt.state_dict()
want_m = [torch.cat([(ref.state[q]['momentum_buffer'] if q in ref.state and 'momentum_buffer' in ref.state[q] else torch.zeros_like(q)).reshape(-1)
                     for q in grp]) for grp in (qs[:2], qs[2:])]
okm = True
for g, m, w in zip(opt.groups, sd["momentum_buffers"], want_m):
    got = torch.cat([m[o:o + p.numel()] for p, o in zip(g.params, g.offs)])
    okm = okm and torch.allclose(got, w, rtol=1e-5, atol=1e-6)
check("state_dict() momentum == reference momentum", okm)

# ---- deferred == immediate
ps2, opt2 = run(True)
check("deferred update bit-identical to the immediate one", all(torch.equal(p, q) for p, q in zip(ps, ps2)))
check("total_norm agrees", abs(float(opt.total_norm()) - float(opt2.total_norm())) == 0.0, float(opt.total_norm()))
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
