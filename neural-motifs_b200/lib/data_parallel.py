"""Data parallelism for the relation model: one process per GPU (torch.distributed, NCCL over
NVLink 5 / NVSwitch), images sharded per rank, parameters resident per rank, ONE gradient
all-reduce per step — replaces the reference's single-process replicate / parallel_apply / Gather
(lib/rel_model.py:549-560, lib/object_detector.py:40-47) which re-broadcasts every parameter each
step and funnels all results through GPU 0. The frozen detector needs no communication."""
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK, WORLD_SIZE, MASTER_*).
    Returns (rank, world_size, local_rank); a no-op single-process setup when WORLD_SIZE is unset."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # NCCL's channel count is left at its default (32 over NVSwitch): the 1.1 GB gradient all-reduce is bandwidth
            # bound and capping it at 8 / 4 channels cost 2.3 / 8.9 ms per step at N=2 (profiles/r02_nccl_channels.json)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


class GradAllReducer(object):
    """Averages the gradients of the trainable parameters over all ranks. Gradients are packed into
    a few large flat buckets (NVSwitch bandwidth is uniform: buckets are sized for launch latency,
    not link count) and reduced asynchronously; `wait()` unpacks them."""

    def __init__(self, params, bucket_bytes=256 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets, cur, size = [], [], 0
        for p in self.params:
            n = p.numel() * 4
            if cur and size + n > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            self.buckets.append(cur)
        self._pending = []

    def start(self):
        if self.world == 1:
            return
        self._pending = []
        for bucket in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            self._pending.append((bucket, flat, work))

    def wait(self):
        if self.world == 1:
            return
        inv = 1.0 / self.world
        for bucket, flat, work in self._pending:
            work.wait()
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p).mul_(inv)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self._pending = []

    def all_reduce(self):
        self.start()
        self.wait()


def count_weighted_loss(sums_and_counts):
    """Data-parallel loss whose gradient, once all-reduced and divided by the world size (lib/fused_optim.FlatSGD folds
    the 1/world into its update), equals the gradient of the reference's loss: there all GPUs' outputs are gathered on
    GPU 0 and ONE mean is taken over the concatenated objects / relations (models/train_rels.py:140-141 after
    `gather_res`, lib/object_detector.py:40-47), i.e. every element weighs 1/N_total — a per-rank mean would weigh the
    elements of a rank that holds fewer of them more. `sums_and_counts`: [(sum-reduced loss tensor, local element
    count), ...]; returns  sum_k  loss_sum_k * world / N_total_k  (one tiny all-reduce of the counts; with one rank, or
    equal counts everywhere, this is exactly the plain per-rank mean)."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return sum(s / max(int(n), 1) for s, n in sums_and_counts)
    dev = sums_and_counts[0][0].device
    # pinned + non_blocking: a pageable `torch.tensor(..., device=dev)` copy is stream-ordered AND host-synchronous, i.e. the
    # host would sit here until the whole forward has drained and only then start queueing the backward
    if dev.type == "cuda":
        import numpy as np
        from lib.pytorch_misc import to_device_async
        tot = to_device_async(np.array([float(n) for _, n in sums_and_counts], np.float32), dev)
    else:
        tot = torch.tensor([float(n) for _, n in sums_and_counts], device=dev, dtype=torch.float32)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    tot = tot.clamp_min(1.0)
    return sum(s * (world / tot[k]) for k, (s, _) in enumerate(sums_and_counts))
