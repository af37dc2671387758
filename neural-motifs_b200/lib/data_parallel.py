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
