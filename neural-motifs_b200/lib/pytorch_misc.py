"""Index / packing helpers used inside forward — same names as the reference's lib/pytorch_misc.py
(`enumerate_by_image` :278, `transpose_packed_sequence_inds` :365, `to_onehot` :110, `arange` :103,
`diagonal_inds` :301, `gather_nd` :255, `random_choose` :347, `Flattener` :80, `clip_grad_norm` :416,
`optimistic_restore` :14).  Checkpoint / printing helpers of that file are caller-side."""
import numpy as np
import torch
from torch import nn


class Flattener(nn.Module):
    """[N, ...] -> [N, -1] (pytorch_misc.py:80-87)."""

    def forward(self, x):
        return x.reshape(x.size(0), -1)


def arange(base_tensor, n=None):
    """LongTensor 0..n-1 on base_tensor's device (pytorch_misc.py:103-107)."""
    return torch.arange(base_tensor.size(0) if n is None else n, device=base_tensor.device, dtype=torch.long)


def to_onehot(vec, num_classes, fill=1000):
    """[N] labels -> [N,num_classes] with +fill at the label and -fill elsewhere (pytorch_misc.py:110-125)."""
    out = torch.full((vec.size(0), num_classes), -float(fill), device=vec.device, dtype=torch.float32)
    out.scatter_(1, vec.view(-1, 1).long(), float(fill))
    return out


def gather_nd(x, index):
    """x [x0,...,x{n-1},dim], index [num,n] -> [num,dim] (pytorch_misc.py:255-275)."""
    nd = x.dim() - 1
    assert nd > 0 and index.dim() == 2 and index.size(1) == nd
    dim = x.size(-1)
    sel = index[:, nd - 1].clone()
    mult = x.size(nd - 1)
    for col in range(nd - 2, -1, -1):
        sel += index[:, col] * mult
        mult *= x.size(col)
    return x.reshape(-1, dim)[sel]


def to_device_async(a, device, dtype=None):
    """numpy array / list / CPU tensor -> tensor on `device` WITHOUT stalling the host. `torch.as_tensor(a, device=cuda)`
    copies from pageable memory: the call is host-synchronous and stream-ordered, i.e. the host waits until everything
    already queued on the stream (the whole backbone) has run — round 2's GPU trace showed the forward after the backbone
    host-bound for that reason alone (3.6 ms of idle in 6 ms). Here the data goes through a pinned staging tensor
    (PyTorch's caching host allocator) and an asynchronous copy."""
    t = torch.as_tensor(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    device = torch.device(device)
    if device.type != "cuda":
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)


def image_segments(im_inds, host=None):
    """Runs of equal image index as a python list [(image, start, end)]. One small D2H read of the
    [N] index vector (the reference does the same, pytorch_misc.py:279) unless the caller already holds
    its host copy (`host`, numpy): a D2H in the middle of forward drains the stream and the GPU then
    idles while the host queues the next kernels."""
    a = np.asarray(host) if host is not None else im_inds.detach().cpu().numpy()
    if a.shape[0] == 0:
        return []
    cuts = np.flatnonzero(a[1:] != a[:-1]) + 1
    starts = np.concatenate(([0], cuts))
    ends = np.concatenate((cuts, [a.shape[0]]))
    return [(int(a[s]), int(s), int(e)) for s, e in zip(starts, ends)]


def enumerate_by_image(im_inds):
    """Generator form, as the reference's (pytorch_misc.py:278-287)."""
    for t in image_segments(im_inds):
        yield t


def diagonal_inds(tensor):
    """Flat indices of the diagonal of the first two dims (pytorch_misc.py:301-312)."""
    assert tensor.dim() >= 2 and tensor.size(0) == tensor.size(1)
    size = tensor.size(0)
    return (size + 1) * torch.arange(size, device=tensor.device, dtype=torch.long)


def random_choose(tensor, num, rng=np.random):
    """Random subset of rows without replacement (pytorch_misc.py:347-362); `rng` is injectable so
    parity runs can draw the same indices as the oracle."""
    num_choose = min(tensor.size(0), num)
    if num_choose == tensor.size(0):
        return tensor
    rand_idx = rng.choice(tensor.size(0), size=num, replace=False)
    rand_idx = to_device_async(rand_idx, tensor.device, torch.long)
    return tensor[rand_idx].contiguous()


def transpose_packed_sequence_inds(lengths):
    """Image-major -> time-major gather indices and per-step batch sizes for descending `lengths`
    (pytorch_misc.py:365-384), vectorised."""
    lengths = np.asarray(lengths, dtype=np.int64)
    starts = np.concatenate(([0], np.cumsum(lengths)[:-1]))
    T = int(lengths[0]) if len(lengths) else 0
    t = np.arange(T)[:, None]
    alive = lengths[None, :] > t                      # [T, B]
    inds = (starts[None, :] + t)[alive]               # row-major = time-major order
    return inds, alive.sum(1).tolist()


def clip_grad_norm(named_parameters, max_norm, clip=False, verbose=False):
    """Global-norm gradient clipping (pytorch_misc.py:416-459) with ONE device reduction instead
    of one host sync per parameter."""
    params = [p for _, p in named_parameters if p.grad is not None]
    if not params:
        return 0.0
    norms = torch._foreach_norm([p.grad for p in params], 2)
    total = torch.linalg.vector_norm(torch.stack(norms), 2)
    clip_coef = max_norm / (total + 1e-6)
    if clip:
        coef = torch.clamp(clip_coef, max=1.0)
        torch._foreach_mul_([p.grad for p in params], coef)
    return total


def optimistic_restore(network, state_dict):
    """Size-matched partial state-dict load (pytorch_misc.py:14-33). Returns True when every key matched."""
    own = network.state_dict()
    mismatch = False
    for name, param in state_dict.items():
        if name not in own:
            mismatch = True
            continue
        if param.size() == own[name].size():
            own[name].copy_(param)
        else:
            mismatch = True
    missing = set(own.keys()) - set(state_dict.keys())
    return not (mismatch or len(missing) > 0)
