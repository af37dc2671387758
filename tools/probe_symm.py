"""Probe (run under torchrun, one rank per GPU): torch symmetric memory over NVSwitch — rendezvous, peer-mapped buffers,
copy-engine P2P bandwidth (single peer and all peers at once), barrier. Prints one JSON line from rank 0."""
import json, os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
out = {"world": world}
try:
    n = 128 << 20                                    # 512 MB of fp32
    stage = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(stage, dist.group.WORLD)
    out["rendezvous"] = "ok"
    out["multicast"] = bool(getattr(hdl, "has_multicast_support", False))
    src = torch.full((n,), float(rank + 1), device=dev)
    peers = [(rank + k) % world for k in range(1, world)]
    bufs = {p: hdl.get_buffer(p, (n,), torch.float32) for p in peers}
    hdl.barrier()
    def timed(fn, reps=5):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]
    # one peer, whole buffer
    ms = timed(lambda: bufs[peers[0]].copy_(src, non_blocking=True))
    out["one_peer_GBs"] = n * 4 / ms / 1e6
    # every peer at once, 1/world of the buffer each, one stream per peer
    streams = [torch.cuda.Stream() for _ in peers]
    shard = n // world
    def fan():
        cur = torch.cuda.current_stream()
        for s, p in zip(streams, peers):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                bufs[p][rank * shard:(rank + 1) * shard].copy_(src[p * shard:(p + 1) * shard], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
    ms = timed(fan)
    out["all_peers_out_GBs"] = shard * 4 * len(peers) / ms / 1e6
    hdl.barrier()
    torch.cuda.synchronize()
    ok = all(float(stage[p * shard]) == p + 1 and float(stage[(p + 1) * shard - 1]) == p + 1 for p in peers)
    out["data_ok"] = bool(ok)
    # does a P2P copy slow a concurrent GEMM? (copy engines: it should not)
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    def gemms():
        for _ in range(20): torch.matmul(a, b)
    base = timed(gemms, 3)
    side = torch.cuda.Stream()
    def gemms_with_copy():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(4): bufs[peers[0]].copy_(src, non_blocking=True)
        gemms()
        torch.cuda.current_stream().wait_stream(side)
    both = timed(gemms_with_copy, 3)
    out["gemm_ms_alone"] = base; out["gemm_plus_2GB_p2p_ms"] = both
    t0 = time.time()
    for _ in range(20): hdl.barrier()
    torch.cuda.synchronize(); out["barrier_us"] = (time.time() - t0) / 20 * 1e6
except Exception as e:                                # noqa
    out["error"] = repr(e)[:500]
allo = [None] * world
dist.all_gather_object(allo, out)
if rank == 0:
    print(json.dumps(allo[0])); 
    if any("error" in o for o in allo): print(json.dumps(allo), file=sys.stderr)
dist.destroy_process_group()
