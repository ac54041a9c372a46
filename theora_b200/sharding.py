"""Multi-GPU plumbing: independent streams shard one-per-rank (weak scaling);
the only collective on the data path is the setup broadcast of the packets,
plus a MAX reduction of the per-rank elapsed time for reporting.  Works over
NCCL (one process per GPU) and over gloo (CPU tests)."""
import torch
import torch.distributed as dist


def stream_assignment(n_streams, world_size):
    """Round-robin: stream i runs on rank i % world_size."""
    return [[i for i in range(n_streams) if i % world_size == r] for r in range(world_size)]


def broadcast_bytes(blob, src=0, device=None):
    """One broadcast of the stream headers + packets from `src` to every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return blob
    dev = device if device is not None else torch.device("cpu")
    rank = dist.get_rank()
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    if rank == src:
        buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def max_over_ranks(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def quiet_barrier(tag):
    """Rendezvous that SLEEPS while it waits (key/value store of the process group, blocking socket wait).
    A NCCL barrier or all-reduce spins a host core per waiting rank; around a host-bound measurement (the
    e2e decode runs T threads per rank on shared cores) the early finishers would steal cores from the
    ranks still measuring.  Falls back to dist.barrier() if the store is not reachable."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    try:
        import datetime
        store = dist.distributed_c10d._get_default_store()
        store.set("%s_%d" % (tag, dist.get_rank()), "1")
        store.wait(["%s_%d" % (tag, r) for r in range(dist.get_world_size())], datetime.timedelta(seconds=1800))
    except Exception:
        dist.barrier()
