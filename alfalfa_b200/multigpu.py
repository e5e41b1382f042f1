"""Multi-GPU plumbing: the decode path shards at GOP / stream granularity -- a key
frame resets all codec state (decoder_state.hh:90) -- so every rank decodes its own GOPs on its own
GPU and there is NO data-path collective.  torch.distributed is used only for the barrier around
the timed region and for reducing the timing / work counters (NCCL on the GPU box, gloo in the CPU
tests)."""
import os
import struct


def rank_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """returns (rank, world, local_rank, dist or None)"""
    rank, world, local = rank_info()
    if world <= 1:
        return rank, world, local, None
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, world, local, dist


def _tensor(dist, local, x):
    import torch
    dev = "cuda:%d" % local if dist.get_backend() == "nccl" else "cpu"
    return torch.tensor([float(x)], dtype=torch.float64, device=dev)


def barrier(dist):
    if dist is not None:
        dist.barrier()


def reduce_max(dist, local, x):
    if dist is None:
        return x
    t = _tensor(dist, local, x)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, local, x):
    if dist is None:
        return x
    t = _tensor(dist, local, x)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(dist, local, units_this_rank, seconds_this_rank):
    """whole-job throughput = units of all ranks / slowest rank's time"""
    return reduce_sum(dist, local, units_this_rank) / reduce_max(dist, local, seconds_this_rank)


def split_gops(ivf):
    """-> (header32, [list of GOPs], each GOP = list of raw IVF frame records (12-byte header + data)).
    Frames before the first key frame are dropped (FilePlayer, player.cc:101-109)."""
    assert ivf[:4] == b"DKIF"
    n = struct.unpack_from("<I", ivf, 24)[0]
    pos, gops = 32, []
    for _ in range(n):
        flen = struct.unpack_from("<I", ivf, pos)[0]
        rec = ivf[pos:pos + 12 + flen]
        key = flen > 0 and not (rec[12] & 1)
        if key:
            gops.append([])
        if gops:
            gops[-1].append(rec)
        pos += 12 + flen
    return ivf[:32], gops


def shard_gop_indices(n_gops, rank, world):
    """round-robin GOP -> rank assignment (every GOP exactly once)"""
    return list(range(rank, n_gops, world))


def shard_ivf(ivf, rank, world):
    """the sub-stream (a valid IVF) holding this rank's GOPs, and their indices in the full stream"""
    hdr, gops = split_gops(ivf)
    mine = shard_gop_indices(len(gops), rank, world)
    recs = [r for g in mine for r in gops[g]]
    h = bytearray(hdr)
    struct.pack_into("<I", h, 24, len(recs))
    return bytes(h) + b"".join(recs), mine
