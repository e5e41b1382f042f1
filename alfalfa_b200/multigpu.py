"""Multi-GPU plumbing: the decode path shards at GOP / stream granularity -- a key
frame resets all codec state (decoder_state.hh:90) -- so every rank decodes its own GOPs on its own
GPU and GOP-aligned shards need NO data-path collective: torch.distributed is used for the barrier
around the timed region and for reducing the timing / work counters (NCCL on the GPU box, gloo in the
CPU tests).  The one real exchange step of the path -- a GOP that continues on another GPU needs the
DecoderState and the reference rasters -- is broadcast_decoder() at the end of this file."""
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


# ------------------------------------------------------------------------------------------------
# A GOP that continues on another GPU (BASELINE.json config 4: "NCCL golden/altref broadcast over
# NVLink"): frames inside a GOP depend on the three reference rasters and the DecoderState, so the
# rank that decoded the first part broadcasts them -- the state as a byte blob, every DISTINCT
# reference raster once (after a key frame all three are one raster) as a uint8 tensor on the
# device (NCCL) or the host (gloo).  One exchange step per boundary; GOP-aligned shards need none.
# ------------------------------------------------------------------------------------------------
def reference_plan(ids):
    """(unique ids in first-seen order, index of each of the 3 references into that list)"""
    uniq = []
    for i in ids:
        if i not in uniq:
            uniq.append(i)
    return uniq, [uniq.index(i) for i in ids]


def broadcast_bytes(dist, local, blob, src):
    """every rank returns src's bytes"""
    import torch
    dev = _device(dist, local)
    n = torch.tensor([len(blob) if blob is not None else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def _device(dist, local):
    import torch
    return torch.device("cuda", local) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_decoder(ctx, decoder, src, dist, local, make_decoder=None):
    """Decoder (state + references) of rank `src` on every rank.  `decoder` is ignored elsewhere.
    Returns src's own decoder on src, a new Decoder on the other ranks (make_decoder(ctx, state_blob,
    (last, golden, alternative) rasters) builds it; default: Decoder.from_state)."""
    import torch
    rank = dist.get_rank()
    dev = _device(dist, local)
    refs, head = None, None
    if rank == src:
        refs = decoder.get_references()
        uniq, index = reference_plan([r.id for r in refs])
        by_id = {r.id: r for r in refs}
        head = struct.pack("<4B", len(uniq), *index) + decoder.get_state().serialize()
    head = broadcast_bytes(dist, local, head, src)
    n_uniq, index = head[0], list(head[1:4])
    nbytes = ctx.frame_bytes
    rasters = []
    for k in range(n_uniq):
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if rank == src:
            by_id[uniq[k]].export_to(t.data_ptr(), nbytes)   # synchronous: the tensor is complete
        dist.broadcast(t, src)
        if rank != src:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)   # the broadcast ran on torch's stream
            fr = ctx.alloc_frame()
            fr.import_from(t.data_ptr(), nbytes)
            rasters.append(fr)
    if rank == src:
        for r in refs:
            r.release()
        return decoder
    if make_decoder is None:
        from .decoder import Decoder, DecoderState

        def make_decoder(c, blob, three):
            return Decoder.from_state(c, DecoderState.deserialize(blob), three)
    out = make_decoder(ctx, head[4:], tuple(rasters[i] for i in index))
    for fr in rasters:
        fr.release()
    return out


# ------------------------------------------------------------------------------------------------
# The same hand-over through the C ABI (include/vp8gpu.h "the exchange step"): ncclBroadcast straight from /
# into the rasters on the context's lane stream -- no torch tensors, no staging copies, no host synchronisation
# on the raster path.  torch.distributed is only used once, to hand the NCCL id to the other ranks.
# ------------------------------------------------------------------------------------------------
class Comm:
    """vp8gpu_comm: one NCCL communicator per context and rank set"""

    def __init__(self, ctx, dist, local):
        import ctypes as C

        import torch
        from . import capi
        self.ctx, self.L = ctx, ctx.L
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            capi.check(self.L.vp8gpu_comm_unique_id(uid), ctx.h, "comm_unique_id")
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=_device(dist, local))
        dist.broadcast(t, 0)
        uid = (C.c_uint8 * 128)(*t.cpu().tolist())
        self.h = C.c_void_p()
        capi.check(self.L.vp8gpu_comm_create(ctx.h, self.rank, self.size, uid, C.byref(self.h)), ctx.h, "comm_create")

    def close(self):
        if self.h:
            self.L.vp8gpu_comm_destroy(self.h)
            self.h = None

    def broadcast_bytes(self, blob, src):
        import ctypes as C
        from . import capi
        n = (C.c_uint32 * 1)(len(blob) if self.rank == src else 0)
        capi.check(self.L.vp8gpu_comm_broadcast_bytes(self.h, src, n, 4), self.ctx.h, "comm_broadcast_bytes")
        buf = (C.c_uint8 * max(1, n[0]))()
        if self.rank == src:
            C.memmove(buf, blob, len(blob))
        capi.check(self.L.vp8gpu_comm_broadcast_bytes(self.h, src, buf, n[0]), self.ctx.h, "comm_broadcast_bytes")
        return bytes(buf[:n[0]])

    def broadcast_frames(self, rasters, src, lane=0):
        """collective; `rasters` = RasterHandles: the ones to send on src, freshly allocated ones elsewhere"""
        import ctypes as C
        from . import capi
        ids = (C.c_int32 * len(rasters))(*[r.id for r in rasters])
        capi.check(self.L.vp8gpu_comm_broadcast_frames(self.h, src, lane, ids, len(rasters)), self.ctx.h, "comm_broadcast_frames")


def broadcast_decoder_capi(ctx, comm, decoder, src):
    """broadcast_decoder over the C ABI communicator: returns src's decoder on src, a new Decoder elsewhere"""
    from .decoder import Decoder, DecoderState
    refs, head = None, b""
    if comm.rank == src:
        refs = decoder.get_references()
        uniq, index = reference_plan([r.id for r in refs])
        by_id = {r.id: r for r in refs}
        head = struct.pack("<4B", len(uniq), *index) + decoder.get_state().serialize()
    head = comm.broadcast_bytes(head, src)
    n_uniq, index = head[0], list(head[1:4])
    if comm.rank == src:
        rasters = [by_id[u] for u in uniq]
        lane = ctx.L.vp8gpu_decoder_lane(decoder.h)
    else:
        rasters = [ctx.alloc_frame() for _ in range(n_uniq)]
        lane = 0
    comm.broadcast_frames(rasters, src, lane)   # queued on the lane stream; later decodes order themselves after it
    if comm.rank == src:
        for r in refs:
            r.release()
        return decoder
    out = Decoder.from_state(ctx, DecoderState.deserialize(head[4:]), tuple(rasters[i] for i in index))
    for fr in rasters:
        fr.release()
    return out
