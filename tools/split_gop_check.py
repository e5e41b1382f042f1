#!/usr/bin/env python3
"""A GOP split across GPUs (BASELINE.json config 4): rank 0 decodes the first part of a single-GOP
stream with live golden / altref references, broadcasts its Decoder (DecoderState blob + the distinct
reference rasters as uint8 tensors, NCCL over NVLink), and the other ranks continue.  Every rank also
decodes the whole stream by itself and checks that the continued decode is bit-identical: every frame
after the hand-over, and the final Decoder (state + the three rasters).

launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
            --master-port 29533 tools/split_gop_check.py [ivf] [split_frame]"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from alfalfa_b200 import Context, Decoder, multigpu
    from alfalfa_b200.decoder import read_ivf
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bench_data", "features1080p_12f.ivf")
    rank, world, local, dist = multigpu.init()
    assert dist is not None, "run under torch.distributed.run with >= 2 ranks"
    w, h, frames = read_ivf(open(path, "rb").read())
    split = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else len(frames) // 2
    ctx = Context(w, h, device=local, max_frames=32)
    # everybody: the whole stream alone (the truth on this rank)
    alone = Decoder(ctx)
    want = []
    for f in frames:
        shown, r = alone.get_frame_output(f)
        want.append(hashlib.sha1(r.display_bytes()).hexdigest())
        r.release()
    # rank 0 decodes the first part, then hands over
    dec = None
    if rank == 0:
        dec = Decoder(ctx)
        for i, f in enumerate(frames[:split]):
            shown, r = dec.get_frame_output(f)
            assert hashlib.sha1(r.display_bytes()).hexdigest() == want[i]
            r.release()
    # the hand-over through the C ABI communicator (ncclBroadcast straight between rasters on the lane stream);
    # --torch uses the round-1 path (uint8 tensors + dist.broadcast + export / import copies) for comparison
    use_torch = "--torch" in sys.argv
    comm = None if use_torch else multigpu.Comm(ctx, dist, local)
    exchange_us = None
    if comm is not None:
        # raster exchange alone, amortised: 50 broadcasts of one raster queued back to back on lane 0
        probe = [ctx.alloc_frame()]
        comm.broadcast_frames(probe, 0)
        ctx.sync()
        multigpu.barrier(dist)
        t0 = time.perf_counter()
        for _ in range(50):
            comm.broadcast_frames(probe, 0)
        ctx.sync()
        exchange_us = (time.perf_counter() - t0) / 50 * 1e6
        probe[0].release()
    multigpu.barrier(dist)
    t0 = time.perf_counter()
    if comm is not None:
        dec = multigpu.broadcast_decoder_capi(ctx, comm, dec, 0)
        t_queued = time.perf_counter() - t0
        ctx.sync()
    else:
        dec = multigpu.broadcast_decoder(ctx, dec, 0, dist, local)
        t_queued = time.perf_counter() - t0
    dt = time.perf_counter() - t0
    multigpu.barrier(dist)
    bad = 0
    for i, f in enumerate(frames[split:], start=split):
        shown, r = dec.get_frame_output(f)
        bad += hashlib.sha1(r.display_bytes()).hexdigest() != want[i]
        r.release()
    equal = dec == alone
    total_bad = multigpu.reduce_sum(dist, local, bad + (0 if equal else 1))
    if rank == 0:
        print(json.dumps({"check": "split GOP across GPUs", "ranks": world, "stream": os.path.basename(path),
                          "frames": len(frames), "split_at": split, "mismatches": int(total_bad),
                          "handover_ms": dt * 1e3, "handover_queued_ms": t_queued * 1e3,
                          "raster_exchange_us": exchange_us, "path": "torch tensors" if use_torch else "C ABI ncclBroadcast",
                          "raster_bytes": ctx.frame_bytes, "backend": dist.get_backend()}))
    del dec, alone
    if comm is not None:
        comm.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if total_bad else 0)


if __name__ == "__main__":
    main()
