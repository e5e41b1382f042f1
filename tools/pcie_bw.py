#!/usr/bin/env python3
"""Device-to-host copy bandwidth of this box into pinned memory, in the shape vp8gpu_decode_ivf uses
(one 1080p frame = three copies of 2.07 / 0.52 / 0.52 MB, several streams): the ceiling of `e2e`, which
copies every decoded frame to the host.  usage: python tools/pcie_bw.py"""
import json
import time

import torch


def main():
    dev = torch.device("cuda", 0)
    sizes = [1920 * 1080, 960 * 540, 960 * 540]
    n_frames, n_streams = 512, 8
    src = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
    dst = torch.empty(n_frames * sum(sizes), dtype=torch.uint8).pin_memory()
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
    out = {}
    for label, ns in (("8_streams", n_streams), ("1_stream", 1)):
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            off = 0
            for f in range(n_frames):
                with torch.cuda.stream(streams[f % ns]):
                    for k, s in enumerate(sizes):
                        dst[off:off + s].copy_(src[k], non_blocking=True)
                        off += s
            torch.cuda.synchronize()
            best = max(best, n_frames * sum(sizes) / (time.perf_counter() - t0) / 1e9)
        out[label] = best
    big_src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    big_dst = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    big_dst.copy_(big_src, non_blocking=True)
    torch.cuda.synchronize()
    out["one_1GiB_copy"] = (1 << 30) / (time.perf_counter() - t0) / 1e9
    print(json.dumps({"d2h_GB_per_s": out, "frame_bytes": sum(sizes),
                      "frames_per_s_ceiling": out["8_streams"] * 1e9 / sum(sizes),
                      "mpix_per_s_ceiling": out["8_streams"] * 1e9 / sum(sizes) * 1920 * 1080 / 1e6}))


if __name__ == "__main__":
    main()
