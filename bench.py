#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 VP8 decode hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): "1080p30 IVF decode, 1 GPU": the synthetic 1080p stream
bench_data/synth1080p_medium_q90.ivf (60 frames = 2 GOPs of 30, ~11 Mbit/s at 30 fps, produced by
the reference's own encoder, tools/make_bench_streams.sh) repeated R times; GOPs are independent
(a key frame resets all codec state), so the repeats are extra GOPs of a longer stream.

One JSON line on stdout (rank 0):
  value      Mpix/s of the device pipeline with the parsed records already resident in HBM
             (all GOP instances advance one frame position per batch: 30 batches per step)
  e2e        Mpix/s through the public C ABI call vp8gpu_decode_ivf with HOST buffers: bitstream
             in host memory -> CPU entropy front end -> H2D records -> kernels -> D2H of every
             shown frame into pinned host memory, all inside the timed region
  roofline   dominant kernel: algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline  the unmodified reference decoder (oracle/_ref/ref_dump) on one host core
--impl reference: the reference's CPU decode on all usable host cores (one process per core, each
decoding the whole clip), same metric / unit / config.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STREAMS = {"medium": "synth1080p_medium_q90.ivf", "easy": "synth1080p_easy_q40.ivf",
           "720p": "synth720p_medium_q90.ivf"}   # 720p: BASELINE.json config 5 (--gop-instances 64 = 64 streams in lock-step)
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def effective_cpus():
    """CPUs this process may actually use: affinity mask and cgroup quota (the GPU box gives the
    container a CPU quota well below the 128 logical CPUs it shows)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def replicate_ivf(data, reps):
    """IVF with the frames of `data` repeated `reps` times (extra GOPs)."""
    import struct
    n = struct.unpack_from("<I", data, 24)[0]
    body = data[32:]
    hdr = bytearray(data[:32])
    struct.pack_into("<I", hdr, 24, n * reps)
    return bytes(hdr) + body * reps


def reference_mpix_per_s(path, procs, reps=1):
    """`procs` concurrent reference decoders (one process per core), each decoding the whole clip."""
    t0 = time.perf_counter()
    ps = [subprocess.Popen([REF_DUMP, "time", path, str(reps)], stdout=subprocess.PIPE, text=True) for _ in range(procs)]
    outs = [json.loads(p.communicate()[0]) for p in ps]
    wall = time.perf_counter() - t0
    frames = sum(o["frames"] for o in outs) * reps
    mpix = outs[0]["width"] * outs[0]["height"] * frames / 1e6
    return mpix / wall, wall, outs


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    path = os.path.join(ROOT, "bench_data", STREAMS[a.workload])
    cores = a.ref_procs or effective_cpus()
    if not os.path.exists(REF_DUMP):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_dump not built"}))
        return
    for _ in range(a.warmup):
        reference_mpix_per_s(path, cores)
    vals, walls = [], []
    for _ in range(a.steps):
        v, wall, _ = reference_mpix_per_s(path, cores)
        vals.append(v)
        walls.append(wall)
    v = statistics.mean(vals)
    sample = "%d processes x 60-frame 1080p clip per step" % cores
    print(json.dumps({
        "impl": "reference", "metric": "decode_throughput", "value": v, "unit": "Mpix/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": statistics.mean(walls) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "IVF decode (%s), reference CPU decoder, C++ fallback build (no yasm)" % STREAMS[a.workload]},
        "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "reference", "sample": sample, "usable_cpus": effective_cpus()},
        "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def ncu_traffic(kernel, gop_instances, path=os.path.join(ROOT, "profiles", "r1c_ncu_full_summary.csv")):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture (profiles/r1c_notes.md; taken at 64 GOP instances, so only reported for that configuration)"""
    if gop_instances != 64 or not os.path.exists(path):
        return None
    import csv
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    r_i, w_i = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    vals = [float(r[r_i]) * scale[units[r_i]] + float(r[w_i]) * scale[units[w_i]] for r in rows[2:] if kernel + "(" in r[0]]
    return sum(vals) / len(vals) if vals else None


def synth_1080p(t, w=1920, h=1080, seed=11):
    """seeded synthetic source: drifting smooth pattern, four translating softly textured tiles, light noise"""
    import numpy as np
    rng = np.random.default_rng(seed)
    tex = rng.integers(0, 256, (256, 256)).astype(np.float32)
    for _ in range(3):
        tex = (tex + np.roll(tex, 1, 0) + np.roll(tex, 1, 1) + np.roll(tex, (1, 1), (0, 1))) / 4
    tex = np.clip(128 + (tex - 128) * 3, 0, 255)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    y = 128 + 50 * np.sin(0.012 * (xx + 3 * t)) * np.cos(0.009 * (yy + 2 * t))
    for k, (vx, vy) in enumerate(((3, 1), (-2, 2), (1, -3), (-3, -2))):
        ox, oy = 200 + 400 * k + vx * t, 150 + 180 * k + vy * t
        y[oy:oy + 256, ox:ox + 256] = tex
    y += np.random.default_rng(seed + 1 + t).integers(-1, 2, (h, w))
    cy, cx = np.mgrid[0:(h + 1) // 2, 0:(w + 1) // 2].astype(np.float32)
    u = 128 + 30 * np.sin(0.01 * (cx + t))
    v = 128 + 30 * np.cos(0.012 * (cy - t))
    return tuple(np.clip(p, 0, 255).astype(np.uint8) for p in (y, u, v))


def bench_encode(a, local):
    """fps of Encoder::encode_with_target_size at 1080p through the C ABI (host planes in, compressed
    frame out, every step incl. H2D of the source and D2H of the records), next to the reference
    encoder (REALTIME_QUALITY, one core) on the same raw frames."""
    import numpy as np

    from alfalfa_b200 import Context, Encoder
    w, h, n = 1920, 1080, a.encode_frames
    src = [synth_1080p(t) for t in range(n)]
    ctx = Context(w, h, device=local, max_frames=32)
    enc = Encoder(ctx)
    enc.encode_with_target_size(*src[0], a.encode_target)  # warm-up (key frame, allocations)
    del enc
    enc = Encoder(ctx)
    sizes, qis, psnrs, times, ssims, lfs = [], [], [], [], [], []
    for t in range(n):
        t0 = time.perf_counter()
        blob, qi = enc.encode_with_target_size(*src[t], a.encode_target)
        times.append(time.perf_counter() - t0)
        sizes.append(len(blob))
        qis.append(qi)
        st = enc.stats()
        ssims.append(st["ssim"])
        lfs.append(st["loop_filter_level"])
        rec = enc.reconstruction()
        ry = rec.planes()[0][:h, :w]
        rec.release()
        mse = float(np.mean((ry.astype(np.float64) - src[t][0].astype(np.float64)) ** 2))
        psnrs.append(99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse))
    launches = ctx.launch_count()
    del enc
    ctx.close()
    out = {"metric": "encode_with_target_size fps @1080p", "frames": n, "target_bytes": a.encode_target,
           "fps": (n - 1) / sum(times[1:]), "key_frame_ms": times[0] * 1e3, "inter_frame_ms": 1e3 * sum(times[1:]) / (n - 1),
           "bytes_per_frame": sum(sizes) / n, "qi": qis, "loop_filter_level": lfs, "psnr_y": sum(psnrs) / n,
           "ssim_y": sum(ssims) / n, "gpu_launches": int(launches),
           "note": "first slice: SAD decisions with zero/left/above vector candidates, 16x16 intra modes, LAST reference, "
                   "SSIM-driven loop-filter search; closed loop verified in tests/test_gpu_encoder.py"}
    ref_enc = os.path.join(ROOT, "oracle", "_ref", "ref_encode")
    if os.path.exists(ref_enc):
        import tempfile
        m = min(n, 6)  # bounded sample: the reference runs at a few fps
        with tempfile.TemporaryDirectory() as d:
            raw = os.path.join(d, "src.yuv")
            with open(raw, "wb") as f:
                for t in range(m):
                    for p in src[t]:
                        f.write(p.tobytes())
            env = dict(os.environ, REF_RAW=raw, REF_TARGET=str(a.encode_target))
            r = subprocess.run([ref_enc, os.path.join(d, "o.ivf"), str(w), str(h), str(m), "1000", "0"], env=env,
                               capture_output=True, text=True)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                out["reference"] = {"fps": j["fps"], "bytes_per_frame": j["bytes"] / m, "psnr_y": j["psnr_y"], "frames": m,
                                    "cores": 1, "kind": "reference", "sample": "unmodified reference encoder, REALTIME_QUALITY, "
                                    "encode_with_target_size, same raw frames, SSIM restated (parity unpinned)"}
            except Exception as e:  # noqa: BLE001
                out["reference"] = {"unavailable": "%s %s" % (e, r.stderr[-200:])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="medium", choices=list(STREAMS))
    ap.add_argument("--gop-instances", type=int, default=64, help="GOPs advanced together in the HBM-resident run")
    ap.add_argument("--replicas", type=int, default=0, help="stream repeats for the end-to-end run (0 = auto)")
    ap.add_argument("--threads", type=int, default=0, help="host workers for the end-to-end run (0 = auto)")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference processes (0 = usable CPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encode", action="store_true", help="skip the 1080p encode section")
    ap.add_argument("--encode-frames", type=int, default=12)
    ap.add_argument("--encode-target", type=int, default=45000, help="bytes per frame for encode_with_target_size")
    ap.add_argument("--host-tokens", action="store_true",
                    help="end-to-end run: DCT partitions decoded by the host workers instead of k_tokens on the device")
    ap.add_argument("--no-output", action="store_true", help="diagnostic: leave decoded frames on the device")
    ap.add_argument("--host-stats", action="store_true", help="diagnostic: print host time accounting to stderr")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 0)

    from alfalfa_b200 import multigpu as M
    if a.impl == "reference":  # CPU arm: rank 0 alone runs it, no process group needed
        rank, world, _ = M.rank_info()
        run_reference_arm(a, rank, world)
        return
    rank, world, local, dist = M.init()
    barrier, max_over_ranks, sum_over_ranks = M.barrier, M.reduce_max, M.reduce_sum

    import numpy as np

    import oracle_lib as O  # only for read_ivf and (rank 0) the cpu_baseline leg
    from alfalfa_b200 import Context, capi

    path = os.path.join(ROOT, "bench_data", STREAMS[a.workload])
    data = open(path, "rb").read()
    w, h, frames = O.read_ivf(data)
    mpix_frame = w * h / 1e6
    L = capi.lib()

    # ---------------- set-up: parse once, build the HBM-resident batches ----------------
    G = a.gop_instances
    gop_len = 30
    n_gops_in_clip = len(frames) // gop_len
    ctx = Context(w, h, device=local, max_frames=G * (gop_len + 1) + 600)
    st, pf = C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    parsed = []  # per clip frame: (desc, mbs, tok, split) numpy copies
    n_mbs = ((w + 15) // 16) * ((h + 15) // 16)
    h2d_per_clip = 0
    h2d_dev_tokens = 0
    z_blocks = 0
    for f in frames:
        capi.check(L.vp8gpu_parse_frame(st, f, len(f), pf), None, "parse")
        d = capi.FrameDesc.from_buffer_copy(bytes(L.vp8gpu_parsed_desc(pf).contents))
        mbs = np.frombuffer(C.string_at(L.vp8gpu_parsed_mbs(pf), n_mbs * 32), dtype=capi.MB_DTYPE).copy()
        tok = (np.frombuffer(C.string_at(L.vp8gpu_parsed_tokens(pf), d.n_tokens * 4), dtype="<u4").copy()
               if d.n_tokens else np.zeros(1, "<u4"))
        sp = (np.frombuffer(C.string_at(L.vp8gpu_parsed_split(pf), d.n_split * 64), dtype="u1").copy()
              if d.n_split else np.zeros(64, "u1"))
        parsed.append((d, mbs, tok, sp))
        h2d_per_clip += n_mbs * 32 + d.n_tokens * 4 + d.n_split * 64 + 512
        h2d_dev_tokens += n_mbs * 32 + d.n_split * 64 + 1536 + len(f)  # records + probabilities + raw partitions
        if d.n_tokens:
            z_blocks += len(np.unique((tok[:d.n_tokens] >> 20) & 31 | (np.repeat(np.arange(n_mbs), mbs["tok_cnt"]) << 5)))
    L.vp8gpu_state_destroy(st)
    L.vp8gpu_parsed_destroy(pf)

    # frame ids per (gop instance, position); references follow Frame::copy_to (frame.cc:272-307)
    batches = (C.c_void_p * gop_len)()
    keep = []
    for pos in range(gop_len):
        jobs = (capi.Job * G)()
        for g in range(G):
            d, mbs, tok, sp = parsed[(g % n_gops_in_clip) * gop_len + pos]
            if pos == 0:
                refs_g = [-1, -1, -1]
                keep.append({"refs": refs_g, "frames": []})
            state = keep[g]
            out = ctx.alloc_frame()
            state["frames"].append(out)
            jobs[g].desc = C.pointer(d)
            jobs[g].mbs = mbs.ctypes.data
            jobs[g].tokens = tok.ctypes.data
            jobs[g].split = sp.ctypes.data
            jobs[g].refs[:] = state["refs"]
            jobs[g].out = out.id
            r = state["refs"]
            if d.key_frame:
                r[0] = r[1] = r[2] = out.id
            else:
                if d.copy_to_alternate == 1:
                    r[2] = r[0]
                elif d.copy_to_alternate == 2:
                    r[2] = r[1]
                if d.copy_to_golden == 1:
                    r[1] = r[0]
                elif d.copy_to_golden == 2:
                    r[1] = r[2]
                if d.refresh_golden:
                    r[1] = out.id
                if d.refresh_alternate:
                    r[2] = out.id
                if d.refresh_last:
                    r[0] = out.id
        b = C.c_void_p()
        capi.check(L.vp8gpu_batch_upload(ctx.h, jobs, G, C.byref(b)), ctx.h, "batch_upload")
        batches[pos] = b
    ctx.sync()

    def resident_step():
        ms = C.c_float(0)
        capi.check(L.vp8gpu_batches_run(ctx.h, 0, batches, gop_len, C.byref(ms)), ctx.h, "batches_run")
        return ms.value

    # ---------------- HBM-resident run: `value` ----------------
    for _ in range(max(a.warmup, 3)):
        resident_step()
    ctx.sync()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier(dist)
    ctx.sync()
    step_ms = [resident_step() for _ in range(a.steps)]
    ctx.sync()
    barrier(dist)
    launches_resident = ctx.launch_count() - launches0
    total_ms = max_over_ranks(dist, local, sum(step_ms))
    frames_per_step = G * gop_len
    value = sum_over_ranks(dist, local, frames_per_step * mpix_frame * a.steps) / (total_ms / 1e3)

    # correctness of what was timed: first GOP instance's last frame == oracle-free self check
    # (bit-exact parity is the job of tests/; here we only make sure the frames are not garbage)
    y, _, _ = keep[0]["frames"][-1].planes()
    assert y.std() > 1.0, "decoded frame looks empty"

    # ---------------- per-kernel timing for the roofline ----------------
    kt = np.zeros((gop_len, 3))
    reps = 3
    for _ in range(reps):
        for pos in range(gop_len):
            ms3 = (C.c_float * 3)()
            capi.check(L.vp8gpu_batch_run_timed(ctx.h, 0, batches[pos], ms3), ctx.h, "batch_run_timed")
            kt[pos] += np.array(list(ms3)) / reps
    k_total = kt.sum(axis=0)  # ms per step per kernel
    names = ["k_inter", "k_intra", "k_loopfilter"]
    dom = int(np.argmax(k_total))
    # algorithmic bytes per launch (DESIGN.md "kernels and their rooflines"); P = 384 bytes per MB
    n_inter = sum(int((p[1]["ref_frame"] != 0).sum()) for p in parsed[:gop_len * n_gops_in_clip]) / n_gops_in_clip
    n_intra = sum(int((p[1]["ref_frame"] == 0).sum()) for p in parsed[:gop_len * n_gops_in_clip]) / n_gops_in_clip
    n_tok = sum(p[0].n_tokens for p in parsed) / n_gops_in_clip
    n_filt = sum(int((p[1]["lf_level"] != 0).sum()) for p in parsed) / n_gops_in_clip
    per_gop_bytes = {
        "k_inter": n_inter * (384 * 2 + 32) + 4 * n_tok * (n_inter / max(n_inter + n_intra, 1)),
        "k_intra": n_intra * (384 + 32) + 4 * n_tok * (n_intra / max(n_inter + n_intra, 1)),
        "k_loopfilter": n_filt * (384 * 2 + 32),
    }
    launches_per_step = {"k_inter": gop_len - 1, "k_intra": gop_len, "k_loopfilter": gop_len}
    peak, peak_src = measured_peaks()
    dname = names[dom]
    bytes_per_launch = per_gop_bytes[dname] * G / launches_per_step[dname]
    avg_launch_ms = k_total[dom] / launches_per_step[dname]
    achieved = bytes_per_launch / (avg_launch_ms / 1e3) / 1e9
    # whole-frame budget of SURVEY.md 8(d): P + I*P + 32 Z + 48 M per frame, charged to the sum of the kernels
    P = 384 * n_mbs
    inter_frames = sum(0 if p[0].key_frame else 1 for p in parsed) / n_gops_in_clip
    pipeline_bytes = G * (gop_len * P + inter_frames * P + 32 * z_blocks / n_gops_in_clip + 48 * n_mbs * gop_len)
    pipeline_gbs = pipeline_bytes / (statistics.mean(step_ms) / 1e3) / 1e9
    per_kernel = {}
    for k, nm in enumerate(names):
        bpl = per_gop_bytes[nm] * G / launches_per_step[nm]
        gbs = bpl / (k_total[k] / launches_per_step[nm] / 1e3) / 1e9
        per_kernel[nm] = {"achieved": gbs, "frac": gbs / peak, "bytes_per_launch": bpl}
    roofline = {"bound": "hbm", "kernel": dname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(dname, G), "peak_source": peak_src,
                "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_launch_ms,
                "kernel_ms_per_step": dict(zip(names, [float(x) for x in k_total])), "kernels": per_kernel,
                "pipeline_achieved": pipeline_gbs, "pipeline_frac": pipeline_gbs / peak}

    # ---------------- end-to-end run through the public API: `e2e` ----------------
    for b in batches:
        L.vp8gpu_batch_free(ctx.h, b)
    for s_ in keep:
        for fr in s_["frames"]:
            fr.release()
    ctx.close()
    # Host workers, shared by the ranks.  With the DCT partitions decoded on the device a worker only
    # walks first partitions, and the number of GOPs in flight (= workers) is what fills the device:
    # four per usable CPU; when the workers parse everything, two (they also wait on DMA / the dispatcher).
    per_cpu = 2 if a.host_tokens else 4
    threads = a.threads or max(2, min(96, per_cpu * effective_cpus() // max(world, 1)))
    # 2 GOPs per repeat.  With device-side tokens a worker runs two GOPs ahead, so give it four: the
    # start-up (one k_tokens latency before the first pixel round) is then a smaller part of the step.
    R = a.replicas or max(16, threads * (1 if a.host_tokens else 2))
    ctx2 = Context(w, h, device=local, max_frames=threads * (10 if a.host_tokens else int(os.environ.get("VP8GPU_TOK_SLOTS", 60)) + 6) + 64)
    ctx2.set_device_tokens(not a.host_tokens)
    dst = C.c_void_p()
    while True:  # the pinned output buffer is 3.1 MB per frame: halve the run if the box cannot pin that much
        out_bytes = ctx2.display_bytes * len(frames) * R
        if L.vp8gpu_host_alloc(C.byref(dst), out_bytes) == 0:
            break
        if R <= 16:
            capi.check(capi.ERR_NOMEM, ctx2.h, "host_alloc")
        R //= 2
    big = replicate_ivf(data, R)
    n_e2e_frames = len(frames) * R
    nd, ns = C.c_uint32(0), C.c_uint32(0)

    def e2e_step():
        t0 = time.perf_counter()
        capi.check(L.vp8gpu_decode_ivf(ctx2.h, big, len(big), threads, None if a.no_output else dst,
                                       0 if a.no_output else out_bytes, C.byref(nd), C.byref(ns)), ctx2.h, "decode_ivf")
        capi.check(L.vp8gpu_ctx_sync(ctx2.h), ctx2.h, "sync")
        return time.perf_counter() - t0

    for _ in range(max(a.warmup, 1)):
        e2e_step()
    barrier(dist)
    e2e_s = [e2e_step() for _ in range(a.steps)]
    barrier(dist)
    if a.host_stats:
        stt = (C.c_double * 8)()
        L.vp8gpu_decode_ivf_stats(ctx2.h, stt)
        print("host stats (last step, s): parse %.3f wait_dispatch %.3f wait_dma %.3f | dispatcher: submit %.3f downloads %.3f "
              "idle %.3f | batches %d frames %d | step wall %.3f" % (*list(stt)[:6], int(stt[6]), int(stt[7]), e2e_s[-1]),
              file=sys.stderr)
    e2e_total = max_over_ranks(dist, local, sum(e2e_s))
    e2e_value = sum_over_ranks(dist, local, n_e2e_frames * mpix_frame * a.steps) / e2e_total
    launches_e2e = ctx2.launch_count()
    clocks = sampler.stop()
    # spot check of the end-to-end output against the HBM-resident output of the same frame
    first = np.frombuffer(C.string_at(dst, w * h), dtype=np.uint8).reshape(h, w)
    assert first.std() > 1.0
    L.vp8gpu_host_free(dst)
    ctx2.close()

    # ---------------- encode: 1080p encode_with_target_size (BASELINE.json config 3) ----------------
    encode = None
    if rank == 0 and not a.no_encode:
        encode = bench_encode(a, local)

    # ---------------- CPU baseline (rank 0, one core, bounded sample) ----------------
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        if os.path.exists(REF_DUMP):
            v, wall, outs = reference_mpix_per_s(path, 1, reps=2)
            o = outs[0]
            cpu = {"value": o["mpix_per_s"], "unit": "Mpix/s", "cores": 1, "kind": "reference",
                   "sample": "60-frame 1080p clip, best of 2, unmodified reference decoder (C++ fallback, no yasm): "
                             "parse %.2fs recon %.2fs loopfilter %.2fs" % (o["parse_s"], o["recon_s"], o["loopfilter_s"])}
        else:
            ph = (C.c_double * 3)()
            n = C.c_uint32()
            O.lib().vp8o_time_ivf(data, len(data), 2, 100000, ph, C.byref(n))
            cpu = {"value": n.value * mpix_frame / sum(ph), "unit": "Mpix/s", "cores": 1, "kind": "port",
                   "sample": "60-frame 1080p clip, best of 2, oracle/vp8_oracle.c"}

    if rank == 0:
        print(json.dumps({
            "metric": "decode_throughput", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%dx%d IVF decode, %s (reference-encoder synthetic clip, 2 GOPs x 30 frames), "
                                   "%d GOP instances per GPU advanced in lock-step, records resident in HBM" % (w, h, STREAMS[a.workload], G),
                       "frames_per_step": frames_per_step, "l2": "working set %.0f MB per step > 126 MB L2, no flush needed"
                       % (frames_per_step * 3.1), "e2e_frames_per_step": n_e2e_frames, "e2e_host_threads": threads,
                       "bit_exact": "tests/test_gpu_parity.py (53/53 golden SHA-1 + per-frame oracle)"},
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": int((h2d_per_clip if a.host_tokens else h2d_dev_tokens) * R),
                    "dct_partitions": "host workers" if a.host_tokens else "k_tokens on the device",
                    "d2h_bytes_per_step": int(out_bytes), "ms_per_step": e2e_total / a.steps * 1e3,
                    "api": "vp8gpu_decode_ivf (host IVF bytes -> pinned host YUV)"},
            "gpu_launches": int(launches_resident), "gpu_launches_e2e": int(launches_e2e),
            "roofline": roofline, "cpu_baseline": cpu, "encode": encode, "clocks": clocks}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
