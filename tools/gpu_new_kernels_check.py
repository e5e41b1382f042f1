#!/usr/bin/env python
"""One short process that takes the code WITHOUT a hardware record through its paces on the device and writes what
happened to gpurun_out/r2z_new_kernels.json after every step (so that a timeout still leaves the steps that finished):
re-encoding (k_reenc_inter, k_reenc_intra, the update_residues / reencode_as_interframe / write_frame host code), the
two-pass key frame (k_enc_rd<true>), an Encoder built from a Decoder in a libvpx state.  Every comparison is against
the unmodified reference's tools in oracle/_ref (byte identity), the same checks as tests/test_gpu_reencode.py and
tests/test_gpu_encoder.py.  Usage: python tools/gpu_new_kernels_check.py [out.json]"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2z_new_kernels.json")
results = {"steps": []}
T0 = time.time()


def step(name, fn):
    t = time.time()
    try:
        r = fn()
        results["steps"].append({"step": name, "ok": bool(r is None or r is True or (isinstance(r, dict) and r.get("ok", True))),
                                 "detail": r if isinstance(r, dict) else None, "s": round(time.time() - t, 2)})
    except Exception as e:  # noqa: BLE001
        results["steps"].append({"step": name, "ok": False, "error": "%s: %s" % (type(e).__name__, e),
                                 "trace": traceback.format_exc()[-1200:], "s": round(time.time() - t, 2)})
    results["elapsed_s"] = round(time.time() - T0, 2)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(results, open(OUT, "w"), indent=1)
    print(json.dumps(results["steps"][-1])[:400], flush=True)


def main():
    import numpy as np
    import oracle_lib as O
    import reencode_worker as W
    import test_gpu_encoder as E
    import test_gpu_reencode as R
    from alfalfa_b200 import Context, Decoder, Encoder

    def product_reencode(w, h, targets, pred, state, kfw, extra):
        ctx = Context(w, h, max_frames=24)
        pd = Decoder(ctx)
        pfs = []
        for c in pred:
            pf = pd.parse_frame(c, keep_labels=True)
            pd.decode_frame(pf)
            pfs.append(pf)
        enc = Encoder.from_decoder(ctx, Decoder.deserialize(ctx, state))
        launches0 = ctx.launch_count()
        frames = enc.reencode(targets, pfs, kfw, extra)
        launches = ctx.launch_count() - launches0
        rx = Decoder.deserialize(ctx, state)
        for c in frames:
            rx.get_frame_output(c)
        in_step = rx == enc.export_decoder()
        ctx.close()
        return frames, bool(in_step), int(launches)

    def reencode_case(w, h, extra, kfw=0.75, n=4):
        targets, pred, state = R.make_case(w, h, n, qi_a=40, qi_b=60)
        want = R.reference_reencode(w, h, targets, pred, state, kfw, extra)
        got, in_step, launches = product_reencode(w, h, targets, pred, state, kfw, extra)
        same = [a == b for a, b in zip(got, want)]
        return {"ok": len(got) == len(want) and all(same) and in_step, "identical": same, "in_step": in_step, "gpu_launches": launches}

    def vector_case(prev, this, nframes):
        pw, ph, prev_chunks = R._golden(R._full_name(prev))
        w, h, chunks = R._golden(R._full_name(this))
        chunks = chunks[:nframes]
        state = R.reference_state_after(w, h, prev_chunks, len(prev_chunks))
        targets = R._decoded_targets(w, h, chunks)
        want = R.reference_reencode(w, h, targets, chunks, state, 0.75, True)
        got, in_step, launches = product_reencode(w, h, targets, chunks, state, 0.75, True)
        same = [a == b for a, b in zip(got, want)]
        return {"ok": len(got) == len(want) and all(same) and in_step, "identical": sum(same), "frames": len(want), "in_step": in_step,
                "gpu_launches": launches}

    def two_pass(w, h, qi, amp, n=2):
        frames = [E._noisy(w, h, t, amp) for t in range(n)]
        os.environ["REF_TWO_PASS"] = "1"
        try:
            want = E.reference_encode(frames, w, h, qi=qi)
        finally:
            del os.environ["REF_TWO_PASS"]
        ctx = Context(w, h, max_frames=16)
        enc = Encoder(ctx)
        enc.set_two_pass(True)
        got = [enc.encode_with_quantizer(*f, qi) for f in frames]
        ctx.close()
        return {"ok": got == want, "identical": [a == b for a, b in zip(got, want)]}

    def any_state(prefix):
        full = R._full_name(prefix)
        w, h, chunks = R._golden(full)
        ctx = Context(w, h, max_frames=24)
        rx = Decoder(ctx)
        for c in chunks[:10]:
            rx.get_frame_output(c)
        enc = Encoder.from_decoder(ctx, rx)
        ok = True
        for t in range(3):
            rx.get_frame_output(enc.encode_with_quantizer(*E.synth(w, h, t), 36 + 8 * t))
            ok = ok and (rx == enc.export_decoder())
        ctx.close()
        return {"ok": bool(ok)}

    step("argument errors + first contact of every re-encoding entry point (64x64)", lambda: W.errors())
    step("extra-frame chunk 176x144 (update_residues, options 2 + 4)", lambda: reencode_case(176, 144, True))
    step("whole chunk 176x144 (reencode_as_interframe + update_residues, options 1 + 4)", lambda: reencode_case(176, 144, False))
    step("libvpx prediction stream 07b5eb1e after 04b68b0a (SPLITMV, golden / altref, intra MBs in inter frames)",
         lambda: vector_case("04b68b0a", "07b5eb1e", 20))
    step("two-pass key frame 176x144 (k_enc_rd<true>)", lambda: two_pass(176, 144, 70, 40))
    step("Encoder from a Decoder after a libvpx stream (07b5eb1e)", lambda: any_state("07b5eb1e"))
    step("whole chunk 640x360", lambda: reencode_case(640, 360, False))
    step("libvpx prediction stream a61782d0 after 353ee97f (352x288)", lambda: vector_case("353ee97f", "a61782d0", 15))
    step("two-pass key frame 640x368", lambda: two_pass(640, 368, 40, 30))
    results["all_ok"] = all(s["ok"] for s in results["steps"])
    json.dump(results, open(OUT, "w"), indent=1)
    print("all ok" if results["all_ok"] else "FAILURES", flush=True)


if __name__ == "__main__":
    main()
