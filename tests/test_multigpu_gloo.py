"""N>1 path on CPU: world_size-2 gloo processes exercise the sharding and the reductions that
bench.py / multi-GPU decode use (GOP-level data parallelism, no data-path collective)."""
import hashlib
import os
import subprocess
import sys

import pytest

import oracle_lib as O
from conftest import GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    """a rendezvous port nobody is using right now (fixed ports collide with lingering sockets)"""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return str(s.getsockname()[1])

WORKER = r'''
import hashlib, json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import oracle_lib as O
from alfalfa_b200 import multigpu as M
rank, world, local, dist = M.init("gloo")
data = open(%(vec)r, "rb").read()
sub, mine = M.shard_ivf(data, rank, world)
# each rank decodes only its own GOPs (here with the CPU oracle standing in for the GPU)
out = O.decode_ivf_display(sub)
w, h, frames = O.read_ivf(sub)
units = len(frames) * w * h / 1e6
seconds = 1.0 + rank          # pretend rank 1 is slower: the job is as slow as its slowest rank
M.barrier(dist)
agg = M.aggregate_throughput(dist, local, units, seconds)
total_frames = M.reduce_sum(dist, local, len(frames))
os.write(1, (json.dumps({"rank": rank, "gops": mine, "frames": len(frames), "sha": hashlib.sha1(out).hexdigest(),
                  "agg": agg, "total_frames": total_frames, "units": units}) + "\n").encode())
dist.destroy_process_group()
'''


def test_two_rank_gop_sharding_and_reductions(tmp_path):
    vec = os.path.join(GOLDEN_DIR, "45502fe01a62b82d498b83dc50824741402436db")  # 30 key frames = 30 GOPs
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "vec": vec})
    port = _free_port()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    rows = sorted((json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")), key=lambda r: r["rank"])
    assert [r["rank"] for r in rows] == [0, 1]
    # every GOP decoded exactly once, round-robin
    assert sorted(rows[0]["gops"] + rows[1]["gops"]) == list(range(30))
    assert rows[0]["gops"] == list(range(0, 30, 2))
    assert rows[0]["total_frames"] == rows[1]["total_frames"] == 30
    # whole-job throughput = all units / slowest rank
    units = rows[0]["units"] + rows[1]["units"]
    assert abs(rows[0]["agg"] - units / 2.0) < 1e-9 and rows[0]["agg"] == rows[1]["agg"]
    # the shards together reproduce the full stream's output (all-key-frame clip: order = interleave)
    from alfalfa_b200 import multigpu as M
    data = open(vec, "rb").read()
    full = O.decode_ivf_display(data)
    w, h, _ = O.read_ivf(data)
    fb = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)
    shards = [O.decode_ivf_display(M.shard_ivf(data, r, 2)[0]) for r in range(2)]
    for g in range(30):
        r, k = g % 2, g // 2
        assert shards[r][k * fb:(k + 1) * fb] == full[g * fb:(g + 1) * fb]
    assert hashlib.sha1(full).hexdigest() == "45502fe01a62b82d498b83dc50824741402436db"


def test_shard_ivf_keeps_gops_whole():
    from alfalfa_b200 import multigpu as M
    data = open(os.path.join(GOLDEN_DIR, "2a4c049c2f8e3a19ee39ffd7074cecd68006a101"), "rb").read()
    _, gops = M.split_gops(data)
    assert len(gops) == 4 and sum(len(g) for g in gops) == 260
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            sub, mine = M.shard_ivf(data, rank, world)
            seen += mine
            _, _, frames = O.read_ivf(sub)
            assert len(frames) == sum(len(gops[g]) for g in mine)
            if frames:
                assert not (frames[0][0] & 1)  # every shard starts with a key frame
        assert sorted(seen) == list(range(len(gops)))


def test_reference_arm_runs_only_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


HANDOVER = r'''
import ctypes as C, hashlib, json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import oracle_lib as O
from alfalfa_b200 import capi, multigpu as M
from alfalfa_b200.decoder import DecoderState
rank, world, local, dist = M.init("gloo")
L = capi.lib()
w, h, frames = O.read_ivf(open(%(vec)r, "rb").read())
frames = [f for f in frames][:12]
SPLIT = 5
NB = 4096


class Raster:                       # host stand-in for a device raster: same export / import contract
    count = 0
    def __init__(self, fill=None):
        Raster.count += 1
        self.id = Raster.count
        self.buf = (C.c_uint8 * NB)(*([fill] * NB)) if fill is not None else (C.c_uint8 * NB)()
    def export_to(self, ptr, n): C.memmove(ptr, self.buf, n)
    def import_from(self, ptr, n): C.memmove(self.buf, ptr, n)
    def release(self): pass


class Ctx:
    frame_bytes = NB
    def alloc_frame(self): return Raster()


def records_digest(state, chunk_list):
    pf = C.c_void_p(); capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    sha = hashlib.sha1()
    for f in chunk_list:
        assert L.vp8gpu_parse_frame(state.h, f, len(f), pf) == 0
        d = L.vp8gpu_parsed_desc(pf).contents
        sha.update(bytes(d)); sha.update(C.string_at(L.vp8gpu_parsed_mbs(pf), d.mb_cols * d.mb_rows * 32))
    return sha.hexdigest()

straight = DecoderState(w, h)
records_digest(straight, frames[:SPLIT])
want_tail = records_digest(straight, frames[SPLIT:])


class Dec:                          # what broadcast_decoder needs from a Decoder
    def __init__(self, state, refs): self.state, self.refs = state, refs
    def get_state(self): return self.state
    def get_references(self): return self.refs

dec = None
if rank == 0:
    st = DecoderState(w, h)
    records_digest(st, frames[:SPLIT])
    shared = Raster(7)              # last and golden are one raster, alternative another
    dec = Dec(st, (shared, shared, Raster(9)))
got = M.broadcast_decoder(Ctx(), dec, 0, dist, local,
                          make_decoder=lambda c, blob, three: Dec(DecoderState.deserialize(blob), three))
tail = records_digest(got.get_state(), frames[SPLIT:])
refs = got.get_references()
os.write(1, (json.dumps({"rank": rank, "tail_ok": tail == want_tail, "state_equal": got.get_state() == straight,
                  "fills": [int(r.buf[0]) for r in refs], "shared": refs[0] is refs[1], "distinct": refs[2] is not refs[0]}) + "\n").encode())
dist.destroy_process_group()
'''


def test_two_rank_decoder_handover_over_gloo(tmp_path):
    """A GOP that continues on another rank: DecoderState blob + distinct reference rasters are broadcast
    (alfalfa_b200.multigpu.broadcast_decoder; NCCL on GPUs, gloo here with host stand-ins for rasters)."""
    import json
    vec = os.path.join(GOLDEN_DIR, "0b546dad90ddefea5085c7751b5fa2f117630b1c")
    script = tmp_path / "handover.py"
    script.write_text(HANDOVER % {"root": ROOT, "vec": vec})
    port = _free_port()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    rows = sorted((json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")), key=lambda r: r["rank"])
    assert len(rows) == 2
    for r in rows:
        assert r["tail_ok"] and r["state_equal"] and r["fills"] == [7, 7, 9] and r["shared"] and r["distinct"]


def test_state_blob_roundtrip_on_every_golden_vector():
    """vp8gpu_state_serialize / _deserialize (DecoderState::serialize, decoder.cc:266-330)"""
    import ctypes as C
    from alfalfa_b200 import capi
    from alfalfa_b200.decoder import DecoderState
    from conftest import golden_vectors
    L = capi.lib()
    for name in golden_vectors():
        w, h, frames = O.read_ivf(open(os.path.join(GOLDEN_DIR, name), "rb").read())
        st = DecoderState(w, h)
        pf = C.c_void_p()
        capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
        started = False
        for f in frames[:25]:
            if not started and (f[0] & 1):
                continue
            started = True
            assert L.vp8gpu_parse_frame(st.h, f, len(f), pf) == 0
        back = DecoderState.deserialize(st.serialize())
        assert back == st and back.hash() == st.hash()
        L.vp8gpu_parsed_destroy(pf)
    with pytest.raises(Exception):
        DecoderState.deserialize(b"nonsense")
