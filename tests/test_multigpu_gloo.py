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
print(json.dumps({"rank": rank, "gops": mine, "frames": len(frames), "sha": hashlib.sha1(out).hexdigest(),
                  "agg": agg, "total_frames": total_frames, "units": units}))
dist.destroy_process_group()
'''


def test_two_rank_gop_sharding_and_reductions(tmp_path):
    vec = os.path.join(GOLDEN_DIR, "45502fe01a62b82d498b83dc50824741402436db")  # 30 key frames = 30 GOPs
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "vec": vec})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
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
