"""World-size-2 gloo run of the multi-GPU host logic: batches dealt round-robin, counters all-reduced."""
import os
import subprocess
import sys

from conftest import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["REPO"])
import torch, torch.distributed as dist
from tokenize_audio_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lens = [(i * 7919) % 400000 + 48000 for i in range(37)]
batches = sharding.bucket_batches(lens, 8)
mine = sharding.shard_for_rank(len(batches), rank, world)
items = [i for b in mine for i in batches[b]]
c = sharding.reduce_counters({"items": len(items), "samples": float(sum(lens[i] for i in items)), "elapsed_max": 1.0 + rank})
if rank == 0:
    assert c["items"] == 37, c
    assert c["samples"] == float(sum(lens)), c
    assert c["elapsed_max"] == float(world), c
    print("OK", c)
dist.destroy_process_group()
"""


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout
