"""A/B of debug knobs on one box (diagnostic): the resident C2 step (two item ranges on two streams, CUDA events) and the
per-kernel profile of a single-stream pass, for the default build and for each `KEY=VALUE[,KEY=VALUE]` argument.
Usage: python tools/ab_knobs.py 17=1 18=1 17=1,18=1 balance_ranges=0"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model

bench.select_workload(os.environ.get("AB_WORKLOAD", "c2"))
K = bench.K_CODEBOOKS
model = MimiB200Model(synth.synth_state_dict(0), device="cuda:0")
clips, lengths, batches = bench.make_workload(0)
audio = sum(sum(len(c) for c in b) for b in clips) / 24000
dev = []
for cl in clips:
    n = max(len(c) for c in cl)
    x = torch.zeros(len(cl), 1, n)
    for i, c in enumerate(cl):
        x[i, 0, :len(c)] = torch.from_numpy(c)
    dev.append((x.cuda(), [len(c) for c in cl]))
model.reserve_workspace(max(len(c) for c in clips), max(len(c) for b in clips for c in b), K)


def one_pass():
    for x, l in dev:
        model.encode(x, num_quantizers=K, valid_lengths=l)


def measure(reps=3):
    one_pass()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_pass()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    model.streams = 1
    one_pass()
    model.profile(True)
    one_pass()
    prof = model.profile_read()
    model.profile(False)
    model.streams = 2
    kinds = {}
    for k, (ms, cnt) in prof.items():
        kinds[k] = round(ms / len(dev), 4)
    return {"ms_per_batch": round(best / len(dev), 4), "x_realtime": round(audio / (best / 1e3), 1), "per_kernel_ms_per_batch": kinds}


settings = [""] + sys.argv[1:]
for s in settings:
    # KEY=VALUE -> debug_set(KEY, VALUE); NAME=VALUE with a non-numeric NAME -> setattr(model, NAME, VALUE) for this measurement
    pairs = [tuple(int(v) for v in kv.split("=")) for kv in s.split(",") if kv and kv.split("=")[0].isdigit()]
    attrs = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv and not kv.split("=")[0].isdigit()]
    old = [(a, getattr(model, a)) for a, _ in attrs]
    for k, v in pairs:
        model.debug_set(k, v)
    for a, v in attrs:
        setattr(model, a, type(getattr(model, a))(v))
    try:
        print(json.dumps({"knobs": s or "default", **measure()}), flush=True)
    finally:
        for k, v in pairs:
            model.debug_set(k, 0)
        for a, v in old:
            setattr(model, a, v)
