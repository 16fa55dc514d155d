"""A/B of the end-to-end path on one box (diagnostic): resident encode vs MimiEncoder.encode_stream with and without per-slot
streams, plus the host time spent inside submit() and result(). Usage: python tools/e2e_ab.py [mode]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model, MimiEncoder

bench.select_workload("c2")
model = MimiB200Model(synth.synth_state_dict(0), device="cuda:0")
if len(sys.argv) > 1:
    model.set_mode(int(sys.argv[1]))
clips, lengths, batches = bench.make_workload(0)
audio = sum(sum(len(c) for c in b) for b in clips) / 24000
nmax = max(len(c) for b in clips for c in b)
dev = []
for cl in clips:
    n = max(len(c) for c in cl)
    x = torch.zeros(len(cl), 1, n)
    for i, c in enumerate(cl):
        x[i, 0, :len(c)] = torch.from_numpy(c)
    dev.append((x.cuda(), [len(c) for c in cl]))
model.reserve_workspace(64, nmax, 8)


def resident():
    for x, l in dev:
        model.encode(x, num_quantizers=8, valid_lengths=l)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(2):
        for x, l in dev:
            model.encode(x, num_quantizers=8, valid_lengths=l)
    torch.cuda.synchronize()
    return 2 * audio / (time.perf_counter() - t)


def stream(slot_streams, threads=None):
    w = MimiEncoder(model, num_quantizers=8)
    w.slot_streams = slot_streams
    if threads:
        w.pack_threads = threads
    w.reserve(64, nmax)
    list(w.encode_stream(clips[:3]))
    torch.cuda.synchronize()
    ts = tr = 0.0
    t = time.perf_counter()
    pend = []
    for _ in range(2):
        for cl in clips:
            t0 = time.perf_counter(); pend.append(w.submit(cl)); ts += time.perf_counter() - t0
            if len(pend) >= 2:
                t0 = time.perf_counter(); w.result(pend.pop(0)); tr += time.perf_counter() - t0
    while pend:
        t0 = time.perf_counter(); w.result(pend.pop(0)); tr += time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return 2 * audio / dt, 1e3 * ts / 16, 1e3 * tr / 16


for rep in range(2):
    print(f"resident            {resident():9.0f} x RT", flush=True)
    for ss in (False, True):
        v, ts, tr = stream(ss)
        print(f"stream slot_streams={ss!s:5} {v:9.0f} x RT   submit {ts:.2f} ms  result {tr:.2f} ms per batch", flush=True)
    v, ts, tr = stream(True, 4)
    print(f"stream slot_streams=True, 4 pack threads {v:9.0f} x RT   submit {ts:.2f} ms  result {tr:.2f} ms", flush=True)
