#!/usr/bin/env python3
"""End-to-end (host buffers in, host codes out) throughput of MimiEncoder.encode_audio_batch on the bench workload for a
few staging configurations (run under gpurun)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model, MimiEncoder

sd = synth.synth_state_dict(0)
model = MimiB200Model(sd, device="cuda:0")
clips, lengths, batches = bench.make_workload(0)
nmax = max(len(c) for cl in clips for c in cl)
model.reserve_workspace(64, nmax, 8)
audio = sum(len(c) for cl in clips for c in cl) / 24000
for chunk_items, threads, first, phased, two in [(16, 8, 8, True, False), (16, 8, 8, True, True), (16, 8, 4, True, True), (16, 8, 16, True, False), (16, 8, 8, True, False)]:
    w = MimiEncoder(model, ragged=True, num_quantizers=8, chunk_items=chunk_items, stage_threads=threads, first_items=first)
    w.phased = phased
    w.two_ranges = two
    w.pack_threads = threads
    w.reserve(64, nmax)
    for i in range(3):
        w.encode_audio_batch(clips[i])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        for i in range(8):
            w.encode_audio_batch(clips[i])
        best = min(best, time.perf_counter() - t)
    print(f"two_ranges={two} phased={phased} chunk_items={chunk_items} threads={threads} first={first}: {1e3*best/8:.2f} ms/step  {audio/best:.0f} x RT", flush=True)
