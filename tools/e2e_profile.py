#!/usr/bin/env python3
"""Host-side timing of the ragged path of MimiEncoder.encode_audio_batch on the bench workload (run under gpurun)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model, MimiEncoder, FRAME_SIZE

sd = synth.synth_state_dict(0)
model = MimiB200Model(sd, device="cuda:0")
sizes = [int(v) for v in os.environ.get("CHUNKS", "16,16,16,16").split(",")]
w = MimiEncoder(model, ragged=True, num_quantizers=8)
clips, lengths, batches = bench.make_workload(0)
nmax = max(len(c) for cl in clips for c in cl)
w.reserve(64, nmax)
model.reserve_workspace(64, nmax, 8)
for i in range(3):
    w.encode_audio_batch(clips[i])
torch.cuda.synchronize()
T = dict(fill=0.0, h2d=0.0, enc=0.0, sync=0.0, asm=0.0, gpu=0.0)
dev = model.device
cs = torch.cuda.Stream(device=dev)
n = 8
t_all = time.perf_counter()
for it in range(n):
    arrs = clips[it % 8]
    L = [len(a) for a in arrs]
    chunks, c = [], 0
    for s in sizes:
        chunks.append(list(range(c, min(c + s, 64)))); c += s
    chunks = [ch for ch in chunks if ch]
    off = coff = 0
    hcs = []
    main = torch.cuda.current_stream(dev)
    cs.wait_stream(main)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for ch in chunks:
        N = max(L[i] for i in ch); Tt = -(-N // FRAME_SIZE)
        t = time.perf_counter()
        buf = w._pinned_view(off, len(ch), N); off += len(ch) * N
        lens = [L[i] for i in ch]
        w._fill(buf, [arrs[i] for i in ch], [min(N, -(-x // FRAME_SIZE) * FRAME_SIZE) for x in lens])
        T["fill"] += time.perf_counter() - t; t = time.perf_counter()
        with torch.cuda.stream(cs):
            x = buf.to(dev, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs)
        main.wait_event(ev); x.record_stream(main)
        T["h2d"] += time.perf_counter() - t; t = time.perf_counter()
        out = model.encode(input_values=x, num_quantizers=8, valid_lengths=lens)
        hc = w._pinned_codes[coff: coff + len(ch) * 8 * Tt].view(len(ch), 8, Tt); coff += len(ch) * 8 * Tt
        hc.copy_(out.audio_codes, non_blocking=True)
        hcs.append(hc)
        T["enc"] += time.perf_counter() - t
    e1.record(main)
    t = time.perf_counter()
    main.synchronize()
    T["sync"] += time.perf_counter() - t; t = time.perf_counter()
    res = []
    for ch, hc in zip(chunks, hcs):
        a = hc.numpy()
        for j, i in enumerate(ch):
            res.append(a[j, :, : int(np.ceil(L[i] / 1920.0))].copy())
    T["asm"] += time.perf_counter() - t
    T["gpu"] += e0.elapsed_time(e1) / 1e3
tot = time.perf_counter() - t_all
audio = sum(sum(len(c) for c in clips[i % 8]) for i in range(n)) / 24000
print(f"chunks={sizes}: total {1e3*tot/n:.1f} ms/step xRT {audio/tot:.0f} | " + " ".join(f"{k} {1e3*v/n:.1f}" for k, v in T.items()))
