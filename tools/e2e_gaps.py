#!/usr/bin/env python3
"""Where the end-to-end time of encode_audio_batch goes: wall time per step vs the sum of kernel times (per-launch CUDA
events) in the same run, next to the same batches encoded from resident inputs (run under gpurun)."""
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
w = MimiEncoder(model, ragged=True, num_quantizers=8, first_items=int(os.environ.get("FIRST", "8")))
w.reserve(64, nmax)
dev_batches = []
for cl in clips:
    n = max(len(c) for c in cl)
    x = torch.zeros((len(cl), 1, n))
    for i, c in enumerate(cl):
        x[i, 0, : len(c)] = torch.from_numpy(c)
    dev_batches.append((x.cuda(), [len(c) for c in cl]))
for i in range(3):
    w.encode_audio_batch(clips[i])
torch.cuda.synchronize()
for prof in (False, True):
    for name in ("resident", "e2e"):
        model.profile(prof)
        t = time.perf_counter()
        for i in range(8):
            if name == "e2e":
                w.encode_audio_batch(clips[i])
            else:
                model.encode(dev_batches[i][0], num_quantizers=8, valid_lengths=dev_batches[i][1])
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t) / 8 * 1e3
        msg = f"{name:9s} profile={prof}: wall {wall:.2f} ms/step"
        if prof:
            pr = model.profile_read()
            msg += f"  kernels {sum(v[0] for v in pr.values()) / 8:.2f} ms/step  front {pr['front_fused'][0] / 8:.2f} ({pr['front_fused'][1] // 8} launches)"
        print(msg, flush=True)
        model.profile(False)
