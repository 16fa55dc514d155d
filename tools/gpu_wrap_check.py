#!/usr/bin/env python3
"""Diagnostic behind tests/test_gpu_modes.py::test_nothing_is_read_before_it_is_written: the workspace is re-allocated and
filled with NaN before a plain encode and before phased wrapper calls with several front-end groups; prints the items whose
codes differ from the clean result (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model, MimiEncoder
m = MimiB200Model(synth.synth_state_dict(0), "cuda:0")
m.streams = 1
rng = np.random.default_rng(7)
clips = [synth.synth_speech(1000 + i, int(n)) for i, n in enumerate(rng.integers(3000, 60000, size=21))]
n = max(len(c) for c in clips)
x = np.zeros((21, 1, n), np.float32)
for i, c in enumerate(clips):
    x[i, 0, : len(c)] = c
lens = [len(c) for c in clips]
xd = torch.from_numpy(x).cuda()
plain = m.encode(xd, num_quantizers=8, valid_lengths=lens).audio_codes.cpu().numpy()
grow = [30]
def poison():
    # new, larger workspace filled with NaN: anything read before it is written shows up
    grow[0] += 4
    m.reserve_workspace(grow[0], n, 8)
    m._workspaces[0].view(torch.float32)[: m._workspaces[0].numel() // 4].fill_(float("nan"))
    torch.cuda.synchronize()
def check(tag, **kw):
    w = MimiEncoder(m, num_quantizers=8, **kw)
    res = w.encode_audio_batch(clips)
    bad = [i for i, a in enumerate(res) if not np.array_equal(a, plain[i, :, : a.shape[1]])]
    print(tag, bad, flush=True)
poison(); p = m.encode(xd, num_quantizers=8, valid_lengths=lens).audio_codes.cpu().numpy(); print("plain on NaN workspace equal", np.array_equal(p, plain), flush=True)
poison(); check("phased first1 on NaN ws", first_items=1)
poison(); check("phased first64 on NaN ws", first_items=64)
m.debug_set(13, 1); poison(); check("phased first1, no tile lists", first_items=1); m.debug_set(13, 0)
m.debug_set(10, 1); poison(); check("phased first1, no flat", first_items=1); m.debug_set(10, 0)
