#!/usr/bin/env python3
"""NaN-poisoned-workspace check (see tests/test_gpu_modes.py::test_nothing_is_read_before_it_is_written) for the opt-in
kernel generations: mode 8 and the fused level-1 residual block (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model, MimiEncoder
m = MimiB200Model(synth.synth_state_dict(0), "cuda:0")
rng = np.random.default_rng(7)
clips = [synth.synth_speech(1000 + i, int(n)) for i, n in enumerate(rng.integers(3000, 60000, size=21))]
lens = [len(c) for c in clips]
x = np.zeros((21, 1, max(lens)), np.float32)
for i, c in enumerate(clips):
    x[i, 0, : len(c)] = c
xd = torch.from_numpy(x).cuda()
def poison():
    for w in m._workspaces.values():
        if w is not None:
            w.view(torch.float32)[: w.numel() // 4].fill_(float("nan"))
    torch.cuda.synchronize()
for tag, setup in (("mode 7", lambda: m.set_mode(7)), ("mode 7 + fused resblock", lambda: (m.set_mode(7), m.debug_set(15, 1))),
                   ("mode 8", lambda: (m.debug_set(15, 0), m.set_mode(8))), ("mode 6", lambda: m.set_mode(6)), ("mode 3", lambda: m.set_mode(3))):
    setup()
    clean = m.encode(xd, num_quantizers=8, valid_lengths=lens).audio_codes.cpu().numpy()
    poison()
    again = m.encode(xd, num_quantizers=8, valid_lengths=lens).audio_codes.cpu().numpy()
    w = MimiEncoder(m, num_quantizers=8, first_items=1)
    poison()
    res = w.encode_audio_batch(clips)
    bad = [i for i, a in enumerate(res) if not np.array_equal(a, clean[i, :, : a.shape[1]])]
    print(tag, "plain on NaN workspace:", "ok" if np.array_equal(again, clean) else "MISMATCH", "| phased:", "ok" if not bad else f"MISMATCH {bad}", flush=True)
