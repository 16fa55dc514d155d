"""Diagnostic: the wrapper's phased batch path against plain encode() calls, per debug knob (which frames differ)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model, MimiEncoder

model = MimiB200Model(synth.synth_state_dict(0), device="cuda:0")
audio = [synth.synth_speech(70, 20000), synth.synth_speech(71, 33333), synth.synth_speech(72, 5000)]
n = max(len(a) for a in audio)
x = np.zeros((3, 1, n), np.float32)
for i, a in enumerate(audio):
    x[i, 0, : len(a)] = a
xd = torch.from_numpy(x).cuda()
lens = [len(a) for a in audio]
model.debug_set(17, 1); model.debug_set(18, 1)
base = model.encode(xd, num_quantizers=32).audio_codes.cpu().numpy()
model.debug_set(17, 0); model.debug_set(18, 0)
for knobs in ((), (17,), (18,), (17, 18)):
    for k in knobs:
        model.debug_set(k, 1)
    strict = model.encode(xd, num_quantizers=32).audio_codes.cpu().numpy()
    rag = model.encode(xd, num_quantizers=32, valid_lengths=lens).audio_codes.cpu().numpy()
    enc = MimiEncoder(model)
    enc.encode_audio_chunk(audio[0])
    wb = enc.encode_audio_batch(audio)
    wb2 = enc.encode_audio_batch(audio)
    for k in knobs:
        model.debug_set(k, 0)
    msg = [f"knobs {knobs}: strict!=base {int((strict != base).sum())}"]
    for i, a in enumerate(lens):
        t = -(-a // 1920)
        msg.append(f"item{i}: rag {np.argwhere((rag[i,:,:t] != base[i,:,:t]).any(0)).ravel().tolist()} "
                   f"wrap {np.argwhere((wb[i] != base[i,:,:t]).any(0)).ravel().tolist()} "
                   f"wrap2 {np.argwhere((wb2[i] != base[i,:,:t]).any(0)).ravel().tolist()}")
    print(" | ".join(msg), "fallbacks", enc.range_fallbacks, flush=True)
