#!/usr/bin/env python3
"""Stage-by-stage bisection of the CUDA path against the oracle on a small input (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import mimi_oracle as O                      # noqa: E402  (debug tool == test infrastructure)
from tokenize_audio_b200 import synth                   # noqa: E402
from tokenize_audio_b200.encoder import MimiB200Model   # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def main():
    sd = synth.synth_state_dict(0)
    t0 = time.time()
    m = MimiB200Model(sd, "cuda:0")
    print("model load %.1fs" % (time.time() - t0), flush=True)
    N = 24000 + 4321
    x = np.stack([synth.synth_speech(11, N), synth.synth_speech(12, N)])[:, None, :]
    taps, margins = {}, []
    codes_o = O.encode(sd, x, 32, taps=taps, margins=margins)
    xd = torch.from_numpy(x).cuda()
    if any(a.startswith("--tc") for a in sys.argv):
        # tensor-core path: raw taps that exist there (down convs 3/6/9, stream z), then the tail
        mode = 1 if "--tc1" in sys.argv else 2 if "--tc2" in sys.argv else 4 if "--tc4" in sys.argv else 5 if "--tc5" in sys.argv else 6 if "--tc6" in sys.argv else 7 if "--tc7" in sys.argv else 8 if "--tc8" in sys.argv else 3
        m.set_mode(mode)
        for a in sys.argv:
            if a.startswith("--dbg="):          # --dbg=KEY:VALUE
                k, v = a[6:].split(":")
                m.debug_set(int(k), int(v))
        m.debug_set(0, 0)
        m.encode(xd, num_quantizers=32)
        torch.cuda.synchronize()
        for ci, nm in {0: "seanet.l0", 3: "seanet.down3", 6: "seanet.down6", 9: "seanet.down9", 13: "seanet.out"}.items():
            if mode >= 3 and ci == 0:
                continue
            t = m.debug_tap(ci).cpu().numpy()
            ref = np.stack([a.T for a in taps[nm]])
            t = t[:, : ref.shape[1]]
            print(f"[tc] conv {ci:2d} {nm:14s} rel {rel(t, ref):.3e} maxabs {np.abs(t - ref).max():.3e} shape {t.shape}", flush=True)
        for l in range(8):
            m.debug_set(0, l + 1)
            m.encode(xd, num_quantizers=32)
            torch.cuda.synchronize()
            t = m.debug_tap(100 + l).cpu().numpy()
            ref = np.stack(taps[f"transformer.layer{l}"])
            print(f"[tc] transformer layer {l} rel {rel(t[:, :ref.shape[1]], ref):.3e}", flush=True)
        m.debug_set(0, 8)
        out, lat = m.encode(xd, num_quantizers=32, return_latent=True)
        torch.cuda.synchronize()
        lat = lat.cpu().numpy()
        ref = np.stack(taps["latent"])
        print(f"[tc] latent rel {rel(lat, ref):.3e} maxabs {np.abs(lat - ref).max():.3e}")
        codes = out.audio_codes.cpu().numpy()
        agree = codes == codes_o
        print("[tc] codes agree all %.5f first8 %.5f" % (agree.mean(), agree[:, :8].mean()))
        print("[tc] per-codebook", np.round(agree.mean(axis=(0, 2)), 3))
        mg = np.stack(margins).reshape(2, 32, -1)
        for b, k, t in np.argwhere(~agree)[:20]:
            print("  mismatch item %d cb %d frame %d: gpu %d oracle %d  oracle top-2 rel margin %.2e" %
                  (b, k, t, codes[b, k, t], codes_o[b, k, t], mg[b, k, t]))
        return
    m.set_mode(False)
    # SEANet taps: run with the pipeline stopped after each conv so in-place buffers hold that conv's output
    names = {0: "seanet.l0", 2: "seanet.res1", 3: "seanet.down3", 5: "seanet.res4", 6: "seanet.down6",
             8: "seanet.res7", 9: "seanet.down9", 11: "seanet.res10", 12: "seanet.down12", 13: "seanet.out"}
    for ci, nm in names.items():
        m.debug_set(1, ci)
        m.debug_set(0, 0)
        m.encode(xd, num_quantizers=32)
        torch.cuda.synchronize()
        t = m.debug_tap(ci).cpu().numpy()              # [B, rows, C]
        ref = np.stack([a.T for a in taps[nm]])        # [B, L, C]
        t = t[:, : ref.shape[1]]
        print(f"conv {ci:2d} {nm:14s} rel {rel(t, ref):.3e} maxabs {np.abs(t - ref).max():.3e} shape {t.shape}", flush=True)
    m.debug_set(1, 13)
    for l in range(8):
        m.debug_set(0, l + 1)
        m.encode(xd, num_quantizers=32)
        torch.cuda.synchronize()
        t = m.debug_tap(100 + l).cpu().numpy()
        ref = np.stack(taps[f"transformer.layer{l}"])
        print(f"transformer layer {l} rel {rel(t[:, :ref.shape[1]], ref):.3e} maxabs {np.abs(t[:, :ref.shape[1]] - ref).max():.3e}", flush=True)
    m.debug_set(0, 8)
    out, lat = m.encode(xd, num_quantizers=32, return_latent=True)
    torch.cuda.synchronize()
    lat = lat.cpu().numpy()
    ref = np.stack(taps["latent"])
    print(f"latent rel {rel(lat, ref):.3e} maxabs {np.abs(lat - ref).max():.3e}")
    codes = out.audio_codes.cpu().numpy()
    agree = codes == codes_o
    print("codes agree all %.5f first8 %.5f" % (agree.mean(), agree[:, :8].mean()))
    print("per-codebook", np.round(agree.mean(axis=(0, 2)), 3))
    bad = np.argwhere(~agree)
    mg = np.stack(margins).reshape(2, 32, -1)          # [B][K][T]
    for b, k, t in bad[:20]:
        print("  mismatch item %d cb %d frame %d: gpu %d oracle %d  oracle top-2 rel margin %.2e" %
              (b, k, t, codes[b, k, t], codes_o[b, k, t], mg[b, k, t]))
    print("launches", m.launch_count)


if __name__ == "__main__":
    main()
