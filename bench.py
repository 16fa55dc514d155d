#!/usr/bin/env python3
"""Benchmark of the Mimi encode hot path (BASELINE.json metric: audio-seconds encoded per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|reference-cuda] [--workload c1..c5|resample|utf8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Default workload (config.workload = "c2"): BASELINE.json configs[1] -- Mimi encode, 8 codebooks, batch 64 of 2-20 s
utterances, length-bucketed (2 s buckets), padded to the longest item of the batch. A pool of 512 synthetic
speech-shaped utterances (durations U(2,20) s, fixed seed) is bucketed into 8 batches of 64; step i encodes
batch i % 8. One "step" = one batch. Every rank works on its own shard (same length multiset, different
audio), no collective on the data path (weak scaling); counters meet in one all_reduce at the end.
The other BASELINE configs are extra measured lines (--workload c1 / c3 / c4 / c5), not what the driver runs.

Printed JSON (rank 0, one line):
  value   audio-seconds per second with the padded batches already resident in HBM (device timed)
  e2e     the same from HOST numpy clips through the wrapper's pipelined call MimiEncoder.encode_stream (submit/result with
          two batches in flight): pinned staging + H2D of every step's input and D2H of its codes inside the timed region.
          e2e_single_call is the strictly synchronous MimiEncoder.encode_audio_batch (the reference's own call shape).
  roofline      dominant kernel, timed live with CUDA events inside the timed region
  cpu_baseline  transformers.MimiModel (the reference's own implementation) on this box's host cores on a
                bounded sample of the same workload (rank 0, N=1 only)
  agreement     codes of the CUDA path vs that CPU run on the same items: overall, per codebook, and every mismatch classified
                with the REFERENCE's own top-2 distance margin (near-tie < 1e-3, cascade of an earlier flip, or unexplained)

--impl reference times the CPU implementation as its own arm (bounded sample per step); --impl reference-cuda is an
informational line: the same stock transformers model on the B200 (what a user of the reference gets today).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tokenize_audio_b200 import sharding, synth  # noqa: E402

SR = 24000
K_CODEBOOKS = 8
BATCH = 64
POOL = 512
N_BATCHES = POOL // BATCH
SEED = 1234 + 2000            # seed = 1234 + config_id*1000 (+ item index)
REF_ITEMS_PER_STEP = 16       # bounded CPU sample per step for the reference arm (every (BATCH/16)-th item of the batch)
CPU_BASELINE_ITEMS = 24       # bounded CPU sample for the cpu_baseline leg
NEAR_TIE = 1e-3

# algorithmic MACs per audio-second of each launch kind (K = 8 codebooks)
MMAC_PER_AUDIO_S = {
    "conv0": 10.752, "seanet_conv1": 147.456, "seanet_conv2": 49.152, "seanet_conv3": 393.216,
    "seanet_conv4": 147.456, "seanet_conv5": 49.152, "seanet_conv6": 393.216, "seanet_conv7": 117.965,
    "seanet_conv8": 39.322, "seanet_conv9": 314.573, "seanet_conv10": 78.643, "seanet_conv11": 26.214,
    "seanet_conv12": 209.715, "seanet_conv13": 39.322, "qkv_gemm": 157.286, "o_proj": 52.429,
    "fc1_gelu": 209.715, "fc2": 209.715, "attention": 51.2, "downsample_conv": 13.107,
    "rvq_input_proj": 3.277, "rvq_fused": 6.5536 * K_CODEBOOKS,
    "front_fused": 10.752 + 147.456 + 49.152,
}
# algorithmic HBM bytes per audio-second for the bandwidth-bound kinds (fp32, channels-last)
HBM_BYTES_PER_AUDIO_S = {
    "conv0": 24000 * 4 + 24000 * 64 * 4,
    "layernorm": 25 * 512 * 4 * 2 * 16,
    # fused 24 kHz front end: waveform in, hi/lo split of the 64-channel activation out
    "front_fused": 24000 * 4 + 24000 * 64 * 4,
}
HBM_BOUND_KINDS = ("conv0", "layernorm", "front_fused")
# level-1 GEMM layers in the default generation (fp16 pairs: 4 B per split element, raw skip tensor fp32): operand in + outputs
LEVEL1_HBM_BYTES_PER_AUDIO_S = {
    "seanet_conv3": 24000 * 64 * 4 + 6000 * 128 * (4 + 4),             # D1: h1 pair in, d1 raw + ELU'd pair out
    "seanet_conv4": 6000 * 128 * 4 + 6000 * 64 * 4,                    # R2a: d1 pair in, r2 pair out
    "seanet_conv5": 6000 * 64 * 4 + 6000 * 128 * 4 + 6000 * 128 * 4,   # R2b: r2 pair in, d1 raw (skip) in, h2 pair out
}


# stdout carries exactly ONE line, the JSON result: everything else a library prints there (NCCL's version banner, a
# stray warning) is sent to stderr by pointing fd 1 at fd 2 for the whole run and writing the line to the saved fd
_REAL_STDOUT = None


def guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: str):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(line, flush=True)
    else:
        os.write(_REAL_STDOUT, (line + "\n").encode())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


# BASELINE.json configs. The default bench line is c2 (the config the metric is quoted on).
WORKLOADS = {
    "c1": {"codebooks": 8, "batch": 1, "pool": 8, "dur": (10.0, 10.0), "sr_in": 16000,
           "desc": "one 10 s utterance at 16 kHz per step, resampled to 24 kHz on the GPU, batch 1 (latency case), 8 clips cycled"},
    "c2": {"codebooks": 8, "batch": 64, "pool": 512, "dur": (2.0, 20.0), "desc": "U(2,20) length-bucketed, 8 batches cycled"},
    "c3": {"codebooks": 8, "batch": 16, "pool": 128, "dur": (30.0, 30.0), "desc": "30 s long-form segments, batch 16, 8 batches cycled"},
    "c4": {"codebooks": 32, "batch": 32, "pool": 256, "dur": (15.0, 15.0), "desc": "15 s items, batch 32, all 32 codebooks, 8 batches cycled"},
    "c5": {"codebooks": 8, "batch": 16, "pool": 128, "dur": (30.0, 30.0), "strings": True,
           "desc": "30 s long-form segments, batch 16, codes -> codes_to_chars UTF-8 strings on the GPU, 8 batches cycled"},
}
WORKLOAD = "c2"
ORDER = "bucketed"            # --order file: batches in pool (file) order, no length bucketing (what the reference scripts do)
STRICT = False                # --strict: every item encoded over the padded batch length (what the reference computes)
DESC = WORKLOADS["c2"]["desc"]
DUR = WORKLOADS["c2"]["dur"]
SR_IN = SR
STRINGS = False


def select_workload(name: str) -> None:
    global WORKLOAD, K_CODEBOOKS, BATCH, POOL, N_BATCHES, DESC, DUR, SR_IN, STRINGS, SEED
    w = WORKLOADS[name]
    WORKLOAD, K_CODEBOOKS, BATCH, POOL, DESC, DUR = name, w["codebooks"], w["batch"], w["pool"], w["desc"], w["dur"]
    SR_IN = w.get("sr_in", SR)
    STRINGS = bool(w.get("strings"))
    N_BATCHES = POOL // BATCH
    SEED = 1234 + 1000 * int(name[1])
    MMAC_PER_AUDIO_S["rvq_fused"] = 6.5536 * K_CODEBOOKS


def make_workload(rank: int):
    """POOL utterance lengths (U(2,20) s for c2) -> 8 length-bucketed batches of BATCH (lists of numpy clips at SR_IN)."""
    rng = np.random.Generator(np.random.PCG64(SEED))
    lengths = [int(v) for v in rng.uniform(DUR[0], DUR[1], size=POOL) * SR_IN]
    if ORDER == "file":
        # "the un-bucketed file-order variant the reference uses" -- consecutive items form a batch
        # (REF/emilia-mimi/process_shard.py:479-510 batches files as they come), so every batch is padded to ~ the longest clip
        batches = [list(range(i, min(i + BATCH, POOL))) for i in range(0, POOL, BATCH)]
    else:
        batches = sharding.bucket_batches(lengths, BATCH, bucket_width=2 * SR_IN)
    # 12 base clips per rank; utterance i = a crop of base clip i % 12 (content does not affect timing)
    top = int(DUR[1] * SR_IN)
    base = [synth.synth_speech(SEED + 100 * rank + j, top, sr=SR_IN) for j in range(12)]
    clips = []
    for b in batches:
        clips.append([base[i % 12][: lengths[i]] if (i // 12) % 2 == 0 else base[i % 12][top - lengths[i]:] for i in b])
    return clips, lengths, batches


def bench_config(extra=None):
    desc = DESC.replace("length-bucketed", "in file order (not bucketed)") if ORDER == "file" else DESC
    mode = ("strict (every item encoded over the padded batch length, exactly the reference's computation)" if STRICT
            else "ragged (padded tails skipped, kept frames identical)")
    cfg = {"workload": WORKLOAD, "codebooks": K_CODEBOOKS, "batch": BATCH, "durations_s": desc,
           "mode": mode, "l2": "inputs+activations per step >> 126 MB L2",
           "weights": "synthetic seed 0 (kyutai/mimi architecture)"}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---- the reference implementation (transformers.MimiModel), CPU or CUDA --------------------------------------------------

def reference_model(sd, device="cpu"):
    from transformers import MimiConfig, MimiModel
    model = MimiModel(MimiConfig()).eval()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    return model.to(device)


def reference_encoder(sd, device="cpu", model=None):
    """The reference path unchanged: transformers.MimiModel fp32, driven the way MimiEncoder.encode_audio_batch does
    (REF/emilia-mimi/process_shard.py:88-140): feature extractor pads, model.encode(input_values, padding_mask), trim."""
    from transformers import EncodecFeatureExtractor
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    model = model if model is not None else reference_model(sd, device)
    fe = EncodecFeatureExtractor()

    def encode_batch(clips):
        with torch.no_grad():
            inputs = fe(raw_audio=clips, sampling_rate=SR, return_tensors="pt", padding=True)
            codes = model.encode(input_values=inputs["input_values"].to(device), padding_mask=inputs["padding_mask"].to(device),
                                 num_quantizers=K_CODEBOOKS).audio_codes
            codes = codes.cpu()
            return [codes[i, :, : int(np.ceil(len(c) / 1920.0))].numpy() for i, c in enumerate(clips)]
    encode_batch.model = model
    return encode_batch


def reference_margins(model, clip, K):
    """Top-2 relative distance margins (d2 - d1) / d1 of every argmin the REFERENCE takes on this clip, from its own
    modules along its own residual chain (modeling_mimi.py:1197-1202, 1262-1280): [K, T]."""
    with torch.no_grad():
        iv = torch.from_numpy(np.asarray(clip, np.float32))[None, None, :]
        emb = model.encoder(iv)
        z = model.encoder_transformer(emb.transpose(1, 2))[0].transpose(1, 2)
        latent = model.downsample(z)
        q = model.quantizer
        T = latent.shape[2]
        out = []
        for rvq, n in ((q.semantic_residual_vector_quantizer, 1), (q.acoustic_residual_vector_quantizer, K - 1)):
            r = rvq.input_proj(latent)
            for layer in rvq.layers[:n]:
                e = layer.codebook.embed
                x = r.permute(0, 2, 1).reshape(-1, e.shape[1])
                d = torch.cdist(x[None], e[None], p=2)[0]
                top2 = d.topk(2, largest=False).values
                idx = d.argmin(-1)
                out.append(((top2[:, 1] - top2[:, 0]) / top2[:, 0].clamp_min(1e-30)))
                r = r - torch.nn.functional.embedding(idx, e).view(1, T, -1).permute(0, 2, 1)
    return torch.stack(out, 0).numpy()


def agreement_report(got, ref, clips=None, model=None, vs="transformers.MimiModel CPU fp32, same items"):
    """codes of the CUDA path vs the reference on the same items: overall fraction, per-codebook fraction, and the list of
    mismatching slots, each classified with the reference's own margin (computed only for items that have a mismatch)."""
    K = ref[0].shape[0]
    same = sum(int((a == b).sum()) for a, b in zip(got, ref))
    tot = sum(a.size for a in ref)
    per_cb = [float(sum(int((a[k] == b[k]).sum()) for a, b in zip(got, ref)) / max(1, sum(b.shape[1] for b in ref))) for k in range(K)]
    flips, unexplained = [], 0
    for i, (a, b) in enumerate(zip(got, ref)):
        bad = a != b
        if not bad.any():
            continue
        mg = reference_margins(model, clips[i], K) if (model is not None and clips is not None) else None
        for k, t in np.argwhere(bad):
            first_bad = int(np.argmax(bad[:, t]))
            margin = float(mg[k, t]) if mg is not None else None
            kind = "cascade" if k > first_bad else ("near-tie" if (margin is not None and margin < NEAR_TIE) else "unexplained")
            unexplained += kind == "unexplained"
            if len(flips) < 64:
                flips.append({"item": i, "codebook": int(k), "frame": int(t), "reference_margin": margin, "kind": kind})
    return {"codes_equal_frac": same / max(tot, 1), "slots": tot, "per_codebook": [round(v, 6) for v in per_cb],
            "mismatches": tot - same, "flips": flips, "unexplained": int(unexplained), "near_tie_threshold": NEAR_TIE, "vs": vs}


def ref_sample(batch_clips):
    """The bounded per-step sample of the reference arm: every (B/16)-th item of the batch (the batch's own length mix)."""
    step = max(1, len(batch_clips) // REF_ITEMS_PER_STEP)
    return batch_clips[::step][:REF_ITEMS_PER_STEP]


def to24k_cpu(clips):
    """16 kHz clips of c1 -> 24 kHz on the host for the CPU reference arm (scipy polyphase; the reference's librosa/soxr is
    not installed, and the arm times the model, not the resampler)."""
    if SR_IN == SR:
        return clips
    from scipy.signal import resample_poly
    import math
    g = math.gcd(SR, SR_IN)
    return [resample_poly(c.astype(np.float64), SR // g, SR_IN // g).astype(np.float32) for c in clips]


def run_reference(args, rank, world, device="cpu"):
    if rank != 0:
        return
    sd = synth.synth_state_dict(0)
    clips, lengths, batches = make_workload(0)
    cuda = device != "cpu"
    tf32_modes = [None]
    try:
        enc = reference_encoder(sd, device)
    except Exception as e:  # transformers missing on this box
        emit(json.dumps({"impl": "reference" if not cuda else "reference-cuda", "unavailable": f"transformers MimiModel not importable: {e}"}))
        return
    cores = os.cpu_count() or 1
    samples = [to24k_cpu(ref_sample(c)) for c in clips]

    def step(i):
        sample = samples[i % N_BATCHES]
        enc(sample)
        return sum(len(c) for c in sample) / SR

    def timed():
        for i in range(args.warmup):
            step(i)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        audio = sum(step(i) for i in range(args.steps))
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return audio / dt, dt

    n_s = len(samples[0])
    sample_txt = (f"{n_s} of the {BATCH} items of each step's batch (every {max(1, BATCH // REF_ITEMS_PER_STEP)}th item of the "
                  f"{'length-sorted' if ORDER == 'bucketed' else 'file-order'} batch, one forward of batch {n_s})" if n_s < BATCH else f"the full {BATCH}-item batch per step")
    if not cuda:
        val, dt = timed()
        emit(json.dumps({
            "impl": "reference", "metric": "audio_seconds_encoded_per_sec", "value": val, "unit": "x_realtime",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(),
            "sample": sample_txt,
            "cpu_baseline": {"value": val, "unit": "x_realtime", "cores": cores, "kind": "reference",
                             "sample": f"{sample_txt}, transformers.MimiModel fp32 CPU, {cores} torch threads"},
            "e2e": {"value": val, "unit": "x_realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return
    # informational: the stock reference on the B200 itself (REF/emilia-mimi/process_shard.py:53-60 runs device="cuda"),
    # cudnn.allow_tf32 at its default (True) and off, with the code agreement of each against the CPU run
    cpu_enc = reference_encoder(sd, "cpu")
    probe = samples[N_BATCHES // 2][:8]
    ref_codes = cpu_enc(probe)
    res = {}
    for name, tf32 in (("allow_tf32_default", True), ("allow_tf32_off", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False if not tf32 else torch.backends.cuda.matmul.allow_tf32
        val, dt = timed()
        got = enc(probe)
        rep = agreement_report(got, ref_codes, vs="transformers.MimiModel CPU fp32, same items")
        res[name] = {"value": val, "ms_per_step": 1e3 * dt / max(args.steps, 1), "codes_equal_frac": rep["codes_equal_frac"],
                     "per_codebook": rep["per_codebook"], "slots": rep["slots"]}
    best = res["allow_tf32_off"]
    emit(json.dumps({
        "impl": "reference-cuda", "metric": "audio_seconds_encoded_per_sec", "value": best["value"], "unit": "x_realtime",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(),
        "sample": sample_txt, "note": "stock transformers.MimiModel.encode on the B200 (eager PyTorch, cuDNN/cuBLAS), host tensors in, "
        "codes back to the host every step; value = the allow_tf32-off run (the one that agrees with the CPU reference)",
        "variants": res, "gpu_launches": None,
    }))


# ---- byte-kernel workloads: resampler and codes -> UTF-8 -------------------------------------------------------------------

def run_byte_kernel(args, local_rank):
    """--workload resample / utf8: the two HBM-side kernels either side of the encoder, timed alone on resident buffers with
    an L2 flush between launches; value = audio-seconds per second, roofline vs the measured copy bandwidth."""
    import ctypes as C
    from tokenize_audio_b200 import _lib, utils
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = utils._Engine.get(dev)
    for kv in args.dbg:
        k, v = kv.split("=")
        _lib.check(eng.lib, eng.h, eng.lib.mimi_b200_debug_set(eng.h, int(k), int(v)), "debug_set")
    peaks = load_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    if args.workload == "resample":
        B, secs, sr_in = 64, 20, 16000
        n_in, n_out = secs * sr_in, secs * SR
        x = torch.randn(B, n_in, device=dev) * 0.1
        y = torch.empty(B, 1, n_out, device=dev)
        ln = (C.c_int64 * B)(*([n_in] * B))

        def launch():
            _lib.check(eng.lib, eng.h, eng.lib.mimi_b200_resample(eng.h, x.data_ptr(), n_in, ln, B, sr_in, SR, y.data_ptr(), n_out, st), "resample")
        audio_s, nbytes = B * secs, B * (n_in + n_out) * 4
        kernel = "resample_kernel (first draft, knob 16)" if "16=1" in args.dbg else "resample_poly_kernel<3,2>"
        flop = B * n_out * 2.0 * 80          # M * V = 2 * 40 FMAs per output sample (68 of them on non-zero taps)
        desc = f"16 kHz -> 24 kHz, batch {B} x {secs} s resident"
    else:
        B, K, T = 4096, 8, 375              # 256 of C5's 16-item batches at once: at C5's own size the launch is 12 us of latency
        codes = torch.randint(0, 2048, (B, K, T), device=dev)
        out = torch.empty(B, T * 28, dtype=torch.uint8, device=dev)
        lens = (C.c_int64 * B)()

        def launch():
            _lib.check(eng.lib, eng.h, eng.lib.mimi_b200_codes_to_utf8(eng.h, codes.data_ptr(), B, K, T, None, 0xE000, 2048, out.data_ptr(), T * 28, lens, st), "utf8")
        audio_s, nbytes, kernel = B * T / 12.5, B * K * T * 8 + B * T * 28, "codes_to_utf8_kernel"
        flop = 0.0
        desc = f"codes [{B},8,375] int64 -> UTF-8 (28 B per frame), resident"
    for _ in range(args.warmup):
        launch()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    times = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ms = float(np.mean(times))
    gbs = nbytes / (ms / 1e3) / 1e9
    roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": None, "kernel": kernel, "avg_launch_ms": ms, "peak_source": peaks["src"],
            "note": "achieved = algorithmic bytes (input + output once) / CUDA-event duration; peak = measured copy bandwidth"}
    if flop:
        roof["fp32_tflops"] = flop / (ms / 1e3) / 1e12
        roof["note"] += ("; this filter spends 80 FMAs per output sample for 6.7 bytes (24 FLOP/B against an fp32 ridge of ~11 FLOP/B): "
                         "the FP32 pipe, not HBM, is the binding roof -- fp32_tflops is the achieved FMA rate x2")
    emit(json.dumps({"metric": "audio_seconds_encoded_per_sec", "value": audio_s / (ms / 1e3), "unit": "x_realtime", "n_gpus": 1,
                     "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                     "vs_baseline": None, "dtype": "f32" if args.workload == "resample" else "int64", "data": "synthetic",
                     "config": {"workload": args.workload, "shape": desc, "l2": "256 MB flush write between launches"},
                     "gpu_launches": args.steps, "clocks": clocks, "roofline": roof}))


def run_b200(args, rank, world, local_rank):
    from tokenize_audio_b200 import utils
    from tokenize_audio_b200.encoder import MimiB200Model, MimiEncoder
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sd = synth.synth_state_dict(0)
    model = MimiB200Model(sd, device=dev)
    if args.mode is not None:
        model.set_mode(args.mode)
    for kv in args.dbg:
        k, v = kv.split("=")
        model.debug_set(int(k), int(v))
    if args.streams:
        model.streams = args.streams
    wrapper = MimiEncoder(model, device=str(dev), ragged=not STRICT, num_quantizers=K_CODEBOOKS)
    clips, lengths, batches = make_workload(rank)
    peaks = load_peaks()
    native = SR_IN != SR

    # resident inputs: padded [B,1,N] batches at 24 kHz already in HBM (for c1: the GPU-resampled utterance)
    dev_batches = []
    for b, cl in zip(batches, clips):
        if native:
            x, l = utils.resample_batch(cl, SR_IN, SR, device=dev)
            dev_batches.append((x.clone(), l))
        else:
            n = max(len(c) for c in cl)
            x = torch.zeros((len(cl), 1, n), dtype=torch.float32)
            for i, c in enumerate(cl):
                x[i, 0, : len(c)] = torch.from_numpy(c)
            dev_batches.append((x.to(dev), [len(c) for c in cl]))
    # one workspace sized for the longest batch up front (a shard driver knows its longest bucket too)
    model.reserve_workspace(BATCH, max(x.shape[2] for x, _ in dev_batches), K_CODEBOOKS)
    wrapper.reserve(BATCH, max(x.shape[2] for x, _ in dev_batches))
    audio_s = [sum(l) / SR for _, l in dev_batches]
    computed_s = [(x.shape[0] * x.shape[2] if STRICT else sum(min(x.shape[2], -(-n // 1920) * 1920) for n in l)) / SR for x, l in dev_batches]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def resident_step(i):
        x, l = dev_batches[i % N_BATCHES]
        codes = model.encode(x, num_quantizers=K_CODEBOOKS, valid_lengths=l if (BATCH > 1 and not STRICT) else None).audio_codes
        if STRINGS:
            return utils.codes_to_utf8_device(codes, [-(-n // 1920) for n in l])
        return codes

    fmt = "utf8" if STRINGS else "int64"

    def e2e_single(i):
        cl = clips[i % N_BATCHES]
        if native:
            return wrapper.encode_native_rate_batch(cl, SR_IN)
        if STRINGS:
            return wrapper.encode_to_strings(cl, num_codebooks=K_CODEBOOKS)
        return wrapper.encode_audio_batch(cl)

    # ---- device-timed, inputs resident ---------------------------------------------------------------
    for i in range(args.warmup):
        resident_step(i)
    barrier()
    # encode() runs a batch of >= 8 items as two item ranges on two streams (the kernels of one range fill the SMs the
    # other's last tiles leave idle), so inside the timed region two kernels are always in flight and a launch's
    # event-to-event duration is not that kernel's own: the per-kernel profile (roofline, ms_per_step_by_kernel) comes from a
    # second pass over the same steps with single-stream launches, right after the timed region.
    split_streams = model.streams if BATCH >= model.min_split_batch else 1
    model.profile(split_streams <= 1)
    launches0 = model.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        resident_step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = model.launch_count - launches0
    if split_streams > 1:
        keep = model.streams
        model.streams = 1
        resident_step(0)
        torch.cuda.synchronize(dev)
        model.profile(True)
        for i in range(args.steps):
            resident_step(i)
        torch.cuda.synchronize(dev)
        model.streams = keep
    prof = model.profile_read()
    model.profile(False)
    total_audio = sum(audio_s[i % N_BATCHES] for i in range(args.steps))
    total_computed = sum(computed_s[i % N_BATCHES] for i in range(args.steps))

    # ---- end to end with host buffers --------------------------------------------------------------------
    # (1) the strictly synchronous reference call shape, one batch per call
    for i in range(min(args.warmup, 3)):
        e2e_single(i)
    barrier()
    t0 = time.perf_counter()
    lat = []
    for i in range(args.steps):
        t1 = time.perf_counter()
        e2e_single(i)
        lat.append(time.perf_counter() - t1)
    torch.cuda.synchronize(dev)
    e2e_single_s = time.perf_counter() - t0
    # (2) the pipelined call: two batches in flight (staging of batch i+1 and D2H of batch i-1 under the kernels of batch i)
    if native:
        e2e_s = e2e_single_s          # the native-rate front door is a synchronous call (c1 is the latency case)
    else:
        for _ in wrapper.encode_stream((clips[i % N_BATCHES] for i in range(min(args.warmup, 3))), fmt=fmt, **({"num_codebooks": K_CODEBOOKS} if STRINGS else {})):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for res in wrapper.encode_stream((clips[i % N_BATCHES] for i in range(args.steps)), fmt=fmt, **({"num_codebooks": K_CODEBOOKS} if STRINGS else {})):
            n_out += len(res)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        assert n_out == BATCH * args.steps
    clocks = sampler.stop() if sampler else None
    if native:
        h2d = float(np.mean([sum(len(c) for c in clips[i % N_BATCHES]) * 4 for i in range(args.steps)]))
    else:
        h2d = float(np.mean([dev_batches[i % N_BATCHES][0].numel() * 4 for i in range(args.steps)]))
    frames_b = [-(-dev_batches[i % N_BATCHES][0].shape[2] // 1920) for i in range(args.steps)]
    d2h = float(np.mean([BATCH * (28 * f if STRINGS else K_CODEBOOKS * f * 8) for f in frames_b]))

    red = sharding.reduce_counters({"audio": total_audio, "ms_max": ms, "e2e_s_max": e2e_s, "e2e1_s_max": e2e_single_s,
                                    "launches": launches}, device=dev)
    if rank != 0:
        return
    value = red["audio"] / (red["ms_max"] / 1e3)
    e2e_value = red["audio"] / red["e2e_s_max"]
    e2e1_value = red["audio"] / red["e2e1_s_max"]

    # ---- roofline of the dominant kernel (live CUDA-event durations from the timed region) --------------
    # The profile is per launch KIND (layer); the ncu launch list is per kernel FUNCTION. In the default generation most
    # layers are launches of one function, the 256-column CTA-pair GEMM: group the kinds by the function that runs them so
    # that "dominant kernel" and its share of the step mean the same thing here and in profiles/*launches*.md.
    # (tc_host.inl: launch_tcp -- 256-column pair tiles for K > 2048, 64 for N = 64, 128 for the rest)
    #  generation 9 also: convs whose taps share input rows run as tap groups, tcp_taps_kernel<BNP, taps per group> (tc_gemm7.cuh),
    #  the front end is front_f16_kernel and the RVQ rvq_f16_kernel unless a debug knob says otherwise)
    fn_of = {"seanet_conv9": 256, "seanet_conv12": 256, "seanet_conv13": 256, "seanet_conv4": 64}
    gemm_kinds = {f"seanet_conv{i}" for i in range(3, 14)} | {"qkv_gemm", "o_proj", "fc1_gelu", "fc2", "downsample_conv", "rvq_input_proj"}
    groups = {}
    cur_mode = args.mode if args.mode is not None else model.DEFAULT_MODE
    default_gen = cur_mode >= 7 and not any(kv.startswith("9=") for kv in args.dbg)
    knob_on = lambda key: any(kv.split("=")[0] == str(key) and kv.split("=")[1] != "0" for kv in args.dbg)
    taps_of = {"seanet_conv3": (128, 2), "seanet_conv6": (128, 2), "downsample_conv": (128, 2), "seanet_conv7": (128, 3),
               "seanet_conv10": (128, 3), "seanet_conv4": (64, 3)} if (cur_mode == 9 and default_gen and not knob_on(20)) else {}
    renamed = {}
    if cur_mode == 9:
        if not knob_on(17):
            renamed["front_fused"] = "front_f16_kernel"
        if not knob_on(19):
            renamed["rvq_fused"] = "rvq_f16_kernel"
    for k, (kms_, kcnt_) in prof.items():
        if k in taps_of:
            g = "tcp_taps_kernel<%d,%d>" % taps_of[k]
        else:
            g = "tcp_gemm_kernel<%d,%d>" % (fn_of.get(k, 128), 3 if cur_mode == 9 else 1) if (default_gen and k in gemm_kinds) else renamed.get(k, k)
        e = groups.setdefault(g, {"ms": 0.0, "count": 0, "kinds": []})
        e["ms"] += kms_; e["count"] += kcnt_; e["kinds"].append(k)
    kind, grp = max(groups.items(), key=lambda kv: kv[1]["ms"])
    kms, kcnt = grp["ms"], grp["count"]
    per_launch_s = kms / 1e3 / kcnt
    audio_per_step = total_computed / args.steps
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath))
    if all(k in MMAC_PER_AUDIO_S and k not in HBM_BOUND_KINDS for k in grp["kinds"]):
        # algorithmic FLOPs of every launch of the kernel in the timed region / their summed duration = FLOPs per launch /
        # average launch duration
        flops = sum(2 * MMAC_PER_AUDIO_S[k] * 1e6 for k in grp["kinds"]) * audio_per_step * args.steps
        achieved = flops / (kms / 1e3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops"]}
    else:
        nbytes = sum(HBM_BYTES_PER_AUDIO_S.get(k, 0.0) for k in grp["kinds"]) * audio_per_step * args.steps
        achieved = nbytes / (kms / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"]}
    # measured DRAM traffic (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per audio-second of the profiled
    # launches, profiles/traffic.json) per average launch of this kernel
    have = [k for k in grp["kinds"] if traffic_tab.get(k, {}).get("dram_bytes_per_audio_s") is not None]
    # (a layer kind that runs once per transformer layer has 8 launches per step, each over the whole batch)
    traffic = (sum(traffic_tab[k]["dram_bytes_per_audio_s"] * prof[k][1] for k in have) * audio_per_step / kcnt) if have else None
    roof.update({"traffic": traffic, "kernel": kind, "share_of_step": kms / sum(v[0] for v in prof.values()),
                 "avg_launch_ms": 1e3 * per_launch_s, "launches_per_step": kcnt / args.steps, "peak_source": peaks["src"]})
    if split_streams > 1:
        roof["profile_pass"] = ("per-kernel durations from a second pass over the same steps with single-stream launches; the "
                                "timed region runs two item ranges on two streams")
    if len(grp["kinds"]) > 1:
        roof["layers"] = sorted(grp["kinds"])
        roof["traffic_covers"] = sorted(have)
    if roof["bound"] == "tensor":
        passes = ("fp16 hi/lo split of both operands (hi*hi + hi*lo + lo*hi, all on kind::f16: 3 bf16-rate tensor passes per MAC)"
                  if cur_mode == 9 else "hi*hi + hi*lo on TF32, lo*hi on bf16 (5 bf16-rate pass units per MAC)")
        roof["precision"] = (f"fp32-equivalent split precision on tcgen05: {passes}; achieved counts algorithmic FLOPs once; "
                             "peak is the measured dense bf16 figure")
    else:
        roof["note"] = "achieved = algorithmic bytes of the launch / CUDA-event duration; peak = measured copy bandwidth"
    # every kernel function of the step with its own roofline reading (the headline `roofline` is the entry with the largest share);
    # the level-1 layers of the 128-column GEMM are HBM-bound, so they are also given against the copy bandwidth
    by_kernel = []
    step_ms = sum(v[0] for v in prof.values())
    for g, e in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
        if e["ms"] / step_ms < 0.01:
            continue
        row = {"kernel": g, "share_of_step": round(e["ms"] / step_ms, 4), "ms_per_step": round(e["ms"] / args.steps, 4)}
        if all(k in MMAC_PER_AUDIO_S and k not in HBM_BOUND_KINDS for k in e["kinds"]):
            tf = sum(2 * MMAC_PER_AUDIO_S[k] * 1e6 for k in e["kinds"]) * audio_per_step * args.steps / (e["ms"] / 1e3) / 1e12
            row.update({"bound": "tensor", "achieved_tflops": round(tf, 1), "frac": round(tf / peaks["tflops"], 4)})
        elif all(k in HBM_BYTES_PER_AUDIO_S for k in e["kinds"]):
            gb = sum(HBM_BYTES_PER_AUDIO_S[k] for k in e["kinds"]) * audio_per_step * args.steps / (e["ms"] / 1e3) / 1e9
            row.update({"bound": "hbm", "achieved_gbs": round(gb, 1), "frac": round(gb / peaks["hbm_gbs"], 4)})
        by_kernel.append(row)
    if cur_mode == 9:
        for k, nbytes in LEVEL1_HBM_BYTES_PER_AUDIO_S.items():
            if k in prof:
                gb = nbytes * audio_per_step * args.steps / (prof[k][0] / 1e3) / 1e9
                by_kernel.append({"kernel": f"GEMM layer {k} alone (HBM view)", "bound": "hbm", "achieved_gbs": round(gb, 1),
                                  "frac": round(gb / peaks["hbm_gbs"], 4), "ms_per_step": round(prof[k][0] / args.steps, 4)})
    breakdown = {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference implementation on host cores -----------
    cpu = None
    agreement = None
    extras = {}
    if world == 1 and not args.no_cpu_baseline:
        try:
            enc = reference_encoder(sd)
            mid = N_BATCHES // 2
            src = clips[mid][:CPU_BASELINE_ITEMS]
            sub = 8
            if native:
                # c1: both sides must see the SAME 24 kHz samples -- take them from the GPU resampler
                sample = [dev_batches[j][0][0, 0, : dev_batches[j][1][0]].cpu().numpy() for j in range(min(4, N_BATCHES))]
                sub = 1
            else:
                sample = src
            t0 = time.perf_counter()
            ref_codes = []
            for j in range(0, len(sample), sub):
                ref_codes += enc(sample[j:j + sub])
            dt = time.perf_counter() - t0
            cores = os.cpu_count() or 1
            cpu = {"value": sum(len(c) for c in sample) / SR / dt, "unit": "x_realtime", "cores": cores, "kind": "reference",
                   "sample": f"{len(sample)} items of batch {mid} ({sum(len(c) for c in sample) / SR:.0f} audio-s) in sub-batches of {sub}, transformers.MimiModel fp32 CPU, {cores} torch threads, {dt:.1f} s"}
            got = []
            for j in range(0, len(sample), sub):
                got += wrapper.encode_audio_batch(sample[j:j + sub])
            agreement = agreement_report(got, ref_codes, sample, enc.model, vs="transformers.MimiModel CPU fp32, same items, same sub-batches")
            if STRINGS:
                # C5: the UTF-8 the GPU wrote for its own codes must be what the reference's codes_to_chars writes for them
                from oracle import chars_oracle
                strs = wrapper.encode_to_strings(sample[:8], num_codebooks=K_CODEBOOKS)
                codes8 = wrapper.encode_audio_batch(sample[:8])
                ok = all(s.encode("utf-8") == chars_oracle.codes_to_utf8(c[:K_CODEBOOKS], 2048) for s, c in zip(strs, codes8))
                extras["utf8_bit_exact"] = bool(ok)
                extras["utf8_bytes_per_item"] = len(strs[0].encode("utf-8"))
            if native:
                # what a different (equally band-limited) resampling filter does to the codes: this kernel vs torchaudio's
                # Kaiser-sinc and scipy's resample_poly on the same 16 kHz clips (the reference's soxr_hq is not installed)
                import math
                import torchaudio.functional as AF
                from scipy.signal import resample_poly
                g = math.gcd(SR, SR_IN)
                base = [c[0] for c in clips[: min(8, N_BATCHES)]]
                ours = [wrapper.encode_native_rate_batch([c], SR_IN)[0] for c in base]
                alt = {"torchaudio_sinc_interp_kaiser": [AF.resample(torch.from_numpy(c), SR_IN, SR, resampling_method="sinc_interp_kaiser").numpy() for c in base],
                       "scipy_resample_poly": [resample_poly(c.astype(np.float64), SR // g, SR_IN // g).astype(np.float32) for c in base]}
                ra = {}
                for nm, ys in alt.items():
                    theirs = [wrapper.encode_audio_chunk(y) for y in ys]
                    t = min(min(a.shape[1], b.shape[1]) for a, b in zip(ours, theirs))
                    eq = np.stack([a[:, :t] == b[:, :t] for a, b in zip(ours, theirs)])
                    ra[nm] = {"codes_equal_frac": float(eq.mean()), "per_codebook": [round(float(v), 4) for v in eq.mean(axis=(0, 2))]}
                # the same measure between the two stand-ins themselves: how far apart two "good" resamplers are on this audio
                ta = [wrapper.encode_audio_chunk(y) for y in alt["torchaudio_sinc_interp_kaiser"]]
                sp = [wrapper.encode_audio_chunk(y) for y in alt["scipy_resample_poly"]]
                t = min(min(a.shape[1], b.shape[1]) for a, b in zip(ta, sp))
                eq = np.stack([a[:, :t] == b[:, :t] for a, b in zip(ta, sp)])
                ra["torchaudio_vs_scipy"] = {"codes_equal_frac": float(eq.mean()), "per_codebook": [round(float(v), 4) for v in eq.mean(axis=(0, 2))]}
                extras["resampler_filter_sensitivity"] = {
                    "what": "code agreement of encode(resample_b200(x16k)) with encode(other_resampler(x16k)), 8 x 10 s clips; "
                            "soxr_hq (the reference's filter) is not installed, these two stand in as a yardstick",
                    **ra}
        except Exception as e:
            cpu = {"value": None, "unit": "x_realtime", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {type(e).__name__}: {e}"}

    api = ("MimiEncoder.encode_native_rate_batch(list[np.ndarray @16 kHz], 16000) -> list[np.ndarray]" if native else
           "MimiEncoder.encode_stream(iter of list[np.ndarray], fmt='utf8') -> list[str] per batch (submit/result, 2 batches in flight)" if STRINGS else
           "MimiEncoder.encode_stream(iter of list[np.ndarray]) -> list[np.ndarray] per batch (submit/result, 2 batches in flight)")
    emit(json.dumps({
        "metric": "audio_seconds_encoded_per_sec", "value": value, "unit": "x_realtime", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": red["ms_max"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "launch": (f"encode() runs the batch as {split_streams} item ranges on {split_streams} streams" if split_streams > 1 else "one stream"),
        "audio_s_per_step": total_audio / args.steps,
        "computed_audio_s_per_step": total_computed / args.steps,
        "padding_waste": sharding.padding_waste(lengths, batches),
        "audio_hours_per_sec": value / 3600.0,
        "e2e": {"value": e2e_value, "unit": "x_realtime", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": api},
        "e2e_single_call": {"value": e2e1_value, "unit": "x_realtime", "latency_ms_median": 1e3 * float(np.median(lat)),
                            "api": "one synchronous wrapper call per batch (MimiEncoder.encode_audio_batch / encode_to_strings / encode_native_rate_batch)"},
        "gpu_launches": int(red["launches"]), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "agreement": agreement, **extras, "roofline_by_kernel": by_kernel, "ms_per_step_by_kernel": breakdown,
    }))


def main():
    global ORDER, STRICT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["resample", "utf8"], help="c2 = the metric's config (default)")
    ap.add_argument("--order", default="bucketed", choices=["bucketed", "file"],
                    help="c2 variant of SURVEY.md 8(d): 'file' = batches in pool order without length bucketing (the reference's own batching)")
    ap.add_argument("--strict", action="store_true",
                    help="c2 variant of SURVEY.md 8(d): encode every item over the padded batch length (no valid_lengths), as the reference does")
    ap.add_argument("--mode", type=int, default=None, help="debug: kernel generation 0 / 7 / 9 (see MimiB200Model.set_mode)")
    ap.add_argument("--streams", type=int, default=0, help="debug: item ranges on side streams inside encode()")
    ap.add_argument("--dbg", action="append", default=[], help="debug: KEY=VALUE for mimi_b200_debug_set (A/B knobs)")
    args = ap.parse_args()
    guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in ("resample", "utf8"):
        if rank == 0:
            run_byte_kernel(args, local_rank)
        return
    select_workload(args.workload)
    ORDER, STRICT = args.order, bool(args.strict)
    if args.mode == 7:
        HBM_BYTES_PER_AUDIO_S["front_fused"] = 24000 * 4 + 24000 * 64 * 6      # TF32 hi (fp32) + bf16 lo
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "reference-cuda":
        run_reference(args, rank, world, device=f"cuda:{local_rank}")
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
