#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the REAL reference implementation.

Run in the build container only (needs `transformers` and /root/reference):

    python tests/golden/make_golden.py

* Mimi encode: ``transformers.MimiModel.encode`` (transformers 5.5.0,
  models/mimi/modeling_mimi.py:1522-1611), CPU fp32, loaded with the seeded synthetic state dict of
  ``tokenize_audio_b200.synth.synth_state_dict(0)`` (real kyutai/mimi weights are not available
  offline). Inputs are stored as int16 PCM (x = pcm / 32768 exactly), outputs as codes [B,K,T]
  and the pre-quantisation latent [B,512,T].
* codes -> unicode: ``/root/reference/pretraining-data/converter.py:17-37`` ``codes_to_chars``.
* feature extractor: ``transformers.EncodecFeatureExtractor`` padding behaviour.

The fixtures travel to the GPU box; this script and /root/reference do not need to.
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from tokenize_audio_b200 import synth  # noqa: E402


def pcm16(seed, n):
    x = synth.synth_speech(seed, n)
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def ref_margins(model, latent, K):
    """Top-2 relative distance margin (d2 - d1) / d1 of every argmin the reference takes, from the reference's own
    modules (MimiEuclideanCodebook.quantize, modeling_mimi.py:1197-1202) along its own residual chain
    (MimiResidualVectorQuantizer.encode, :1262-1280). [B, K, T] fp32: lets a test classify a code flip as a near-tie
    without re-running anything."""
    q = model.quantizer
    B, _, T = latent.shape
    margins, codes = [], []
    for rvq, n in ((q.semantic_residual_vector_quantizer, 1), (q.acoustic_residual_vector_quantizer, K - 1)):
        r = rvq.input_proj(latent)
        for layer in rvq.layers[:n]:
            emb = layer.codebook.embed
            x = r.permute(0, 2, 1).reshape(-1, emb.shape[1])
            d = torch.cdist(x[None], emb[None], p=2)[0]
            top2 = d.topk(2, largest=False).values
            idx = d.argmin(-1)
            margins.append(((top2[:, 1] - top2[:, 0]) / top2[:, 0].clamp_min(1e-30)).view(B, T))
            codes.append(idx.view(B, T))
            r = r - torch.nn.functional.embedding(idx, emb).view(B, T, -1).permute(0, 2, 1)
    return torch.stack(margins, 1).numpy().astype(np.float32), torch.stack(codes, 1).numpy()


def main():
    from transformers import EncodecFeatureExtractor, MimiConfig, MimiModel

    only = set(sys.argv[1:])          # fixture names to (re)generate; none = all
    torch.manual_seed(0)
    fe = EncodecFeatureExtractor()
    models = {}

    def load(variant):
        if variant not in models:
            sd = synth.synth_state_dict(0) if variant == "seed0" else synth.variant_state_dict(variant)
            m = MimiModel(MimiConfig()).eval()
            missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
            assert not unexpected
            assert all(k.startswith(("decoder", "upsample")) or "output_proj" in k for k in missing), missing
            models[variant] = (m, synth.state_dict_digest(sd))
        return models[variant]

    model, digest = load("seed0")

    def run(name, pcms, K, variant="seed0"):
        if only and name not in only:
            return
        model, digest = load(variant)
        audio = [p.astype(np.float32) / np.float32(32768.0) for p in pcms]
        if len(audio) > 1:      # REF/emilia-mimi/process_shard.py:113-118
            inputs = fe(raw_audio=audio, sampling_rate=24000, return_tensors="pt", padding=True)
        else:                   # REF/emilia-mimi/process_shard.py:75-79 (padding=None -> True)
            inputs = fe(raw_audio=audio[0], sampling_rate=24000, return_tensors="pt")
        iv, pm = inputs["input_values"], inputs["padding_mask"]
        with torch.no_grad():
            out = model.encode(iv, pm, num_quantizers=K)
            emb = model.encoder(iv)
            z = model.encoder_transformer(emb.transpose(1, 2))[0].transpose(1, 2)
            latent = model.downsample(z)
            # the same call without a mask / with None must give the same codes (mask is unused)
            out2 = model.encode(iv, num_quantizers=K)
        assert torch.equal(out.audio_codes, out2.audio_codes)
        codes = out.audio_codes.numpy()
        assert codes.dtype == np.int64 and codes.max() < 2048
        extra = {}
        if variant != "seed0" or name.startswith(("mimi_c3", "mimi_c4")):
            with torch.no_grad():
                mg, codes_again = ref_margins(model, latent, K)
            assert (codes_again == codes).mean() >= 0.9999, (codes_again == codes).mean()
            extra = {"margins": mg, "weights_variant": variant}
        n_max = max(len(p) for p in pcms)
        pcm = np.zeros((len(pcms), n_max), np.int16)
        for i, p in enumerate(pcms):
            pcm[i, :len(p)] = p
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(
            path, pcm=pcm, lengths=np.array([len(p) for p in pcms], np.int64),
            input_values_sha256=hashlib.sha256(iv.numpy().tobytes()).hexdigest(),
            padding_mask_sum=pm.sum(-1).numpy().astype(np.int64),
            codes=codes.astype(np.int16), latent=latent.numpy().astype(np.float32),
            seanet_out_first=emb.numpy()[0, :, :8].astype(np.float32),
            num_quantizers=np.int64(K), weights_digest=digest, weights_seed=np.int64(0), **extra)
        print(name, "codes", codes.shape, "latent", tuple(latent.shape), os.path.getsize(path) // 1024, "KiB")

    # C1-like: B=1, ragged tail (N % 1920 != 0), all 32 codebooks
    run("mimi_b1_k32", [pcm16(101, 24000 + 18000 + 777)], 32)
    # C2-like: padded batch of 3 with a padding mask, 8 codebooks (odd T25 for the short items)
    run("mimi_b3_pad_k8", [pcm16(201, 30000), pcm16(202, 47999), pcm16(203, 12345)], 8)
    # C3-like: T25 = 275 > 250 exercises the sliding window, 8 codebooks
    run("mimi_long_k8", [pcm16(301, 275 * 960)], 8)

    # BASELINE config 3 at full length: two ragged 30 s / 22 s items (T25 = 750 = three attention windows), 8 codebooks
    run("mimi_c3_k8", [pcm16(501, 720000), pcm16(502, 22 * 24000 + 777)], 8)
    # BASELINE config 4 at full length: 2 x 15 s (360000 samples: the last frame is partial), all 32 codebooks
    run("mimi_c4_k32", [pcm16(511, 360000), pcm16(512, 360000)], 32)
    # exact ties: duplicated and all-zero codebook rows that win (lowest-index rule, TF:1200-1201)
    run("mimi_ties_k32", [pcm16(521, 6 * 24000 + 100), pcm16(522, 4 * 24000)], 32, variant="ties")
    # second weight draw with heavy tails and x50 outlier channels (T25 = 302 > window on the long item)
    run("mimi_heavy_k32", [pcm16(531, 290000), pcm16(532, 170000)], 32, variant="heavy")

    if not only or "audio_to_str" in only:
        # utils.audio_to_str (REF/emilia-mimi/utils.py:58-69; the module itself needs librosa, so its three steps are
        # run here): un-batched encode of a [1,1,N] tensor, audio_codes[0][0][:8], converter.codes_to_chars
        spec = importlib.util.spec_from_file_location("ref_converter", "/root/reference/pretraining-data/converter.py")
        conv = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(conv)
        p = pcm16(541, 3 * 24000 + 1234)
        a = p.astype(np.float32) / np.float32(32768.0)
        with torch.no_grad():
            audio_codes = model.encode(torch.tensor(a).unsqueeze(0).unsqueeze(1))
        s = conv.codes_to_chars(audio_codes[0][0][:8, :].numpy(), 2048, copy_before_conversion=True, unicode_offset=0xE000)
        np.savez_compressed(os.path.join(HERE, "audio_to_str.npz"), pcm=p, utf8=np.frombuffer(s.encode("utf-8"), np.uint8),
                            weights_digest=digest)
        print("audio_to_str", len(s), "chars")

    if not only or "mimi_decode" in only:
        # decode direction (MimiModel.decode, modeling_mimi.py:1613-1679) on the seed-0 encode weights + seed-0 decode weights:
        # random codes, K = 8 (what str_to_audio passes) and K = 32, T25 = 2T beyond the attention window for the first
        sd_full = {**synth.synth_state_dict(0), **synth.decoder_state_dict(0)}
        mfull = MimiModel(MimiConfig()).eval()
        missing, unexpected = mfull.load_state_dict({k: torch.from_numpy(v) for k, v in sd_full.items()}, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        g = np.random.Generator(np.random.PCG64(11))
        cases = {"weights_digest": synth.state_dict_digest(sd_full)}
        for tag, B, K, T in (("k8", 2, 8, 140), ("k32", 1, 32, 9), ("k1", 1, 1, 5)):
            codes = g.integers(0, 2048, size=(B, K, T), dtype=np.int64)
            with torch.no_grad():
                audio = mfull.decode(torch.from_numpy(codes)).audio_values.numpy()
            assert audio.shape == (B, 1, 1920 * T)
            cases[f"{tag}_codes"] = codes.astype(np.int16)
            cases[f"{tag}_audio"] = audio.astype(np.float32)
        # str_to_audio (REF/emilia-mimi/utils.py:72-81): string -> chars_to_codes -> decode(codes[None]).audio_values[0]
        spec = importlib.util.spec_from_file_location("ref_converter", "/root/reference/pretraining-data/converter.py")
        conv = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(conv)
        s8 = conv.codes_to_chars(cases["k8_codes"][0].astype(np.int64), 2048, copy_before_conversion=True, unicode_offset=0xE000)
        cases["k8_item0_utf8"] = np.frombuffer(s8.encode("utf-8"), np.uint8)
        np.savez_compressed(os.path.join(HERE, "mimi_decode.npz"), **cases)
        print("mimi_decode", {k: getattr(v, "shape", v) for k, v in cases.items()})

    if only and not ({"encoded_length", "codes_to_chars"} & only):
        return
    # get_encoded_length known answers (TF:1490-1503)
    lens = np.array([240777, 150000, 1, 1919, 1920, 1921, 47999, 12345, 264000], np.int64)
    enc = model.get_encoded_length(torch.from_numpy(lens)).numpy()
    np.savez_compressed(os.path.join(HERE, "encoded_length.npz"), lengths=lens, frames=enc.astype(np.int64))
    print("encoded_length", dict(zip(lens.tolist(), enc.tolist())))

    # codes -> unicode from the reference's own converter (imports numpy+torch only)
    spec = importlib.util.spec_from_file_location("ref_converter", "/root/reference/pretraining-data/converter.py")
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    g = np.random.Generator(np.random.PCG64(7))
    cases = {}
    for tag, K, T in (("k8_t375", 8, 375), ("k8_t3", 8, 3), ("k32_t17", 32, 17), ("k1_t5", 1, 5), ("k8_t0", 8, 0)):
        codes = g.integers(0, 2048, size=(K, T), dtype=np.int64)
        if T:
            codes[:, 0] = 0
            codes[:, -1] = 2047
        s = conv.codes_to_chars(codes, 2048, copy_before_conversion=True, unicode_offset=0xE000)
        back = np.array(conv.chars_to_codes(s, K, 2048, unicode_offset=0xE000), np.int64).reshape(K, T) if T else codes
        assert np.array_equal(back, codes)
        cases[f"{tag}_codes"] = codes.astype(np.int16)
        cases[f"{tag}_utf8"] = np.frombuffer(s.encode("utf-8"), np.uint8)
    np.savez_compressed(os.path.join(HERE, "codes_to_chars.npz"), **cases)
    print("codes_to_chars", {k: v.shape for k, v in cases.items()})


if __name__ == "__main__":
    main()
