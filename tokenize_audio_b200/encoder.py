"""Host-side mirror of the reference interface for the Mimi encode path.

* :class:`MimiB200Model` keeps the call signature of ``transformers.MimiModel.encode``
  (transformers/models/mimi/modeling_mimi.py:1522-1531): ``encode(input_values[B,1,N], padding_mask=None,
  num_quantizers=None, ...) -> MimiEncoderOutput`` with ``audio_codes`` int64 ``[B,K,T]``, T = ceil(N/1920).
* :class:`MimiEncoder` mirrors the wrapper class every ``*-mimi/process_*.py`` script copies
  (REF/emilia-mimi/process_shard.py:50-140): ``encode_audio_chunk`` / ``encode_audio_batch``.
* :class:`EncodecFeatureExtractorLite` mirrors what those scripts use of ``EncodecFeatureExtractor``
  (transformers/models/encodec/feature_extraction_encodec.py:81-202): fp32 cast, zero right-padding to the
  longest item, ``padding_mask``.

PyTorch is only the plumbing here (device memory, streams); all arithmetic is in libmimi_b200.so, and
nothing in this module falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .synth import SEANET_CONVS

FRAME_SIZE = 1920
NUM_QUANTIZERS = 32
NUM_SEMANTIC_QUANTIZERS = 1


class MimiDecoderOutput(tuple):
    """Tuple-compatible stand-in for transformers' ``MimiDecoderOutput`` (modeling_mimi.py:196-211):
    ``out.audio_values``, ``out[0]``, ``out.decoder_past_key_values``."""

    def __new__(cls, audio_values, decoder_past_key_values=None):
        return super().__new__(cls, (audio_values, decoder_past_key_values))

    audio_values = property(lambda self: self[0])
    decoder_past_key_values = property(lambda self: self[1])


class MimiEncoderOutput(tuple):
    """Tuple-compatible stand-in for transformers' ``MimiEncoderOutput`` (modeling_mimi.py:60-75):
    ``out.audio_codes``, ``out[0]``, ``out.encoder_past_key_values``, ``out.padding_cache``."""

    def __new__(cls, audio_codes, encoder_past_key_values=None, padding_cache=None):
        return super().__new__(cls, (audio_codes, encoder_past_key_values, padding_cache))

    audio_codes = property(lambda self: self[0])
    encoder_past_key_values = property(lambda self: self[1])
    padding_cache = property(lambda self: self[2])


def _as_f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _rope_inv_freq() -> np.ndarray:
    # MimiRotaryEmbedding.compute_default_rope_parameters (modeling_mimi.py:538-560), same torch ops
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.int64).to(dtype=torch.float) / 64))
    return np.ascontiguousarray(inv.numpy().astype(np.float32))


class MimiB200Model:
    """B200-native drop-in for the encode half of ``transformers.MimiModel``."""

    def __init__(self, state_dict: Dict[str, Union[np.ndarray, torch.Tensor]], device: Union[str, int, torch.device] = "cuda"):
        if not torch.cuda.is_available():
            raise _lib.MimiB200Error("MimiB200Model needs a CUDA device (B200); there is no CPU fallback")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.MimiB200Error(f"MimiB200Model only runs on CUDA devices, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._lib = _lib.load_library()
        self._lock = threading.Lock()
        self._workspaces: Dict[object, Optional[torch.Tensor]] = {}
        self._mode = self.DEFAULT_MODE
        # encode() runs a batch of >= min_split_batch items as `streams` contiguous item ranges on side streams: every
        # kernel is a persistent one-CTA-per-SM grid whose last tiles leave SMs idle, the other range's kernels fill them
        # (+4-5 % on C2 / C3 / C4; three ranges are slower than one)
        self.streams = 2
        self.balance_ranges = True        # ranges of equal total length rather than equal item count (A/B switch)
        self.min_split_batch = 8
        self._side_streams = []
        self._last_split = False
        self.ragged_from_mask = False     # True: use padding_mask row sums as valid lengths (ragged mode)
        h = C.c_void_p()
        rc = self._lib.mimi_b200_create(C.byref(h), self.device.index)
        _lib.check(self._lib, None, rc, "mimi_b200_create")
        self._h = h
        self._load(state_dict)
        self.has_decoder = self._load_decoder(state_dict)

    # -- construction helpers ---------------------------------------------------------------------
    @classmethod
    def from_transformers(cls, model, device="cuda") -> "MimiB200Model":
        """Build from a loaded ``transformers.MimiModel`` (e.g. ``MimiModel.from_pretrained("kyutai/mimi")``)."""
        return cls(model.state_dict(), device=device)

    @classmethod
    def from_pretrained(cls, path: str, device="cuda") -> "MimiB200Model":
        """Load a local ``kyutai/mimi`` checkpoint: a ``.safetensors`` file, or a directory holding
        ``model.safetensors`` (the layout ``MimiModel.from_pretrained`` reads, REF/emilia-mimi/process_shard.py:58)."""
        import os
        from safetensors.numpy import load_file
        if os.path.isdir(path):
            path = os.path.join(path, "model.safetensors")
        return cls(load_file(path), device=device)

    def _load(self, sd) -> None:
        keep: List[np.ndarray] = []

        def ptr(name: str, shape) -> C.c_void_p:
            if name not in sd:
                raise KeyError(f"state dict is missing '{name}'")
            a = _as_f32(sd[name])
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"'{name}' has shape {tuple(a.shape)}, expected {tuple(shape)}")
            keep.append(a)
            return C.c_void_p(a.ctypes.data)

        w = _lib.Weights()
        for i, (name, cin, cout, k, _s) in enumerate(SEANET_CONVS):
            w.conv_weight[i] = ptr(f"{name}.conv.weight", (cout, cin, k))
            w.conv_bias[i] = ptr(f"{name}.conv.bias", (cout,))
        for l in range(8):
            p = f"encoder_transformer.layers.{l}"
            lw = w.layer[l]
            lw.input_layernorm_weight = ptr(f"{p}.input_layernorm.weight", (512,))
            lw.input_layernorm_bias = ptr(f"{p}.input_layernorm.bias", (512,))
            lw.q_proj_weight = ptr(f"{p}.self_attn.q_proj.weight", (512, 512))
            lw.k_proj_weight = ptr(f"{p}.self_attn.k_proj.weight", (512, 512))
            lw.v_proj_weight = ptr(f"{p}.self_attn.v_proj.weight", (512, 512))
            lw.o_proj_weight = ptr(f"{p}.self_attn.o_proj.weight", (512, 512))
            lw.self_attn_layer_scale = ptr(f"{p}.self_attn_layer_scale.scale", (512,))
            lw.post_attention_layernorm_weight = ptr(f"{p}.post_attention_layernorm.weight", (512,))
            lw.post_attention_layernorm_bias = ptr(f"{p}.post_attention_layernorm.bias", (512,))
            lw.fc1_weight = ptr(f"{p}.mlp.fc1.weight", (2048, 512))
            lw.fc2_weight = ptr(f"{p}.mlp.fc2.weight", (512, 2048))
            lw.mlp_layer_scale = ptr(f"{p}.mlp_layer_scale.scale", (512,))
        w.downsample_weight = ptr("downsample.conv.weight", (512, 512, 4))
        w.semantic_input_proj_weight = ptr("quantizer.semantic_residual_vector_quantizer.input_proj.weight", (256, 512, 1))
        w.acoustic_input_proj_weight = ptr("quantizer.acoustic_residual_vector_quantizer.input_proj.weight", (256, 512, 1))
        for s in range(32):
            which, idx = ("semantic", 0) if s == 0 else ("acoustic", s - 1)
            q = f"quantizer.{which}_residual_vector_quantizer.layers.{idx}.codebook"
            w.embed_sum[s] = ptr(f"{q}.embed_sum", (2048, 256))
            w.cluster_usage[s] = ptr(f"{q}.cluster_usage", (2048,))
        inv = _rope_inv_freq()
        keep.append(inv)
        w.rope_inv_freq = C.c_void_p(inv.ctypes.data)
        with torch.cuda.device(self.device):
            rc = self._lib.mimi_b200_load_weights(self._h, C.byref(w))
        _lib.check(self._lib, self._h, rc, "mimi_b200_load_weights")
        del keep

    def _load_decoder(self, sd) -> bool:
        """Decode-side tensors (output projections, upsample, decoder transformer, SEANet decoder) when the state dict has
        them -- a full ``kyutai/mimi`` checkpoint does; an encode-only dict leaves :meth:`decode` unavailable."""
        from .synth import SEANET_DECODER_UPS
        if "decoder.layers.0.conv.weight" not in sd:
            return False
        keep: List[np.ndarray] = []

        def ptr(name: str, shape) -> C.c_void_p:
            if name not in sd:
                raise KeyError(f"state dict is missing '{name}'")
            a = _as_f32(sd[name])
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"'{name}' has shape {tuple(a.shape)}, expected {tuple(shape)}")
            keep.append(a)
            return C.c_void_p(a.ctypes.data)

        w = _lib.DecoderWeights()
        w.semantic_output_proj_weight = ptr("quantizer.semantic_residual_vector_quantizer.output_proj.weight", (512, 256, 1))
        w.acoustic_output_proj_weight = ptr("quantizer.acoustic_residual_vector_quantizer.output_proj.weight", (512, 256, 1))
        w.upsample_weight = ptr("upsample.conv.weight", (512, 1, 4))
        for l in range(8):
            p = f"decoder_transformer.layers.{l}"
            lw = w.layer[l]
            lw.input_layernorm_weight = ptr(f"{p}.input_layernorm.weight", (512,))
            lw.input_layernorm_bias = ptr(f"{p}.input_layernorm.bias", (512,))
            lw.q_proj_weight = ptr(f"{p}.self_attn.q_proj.weight", (512, 512))
            lw.k_proj_weight = ptr(f"{p}.self_attn.k_proj.weight", (512, 512))
            lw.v_proj_weight = ptr(f"{p}.self_attn.v_proj.weight", (512, 512))
            lw.o_proj_weight = ptr(f"{p}.self_attn.o_proj.weight", (512, 512))
            lw.self_attn_layer_scale = ptr(f"{p}.self_attn_layer_scale.scale", (512,))
            lw.post_attention_layernorm_weight = ptr(f"{p}.post_attention_layernorm.weight", (512,))
            lw.post_attention_layernorm_bias = ptr(f"{p}.post_attention_layernorm.bias", (512,))
            lw.fc1_weight = ptr(f"{p}.mlp.fc1.weight", (2048, 512))
            lw.fc2_weight = ptr(f"{p}.mlp.fc2.weight", (512, 2048))
            lw.mlp_layer_scale = ptr(f"{p}.mlp_layer_scale.scale", (512,))
        w.conv_in_weight = ptr("decoder.layers.0.conv.weight", (1024, 512, 7))
        w.conv_in_bias = ptr("decoder.layers.0.conv.bias", (1024,))
        for i, (name, cin, cout, r, res) in enumerate(SEANET_DECODER_UPS):
            w.up_weight[i] = ptr(f"{name}.conv.weight", (cin, cout, 2 * r))
            w.up_bias[i] = ptr(f"{name}.conv.bias", (cout,))
            w.res_a_weight[i] = ptr(f"{res}.block.1.conv.weight", (cout // 2, cout, 3))
            w.res_a_bias[i] = ptr(f"{res}.block.1.conv.bias", (cout // 2,))
            w.res_b_weight[i] = ptr(f"{res}.block.3.conv.weight", (cout, cout // 2, 1))
            w.res_b_bias[i] = ptr(f"{res}.block.3.conv.bias", (cout,))
        w.conv_out_weight = ptr("decoder.layers.14.conv.weight", (1, 64, 3))
        w.conv_out_bias = ptr("decoder.layers.14.conv.bias", (1,))
        with torch.cuda.device(self.device):
            rc = self._lib.mimi_b200_load_decoder_weights(self._h, C.byref(w))
        _lib.check(self._lib, self._h, rc, "mimi_b200_load_decoder_weights")
        del keep
        return True

    def decode(self, audio_codes: torch.Tensor, padding_mask: Optional[torch.Tensor] = None, decoder_past_key_values=None,
               return_dict: Optional[bool] = None):
        """Same contract as ``MimiModel.decode`` (modeling_mimi.py:1613-1679): ``audio_codes`` ``[B, K, T]`` (K = 1..32) ->
        ``MimiDecoderOutput`` with ``audio_values`` ``[B, 1, 1920*T]`` fp32, truncated to ``padding_mask.shape[-1]`` when a
        shorter mask is given. Runs on the exact-fp32 FFMA kernels (the reference's round-trip spot-check path,
        REF/emilia-mimi/utils.py:72-81)."""
        if decoder_past_key_values is not None:
            raise NotImplementedError("decoder_past_key_values (streaming decode) is not supported")
        if not self.has_decoder:
            raise _lib.MimiB200Error("this model was built from an encode-only state dict: decode() needs the decoder tensors")
        if audio_codes.dim() != 3:
            raise ValueError(f"audio_codes must be [batch, num_quantizers, codes_length], got {tuple(audio_codes.shape)}")
        if not audio_codes.is_cuda or audio_codes.device != self.device:
            raise _lib.MimiB200Error(f"audio_codes must live on {self.device} (no CPU fallback)")
        B, K, T = audio_codes.shape
        if K < 1 or K > NUM_QUANTIZERS:
            raise ValueError(f"audio_codes must hold between 1 and {NUM_QUANTIZERS} codebooks, got {K}")
        codes = audio_codes.detach().to(torch.int64).contiguous()
        if codes.numel():
            lo, hi = int(codes.min()), int(codes.max())
            if lo < 0 or hi >= 2048:      # F.embedding in the reference raises IndexError here
                raise IndexError(f"audio_codes must lie in [0, 2048), got values in [{lo}, {hi}]")
        audio = torch.empty((B, 1, FRAME_SIZE * T), dtype=torch.float32, device=self.device)
        if B > 0 and T > 0:
            with self._lock, torch.cuda.device(self.device):
                nbytes = C.c_size_t()
                rc = self._lib.mimi_b200_decode_workspace_bytes(self._h, B, T, C.byref(nbytes))
                _lib.check(self._lib, self._h, rc, "mimi_b200_decode_workspace_bytes")
                ws = self._ws(nbytes.value, "decode")
                rc = self._lib.mimi_b200_decode(self._h, codes.data_ptr(), B, K, T, audio.data_ptr(), ws.data_ptr(), ws.numel(),
                                                torch.cuda.current_stream(self.device).cuda_stream)
                _lib.check(self._lib, self._h, rc, "mimi_b200_decode")
        if padding_mask is not None and padding_mask.shape[-1] < audio.shape[-1]:
            audio = audio[..., : padding_mask.shape[-1]]
        out = MimiDecoderOutput(audio, None)
        return tuple(out) if return_dict is False else out

    # -- nn.Module-flavoured no-ops so reference scripts keep working -----------------------------------
    def to(self, device):
        if torch.device(device).type != "cuda":
            raise _lib.MimiB200Error("MimiB200Model cannot move to a non-CUDA device (no CPU fallback)")
        return self

    def eval(self):
        return self

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mimi_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the hot path --------------------------------------------------------------------------------
    def get_encoded_length(self, input_length):
        """MimiModel.get_encoded_length (modeling_mimi.py:1490-1503)."""
        if isinstance(input_length, torch.Tensor):
            return torch.div(input_length + (FRAME_SIZE - 1), FRAME_SIZE, rounding_mode="floor")
        return int(self._lib.mimi_b200_encoded_frames(int(input_length)))

    def _ws(self, nbytes: int, slot: int = 0) -> torch.Tensor:
        """Workspace `slot` (one per concurrent encode: slot 0 for plain calls, more when a wrapper runs item ranges on
        several streams), grown on demand."""
        w = self._workspaces.get(slot)
        if w is None or w.numel() < nbytes:
            self._workspaces[slot] = None
            w = self._workspaces[slot] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return w

    def reserve_workspace(self, batch: int, num_samples: int, num_quantizers: int = NUM_QUANTIZERS, slot: int = 0) -> int:
        """Pre-size the activation workspace for the largest ``[batch, 1, num_samples]`` input that will be
        encoded, so no later call has to grow it (a multi-GB device allocation in the middle of a shard).
        Returns the size in bytes."""
        with self._lock, torch.cuda.device(self.device):
            nbytes = C.c_size_t()
            rc = self._lib.mimi_b200_workspace_bytes(self._h, int(batch), int(num_samples), int(num_quantizers), C.byref(nbytes))
            _lib.check(self._lib, self._h, rc, "mimi_b200_workspace_bytes")
            self._ws(nbytes.value + (1 << 20), slot)       # + rounding slack for the per-stream split of encode()
        return int(nbytes.value)

    def encode(self, input_values: torch.Tensor, padding_mask: Optional[torch.Tensor] = None,
               num_quantizers: Optional[float] = None, encoder_past_key_values=None, padding_cache=None,
               use_streaming: Optional[bool] = None, return_dict: Optional[bool] = None,
               valid_lengths: Optional[Sequence[int]] = None, return_latent: bool = False,
               staged_groups=None, workspace_slot: int = 0, on_front_done=None):
        """Same contract as ``MimiModel.encode``. Extras (keyword-only in spirit): ``valid_lengths`` switches
        on ragged mode (skip work past each item's last kept frame), ``return_latent`` also returns the
        pre-quantisation latent ``[B,512,T]`` for parity checks. ``staged_groups`` = iterable of ``(b0, b1, land)``
        partitioning ``[0, B)`` in order: ``input_values`` is a device buffer still being filled, and ``land()`` is
        called right before the front end of items ``[b0, b1)`` is launched -- it stages those items (fill, H2D, make the
        current stream wait for the copy). The GPU then works on the first groups while the host stages the later
        ones (``mimi_b200_encode_phase``); the result is that of a plain call on the complete buffer. ``on_front_done()``
        is called once the last reader of ``input_values`` (the front end) has been queued."""
        if encoder_past_key_values is not None or padding_cache is not None or use_streaming:
            raise NotImplementedError("streaming / cache arguments are not supported (the reference scripts never pass them)")
        K = NUM_QUANTIZERS if num_quantizers is None else num_quantizers
        if K > NUM_QUANTIZERS:   # modeling_mimi.py:1562-1565
            raise ValueError(
                f"The number of quantizers (i.e codebooks) asked should be lower than the total number of quantizers {NUM_QUANTIZERS}, but is currently {K}.")
        if input_values.dim() != 3:
            raise ValueError(f"input_values must be [batch, channels, length], got {tuple(input_values.shape)}")
        B, channels, N = input_values.shape
        if channels < 1 or channels > 2:   # modeling_mimi.py:1569-1570
            raise ValueError(f"Number of audio channels must be 1 or 2, but got {channels}")
        if channels != 1:
            # MimiConfig.audio_channels == 1: the first conv takes one channel (torch would raise here too)
            raise RuntimeError("expected input_values with 1 audio channel (kyutai/mimi is mono)")
        if K < NUM_SEMANTIC_QUANTIZERS:    # modeling_mimi.py:1324-1327
            raise ValueError(
                f"The number of quantizers (i.e codebooks) asked should be higher than the number of semantic quantizers {NUM_SEMANTIC_QUANTIZERS}, but is currently {K}.")
        K = int(K)
        if not input_values.is_cuda:
            raise _lib.MimiB200Error("input_values must live on the CUDA device (no CPU fallback)")
        if input_values.device != self.device:
            raise _lib.MimiB200Error(f"input_values is on {input_values.device}, model on {self.device}")
        x = input_values.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        T = -(-N // FRAME_SIZE)
        if valid_lengths is None and self.ragged_from_mask and padding_mask is not None:
            valid_lengths = padding_mask.reshape(B, -1, N)[:, 0].sum(-1).tolist()
        vl = None
        if valid_lengths is not None:
            vl = (C.c_int64 * B)(*[int(v) for v in valid_lengths])
        codes = torch.empty((B, K, T), dtype=torch.int64, device=self.device)
        latent = torch.empty((B, 512, T), dtype=torch.float32, device=self.device) if return_latent else None
        self._last_B = B
        self._last_split = False
        if B > 0 and N > 0:
            with self._lock, torch.cuda.device(self.device):
                nbytes = C.c_size_t()
                rc = self._lib.mimi_b200_workspace_bytes(self._h, B, N, K, C.byref(nbytes))
                _lib.check(self._lib, self._h, rc, "mimi_b200_workspace_bytes")
                ws = self._ws(nbytes.value, workspace_slot)
                stream = torch.cuda.current_stream(self.device).cuda_stream
                lat_ptr = latent.data_ptr() if latent is not None else None
                if staged_groups is None and self.streams > 1 and B >= self.min_split_batch and workspace_slot == 0:
                    self._encode_multi_stream(x, B, N, vl, K, codes, latent)
                    self._last_split = True
                elif staged_groups is None:
                    rc = self._lib.mimi_b200_encode(self._h, x.data_ptr(), B, N, vl, K, codes.data_ptr(), lat_ptr,
                                                    ws.data_ptr(), ws.numel(), stream)
                    _lib.check(self._lib, self._h, rc, "mimi_b200_encode")
                else:
                    def phase(ph, b0, b1):
                        rc = self._lib.mimi_b200_encode_phase(self._h, ph, b0, b1, x.data_ptr(), B, N, vl, K, codes.data_ptr(),
                                                              lat_ptr, ws.data_ptr(), ws.numel(), stream)
                        _lib.check(self._lib, self._h, rc, "mimi_b200_encode_phase")
                    phase(_lib.PHASE_BEGIN, 0, 0)
                    nxt = 0
                    for b0, b1, land in staged_groups:
                        if b0 != nxt or b1 <= b0 or b1 > B:
                            raise ValueError("staged_groups must partition [0, B) in order")
                        land()
                        phase(_lib.PHASE_FRONT, b0, b1)
                        nxt = b1
                    if nxt != B:
                        raise ValueError("staged_groups must partition [0, B) in order")
                    if on_front_done is not None:
                        on_front_done()         # every reader of input_values has been queued: the caller may record an event
                    phase(_lib.PHASE_FINISH, 0, 0)
        out = MimiEncoderOutput(codes, None, None)
        if return_latent:
            return out, latent
        if return_dict is False:
            return tuple(out)
        return out

    @staticmethod
    def _range_bounds(B: int, n: int, vl) -> List[int]:
        """Item ranges of (nearly) equal WORK: with ``valid_lengths`` the split points sit where the cumulative length crosses
        j / n of the total (a length-bucketed batch that straddles two buckets has its short items first: equal item counts gave
        the two streams 45 % / 55 % of the samples), without them equal item counts."""
        if vl is None or n <= 1:
            return [B * j // n for j in range(n + 1)]
        cum, acc = [], 0
        for v in vl:
            acc += max(int(v), 1)
            cum.append(acc)
        bounds = [0]
        for j in range(1, n):
            target = acc * j / n
            k = min(range(B), key=lambda i: abs(cum[i] - target)) + 1
            bounds.append(min(max(k, bounds[-1] + 1), B - (n - j)))
        bounds.append(B)
        return bounds

    def _encode_multi_stream(self, x, B, N, vl, K, codes, latent) -> None:
        """The batch as ``self.streams`` contiguous item ranges, each encoded on its own stream with its own workspace.
        Items are independent, so the results are those of one call; the point is that every kernel is a persistent grid
        with one CTA per SM whose last tiles leave SMs idle (15-30 % of a GEMM launch): the CTAs of the other range's
        kernel fill them."""
        n = self.streams
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if len(self._side_streams) < n:
            self._side_streams = [torch.cuda.Stream(device=dev) for _ in range(n)]
        bounds = self._range_bounds(B, n, vl if self.balance_ranges else None)
        sizes = []
        for j in range(n):
            nb = C.c_size_t()
            rc = self._lib.mimi_b200_workspace_bytes(self._h, bounds[j + 1] - bounds[j], N, K, C.byref(nb))
            _lib.check(self._lib, self._h, rc, "mimi_b200_workspace_bytes")
            sizes.append((int(nb.value) + 1023) // 1024 * 1024)
        ws = self._ws(sum(sizes))
        off = 0
        for j in range(n):
            b0, b1 = bounds[j], bounds[j + 1]
            st = self._side_streams[j]
            st.wait_stream(main)
            sub_vl = None if vl is None else (C.c_int64 * (b1 - b0))(*list(vl)[b0:b1])
            rc = self._lib.mimi_b200_encode(
                self._h, x[b0:].data_ptr(), b1 - b0, N, sub_vl, K, codes[b0:].data_ptr(),
                latent[b0:].data_ptr() if latent is not None else None, ws.data_ptr() + off, sizes[j], st.cuda_stream)
            _lib.check(self._lib, self._h, rc, "mimi_b200_encode")
            off += sizes[j]
        for j in range(n):
            main.wait_stream(self._side_streams[j])

    # -- parity helpers --------------------------------------------------------------------------------
    @property
    def supports_phased(self) -> bool:
        """mimi_b200_encode_phase needs the fused front end (the tensor-core generations 7 and 9)."""
        return self._mode in (7, 9)

    def debug_set(self, key: int, value: int) -> None:
        if key == 3:
            self._mode = int(value)
        _lib.check(self._lib, self._h, self._lib.mimi_b200_debug_set(self._h, key, value), "mimi_b200_debug_set")

    def debug_tap(self, which: int) -> torch.Tensor:
        """Channels-last ``[B, rows, C]`` copy of an internal activation of the last encode call (which must have run as
        one item range: batches of ``min_split_batch`` items or more need ``model.streams = 1`` for that)."""
        if self._last_split:
            raise _lib.MimiB200Error("debug_tap: the last encode ran as several item ranges; set model.streams = 1 first")
        rows, ch = C.c_int64(), C.c_int()
        rc = self._lib.mimi_b200_debug_tap(self._h, which, None, 0, C.byref(rows), C.byref(ch), None)
        _lib.check(self._lib, self._h, rc, "mimi_b200_debug_tap")
        B = self._last_B
        out = torch.empty((B, rows.value, ch.value), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.mimi_b200_debug_tap(self._h, which, out.data_ptr(), out.numel(), None, None, stream)
        _lib.check(self._lib, self._h, rc, "mimi_b200_debug_tap")
        return out

    LAUNCH_KINDS = (["conv0"] + [f"seanet_conv{i}" for i in range(1, 14)] +
                    ["layernorm", "qkv_gemm", "attention", "o_proj", "fc1_gelu", "fc2", "downsample_conv",
                     "rvq_input_proj", "rvq_fused", "latent_transpose", "code_fill", "halo_zero", "pad_split",
                     "front_fused"])

    DEFAULT_MODE = 9
    RANGE_SAFE_MODE = 7

    def range_overflow(self, reset: bool = False) -> bool:
        """True if an encode of the fp16 generation (mode 9, the default) that has COMPLETED on this device had to clamp an
        activation to fp16's range (|x| > 65504) -- its codes are then not the reference's. Synchronise the stream of the
        encode before asking. ``reset=True`` also clears the flag (synchronises the device). ``MimiEncoder`` checks this for
        every batch and re-encodes with the range-safe TF32 generation (mode 7); callers of :meth:`encode` do the same with
        ``set_mode(MimiB200Model.RANGE_SAFE_MODE)``."""
        with torch.cuda.device(self.device):
            return bool(self._lib.mimi_b200_range_overflow(self._h, 1 if reset else 0))

    def set_mode(self, tensor_cores) -> None:
        """True / 9 (default): fused 24 kHz front end + CTA-pair (cta_group::2) tcgen05 GEMM with every operand as an fp16 hi/lo
        pair, all three products on kind::f16 (3 tensor passes, 4 bytes per activation element; fp16 range, see
        range_overflow) + tcgen05 attention + tensor-core RVQ; 7: the same with TF32 hi (fp32) and bf16 lo operands (hi*hi and
        hi*lo on kind::tf32, lo*hi on kind::f16: 5 pass units, fp32 range); False / 0: all-fp32 FFMA (bisection baseline)."""
        mode = (self.DEFAULT_MODE if tensor_cores else 0) if isinstance(tensor_cores, bool) else int(tensor_cores)
        self.debug_set(3, mode)

    def profile(self, on: bool) -> None:
        """Switch per-launch CUDA-event profiling on/off (resets the counters)."""
        self.debug_set(2, 1 if on else 0)

    def profile_read(self) -> Dict[str, "tuple[float, int]"]:
        """{launch kind: (total ms, launches)} since the last read; synchronises."""
        n = len(self.LAUNCH_KINDS)
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(self._lib, self._h, self._lib.mimi_b200_profile_read(self._h, n, ms, cnt), "mimi_b200_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.LAUNCH_KINDS) if cnt[i]}

    @property
    def launch_count(self) -> int:
        return int(self._lib.mimi_b200_launch_count(self._h))


class EncodecFeatureExtractorLite:
    """The part of ``EncodecFeatureExtractor.__call__`` the reference scripts use: fp32 cast, zero
    right-padding to the longest item, int32 ``padding_mask``; ``ValueError`` on a wrong sampling rate
    (feature_extraction_encodec.py:122-128)."""

    sampling_rate = 24000
    padding_value = 0.0

    def __call__(self, raw_audio, sampling_rate: Optional[int] = None, return_tensors: Optional[str] = "pt",
                 padding=None):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor: {self} was trained using a sampling rate of"
                f" {self.sampling_rate}. Please make sure that the provided audio input was sampled with"
                f" {self.sampling_rate} and not {sampling_rate}.")
        batched = isinstance(raw_audio, (list, tuple)) and len(raw_audio) > 0 and isinstance(raw_audio[0], (np.ndarray, list, tuple))
        items = [np.asarray(a, dtype=np.float32) for a in (raw_audio if batched else [raw_audio])]
        for a in items:
            if a.ndim != 1:
                raise ValueError(f"Expected mono audio but example has {a.shape[-1]} channels")
        n = max(len(a) for a in items)
        iv = np.zeros((len(items), 1, n), np.float32)
        pm = np.zeros((len(items), n), np.int32)
        for i, a in enumerate(items):
            iv[i, 0, : len(a)] = a
            pm[i, : len(a)] = 1
        if return_tensors == "pt":
            return {"input_values": torch.from_numpy(iv), "padding_mask": torch.from_numpy(pm)}
        return {"input_values": iv, "padding_mask": pm}


class _Staging:
    """One slot of the wrapper's double-buffered staging: pinned input rows, their device landing buffer, a pinned
    result buffer, and the events that say when each may be reused."""

    def __init__(self):
        self.pinned: Optional[torch.Tensor] = None         # fp32 samples (pinned host)
        self.dev_in: Optional[torch.Tensor] = None         # same size, device
        self.pinned_out: Optional[torch.Tensor] = None     # bytes (pinned host): codes int64 / uint16 or UTF-8
        self.h2d_done: Optional[torch.cuda.Event] = None   # last H2D copy out of `pinned` has completed
        self.front_done: Optional[torch.cuda.Event] = None  # last reader of `dev_in` (the front end) has completed
        self.pending: Optional["PendingBatch"] = None      # result not collected yet: `pinned_out` is still in use


class PendingBatch:
    """Handle returned by :meth:`MimiEncoder.submit`; pass it to :meth:`MimiEncoder.result`."""

    def __init__(self, slot, done, view, lengths, frames, fmt, extra=None):
        self._slot, self._done, self._view = slot, done, view
        self.lengths, self.frames, self.format, self._extra = lengths, frames, fmt, extra
        self._consumed = False
        self._redo = None          # (audio_arrays, sample_rate, kwargs, serial): what result() needs to re-encode range-safely

    def ready(self) -> bool:
        return self._done is None or self._done.query()


class MimiEncoder:
    """Same methods as the reference's ``MimiEncoder`` wrapper (REF/emilia-mimi/process_shard.py:50-140).

    ``model`` may be a :class:`MimiB200Model`, a state dict, a checkpoint path, or a loaded
    ``transformers.MimiModel`` whose weights are taken over. ``num_quantizers=None`` returns all 32
    codebooks like the reference (whose callers then slice ``[:8]``); pass 8 to compute only those.

    Thread safety: every public method takes the wrapper's lock, so one instance may be shared by the worker threads of
    REF/yodas2-mimi/process_shard.py:690-717 (calls are serialised; the staging buffers are shared state).

    Pipelined use: ``submit()`` stages a batch and queues its GPU work without waiting, ``result()`` collects it;
    ``encode_stream(batches)`` keeps two batches in flight so that the staging of batch i+1 and the device->host copy of
    batch i-1 run under the kernels of batch i. ``encode_audio_batch`` is ``result(submit(...))``."""

    DEPTH = 2

    def __init__(self, model, device: str = "cuda", ragged: bool = True, num_quantizers: Optional[int] = None,
                 chunk_items: int = 16, stage_threads: int = 1, first_items: Optional[int] = None):
        self.device = device
        self.feature_extractor = EncodecFeatureExtractorLite()
        if isinstance(model, MimiB200Model):
            self.model = model
        elif isinstance(model, str):
            self.model = MimiB200Model.from_pretrained(model, device=device)
        elif isinstance(model, dict):
            self.model = MimiB200Model(model, device=device)
        else:
            self.model = MimiB200Model.from_transformers(model, device=device)
        self.ragged = ragged
        self.num_quantizers = num_quantizers
        # ragged mode encodes a batch as sub-batches of `chunk_items` items so that host staging of the next
        # sub-batch overlaps the GPU work of the previous one (items are independent, results identical)
        self.chunk_items = max(1, int(chunk_items))
        self.first_items = max(1, self.chunk_items // 2) if first_items is None else max(1, int(first_items))
        # phased=True: one padded batch staged group by group under the running front end (mimi_b200_encode_phase);
        # False: independent sub-batches (the only choice for kernel generations without the fused front end)
        self.phased = True
        # memcpy threads of mimi_b200_host_pack (1: torch copies). With one process per GPU every rank has its own pool:
        # never more threads than this rank's share of the host cores (8 ranks x 8 threads on 32 cores cost 9 % at 8 GPUs)
        ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE") or os.environ.get("WORLD_SIZE") or 1)
        self.pack_threads = max(1, min(8, (os.cpu_count() or 1) // max(1, ranks_here)))
        # every staging slot encodes on its own stream with its own workspace: the kernels of two batches in flight
        # interleave on the GPU, and the SMs that the last tiles of one batch's persistent kernels leave idle run the other's
        # (what MimiB200Model.encode does with two item ranges of ONE resident batch). Costs a second workspace.
        self.slot_streams = True
        self._slot_stream: List[Optional[torch.cuda.Stream]] = [None] * self.DEPTH
        self._lock = threading.RLock()
        self._spare = _Staging()          # staging of the (rare) range-safe re-encode
        self._serial = 0
        self._suspect_upto = -1           # batches submitted up to this serial ran while the fp16 range flag was up
        self.range_fallbacks = 0          # batches re-encoded with the range-safe generation so far
        self._slots = [_Staging() for _ in range(self.DEPTH)]
        self._cur = 0
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._pool = None
        if stage_threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=int(stage_threads), thread_name_prefix="mimi-stage")

    # -- staging buffers -------------------------------------------------------------------------------------------
    def reserve(self, batch: int, max_samples: int) -> None:
        """Pre-size the pinned staging buffers (and the model workspace) for batches of up to ``batch`` items
        of up to ``max_samples`` samples, so that no later call has to pin fresh host memory (tens of ms)."""
        K = NUM_QUANTIZERS if self.num_quantizers is None else int(self.num_quantizers)
        with self._lock:
            for j, slot in enumerate(self._slots):
                self._grow(slot, batch * max_samples, batch * K * (-(-max_samples // FRAME_SIZE)) * 8)
                if j == 0 or (self.slot_streams and self._two_workspaces_fit(batch, max_samples, K)):
                    self.model.reserve_workspace(batch, max_samples, K, slot=j)

    def _grow(self, slot: _Staging, samples: int, out_bytes: int) -> None:
        if slot.pinned is None or slot.pinned.numel() < samples:
            if slot.h2d_done is not None:
                slot.h2d_done.synchronize()
            if slot.front_done is not None:
                slot.front_done.synchronize()
            slot.pinned = slot.dev_in = None
            slot.pinned = torch.empty(max(int(samples * 1.25), 1), dtype=torch.float32).pin_memory()
            # device-side landing buffer of the same size: no allocator traffic on the hot path
            slot.dev_in = torch.empty(slot.pinned.numel(), dtype=torch.float32, device=self.model.device)
        if slot.pinned_out is None or slot.pinned_out.numel() < out_bytes:
            slot.pinned_out = None
            slot.pinned_out = torch.empty(max(int(out_bytes * 1.25), 16), dtype=torch.uint8).pin_memory()

    def _next_slot(self) -> _Staging:
        slot = self._slots[self._cur]
        slot.index = self._cur
        self._cur = (self._cur + 1) % self.DEPTH
        if slot.pending is not None and not slot.pending._consumed:
            raise _lib.MimiB200Error(f"more than {self.DEPTH} batches in flight: collect an earlier submit() with result() first")
        slot.pending = None
        if slot.h2d_done is not None:
            slot.h2d_done.synchronize()       # the previous user's copies have left the pinned rows
        return slot

    @staticmethod
    def _sub_batches(B: int, chunk_items: int, first_items: Optional[int] = None) -> List[List[int]]:
        """Item indices of the pipelined sub-batches of one ``encode_audio_batch`` call: a small first one gets the
        GPU going while the rest is still being staged, then at most two large ones (fewer, larger launches keep the
        persistent kernels efficient)."""
        bounds = [0, min(B, max(1, chunk_items // 2 if first_items is None else first_items))]
        while bounds[-1] < B:
            bounds.append(min(B, bounds[-1] + max(chunk_items, (B - bounds[1] + 1) // 2)))
        return [list(range(a, b)) for a, b in zip(bounds[:-1], bounds[1:])]

    def _check_rate(self, sample_rate: int) -> None:
        if sample_rate != self.feature_extractor.sampling_rate:
            self.feature_extractor(raw_audio=np.zeros(1, np.float32), sampling_rate=sample_rate)   # raises ValueError

    def _fill(self, buf: torch.Tensor, audio_arrays: Sequence[np.ndarray], zero_to: Sequence[int]) -> None:
        """fp32 cast + right zero-padding of every item into the pinned ``[B,1,N]`` view (what
        ``feature_extractor(..., padding=True)`` does, REF/emilia-mimi/process_shard.py:113-118). ``zero_to[i]``
        is how far item i's padding has to be materialised (N in strict mode; the end of its last kept frame
        in ragged mode, nothing beyond that is ever read)."""
        B = len(audio_arrays)
        arrs = [np.asarray(a) for a in audio_arrays]
        if self.pack_threads > 1 and B > 0 and all(a.ndim == 1 and a.dtype == np.float32 and a.flags.c_contiguous for a in arrs):
            # native gather: memcpy threads of the extension (the GIL is released for the duration of the call)
            N = buf.shape[-1]
            if any(a.shape[0] > N for a in arrs):
                raise ValueError("clip longer than the staging row")
            src = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
            lens = (C.c_int64 * B)(*[a.shape[0] for a in arrs])
            zto = (C.c_int64 * B)(*[int(z) for z in zero_to])
            rc = self.model._lib.mimi_b200_host_pack(buf.data_ptr(), buf.stride(0), src, lens, zto, B, self.pack_threads)
            _lib.check(self.model._lib, None, rc, "mimi_b200_host_pack")
            return

        def one(i):
            a = np.asarray(audio_arrays[i])
            if a.ndim != 1:
                raise ValueError(f"Expected mono audio but example has {a.shape[-1]} channels")
            n = a.shape[0]
            buf[i, 0, :n].copy_(torch.from_numpy(np.ascontiguousarray(a)))       # casts float64 -> float32
            if n < zero_to[i]:
                buf[i, 0, n:zero_to[i]].zero_()
        if self._pool is None or len(audio_arrays) < 4:
            for i in range(len(audio_arrays)):
                one(i)
        else:
            list(self._pool.map(one, range(len(audio_arrays))))

    def _streams(self):
        dev = self.model.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        return torch.cuda.current_stream(dev), self._copy_stream

    def _land(self, slot: _Staging, buf: torch.Tensor, x: torch.Tensor, audio_arrays, zero_to, b0: int, b1: int,
              consumer: torch.cuda.Stream) -> None:
        """Stage items [b0, b1): fill their pinned rows, copy them to the device on the copy stream, make ``consumer``
        wait for the copy."""
        self._fill(buf[b0:b1], audio_arrays[b0:b1], zero_to[b0:b1])
        with torch.cuda.stream(self._copy_stream):
            x[b0:b1].copy_(buf[b0:b1], non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(self._copy_stream)
        slot.h2d_done = landed
        consumer.wait_event(landed)

    def _stage(self, slot: _Staging, audio_arrays: Sequence[np.ndarray], sample_rate: int) -> torch.Tensor:
        """One pinned staging buffer and one H2D copy -> ``input_values [B,1,N]`` on the device. The
        ``padding_mask`` is never shipped: the model ignores it and the lengths are known here."""
        self._check_rate(sample_rate)
        B = len(audio_arrays)
        N = max(len(a) for a in audio_arrays)
        self._grow(slot, B * N, 0)
        main, copy = self._streams()
        if slot.front_done is not None:
            copy.wait_event(slot.front_done)     # the previous reader of the landing buffer is done
        buf = slot.pinned[: B * N].view(B, 1, N)
        x = slot.dev_in[: B * N].view(B, 1, N)
        self._land(slot, buf, x, list(audio_arrays), [N] * B, 0, B, main)
        return x

    # -- the reference's two methods ------------------------------------------------------------------------------------
    def encode_audio_chunk(self, audio_array: np.ndarray, sample_rate: int = 24000) -> np.ndarray:
        """REF/emilia-mimi/process_shard.py:63-86: one utterance -> codes ``[K, T]`` (numpy int64)."""
        with self._lock:
            return self.result(self.submit([audio_array], sample_rate))[0]

    def encode_audio_batch(self, audio_arrays: List[np.ndarray], sample_rate: int = 24000,
                           dtype=np.int64) -> List[np.ndarray]:
        """REF/emilia-mimi/process_shard.py:88-140: pad to the longest, encode, trim item i to
        ceil(len_i / 1920) frames. With ``ragged=True`` (default) the padded tails are not computed and the
        batch is staged group by group under the running front end (``mimi_b200_encode_phase``); the kept frames are the
        same either way. One device->host copy per batch instead of one per item. ``dtype=np.uint16`` returns the
        storage format of REF/yodas2-mimi/process_shard.py:519-523 (cast on the GPU, a quarter of the D2H bytes)."""
        if len(audio_arrays) == 0:
            return []
        fmt = "uint16" if np.dtype(dtype) == np.uint16 else "int64"
        with self._lock:            # submit + result as one step: concurrent callers never exceed the staging depth
            return self.result(self.submit(audio_arrays, sample_rate, fmt))

    # -- pipelined API ---------------------------------------------------------------------------------------------
    def submit(self, audio_arrays: List[np.ndarray], sample_rate: int = 24000, fmt: str = "int64",
               num_codebooks: Optional[int] = None, codebook_size: int = 2048, unicode_offset: int = 0xE000,
               _slot: Optional[_Staging] = None) -> PendingBatch:
        """Stage ``audio_arrays`` (host numpy clips at 24 kHz), queue the encode and the device->host copy of the result on
        the current stream, and return without waiting. ``fmt``: ``"int64"`` (``encode_audio_batch``'s arrays),
        ``"uint16"`` or ``"utf8"`` (``codes_to_chars`` strings of the first ``num_codebooks`` codebooks). At most
        ``DEPTH`` batches may be uncollected."""
        from . import utils
        if fmt not in ("int64", "uint16", "utf8"):
            raise ValueError(f"unknown result format '{fmt}'")
        with self._lock, torch.no_grad():
            self._check_rate(sample_rate)
            B = len(audio_arrays)
            lengths = [len(a) for a in audio_arrays]
            K = NUM_QUANTIZERS if self.num_quantizers is None else int(self.num_quantizers)
            if fmt == "utf8":
                K = int(num_codebooks) if num_codebooks is not None else min(K, 8)
                utils.validate_unicode_offset(unicode_offset, K, codebook_size)
            if B == 0:
                return PendingBatch(None, None, None, [], [], fmt)
            slot = self._next_slot() if _slot is None else _slot
            caller, copy = self._streams()
            main, ws_slot = caller, 0
            if self.slot_streams and _slot is None and self._two_workspaces_fit(B, max(lengths), K):
                if self._slot_stream[slot.index] is None:
                    self._slot_stream[slot.index] = torch.cuda.Stream(device=self.model.device)
                main, ws_slot = self._slot_stream[slot.index], slot.index
                main.wait_stream(caller)
            frames = [-(-n // FRAME_SIZE) for n in lengths]
            N = max(lengths)
            T = -(-N // FRAME_SIZE)
            phased = self.ragged and B > 1 and self.phased and self.model.supports_phased
            if self.ragged and B > 1 and not phased:
                # kernel generations without the phased call: independent sub-batches, synchronous
                arrs = self._encode_sub_batched(slot, audio_arrays, lengths, K)
                if fmt == "uint16":
                    arrs = [a.astype(np.uint16) for a in arrs]
                elif fmt == "utf8":
                    arrs = [utils.codes_to_chars(a[:K], codebook_size, unicode_offset=unicode_offset) for a in arrs]
                p = PendingBatch(None, None, None, lengths, frames, fmt, extra=arrs)
                return p
            with torch.cuda.stream(main):
                if not phased:
                    x = self._stage(slot, audio_arrays, sample_rate)
                    out = self.model.encode(x, None, num_quantizers=K, workspace_slot=ws_slot)     # B == 1 or strict mode: all T frames
                    slot.front_done = torch.cuda.Event()
                    slot.front_done.record(main)
                else:
                    # One padded [B,1,N] batch, exactly the reference's, staged group by group: the 24 kHz front end (the only
                    # part where items are independent) starts on group g as soon as its samples have landed while the host
                    # is still filling group g+1; everything after the front end runs once over the whole batch.
                    self._grow(slot, B * N, 0)
                    buf = slot.pinned[: B * N].view(B, 1, N)
                    x = slot.dev_in[: B * N].view(B, 1, N)
                    if slot.front_done is not None:
                        copy.wait_event(slot.front_done)     # the previous reader of this landing buffer is done
                    zero_to = [min(N, f * FRAME_SIZE) for f in frames]
                    arrs = list(audio_arrays)
                    groups = [(g[0], g[-1] + 1, (lambda b0=g[0], b1=g[-1] + 1: self._land(slot, buf, x, arrs, zero_to, b0, b1, main)))
                              for g in self._front_groups(B, self.first_items)]
                    out = self.model.encode(input_values=x, padding_mask=None, num_quantizers=K, valid_lengths=lengths,
                                            staged_groups=groups, on_front_done=lambda: self._mark_front_done(slot, main),
                                            workspace_slot=ws_slot)
                codes = out.audio_codes
                # result -> pinned host memory, asynchronously on the encode stream
                if fmt == "int64":
                    dev_res, shape, dt = codes, (B, K, T), torch.int64
                elif fmt == "uint16":
                    dev_res, shape, dt = utils.codes_to_uint16(codes), (B, K, T), torch.uint16
                else:
                    dev_res, blens = utils.codes_to_utf8_device(codes[:, :K], frames, codebook_size, unicode_offset)
                    shape, dt = tuple(dev_res.shape), torch.uint8
                nbytes = dev_res.numel() * dev_res.element_size()
                self._grow(slot, 0, nbytes)
                view = slot.pinned_out[:nbytes].view(dt).view(shape)
                view.copy_(dev_res, non_blocking=True)
                done = torch.cuda.Event()
                done.record(main)
            p = PendingBatch(slot, done, view, lengths, frames, fmt, extra=blens if fmt == "utf8" else None)
            self._serial += 1
            p._redo = (list(audio_arrays), sample_rate,
                       dict(num_codebooks=num_codebooks, codebook_size=codebook_size, unicode_offset=unicode_offset), self._serial)
            if _slot is None:
                slot.pending = p
            return p

    def _two_workspaces_fit(self, B: int, N: int, K: int) -> bool:
        """One workspace per staging slot costs twice the activation memory (~23 MB per padded audio-second each): only when
        both fit comfortably (a B = 64 x 60 s batch needs 90 GB per workspace -- that one runs on a single stream)."""
        nbytes = C.c_size_t()
        with self.model._lock, torch.cuda.device(self.model.device):
            rc = self.model._lib.mimi_b200_workspace_bytes(self.model._h, int(B), int(N), int(K), C.byref(nbytes))
        _lib.check(self.model._lib, self.model._h, rc, "mimi_b200_workspace_bytes")
        total = torch.cuda.get_device_properties(self.model.device).total_memory
        return 2 * nbytes.value <= 0.6 * total

    def _mark_front_done(self, slot: _Staging, stream: torch.cuda.Stream) -> None:
        slot.front_done = torch.cuda.Event()
        slot.front_done.record(stream)

    def result(self, pending: PendingBatch):
        """Wait for a submitted batch and return what ``encode_audio_batch`` returns for it (a list of ``[K, T_i]`` arrays,
        or of strings for ``fmt="utf8"``). The arrays are copies: the staging slot is free again afterwards."""
        with self._lock:
            if pending._consumed:
                raise _lib.MimiB200Error("result() was already collected for this batch")
            pending._consumed = True
            if pending._done is None:
                return pending._extra if pending._extra is not None else []
            pending._done.synchronize()
            if self.model._mode == MimiB200Model.DEFAULT_MODE and pending._redo is not None and (
                    pending._redo[3] <= self._suspect_upto or self.model.range_overflow()):
                return self._redo_range_safe(pending)
            if pending.format == "utf8":
                host = pending._view.numpy()
                return [host[i, : pending._extra[i]].tobytes().decode("utf-8") for i in range(len(pending.lengths))]
            arr = pending._view.numpy()
            return [arr[i, :, :f].copy() for i, f in enumerate(pending.frames)]

    def _redo_range_safe(self, pending: PendingBatch):
        """An activation left fp16's range somewhere in this batch (or in one that was in flight with it): encode it again
        with the TF32 generation, which has fp32's range. Rare by construction (|x| > 65504); costs a device synchronisation."""
        audio, sample_rate, kw, serial = pending._redo
        if self.model.range_overflow():
            self._suspect_upto = self._serial           # everything submitted so far ran under the raised flag
            self.model.range_overflow(reset=True)
        torch.cuda.synchronize(self.model.device)
        self.range_fallbacks += 1
        self.model.set_mode(MimiB200Model.RANGE_SAFE_MODE)
        try:
            p2 = self.submit(audio, sample_rate, pending.format, _slot=self._spare, **kw)
            p2._redo = None
            return self.result(p2)
        finally:
            torch.cuda.synchronize(self.model.device)
            self.model.set_mode(MimiB200Model.DEFAULT_MODE)

    def encode_stream(self, batches, sample_rate: int = 24000, fmt: str = "int64", **kw):
        """Generator over the results of ``batches`` (an iterable of lists of clips), in order, with two batches in flight:
        the host stages batch i+1 and copies batch i-1's codes back while the GPU encodes batch i. Each yielded item is
        what ``encode_audio_batch`` would return for that batch."""
        from collections import deque
        queue = deque()
        for batch in batches:
            queue.append(self.submit(batch, sample_rate, fmt, **kw))
            if len(queue) >= self.DEPTH:
                yield self.result(queue.popleft())
        while queue:
            yield self.result(queue.popleft())

    @staticmethod
    def _front_groups(B: int, first_items: int) -> List[List[int]]:
        """Item groups of the phased encode: a small first group gets the GPU going, every later group is twice as large
        (its staging is hidden behind the front end of the groups before it)."""
        bounds = [0, min(B, max(1, int(first_items)))]
        while bounds[-1] < B:
            bounds.append(min(B, bounds[-1] + 2 * (bounds[-1] - bounds[-2])))
        return [list(range(a, b)) for a, b in zip(bounds[:-1], bounds[1:])]

    def _encode_sub_batched(self, slot: _Staging, audio_arrays, original_lengths, K) -> List[np.ndarray]:
        """Ragged path for kernel generations without the phased call: the batch goes through the GPU as independent
        sub-batches (staging of sub-batch j+1 overlaps the encode of sub-batch j). Synchronous."""
        B = len(audio_arrays)
        chunks = self._sub_batches(B, self.chunk_items, self.first_items)
        # a sub-batch is padded to whole frames (but never beyond the full batch length): every item then sees the
        # same zeros behind its last kept frame as in the reference's single padded batch
        n_full = max(original_lengths)
        n_max = [min(n_full, -(-max(original_lengths[i] for i in ch) // FRAME_SIZE) * FRAME_SIZE) for ch in chunks]
        total = sum(len(ch) * n for ch, n in zip(chunks, n_max))
        t_max = [-(-n // FRAME_SIZE) for n in n_max]
        total_codes = sum(len(ch) * K * t for ch, t in zip(chunks, t_max))
        self._grow(slot, total, total_codes * 8)
        off = coff = 0
        host_codes = []
        main, copy = self._streams()
        copy.wait_stream(main)
        pinned_codes = slot.pinned_out[: total_codes * 8].view(torch.int64)
        for ch, N, T in zip(chunks, n_max, t_max):
            buf = slot.pinned[off: off + len(ch) * N].view(len(ch), 1, N)
            x = slot.dev_in[off: off + len(ch) * N].view(len(ch), 1, N)
            off += len(ch) * N
            lens = [original_lengths[i] for i in ch]
            # H2D on its own stream: the copy of sub-batch j+1 runs under the kernels of sub-batch j
            self._land(slot, buf, x, [audio_arrays[i] for i in ch], [min(N, -(-n // FRAME_SIZE) * FRAME_SIZE) for n in lens],
                       0, len(ch), main)
            out = self.model.encode(input_values=x, padding_mask=None, num_quantizers=K, valid_lengths=lens)
            hc = pinned_codes[coff: coff + len(ch) * K * T].view(len(ch), K, T)
            coff += len(ch) * K * T
            hc.copy_(out.audio_codes, non_blocking=True)
            host_codes.append(hc)
        main.synchronize()
        result: List[np.ndarray] = []
        for ch, hc in zip(chunks, host_codes):
            arr = hc.numpy()
            for j, i in enumerate(ch):
                result.append(arr[j, :, : -(-original_lengths[i] // FRAME_SIZE)].copy())
        return result

    # -- the callers and data formats either side of the path ---------------------------------
    def encode_native_rate_batch(self, audio_arrays: List[np.ndarray], sample_rate: int) -> List[np.ndarray]:
        """GPU front door (8f rank 1): what the reference scripts do with ``librosa.resample`` on the CPU before
        calling the wrapper (REF/librispeech-mimi/process_librispeech_train.py:189-192, REF/*/utils.py:84-87) --
        clips at their native rate go to the GPU as they are (16 kHz audio is 1.5x fewer H2D bytes), are resampled
        to 24 kHz by the polyphase kernel straight into the zero-padded ``[B,1,N]`` layout, and encoded ragged.
        Returns what ``encode_audio_batch`` returns for the resampled clips.

        The resampling filter is NOT the reference's soxr_hq (``utils.resample_audio`` explains; parity unpinned): use
        this entry point for throughput, and host-side ``librosa`` + ``encode_audio_batch`` where the reference's exact
        tokens are required."""
        from . import utils
        if len(audio_arrays) == 0:
            return []
        if sample_rate == self.feature_extractor.sampling_rate:
            return self.encode_audio_batch(audio_arrays, sample_rate)
        K = self.num_quantizers
        with self._lock, torch.no_grad():
            x, lens = utils.resample_batch(audio_arrays, sample_rate, self.feature_extractor.sampling_rate,
                                           device=self.model.device)
            out = self.model.encode(x, None, num_quantizers=K, valid_lengths=lens if self.ragged else None)
            codes = out.audio_codes.cpu().numpy()
        return [codes[i, :, : -(-n // FRAME_SIZE)].copy() for i, n in enumerate(lens)]

    def encode_to_strings(self, audio_arrays: List[np.ndarray], sample_rate: int = 24000, num_codebooks: int = 8,
                          codebook_size: int = 2048, unicode_offset: int = 0xE000,
                          audio_tags: Optional[Sequence[str]] = None) -> List[str]:
        """Codes -> storage format on the GPU (8f rank 2): the ``audio_codes[:8]`` + ``codes_to_chars`` step every
        script performs per item on the host (REF/emilia-mimi/process_shard.py:505-506), as one UTF-8 kernel and
        one device->host copy per batch. Returns the unicode strings ``codes_to_chars`` would build;
        ``audio_tags=("<|audio_start|>", "<|audio_end|>")`` wraps each one the way the pretraining-data builders do
        (REF/pretraining-data/prepare_pretraining_data.py:255-271). ``num_codebooks=1`` gives the semantic-only strings of
        REF/yodas2-mimi/build_yodas2_mm_semantic.py:169-195 (every 8th character of the 8-codebook string)."""
        if len(audio_arrays) == 0:
            return []
        with self._lock:
            strs = self.result(self.submit(audio_arrays, sample_rate, "utf8", num_codebooks=num_codebooks,
                                           codebook_size=codebook_size, unicode_offset=unicode_offset))
        if audio_tags is not None:
            a, b = audio_tags
            strs = [f"{a}{s}{b}" for s in strs]
        return strs

    def encode_long_audio(self, audio_array: np.ndarray, sample_rate: int = 24000,
                          max_chunk_duration: Optional[float] = None) -> np.ndarray:
        """Long-form audio (8f rank 3). ``max_chunk_duration=None``: the whole signal in ONE encode call -- exactly
        ``MimiModel.encode`` on the unsplit waveform, i.e. one continuous stream (every conv halo and the 250-frame
        attention window carry across what would have been piece boundaries). The engine takes items of up to 65 536
        positions at 25 Hz (43 min) and needs ~23 MB of workspace per audio-second (33 MB in the TF32 generation).
        (The reference itself counts frames with a float32 division, modeling_mimi.py:279-280, exact only below 2^24 samples =
        11.6 min; past that this engine keeps the exact integer ceil(L / stride), so only there can the two differ.)
        ``max_chunk_duration=s``: what REF/yodas2-mimi/process_shard.py:459-493 does -- pieces of at most ``s`` seconds
        encoded independently (context reset at every cut) and concatenated along time; bit-identical to that loop."""
        audio_array = np.asarray(audio_array)
        if max_chunk_duration is None:
            return self.encode_audio_chunk(audio_array, sample_rate)
        max_samples = int(max_chunk_duration * sample_rate)
        if max_samples <= 0:
            raise ValueError("max_chunk_duration must be positive")
        pieces = [audio_array[i:i + max_samples] for i in range(0, len(audio_array), max_samples)]
        return np.concatenate([self.encode_audio_chunk(p, sample_rate) for p in pieces], axis=1)
