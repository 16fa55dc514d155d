/*
 * mimi_b200.h -- C ABI of the B200-native Mimi encode hot path (waveform -> discrete codes).
 *
 * This is the drop-in boundary. The reference (potsawee/tokenize-audio) is pure
 * Python and reaches the GPU through `transformers.MimiModel.encode`; there is no reference FFI, so
 * every entry point below cites the reference-side Python interface it stands in for:
 *
 *   mimi_b200_encode          <- MimiModel.encode            transformers/models/mimi/modeling_mimi.py:1522-1611
 *                                called at REF/emilia-mimi/process_shard.py:81-84, :124-127,
 *                                REF/emilia-mimi/utils.py:63-66 (and the 9 other MimiEncoder copies)
 *   mimi_b200_load_weights    <- MimiModel.from_pretrained   REF/emilia-mimi/process_shard.py:57-59
 *                                (+ lazy MimiEuclideanCodebook.embed, modeling_mimi.py:1191-1195)
 *   mimi_b200_resample        <- utils.resample_audio        REF/emilia-mimi/utils.py:84-87 (librosa.resample)
 *   mimi_b200_codes_to_utf8   <- utils.codes_to_chars        REF/emilia-mimi/utils.py:18-37,
 *                                REF/pretraining-data/converter.py:17-37
 *   mimi_b200_encoded_frames  <- MimiModel.get_encoded_length modeling_mimi.py:1490-1503
 *   mimi_b200_decode          <- MimiModel.decode            modeling_mimi.py:1613-1679, called by utils.str_to_audio,
 *                                REF/emilia-mimi/utils.py:72-81 (round-trip spot checks)
 *
 * Conventions: plain pointers and sizes only, no C++/torch types, no exceptions across the boundary.
 * Every function returns an int status (MIMI_B200_OK == 0). `d_` pointers are device memory on the
 * handle's device, `h_` pointers are host memory. The caller owns inputs, outputs and the workspace;
 * the library owns its packed weights. Calls are asynchronous on `stream` (a cudaStream_t passed as
 * void*) and NOT re-entrant per handle. There is no CPU fallback: without a CUDA device every call
 * fails with MIMI_B200_ERR_CUDA.
 */
#ifndef MIMI_B200_H
#define MIMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMI_B200_ABI_VERSION 1

enum {
  MIMI_B200_OK = 0,
  MIMI_B200_ERR_ARG = 1,        /* bad argument (message via mimi_b200_last_error) */
  MIMI_B200_ERR_CUDA = 2,       /* CUDA runtime / driver error */
  MIMI_B200_ERR_STATE = 3,      /* e.g. encode before load_weights */
  MIMI_B200_ERR_WORKSPACE = 4   /* workspace too small */
};

/* fixed architecture of kyutai/mimi == transformers MimiConfig() defaults */
#define MIMI_B200_SAMPLE_RATE 24000
#define MIMI_B200_FRAME_SIZE 1920      /* 24 kHz samples per code frame (12.5 Hz) */
#define MIMI_B200_HIDDEN 512
#define MIMI_B200_NUM_CONVS 14
#define MIMI_B200_NUM_LAYERS 8
#define MIMI_B200_MAX_QUANTIZERS 32
#define MIMI_B200_CODEBOOK_SIZE 2048
#define MIMI_B200_CODEBOOK_DIM 256

typedef struct mimi_b200 mimi_b200_t;

/* One transformer layer, names as in MimiModel.state_dict(): encoder_transformer.layers.{l}.* */
typedef struct {
  const float* input_layernorm_weight;          /* [512] */
  const float* input_layernorm_bias;            /* [512] */
  const float* q_proj_weight;                   /* [512,512] (out,in) */
  const float* k_proj_weight;
  const float* v_proj_weight;
  const float* o_proj_weight;
  const float* self_attn_layer_scale;           /* [512] */
  const float* post_attention_layernorm_weight; /* [512] */
  const float* post_attention_layernorm_bias;
  const float* fc1_weight;                      /* [2048,512] */
  const float* fc2_weight;                      /* [512,2048] */
  const float* mlp_layer_scale;                 /* [512] */
} mimi_b200_layer_weights_t;

/* Host fp32 arrays in state-dict layout. The library copies and repacks; the caller keeps ownership. */
typedef struct {
  /* encoder.layers.{0, 1.block.1, 1.block.3, 3, 4.block.1, 4.block.3, 6, 7.block.1, 7.block.3, 9,
     10.block.1, 10.block.3, 12, 14}.conv.{weight [C_out,C_in,k], bias [C_out]} in execution order */
  const float* conv_weight[MIMI_B200_NUM_CONVS];
  const float* conv_bias[MIMI_B200_NUM_CONVS];
  mimi_b200_layer_weights_t layer[MIMI_B200_NUM_LAYERS];
  const float* downsample_weight;               /* downsample.conv.weight [512,512,4] */
  const float* semantic_input_proj_weight;      /* quantizer.semantic_...input_proj.weight [256,512,1] */
  const float* acoustic_input_proj_weight;      /* quantizer.acoustic_...input_proj.weight [256,512,1] */
  /* index 0 = semantic layer 0, 1..31 = acoustic layers 0..30 */
  const float* embed_sum[MIMI_B200_MAX_QUANTIZERS];      /* [2048,256] */
  const float* cluster_usage[MIMI_B200_MAX_QUANTIZERS];  /* [2048] */
  /* optional [32] rotary inverse frequencies exactly as MimiRotaryEmbedding computes them
     (modeling_mimi.py:538-560); NULL = computed by the library */
  const float* rope_inv_freq;
} mimi_b200_weights_t;

/* Decode-side tensors of MimiModel.state_dict(), host fp32. */
typedef struct {
  const float* semantic_output_proj_weight;     /* quantizer.semantic_residual_vector_quantizer.output_proj.weight [512,256,1] */
  const float* acoustic_output_proj_weight;     /* quantizer.acoustic_residual_vector_quantizer.output_proj.weight [512,256,1] */
  const float* upsample_weight;                 /* upsample.conv.weight [512,1,4] (depthwise ConvTranspose1d, stride 2) */
  mimi_b200_layer_weights_t layer[MIMI_B200_NUM_LAYERS];   /* decoder_transformer.layers.{l}.* */
  const float* conv_in_weight;                  /* decoder.layers.0.conv.weight [1024,512,7] */
  const float* conv_in_bias;                    /* [1024] */
  /* the four upsampling stages, ratios 8, 6, 5, 4: decoder.layers.{2,5,8,11}.conv (ConvTranspose1d weight [C_in,C_out,2r],
     bias [C_out]) and their residual blocks decoder.layers.{3,6,9,12}.block.{1,3}.conv (weight [C_out,C_in,k], bias) */
  const float* up_weight[4];
  const float* up_bias[4];
  const float* res_a_weight[4];
  const float* res_a_bias[4];
  const float* res_b_weight[4];
  const float* res_b_bias[4];
  const float* conv_out_weight;                 /* decoder.layers.14.conv.weight [1,64,3] */
  const float* conv_out_bias;                   /* [1] */
} mimi_b200_decoder_weights_t;

/* Library / ABI identification. */
int mimi_b200_abi_version(void);

/* Create an engine on CUDA device `device_ordinal`. Fails (no CPU fallback) if there is none. */
int mimi_b200_create(mimi_b200_t** out, int device_ordinal);
void mimi_b200_destroy(mimi_b200_t* h);

/* Last error message for this handle (or for create() when h == NULL); owned by the library. */
const char* mimi_b200_last_error(const mimi_b200_t* h);

int mimi_b200_load_weights(mimi_b200_t* h, const mimi_b200_weights_t* host_weights);

/* T = ceil(N / 1920): frames produced for N input samples (get_encoded_length). */
int64_t mimi_b200_encoded_frames(int64_t n_samples);

/* Bytes of scratch `mimi_b200_encode` needs for a [B,1,N] batch with K codebooks. */
int mimi_b200_workspace_bytes(mimi_b200_t* h, int B, int64_t N, int K, size_t* out_bytes);

/*
 * Encode d_input [B,1,N] fp32 (24 kHz) into d_codes [B,K,T] int64, T = ceil(N/1920).
 *  h_valid_len  NULL  : strict mode -- every item is encoded over all N samples, exactly what
 *                       MimiModel.encode does with a padded batch (the padding mask is ignored there).
 *               [B]   : ragged mode -- item i only needs frames < ceil(h_valid_len[i]/1920) (what
 *                       MimiEncoder.encode_audio_batch keeps, REF/emilia-mimi/process_shard.py:132-139);
 *                       work on the padded tail is skipped, kept frames are identical to strict mode
 *                       (the network is causal), all other frames are written as 0.
 *  d_latent_opt NULL or [B,512,T] fp32: pre-quantisation latent (downsample conv output) for parity dumps.
 */
int mimi_b200_encode(mimi_b200_t* h, const float* d_input, int B, int64_t N,
                     const int64_t* h_valid_len, int K, int64_t* d_codes, float* d_latent_opt,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * The same encode in phases, so that a host that stages its batch group by group (fill pinned memory -> H2D) can start
 * the GPU on the first group while it is still staging the others. Items are independent only in the 24 kHz front end
 * (everything after it runs once over the whole batch), so the phases are:
 *   MIMI_B200_PHASE_BEGIN   lengths to the device, halo rows, zeroed codes (needs no input samples);
 *   MIMI_B200_PHASE_FRONT   fused front end for items [b0, b1) -- their samples must have landed in d_input;
 *   MIMI_B200_PHASE_FINISH  the rest of the pipeline for the whole batch.
 * Every phase takes the arguments of mimi_b200_encode, identical from call to call; BEGIN, FRONT over a partition of
 * [0, B), FINISH on one stream is exactly mimi_b200_encode. This is what MimiEncoder.encode_audio_batch
 * (REF/emilia-mimi/process_shard.py:88-140) does with its pinned staging buffer. Needs a tensor-core generation (7 or 9).
 */
#define MIMI_B200_PHASE_BEGIN 1
#define MIMI_B200_PHASE_FRONT 2
#define MIMI_B200_PHASE_FINISH 3
int mimi_b200_encode_phase(mimi_b200_t* h, int phase, int b0, int b1, const float* d_input, int B, int64_t N,
                           const int64_t* h_valid_len, int K, int64_t* d_codes, float* d_latent_opt,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/* Debug/parity taps: copy an internal activation of the LAST encode call into d_out (channels-last
   [B, rows, C] fp32). `which`: 0..13 = output of SEANet conv i (after residual add for block.3 convs),
   100+l = transformer layer l output; 200 = the latent, 201 = the RVQ input projections [sem | aco] (in generation 9 the
   acoustic half holds the last stage's residual after an encode with more than two codebooks: rvq_f16.cuh keeps the fp32
   residual there). Returns rows/C through the out params. */
int mimi_b200_debug_tap(mimi_b200_t* h, int which, float* d_out, size_t out_capacity_floats,
                        int64_t* rows_per_item, int* channels, void* stream);

/* Debug knobs (parity bisection and A/B measurements only; MimiB200Model.debug_set):
     0  number of transformer layers to run (default 8)
     1  index of the last SEANet conv to run (default 13; smaller values stop the pipeline there, on the fp32 FFMA path,
        and leave d_codes untouched)
     2  per-launch CUDA-event profiling on/off (resets the profile)
     3  kernel generation ("mode"), default 9:
          9  fused 24 kHz front end (front_f16.cuh) + CTA-pair tcgen05 GEMM (tc_gemm5.cuh; tc_gemm7.cuh for the convs whose taps
             share input rows) + tcgen05 attention + tensor-core RVQ (rvq_f16.cuh), with every GEMM operand as an fp16
             hi/lo pair (activations: hi + lo/2048, weights row-scaled by a power of two): hi*hi + hi*lo + lo*hi all on
             kind::f16, 3 tensor passes, 4 bytes per activation element. fp16 ends at 65504 (see mimi_b200_range_overflow)
          7  the same with TF32 hi (fp32) and bf16 lo operands: hi*hi and hi*lo on kind::tf32, lo*hi on kind::f16 (5 pass
             units, 6 bytes per element); fp32 range -- the fallback of mode 9
          0  every layer on fp32 FFMA (exact-fp32 bisection baseline; also what the decode direction runs on)
     5  k-blocks per accumulation chunk of the GEMM (0 = default 4, i.e. K = 128 inside TMEM between drains)
     9  1 = pair tiles of 256 columns wherever N allows (default 0: 256 columns only for K > 2048, 128 otherwise)
    10  1 = never flatten the row dimension of the linears across items
    11  1 = k-blocks in linear order (default 0: taps grouped by tau mod s so that re-read input rows hit L2)
    13  1 = walk the mt_max x B tile grid instead of the compact tile lists of a ragged call
    16  1 = first-draft one-thread-per-output resampler instead of resample_poly_kernel
    17  1 = generation 9 runs round 1's front end (front_fused.cuh: TF32 internals) instead of front_f16.cuh
    18  1 = the attention kernel walks the mt_max x B unit grid instead of the compact tile list of a ragged call
    19  1 = generation 9 runs the TF32 RVQ kernel (rvq_tc.cuh) instead of the fp16-pair one (rvq_f16.cuh)
    21  1 = fc1's GELU through erff instead of gelu_fast (common.cuh; same function to <= 4e-7 absolute)
    20  1 = convs never run as tap groups (tc_gemm7.cuh): every k-block of the im2col view is fetched on its own (tc_gemm5.cuh) */
int mimi_b200_debug_set(mimi_b200_t* h, int key, int value);

/* Read and reset the per-launch profile gathered since profiling was switched on: for launch kind
   id < max_ids, sum_ms[id] = total device time between the previous launch's end and this launch's end on
   the encode stream, count[id] = launches. Kinds: 0 conv0; 1..13 SEANet conv i (gather-GEMM); 14
   LayerNorm; 15 QKV; 16 attention; 17 o_proj; 18 fc1; 19 fc2; 20 downsample conv; 21 RVQ input_proj;
   22 fused RVQ; 23 latent transpose; 24 code fill; 25 halo zeroing; 26 replicate-pad + split; 27 fused front end. Synchronises with the last recorded launch. */
int mimi_b200_profile_read(mimi_b200_t* h, int max_ids, double* sum_ms, int64_t* count);

/* Unit-test hook for the tensor-core GEMM kernel: d_out[M][N] = act(d_a[M][K] * h_w[N][K]^T + bias) through the
   encoder's own operand split + TMA + tcgen05 path (N % 64 == 0, K % 32 == 0; act 1 = GELU(erf)), in the tensor-core
   generation selected by debug_set key 3 (7 or 9). Synchronises the stream. */
int mimi_b200_debug_tc_gemm(mimi_b200_t* h, const float* d_a, const float* h_w, const float* d_bias_opt, int M,
                            int N, int K, int act, float* d_out, void* stream);

/* Hardware probe (unit test only; the fused front end relies on the property): one 128 x 64 single-pass TF32 tile whose SWIZZLE_128B A-operand
   descriptor starts `shift` rows (0..8) into the staged tile: d_out[m][n] = sum_k d_a[m + shift][k] * h_w[n][k].
   d_a has 136 rows of K floats (K % 32 == 0), h_w 64 rows. base_mode 1 also sets the descriptor's
   base-offset field to (start >> 7) & 7. Synchronises the stream. */
int mimi_b200_debug_shift_probe(mimi_b200_t* h, const float* d_a, const float* h_w, int K, int shift, int base_mode,
                                float* d_out, void* stream);

/* Output length of the resampler for n input samples: ceil(n * sr_out / sr_in) (librosa fix=True). */
int64_t mimi_b200_resample_out_len(int64_t n_in, int sr_in, int sr_out);

/*
 * Polyphase FIR resampler. d_in [B, in_stride] fp32 rows with h_len[i] valid samples each, d_out
 * [B, out_stride] fp32; row i gets resample_out_len(h_len[i]) samples followed by zeros up to
 * out_stride (so the result is directly the zero-right-padded [B,1,N] batch the encoder takes).
 * sr_in == sr_out copies.
 */
int mimi_b200_resample(mimi_b200_t* h, const float* d_in, int64_t in_stride, const int64_t* h_len, int B,
                       int sr_in, int sr_out, float* d_out, int64_t out_stride, void* stream);

/* Bytes of UTF-8 one frame of K codebooks takes (e.g. 28 for K=8, offset 0xE000, size 2048);
   returns -1 for an offset whose range touches the surrogates U+D800..DFFF (converter.py:68-81). */
int64_t mimi_b200_utf8_bytes_per_frame(int K, uint32_t unicode_offset, int codebook_size);

/*
 * codes -> UTF-8 of codes_to_chars: item i, frame t < h_frames[i] (NULL = all T), codebook k becomes
 * code point unicode_offset + k*codebook_size + d_codes[i,k,t], frame-major / codebook-minor.
 * d_out is [B, out_stride] bytes; h_out_len_opt[i] (host, may be NULL) receives the byte length of row i.
 */
int mimi_b200_codes_to_utf8(mimi_b200_t* h, const int64_t* d_codes, int B, int K, int64_t T,
                            const int64_t* h_frames, uint32_t unicode_offset, int codebook_size,
                            uint8_t* d_out, int64_t out_stride, int64_t* h_out_len_opt, void* stream);

/*
 * Decode direction: d_codes [B,K,T] int64 (K = 1..32 codebooks, as MimiModel.decode accepts) -> d_audio [B,1,1920*T] fp32
 * at 24 kHz. Needs load_weights (codebooks) and load_decoder_weights. Runs on the exact-fp32 FFMA kernels: this is the
 * reference's round-trip spot-check path (REF/emilia-mimi/utils.py:72-81), not a throughput path. Codes outside [0, 2048)
 * contribute nothing (the caller validates them; F.embedding would raise).
 */
int mimi_b200_load_decoder_weights(mimi_b200_t* h, const mimi_b200_decoder_weights_t* host_weights);
int mimi_b200_decode_workspace_bytes(mimi_b200_t* h, int B, int64_t T, size_t* out_bytes);
int mimi_b200_decode(mimi_b200_t* h, const int64_t* d_codes, int B, int K, int64_t T, float* d_audio, void* d_workspace,
                     size_t workspace_bytes, void* stream);

/*
 * codes int64 -> uint16, n elements of any shape (the `codes.astype(np.uint16)` storage format of
 * REF/yodas2-mimi/process_shard.py:519-523, done before the device->host copy: 2 instead of 8 bytes per code cross PCIe).
 */
int mimi_b200_codes_pack_u16(mimi_b200_t* h, const int64_t* d_codes, int64_t n, uint16_t* d_out, void* stream);

/*
 * Host-side staging helper of the wrapper (no device work): gathers n ragged fp32 clips into the rows of a (pinned)
 * [n, dst_stride] buffer -- row i = h_src[i][0 .. h_len[i]) followed by zeros up to h_zero_to[i] (<= dst_stride) -- with a
 * small pool of memcpy threads. This is the right-zero-padding of EncodecFeatureExtractor.__call__
 * (feature_extraction_encodec.py:81-202) as REF/emilia-mimi/process_shard.py:113-118 uses it, minus the Python loop;
 * a single core's memcpy (~10 GB/s) would otherwise bound the end-to-end rate.
 */
int mimi_b200_host_pack(float* h_dst, int64_t dst_stride, const float* const* h_src, const int64_t* h_len,
                        const int64_t* h_zero_to, int n, int n_threads);

/*
 * Range guard of the default (fp16 split, "mode 9") kernel generation: activations are carried as fp16 hi/lo pairs, which
 * saturate at 65504 (MimiModel itself runs fp32 and has no such limit). Kernels clamp and raise a device flag; every mode-9
 * encode copies the flag to the host behind its last kernel. Returns the flag as of the last COMPLETED encode on this handle's
 * device (synchronise the encode's stream first); reset != 0 also clears it (synchronises the device). A caller that sees 1
 * re-encodes with debug_set(3, 7), the TF32 generation, which has fp32's range -- MimiEncoder does that by itself.
 */
int mimi_b200_range_overflow(mimi_b200_t* h, int reset);

/* Number of kernels this handle has launched since creation (bench.py reports it as gpu_launches). */
int64_t mimi_b200_launch_count(const mimi_b200_t* h);

#ifdef __cplusplus
}
#endif
#endif /* MIMI_B200_H */
