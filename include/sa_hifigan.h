/* sa_hifigan.h -- C ABI of libsatools_hifigan.so, the B200 (sm_100a) implementation of the
 * SA-toolkit HiFi-GAN generator forward (the synthesis hot path behind model.convert()).
 *
 * It is a sibling of the reference's only native extension, satools/csrc (`_satools`,
 * pybind11 + libtorch + Kaldi; conventions at satools/csrc/matrix.cc:3-70 and
 * satools/satools/chain/objf.py:56-77: caller allocates contiguous outputs, callee writes
 * in place through data pointers).  Same convention here, but plain C: no torch types, no
 * exit(), every call returns 0 on success or a negative sa_status and leaves a message in
 * sa_hifigan_last_error() (thread local).
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   sa_hifigan_create       CoreHifiGan.__init__            satools/satools/hifigan/archi.py:22-75
 *   sa_hifigan_set_weight   Module.load_state_dict of the   satools/satools/infer_helper.py:57-58
 *                           291 {weight_g,weight_v,bias}    (keys from archi.py:40-72, nn.py:96-166)
 *   sa_hifigan_finalize     weight_norm fold (recomputed by a pre-forward hook on every call in
 *                           the reference, archi.py:4,40,50,70; folded once here) and
 *                           remove_weight_norm              archi.py:109-115
 *   sa_hifigan_forward      CoreHifiGan.forward             archi.py:77-107  (called from
 *                           Net._forward, egs/vc/libritts/local/tuning/hifigan.py:99-100)
 *   sa_hifigan_synthesize_host   the H2D / convert / D2H sequence of the anonymize pipeline
 *                           satools/satools/bin/pipeline.py:104-107,148-149
 *   sa_hifigan_check        torch.cuda.synchronize() + error check (the reference gets CUDA errors
 *                           as Python exceptions from torch)
 *   sa_hifigan_set_profiling / sa_hifigan_get_profile  (no reference counterpart; the reference has
 *                           no profiler hooks) per-launch CUDA-event timing
 *   sa_hifigan_chain_timing  (no reference counterpart) in-kernel cycle accounting of the fused kernels
 *   sa_hifigan_set_debug_tap  (no reference counterpart; exposes stage activations so the
 *                           parity tests can localise a mismatch)
 *
 * Threading: a handle belongs to one process and one device and is not thread safe (the
 * reference runs one process per GPU slot with a single issuing thread, bin/anonymize:85-93).
 * All device work is enqueued on the caller's stream; nothing is allocated per call: the caller
 * owns x, y and the workspace (so torch.cuda.empty_cache() at pipeline.py:183-184 is harmless).
 */
#ifndef SA_HIFIGAN_H_
#define SA_HIFIGAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SA_HIFIGAN_ABI_VERSION 1
#define SA_HIFIGAN_MAX_STAGES 8
#define SA_HIFIGAN_MAX_RB 4

typedef struct sa_hifigan sa_hifigan;

/* Constructor arguments of CoreHifiGan (archi.py:22-32). */
typedef struct sa_hifigan_cfg {
  int32_t input_dim;                                /* imput_dim: 256 BN + 1 F0 + n speakers (504) */
  int32_t initial_channels;                         /* upsample_initial_channel (512) */
  int32_t n_stages;                                 /* len(upsample_rates) (5) */
  int32_t upsample_rates[SA_HIFIGAN_MAX_STAGES];    /* 5,4,4,2,2 */
  int32_t upsample_kernels[SA_HIFIGAN_MAX_STAGES];  /* 11,8,8,4,4 */
  int32_t n_resblocks;                              /* len(resblock_kernel_sizes) (3) */
  int32_t resblock_kernels[SA_HIFIGAN_MAX_RB];      /* 3,7,11 */
  int32_t n_dilations;                              /* dilations per ResBlock1 (3) */
  int32_t resblock_dilations[SA_HIFIGAN_MAX_RB][SA_HIFIGAN_MAX_RB]; /* 1,3,5 for each block */
  int32_t device;                                   /* CUDA ordinal; -1 = current device */
} sa_hifigan_cfg;

typedef enum sa_dtype {
  SA_DTYPE_F32 = 0,
  SA_DTYPE_F16 = 1,
  SA_DTYPE_BF16 = 2,
  SA_DTYPE_F64 = 3,
  SA_DTYPE_PCM16 = 4        /* output only: clamp(round(y * 32767)) as int16 (pipeline.py:160 PCM_S 16) */
} sa_dtype;

/* Arithmetic of the contraction stages.  Accumulation and the residual stream are fp32 in
 * every mode. */
typedef enum sa_precision {
  SA_PRECISION_FP32 = 0,    /* fp32 operands on the CUDA cores: the parity mode */
  SA_PRECISION_FP16 = 1,    /* fp16 operands on tcgen05 tensor cores (what the reference's CUDA
                               path uses under torch.amp.autocast, hifigan.py:99) */
  SA_PRECISION_BF16 = 2     /* bf16 operands on tcgen05 tensor cores (BASELINE config 3) */
} sa_precision;

typedef enum sa_status {
  SA_OK = 0,
  SA_ERR_INVALID_ARG = -1,
  SA_ERR_BAD_KEY = -2,
  SA_ERR_BAD_SHAPE = -3,
  SA_ERR_MISSING_WEIGHT = -4,
  SA_ERR_NOT_FINALIZED = -5,
  SA_ERR_WORKSPACE = -6,
  SA_ERR_CUDA = -7,
  SA_ERR_UNSUPPORTED = -8
} sa_status;

/* Which activation sa_hifigan_set_debug_tap exposes. */
typedef enum sa_debug_tap {
  SA_TAP_CONV_PRE = 0,      /* conv_pre output [B,512,T]                (archi.py:78) */
  SA_TAP_STAGE0 = 1         /* + i: output of stage i, xs/3 [B,C_i,L_i] (archi.py:86) */
} sa_debug_tap;

int sa_hifigan_abi_version(void);
const char* sa_hifigan_last_error(void);

int sa_hifigan_default_cfg(sa_hifigan_cfg* cfg);
int sa_hifigan_create(const sa_hifigan_cfg* cfg, sa_hifigan** out);
void sa_hifigan_destroy(sa_hifigan* h);

/* key: a reference state-dict key below the generator, e.g. "conv_pre.weight_g",
 * "ups.3.weight_v", "resblocks.7.convs1.2.bias", "conv_post.weight" (folded, after
 * remove_weight_norm).  data: host or device pointer, contiguous, `dtype` elements of the
 * given shape.  The data is copied; the caller may free it on return. */
int sa_hifigan_set_weight(sa_hifigan* h, const char* key, const void* data,
                          const int64_t* shape, int32_t ndim, int32_t dtype);

/* Fold weight-norm (w = g * v / ||v||_2 over all dims but 0, fp32), pack every conv into
 * the layout of the kernels of `precision`, upload.  May be called again after new
 * sa_hifigan_set_weight calls or to switch precision. */
int sa_hifigan_finalize(sa_hifigan* h, int32_t precision);

/* Frames T -> samples 320*T+1 (archi.py:88: ReflectionPad1d((1,0)) adds one sample). */
int64_t sa_hifigan_output_length(const sa_hifigan* h, int64_t T);

/* Bytes of device workspace sa_hifigan_forward needs for a batch of B items of T frames. */
size_t sa_hifigan_workspace_bytes(const sa_hifigan* h, int32_t B, int32_t T);

/* x: device, fp32, contiguous [B, input_dim, T], 16-byte aligned.
 * frames_per_item: NULL, or host int32[B] with the true frame count of each item; the
 *   tensor-core modes then only compute the tiles that can reach the first
 *   320*frames_per_item[b]+1 samples of item b (true length + 24 frames at every layer; the
 *   receptive field of the generator is 20 frames), which stay bit-identical to the padded
 *   run; the rest of item b's output is not written (the pipeline trims it, pipeline.py:156).
 *   The array is copied during the call.  NULL reproduces the reference's padded semantics.
 * y: device, contiguous [B, 1, 320*T+1] of y_dtype (F32, F16 or PCM16).
 * stream: a cudaStream_t passed as void* (NULL = legacy default stream). */
int sa_hifigan_forward(sa_hifigan* h, const float* x, int32_t B, int32_t T,
                       const int32_t* frames_per_item, void* y, int32_t y_dtype,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The same forward fed with the conditioning PARTS instead of the concatenated tensor: what
 * Net._forward (egs/vc/libritts/local/tuning/hifigan.py:83-97) builds with
 *   x = cat([bn, F.interpolate(f0, T), F.interpolate(spk_id.unsqueeze(2), T)], dim=1)
 * bn [B, n_bn, T], f0 [B, 1, T] (already normalised / transformed / interpolated to T) and
 * spk [B, n_spk] (one-hot rows, constant in time), device fp32, n_bn + 1 + n_spk == input_dim.
 * The [B, input_dim, T] tensor is never materialised; the result is bit-identical to
 * sa_hifigan_forward on the concatenation.  Tensor-core modes only (fp32 parity mode:
 * SA_ERR_UNSUPPORTED, assemble x). */
int sa_hifigan_forward_parts(sa_hifigan* h, const float* bn, int32_t n_bn, const float* f0,
                             const float* spk, int32_t n_spk, int32_t B, int32_t T,
                             const int32_t* frames_per_item, void* y, int32_t y_dtype,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Same with HOST buffers (pinned for full speed): copies x to the device, runs the forward,
 * copies y back and waits.  dev_scratch must hold
 * sa_hifigan_host_scratch_bytes(h,B,T,y_dtype) bytes (staging for x and y + the workspace). */
size_t sa_hifigan_host_scratch_bytes(const sa_hifigan* h, int32_t B, int32_t T, int32_t y_dtype);
int sa_hifigan_synthesize_host(sa_hifigan* h, const float* x_host, int32_t B, int32_t T,
                               const int32_t* frames_per_item, void* y_host, int32_t y_dtype,
                               void* dev_scratch, size_t dev_scratch_bytes, void* stream);
/* Stream-ordered form of the same sequence: enqueues H2D copy, forward and D2H copy on `stream`
 * and returns without waiting.  x_host, y_host and dev_scratch stay owned by the call until the
 * caller has synchronized `stream`.  The copies run on `stream`; the kernels of all in-flight
 * calls of a handle run in submission order on one internal stream (event-linked to `stream`),
 * so two batches never compete for the SMs.  Two streams used alternately (each with its own buffers)
 * overlap the copies of one batch with the kernels of the other -- the pipeline's DataLoader
 * prefetch (bin/pipeline.py:91-101) moved to the device side. */
int sa_hifigan_synthesize_host_async(sa_hifigan* h, const float* x_host, int32_t B, int32_t T,
                                     const int32_t* frames_per_item, void* y_host, int32_t y_dtype,
                                     void* dev_scratch, size_t dev_scratch_bytes, void* stream);

/* The stream-ordered host entry fed with the conditioning parts (see sa_hifigan_forward_parts):
 * bn_host [B, n_bn, T], f0_host [B, 1, T], spk_host [B, n_spk] on the host (pinned for overlap):
 * (n_bn + 1) / input_dim of the H2D bytes of the concatenated tensor.  Same dev_scratch size. */
int sa_hifigan_synthesize_host_parts_async(sa_hifigan* h, const float* bn_host, int32_t n_bn,
                                           const float* f0_host, const float* spk_host,
                                           int32_t n_spk, int32_t B, int32_t T,
                                           const int32_t* frames_per_item, void* y_host,
                                           int32_t y_dtype, void* dev_scratch,
                                           size_t dev_scratch_bytes, void* stream);

/* Compact conditioning.  What Net._forward concatenates is highly redundant: the ASR bottleneck
 * features are rows of the extractor's VQ codebook (48 codewords, satools/satools/chain/nn.py:427-459, selected in
 * egs/asr/librispeech/local/chain/tuning/tdnnf_wav2vec2_vq.py:290-314) and the speaker block is a one-hot that is constant
 * in time (egs/vc/libritts/local/tuning/hifigan.py:94-97).  sa_hifigan_set_codebook uploads the codebook [n_codes, dim]
 * (host or device pointer, fp32, n_codes <= 255, dim + 1 + n_speakers == input_dim); sa_hifigan_forward_vq then takes
 *   vq_idx  uint8 [B, T]   code index per frame (>= n_codes: a zero vector, the pipeline's padding frames)
 *   f0      fp32  [B, T]   normalised / transformed F0 at T frames (channel `dim`)
 *   spk_ids int32 [B]      target speaker per item (channel dim + 1 + id is 1, constant in time)
 * all on the device: 5 bytes per frame + 4 per item instead of 4 * input_dim bytes per frame.  Bit-identical to
 * sa_hifigan_forward on the assembled tensor when its BN rows are exact codewords.  Tensor-core modes only. */
int sa_hifigan_set_codebook(sa_hifigan* h, const float* codebook, int32_t n_codes, int32_t dim);
int sa_hifigan_forward_vq(sa_hifigan* h, const uint8_t* vq_idx, const float* f0, const int32_t* spk_ids,
                          int32_t B, int32_t T, const int32_t* frames_per_item, void* y, int32_t y_dtype,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Nearest-codeword assignment (the last step of the ASR-BN extractor): what VectorQuantizerEMA.forward
 * does in eval mode (satools/satools/chain/nn.py:402-477) at the end of extract_bn
 * (egs/asr/librispeech/local/chain/tuning/tdnnf_wav2vec2_vq.py:96-112,312), against the codebook of sa_hifigan_set_codebook.
 *   bn        fp32 [n_rows, dim]  the pre-quantisation bottleneck rows (the [N, T, C] tensor flattened as inputs.view(-1, C))
 *   vq_idx    uint8 [n_rows]      encoding_indices (first minimum of |x|^2 + |e|^2 - 2 x.e, chain/nn.py:423-436)
 *   quantized fp32 [n_rows, dim]  or NULL: the returned tensor, inputs + (codeword - inputs) in fp32 (chain/nn.py:448-456)
 * all on the device.  vq_idx [B, T] is exactly what sa_hifigan_forward_vq consumes, so the dense [B, dim, T] feature tensor
 * never has to exist.  dim <= 512. */
int sa_hifigan_vq_assign(sa_hifigan* h, const float* bn, int64_t n_rows, uint8_t* vq_idx,
                         float* quantized, void* stream);

/* Stream-ordered host entries with TRIMMED output (the reference copies the whole padded batch back and
 * trims on the host, bin/pipeline.py:148-156): frames_per_item is required; after the forward only the kept samples of
 * every item, 320 * frames_per_item[b] + 1, are copied to y_host, packed one item after the other (item b starts at
 * sample sum_{i<b} (320 * frames_per_item[i] + 1)).  With y_dtype = SA_DTYPE_PCM16 that is what torchaudio.save(...,
 * encoding="PCM_S", bits_per_sample=16) writes (pipeline.py:160).  Same dev_scratch size as the untrimmed entries.
 * The _vq_ form takes the compact conditioning from host memory (pinned for overlap). */
int sa_hifigan_synthesize_host_trimmed_async(sa_hifigan* h, const float* x_host, int32_t B, int32_t T,
                                             const int32_t* frames_per_item, void* y_host,
                                             int32_t y_dtype, void* dev_scratch,
                                             size_t dev_scratch_bytes, void* stream);
int sa_hifigan_synthesize_host_vq_trimmed_async(sa_hifigan* h, const uint8_t* vq_host,
                                                const float* f0_host, const int32_t* spk_ids_host,
                                                int32_t B, int32_t T, const int32_t* frames_per_item,
                                                void* y_host, int32_t y_dtype, void* dev_scratch,
                                                size_t dev_scratch_bytes, void* stream);

/* Debug tap: while `out` is non-NULL every following forward also writes activation `tap`
 * as fp32 [B,C,L] to `out` (device memory, caller sized: conv_pre B*initial_channels*T,
 * stage i B*C_i*L_i floats).  out = NULL switches it off. */
int sa_hifigan_set_debug_tap(sa_hifigan* h, int32_t tap, float* out);

/* Per-launch device timing (off by default; costs two CUDA events per launch).  While it is
 * on, every forward records an event before each kernel launch.  sa_hifigan_get_profile waits
 * for the stream and returns, for each launch of the most recent forward in launch order, its
 * duration in milliseconds and a tag = 16 * section + kind, where section 0 is conv_pre
 * (and input packing), 1 + i is stage i, n_stages + 1 is the tail, and kind 0 is the section's
 * own conv (conv_pre / upsampler / conv_post), 1 + j a conv of ResBlock j, 15 anything else.
 * Returns the number of launches (<= max_n entries are written). */
int sa_hifigan_set_profiling(sa_hifigan* h, int32_t enable);
int sa_hifigan_get_profile(sa_hifigan* h, float* ms, int32_t* tags, int32_t max_n);

/* Wait for `stream` and report any device-side failure of the work enqueued so far (a CUDA
 * error, or a tensor-core kernel that gave up waiting on a barrier).  0 = all good. */
int sa_hifigan_check(sa_hifigan* h, void* stream);

/* Diagnostics (enabled by SATOOLS_B200_CHAIN_TIMING=1 in the environment at finalize time): cycle
 * counters of the fused ResBlock kernels of the most recent forward, 16 int64 per launch in launch
 * order (MMA warp: total, wait-activations, wait-weights, issue; one epilogue warp: total, load+stage x,
 * wait-accumulator, work, of which TMEM loads, of which fence+arrive; rest unused), summed over CTAs.
 * Returns the number of launches written. */
int sa_hifigan_chain_timing(sa_hifigan* h, int64_t* out, int32_t max_launches);

/* Kernel launches enqueued by the most recent forward on this handle. */
int64_t sa_hifigan_last_launch_count(const sa_hifigan* h);

#ifdef __cplusplus
}
#endif
#endif /* SA_HIFIGAN_H_ */
