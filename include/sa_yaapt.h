/* C ABI of the YAAPT front end on the GPU (row N2 of SURVEY.md section 8f, first step).
 *
 * What it replaces in the reference (/root/reference/satools/satools/hifigan/yaapt.py), for a whole batch at once:
 *   _yaapt lines 873-880                zero padding by frame_length / 2, the squared ("nonlinear") signal
 *   SignalObj.filtered_version 42-52    torchaudio lowpass_biquad(bp_low) -> highpass_biquad(bp_high), each clamped to [-1, 1]
 *   nlfer 148-176                       Hann-windowed frames, |DFT| summed over the F0 band
 *   PitchObj.set_energy 124-127         energy / mean(energy), voiced = energy > nlfer_thresh1
 * The reference runs this per utterance on one CPU thread (yaapt.py:27, 947-952); the outputs are what its spectral and
 * temporal trackers (spec_track, time_track) read: SignalObj.filtered of both signals, PitchObj.energy / vuv / mean_energy --
 * plus spec_track itself: its per-frame part (SHC vectors and the candidates `peaks` picks from them, sa_yaapt_shc) and its
 * per-utterance part (sa_yaapt_spec_track), and the temporal tracker with the final stages (time_track, refine, dynamic:
 * sa_yaapt_track).  Together: everything `yaapt()` computes.
 *
 * Same conventions as sa_hifigan.h: plain pointers and sizes, 0 = success, negative = error with the text in
 * sa_yaapt_last_error(); all tensor pointers are DEVICE pointers, `stream` is a cudaStream_t (NULL = default stream).
 */
#ifndef SA_YAAPT_H
#define SA_YAAPT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The options of `_yaapt` this part reads (yaapt.py:818-832), same names, same defaults. */
typedef struct sa_yaapt_params {
  double sr;            /* 16000 */
  double frame_length;  /* 35 ms */
  double frame_space;   /* 10 ms (bin/pipeline.py passes 20) */
  double f0_min;        /* 60 Hz */
  double f0_max;        /* 400 Hz */
  double fft_length;    /* 8192 */
  double bp_low;        /* 50 Hz   (handed to the LOW-pass, as the reference does) */
  double bp_high;       /* 1500 Hz (handed to the HIGH-pass) */
  double nlfer_thresh1; /* 0.75 */
  double shc_numharms;  /* 3    harmonics in the SHC product besides the fundamental (spec_track) */
  double shc_window;    /* 40 Hz  SHC window length */
  double shc_pwidth;    /* 50 Hz  peak-picking width: max_SHC = floor((f0_max + 2 shc_pwidth) / (sr / fft_length)) */
  double shc_maxpeaks;  /* 4    candidates per frame (peaks) */
  double shc_thresh1;   /* 5.0  */
  double shc_thresh2;   /* 1.25 */
  double f0_double;     /* 150 Hz */
  double f0_half;       /* 150 Hz */
  double merit_extra;   /* 0.4  */
  double median_value;  /* 7    order of the median filters (spec_track uses median_value - 2) */
  double dp5_k1;        /* 11   weight of the transition costs in dynamic5 */
  double spec_pitch_min_std; /* 0.05 */
  double tda_frame_length;   /* 35 ms (bin/pipeline.py passes 25): frame length of the time-domain analysis */
  double nccf_thresh1;       /* 0.3  (bin/pipeline.py passes 0.25) */
  double nccf_thresh2;       /* 0.9  */
  double nccf_maxcands;      /* 3    */
  double nccf_pwidth;        /* 5    */
  double merit_boost;        /* 0.2  */
  double nlfer_thresh2;      /* 0.1  */
  double merit_pivot;        /* 0.99 */
  double dp_w1, dp_w2, dp_w3, dp_w4;  /* 0.15, 0.5, 0.1, 0.9 */
} sa_yaapt_params;

const char* sa_yaapt_last_error(void);
int sa_yaapt_default_params(sa_yaapt_params* p);

/* Geometry for an utterance of n_samples (before padding): samples of the padded signal (n_samples + 2 pad) and NLFER frames
 * (`len(samples)` of nlfer, yaapt.py:164-166).  Negative on bad arguments. */
int64_t sa_yaapt_padded_length(const sa_yaapt_params* p, int64_t n_samples);
int64_t sa_yaapt_num_frames(const sa_yaapt_params* p, int64_t n_samples);

/* Scratch bytes of sa_yaapt_frontend for B utterances of at most n_max samples. */
size_t sa_yaapt_frontend_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max);

/* wav        [B, n_max] float32, item b valid in [0, lengths[b]) (lengths: HOST array, NULL = all n_max)
 * filtered   [B, n_max + 2 pad] float32   SignalObj.filtered of the padded signal      (zero beyond the item's padded length)
 * filtered_nl[B, n_max + 2 pad] float32   SignalObj.filtered of the squared signal
 * energy     [B, F_max] float32           PitchObj.energy (normalised by the item's mean; 0 beyond the item's frames)
 * vuv        [B, F_max] uint8             PitchObj.vuv
 * mean_energy[B] float32                  PitchObj.mean_energy
 * F_max = sa_yaapt_num_frames(p, n_max).  Any output pointer may be NULL (not written). */
int sa_yaapt_frontend(const sa_yaapt_params* p, const float* wav, int32_t B, int64_t n_max, const int32_t* lengths,
                      float* filtered, float* filtered_nl, float* energy, uint8_t* vuv, float* mean_energy, void* workspace,
                      size_t workspace_bytes, void* stream);

/* spec_track lines 184-231 (up to the call of `peaks`): the spectral harmonics correlation of every VOICED frame of the
 * squared signal -- 2 frame_size samples x Kaiser(beta 0.5) window, mean removed, |DFT_nfft|,
 * SHC[k] = sum_c prod_{h = 1 .. numharms + 1} |X|[h k + c - half_window],  k in [min_SHC, max_SHC] stored at index k - 1.
 * sa_yaapt_shc_length = max_SHC = the length of the vector `peaks` receives (256 for the defaults).
 * filtered_nl [B, n_max + 2 pad] and vuv [B, F_max] are sa_yaapt_frontend's outputs; shc [B, F_max, max_SHC] float32 (rows
 * of unvoiced frames and frames beyond an item's count are zero).
 * cand_pitch / cand_merit [B, maxpeaks, F_max] float32: what `peaks` (yaapt.py:383-497) returns for each of those vectors,
 * laid out like spec_track's cand_pitch / cand_merit (0 / 1 for unvoiced frames, lines 204-205).  shc, cand_pitch and
 * cand_merit may each be NULL (not written); cand_pitch and cand_merit go together. */
int64_t sa_yaapt_shc_length(const sa_yaapt_params* p);
size_t sa_yaapt_shc_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max);
int sa_yaapt_shc(const sa_yaapt_params* p, const float* filtered_nl, int32_t B, int64_t n_max, const int32_t* lengths,
                 const uint8_t* vuv, float* shc, float* cand_pitch, float* cand_merit, void* workspace, size_t workspace_bytes,
                 void* stream);

/* The rest of spec_track (yaapt.py:233-316), per utterance: voiced candidates, lowest-merit smoothing (medfilt), dynamic5 / path1
 * over the candidates, median filter, pitch_avg / pitch_std, end-point fixes, the linear re-sampling of the non-zero values and
 * the copy of elements 2, 3 into 0, 1.  cand_pitch / cand_merit [B, maxpeaks, F_max] are sa_yaapt_shc's outputs.
 * spec_pitch [B, F_max] float32 (zero beyond an item's frames), pitch_std [B] float32: what spec_track returns.
 * Items with fewer than four frames get NaN in pitch_std and zeros in spec_pitch: the reference raises IndexError there
 * (yaapt.py:311); the Python mirror raises the same. */
size_t sa_yaapt_spec_track_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max);
int sa_yaapt_spec_track(const sa_yaapt_params* p, const float* cand_pitch, const float* cand_merit, int32_t B, int64_t n_max,
                        const int32_t* lengths, float* spec_pitch, float* pitch_std, void* workspace, size_t workspace_bytes,
                        void* stream);

/* The temporal tracker and the final stages: time_track (yaapt.py:681-731, with crs_corr 577-602 and cmp_rate 609-673) on
 * the filtered signal and on the filtered squared signal, the zero padding of _yaapt (lines 921-931), refine (732-787) and
 * dynamic (321-372).  Inputs are the outputs of sa_yaapt_frontend (filtered, filtered_nl, energy, vuv) and of
 * sa_yaapt_spec_track (spec_pitch, pitch_std).  final_pitch [B, F_max] float32: `pitch.samp_values`, what yaapt() returns per
 * utterance (0 = unvoiced; zero beyond an item's frames).  Reference quirks reproduced on purpose: crs_corr removes the frame
 * mean in place from a view of the signal buffer, so later overlapping frames see shifted samples; cmp_rate looks at the first
 * local maximum of the NCCF only. */
size_t sa_yaapt_track_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max);
int sa_yaapt_track(const sa_yaapt_params* p, const float* filtered, const float* filtered_nl, const float* energy, const uint8_t* vuv,
                   const float* spec_pitch, const float* pitch_std, int32_t B, int64_t n_max, const int32_t* lengths,
                   float* final_pitch, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
