/*
 * hpss_b200.h -- C ABI of the B200-native HPSS feature front-end.
 *
 * Drop-in boundary for the feature path of mrinmoy-iitg/SM_HPSS_MTL
 * (lib/preprocessing.py: get_featuregram :355-457, get_feature_patches :137-292,
 * get_data_stats :461-586, scale_data :590-614; lib/cython_impl/tools.pyx:21-38,
 * 138-166).  The reference implements this path by calling librosa / scipy /
 * sklearn on the CPU; every entry point below names the reference call it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no C++/torch types.
 *   - every function returns an int status (HPSS_OK == 0); the message of the last
 *     failure on the calling thread is available from hpss_last_error().
 *   - pointers named *_dev are device pointers on the context's GPU, *_host are
 *     host pointers.  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  Device entry points enqueue work on `stream` and return without
 *     synchronising the host.
 *   - A *batch* describes many independent clips processed by one launch.
 *     Waveforms are concatenated back to back (clip c starts at sample_off[c]);
 *     every per-frame array is stored clip after clip, each clip a C-ordered
 *     (rows, T_c) matrix exactly as the reference's numpy arrays:
 *         element (c, r, t)  at  rows * frame_off[c] + r * T_c + t.
 *     The reference's one-file-per-call signature is the n_clips == 1 case.
 *   - No CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef HPSS_B200_H
#define HPSS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define HPSS_API
#else
#define HPSS_API __attribute__((visibility("default")))
#endif

enum hpss_status {
    HPSS_OK = 0,
    HPSS_ERR_INVALID = 1,        /* bad argument / shape                                  */
    HPSS_ERR_SHORT_SIGNAL = 2,   /* n_fft > len(y): librosa.stft raises ParameterError     */
    HPSS_ERR_UNSUPPORTED = 3,    /* e.g. n_fft/2 has a prime factor other than 2, 3, 5     */
    HPSS_ERR_CUDA = 4,           /* CUDA runtime error (message has the CUDA string)       */
    HPSS_ERR_NEGATIVE = 5,       /* softmask input < 0: librosa.util.softmask raises       */
    HPSS_ERR_NOMEM = 6,
    HPSS_ERR_NONFINITE = 7       /* non-finite audio: librosa.util.valid_audio raises      */
};
/* HPSS_ERR_NEGATIVE / HPSS_ERR_NONFINITE are data-dependent: device entry points never synchronise, so the
 * kernels that see the data (the signal preparation, hpss_validate_audio / hpss_validate_nonneg,
 * hpss_featuregram_from_spec) set a bit in a device status word and hpss_ctx_check(), called where the
 * caller synchronises anyway, turns it into the status.  The host-buffer entries (hpss_featuregram_host,
 * hpss_pipeline_run) validate their input and return these codes themselves. */

/* sample formats of decoded audio handed to the signal preparation */
enum hpss_pcm_format {
    HPSS_PCM_F32 = 0,            /* float32 samples (what librosa.load returns)             */
    HPSS_PCM_S16 = 1             /* 16-bit PCM as stored in MUSAN's wav files; x / 32768 on the device */
};

/* Feature families of get_featuregram's name dispatch (lib/preprocessing.py:378-444).
 * The dispatch is by startswith(), so e.g. LogMelHarmSpec / LogMelPercSpec /
 * LogMelHarmPercSpec all map to HPSS_FEAT_LOGMEL_HARMPERC: both streams are always
 * computed and stacked harmonic rows first, then percussive rows (:411,:423,:433,:443). */
enum hpss_feature {
    HPSS_FEAT_SPEC = 0,             /* |STFT|                              (F, T)  :378-382 */
    HPSS_FEAT_LOGSPEC = 1,          /* power_to_db(|STFT|^2)               (F, T)  :384-389 */
    HPSS_FEAT_MELSPEC = 2,          /* mel_sr(|STFT|^2)                    (M, T)  :391-395 */
    HPSS_FEAT_LOGMELSPEC = 3,       /* power_to_db(mel_sr(|STFT|^2)^2)     (M, T)  :397-402 */
    HPSS_FEAT_HARMPERC = 4,         /* [H; P]                              (2F, T) :426-434 */
    HPSS_FEAT_LOG_HARMPERC = 5,     /* [power_to_db(H^2); power_to_db(P^2)](2F, T) :436-444 */
    HPSS_FEAT_MEL_HARMPERC = 6,     /* [mel(H); mel(P)]                    (2M, T) :404-412 */
    HPSS_FEAT_LOGMEL_HARMPERC = 7   /* [power_to_db(mel(H)^2); ...(P)]     (2M, T) :414-424 */
};

typedef struct hpss_params {
    int32_t n_fft;        /* FFT size (even; n_fft/2 = 2^a 3^b 5^c)                          */
    int32_t win_length;   /* int(Tw*fs/1000); periodic Hann, centre-padded to n_fft          */
    int32_t hop_length;   /* int(Ts*fs/1000)                                                 */
    int32_t l_harm;       /* harmonic median length  (time axis),  PARAMS['l_harm'][Model]   */
    int32_t l_perc;       /* percussive median length (freq axis), PARAMS['l_perc'][Model]   */
    int32_t n_mels;       /* mel bands for the MEL* features                                 */
    int32_t mel_sr;       /* sample rate of the mel basis: 22050 on the HPSS branches (the
                             reference omits sr= there), fs on MELSPEC / LOGMELSPEC          */
    int32_t feature;      /* enum hpss_feature                                               */
    float   amin;         /* power_to_db amin  (1e-10)                                       */
    float   top_db;       /* power_to_db top_db (80); < 0 disables the clip                  */
} hpss_params;

typedef struct hpss_ctx hpss_ctx;       /* per-device context: tables + workspace            */
typedef struct hpss_batch hpss_batch;   /* clip layout of one batch                          */
typedef struct hpss_pipeline hpss_pipeline;   /* host-buffer pipeline: chunking + device slots */

/* ---- library ------------------------------------------------------------------------ */
HPSS_API const char* hpss_version(void);
HPSS_API const char* hpss_last_error(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches). */
HPSS_API uint64_t hpss_launch_count(void);

HPSS_API int hpss_ctx_create(int device, hpss_ctx** ctx);
HPSS_API int hpss_ctx_destroy(hpss_ctx* ctx);
HPSS_API int hpss_ctx_device(const hpss_ctx* ctx);
/* bytes of device workspace currently held by the context */
HPSS_API uint64_t hpss_ctx_workspace_bytes(const hpss_ctx* ctx);
/* Synchronises `stream` and returns HPSS_ERR_NONFINITE / HPSS_ERR_NEGATIVE if a kernel flagged such input since
 * the last check (the status word is cleared), HPSS_OK otherwise.  The context owns one scratch workspace:
 * entry points called from different streams or threads are serialised on the device through an event, so
 * concurrent callers are safe but do not overlap. */
HPSS_API int hpss_ctx_check(hpss_ctx* ctx, void* stream);
/* librosa.util.valid_audio (stft raises "Audio buffer is not finite everywhere") and the non-negativity check of
 * librosa.util.softmask as explicit passes for callers that hand device buffers of unknown content to the stage
 * entry points; the outcome is reported by hpss_ctx_check. */
HPSS_API int hpss_validate_audio(hpss_ctx* ctx, const float* wave_dev, int64_t n, void* stream);
HPSS_API int hpss_validate_nonneg(hpss_ctx* ctx, const float* x_dev, int64_t n, void* stream);

/* pinned host memory for hpss_featuregram_host callers */
HPSS_API int hpss_host_alloc(void** ptr, uint64_t bytes);
HPSS_API int hpss_host_free(void* ptr);

/* ---- batch layout --------------------------------------------------------------------
 * Either from waveform lengths (T_c = 1 + (L_c - n_fft) / hop, librosa.util.frame with
 * center=False; fails with HPSS_ERR_SHORT_SIGNAL if any L_c < n_fft like librosa.stft),
 * or directly from per-clip frame counts when the caller already owns spectrograms
 * (DAFx12_Speech_Music_Detection_B3_MTL_v2.py:230-246 passes a precomputed Spec). */
HPSS_API int hpss_batch_from_samples(hpss_ctx* ctx, const int64_t* clip_len_host, int32_t n_clips,
                                     int32_t n_fft, int32_t hop_length, hpss_batch** batch);
HPSS_API int hpss_batch_from_frames(hpss_ctx* ctx, const int64_t* clip_frames_host, int32_t n_clips,
                                    hpss_batch** batch);
HPSS_API int hpss_batch_destroy(hpss_batch* batch);
HPSS_API int32_t hpss_batch_n_clips(const hpss_batch* batch);
HPSS_API int64_t hpss_batch_total_frames(const hpss_batch* batch);
HPSS_API int64_t hpss_batch_total_samples(const hpss_batch* batch);
/* copies n_clips+1 exclusive prefix sums */
HPSS_API int hpss_batch_frame_offsets(const hpss_batch* batch, int64_t* out_host);
HPSS_API int hpss_batch_sample_offsets(const hpss_batch* batch, int64_t* out_host);

/* ---- tables (host side, double precision internally) ---------------------------------
 * librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm='slaney')
 * -> float32 (n_mels, 1+n_fft/2), as melspectrogram builds it (lib/preprocessing.py:394,
 * 400, 409-410, 419, 421). */
HPSS_API int hpss_mel_filterbank(int32_t sr, int32_t n_fft, int32_t n_mels, float* out_host);
/* scipy.signal.get_window('hann', win_length, fftbins=True) centre-padded to n_fft
 * (librosa.stft window handling); float32 rounding of the float64 window. */
HPSS_API int hpss_stft_window(int32_t n_fft, int32_t win_length, float* out_host);

/* ---- K1: librosa.core.stft(center=False) + np.abs  (lib/preprocessing.py:381,387,407,
 * 417,429,439).  S_dev: (F, T_c) per clip, F = n_fft/2+1.  cplx_dev (optional, may be NULL):
 * interleaved complex64 with the same layout.  power != 0 writes |X|^2 instead of |X|. */
HPSS_API int hpss_stft_mag(hpss_ctx* ctx, const hpss_batch* batch, const float* wave_dev,
                           int32_t n_fft, int32_t win_length, int32_t hop_length, int32_t power,
                           float* S_dev, float* cplx_dev, void* stream);

/* ---- K2: scipy.ndimage.median_filter(S, size=(1,k) | (k,1), mode='reflect') as called
 * by librosa.decompose.hpss (lib/preprocessing.py:408,418,430,440).  Bit-exact selection:
 * window offsets -k/2 .. k-1-k/2, half-sample-symmetric reflection (any overshoot),
 * element of rank k/2.  rows = F of the (rows, T_c) matrices. */
HPSS_API int hpss_median_time(hpss_ctx* ctx, const hpss_batch* batch, const float* S_dev,
                              int32_t rows, int32_t k, float* out_dev, void* stream);
HPSS_API int hpss_median_freq(hpss_ctx* ctx, const hpss_batch* batch, const float* S_dev,
                              int32_t rows, int32_t k, float* out_dev, void* stream);

/* ---- K3: librosa.util.softmask(power=2, split_zeros=True) x2, S*mask, then per stream
 * np.dot(mel, .) and power_to_db(.**2) without the top_db clip (lib/preprocessing.py:
 * 408-412, 418-424, 430-434, 440-444).
 *   mel_dev == NULL  -> identity projection, out rows = 2*rows  (HARMPERC / LOG_HARMPERC)
 *   mel_dev != NULL  -> dense float32 (n_mels, rows) basis,  out rows = 2*n_mels
 *   log_power == 1   -> 10*log10(max(amin, x*x))  (the reference always calls power_to_db(x**2));
 *   log_power == 2   -> 10*log10(max(amin, x));    either way clip_max_dev[2*c+s] (ordered-uint
 *                       encoding, zero-initialised by this call) receives the per-clip,
 *                       per-stream maximum needed by top_db.
 * harm_dev/perc_dev may be NULL together: plain single-stream mode on S (SPEC family),
 * with pre_square != 0 squaring S first (MELSPEC uses the power spectrogram). */
HPSS_API int hpss_mask_mel_log(hpss_ctx* ctx, const hpss_batch* batch, const float* S_dev,
                               const float* harm_dev, const float* perc_dev, int32_t rows,
                               const float* mel_dev, int32_t n_mels, int32_t pre_square,
                               int32_t log_power, float amin, float* out_dev,
                               uint32_t* clip_max_dev, void* stream);

/* Same with the basis librosa.feature.melspectrogram itself would build: the cached Slaney filterbank for
 * (mel_sr, n_fft = 2*(rows-1), n_mels) (lib/preprocessing.py:409-410, 419, 421 -> sr = 22050; :394, :400 ->
 * sr = 16000).  Lets the library use its banded single-sweep kernel; bit-identical to hpss_mask_mel_log
 * called with hpss_mel_filterbank(mel_sr, 2*(rows-1), n_mels). */
HPSS_API int hpss_mask_mel_log_sr(hpss_ctx* ctx, const hpss_batch* batch, const float* S_dev,
                                  const float* harm_dev, const float* perc_dev, int32_t rows,
                                  int32_t mel_sr, int32_t n_mels, int32_t pre_square,
                                  int32_t log_power, float amin, float* out_dev,
                                  uint32_t* clip_max_dev, void* stream);

/* ---- K3b: the top_db part of librosa.core.power_to_db: x = max(x, max_clip_stream - top_db)
 * (lib/preprocessing.py:388,401,420,422,441-442).  out rows = n_streams * rows_per_stream. */
HPSS_API int hpss_topdb_clip(hpss_ctx* ctx, const hpss_batch* batch, float* out_dev,
                             int32_t rows_per_stream, int32_t n_streams,
                             const uint32_t* clip_max_dev, float top_db, void* stream);

/* ---- fused convenience entry: body of get_featuregram after the signal is loaded
 * (lib/preprocessing.py:378-444).  out_dev rows = hpss_feature_rows(params). */
HPSS_API int32_t hpss_feature_rows(const hpss_params* params);
HPSS_API int hpss_featuregram(hpss_ctx* ctx, const hpss_batch* batch, const float* wave_dev,
                              const hpss_params* params, float* out_dev, void* stream);
/* same from a precomputed magnitude spectrogram (DAFx12 ...v2.py:230-246) */
HPSS_API int hpss_featuregram_from_spec(hpss_ctx* ctx, const hpss_batch* batch, const float* S_dev,
                                        int32_t rows, const hpss_params* params, float* out_dev,
                                        void* stream);
/* Host-buffer entry (what a caller of the reference's numpy API binds): waveform in
 * host memory -> features in host memory; H2D, kernels and D2H are pipelined over clip
 * chunks on the context's own streams; returns after the result is complete.
 * Buffers from hpss_host_alloc (pinned) give full PCIe rate. */
HPSS_API int hpss_featuregram_host(hpss_ctx* ctx, const hpss_batch* batch, const float* wave_host,
                                   const hpss_params* params, float* out_host);

/* Host-buffer pipeline as an object (what hpss_featuregram_host runs internally), for corpus passes: clips are cut
 * into ~n_chunks_hint chunks (0 = default 16; never splitting a clip); chunk i+1 uploads while chunk i computes and
 * chunk i-1 downloads.
 *   prepare == 0: pcm_host is the prepared float32 waveform (pcm_format must be HPSS_PCM_F32).
 *   prepare != 0: pcm_host is the decoded file (float32 or int16 PCM); load_and_preprocess_signal's normalise /
 *                 silence removal / normalise (lib/preprocessing.py:332-348, tools.pyx:42-134) runs on the device
 *                 with frame_length = params->win_length, hop = params->hop_length, sampling rate fs.
 * hpss_pipeline_run: feat_host (may be NULL: features are not downloaded) receives hpss_feature_rows(params) x
 * total_frames floats in the batch layout; when clip_class_host / moments_host are given, the raw moments of
 * get_data_stats ([n_classes*D sums | D sums of squares | n_classes frame counts | non-finite count], float64) are
 * ADDED to moments_host -- the corpus pass of get_data_stats then moves 8 KB back instead of the features.
 * Returns HPSS_ERR_NONFINITE for non-finite audio (librosa.util.valid_audio). */
HPSS_API int hpss_pipeline_create(hpss_ctx* ctx, const int64_t* clip_len_host, int32_t n_clips,
                                  const hpss_params* params, int32_t pcm_format, int32_t prepare, int32_t fs,
                                  double alpha, double beta, int32_t n_chunks_hint, hpss_pipeline** pipeline);
HPSS_API int hpss_pipeline_destroy(hpss_pipeline* pipeline);
HPSS_API int64_t hpss_pipeline_total_frames(const hpss_pipeline* pipeline);
HPSS_API int32_t hpss_pipeline_n_chunks(const hpss_pipeline* pipeline);
HPSS_API int hpss_pipeline_frame_offsets(const hpss_pipeline* pipeline, int64_t* out_host);
HPSS_API int hpss_pipeline_run(hpss_pipeline* pipeline, const void* pcm_host, float* feat_host,
                               const int32_t* clip_class_host, int32_t n_classes, double* moments_host);

/* ---- N2: signal preparation = load_and_preprocess_signal after the decode (lib/preprocessing.py:332-348):
 * normalize_signal (:114-132), librosa.feature.rms(frame_length=win_length, hop_length, center=True) (:337),
 * removeSilence (lib/cython_impl/tools.pyx:42-134: float32 threshold alpha*max, 5-tap median of the markers,
 * stretches longer than beta seconds, nothing removed unless MORE than one stretch qualifies, kept samples packed
 * to the front of a buffer of ones that keeps the original length), the doubling of clips shorter than 0.1 s
 * (:345-347) and the second normalize_signal (:348), for n_clips files per call.
 *   pcm_dev: the files back to back, float32 or int16 (hpss_pcm_format); out_dev: the prepared float32 signals back
 *   to back, clip c has hpss_prep_out_length(len_c, fs) samples (== len_c unless len_c / fs < 0.1).
 *   Optional outputs (NULL to skip): frame_marker_dev int32 per RMS frame (clip c owns hpss_prep_num_frames(len_c)
 *   entries, clips back to back) = the reference's frame_silMarker; sample_marker_dev uint8 per input sample = its
 *   sample_silMarker; n_sil_dev int32 per clip = stretches that qualified.
 * Non-finite samples set the NONFINITE bit (hpss_ctx_check). */
HPSS_API int64_t hpss_prep_out_length(int64_t n_samples, int32_t fs);
HPSS_API int64_t hpss_prep_num_frames(int64_t n_samples, int32_t win_length, int32_t hop_length);
HPSS_API int hpss_prep_signals(hpss_ctx* ctx, const void* pcm_dev, int32_t pcm_format, const int64_t* clip_len_host,
                               int32_t n_clips, int32_t fs, int32_t win_length, int32_t hop_length, double alpha,
                               double beta, float* out_dev, int32_t* frame_marker_dev, uint8_t* sample_marker_dev,
                               int32_t* n_sil_dev, void* stream);
/* mix_signals (lib/preprocessing.py:297-325) for n_pairs (speech, music) pairs: the music is looped to the speech
 * length, scaled to the target speech-to-music ratio target_db_host[p] (dB), both weights divided by their sum, the
 * mix normalised (normalize_signal).  sp_dev / mu_dev: prepared signals back to back with the given lengths; out_dev:
 * the mixes back to back, pair p has sp_len_host[p] samples.  float64 arithmetic, one rounding to float32. */
HPSS_API int hpss_mix_signals(hpss_ctx* ctx, const float* sp_dev, const int64_t* sp_len_host, const float* mu_dev,
                              const int64_t* mu_len_host, const double* target_db_host, int32_t n_pairs,
                              float* out_dev, void* stream);

/* ---- K5: raw moments for get_data_stats (lib/preprocessing.py:461-586).
 * feat_dev: (D, T_c) per clip.  clip_class_host[c] in [0, n_classes), n_clips entries (must equal the batch's clip
 * count; the classes stay on the device and are uploaded again only when they change).  Accumulates, in
 * float64, sum_dev[class*D + d] += sum_t x, sumsq_dev[d] += sum_t x^2 (all classes),
 * count_dev[class] += T_c, nonfinite_dev[0] += number of non-finite values.  The caller
 * zeroes the accumulators once, calls this per batch, all-reduces them across ranks (the
 * only collective on the path) and finishes with hpss_stats_finalize. */
HPSS_API int hpss_moments(hpss_ctx* ctx, const hpss_batch* batch, const float* feat_dev, int32_t D,
                          const int32_t* clip_class_host, int32_t n_clips, int32_t n_classes, double* sum_dev,
                          double* sumsq_dev, double* count_dev, double* nonfinite_dev, void* stream);
/* K3b + K5 in one pass over the features: hpss_topdb_clip followed by hpss_moments of the clipped values. */
HPSS_API int hpss_topdb_moments(hpss_ctx* ctx, const hpss_batch* batch, float* out_dev, int32_t rows_per_stream,
                                int32_t n_streams, const uint32_t* clip_max_dev, float top_db,
                                const int32_t* clip_class_host, int32_t n_clips, int32_t n_classes, double* sum_dev,
                                double* sumsq_dev, double* count_dev, double* nonfinite_dev, void* stream);
/* hpss_featuregram followed by hpss_moments in one call; when the feature has a top_db clip the
 * clip and the moment accumulation share a single pass over the features (K3b + K5 fused). */
HPSS_API int hpss_featuregram_moments(hpss_ctx* ctx, const hpss_batch* batch, const float* wave_dev,
                                      const hpss_params* params, float* out_dev, const int32_t* clip_class_host,
                                      int32_t n_clips, int32_t n_classes, double* sum_dev, double* sumsq_dev,
                                      double* count_dev, double* nonfinite_dev, void* stream);
/* class means -> unweighted mean of class means (:530-536); stdev = sqrt(sum (x-mean)^2 /
 * (N-1)) (:575-583) from the raw moments, float64 -> float32 (:586). Host arrays. */
HPSS_API int hpss_stats_finalize(const double* sum_host, const double* sumsq_host,
                                 const double* count_host, int32_t D, int32_t n_classes,
                                 float* mean_out, float* stdev_out);

/* ---- scale_data: (x - mean) / (stdev + eps) -> float64 (lib/cython_impl/tools.pyx:138-166
 * uses eps = 1e-10; lib/preprocessing.py:590-614 uses eps = 0). */
HPSS_API int hpss_scale_data(hpss_ctx* ctx, const hpss_batch* batch, const float* feat_dev, int32_t D,
                             const float* mean_dev, const float* stdev_dev, double eps,
                             double* out_dev, void* stream);
/* The pure-Python scale_data (lib/preprocessing.py:590-614) on float32 input evaluates (FV - mean) / stdev in float32:
 * two rounded operations, float32 result -- bit-identical to numpy. */
HPSS_API int hpss_scale_data_f32(hpss_ctx* ctx, const hpss_batch* batch, const float* feat_dev, int32_t D,
                                 const float* mean_dev, const float* stdev_dev, float* out_dev, void* stream);

/* ---- N1: get_feature_patches (lib/preprocessing.py:137-292 + tools.pyx:21-38) for one
 * clip-batch: optional per-clip per-row standardisation (sklearn StandardScaler: mean,
 * std ddof=0 in float64, constant rows -> scale 1, float32 update) followed by the patch gather.
 *   hpss_row_standardize: in place on feat_dev (D, T_c) per clip.
 *   hpss_extract_patches: clip `clip` only; patch p covers frames
 *       [p*shift, p*shift + patch_size), p < n_patches = len(range(W/2, T - W/2, shift));
 *       out float64 (n_patches, D, patch_size). */
HPSS_API int hpss_row_standardize(hpss_ctx* ctx, const hpss_batch* batch, float* feat_dev, int32_t D,
                                  void* stream);
HPSS_API int64_t hpss_num_patches(int64_t n_frames, int32_t patch_size, int32_t patch_shift);
HPSS_API int hpss_extract_patches(hpss_ctx* ctx, const float* feat_dev, int32_t D, int64_t n_frames,
                                  int32_t patch_size, int32_t patch_shift, double* out_dev,
                                  void* stream);

/* N1, device resident: the model-ready patch tensor of a whole batch of featuregrams in one call -- what
 * get_feature_patches (lib/preprocessing.py:137-292) followed by the generator's batch assembly builds on the host.
 *   feat_dev   (D, T_c) float32 per clip in the batch layout (hpss_featuregram output); standardize != 0 first applies
 *              hpss_row_standardize IN PLACE (the reference's StandardScaler(copy=False) also mutates its input).
 *   rows       [row0, row0 + n_rows): all rows for *HarmPercSpec / the plain names (the per-stream scalers act per
 *              row, so standardising both halves at once equals :211-224), [0, D/2) for *HarmSpec, [D/2, D) for *PercSpec.
 *   patches    clip c contributes hpss_num_patches_tiled(T_c, W, shift) patches (clips shorter than W are tiled
 *              along time like :139-142); hpss_patch_offsets returns the n_clips+1 prefix sums.
 *   out_dev    time_major == 0: (n_patches, n_rows, W)  (CNN input, lib/proposed_architectures.py:451);
 *              time_major != 0: (n_patches, W, n_rows)  (the transpose the TCNs take, Proposed_Work_Results.py:235-236);
 *              float32 (out_f64 == 0, what the model consumes) or float64 (the reference's dtype, tools.pyx:27). */
HPSS_API int64_t hpss_num_patches_tiled(int64_t n_frames, int32_t patch_size, int32_t patch_shift);
HPSS_API int hpss_patch_offsets(const hpss_batch* batch, int32_t patch_size, int32_t patch_shift, int64_t* out_host);
HPSS_API int hpss_patch_tensor(hpss_ctx* ctx, const hpss_batch* batch, float* feat_dev, int32_t D, int32_t standardize,
                               int32_t row0, int32_t n_rows, int32_t patch_size, int32_t patch_shift,
                               int32_t time_major, int32_t out_f64, void* out_dev, void* stream);
/* The same gather for float64 featuregrams -- with frame_level_scaling the generators hand get_feature_patches the
 * float64 output of the Cython scale_data (Proposed_Work_Results.py:94-95) and no standardisation happens: exact
 * float64 copies. */
HPSS_API int hpss_patch_tensor_f64(hpss_ctx* ctx, const hpss_batch* batch, const double* feat_dev, int32_t D,
                                   int32_t row0, int32_t n_rows, int32_t patch_size, int32_t patch_shift,
                                   int32_t time_major, double* out_dev, void* stream);

/* get_data_stats drops, per file, the feature rows that hold a NaN or Inf (lib/preprocessing.py:507-508):
 * flags_dev[c * D + d] = 1 when row d of clip c has a non-finite value.  Only needed when hpss_moments reported
 * non-finite values. */
HPSS_API int hpss_row_nonfinite(hpss_ctx* ctx, const hpss_batch* batch, const float* feat_dev, int32_t D,
                                uint8_t* flags_dev, void* stream);

/* N4: get_data_statistics (lib/cython_impl/tools.pyx:169-211): per-patch statistic vectors of a (n_patches, n_feat,
 * n_frames) float64 patch array.  stat: 0 mean, 1 variance (ddof 0), 2 scipy.stats.skew, 3 scipy.stats.kurtosis
 * (biased, Fisher).  axis 0 reduces over the features -> (n_patches, n_frames) ("percussive"); axis 1 over the
 * frames -> (n_patches, n_feat) ("harmonic").  Constant vectors give NaN for skew / kurtosis as scipy >= 1.9 does. */
HPSS_API int hpss_patch_statistics(hpss_ctx* ctx, const double* patches_dev, int64_t n_patches, int32_t n_feat,
                                   int32_t n_frames, int32_t stat, int32_t axis, double* out_dev, void* stream);

/* ---- extension: MFCC.  The reference has no MFCC / DCT anywhere; BASELINE.json's north_star
 * names it, so it is offered as what librosa.feature.mfcc would add after the reference's log-mel step
 * (lib/preprocessing.py:119-120): scipy.fftpack.dct(S_db, axis=0, type=2, norm='ortho')[:n_mfcc] per stream.
 * Parity is pinned to scipy's DCT, not to the reference.
 *   feat_dev: (n_streams * rows_per_stream, T_c) float32 per clip in the batch layout (hpss_featuregram output);
 *   out_dev : (n_streams * n_mfcc, T_c) float32 per clip, same layout; must not alias feat_dev;
 *   1 <= n_mfcc <= min(rows_per_stream, 64).
 *   hpss_dct_basis: the (n_mfcc, n_mels) float32 basis itself (host memory), for inspection / tests. */
HPSS_API int hpss_dct_mfcc(hpss_ctx* ctx, const hpss_batch* batch, const float* feat_dev,
                           int32_t rows_per_stream, int32_t n_streams, int32_t n_mfcc, float* out_dev,
                           void* stream);
HPSS_API int hpss_dct_basis(int32_t n_mels, int32_t n_mfcc, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* HPSS_B200_H */
