/* logmel.h -- C ABI of the B200-native log-mel feature frontend (liblogmel_b200.so).
 *
 * This is the drop-in boundary for the one hot path of AdamBeedell/MLX8-WS-Audio-Transformer:
 * waveform -> log-mel spectrogram.  Each entry point names the reference interface it stands
 * in for.  The arithmetic the reference reaches lives in third-party Python libraries, so the
 * "FFI" a maintainer binds is a ctypes stub (INTEGRATION.md shows it):
 *
 *   lm_create / lm_forward  with LM_LOG10_CLAMP_WHISPER_NORM
 *       = WhisperFeatureExtractor._torch_extract_fbank_features
 *         (transformers/models/whisper/feature_extraction_whisper.py:135-164; NumPy twin :105-133),
 *         reached from /root/reference/AB/fineTune.py:88, AB/fineTuneMidi.py:88,
 *         AB/wavToWhisper.py:55, AB/fineTuneMidiTester.py:33, .charles/music2midi/model.py:100-104.
 *   lm_create / lm_forward  with LM_LOG_NONE or LM_LN_PLUS_EPS
 *       = torchaudio.transforms.MelSpectrogram.forward (+ torch.log(mel + 1e-6))
 *         (torchaudio/transforms/_transforms.py:621-631, 407-419;
 *          torchaudio/functional/functional.py:123-144),
 *         reached from /root/reference/.charles/spectrogram.py:79-87, 161-162, 299-300, 306-307.
 *   the zero padding / truncation of each clip to n_samples
 *       = SequenceFeatureExtractor.pad (transformers/feature_extraction_sequence_utils.py:276-277)
 *         and /root/reference/.charles/spectrogram.py:152-157, expressed here by `d_lengths`.
 *
 * Conventions
 *   - Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *   - Every function returns 0 on success, a negative lm_status for a rejected argument
 *     (checked before anything is launched) or a positive cudaError_t.  lm_last_error()
 *     returns a thread-local message for the last failure.
 *   - Device entry points are asynchronous on the caller's stream and never allocate;
 *     the caller owns all buffers.  A handle is immutable after lm_create and may be shared
 *     by host threads; it belongs to one device.
 *   - There is no CPU fallback: without a CUDA device lm_create fails.
 */
#ifndef LOGMEL_B200_H
#define LOGMEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_ABI_VERSION 1

typedef struct lm_handle lm_handle;

typedef enum lm_log_mode {
  LM_LOG_NONE = 0,                 /* mel power, as MelSpectrogram.forward                       */
  LM_LOG10_CLAMP_WHISPER_NORM = 1, /* log10(max(M, floor)); max(S, clipmax - 8); (S + 4) / 4     */
  LM_LN_PLUS_EPS = 2,              /* ln(M + eps), as spectrogram.py:162                          */
  LM_LOG10_CLAMP = 3               /* log10(max(M, floor)) without the per-clip normalisation     */
} lm_log_mode;

typedef enum lm_status {
  LM_OK = 0,
  LM_ERR_NULL = -1,          /* a required pointer is NULL                                        */
  LM_ERR_GEOMETRY = -2,      /* (n_fft, hop) pair has no kernel: supported are (400,160),
                                (1024,512), (1024,128)                                            */
  LM_ERR_FBANK = -3,         /* n_mels out of range or filter bank not banded                     */
  LM_ERR_SHAPE = -4,         /* batch / n_samples / stride out of range (n_samples <= n_fft/2
                                cannot be reflect padded, exactly as torch.stft refuses it)       */
  LM_ERR_SCRATCH = -5,       /* scratch buffer too small (see lm_scratch_bytes)                   */
  LM_ERR_NO_DEVICE = -6,     /* no CUDA device / wrong architecture                               */
  LM_ERR_MODE = -7           /* unknown lm_log_mode                                               */
} lm_status;

typedef struct lm_config {
  int32_t n_fft;             /* 400 (Whisper) or 1024 (.charles/spectrogram.py N_FFT)             */
  int32_t hop;               /* 160, or 512 / 128                                                 */
  int32_t n_mels;            /* 1..128                                                            */
  int32_t log_mode;          /* lm_log_mode                                                       */
  float log_param;           /* floor (1e-10) for the LOG10 modes, eps (1e-6) for LN_PLUS_EPS     */
  int32_t drop_last;         /* 1: emit n_samples/hop frames (Whisper drops frame 3000),
                                0: emit 1 + n_samples/hop frames (torchaudio)                     */
  int32_t device;            /* CUDA device ordinal                                               */
  int32_t variant;           /* 0 = default: for the Whisper normalisation with the 80- / 128-filter
                                Slaney bank, 16-byte-aligned clips of >= 8 tiles (2.5 s) run the thread-per-frame
                                kernel (tensor memory as transpose scratch) at every batch size, everything else the
                                CTA-tiled kernels (n_fft 400: warp-specialised CTA, two frames per lane);
                                1 = CTA-tiled, one frame per lane (scalar FP32 path), 2 = CTA-tiled,
                                two frames per lane, phase-synchronous CTA, 3 = thread-per-frame
                                kernel for every batch size -- tuning / cross-check knobs.
                                The kernels agree to float32 rounding (<= 1e-6), not bit for bit.  */
  const float* fbank;        /* host, [n_fft/2+1][n_mels] row-major float32 (mel_filters cast to
                                f32 / MelScale.fb); copied by lm_create                           */
  const float* window;       /* host, [n_fft] float32, or NULL for the periodic Hann window      */
} lm_config;

int lm_version(void);
const char* lm_last_error(void);

int lm_create(lm_handle** out, const lm_config* cfg);
void lm_destroy(lm_handle* h);

/* frames produced for clips padded/truncated to n_samples */
int64_t lm_num_frames(const lm_handle* h, int64_t n_samples);
/* bytes of device scratch lm_forward needs for `batch` clips */
size_t lm_scratch_bytes(const lm_handle* h, int64_t batch);

/* Device-resident forward.
 *   d_wave      [batch] clips, clip i starts at d_wave + i * clip_stride (floats), float32
 *   n_samples   L: every clip is treated as right zero-padded / truncated to L samples
 *   d_lengths   NULL, or int32 [batch]: samples actually present in clip i (<= clip_stride);
 *               samples at or past min(d_lengths[i], L) read as zero
 *   d_out       float32 [batch][n_mels][lm_num_frames(L)], contiguous
 *   d_clip_max  NULL, or float32 [batch]: receives each clip's max log10 (WHISPER_NORM only)
 *   d_scratch   lm_scratch_bytes(batch) bytes, 16-byte aligned
 *   stream      cudaStream_t (as void*), 0 for the default stream
 */
int lm_forward(lm_handle* h, const float* d_wave, int64_t batch, int64_t clip_stride,
               int64_t n_samples, const int32_t* d_lengths, float* d_out, float* d_clip_max,
               void* d_scratch, size_t scratch_bytes, void* stream);

/* Fused 16-bit PCM ingest: the same operator for clips that are still what a WAV file
 * holds -- int16 frames of `channels` (1 or 2) interleaved channels -- so that the s16 -> float32
 * conversion of torchaudio.load (/root/reference/AB/wavToWhisper.py:52, AB/memoToWav.py:19 writes s16
 * mono) and the stereo -> mono mean of /root/reference/.charles/spectrogram.py:147-148 happen in the
 * kernel's tile loader: sample = sum_c pcm[frame][c] / (32768 * channels), bit-identical to the float
 * path on the converted waveform.  clip_stride, n_samples and d_lengths count FRAMES.  Halves (mono) the
 * bytes read from HBM / sent over PCIe.  Every clip should start 16-byte aligned
 * (clip_stride * channels % 8 == 0) for the vectorised loader; other layouts are read sample by sample.
 */
int lm_forward_pcm16(lm_handle* h, const int16_t* d_pcm, int32_t channels, int64_t batch,
                     int64_t clip_stride, int64_t n_samples, const int32_t* d_lengths, float* d_out,
                     float* d_clip_max, void* d_scratch, size_t scratch_bytes, void* stream);
int lm_forward_host_pcm16(lm_handle* h, const int16_t* h_pcm, int32_t channels, int64_t batch,
                          int64_t clip_stride, int64_t n_samples, const int32_t* h_lengths, float* h_out);

/* Host-buffer forward: the same operator for HOST waveforms and HOST features.  The batch is
 * cut into chunks that are copied to the device, transformed and copied back on three
 * streams so that H2D, compute and D2H overlap; device staging buffers belong to the handle
 * (calls on one handle are serialised).  h_wave / h_out should be pinned for full PCIe speed
 * (lm_host_register pins a caller buffer in place).  Returns after h_out is complete.
 */
int lm_forward_host(lm_handle* h, const float* h_wave, int64_t batch, int64_t clip_stride,
                    int64_t n_samples, const int32_t* h_lengths, float* h_out);

int lm_host_register(void* p, size_t bytes);
int lm_host_unregister(void* p);

/* Introspection used by bench.py / tests: number of kernels launched by this library in the
 * calling process; SM count, resident CTAs, dynamic shared memory, threads and frames per tile of the
 * handle's CTA-tiled kernel (the thread-per-frame kernel, when lm_kernel_name says a launch takes it, is
 * always one 256-thread CTA per SM with 32-frame tiles). */
int64_t lm_launch_count(void);
int lm_kernel_info(const lm_handle* h, int32_t* n_sm, int32_t* ctas_per_sm, int32_t* smem_bytes,
                   int32_t* threads, int32_t* frames_per_tile);
/* name (as ncu / cuobjdump print it) of the kernel lm_forward launches for `batch` 16-byte aligned
 * clips of n_samples; the string belongs to the handle */
const char* lm_kernel_name(const lm_handle* h, int64_t batch, int64_t n_samples);

#ifdef __cplusplus
}
#endif
#endif /* LOGMEL_B200_H */
