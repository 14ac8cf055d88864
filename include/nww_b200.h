/* nww_b200.h — C ABI of libnwwb200.so, the B200 (sm_100a) engine for the nanowakeword
 * per-window hot path:  int16 PCM -> log-mel -> classifier head -> sigmoid score.
 *
 * This is the drop-in boundary.  The reference (arcosoph/nanowakeword v3.0.0) has no native
 * code; its only seam is the onnxruntime.InferenceSession duck type that NanoInterpreter
 * drives (reference nanowakeword/interpreter/nanointerpreter.py:165-167, 677-682, 783,
 * 955-959) and that _RemoteSession already replaces
 * (nanowakeword/interpreter/remote_verifier.py:490-648).  Each entry point below cites the
 * reference call it stands in for; INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions: every function returns 0 on success or a negative NWW_E* code and records a
 * message retrievable with nww_last_error() (thread-local).  Pointers named *_dev are CUDA
 * device pointers on the engine's device, *_host are host pointers (pinned memory makes the
 * copies asynchronous).  The engine owns weights, tables and workspaces; the caller owns all
 * I/O buffers.  Calls on one engine are serialised by an internal mutex AND ordered on the
 * device: the engine's workspaces and stream rings are shared between calls, so every entry
 * point makes its CUDA stream wait for the work the previous call enqueued (an event chain),
 * whichever stream that was — calls on different streams behave like calls on one stream.
 * There is no CPU fallback: without a CUDA device nww_create fails.
 */
#ifndef NWW_B200_H
#define NWW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NWW_OK 0
#define NWW_EINVAL (-1)      /* bad argument / malformed spec or weight blob            */
#define NWW_ECUDA (-2)       /* a CUDA runtime call failed (message has the CUDA error)  */
#define NWW_EUNSUPPORTED (-3) /* architecture / geometry not built into this library       */

/* classifier backbones (reference nanowakeword/modules/architectures.py) */
#define NWW_ARCH_DNN 0       /* Net + FCNBlock            :102-126 */
#define NWW_ARCH_CNN 1       /* CNNModel                  :51-80   */
#define NWW_ARCH_TCN 2       /* TCNModel / TemporalBlock  :290-362 */
#define NWW_ARCH_BCRESNET 3  /* BcResNetModel             :620-687 */
#define NWW_ARCH_CRNN_GRU 4  /* CRNNModel, rnn_type=gru   :209-287 */
#define NWW_ARCH_E2E_MELCNN 5 /* E2E_MelSpectrogram_CNN   :820-888 */
#define NWW_ARCH_GRU 6       /* GRUModel (bidirectional)  :129-146 */
#define NWW_ARCH_LSTM 7      /* LSTMModel :83-99, and RNNModel :149-161 (a bidirectional LSTM, 64 hidden units) */
#define NWW_ARCH_QUARTZNET 8 /* QuartzNetModel / QuartzNetBlock :366-437 */
#define NWW_ARCH_E2E_QUARTZNET 9 /* E2ERawQuartzNet :796-817 = RawAudioFrontend :695-714 on the audio itself + QuartzNetModel;
                                    no log-mel: the spec's geometry is ignored and mel dumps are refused */
#define NWW_ARCH_E2E_CNN 10  /* E2ERawCNN :777-793 = RawAudioFrontend + RawAudioBackbone :738-774 (raw audio, no log-mel) */

#define NWW_ACT_RELU 0       /* reference nanowakeword/modules/model.py:81-87 */
#define NWW_ACT_GELU 1
#define NWW_ACT_SILU 2

#define NWW_GEOM_NS40X98 0   /* frame 400 -> n_fft 512, hop 160, 40 mels, not centred -> (40, 98)  */
#define NWW_GEOM_REF64X101 1 /* n_fft = win = 400, hop 160, 64 mels, centred reflect -> (64, 101)  */

#define NWW_FRONTEND_FP64 0  /* FFT / power / mel accumulation in double (default, exact)  */
#define NWW_FRONTEND_FP32 1  /* experimental fast mode                                      */

typedef struct nww_engine nww_engine;

typedef struct nww_spec {
    uint32_t struct_size;        /* sizeof(nww_spec), for forward compatibility */
    int32_t arch;                /* NWW_ARCH_*  */
    int32_t activation;          /* NWW_ACT_*   */
    int32_t geometry;            /* NWW_GEOM_*  */
    int32_t n_fft, win_length, hop_length, n_mels, center, clip_samples; /* checked against geometry */
    int32_t frontend_precision;  /* NWW_FRONTEND_* */
    int32_t chunk_windows;       /* windows per internal launch chunk; 0 = default */
    int32_t reserved[8];
} nww_spec;

typedef struct nww_info_t {
    int32_t device, sm_count;
    int32_t n_mels, n_frames, clip_samples, feature_dim, embedding_dim;
    int32_t chunk_windows;
    int64_t kernel_launches;     /* kernels launched by this engine so far */
    int64_t windows_scored;
} nww_info_t;

/* Build an engine on CUDA device `device` from a spec and a packed weight blob
 * (format: nanowakeword_b200/csrc/nww_blob.h).  Replaces onnxruntime.InferenceSession(path,
 * sess_options, providers=["CPUExecutionProvider"]) — nanointerpreter.py:955-959. */
int nww_create(const nww_spec* spec, const void* weights, size_t weights_size, int device, nww_engine** out);
void nww_destroy(nww_engine* e);
const char* nww_last_error(void);
int nww_get_info(nww_engine* e, nww_info_t* info);

/* Score n_windows windows of clip_samples int16 samples each, resident in device memory
 * (16-byte aligned), writing one float32 probability per window.  `stream` is a
 * cudaStream_t (NULL = the legacy default stream, as everywhere in CUDA); the call returns
 * after enqueueing and the outputs are ordered on that stream.
 * Optional outputs (NULL to skip): mel_dev (n, n_mels, n_frames) log-mel in dB,
 * logits_dev (n), emb_dev (n, embedding_dim).
 * Replaces session.run(None, {"input": clip}) on the exported graph
 * [mel -> dB ->] backbone -> classifier -> sigmoid -> view(-1,1,1)
 * — nanointerpreter.py:783, _export/onnx.py:164-172. */
int nww_run_windows(nww_engine* e, const int16_t* pcm_dev, int64_t n_windows, float* scores_dev, float* mel_dev,
                    float* logits_dev, float* emb_dev, void* stream);

/* Same computation from float32 PCM already scaled by 1/32768 — the tensor the reference
 * feeds its session (nanointerpreter.py:750, 771-775).  The samples are used AS THEY ARE (no
 * rounding to the int16 grid, no clipping at +-1): FP64 front end on the float samples (or the
 * raw-audio layers for the E2ERaw* models), then the same head kernels.  For audio that is on
 * the int16 grid the int16 entry points compute the same thing from half the bytes. */
int nww_run_windows_f32(nww_engine* e, const float* pcm_dev, int64_t n_windows, float* scores_dev, float* mel_dev,
                        float* logits_dev, float* emb_dev, void* stream);

/* End-to-end variant on HOST buffers: chunked host->device copies overlapped with compute
 * and one device->host copy of the scores; returns when scores_host is complete.
 * This is what B200Session.run() (the session duck type) calls. */
int nww_run_windows_host(nww_engine* e, const int16_t* pcm_host, int64_t n_windows, float* scores_host);

/* ---- many independent audio streams ---------------------------------------------------------
 * The reference keeps, per model, a deque(maxlen=clip_samples) of the most recent samples and a
 * cumulative sample counter, and on every predict(chunk) re-scores the last clip_samples once the
 * counter has reached clip_samples (nanointerpreter.py:176-183 create, :750-756 append + window,
 * :785-786 score 0.0 before that, :719-733 reset).  These entry points hold that state for
 * n_streams streams in device memory (one int16 ring per stream) and advance all of them by one
 * chunk per call.  Warm-up zeroing of the first five predictions and the patience / debounce
 * filters (:789-790, :1034-1064) stay on the host side (nanowakeword_b200/streams.py).
 *
 * nww_stream_open    allocate and zero the rings (re-opening discards the previous set).
 * nww_stream_push    chunks_dev is (n_streams, chunk_len) int16 on the device: every stream gets
 *                    chunk_len new samples (any positive length); scores_dev (n_streams) receives the
 *                    probability of each stream's last clip_samples, or 0 while a stream has received
 *                    fewer than clip_samples since it was opened / reset.  Ordered on `stream`.
 * nww_stream_push_host  same with host buffers (H2D of the chunks, D2H of the scores; synchronous).  Large banks are cut
 *                       into up to four pieces so that a piece's copy runs behind the previous piece's kernels.
 * nww_stream_reset   ids_host == NULL resets every stream, else the n_ids listed streams.
 * nww_stream_push_select[_host]  like nww_stream_push[_host], but only the n_ids streams listed in ids (distinct
 *                    indices, any order) are SCORED: every stream still receives its chunk (PCM ring and log-mel ring
 *                    advance), the streams not listed report 0.  This is the cascade's verifier stage
 *                    (nanointerpreter.py:758-769: the verifier is skipped, score 0.0, when the gate score is below
 *                    gate_threshold) for many streams: a gated-off stream costs the ingest step only.
 */
int nww_stream_open(nww_engine* e, int64_t n_streams);
int nww_stream_push(nww_engine* e, const int16_t* chunks_dev, int32_t chunk_len, float* scores_dev, void* stream);
int nww_stream_push_host(nww_engine* e, const int16_t* chunks_host, int32_t chunk_len, float* scores_host);
int nww_stream_push_select(nww_engine* e, const int16_t* chunks_dev, int32_t chunk_len, const int64_t* ids_dev, int64_t n_ids,
                           float* scores_dev, void* stream);
int nww_stream_push_select_host(nww_engine* e, const int16_t* chunks_host, int32_t chunk_len, const int64_t* ids_host, int64_t n_ids,
                                float* scores_host);
int nww_stream_reset(nww_engine* e, const int64_t* ids_host, int64_t n_ids);
int nww_stream_close(nww_engine* e);

/* Front end only: log-mel in dB, (n, n_mels, n_frames) or, if time_major, (n, n_frames, n_mels).
 * Replaces MelSpectrogram + AmplitudeToDB — architectures.py:830-837, 873-875;
 * _export/onnx.py:66-83. */
int nww_logmel(nww_engine* e, const int16_t* pcm_dev, int64_t n_windows, float* mel_dev, int time_major, void* stream);

/* Per-stage device timing with CUDA events recorded on the launching stream (used by bench.py
 * for the roofline of the dominant kernel).  Stage A = the per-window kernels (front end +
 * head body), stage B = the dense tail.  nww_get_profile() synchronises the recorded events,
 * returns the totals since the last call and clears them. */
typedef struct nww_profile_t {
    double stage_a_ms, stage_b_ms;
    int64_t stage_a_spans, stage_b_spans;   /* timed launch groups (one per chunk)   */
    int64_t stage_a_windows, stage_b_windows;
} nww_profile_t;
int nww_set_profiling(nww_engine* e, int enable);
int nww_get_profile(nww_engine* e, nww_profile_t* out);

/* Pipe micro-benchmark on `device`: kind 0 = FP32 FFMA, 1 = FP64 DFMA; *tflops receives the measured peak
 * (independent FMA chains, CUDA events, best of 5; 1 FMA = 2 flop).  bench.py quotes the stage kernel's FP32 /
 * FP64 work against these measured roofs ("FP32 CUDA-core peak: builder to measure"). */
int nww_microbench(int device, int kind, double* tflops);

/* Block until everything enqueued on the engine's own streams has finished. */
int nww_synchronize(nww_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* NWW_B200_H */
