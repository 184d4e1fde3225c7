/* libars_b200 -- C ABI of the B200 (sm_100a) render path of Audio Raytracing Studio.
 *
 * The reference (CipherCorePro/Audio-Raytracing-Studio, raytracer_studio.py = "rs.py") has no
 * FFI of its own: its boundary is the Python module namespace.  Each entry point below replaces
 * the array work of one reference function (cited per function); the Python drop-in
 * `ars_b200/raytracer_studio.py` binds them with ctypes and keeps the reference's signatures.
 *
 * Conventions
 *   - plain pointers and sizes only; arrays are C-contiguous, frames-major ("(frames, channels)"),
 *     float32 unless stated; the caller owns every buffer;
 *   - every function returns 0 on success, non-zero on failure; ars_last_error() gives the
 *     message of the calling thread's last failure; nothing throws across the ABI;
 *   - there is NO CPU path: without a CUDA device (sm_100) ars_init() fails and every other
 *     call returns ARS_ERR_NO_DEVICE;
 *   - calls are serialised on one internal CUDA stream and are safe to issue from any thread;
 *   - "_dev" variants take device pointers, enqueue on the library stream and return without
 *     waiting (ars_sync() waits); all other variants take host pointers and copy both ways.
 */
#ifndef ARS_B200_H
#define ARS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ARS_API __attribute__((visibility("default")))
#else
#define ARS_API
#endif

#define ARS_OK 0
#define ARS_ERR_ARG 1
#define ARS_ERR_CUDA 2
#define ARS_ERR_NO_DEVICE 3
#define ARS_ERR_INTERNAL 4

/* rs.py:37-42 CHANNEL_LAYOUTS, in the order the reference lists them */
#define ARS_LAYOUT_STEREO 0   /* "Stereo"               2 ch: FL FR */
#define ARS_LAYOUT_5_1 1      /* "5.1 (Standard)"       6 ch: FL FR C LFE RL RR */
#define ARS_LAYOUT_7_1 2      /* "7.1 (Surround)"       8 ch: ... SL SR  (12 ms, x0.7) */
#define ARS_LAYOUT_5_1_2 3    /* "5.1.2 (Atmos Light)"  8 ch: ... TFL TFR (18 ms, x0.6*z) */

#define ARS_LUFS_OK 0         /* metrics.lufs holds a value (may be -inf)                     */
#define ARS_LUFS_NONE 1       /* shorter than one 400 ms block: the reference reports None    */
#define ARS_LUFS_SKIPPED 2    /* loudness was not requested                                   */

typedef struct ArsMetrics {   /* calculate_audio_metrics, rs.py:674-698 */
    double lufs;              /* integrated loudness of mean(ch0, ch1)  (rs.py:687-691)       */
    int32_t lufs_status;      /* ARS_LUFS_*                                                   */
    int32_t reserved;
    double peak_linear;       /* max |x| over all channels (rs.py:695)                        */
    double rms_linear;        /* sqrt(mean(x^2)) over all channels (rs.py:696)                */
    double true_peak_dbfs;    /* 20 log10(peak) or -inf (rs.py:697)                           */
    double rms_dbfs;          /* 20 log10(rms)  or -inf (rs.py:698)                           */
    /* add-on (not in the reference, whose "true peak" above is the sample peak): 4x-oversampled peak after ITU-R
     * BS.1770-4 Annex 2, over every output channel; filled when ArsRenderParams.want_lufs has bit 1 (value 2) set */
    double true_peak_4x_dbfs;
    int32_t true_peak_4x_status;  /* 0: computed, 1: not requested                                  */
    int32_t reserved2;
} ArsMetrics;

/* Random draws of generate_impulse_response_split_3d replayed by the caller
 * (np.random.randint / uniform in the reference's call order, rs.py:262,264,285). */
typedef struct ArsIrDraws {
    const int64_t* tap_delay;   /* ntaps accepted delays (0 < d < split), draw order          */
    const double* tap_base;     /* ntaps base strengths, uniform(0.3, 0.8)                    */
    int32_t ntaps;
    int32_t reserved;
    const double* noise;        /* late_len raw tail noise, uniform(-1, 1); values must lie in [-1, 1]: the folded-air
                                 * route bounds the non-zero extent of the tail from that range                 */
    int64_t noise_len;          /* = length - split                                            */
} ArsIrDraws;

/* Scalars of one render, already shaped by the host prologue (rs.py:157-236). */
typedef struct ArsRenderParams {
    double rate;                /* sample rate                                                 */
    int32_t external_ir;        /* 0: procedural hall (rs.py:1047-1056)  1: stereo IR (1027-1045) */
    int32_t layout;             /* ARS_LAYOUT_*                                                */
    /* procedural IR (ignored when external_ir) -- arguments of rs.py:238 */
    double ir_duration, ir_max_delay, absorption, directionality, ir_split_time, diffusion;
    /* mix / EQ -- arguments of rs.py:338 / 410 (levels already adapted, rs.py:168-182) */
    double early_level, late_level, dry_wet, kill_start, bass_gain, treble_gain, air_absorption;
    /* position -- rs.py:464, 517 */
    double x, y, z;
    int32_t want_lufs;          /* bit 0: also run the loudness meter; bit 1: also the 4x-oversampled true peak */
    int32_t reserved;
} ArsRenderParams;

/* ---- life cycle ---------------------------------------------------------------------------- */
ARS_API int ars_init(int device);                 /* device < 0: device 0.  Idempotent.                 */
ARS_API void ars_shutdown(void);
ARS_API int ars_sync(void);                       /* wait for everything enqueued by _dev calls         */
ARS_API const char* ars_last_error(void);
ARS_API const char* ars_version(void);
ARS_API uint64_t ars_launch_count(void);          /* kernels this library has launched so far           */
ARS_API uint64_t ars_air_fold_count(void);        /* convolution stages that took the folded-air route  */
ARS_API uint64_t ars_olsb_count(void);            /* ... that took the big-block overlap-save route      */
ARS_API uint64_t ars_head_start_count(void);      /* asynchronous renders whose head overlapped the tail of the render before */
ARS_API uint64_t ars_meter_stream_count(void);    /* asynchronous renders whose loudness meter ran on the meter stream (option "loud_stream") */
ARS_API uint64_t ars_tail_overlap_count(void);    /* asynchronous renders whose last passes did not wait for the final pass of the render before (option "tail_overlap") */
ARS_API void* ars_stream(void);                   /* the library's cudaStream_t (for event timing)      */
/* Options: "upols" (1 = use the partitioned overlap-save convolution whenever a render has no exact-N
 * spectral mask, i.e. air <= 0.01 and both EQ gains ~ 1 [default]; 0 = always the N-point spectral filter),
 * "upols_logf" (12 | 13: points per overlap-save transform = 2^logf, hop = half of it),
 * "sparse_ir" (1 = IR spectra of procedural / sparse IRs through the cached overlap-save route [default]; 0 = always
 * two M-point transforms),
 * "air_fold" (1 = a render whose only spectral mask is the air-absorption ramp [air > 0.01, EQ gains ~ 1] folds the
 * ramp into the impulse response and runs as one overlap-save convolution [default]; 0 = exact N-point filter),
 * "air_fold_eps_e9" (bound on the late path's transfer-function error the fold may introduce, in 1e-9; default 2000),
 * "air_fold_max_taps" (longest half-length of the truncated air kernel [default 131072: every air value at 2e-6];
 * a longer one -- or, above 32768 taps, one whose fold would cost more than half the N-point route -- takes that route),
 * "side_stream" (1 = folded-air renders run IR synthesis, fold and IR spectra on a second stream next to the delay-line
 * transform [default]),
 * "ols_r2" (1 = the 8192-point overlap-save transforms run as a radix-2 stage folded into the window load / output
 * store plus two 4096-point transforms; 0 = one four-stage 8192-point tile [default: measured faster]),
 * "lufs_fused" (1 = one-pass loudness meter: both K-weighting stages and the hop energies in one kernel [default];
 * 0 = one pass per stage and step), "lufs_from_stage" (1 = inside a render the meter recomputes its feed from the
 * convolution stage's output and runs next to the final pass; 0 = the final pass writes a feed array [default: measured
 * faster, both kernels are bound by issue slots]),
 * "mac_tiled_min" (partition count above which dense IRs use the register-tiled multiply-accumulate kernel),
 * "olsb" (1 = mask-free and folded-air convolutions whose taps fit a quarter of a two-pass transform run as ONE-partition
 * overlap-save over 2^18..2^22-point blocks: strided forward pass, fused middle pass [contiguous forward x IR spectrum x
 * contiguous inverse], strided inverse pass [default]; 0 = always the 4096-frame partitioned form),
 * "olsb_logf" (0 = block length chosen from the tap count [default] | 18..22), "olsb_stripe" (0 = transforms per
 * L2-resident stripe chosen from the SM count [default] | n), "stream_hints" (bit mask, default 15: data touched once moves
 * with the evict-first cache policy -- 1 signal frames read by the first pass, 2 output frames stored by the last pass,
 * 4 PCM / float frames stored and 8 stereo frames read by the final pass),
 * "final_lean" (1 [default] | 2 = the final pass of the 5.1-based layouts runs its lean frame loop -- packed FP32x2 guard
 * division behind one range test per frame, 32-bit offsets, one | two frames per step -- whenever only the stereo peak
 * guard can be active; 0 = the general loop.  Bit-identical results either way),
 * "final_in_meter" (1 = a whole render that wants the loudness runs final pass + one-pass meter as ONE kernel: a meter
 * CTA produces its block's frames itself and filters the feed from shared memory; 0 = final pass, then meter [default:
 * measured faster -- 109 + 77 us against 214 us on the 300 s render; the results are bit-identical]),
 * "head_start" (1 [default] | 0: see ars_render_dev_async),
 * "loud_stream" (1 [default] | 0: see ars_render_dev_async), "tail_overlap" (0 [default] | 1: likewise).
 * Environment (read once per process, experiments): ARS_MID_PIPE (3 [default] | 2 = the plain middle pass runs as a persistent
 * kernel that fetches its next tile with cp.async.bulk, 3 | 2 CTAs per SM; 0 = one tile per CTA), ARS_LAST_PIPE (3 [default] |
 * 2 | 6 = the last pass of 2^18-point blocks likewise, double-buffered; 0 = one tile per CTA), ARS_MID_NT (256 | 512),
 * ARS_LOUD_PRIO (1 = the meter stream of "loud_stream" gets a high priority), ARS_LUFS_FEED_CTAS (1..8, default 6: CTAs per
 * SM of the one-pass meter over a materialised feed); measurements: profiles/r02_stream_options_ab.txt. */
ARS_API int ars_set_option(const char* key, int32_t value);
/* Page-locked host blocks for results: the Python layer wraps them as numpy arrays (freed through ars_host_free when the
 * array dies; the library keeps up to 4 GiB of freed blocks for the next render).  Copies from / to PAGEABLE host memory
 * of 4 MiB and more go through a ring of pinned slots filled by a few worker threads (option "host_staging", default 1). */
ARS_API void* ars_host_alloc(int64_t bytes);
ARS_API void ars_host_free(void* p);
/* CUDA-event stopwatch on the library stream: begin records, end records + waits + reports ms. */
ARS_API int ars_timer_begin(void);
ARS_API int ars_timer_end(float* ms);
/* Per-launch event timing: between begin and end every kernel (or short chain of small kernels) the library launches
 * is bracketed by two CUDA events on its own stream.  end reports the FFT pass kernels' count, summed duration and
 * algorithmic bytes; ars_profile_report() then returns a JSON array with one {"name", "launches", "ms", "bytes"} object
 * per kernel name of that session (valid until the next session ends).  Turn "side_stream" and "olsb_lanes" off around
 * a session to time the kernels one after another. */
ARS_API int ars_profile_begin(void);
ARS_API int ars_profile_end(int64_t* launches, double* ms, double* bytes);
ARS_API const char* ars_profile_report(void);

/* ---- stage entry points (host pointers), one per reference function ------------------------- */

/* generate_impulse_response_split_3d, rs.py:238-305.  early/late: float32[length],
 * length = max(1, (int64)(ir_duration * rate)).  Returns ARS_ERR_ARG if `length` disagrees. */
ARS_API int ars_ir_synth(double rate, double ir_duration, double ir_max_delay, double absorption,
                 double directionality, double ir_split_time, double diffusion, const ArsIrDraws* draws,
                 float* early, float* late, int64_t length);
/* integer geometry used by the caller to size the draws: rs.py:249,254-255,259,271-272 */
ARS_API int ars_ir_geometry(double rate, double ir_duration, double ir_max_delay, double ir_split_time,
                    int64_t* length, int64_t* split, int64_t* tap_hi, int64_t* late_len);

/* apply_simple_lp_filter, rs.py:310-333, on an (n, 2) signal; caller applies the `< 0.01` gate. */
ARS_API int ars_air_filter(const float* sig, int64_t n, double rate, double air, float* out);

/* scipy.signal.resample(x, num, axis=0) of an (n, 2) signal -> (num, 2): the IR rate conversion of
 * apply_raytrace_convolution_3d, rs.py:1037-1040 (exact n- and num-point DFTs on the GPU). */
ARS_API int ars_resample(const float* sig, int64_t n, int64_t num, float* out);

/* dynamic_dry_wet_mix, rs.py:84-121; out has max(n_dry, n_wet) frames of `ch` channels. */
ARS_API int ars_dry_wet_mix(const float* dry, int64_t n_dry, const float* wet, int64_t n_wet, int32_t ch,
                    double dry_wet, double kill_start, float* out);

/* convolve_audio_split_3d, rs.py:338-408.  data: (n, cin) with cin == 1 (duplicated), 2, or > 2
 * (first two used); early/late: float32[L_early], float32[L_late] (either may be NULL/0 =
 * "zeros(1)").  out: (n_out, 2), n_out = ars_convolve_out_len(n, L_early, L_late). */
ARS_API int64_t ars_convolve_out_len(int64_t n, int64_t len_early, int64_t len_late);
ARS_API int ars_convolve_split(const float* data, int64_t n, int32_t cin, const float* early, int64_t len_early,
                       const float* late, int64_t len_late, double early_level, double late_level,
                       double dry_wet, double bass_gain, double treble_gain, double rate, double kill_start,
                       double air_absorption, float* out);

/* convolve_audio_external_ir, rs.py:410-462.  ir: (L, 2).  out: (n + L - 1, 2). */
ARS_API int ars_convolve_external(const float* data, int64_t n, int32_t cin, const float* ir, int64_t L,
                          double dry_wet, double bass_gain, double treble_gain, double rate,
                          double kill_start, float* out);

/* apply_surround_panning_3d, rs.py:464-501.  stereo: (n, 2); out: (n, 6). */
ARS_API int ars_pan(const float* stereo, int64_t n, double x, double y, double z, float* out);

/* apply_delay, rs.py:507-515.  (n, ch) -> (n, ch). */
ARS_API int ars_delay(const float* sig, int64_t n, int32_t ch, int64_t delay_samples, float* out);

/* map_channels, rs.py:517-563.  six: (n, 6); out: (n, C), C = ars_layout_channels(layout). */
ARS_API int ars_layout_channels(int32_t layout);
ARS_API int ars_map_channels(const float* six, int64_t n, int32_t layout, double rate, double z, float* out);

/* calculate_audio_metrics, rs.py:674-698.  data: (n, ch). */
ARS_API int ars_metrics(const float* data, int64_t n, int32_t ch, double rate, int32_t want_lufs, ArsMetrics* out);

/* Numerics of the A/B report (run_audio_profiler_v4, rs.py:769-798): per-channel RMS sqrt(mean(x^2)) of an
 * (n, ch <= 8) array and the RMS of the side signal (ch0 - ch1) * 0.5 (0 for mono).  rms_out: ch floats. */
/* 4x-oversampled true peak (dBTP) of an (n, ch) array: BS.1770-4 Annex 2 polyphase FIR, maximum over all channels
 * (add-on; the reference reports the sample peak only) */
ARS_API int ars_true_peak_4x(const float* data, int64_t n, int32_t ch, double* dbtp);
ARS_API int ars_channel_rms(const float* data, int64_t n, int32_t ch, float* rms_out, float* side_rms_out);
/* Spectrogram of the visualiser (rs.py:626-634): scipy.signal.spectrogram(data[:, 0], fs=rate, window='hann', nperseg,
 * noverlap = nperseg / 2) with scipy's defaults (constant detrend, one-sided density).  nperseg: a power of two in
 * 2..8192.  sxx_out: (nperseg / 2 + 1) rows x ars_spectrogram_segments(n, nperseg) columns, row-major float32. */
ARS_API int64_t ars_spectrogram_segments(int64_t n, int32_t nperseg);
ARS_API int ars_spectrogram(const float* data, int64_t n, int32_t ch, double rate, int32_t nperseg, float* sxx_out);

/* clip + scrub + float->PCM16, rs.py:1082-1084 (libsndfile rule lrintf(x * 32767)). */
ARS_API int ars_pcm16(const float* data, int64_t count, int16_t* out);

/* ---- the whole render on arrays: compute part of apply_raytrace_convolution_3d, rs.py:1020-1084
 * in: (n, cin) float32.  ext_ir: (L, 2) when params->external_ir, else NULL and `draws` feeds the
 * procedural IR.  Outputs (any may be NULL): out_stereo (N, 2) = the convolution stage result,
 * out_f32 (N, C) pre-clip final, out_pcm (N, C) int16, metrics.  N = ars_render_out_len(...). */
ARS_API int64_t ars_render_out_len(const ArsRenderParams* p, int64_t n, int64_t ext_ir_len);
ARS_API int ars_render(const ArsRenderParams* p, const float* in, int64_t n, int32_t cin, const float* ext_ir,
               int64_t ext_ir_len, const ArsIrDraws* draws, float* out_stereo, float* out_f32,
               int16_t* out_pcm, ArsMetrics* metrics);
/* Same with every array pointer (including those inside `draws`) on the device.  Asynchronous
 * unless `metrics` is non-NULL (the loudness gate needs the block energies on the host). */
/* (device pointers of the *_dev entry points must be 8-byte aligned: stereo frames are read as float2) */
ARS_API int ars_render_dev(const ArsRenderParams* p, const float* d_in, int64_t n, int32_t cin, const float* d_ext_ir,
                   int64_t ext_ir_len, const ArsIrDraws* d_draws, float* d_out_stereo, float* d_out_f32,
                   int16_t* d_out_pcm, ArsMetrics* metrics);
/* The same, never waiting: with `metrics` non-NULL the render's state block goes to a pinned slot and *metrics is filled
 * in by the next call that waits for the library stream anyway -- ars_sync() or ars_timer_end() -- until which it must
 * stay valid.  Lets a caller enqueue render after render (a batch of device-resident clips) without the GPU idling while
 * the host reads 80 bytes back and prepares the next one.
 * Head start (option "head_start", default 1): when the call right before this one was a render of the same geometry (same
 * parameters, frame count, channel count, IR length), the part of this render that needs nothing from that one -- IR
 * synthesis / fold / IR spectrum and the first pass of every transform -- is ordered after that render's CONVOLUTION, not
 * after its tail, and runs next to its final pass and loudness meter on the library's internal streams.  The input arrays
 * (d_in, d_draws->noise, d_ext_ir) must therefore be COMPLETE when the call is made (e.g. produced before an ars_sync() or
 * a device synchronisation), not merely ordered on ars_stream(); the outputs are ordered on ars_stream() as always.
 * Meter stream (option "loud_stream", default 1): the loudness meter of such a render and the read-back of its state block
 * run on a third internal stream behind its final pass, over one of two alternating feed buffers / state blocks, so the
 * last pass of the NEXT render follows this render's final pass directly instead of its meter and gating.  The PCM / float
 * frames are ordered on ars_stream() as always; the metrics arrive with the next waiting call as before, and every other
 * entry point first orders ars_stream() after the meter stream.  Measured on the 300 s render: 0.4636 -> 0.437 ms.  The
 * stream, the second feed buffer and the second state block are created on first use (3-90 ms, once): warm a timed loop up
 * through this call.
 * Tail overlap (option "tail_overlap", default 0; needs the meter stream): the stage output alternates between two buffers
 * as well and the meter stream zeroes a state block behind its read-back, so the last passes of a render wait for the render
 * BEFORE the previous one to be through with the slot, not for the previous render's final pass: the convolution of render
 * k+1 overlaps the whole tail of render k (measured: 0.4367 ms, no better than the meter stream alone). */
ARS_API int ars_render_dev_async(const ArsRenderParams* p, const float* d_in, int64_t n, int32_t cin, const float* d_ext_ir,
                         int64_t ext_ir_len, const ArsIrDraws* d_draws, float* d_out_stereo, float* d_out_f32,
                         int16_t* d_out_pcm, ArsMetrics* metrics);


/* ---- a batch of independent renders (host buffers; pin them for full overlap) -----------------
 * Clip i+1's host->device copy and clip i-1's device->host copy overlap clip i's compute
 * (two buffer slots, one copy stream per direction).  Results are identical to calling
 * ars_render on every clip in turn.  Returns when every output buffer is filled. */
typedef struct ArsClip {
    const ArsRenderParams* params;
    const float* in;            /* (n, cin) float32                                            */
    int64_t n;
    int32_t cin;
    int32_t reserved;
    const float* ext_ir;        /* (ext_ir_len, 2) when params->external_ir, else NULL         */
    int64_t ext_ir_len;
    const ArsIrDraws* draws;    /* procedural-IR draws (host pointers), NULL for external IR   */
    float* out_f32;             /* (N, C) or NULL                                              */
    int16_t* out_pcm;           /* (N, C) or NULL                                              */
    ArsMetrics* metrics;        /* or NULL                                                     */
} ArsClip;
ARS_API int ars_render_batch(const ArsClip* clips, int32_t count);


/* ---- one long mask-free render split by overlap-save block ranges (one rank of a multi-GPU render) ------
 * Only renders without an exact-N spectral mask (air <= 0.01, both EQ gains ~ 1) shard this way; the others need
 * the global N-point transform and run on one GPU (ars_render).  All pointers are device pointers.
 *
 * The per-render device state is an opaque block of ars_state_bytes() bytes, zeroed by the caller.  Its words
 * are what the ranks reduce between the phases (all values are bit patterns of non-negative floats, so an
 * integer MAX is a float MAX):
 *   uint32 [0..3] max |y|, max |L|, max |R|, max |float32(L+R)| of the convolution output  -> MAX after convolve
 *   uint32 [4] max |six| (pan)                                                       -> MAX after tail phase 0
 *   uint32 [5] max |out| (Stereo map only)                                           -> MAX after tail phase 1
 *   uint32 [8] peak, [9] loudness-feed peak -> MAX ; double at byte 48: sum of squares -> SUM  (after phase 2)
 *   double at byte 56: integrated loudness (written by ars_loudness_dev)
 * ars_long_plan says how the render splits: blocks of block_frames output frames (one overlap-save transform of the
 * big-block route, or 4096 frames of the partitioned route).  Rank r computes output blocks [block_lo, block_hi) from
 * an input slice that starts at absolute frame x_frame0 and holds x_frames frames; the slice must reach halo_frames
 * frames before frame block_lo * block_frames (or to frame 0).  y / outputs are slices too.  Every rank must use the same
 * library options, so that every rank takes the same route: the result is then bit-identical to the one-GPU render. */
typedef struct ArsLongPlan {
    int64_t block_frames;       /* output frames per block                                                   */
    int64_t n_blocks;           /* blocks of the whole render: ceil(frames_out / block_frames)               */
    int64_t halo_frames;        /* input frames needed in front of a block range                             */
    int64_t frames_out;         /* n_total + ir_len - 1                                                      */
    int32_t hop_count;          /* 100 ms hops of the loudness meter (ars_long_loudness_*), 0: clip too short */
    int32_t route;              /* 0: partitioned overlap-save, 1: big-block overlap-save                     */
} ArsLongPlan;
ARS_API int64_t ars_state_bytes(void);
ARS_API int64_t ars_ols_block_frames(void);       /* block length of the partitioned route (route 0)         */
ARS_API int ars_long_plan(const ArsRenderParams* p, int64_t n_total, int64_t ir_len, ArsLongPlan* out);
ARS_API int ars_long_convolve_dev(const ArsRenderParams* p, const float* d_x, int64_t x_frame0, int64_t x_frames,
                                  int64_t n_total, int32_t cin, const float* d_ir0, int64_t L0, const float* d_ir1,
                                  int64_t L1, int64_t block_lo, int64_t block_hi, float* d_y, int64_t y_frame0,
                                  void* d_state);
/* phase 0: pan maximum; 1: map maximum (Stereo layout only, no-op otherwise); 2: final pass writing the PCM /
 * float / loudness-feed slices for frames [frame_lo, frame_hi) (outputs are indexed from frame_lo). */
ARS_API int ars_long_tail_dev(const ArsRenderParams* p, int32_t phase, const float* d_y, int64_t y_frame0,
                              int64_t frame_lo, int64_t frame_hi, int64_t N_total, void* d_state, float* d_out_f32,
                              int16_t* d_out_pcm, float* d_mono);
/* The loudness meter split over the ranks: a rank leaves the hop energies of ITS frames [frame_lo, frame_hi) -- fed from
 * the stage output y, whose slice must reach 3 x 8192 frames before frame_lo (or to frame 0) for the filters' warm-up --
 * in d_hops[hop_count] (absolute hop index, zeros elsewhere) and max-merges the feed's peak into state word [9]; the
 * ranks SUM the vectors, then every rank (or one) gates: the loudness lands in the state block. */
ARS_API int ars_long_loudness_hops_dev(const ArsRenderParams* p, const float* d_y, int64_t y_frame0, int64_t frame_lo,
                                       int64_t frame_hi, int64_t N_total, void* d_state, double* d_hops, int32_t n_hops);
ARS_API int ars_long_loudness_gate_dev(const ArsRenderParams* p, const double* d_hops, int32_t n_hops, int64_t N_total,
                                       void* d_state, int32_t* lufs_status);
/* loudness meter on a (gathered) mono feed; enqueued, result lands in the state block */
ARS_API int ars_loudness_dev(const float* d_mono, int64_t N, double rate, void* d_state, int32_t* lufs_status);
/* read the state block back and turn it into metrics (synchronises) */
ARS_API int ars_state_metrics(const void* d_state, int64_t sample_count, int32_t lufs_status, ArsMetrics* out);

/* ---- the whole-render array of a block-sharded render, filled over NVLink by the ranks themselves -------------------
 * (the reference returns ONE (frames, channels) array, rs.py:1082-1084; with one process per GPU the segments have to
 * reach the rank that holds it.)  Rank 0 allocates the array with ars_peer_alloc and passes the 64-byte handle to the
 * other processes of the node (any byte transport: a broadcast); they map it with ars_peer_open (CUDA IPC, peer access
 * over NVLink enabled on the way) and PUSH their PCM segment straight into its place with ars_peer_push -- a device-to-
 * device copy over the peer mapping on `stream` (NULL: the library stream), no receive side and no staging buffer.  The
 * owner may read the array once every pushing rank has passed a collective that it enqueued behind its push.
 * ars_peer_close unmaps (non-owners), ars_peer_free releases (owner, after every rank has unmapped). */
#define ARS_PEER_HANDLE_BYTES 64
ARS_API int ars_peer_alloc(int64_t bytes, void** d_ptr, unsigned char* handle);
ARS_API int ars_peer_open(const unsigned char* handle, void** d_ptr);
ARS_API int ars_peer_push(void* d_peer_dst, const void* d_src, int64_t bytes, void* stream);
ARS_API int ars_peer_close(void* d_ptr);
ARS_API int ars_peer_free(void* d_ptr);

#ifdef __cplusplus
}
#endif
#endif /* ARS_B200_H */
