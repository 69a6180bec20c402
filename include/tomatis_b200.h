/* tomatis_b200.h -- C ABI of libtomatis_b200.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (xyjk0511/tomatis-audio-processor) has no FFI boundary of its own: its contract is
 * the Python `process(in_path, out_path, ...)` functions of
 *     src/process_tomatis.py:160-178, src/process_tomatis_xfade.py:55-71,
 *     src/process_tomatis_adaptive.py:157-172
 * whose *inside* (everything between "samples read" and "samples written") this library
 * replaces.  Each entry point below names the reference code it stands in for.  The host side
 * (tomatis_audio_processor_b200/*.py) binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (TMT_OK) or a negative
 * error code and never throws; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 * "device pointer" arguments are owned by the caller (e.g. torch tensors).  Audio is interleaved
 * stereo float32, one sample-frame ("sf") = {L, R}.  n_fft = 4096, hop = 2048 (the reference's
 * defaults, src/process_tomatis.py:174-175) are the supported geometry.
 */
#ifndef TOMATIS_B200_H
#define TOMATIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMT_OK 0
#define TMT_ERR_INVALID (-1)     /* bad argument */
#define TMT_ERR_CUDA (-2)        /* CUDA runtime error, see tmt_last_error */
#define TMT_ERR_UNSUPPORTED (-3) /* geometry / parameter outside what the kernels implement */
#define TMT_ERR_NOMEM (-4)

/* Framing modes (how frames and output blocks are laid over the file). */
#define TMT_FRAMING_STREAMING 0 /* standard + xfade: first frame at -n_fft/2, tail zero pad `pad_end`,      \
                                   out/(sum w^2 + 1e-12), per-chunk limiter                                 \
                                   (src/process_tomatis.py:270-272,310-312,419-426,447-453) */
#define TMT_FRAMING_WHOLEFILE 1 /* adaptive: frames at k*hop with 0 <= k*hop < total, y/max(sum w^2,1e-8),   \
                                   one global limiter (src/process_tomatis_adaptive.py:298-345) */

#define TMT_FRAMING_EQ_PAD 2    /* static EQ: n_fft/2 zeros on both sides, frames at -n_fft/2 + k*hop while they  \
                                   fit, EVERY position the frames cover is output (length (n_frames+1)*hop, shifted by n_fft/2),  \
                                   out/(sum w^2 + 1e-12), one global peak (src/layer2_apply_eq.py:100-215) */
#define TMT_FRAMING_EQ_NOPAD 3  /* same without padding (--no_pad): frames at k*hop */

/* Gate automata. */
#define TMT_GATE_UPDELAY 0 /* hysteresis + up-delay, src/process_tomatis.py:373-385 */
#define TMT_GATE_MINHOLD 1 /* hysteresis + min-hold, src/process_tomatis_adaptive.py:87-121 */

/* Per-frame / per-track arrays owned by a plan (see tmt_plan_read / tmt_plan_write). */
#define TMT_ARR_MEANSQ_F32 0   /* float  [frames]  np.mean(mono*mono) of each frame, bit-exact   */
#define TMT_ARR_MEANSQ_F64 1   /* double [frames]  same in the adaptive float64 branch            */
#define TMT_ARR_GATE_F64 2     /* double [frames]  caller-supplied gate input (levels in dBFS)    */
#define TMT_ARR_STATE 3        /* uint8  [frames]  1 = C1, 2 = C2                                 */
#define TMT_ARR_ROW 4          /* uint16 [frames]  gain-table row (crossfade counter k)           */
#define TMT_ARR_C2_COUNT 5     /* int32  [tracks]  number of C2 frames                            */
#define TMT_ARR_CHUNK_PEAK 6   /* float  [chunks]  max |y| of each limiter chunk before limiting  */
#define TMT_ARR_INPUT_PEAK 7   /* float  [tracks]  max |x|                                        */
#define TMT_ARR_HOPSUM_F32 8   /* float  [frames + tracks] pairwise sum of each hop block         */
#define TMT_ARR_HOPSUM_F64 9   /* double [frames + tracks]                                        */
#define TMT_ARR_BISECT_T 10        /* double [tracks]       threshold found by tmt_plan_bisect                     */
#define TMT_ARR_BISECT_ITERS 11    /* int32  [tracks]       search steps taken                                     */
#define TMT_ARR_BISECT_TRACE_T 12  /* double [tracks * 32]  T_mid of each step                                     */
#define TMT_ARR_BISECT_TRACE_C2 13 /* int32  [tracks * 32]  C2 frame count of each step                            */

typedef struct tmt_engine tmt_engine; /* per device: window, twiddles, gain rows */
typedef struct tmt_plan tmt_plan;     /* per batch of tracks (or file shard): geometry + scratch */

/* One track, or one time shard of a long file.  Positions are absolute sample-frame indices in the
 * file; the in/out device buffers may cover only a window of it (halo sharding). */
typedef struct tmt_track_desc {
    const void* pcm_in; /* device, float32 [in_len][2], element 0 is file position in_origin  */
    void* pcm_out;      /* device, float32 [out_len][2], element 0 is file position out_origin; must not
                           overlap pcm_in (work units read their neighbours' input while those write output);
                           may be NULL with out_len = 0 for a plan that only runs levels / gate (analysis) */
    int64_t total;      /* file length in sample-frames (zero padding / clipping refer to it)  */
    int64_t in_origin, in_len;
    int64_t out_origin, out_len;
    int64_t block_lo, block_hi; /* output hop-blocks [lo,hi) this plan produces; hi < 0 = all  */
} tmt_track_desc;

int tmt_version(void);
const char* tmt_error_string(int code);
/* Last error text of the calling thread (valid until the next failing call on that thread). */
const char* tmt_last_error(void);

/* ---- engine ------------------------------------------------------------------------------- */
/* The library is built once per frame size the fused kernels serve (src/process_tomatis.py:174-175, `--n_fft/--hop`):
 * libtomatis_b200.so = 4096 / 2048 (the reference's defaults), libtomatis_b200_n2048.so = 2048 / 1024 (the documented faster
 * setting); both export exactly this header.  A build rejects every other (n_fft, hop) with TMT_ERR_UNSUPPORTED; other sizes
 * go through the tmt_generic_* entry points. */
int tmt_engine_create(tmt_engine** out, int device, int n_fft, int hop);
int tmt_engine_destroy(tmt_engine* e);
/* Analysis = synthesis window, `np.hanning(n_fft).astype(float32)` (src/process_tomatis.py:266).
 * Host pointer. */
int tmt_engine_set_window(tmt_engine* e, const float* win, int n);
/* Gain table: n_rows rows of n_fft/2+1 linear gains in natural bin order, as produced by
 * db_to_lin(build_tilt_gain_db(...)) (src/process_tomatis.py:105-158); row r is used by frames whose
 * TMT_ARR_ROW value is r (replaces the per-frame `gain = g1 if state == 1 else g2` / dB-domain mix,
 * src/process_tomatis.py:392, _xfade.py:270-274, _adaptive.py:302-304).  Host pointer. */
int tmt_engine_set_gain_rows(tmt_engine* e, const float* rows, int n_rows, int n_bins);

/* ---- plan --------------------------------------------------------------------------------- */
/* unit_blocks: output hop-blocks per CTA work unit of the STFT kernel (0 = default). */
int tmt_plan_create(tmt_engine* e, tmt_plan** out, int framing, int n_tracks, const tmt_track_desc* tracks,
                    int unit_blocks);
int tmt_plan_destroy(tmt_plan* p);
/* Re-point the audio buffers (same geometry), e.g. to ping-pong staging buffers. */
int tmt_plan_set_buffers(tmt_plan* p, int track, const void* pcm_in, void* pcm_out);

/* Restrict / widen what tmt_plan_levels computes for one track: hop-block sums for blocks [hb_lo, hb_hi), mean squares for
 * frames [f_lo, f_hi).  A time shard sums only the hop blocks it owns (no halo needed), all-reduces TMT_ARR_HOPSUM_* and
 * then derives every frame's mean square locally, while its halo hand-off is still in flight. */
int tmt_plan_set_level_ranges(tmt_plan* p, int track, int hb_lo, int hb_hi, int f_lo, int f_hi);

int tmt_plan_total_frames(const tmt_plan* p);  /* sum of n_frames over tracks                  */
int tmt_plan_total_chunks(const tmt_plan* p);
int tmt_plan_total_units(const tmt_plan* p);
/* Limiter chunks the plan does not produce completely (a time shard that owns part of a chunk): their peaks have to be combined
 * across shards before tmt_plan_limiter; 0 means tmt_plan_stft_limited alone finishes the job. */
int tmt_plan_unfusable_chunks(const tmt_plan* p);
/* Dev counters of the fused limiter since the last reset: out2[0] = nanoseconds CTAs spent rescaling finished chunks (summed over
 * CTAs), out2[1] = number of rescales.  Synchronises the device. */
int tmt_plan_debug_counters(tmt_plan* p, uint64_t* out2, int reset);
int tmt_plan_track_frames(const tmt_plan* p, int track);      /* n_frames of one track          */
int tmt_plan_track_frame_base(const tmt_plan* p, int track);  /* its offset in per-frame arrays */
int tmt_plan_track_chunks(const tmt_plan* p, int track);
int tmt_plan_track_chunk_base(const tmt_plan* p, int track);
/* Limiter chunk c of `track`: file sample range [*s0, *s1) (already clipped to [0,total)). */
int tmt_plan_chunk_range(const tmt_plan* p, int track, int c, int64_t* s0, int64_t* s1);
/* The same geometry in bulk (one call per plan instead of one per track / chunk): per-track arrays of n_tracks int32 each
 * (NULL skips one), and every chunk range of the plan as [total_chunks][2] int64 (s0, s1) in plan order. */
int tmt_plan_geometry(const tmt_plan* p, int32_t* n_frames, int32_t* frame_base, int32_t* n_chunks, int32_t* chunk_base);
int tmt_plan_chunk_ranges(const tmt_plan* p, int64_t* ranges);

/* Copy `count` elements starting at `offset` of a plan array to/from `ptr` (host if is_device==0,
 * else device); synchronises `stream` for host copies. */
int tmt_plan_read(tmt_plan* p, int which, int64_t offset, int64_t count, void* ptr, int is_device, void* stream);
/* Several whole arrays in one call: n array ids, n host destinations (each large enough for its array); device-to-host copies are
 * queued behind the plan's work on `stream` and the call waits once.  A single-file call of the streaming modes needs four
 * arrays back (mean squares, states, rows, chunk peaks): one wait instead of four. */
int tmt_plan_read_many(tmt_plan* p, int n, const int32_t* which, void* const* dst, void* stream);
int tmt_plan_write(tmt_plan* p, int which, int64_t offset, int64_t count, const void* ptr, int is_device,
                   void* stream);

/* ---- kernels ------------------------------------------------------------------------------ */
/* max |x| per track -> TMT_ARR_INPUT_PEAK (src/process_tomatis_adaptive.py:201). */
int tmt_plan_input_peaks(tmt_plan* p, void* stream);

/* K2a.  Per-frame mean square m = np.mean(mono*mono), mono = sqrt(mean(frame**2, axis=1)), bit-exact
 * with NumPy's pairwise summation (rms_dbfs + caller, src/process_tomatis.py:43-52,370;
 * compute_frame_levels, src/process_tomatis_adaptive.py:57-84).  in_scale (host, per track, may be
 * NULL = 1) is the adaptive pre-attenuation x*atten_lin applied in float32 before squaring
 * (src/process_tomatis_adaptive.py:215); flags: TMT_LEVELS_F64 selects the float64 branch of that file,
 * TMT_LEVELS_MONO the single-channel level formula.
 * Result in TMT_ARR_MEANSQ_F32 / _F64. */
#define TMT_LEVELS_F64 1  /* float64 branch of the adaptive mode */
#define TMT_LEVELS_HOPSUM_ONLY 8  /* only the hop-block sums (TMT_ARR_HOPSUM_*), no frame mean squares */
#define TMT_LEVELS_MEANSQ_ONLY 16 /* only m[k] = (H[k] + H[k+1]) / n_fft from the hop-block sums already in the plan */
#define TMT_LEVELS_MONO 2 /* single-channel file carried in the L lane (R = 0): mono = sqrt(x*x), _adaptive.py:74,180-181 */
#define TMT_LEVELS_LEFT 32  /* level of the left channel alone, np.mean(x*x) (src/analyze_stereo_state.py:16-19,112) */
#define TMT_LEVELS_RIGHT 64 /* level of the right channel alone (src/analyze_stereo_state.py:113) */
#define TMT_LEVELS_POWER_EPS 128 /* mono = sqrt(0.5*(l*l + r*r) + 1e-12): calibration front end (src/calibrate_to_baseline_v2.py:8-15) */
int tmt_plan_levels(tmt_plan* p, int flags, const float* in_scale, void* stream);

/* Files with more than two channels (adaptive mode: the reference's `for c in range(ch)` loop, src/process_tomatis_adaptive.py:
 * 307-313).  The caller cuts the file into channel pairs (tmt_channels_split), makes every pair a track of one plan and runs the
 * stereo path on them; this entry point computes the level that all of them share -- mono = sqrt(mean(frame**2, axis=1)) over ALL
 * channels (:74; NumPy's pairwise order along the channel axis, sum / channels in the array's dtype) -- from the interleaved file
 * x (device, float32 [total][channels]) and stores the hop-block sums / mean squares into every track of the plan.  Tracks must
 * have identical geometry.  flags: TMT_LEVELS_F64, TMT_LEVELS_HOPSUM_ONLY.  in_scale: as tmt_plan_levels (entry 0 is used). */
int tmt_plan_levels_multichannel(tmt_plan* p, int flags, const float* in_scale, const float* x, int channels, void* stream);
/* interleaved float32 [total][channels] -> ceil(channels / 2) stereo planes [pair][total][2] (an odd last channel gets a silent
 * partner), and back (the partner is dropped).  Device pointers. */
int tmt_channels_split(const float* in, int64_t total, int channels, float* pairs, void* stream);
int tmt_channels_merge(const float* pairs, int64_t total, int channels, float* out, void* stream);

/* K2b.  Gate automaton + crossfade counter as a block-level scan over frames.
 * gate_input: TMT_ARR_MEANSQ_F32, TMT_ARR_MEANSQ_F64 or TMT_ARR_GATE_F64.  Frame is "hi" when
 * value >= on[track], "lo" when value <= off[track] (host arrays of n_tracks doubles; thresholds in the
 * domain of the gate input; both NULL: keep the thresholds already on the device).  param = consecutive hi frames needed to switch up (UPDELAY:
 * ceil(up_delay_samples/hop)+1) or min_hold_frames (MINHOLD).  xfade_frames = 0 -> hard switching.
 * alpha_init_to_target != 0: the counter starts at the first frame's target (adaptive,
 * src/process_tomatis_adaptive.py:257) instead of 0 (xfade, _xfade.py:171).
 * count_only != 0: only TMT_ARR_C2_COUNT is produced (threshold bisection,
 * src/process_tomatis_adaptive.py:136-152). */
int tmt_plan_gate(tmt_plan* p, int automaton, int gate_input, const double* on, const double* off, int param,
                  int xfade_frames, int alpha_init_to_target, int count_only, void* stream);

/* find_optimal_threshold (src/process_tomatis_adaptive.py:124-154) in one launch: per track, up to max_iter (<= 32; the reference
 * uses 30) bisection steps between t_low and t_high (host arrays of n_tracks doubles: the 5th / 95th percentile of the valid levels),
 * each step a count-only TMT_GATE_MINHOLD scan of TMT_ARR_GATE_F64 with thresholds T_mid +- hyst_db / 2; `start` is the reference's
 * initial best_T (the median), `active[t] == 0` skips the search for a track (no valid level) and keeps start[t].  All scalar
 * arithmetic is float64 in the reference's order.  The thresholds best_T +- hyst_db / 2 are left in the plan's device-side on / off
 * arrays: follow with tmt_plan_gate(..., on = NULL, off = NULL, ...) for the final states.  Results: TMT_ARR_BISECT_*.
 * Tracks of more than 16384 frames: TMT_ERR_UNSUPPORTED (drive tmt_plan_gate(count_only) from the host instead). */
int tmt_plan_bisect(tmt_plan* p, const double* t_low, const double* t_high, const double* start, const int32_t* active,
                    double hyst_db, double target_c2, int hold_frames, int max_iter, void* stream);
/* tmt_plan_bisect followed by the final gate with the thresholds found -- tmt_plan_gate(TMT_GATE_MINHOLD, TMT_ARR_GATE_F64, NULL, NULL,
 * hold_frames, xfade_frames, alpha_init_to_target, 0) -- in the same launch where the search runs on the register-resident path
 * (<= 16 automaton states), as a second launch otherwise: TMT_ARR_STATE, TMT_ARR_ROW and TMT_ARR_C2_COUNT are final on return
 * (simulate_gate + the alpha follower, src/process_tomatis_adaptive.py:226,253-265). */
int tmt_plan_bisect_gate(tmt_plan* p, const double* t_low, const double* t_high, const double* start, const int32_t* active,
                         double hyst_db, double target_c2, int hold_frames, int max_iter, int xfade_frames, int alpha_init_to_target,
                         void* stream);

/* K1+K3+K4 fused: frame gather + Hann window + 4096-pt FFT (stereo packed as L+iR) + gain row +
 * inverse FFT + synthesis window + overlap-add (each output sample written exactly once, no atomics)
 * + normalisation + optional output gain, and max|y| per limiter chunk into TMT_ARR_CHUNK_PEAK
 * (process_available_frames / flush, src/process_tomatis.py:359-426,447-453; _adaptive.py:298-332).
 * post_gain: linear output gain 10^(output_gain_db/20) (src/process_tomatis.py:349-350), 1 = none.
 * skip_edges != 0: leave the single-frame edge blocks to tmt_plan_edge_frames. */
int tmt_plan_stft(tmt_plan* p, float post_gain, int skip_edges, void* stream);

/* Fused K4 + limiter.  Same as tmt_plan_stft(..., skip_edges = 1, ...) followed by tmt_plan_limiter, in one kernel: the
 * CTA that completes the last work unit of a limiter chunk rescales that chunk while the rest of the GPU keeps computing
 * (write_clamped after each flush, src/process_tomatis.py:419-426,331-357).  Order of one pass:
 * tmt_plan_clear_peaks -> tmt_plan_edge_frames -> tmt_plan_stft_limited (the edge blocks and their peaks must already be in
 * place when a chunk is finished).  Chunks this plan produces only partly (time shards) and whole-file chunks (adaptive) are
 * left to a trailing limiter_kernel launch, so call it only when their peaks are final on this rank. */
int tmt_plan_stft_limited(tmt_plan* p, float post_gain, float limit, void* stream);
/* Zero TMT_ARR_CHUNK_PEAK and the fused limiter's counters (tmt_plan_stft does this itself). */
int tmt_plan_clear_peaks(tmt_plan* p, void* stream);

/* fp64 recomputation of the two ill-conditioned edge blocks (first hop of WHOLEFILE framing, tail block
 * of both framings: a single frame divided by w^2 -> 0).  Call after tmt_plan_stft(...,
 * skip_edges = 1, ...) and before tmt_plan_limiter.  in_scale / out_scale (host, per track, may be NULL):
 * the adaptive pre-attenuation x*atten_lin and its restore y*restore_lin, applied in float32 exactly as
 * src/process_tomatis_adaptive.py:215,335-337 (the fp32 body kernel folds them away because they cancel).
 * pipeline_f64: adaptive float64 branch (no float32 roundings around the FFT). */
int tmt_plan_edge_frames(tmt_plan* p, float post_gain, const float* in_scale, const float* out_scale,
                         int pipeline_f64, void* stream);

/* Peak limiter: every chunk whose peak exceeds `limit` is scaled by limit/peak in place
 * (write_clamped, src/process_tomatis.py:352-355; global variant _adaptive.py:341-345). */
int tmt_plan_limiter(tmt_plan* p, float limit, void* stream);

/* tmt_plan_stft (chunk peaks cleared, no limiter) with the fp64 edge frames of tmt_plan_edge_frames running BESIDE the STFT kernel
 * on the engine's side stream (forked from and joined back into `stream`): the two write disjoint blocks.  For jobs that limit in
 * a separate pass anyway (adaptive mode, single short files), where the latency-bound edge kernel is a large part of the call. */
int tmt_plan_stft_with_edges(tmt_plan* p, float post_gain, const float* in_scale, const float* out_scale, int pipeline_f64,
                             void* stream);

/* Convenience: levels (f32) -> gate -> edge frames -> stft with fused limiter on one stream, standard/xfade parameters. */
int tmt_plan_run_streaming(tmt_plan* p, double m_on, double m_off, int run_frames, int xfade_frames,
                           float post_gain, float limit, void* stream);

/* The same with integer PCM input: `pcm` holds the plan's tracks one after the other (track t at pcm + t * track_stride_bytes;
 * int16 or packed 24-bit interleaved stereo, TMT_PCM_*), and ONE pass converts the samples into the plan's float input buffers
 * (soundfile's read, src/process_tomatis.py:434) and sums the hop blocks for the levels (src/process_tomatis.py:370) -- no separate
 * levels pass over the float samples.  tmt_plan_pcm_levels is that pass alone (= tmt_pcm_to_float + tmt_plan_levels with
 * TMT_LEVELS_HOPSUM_ONLY, bit for bit).  Whole-track plans only. */
int tmt_plan_pcm_levels(tmt_plan* p, const void* pcm, int64_t track_stride_bytes, int format, void* stream);
int tmt_plan_run_streaming_pcm(tmt_plan* p, const void* pcm, int64_t track_stride_bytes, int format, double m_on, double m_off,
                               int run_frames, int xfade_frames, float post_gain, float limit, void* stream);

/* ---- PCM edge (device pointers) -------------------------------------------------------------- */
#define TMT_PCM_S16 0 /* int16 little endian                  -> value / 32768    */
#define TMT_PCM_S24 1 /* packed 3-byte little endian (PCM_24)  -> value / 8388608  */
/* Integer PCM -> float32 exactly as soundfile's read(dtype='float32') hands it to the reference
 * (src/process_tomatis.py:434, _adaptive.py:179).  n_values = samples x channels. */
int tmt_pcm_to_float(const void* pcm, int format, int64_t n_values, float* out, void* stream);
/* float32 -> PCM_24 as the reference's output files store it (subtype='PCM_24', src/process_tomatis.py:243,
 * _adaptive.py:351) through python-soundfile, i.e. libsndfile's FLAC conversion with clipping on: lrintf(x * 2^23), pinned to
 * [-2^23, 2^23 - 1] (libsndfile src/flac.c f2flac24_clip_array; a WAV PCM_24 file holds floor(x * 2^23) instead, src/pcm.c). */
int tmt_float_to_pcm(const float* in, int format, int64_t n_values, void* pcm, void* stream);

/* y = dequantise(quantise_PCM24(y)) * scale in place: the PCM_24 write / read round trip of the static EQ's gain-protect
 * pass (src/layer2_apply_eq.py:220-234). */
int tmt_requantise_scale(float* y, int64_t n_values, float scale, void* stream);

/* ---- validators ------------------------------------------------------------ */
/* Conditional spectrum of an input / output file pair: for every listed frame f (positions [f*hop, f*hop + n_fft), which
 * must lie inside the file) ratio_f[k] = mean_c |rfft(y_c * hann)[k]| / max(mean_c |rfft(x_c * hann)[k]|, 1e-10),
 * then median over the frames per bin -> median_out (host, n_fft/2 + 1 floats); the caller takes 20*log10(. + 1e-12).
 * Replaces the frame loops of compute_conditional_spectrum (src/validate_layer1.py:338-374) and, with
 * anchor_bin_lo <= anchor_bin_hi (each frame's ratio divided by its mean over those bins when that mean is > 0), of
 * compute_conditional_spectrum_v2 (src/verify_tomatis_15db_v2.py:307-354); pass anchor_bin_hi < anchor_bin_lo for none.
 * x, y: device, float32 [total][2]; frames: host.  Synchronises the stream before returning. */
int tmt_cond_spectrum(tmt_engine* e, const void* x, const void* y, int64_t total, const int32_t* frames, int n_frames,
                      int anchor_bin_lo, int anchor_bin_hi, float* median_out, void* stream);

/* ---- calibration front end (src/calibrate_to_baseline_v2.py) ---------------- */
/* Envelope for the delay estimate: power_mono(x) (:8-11) resampled like scipy.signal.resample_poly(env, up, down) applies
 * its FIR (upfirdn with zero extension, :60,73), then the mean removed (:61,74).  x: device float32 [n_in][2];
 * h: host, the zero-padded float32 filter resample_poly builds (firwin with a Kaiser(5) window, times up), len_h taps;
 * out[j] = upfirdn(h, env, up, down)[j + n_pre_remove], j < n_out (device).  Synchronises the stream. */
int tmt_calib_envelope_decimate(tmt_engine* e, const void* x, int64_t n_in, const float* h, int len_h, int up, int down,
                                int64_t n_pre_remove, int64_t n_out, float* out, void* stream);
/* corr[k] = sum_j a[k + j] * b[j], k = 0 .. na - nb: fftconvolve(a, b[::-1], mode="valid") (:77-78), computed directly with
 * float32 products and double accumulation.  All pointers device. */
int tmt_calib_xcorr_valid(tmt_engine* e, const float* a, int64_t na, const float* b, int64_t nb, float* corr, void* stream);
/* stft_band_tilt's two band energies (:17-30) for the frames [i*hop, i*hop + n_fft), i < n_frames, of x (device):
 * sum of |rfft(power_mono(frame) * hann)|^2 over bins [lo0, lo1) and [hi0, hi1) -> e_lo / e_hi (host, float32).  The caller
 * finishes 10*log10((e_hi + 1e-12) / (e_lo + 1e-12) + 1e-12).  Frame levels of the same frames: tmt_plan_levels with
 * TMT_LEVELS_POWER_EPS on a TMT_FRAMING_EQ_NOPAD plan.  Synchronises the stream. */
int tmt_calib_band_energies(tmt_engine* e, const void* x, int64_t total, int n_frames, int lo0, int lo1, int hi0, int hi1,
                            float* e_lo, float* e_hi, void* stream);
/* The grid search's inner loop (:253-262): simulate_state (:88-112) over n frames at positions start[] for n_combos
 * parameter sets at once (one thread each): thresholds on/off as float32 (NumPy compares a float32 level with a Python
 * float in float32), up-delay in samples; mismatches against want[] (1 = C1, 2 = C2) and number of state switches per set;
 * states (optional, [n_combos][n]) receives every frame's state.  All pointers host.  Synchronises the stream. */
int tmt_calib_gate_grid(tmt_engine* e, const float* level, const int64_t* start, const uint8_t* want, int n, const float* on,
                        const float* off, const int64_t* delay, int n_combos, int32_t* mismatches, int32_t* switches,
                        uint8_t* states, void* stream);

/* ---- general FFT sizes (verified on hardware in round 2; on by default, TMT_GENERIC_FFT=0 turns it off).  The reference exposes --n_fft / --hop (src/process_tomatis.py:509-510, _adaptive.py:396-397,
 * _xfade.py:388-389); the fused kernels implement 4096 / 2048.  Plain path for power-of-two n_fft in [128, 8192], hop in
 * [1, n_fft]: frame k covers [first_start + k*hop, + n_fft), zeros outside [0, total).  All pointers device.
 * flavour: 0 streaming (standard / xfade), 1 adaptive float32 pipeline, 2 adaptive float64 pipeline. */
/* np.mean(mono*mono) per frame in NumPy's pairwise order (src/process_tomatis.py:370,51; _adaptive.py:74-76); out: float32 or
 * float64 [n_frames]. */
int tmt_generic_meansq(tmt_engine* e, const void* x, int64_t total, int64_t first_start, int n_fft, int hop, int n_frames, int use_f64,
                       float in_scale, int mono_file, void* meansq_out, void* stream);
/* window -> FFT -> gain row rows[k] -> IFFT -> window per frame (src/process_tomatis.py:394-398, _adaptive.py:307-313), double
 * precision with the reference's float32 roundings around it; frames_out: float2 (flavours 0, 1) or double2 (2) [n_frames][n_fft]. */
int tmt_generic_frames(tmt_engine* e, const void* x, int64_t total, int64_t first_start, int n_fft, int hop, int n_frames, const float* win,
                       const float* gains, const uint16_t* rows, float in_scale, int flavour, void* frames_out, void* stream);
/* Overlap-add in the reference's frame order + window-energy normalisation (src/process_tomatis.py:400-406,422;
 * _adaptive.py:316-332), times post_scale (output gain / restored pre-attenuation); y_out: float2 or double2 [total]. */
int tmt_generic_overlap_add(tmt_engine* e, const void* frames, int flavour, int64_t total, int64_t first_start, int n_fft, int hop, int n_frames,
                            const float* win, float post_scale, void* y_out, void* stream);
/* write_clamped's limiter (src/process_tomatis.py:352-355; _adaptive.py:341-345) over the chunks bounds[c] = (s0, s1): peak per
 * chunk into peaks_out (float32 / float64 [n_chunks]), chunks above `limit` scaled by limit / peak in place. */
int tmt_generic_limit(tmt_engine* e, void* y, int use_f64, const int64_t* bounds, int n_chunks, double limit, void* peaks_out, void* stream);
int tmt_generic_to_float(tmt_engine* e, const void* y_f64, int64_t n, void* out_f32, void* stream);

/* ---- one long file over several GPUs: the per-pass exchange through peer memory ------------------------
 * The reference processes a file front to back in one process (src/process_tomatis.py:359-453); time shards need, per pass,
 * every rank's hop-block sums on every rank (for the identical gate scan) and one hop of samples from each neighbour.  Instead
 * of two collectives each rank owns an exchange buffer that the other ranks' processes map through CUDA IPC:
 *   tmt_peer_alloc / tmt_peer_open / tmt_peer_close / tmt_peer_free   buffer of tmt_peer_bytes(n_hop_blocks) bytes + its 64-byte handle;
 *   tmt_plan_peer_publish   after tmt_plan_levels(HOPSUM_ONLY): one kernel stores the rank's own hop sums into every rank's buffer,
 *                           its first / last hop (device pointers, 2048 sample-frames each) into the neighbours' halo slots, and
 *                           raises this rank's flag everywhere (system-scope release).  bases: host array of `world` device pointers
 *                           (entry `rank` = the rank's own buffer);
 *   tmt_plan_peer_wait      one kernel waits for all flags of the rank's own buffer (acquire; gives up after timeout_s and records
 *                           1 + the missing rank in the status word, tmt_peer_status), then copies the assembled hop sums into the
 *                           plan (TMT_ARR_HOPSUM_F32) and the halos into the first `left` / last `right` sample-frames of `window`.
 * Every rank must call publish and wait once per pass, in that order.  One-track plans only. */
size_t tmt_peer_bytes(int n_hop_blocks);
int tmt_peer_alloc(int device, size_t bytes, void** ptr, unsigned char* handle64);
int tmt_peer_open(int device, const unsigned char* handle64, void** ptr);
int tmt_peer_close(int device, void* ptr);
int tmt_peer_free(int device, void* ptr);
int tmt_peer_status(int device, const void* base, int32_t* status);
int tmt_plan_peer_publish(tmt_plan* p, int rank, int world, void* const* bases, const void* first_hop, const void* last_hop, void* stream);
int tmt_plan_peer_wait(tmt_plan* p, int world, void* base, void* window, int64_t window_len, int left, int right, double timeout_s,
                       void* stream);

/* Number of kernel launches issued by this plan since creation (bench.py's gpu_launches). */
int64_t tmt_plan_launch_count(const tmt_plan* p);

#ifdef __cplusplus
}
#endif
#endif /* TOMATIS_B200_H */
