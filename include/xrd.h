/*
 * xrd.h -- C ABI of the B200-native xritdemod sample-stream hot path.
 *
 * Drop-in boundary for the reference's processSamples() loop
 * (reference demodulator/src/demodulator.cpp:100-168): decimating FIR -> AGC -> RRC FIR ->
 * Costas loop -> Mueller&Mueller clock recovery -> complex soft symbols.  The reference has
 * no FFI layer of its own; its "operator API" is three in-process C++ seams,
 * and every entry point below names the seam it replaces.  Plain pointers and sizes only, no
 * exceptions cross this boundary: functions return 0 (XRD_OK) or a negative xrd_status, and
 * xrd_last_error() gives the message.
 *
 * All sample buffers are interleaved complex float (I,Q) = std::complex<float>, unless a
 * sample `type` says otherwise (FrontendDevice.h:11-13).
 *
 * There is NO CPU fallback: every compute entry point runs hand-written sm_100a CUDA kernels and
 * fails with XRD_E_CUDA when no usable device is present.
 */
#ifndef XRD_H_
#define XRD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    XRD_OK = 0,
    XRD_E_ARG = -1,      /* bad argument */
    XRD_E_CUDA = -2,     /* CUDA runtime / launch failure, or no device */
    XRD_E_NOMEM = -3,
    XRD_E_OVERFLOW = -4, /* output capacity too small / FIFO overflow */
    XRD_E_STATE = -5
} xrd_status;

/* sample types of the upstream seam: FrontendDevice.h:11-13 */
#define XRD_FLOATIQ 0
#define XRD_S16IQ 1
#define XRD_S8IQ 2
/* the two formats the reference's front ends convert themselves before the callback (accepted here so that a
 * front end can hand over its raw buffer):
 *   XRD_U8IQ    unsigned 8-bit, (v - 128) / 128.f          SpyServerFrontend.cpp:396-434 (:406)
 *   XRD_RTLU8IQ RTL-SDR: lut[v] = (v - 128) / 127.f, then the one-pole DC blocker of RtlFrontend.cpp:104-116
 *               (alpha from the sample rate, :57; its `i % 1` quirk -- one average for I and Q -- is kept).
 *               The blocker is a serial filter with state: xrd_add_samples only. */
#define XRD_U8IQ 3
#define XRD_RTLU8IQ 4

/* ---------------------------------------------------------------------------------------
 * Demodulator: replaces the globals + processSamples() of demodulator.cpp:31-52,100-168
 * ------------------------------------------------------------------------------------- */
typedef struct xrd_demod xrd_demod;

/* Field defaults (xrd_config_defaults) are the reference's Parameters.h:16-37 and the
 * parameter derivation of demodulator.cpp:220,436-450. */
typedef struct {
    uint32_t sample_rate;        /* device sample rate (Hz)                                  */
    uint32_t symbol_rate;        /* HRIT 927000 / LRIT 293883   Parameters.h:18,23           */
    uint32_t decimation;         /* baseDecimation; 1 = decimator skipped (demodulator.cpp:136) */
    uint32_t rrc_taps;           /* RRC_TAPS 63                                              */
    int32_t loop_order;          /* LOOP_ORDER 2 (only 2 is implemented: BPSK)               */
    float rrc_alpha;             /* 0.3 HRIT / 0.5 LRIT                                      */
    float pll_alpha;             /* Costas loop bandwidth; default CLOCK_ALPHA (demodulator.cpp:220) */
    float clock_alpha;           /* CLOCK_ALPHA 0.0037: gain_mu; gain_omega = alpha^2/4      */
    float clock_mu;              /* CLOCK_MU 0.5                                             */
    float clock_omega_limit;     /* CLOCK_OMEGA_LIMIT 0.005                                  */
    float agc_rate, agc_ref, agc_gain, agc_max_gain; /* 0.01, 0.5, 1.0, 4000                 */
    int32_t device_ordinal;      /* CUDA device                                              */
    int32_t n_channels;          /* independent IQ streams with these parameters (>= 1)      */
} xrd_config;

/* hrit != 0: setHRITMode (demodulator.cpp:188-197); else setLRITMode (:177-186) */
void xrd_config_defaults(xrd_config *cfg, int hrit);

int xrd_create(const xrd_config *cfg, xrd_demod **out);
void xrd_destroy(xrd_demod *d);
const char *xrd_last_error(const xrd_demod *d); /* d may be NULL: error of the last failed create */

/* == onSamplesAvailable(void *data, int length, int type)   demodulator.cpp:54-74
 * Copies `n_complex` samples of `type` into the channel's host FIFO (capacity FIFO_SIZE,
 * Parameters.h:57); the caller's buffer is borrowed only for the call.  Thread-safe against
 * xrd_process.  Returns XRD_E_OVERFLOW (and drops the samples) when the FIFO is full. */
int xrd_add_samples(xrd_demod *d, int channel, const void *data, int n_complex, int type);

/* symbol sink == SymbolManager::add(std::complex<float>*, int)   SymbolManager.cpp:94-107 */
typedef void (*xrd_symbols_cb)(void *user, int channel, const float *cf32_symbols, int n_symbols);

/* == processSamples()   demodulator.cpp:100-168
 * Runs the chain on everything queued (if at least min_samples complex samples are queued per
 * channel; the reference's threshold is 32768, demodulator.cpp:113) and hands the symbols to
 * `cb`.  Returns the number of complex samples consumed per channel, 0 if below threshold.
 * With decimation > 1 the n % decimation trailing samples stay queued for the next call; the reference pops and
 * drops them (demodulator.cpp:124-128,137), which shifts its decimation phase after every chunk whose length is not
 * a multiple of the decimation -- a defect this path does not reproduce.
 * On an error (< 0) nothing has been dequeued and `cb` has not run, but the loop state may have advanced:
 * xrd_reset (or xrd_checkpoint_load) before feeding the same samples again. */
int64_t xrd_process(xrd_demod *d, int64_t min_samples, xrd_symbols_cb cb, void *user);

/* One-shot over HOST buffers (state carried across calls, like consecutive processSamples()
 * calls): iq holds n_channels blocks of n_complex samples of `type` (channel-major);
 * symbols (cf32) are written to sym_out[channel * cap ...], counts to n_sym[channel].
 * n_complex must be a multiple of the decimation (the reference silently drops the remainder,
 * demodulator.cpp:137; here it is an error).  `type`: XRD_FLOATIQ, XRD_S16IQ, XRD_S8IQ or XRD_U8IQ. */
int xrd_demod_batch(xrd_demod *d, const void *iq, size_t n_complex, int type, float *sym_out, size_t cap,
                    int64_t *n_sym);

/* Same, device-resident: iq_dev / sym_dev are device pointers on cfg.device_ordinal; n_sym is a
 * host array.  This is the call the throughput metric is quoted on. */
int xrd_demod_device(xrd_demod *d, const void *iq_dev, size_t n_complex, int type, float *sym_dev, size_t cap,
                     int64_t *n_sym);

/* xrd_demod_batch with the egress of SymbolManager::process (SymbolManager.cpp:37-52) fused into the last kernel:
 * soft_out[channel * cap ...] receives one int8 soft symbol per recovered symbol (Re(s)*127, clamp [-128,127],
 * C cast) -- the byte stream the reference sends to the decoder (decoder/src/newdecoder.cpp:213-216 reads it in
 * 16384-byte frames).  No cf32 symbols cross PCIe: 1 byte per symbol instead of 8. */
int xrd_demod_batch_i8(xrd_demod *d, const void *iq, size_t n_complex, int type, int8_t *soft_out, size_t cap,
                       int64_t *n_sym);

/* int8 soft symbols == SymbolManager::process   SymbolManager.cpp:43-46
 * (Re(s)*127, clamp [-128,127], C cast).  Host buffers. */
int xrd_soft_i8(xrd_demod *d, const float *sym_cf32, size_t n_symbols, int8_t *out);

/* Loop state of one channel: what the five libSatHelper objects hold between Work() calls.
 * Doubles as checkpoint (the reference never serialises it). */
typedef struct {
    float agc_gain;
    float costas_phase, costas_freq;
    float mm_mu, mm_omega;
    float mm_p0[2], mm_p1[2];   /* interpolants of the last two symbols */
    int64_t mm_next;            /* next interpolation base relative to the end of consumed input (<0: in retained tail) */
    uint64_t n_in, n_sym;       /* totals since create / last set */
} xrd_loop_state;
int xrd_get_state(xrd_demod *d, int channel, xrd_loop_state *st);
/* Puts the loop variables of one channel back (the reference's state lives in the five operator objects,
 * demodulator.cpp:32-36).  Filter histories and the M&M sample tail are not part of xrd_loop_state: use the
 * checkpoint calls below to continue a stream exactly. */
int xrd_set_state(xrd_demod *d, int channel, const xrd_loop_state *st);

/* Checkpoint / resume of the whole demodulator (all channels): loop variables, decimator and RRC histories,
 * the M&M sample tail, totals.  A demodulator created with the same xrd_config that loads the blob continues the
 * stream bit for bit where the saved one stopped (queued FIFO samples are input, not state, and are not saved).
 * xrd_checkpoint_load returns XRD_E_STATE when the blob was saved under another configuration. */
size_t xrd_checkpoint_size(const xrd_demod *d);
int xrd_checkpoint_save(xrd_demod *d, void *blob, size_t cap);
int xrd_checkpoint_load(xrd_demod *d, const void *blob, size_t bytes);

/* Upper bound of the symbols one call over n_complex input samples per channel can produce (the capacity
 * xrd_demod_batch / xrd_demod_device / xrd_demod_batch_i8 never overflow with); derived from omega, its
 * relative limit and gain_mu exactly as the M&M stage sizes its own staging. */
int64_t xrd_symbol_capacity(const xrd_demod *d, size_t n_complex);

/* Diagnostics of the last chain call on one channel, produced on the device by the pass that writes the symbols.
 *   frame / n_frame  what processSamples() hands to DiagManager::addSamples after every chunk -- the first
 *                    min(symbols, 1024) FLOATS of the interleaved complex symbol buffer (demodulator.cpp:161-163) --
 *                    in the int8 form DiagManager puts on its UDP socket: v * 128, clamp [-128, 127], C cast
 *                    (DiagManager.cpp:31-47).  A caller sends `frame` as is once 1024 bytes have accumulated.
 *   mean_*           E|I|, E[I^2], E[Q^2] over all symbols of the call
 *   snr_db           10 log10(E|I|^2 / (E[I^2] - E|I|^2)): signal to noise on the decision axis (the reference's
 *                    GNU Radio prototype displays an RMS-ratio SNR, demod_tcp_qt.py:263-298; the C++ program none)
 *   lock             E[I^2] / (E[I^2] + E[Q^2]): -> 1 when the Costas loop holds the BPSK constellation on I,
 *                    0.5 without carrier lock */
typedef struct {
    int32_t n_frame;
    int8_t frame[1024];
    uint64_t n_symbols;
    double mean_abs_i, mean_sq_i, mean_sq_q;
    float snr_db, lock;
} xrd_diag;
int xrd_get_diag(xrd_demod *d, int channel, xrd_diag *out);

/* Back to the just-created state (loop states, filter histories, queued samples, totals):
 * what deleting and re-constructing the five operators does in the reference
 * (demodulator.cpp:446-450, 503-523), without giving device buffers back. */
int xrd_reset(xrd_demod *d);

/* The cudaStream_t every kernel and copy of this demodulator is issued on (for callers that
 * time with CUDA events or order their own device work against it). */
void *xrd_stream(xrd_demod *d);

/* Segmentation and kernel choice of the time-parallel loops; 0 keeps the default everywhere.
 * Results do not depend on these values -- hand-offs are certified bitwise -- only speed does
 * (tests force small segments and every kernel shape through this). */
typedef struct {
    int32_t agc_seg, agc_warm;         /* AGC segment / speculative warm-up (samples) */
    int32_t costas_seg, costas_warm;   /* Costas segment / warm-up (samples) */
    int64_t mm_seg, mm_warm;           /* M&M segment / warm-up (samples) */
    int32_t mm_lanes;                  /* lanes of the M&M chain kernel: 128, 256, 512 or 1024 (default) */
    int32_t mm_kernel;                 /* 1: 32-bit fixed-point chain kernel where valid (default); 2: generic 64-bit kernel */
    int32_t mm_rerun;                  /* certified M&M re-runs: 1 walk relative to the recorded trajectory (default),
                                          2 chain-kernel re-runs */
    int32_t mm_walk_lanes;             /* lanes of that walk: 128, 256 or 512 (default) */
    int32_t loop_kernel;               /* AGC/Costas first pass: 1 one thread per segment, 2 window-Newton warp chains
                                          (default), 3..7 window-Newton CTA chains with (slots per thread, warps) =
                                          (1,4) (2,4) (1,2) (2,2) (2,8) */
    int32_t rerun_kernel;              /* AGC/Costas certified re-runs: 2..7 as above (default 4) */
    int32_t h2d_pieces;                /* host-input calls: copy/compute pieces, 1..16 (default 2) */
    int32_t h2d_piece_min_ki;          /* minimum piece, Ki samples (default 16 M samples) */
    int32_t costas_chains_per_sm, agc_chains_per_sm;   /* segments per SM of the first pass (defaults 8, 16) */
    int32_t agc_kernel;                /* AGC first pass only, overriding loop_kernel (same values) */
    int32_t chase;                     /* 1 (default): a certified AGC/Costas re-run that has not merged at the end of
                                          its segment continues into the next one; 2: stops there (one more round) */
    int32_t guided;                    /* 1 (default): the Costas first pass records its trajectory (the state before every
                                          sample, 8 bytes per sample of device memory) and certified re-runs take their
                                          proposals from it; 2: re-runs extrapolate as the first pass does */
} xrd_tuning;
int xrd_set_tuning(xrd_demod *d, const xrd_tuning *t);

/* counters since create (diagnostics; bench.py reports gpu_launches from here) */
typedef struct {
    uint64_t kernel_launches;
    uint64_t agc_rounds, costas_rounds, mm_rounds;       /* fix-up rounds run */
    uint64_t agc_redo, costas_redo, mm_redo;             /* segments re-run */
    uint64_t mm_windows, mm_iters;                       /* fixed-point windows / iterations */
    uint64_t agc_iters, costas_iters;                    /* window-Newton iterations (all warps) */
    float ms_fir_dec, ms_agc, ms_fir_rrc, ms_costas, ms_mm;  /* device time of the last call (CUDA events) */
    uint64_t mm_bail;                                    /* relative (delta) re-runs that fell back to the chain kernel */
} xrd_stats;
int xrd_get_stats(xrd_demod *d, xrd_stats *s);

/* ---------------------------------------------------------------------------------------
 * Tap designers: SatHelper::Filters::RRC / lowPass   demodulator.cpp:443-444
 * ------------------------------------------------------------------------------------- */
/* returns the number of taps written (ntaps | 1), or < 0 */
int xrd_design_rrc(double gain, double sample_rate, double symbol_rate, double alpha, int ntaps, float *taps, int cap);
/* Hamming-window low-pass; returns the number of taps written, or < 0 (cap too small: -needed) */
int xrd_design_lowpass(double gain, double sample_rate, double cutoff, double transition_width, float *taps, int cap);
/* MMSE interpolator table of ClockRecovery: 129 rows x 8 taps */
void xrd_mmse_table(float *table129x8);
/* Costas loop gains from the loop bandwidth (damping sqrt(2)/2) */
void xrd_costas_gains(float loop_bw, float *alpha, float *beta);

/* ---------------------------------------------------------------------------------------
 * Stage operators: SatHelper::{FirFilter,AGC,CostasLoop,ClockRecovery}
 * ctor arguments as in demodulator.cpp:446-450; work == Work(in, out, length) on HOST buffers,
 * state carried across calls.  Each work() is one GPU stage (H2D, kernels, D2H).
 * ------------------------------------------------------------------------------------- */
typedef struct xrd_stage xrd_stage;
int xrd_fir_create(int device, unsigned decimation, const float *taps, int ntaps, xrd_stage **out);
int xrd_agc_create(int device, float rate, float reference, float gain, float max_gain, xrd_stage **out);
int xrd_costas_create(int device, float loop_bw, int order, xrd_stage **out);
int xrd_clock_recovery_create(int device, float omega, float gain_omega, float mu, float gain_mu,
                              float omega_rel_limit, xrd_stage **out);
/* FirFilter: length = OUTPUT count, consumes length*decimation inputs (demodulator.cpp:137-138).
 * ClockRecovery: returns the number of symbols written to out (>= 0).  Others return 0. */
int xrd_stage_work(xrd_stage *s, const float *in, float *out, int length);
/* segmentation override for a stage (samples); 0 keeps defaults */
int xrd_stage_set_tuning(xrd_stage *s, int64_t seg, int64_t warm);
/* AGC / Costas kernel choice as xrd_tuning.loop_kernel (tests; results never depend on it) */
int xrd_stage_set_loop_kernel(xrd_stage *s, int kernel);
void xrd_stage_destroy(xrd_stage *s);
const char *xrd_stage_last_error(const xrd_stage *s);

/* ---------------------------------------------------------------------------------------
 * Decoder front half: what decoder/src/newdecoder.cpp:212-290 does first with the soft-symbol
 * byte stream of this path (xrd_demod_batch_i8) -- SatHelper::Correlator (:76, :218-247), frame alignment (:250-264),
 * PacketFixer for the 180 degree ambiguity (:268-270), Viterbi27 with the last 64 soft bytes of the previous frame in
 * front (:273-296), NRZ-M decoding for HRIT (:283-285).  Integer work, bit-exact against the oracle.  Everything
 * behind it (de-randomiser, Reed-Solomon, channel demux) stays in the decoder.
 * ------------------------------------------------------------------------------------- */
typedef struct xrd_decoder_front xrd_decoder_front;
/* lrit != 0: LRIT sync words and phase fix; else HRIT sync words and NRZ-M decoding (newdecoder.cpp:147-153).
 * soft_mode: how the Viterbi metric reads a soft byte.  0 = raw, as the reference call chain hands it over
 * (newdecoder.cpp:215-216,281 pass SymbolManager's int8 bytes to Viterbi27::decode untouched: the sign of every symbol
 * is read correctly, its confidence mirrored within each half); 1 = as the signed symbol it is (+127 surest coded 0,
 * -128 surest coded 1) -- for the case that libSatHelper converts internally, which could not be checked here. */
#define XRD_SOFT_RAW 0
#define XRD_SOFT_SIGNED 1
int xrd_decoder_front_create(int device, int lrit, int soft_mode, xrd_decoder_front **out);
void xrd_decoder_front_destroy(xrd_decoder_front *f);
int xrd_decoder_front_reset(xrd_decoder_front *f);   /* forget the previous frame's tail (lastFrameEnd := 128) */

/* == Correlator::correlate(data, length) followed by getHighestCorrelation / getHighestCorrelationPosition /
 * getCorrelationWordNumber (newdecoder.cpp:224,238-240): the first position with the strictly highest number of hard
 * decisions that agree with one of the two sync words (0 and 180 degrees). */
int xrd_correlate(xrd_decoder_front *f, const uint8_t *data, uint32_t length, uint32_t *highest, uint32_t *position,
                  uint32_t *word);

typedef struct {
    int64_t offset;        /* of the frame's first soft byte in the stream handed to the call */
    int32_t correlation;   /* sync-word agreement, >= MINCORRELATIONBITS (46) */
    int32_t word;          /* 0: 0 degrees, 1: 180 degrees (phaseShift, newdecoder.cpp:241) */
    int32_t bit_errors;    /* Viterbi27::GetBER: coded bits the decoder corrected */
    int32_t reserved;
} xrd_frame_meta;

/* The loop body of newdecoder.cpp:212-300 over a buffered stretch of the stream (`soft`: n bytes as SymbolManager sends
 * them): chunks of 16384 bytes, dropped when the best correlation is below 46, re-aligned on the sync word otherwise;
 * every aligned frame is phase-fixed (LRIT), Viterbi-decoded and NRZ-M decoded (HRIT) into 1024 bytes (the decoder's
 * `vitdecData`: sync marker first).  frames_out: cap x 1024 bytes.  *consumed: bytes of the stream that are done with
 * (feed the rest again, in front of what arrives next).  The flywheel shortcut of :222-236 -- searching only the first
 * 1024 bytes while locked -- is a CPU economy with the same result and is not reproduced. */
int xrd_decoder_front_run(xrd_decoder_front *f, const int8_t *soft, size_t n, uint8_t *frames_out, xrd_frame_meta *meta_out,
                          size_t cap, size_t *n_frames, size_t *consumed);

/* library / device probe: 0 if a usable sm_100 device is present */
int xrd_device_check(int device, char *name, int name_cap, int *sm_count, int *cc_major, int *cc_minor);
const char *xrd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* XRD_H_ */
