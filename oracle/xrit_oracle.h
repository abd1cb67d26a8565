/*
 * xrit_oracle.h -- CPU oracle for the xritdemod sample-stream hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under xritdemod_b200/ (the product) may
 * include, link or dlopen this.  Allowed users: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (opensatelliteproject/xritdemod @ 630ea9df)
 * keeps the arithmetic of this path in libSatHelper
 * (github.com/opensatelliteproject/libsathelper, cloned un-pinned at build time,
 * reference Makefile:52-59) which is absent from /root/reference, and the
 * reference ships no tests or golden vectors (Makefile:91-92).  This file
 * restates the published algorithms of the GNU Radio 3.7 blocks that the
 * reference itself names as the definition of each stage
 * (demodulator/demod_tcp_qt.py:95-96,261-276) with the constructor arguments
 * of the C++ call sites (demodulator/src/demodulator.cpp:443-450,
 * demodulator/src/Parameters.h:16-37) and the stage order of
 * processSamples() (demodulator.cpp:135-157).
 *
 * All sample buffers are interleaved complex float (I,Q), as
 * std::complex<float> is laid out in the reference.
 */
#ifndef XRIT_ORACLE_H_
#define XRIT_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XO_MMSE_NTAPS 8
#define XO_MMSE_NSTEPS 128

/* ---- tap designers (demodulator.cpp:443-444) ---- */
/* Filters::RRC == firdes.root_raised_cosine; returns ntaps|1, writes that many floats */
int xo_rrc_taps(double gain, double fs, double sym_rate, double alpha, int ntaps, float *taps);
/* Filters::lowPass == firdes.low_pass with a Hamming window */
int xo_lowpass_ntaps(double fs, double transition_width);
int xo_lowpass_taps(double gain, double fs, double cutoff, double transition_width, float *taps);
/* mmse_fir_interpolator table: 129 rows x 8 taps */
void xo_mmse_table(float *table);
/* control_loop gains from loop bandwidth (damping sqrt(2)/2) */
void xo_costas_gains(float loop_bw, float *alpha, float *beta);

/* NCO sin/cos: fully specified FP32 routine (default) or libm (xo_set_libm_sincos(1)) */
void xo_sincosf(float x, float *sn, float *cs);
void xo_sincosf_array(const float *x, int64_t n, float *sn, float *cs);
void xo_set_libm_sincos(int on);
/* cross-check only: FIR tap sums in a 4-lane SIMD order without FMA instead of the serial fmaf order */
void xo_set_fir_simd(int on);

/* ---- stage operators: SatHelper::{FirFilter,AGC,CostasLoop,ClockRecovery} ---- */
typedef struct xo_fir xo_fir;
xo_fir *xo_fir_new(unsigned decimation, const float *taps, int ntaps);
void xo_fir_free(xo_fir *f);
/* n_out output samples; consumes n_out*decimation input samples (demodulator.cpp:137-138) */
void xo_fir_work(xo_fir *f, const float *in, float *out, int n_out);

typedef struct xo_agc xo_agc;
xo_agc *xo_agc_new(float rate, float reference, float gain, float max_gain);
void xo_agc_free(xo_agc *a);
void xo_agc_work(xo_agc *a, const float *in, float *out, int n);
float xo_agc_gain(const xo_agc *a);
void xo_agc_set_gain(xo_agc *a, float g);

typedef struct xo_costas xo_costas;
xo_costas *xo_costas_new(float loop_bw, int order);
void xo_costas_free(xo_costas *c);
void xo_costas_work(xo_costas *c, const float *in, float *out, int n);
void xo_costas_get(const xo_costas *c, float *phase, float *freq);
void xo_costas_set(xo_costas *c, float phase, float freq);

typedef struct xo_mm xo_mm;
typedef struct {
    float mu, omega;
    float p0[2], p1[2], p2[2];
    float c0[2], c1[2], c2[2];
    int64_t next_index; /* absolute stream index of the next interpolation base */
} xo_mm_state;
xo_mm *xo_mm_new(float omega, float gain_omega, float mu, float gain_mu, float omega_rel_limit);
void xo_mm_free(xo_mm *m);
/* returns number of symbols written to out (cf32); chunk-invariant streaming semantics */
int xo_mm_work(xo_mm *m, const float *in, float *out, int n);
void xo_mm_get(const xo_mm *m, xo_mm_state *st);
/* tests only: per-symbol trace of (base index, mu, omega) before and clipped mm of each symbol */
void xo_mm_trace(xo_mm *m, int64_t *ii, float *mu, float *omega, float *mm, int64_t cap);
void xo_mm_set(xo_mm *m, const xo_mm_state *st);

/* ---- the chain: processSamples() (demodulator.cpp:100-168) ---- */
typedef struct {
    uint32_t sample_rate;   /* device sample rate */
    uint32_t symbol_rate;
    uint32_t decimation;    /* baseDecimation; 1 => decimator skipped (demodulator.cpp:136) */
    uint32_t rrc_taps;      /* RRC_TAPS 63 */
    int32_t  loop_order;    /* LOOP_ORDER 2 */
    float rrc_alpha;
    float pll_alpha;        /* Costas loop bandwidth; defaults to CLOCK_ALPHA (demodulator.cpp:220) */
    float clock_alpha;      /* CLOCK_ALPHA 0.0037 (gain_mu; gain_omega = alpha^2/4) */
    float clock_mu;         /* 0.5 */
    float clock_omega_limit;/* 0.005 */
    float agc_rate, agc_ref, agc_gain, agc_max_gain;
} xo_config;

void xo_config_defaults(xo_config *cfg, int hrit);
typedef struct xo_chain xo_chain;
xo_chain *xo_chain_new(const xo_config *cfg);
void xo_chain_free(xo_chain *c);
/* n_complex input samples (must be a multiple of decimation); returns #symbols (cf32) written */
int64_t xo_chain_process(xo_chain *c, const float *iq, int64_t n_complex, float *sym_out, int64_t cap);
/* per-stage taps for inspection: out = 0 costas, 1 agc, 2 rrc, 3 decimated input (NULL to skip) */
int64_t xo_chain_process_tap(xo_chain *c, const float *iq, int64_t n_complex, float *sym_out, int64_t cap,
                             float *dec_out, float *agc_out, float *rrc_out, float *costas_out);
float xo_chain_sps(const xo_chain *c);

/* SymbolManager int8 rule (SymbolManager.cpp:43-46): Re(s)*127, clamp, C cast */
void xo_soft_i8(const float *sym_cf32, int64_t n, int8_t *out);
/* DiagManager's int8 rule (DiagManager.cpp:35-42) for the floats of the diagnostic tap (demodulator.cpp:161-163) */
void xo_diag_i8(const float *v, int64_t n, int8_t *out);
/* onSamplesAvailable conversions (demodulator.cpp:57-70) */
void xo_convert_s16(const int16_t *in, int64_t n_complex, float *out);
void xo_convert_s8(const int8_t *in, int64_t n_complex, float *out);
/* the u8 formats two front ends convert themselves: SpyServerFrontend.cpp:404-407, RtlFrontend.cpp:27,57,104-116 */
void xo_convert_u8(const uint8_t *in, int64_t n_complex, float *out);
float xo_rtl_alpha(uint32_t sample_rate);
void xo_convert_rtl_u8(const uint8_t *in, int64_t n_complex, float alpha, float *avg, float *out);

/* ---- decoder front half (decoder/src/newdecoder.cpp:212-290): see the comments in xrit_oracle.c ---- */
void xo_conv_encode(const uint8_t *bits, int64_t n, unsigned *state, uint8_t *coded);
void xo_nrzm_encode(const uint8_t *bits, int64_t n, uint8_t *last, uint8_t *out);
void xo_nrzm_decode_bytes(uint8_t *data, int64_t n);
void xo_correlate(const uint8_t *data, uint32_t length, const uint64_t *words, int n_words, uint32_t *highest,
                  uint32_t *position, uint32_t *word);
void xo_fix_packet_180(uint8_t *data, int64_t n);
int xo_viterbi27_decode(const uint8_t *soft, int n_bits, int soft_mode, uint8_t *out_bytes);
int64_t xo_decoder_front(const uint8_t *stream, int64_t n, int lrit, int soft_mode, uint8_t *last_end, uint8_t *frames,
                         int32_t *meta, int64_t cap, int64_t *consumed);

#ifdef __cplusplus
}
#endif
#endif
