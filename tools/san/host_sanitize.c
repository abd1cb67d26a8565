/* Host-side sanitizer run (ASan + UBSan) over the CPU code of this repository: the oracle (oracle/xrit_oracle.c)
 * and the synthetic IQ source (xritdemod_b200/csrc/siggen.c), compiled into this driver with
 *   gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off ...  (tools/san/run.sh)
 * It drives every public entry point of the oracle on generated signals: tap designers, the five stage operators in
 * ragged calls, the chain (plain, decimated, S16 / S8 / U8 / RTL ingest), int8 rules, and the decoder front half on the
 * chain's own soft symbols and on an encoded frame stream.  Exit code 0 and no sanitizer report = clean. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/xrit_oracle.h"

typedef struct {
    double sample_rate, symbol_rate, rrc_alpha, timing_offset, carrier_hz, phase0, amp_start, amp_end;
    uint64_t ramp_len;
    double esn0_db;
    int32_t noise, reserved;
    uint64_t seed;
} xrd_sig_params;
void xrd_siggen_cf32(const xrd_sig_params *p, uint64_t start, uint64_t n, float *out);
void xrd_siggen_bits(uint64_t seed, int64_t k_start, int64_t n, int8_t *out);
void xrd_cf32_to_s16(const float *in, uint64_t n_complex, int16_t *out);
void xrd_cf32_to_u8(const float *in, uint64_t n_complex, uint8_t *out);
void xrd_cf32_to_s8(const float *in, uint64_t n_complex, int8_t *out);
void xrd_siggen_set_threads(int n);

static void *xmalloc(size_t n)
{
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "out of memory\n"); exit(2); }
    return p;
}

int main(void)
{
    const int64_t N = 400000;
    xrd_siggen_set_threads(2);
    xrd_sig_params p = {2.5e6, 927000.0, 0.3, 0.37, -48.0, 0.4, 0.2, 0.5, 100000, 12.0, 1, 0, 77};
    float *x = xmalloc(sizeof(float) * 2 * N);
    xrd_siggen_cf32(&p, 0, (uint64_t)N, x);

    /* designers (arguments as xo_chain_new derives them from the config, so that the stage calls below can be compared
     * with the chain bit for bit) */
    xo_config cfg;
    xo_config_defaults(&cfg, 1);
    const float circuit_rate = (float)cfg.sample_rate / 1.0f;
    const float sps = circuit_rate / (float)cfg.symbol_rate;
    float taps[512];
    int nt = xo_rrc_taps(1, circuit_rate, cfg.symbol_rate, cfg.rrc_alpha, (int)cfg.rrc_taps, taps);
    int nl = xo_lowpass_ntaps(10e6, 927000.0 / 2);
    float *lp = xmalloc(sizeof(float) * (size_t)(nl > 0 ? nl : 1));
    xo_lowpass_taps(1.0, 10e6, 927000.0, 927000.0 / 2, lp);
    float table[129 * 8];
    xo_mmse_table(table);
    printf("rrc taps %d, low-pass taps %d, mmse[64][3] %.6f\n", nt, nl, table[64 * 8 + 3]);

    /* stage operators in ragged calls */
    float *a = xmalloc(sizeof(float) * 2 * N), *b = xmalloc(sizeof(float) * 2 * N), *c = xmalloc(sizeof(float) * 2 * N);
    float *s = xmalloc(sizeof(float) * 2 * N);
    xo_agc *agc = xo_agc_new(cfg.agc_rate, cfg.agc_ref, cfg.agc_gain, cfg.agc_max_gain);
    xo_fir *fir = xo_fir_new(1, taps, nt);
    xo_costas *cos_ = xo_costas_new(cfg.pll_alpha, cfg.loop_order);
    xo_mm *mm = xo_mm_new(sps, (cfg.clock_alpha * cfg.clock_alpha) / 4.0f, cfg.clock_mu, cfg.clock_alpha, cfg.clock_omega_limit);
    const int64_t cuts[] = {0, 1, 17, 4096, 70001, 262144, 262145, N};
    int64_t nsym = 0;
    for (int i = 0; i + 1 < (int)(sizeof cuts / sizeof cuts[0]); i++) {
        const int64_t o = cuts[i], m = cuts[i + 1] - cuts[i];
        xo_agc_work(agc, x + 2 * o, a + 2 * o, (int)m);
        xo_fir_work(fir, a + 2 * o, b + 2 * o, (int)m);
        xo_costas_work(cos_, b + 2 * o, c + 2 * o, (int)m);
        nsym += xo_mm_work(mm, c + 2 * o, s + 2 * nsym, (int)m);
    }
    printf("stage operators: %lld symbols, agc gain %.4f\n", (long long)nsym, xo_agc_gain(agc));
    xo_agc_free(agc); xo_fir_free(fir); xo_costas_free(cos_); xo_mm_free(mm);

    /* chain, one call vs the ragged stage calls */
    xo_chain *ch = xo_chain_new(&cfg);
    float *sym = xmalloc(sizeof(float) * 2 * N);
    int64_t n1 = xo_chain_process(ch, x, N, sym, N);
    const int same = (n1 == nsym && memcmp(sym, s, sizeof(float) * 2 * (size_t)n1) == 0);
    printf("chain: %lld symbols, %s the stage calls\n", (long long)n1, same ? "bit-equal to" : "DIFFERENT from");
    if (!same) return 1;
    xo_chain_free(ch);
    xo_set_libm_sincos(1); xo_set_fir_simd(1);
    ch = xo_chain_new(&cfg);
    printf("chain with libm sincos + SIMD-order FIR: %lld symbols\n", (long long)xo_chain_process(ch, x, N, sym, N));
    xo_chain_free(ch);
    xo_set_libm_sincos(0); xo_set_fir_simd(0);

    /* integer ingest */
    int16_t *x16 = xmalloc(sizeof(int16_t) * 2 * N);
    int8_t *x8 = xmalloc(2 * N);
    uint8_t *xu = xmalloc(2 * N);
    float *xf = xmalloc(sizeof(float) * 2 * N);
    xrd_cf32_to_s16(x, (uint64_t)N, x16); xrd_cf32_to_s8(x, (uint64_t)N, x8); xrd_cf32_to_u8(x, (uint64_t)N, xu);
    xo_convert_s16(x16, N, xf);
    ch = xo_chain_new(&cfg); int64_t k16 = xo_chain_process(ch, xf, N, sym, N); xo_chain_free(ch);
    xo_convert_s8(x8, N, xf);
    ch = xo_chain_new(&cfg); int64_t k8 = xo_chain_process(ch, xf, N, sym, N); xo_chain_free(ch);
    xo_convert_u8(xu, N, xf);
    ch = xo_chain_new(&cfg); int64_t ku = xo_chain_process(ch, xf, N, sym, N); xo_chain_free(ch);
    float avg = 0.f;
    xo_convert_rtl_u8(xu, N, xo_rtl_alpha(2500000), &avg, xf);
    ch = xo_chain_new(&cfg); int64_t kr = xo_chain_process(ch, xf, N, sym, N); xo_chain_free(ch);
    printf("ingest s16 %lld s8 %lld u8 %lld rtl %lld symbols\n", (long long)k16, (long long)k8, (long long)ku, (long long)kr);

    /* decimated chain (10 Msps, D = 4) */
    xrd_sig_params p10 = p;
    p10.sample_rate = 10e6;
    xrd_siggen_cf32(&p10, 0, (uint64_t)N, x);
    xo_config c4 = cfg;
    c4.sample_rate = 10000000; c4.decimation = 4;
    ch = xo_chain_new(&c4);
    printf("decimated chain: %lld symbols\n", (long long)xo_chain_process(ch, x, N, sym, N));
    xo_chain_free(ch);

    /* int8 rules + decoder front half on the chain's soft symbols */
    xrd_siggen_cf32(&p, 0, (uint64_t)N, x);
    ch = xo_chain_new(&cfg);
    n1 = xo_chain_process(ch, x, N, sym, N);
    xo_chain_free(ch);
    int8_t *soft = xmalloc((size_t)n1 + 16);
    xo_soft_i8(sym, n1, soft);
    int8_t dg[1024];
    xo_diag_i8(sym, 1024, dg);
    uint8_t last_end[64];
    memset(last_end, 128, sizeof last_end);
    const int64_t cap = 16;
    uint8_t *frames = xmalloc((size_t)cap * 1024);
    int32_t meta[4 * 16];
    int64_t consumed = 0;
    int64_t nf = xo_decoder_front((const uint8_t *)soft, n1, 0, 1, last_end, frames, meta, cap, &consumed);
    printf("decoder front on noise-like payload: %lld frames, consumed %lld of %lld\n", (long long)nf, (long long)consumed, (long long)n1);

    /* ... and on a properly framed, encoded stream: sync marker + payload, r = 1/2 k = 7, hard symbols */
    const int frames_n = 5, fbits = 8192;
    uint8_t *bits = xmalloc((size_t)frames_n * fbits);
    int8_t *rb = xmalloc((size_t)frames_n * fbits);
    xrd_siggen_bits(5, 0, (int64_t)frames_n * fbits, rb);
    const uint32_t asm_ = 0x1ACFFC1D;
    for (int f = 0; f < frames_n; f++)
        for (int i = 0; i < fbits; i++)
            bits[f * fbits + i] = (i < 32) ? (uint8_t)((asm_ >> (31 - i)) & 1) : (uint8_t)(rb[f * fbits + i] > 0);
    uint8_t *coded = xmalloc((size_t)frames_n * fbits * 2);
    unsigned st = 0;
    xo_conv_encode(bits, (int64_t)frames_n * fbits, &st, coded);
    uint8_t *stream = xmalloc((size_t)frames_n * fbits * 2);
    for (int64_t i = 0; i < (int64_t)frames_n * fbits * 2; i++) stream[i] = coded[i] ? 0x9C : 0x64;   /* -100 / +100 as int8 */
    memset(last_end, 128, sizeof last_end);
    nf = xo_decoder_front(stream, (int64_t)frames_n * fbits * 2, 1, 1, last_end, frames, meta, cap, &consumed);
    uint32_t hi = 0, pos = 0, word = 0;
    const uint64_t uw[2] = {0xfca2b63db00d9794ULL, 0x035d49c24ff2686bULL};
    xo_correlate(stream, 40000, uw, 2, &hi, &pos, &word);
    printf("decoder front on an encoded stream: %lld frames (bit errors of the first: %d), correlate %u at %u\n",
           (long long)nf, nf > 0 ? meta[3] : -1, hi, pos);
    uint8_t tmp[64];
    memcpy(tmp, frames, 64);
    xo_nrzm_decode_bytes(tmp, 64);
    xo_fix_packet_180(tmp, 64);

    free(x); free(lp); free(a); free(b); free(c); free(s); free(sym); free(x16); free(x8); free(xu); free(xf); free(soft);
    free(frames); free(bits); free(rb); free(coded); free(stream);
    printf("host sanitizer run complete\n");
    return 0;
}
