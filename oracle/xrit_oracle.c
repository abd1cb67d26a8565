/*
 * xrit_oracle.c -- scalar FP32 CPU restatement of the xritdemod demodulator hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see xrit_oracle.h).  PARITY UNPINNED: restated from the
 * GNU Radio 3.7 block semantics the reference binds each stage to
 * (demodulator/demod_tcp_qt.py:95-96,261-276); libSatHelper, which holds the
 * reference's own arithmetic, is not vendored (reference Makefile:52-59).
 *
 * Build with -ffp-contract=off: every fused multiply-add below is an explicit fmaf()
 * so the summation order is part of the definition (the CUDA path mirrors it).
 */
#include "xrit_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define XO_TWOPI (2.0 * M_PI)

/* ------------------------------------------------------------------------- */
/* Tap designers                                                             */
/* ------------------------------------------------------------------------- */

/* Filters::RRC(gain, fs, symRate, alpha, ntaps) -- call site demodulator.cpp:443;
 * algorithm = firdes.root_raised_cosine (demod_tcp_qt.py:95-96). */
int xo_rrc_taps(double gain, double fs, double sym_rate, double alpha, int ntaps, float *taps)
{
    ntaps |= 1;
    double spb = fs / sym_rate;
    double scale = 0.0;
    for (int i = 0; i < ntaps; i++) {
        double x1, x2, x3, num, den;
        double xindx = i - ntaps / 2;
        x1 = M_PI * xindx / spb;
        x2 = 4 * alpha * xindx / spb;
        x3 = x2 * x2 - 1;
        if (fabs(x3) >= 0.000001) {
            if (i != ntaps / 2)
                num = cos((1 + alpha) * x1) + sin((1 - alpha) * x1) / (4 * alpha * xindx / spb);
            else
                num = cos((1 + alpha) * x1) + (1 - alpha) * M_PI / (4 * alpha);
            den = x3 * M_PI;
        } else {
            if (alpha == 1) {
                taps[i] = -1;
                scale += taps[i];
                continue;
            }
            x3 = (1 - alpha) * x1;
            x2 = (1 + alpha) * x1;
            num = (sin(x2) * (1 + alpha) * M_PI
                   - cos(x3) * ((1 - alpha) * M_PI * spb) / (4 * alpha * xindx)
                   + sin(x3) * spb * spb / (4 * alpha * xindx * xindx));
            den = -32 * M_PI * alpha * alpha * xindx / spb;
        }
        taps[i] = (float)(4 * alpha * num / den);
        scale += taps[i];
    }
    for (int i = 0; i < ntaps; i++)
        taps[i] = (float)(taps[i] * gain / scale);
    return ntaps;
}

/* firdes.compute_ntaps for WIN_HAMMING (max attenuation 53 dB) */
int xo_lowpass_ntaps(double fs, double transition_width)
{
    int ntaps = (int)(53.0 * fs / (22.0 * transition_width));
    if ((ntaps & 1) == 0)
        ntaps++;
    return ntaps;
}

/* Filters::lowPass(gain, fs, cutoff, tw, HAMMING, beta) -- call site demodulator.cpp:444;
 * algorithm = firdes.low_pass (demod_tcp_qt.py:261-262). */
int xo_lowpass_taps(double gain, double fs, double cutoff, double transition_width, float *taps)
{
    int ntaps = xo_lowpass_ntaps(fs, transition_width);
    int M = (ntaps - 1) / 2;
    double fwT0 = 2 * M_PI * cutoff / fs;
    for (int n = -M; n <= M; n++) {
        float w = (float)(0.54 - 0.46 * cos((2 * M_PI * (n + M)) / (ntaps - 1)));
        if (n == 0)
            taps[n + M] = (float)(fwT0 / M_PI * w);
        else
            taps[n + M] = (float)(sin(n * fwT0) / (n * M_PI) * w);
    }
    double fmax = taps[0 + M];
    for (int n = 1; n <= M; n++)
        fmax += 2 * taps[n + M];
    gain /= fmax;
    for (int i = 0; i < ntaps; i++)
        taps[i] = (float)(taps[i] * gain);
    return ntaps;
}

static double xo_sinc(double x)
{
    if (fabs(x) < 1e-12)
        return 1.0;
    return sin(M_PI * x) / (M_PI * x);
}

/* mmse_fir_interpolator_cc tap table (GNU Radio interpolator_taps.h): the 8-tap
 * least-squares fractional-delay filter over |f| <= 0.25, NSTEPS = 128, printed
 * with %.5e upstream. */
void xo_mmse_table(float *table)
{
    const double B = 0.25;
    const int N = XO_MMSE_NTAPS;
    for (int k = 0; k <= XO_MMSE_NSTEPS; k++) {
        double mu = (double)k / XO_MMSE_NSTEPS;
        double A[8][9];
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++)
                A[i][j] = xo_sinc(2 * B * ((i - 4) - (j - 4)));
            A[i][N] = xo_sinc(2 * B * ((i - 4) + mu));
        }
        /* Gaussian elimination with partial pivoting */
        for (int c = 0; c < N; c++) {
            int piv = c;
            for (int r = c + 1; r < N; r++)
                if (fabs(A[r][c]) > fabs(A[piv][c]))
                    piv = r;
            if (piv != c)
                for (int j = 0; j <= N; j++) {
                    double t = A[c][j];
                    A[c][j] = A[piv][j];
                    A[piv][j] = t;
                }
            for (int r = 0; r < N; r++) {
                if (r == c)
                    continue;
                double f = A[r][c] / A[c][c];
                for (int j = c; j <= N; j++)
                    A[r][j] -= f * A[c][j];
            }
        }
        for (int j = 0; j < N; j++) {
            double h = A[j][N] / A[j][j];
            char buf[64];
            snprintf(buf, sizeof buf, "%.5e", h);
            table[k * N + j] = (float)strtod(buf, NULL);
        }
    }
    /* end rows are exact unit impulses upstream */
    for (int j = 0; j < N; j++) {
        table[0 * N + j] = (j == 4) ? 1.0f : 0.0f;
        table[XO_MMSE_NSTEPS * N + j] = (j == 3) ? 1.0f : 0.0f;
    }
}

/* gr::blocks::control_loop::update_gains, damping = sqrt(2)/2 */
void xo_costas_gains(float loop_bw, float *alpha, float *beta)
{
    float damping = sqrtf(2.0f) / 2.0f;
    float denom = (1.0f + 2.0f * damping * loop_bw + loop_bw * loop_bw);
    *alpha = (4 * damping * loop_bw) / denom;
    *beta = (4 * loop_bw * loop_bw) / denom;
}

/* ------------------------------------------------------------------------- */
/* NCO sine/cosine                                                           */
/* ------------------------------------------------------------------------- */
/*
 * The reference NCO is libm sinf/cosf (via libSatHelper), whose last-bit results differ
 * between libm builds.  To make the oracle a *complete* definition that another
 * implementation can reproduce bit for bit, the default NCO is the fully specified FP32
 * routine below: Cody-Waite reduction by pi/2 (3 constants, exact for |x| <= 2*pi plus
 * slack, which is all CostasLoop can produce after its phase wrap) and the Cephes
 * sinf/cosf minimax polynomials on |r| <= pi/4, every operation a single IEEE
 * round-to-nearest mul/add/fma in the order written.  It agrees with glibc sinf/cosf
 * to <= 2 ulp (tests/test_oracle.py).  xo_set_libm_sincos(1) switches the Costas loop
 * to libm sinf/cosf for cross-checks.
 */
static int xo_use_libm_sincos = 0;
void xo_set_libm_sincos(int on) { xo_use_libm_sincos = on; }

#define XO_TWO_OVER_PI 0.636619772367581343f
#define XO_PIO2_HI 1.5703125f                 /* 8 significant bits: q*HI exact for |q| < 2^16 */
#define XO_PIO2_MID 4.837512969970703125e-4f
#define XO_PIO2_LO 7.54978995489188216e-8f

void xo_sincosf(float x, float *sn, float *cs)
{
    float q = rintf(x * XO_TWO_OVER_PI);
    float r = fmaf(-q, XO_PIO2_HI, x);
    r = fmaf(-q, XO_PIO2_MID, r);
    r = fmaf(-q, XO_PIO2_LO, r);
    float z = r * r;
    /* sin(r) = r + r*z*(S1 + z*(S2 + z*S3)) */
    float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    float s = fmaf(ps * z, r, r);
    /* cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + z*C3)) */
    float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    int n = (int)q & 3;
    float so = (n & 1) ? c : s;
    float co = (n & 1) ? s : c;
    if (n & 2)
        so = -so;
    if ((n + 1) & 2)
        co = -co;
    *sn = so;
    *cs = co;
}

void xo_sincosf_array(const float *x, int64_t n, float *sn, float *cs)
{
    for (int64_t i = 0; i < n; i++)
        xo_sincosf(x[i], sn + i, cs + i);
}

/* gr::branchless_clip (used by costas_loop_cc and clock_recovery_mm_cc) */
static inline float xo_clip(float x, float clip)
{
    float x1 = fabsf(x + clip);
    float x2 = fabsf(x - clip);
    x1 -= x2;
    return 0.5f * x1;
}

/* ------------------------------------------------------------------------- */
/* FirFilter (fir_filter_ccf): real taps, complex samples, optional decimation */
/* ------------------------------------------------------------------------- */
struct xo_fir {
    unsigned decim;
    int ntaps;
    float *taps;
    float *hist;  /* (ntaps-1) complex samples preceding the next input */
    float *work;
    size_t work_cap; /* complex samples */
};

xo_fir *xo_fir_new(unsigned decimation, const float *taps, int ntaps)
{
    xo_fir *f = (xo_fir *)calloc(1, sizeof *f);
    f->decim = decimation ? decimation : 1;
    f->ntaps = ntaps;
    f->taps = (float *)malloc(sizeof(float) * (size_t)ntaps);
    memcpy(f->taps, taps, sizeof(float) * (size_t)ntaps);
    f->hist = (float *)calloc((size_t)(ntaps > 1 ? ntaps - 1 : 1) * 2, sizeof(float));
    return f;
}

void xo_fir_free(xo_fir *f)
{
    if (!f)
        return;
    free(f->taps);
    free(f->hist);
    free(f->work);
    free(f);
}

/*
 * Alternative summation order for cross-checks only (xo_set_fir_simd(1)): libSatHelper's FirFilter
 * evaluates the tap sum with a SIMD dot product, whose order is not the serial one.  This restates
 * the common 4-lane form -- lane l accumulates taps l, l+4, l+8, ... with separate multiply and add
 * (no FMA), the four lanes are added pairwise at the end -- so that tests can measure how far a
 * different but equally valid order moves the chain output (tests/test_oracle.py).
 */
static int xo_fir_simd = 0;
void xo_set_fir_simd(int on) { xo_fir_simd = on; }

static void xo_fir_dot_simd(const float *taps, int T, const float *x /* x[-2k] = sample k back */, float *o)
{
    float ar[4] = {0, 0, 0, 0}, ai[4] = {0, 0, 0, 0};
    for (int k = 0; k < T; k++) {
        const float pr = taps[k] * x[-2 * k], pi = taps[k] * x[-2 * k + 1];
        ar[k & 3] = ar[k & 3] + pr;
        ai[k & 3] = ai[k & 3] + pi;
    }
    o[0] = (ar[0] + ar[1]) + (ar[2] + ar[3]);
    o[1] = (ai[0] + ai[1]) + (ai[2] + ai[3]);
}

/* out[i] = sum_{k=0}^{T-1} taps[k] * x[i*D - k], accumulated k = 0..T-1 with fmaf. */
void xo_fir_work(xo_fir *f, const float *in, float *out, int n_out)
{
    const int T = f->ntaps, H = T - 1;
    const size_t D = f->decim;
    const size_t n_in = (size_t)n_out * D;
    if (n_out <= 0)
        return;
    if (f->work_cap < n_in + (size_t)H) {
        free(f->work);
        f->work_cap = n_in + (size_t)H;
        f->work = (float *)malloc(f->work_cap * 2 * sizeof(float));
    }
    float *w = f->work;
    memcpy(w, f->hist, sizeof(float) * 2 * (size_t)H);
    memcpy(w + 2 * (size_t)H, in, sizeof(float) * 2 * n_in);
    const float *taps = f->taps;
    if (xo_fir_simd) {
        for (size_t i = 0; i < (size_t)n_out; i++)
            xo_fir_dot_simd(taps, T, w + 2 * (i * D + (size_t)H), out + 2 * i);
    } else if (D == 1) {
        enum { BLK = 64 };
        size_t i = 0;
        for (; i + BLK <= (size_t)n_out; i += BLK) {
            float acc[2 * BLK];
            for (int b = 0; b < 2 * BLK; b++)
                acc[b] = 0.0f;
            const float *base = w + 2 * (i + (size_t)H);
            for (int k = 0; k < T; k++) {
                const float h = taps[k];
                const float *x = base - 2 * (size_t)k;
                for (int b = 0; b < 2 * BLK; b++)
                    acc[b] = fmaf(h, x[b], acc[b]);
            }
            memcpy(out + 2 * i, acc, sizeof acc);
        }
        for (; i < (size_t)n_out; i++) {
            float ar = 0.0f, ai = 0.0f;
            const float *x = w + 2 * (i + (size_t)H);
            for (int k = 0; k < T; k++) {
                ar = fmaf(taps[k], x[-2 * k], ar);
                ai = fmaf(taps[k], x[-2 * k + 1], ai);
            }
            out[2 * i] = ar;
            out[2 * i + 1] = ai;
        }
    } else {
        for (size_t i = 0; i < (size_t)n_out; i++) {
            float ar = 0.0f, ai = 0.0f;
            const float *x = w + 2 * (i * D + (size_t)H);
            for (int k = 0; k < T; k++) {
                ar = fmaf(taps[k], x[-2 * k], ar);
                ai = fmaf(taps[k], x[-2 * k + 1], ai);
            }
            out[2 * i] = ar;
            out[2 * i + 1] = ai;
        }
    }
    memcpy(f->hist, w + 2 * n_in, sizeof(float) * 2 * (size_t)H);
}

/* ------------------------------------------------------------------------- */
/* AGC (analog.agc_cc + set_max_gain) -- ctor demodulator.cpp:447             */
/* ------------------------------------------------------------------------- */
struct xo_agc {
    float rate, reference, gain, max_gain;
};

xo_agc *xo_agc_new(float rate, float reference, float gain, float max_gain)
{
    xo_agc *a = (xo_agc *)calloc(1, sizeof *a);
    a->rate = rate;
    a->reference = reference;
    a->gain = gain;
    a->max_gain = max_gain;
    return a;
}

void xo_agc_free(xo_agc *a) { free(a); }
float xo_agc_gain(const xo_agc *a) { return a->gain; }
void xo_agc_set_gain(xo_agc *a, float g) { a->gain = g; }

void xo_agc_work(xo_agc *a, const float *in, float *out, int n)
{
    float g = a->gain;
    const float rate = a->rate, ref = a->reference, mx = a->max_gain;
    for (int i = 0; i < n; i++) {
        float yr = in[2 * i] * g;
        float yi = in[2 * i + 1] * g;
        out[2 * i] = yr;
        out[2 * i + 1] = yi;
        g += rate * (ref - sqrtf(yr * yr + yi * yi));
        if (mx > 0.0f && g > mx)
            g = mx;
    }
    a->gain = g;
}

/* ------------------------------------------------------------------------- */
/* CostasLoop (digital.costas_loop_cc(bw, 2)) -- ctor demodulator.cpp:448      */
/* ------------------------------------------------------------------------- */
struct xo_costas {
    float phase, freq, alpha, beta, max_freq, min_freq;
    int order;
};

xo_costas *xo_costas_new(float loop_bw, int order)
{
    xo_costas *c = (xo_costas *)calloc(1, sizeof *c);
    xo_costas_gains(loop_bw, &c->alpha, &c->beta);
    c->max_freq = 1.0f;
    c->min_freq = -1.0f;
    c->order = order;
    return c;
}

void xo_costas_free(xo_costas *c) { free(c); }
void xo_costas_get(const xo_costas *c, float *phase, float *freq)
{
    *phase = c->phase;
    *freq = c->freq;
}
void xo_costas_set(xo_costas *c, float phase, float freq)
{
    c->phase = phase;
    c->freq = freq;
}

void xo_costas_work(xo_costas *c, const float *in, float *out, int n)
{
    float phase = c->phase, freq = c->freq;
    const float alpha = c->alpha, beta = c->beta;
    for (int i = 0; i < n; i++) {
        float np = -phase;
        float cs, sn;
        if (xo_use_libm_sincos) {
            cs = cosf(np);
            sn = sinf(np);
        } else {
            xo_sincosf(np, &sn, &cs);
        }
        float xr = in[2 * i], xi = in[2 * i + 1];
        float yr = xr * cs - xi * sn;
        float yi = xr * sn + xi * cs;
        out[2 * i] = yr;
        out[2 * i + 1] = yi;
        float err = yr * yi;           /* phase_detector_2 */
        err = xo_clip(err, 1.0f);
        freq = freq + beta * err;      /* advance_loop */
        phase = phase + freq + alpha * err;
        while (phase > XO_TWOPI)       /* phase_wrap (double constant upstream) */
            phase = (float)(phase - XO_TWOPI);
        while (phase < -XO_TWOPI)
            phase = (float)(phase + XO_TWOPI);
        if (freq > c->max_freq)        /* frequency_limit */
            freq = c->max_freq;
        else if (freq < c->min_freq)
            freq = c->min_freq;
    }
    c->phase = phase;
    c->freq = freq;
}

/* ------------------------------------------------------------------------- */
/* ClockRecovery (digital.clock_recovery_mm_cc) -- ctor demodulator.cpp:449    */
/* ------------------------------------------------------------------------- */
struct xo_mm {
    float mu, omega, omega_mid, omega_lim, gain_omega, gain_mu;
    float p0[2], p1[2], p2[2], c0[2], c1[2], c2[2];
    float table[(XO_MMSE_NSTEPS + 1) * XO_MMSE_NTAPS];
    float *buf;        /* unconsumed tail + current chunk */
    size_t buf_cap;    /* complex samples */
    size_t tail;       /* complex samples currently retained in buf */
    int64_t skip;      /* samples of future input to skip before the next base */
    int64_t consumed;  /* absolute index of buf[0] */
    /* optional per-symbol trace (tests only) */
    int64_t *tr_ii;
    float *tr_mu, *tr_omega, *tr_mm;
    int64_t tr_cap, tr_n;
};

xo_mm *xo_mm_new(float omega, float gain_omega, float mu, float gain_mu, float omega_rel_limit)
{
    xo_mm *m = (xo_mm *)calloc(1, sizeof *m);
    m->omega = omega;
    m->omega_mid = omega;
    m->omega_lim = omega_rel_limit * omega;
    m->mu = mu;
    m->gain_omega = gain_omega;
    m->gain_mu = gain_mu;
    xo_mmse_table(m->table);
    return m;
}

void xo_mm_free(xo_mm *m)
{
    if (!m)
        return;
    free(m->buf);
    free(m);
}

/* record (base index, mu, omega before the symbol; clipped mm of the symbol) for the
 * next `cap` symbols */
void xo_mm_trace(xo_mm *m, int64_t *ii, float *mu, float *omega, float *mm, int64_t cap)
{
    m->tr_ii = ii;
    m->tr_mu = mu;
    m->tr_omega = omega;
    m->tr_mm = mm;
    m->tr_cap = cap;
    m->tr_n = 0;
}

void xo_mm_get(const xo_mm *m, xo_mm_state *st)
{
    st->mu = m->mu;
    st->omega = m->omega;
    memcpy(st->p0, m->p0, sizeof st->p0);
    memcpy(st->p1, m->p1, sizeof st->p1);
    memcpy(st->p2, m->p2, sizeof st->p2);
    memcpy(st->c0, m->c0, sizeof st->c0);
    memcpy(st->c1, m->c1, sizeof st->c1);
    memcpy(st->c2, m->c2, sizeof st->c2);
    st->next_index = m->consumed + m->skip;
}

/* only valid when the retained tail is empty or the caller re-feeds from next_index */
void xo_mm_set(xo_mm *m, const xo_mm_state *st)
{
    m->mu = st->mu;
    m->omega = st->omega;
    memcpy(m->p0, st->p0, sizeof st->p0);
    memcpy(m->p1, st->p1, sizeof st->p1);
    memcpy(m->p2, st->p2, sizeof st->p2);
    memcpy(m->c0, st->c0, sizeof st->c0);
    memcpy(m->c1, st->c1, sizeof st->c1);
    memcpy(m->c2, st->c2, sizeof st->c2);
    m->tail = 0;
    m->skip = 0;
    m->consumed = st->next_index;
}

/* mmse_fir_interpolator_cc::interpolate: row rint(mu*128), taps applied time-reversed.
 * Summation order (part of the definition here, mirrored by the CUDA kernel):
 *   a_l = fmaf(t[l+4], x[l+4], t[l]*x[l]), l = 0..3;  result = (a0 + a1) + (a2 + a3)
 * where t[j] = row[7-j] -- the 4-lane SIMD dot product shape of the upstream kernel. */
static inline void xo_mm_interp(const float *table, const float *x, float mu, float *out)
{
    int k = (int)rintf(mu * XO_MMSE_NSTEPS);
    const float *row = table + k * XO_MMSE_NTAPS;
    float ar[4], ai[4];
    for (int l = 0; l < 4; l++) {
        float t0 = row[7 - l], t1 = row[7 - (l + 4)];
        ar[l] = fmaf(t1, x[2 * (l + 4)], t0 * x[2 * l]);
        ai[l] = fmaf(t1, x[2 * (l + 4) + 1], t0 * x[2 * l + 1]);
    }
    out[0] = (ar[0] + ar[1]) + (ar[2] + ar[3]);
    out[1] = (ai[0] + ai[1]) + (ai[2] + ai[3]);
}

int xo_mm_work(xo_mm *m, const float *in, float *out, int n)
{
    /* append the chunk behind the retained tail */
    size_t need = m->tail + (size_t)n;
    if (m->buf_cap < need) {
        m->buf_cap = need + 64;
        m->buf = (float *)realloc(m->buf, m->buf_cap * 2 * sizeof(float));
    }
    memcpy(m->buf + 2 * m->tail, in, sizeof(float) * 2 * (size_t)n);
    const int64_t nbuf = (int64_t)need;
    const float *buf = m->buf;

    int64_t ii = m->skip;
    int oo = 0;
    float mu = m->mu, omega = m->omega;
    float p0r = m->p0[0], p0i = m->p0[1], p1r = m->p1[0], p1i = m->p1[1], p2r = m->p2[0], p2i = m->p2[1];
    float c0r = m->c0[0], c0i = m->c0[1], c1r = m->c1[0], c1i = m->c1[1], c2r = m->c2[0], c2i = m->c2[1];

    while (ii + XO_MMSE_NTAPS <= nbuf) {
        float p[2];
        if (m->tr_n < m->tr_cap) {
            m->tr_ii[m->tr_n] = m->consumed + ii;
            m->tr_mu[m->tr_n] = mu;
            m->tr_omega[m->tr_n] = omega;
        }
        p2r = p1r; p2i = p1i;
        p1r = p0r; p1i = p0i;
        xo_mm_interp(m->table, buf + 2 * ii, mu, p);
        p0r = p[0]; p0i = p[1];

        c2r = c1r; c2i = c1i;
        c1r = c0r; c1i = c0i;
        c0r = p0r > 0.0f ? 1.0f : 0.0f;    /* slicer_0deg: 0/1, not +-1 */
        c0i = p0i > 0.0f ? 1.0f : 0.0f;

        /* x = (c0 - c2) * conj(p1);  y = (p0 - p2) * conj(c1);  mm = Re(y - x) */
        float ar = c0r - c2r, ai = c0i - c2i;
        float xr = ar * p1r + ai * p1i;
        float br = p0r - p2r, bi = p0i - p2i;
        float yr = br * c1r + bi * c1i;
        float mm_val = yr - xr;

        out[2 * oo] = p0r;
        out[2 * oo + 1] = p0i;
        oo++;

        mm_val = xo_clip(mm_val, 1.0f);
        if (m->tr_n < m->tr_cap)
            m->tr_mm[m->tr_n++] = mm_val;
        omega = omega + m->gain_omega * mm_val;
        omega = m->omega_mid + xo_clip(omega - m->omega_mid, m->omega_lim);
        mu = mu + omega + m->gain_mu * mm_val;
        float fl = floorf(mu);
        ii += (int64_t)(int)fl;
        mu -= fl;
        if (ii < 0)
            ii = 0;
    }

    m->mu = mu; m->omega = omega;
    m->p0[0] = p0r; m->p0[1] = p0i; m->p1[0] = p1r; m->p1[1] = p1i; m->p2[0] = p2r; m->p2[1] = p2i;
    m->c0[0] = c0r; m->c0[1] = c0i; m->c1[0] = c1r; m->c1[1] = c1i; m->c2[0] = c2r; m->c2[1] = c2i;

    /* retain the unconsumed tail */
    if (ii >= nbuf) {
        m->skip = ii - nbuf;
        m->consumed += nbuf;
        m->tail = 0;
    } else {
        size_t keep = (size_t)(nbuf - ii);
        memmove(m->buf, m->buf + 2 * ii, sizeof(float) * 2 * keep);
        m->tail = keep;
        m->consumed += ii;
        m->skip = 0;
    }
    return oo;
}

/* ------------------------------------------------------------------------- */
/* Chain: processSamples() (demodulator.cpp:100-168)                           */
/* ------------------------------------------------------------------------- */
struct xo_chain {
    xo_config cfg;
    float sps;
    xo_fir *decimator, *rrc;
    xo_agc *agc;
    xo_costas *costas;
    xo_mm *mm;
    float *b0, *b1;
    size_t cap;
};

void xo_config_defaults(xo_config *cfg, int hrit)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->sample_rate = hrit ? 2500000u : 1250000u;
    cfg->symbol_rate = hrit ? 927000u : 293883u;    /* Parameters.h:18,23 */
    cfg->rrc_alpha = hrit ? 0.3f : 0.5f;            /* Parameters.h:19,24 */
    cfg->decimation = 1;
    cfg->rrc_taps = 63;                             /* Parameters.h:28 */
    cfg->loop_order = 2;                            /* Parameters.h:27 */
    cfg->pll_alpha = 0.0037f;                       /* demodulator.cpp:220 (= CLOCK_ALPHA) */
    cfg->clock_alpha = 0.0037f;                     /* Parameters.h:30 */
    cfg->clock_mu = 0.5f;
    cfg->clock_omega_limit = 0.005f;
    cfg->agc_rate = 0.01f;
    cfg->agc_ref = 0.5f;
    cfg->agc_gain = 1.0f;
    cfg->agc_max_gain = 4000.0f;
}

xo_chain *xo_chain_new(const xo_config *cfg)
{
    xo_chain *c = (xo_chain *)calloc(1, sizeof *c);
    c->cfg = *cfg;
    unsigned D = cfg->decimation ? cfg->decimation : 1;
    /* demodulator.cpp:436-437: both divisions are float */
    float circuit_rate = (float)cfg->sample_rate / ((float)D);
    float sps = circuit_rate / ((float)cfg->symbol_rate);
    c->sps = sps;

    float rrc[4096];
    int nrrc = xo_rrc_taps(1, circuit_rate, cfg->symbol_rate, cfg->rrc_alpha, (int)cfg->rrc_taps, rrc);
    c->rrc = xo_fir_new(1, rrc, nrrc);
    if (D > 1) {
        int nlp = xo_lowpass_ntaps((double)cfg->sample_rate, 100e3);
        float *lp = (float *)malloc(sizeof(float) * (size_t)nlp);
        xo_lowpass_taps(1, (double)cfg->sample_rate, circuit_rate / 2, 100e3, lp);
        c->decimator = xo_fir_new(D, lp, nlp);
        free(lp);
    }
    c->agc = xo_agc_new(cfg->agc_rate, cfg->agc_ref, cfg->agc_gain, cfg->agc_max_gain);
    c->costas = xo_costas_new(cfg->pll_alpha, cfg->loop_order);
    /* CLOCK_GAIN_OMEGA = (CLOCK_ALPHA * CLOCK_ALPHA) / 4.0f  (Parameters.h:33) */
    float gain_omega = (cfg->clock_alpha * cfg->clock_alpha) / 4.0f;
    c->mm = xo_mm_new(sps, gain_omega, cfg->clock_mu, cfg->clock_alpha, cfg->clock_omega_limit);
    return c;
}

void xo_chain_free(xo_chain *c)
{
    if (!c)
        return;
    xo_fir_free(c->decimator);
    xo_fir_free(c->rrc);
    xo_agc_free(c->agc);
    xo_costas_free(c->costas);
    xo_mm_free(c->mm);
    free(c->b0);
    free(c->b1);
    free(c);
}

float xo_chain_sps(const xo_chain *c) { return c->sps; }

int64_t xo_chain_process_tap(xo_chain *c, const float *iq, int64_t n_complex, float *sym_out, int64_t cap,
                             float *dec_out, float *agc_out, float *rrc_out, float *costas_out)
{
    /* bounded sub-chunks keep the ping-pong buffers cache-sized; results are chunk-invariant */
    const int64_t CH = 1 << 16;
    unsigned D = c->decimator ? c->cfg.decimation : 1;
    if (!c->b0) {
        c->cap = (size_t)CH;
        c->b0 = (float *)malloc(sizeof(float) * 2 * c->cap);
        c->b1 = (float *)malloc(sizeof(float) * 2 * c->cap);
    }
    int64_t total = 0, done = 0, dpos = 0;
    n_complex -= n_complex % D;
    while (done < n_complex) {
        int64_t n = n_complex - done;
        if (n > CH)
            n = CH - (CH % D);
        const float *src = iq + 2 * done;
        int64_t len = n;
        float *ba = c->b0, *bb = c->b1, *t;
        if (D > 1) {
            len /= D;
            xo_fir_work(c->decimator, src, bb, (int)len);
            t = ba; ba = bb; bb = t;
            src = ba;
            if (dec_out)
                memcpy(dec_out + 2 * dpos, ba, sizeof(float) * 2 * (size_t)len);
        }
        xo_agc_work(c->agc, src, bb, (int)len);
        t = ba; ba = bb; bb = t;
        if (agc_out)
            memcpy(agc_out + 2 * dpos, ba, sizeof(float) * 2 * (size_t)len);
        xo_fir_work(c->rrc, ba, bb, (int)len);
        t = ba; ba = bb; bb = t;
        if (rrc_out)
            memcpy(rrc_out + 2 * dpos, ba, sizeof(float) * 2 * (size_t)len);
        xo_costas_work(c->costas, ba, bb, (int)len);
        t = ba; ba = bb; bb = t;
        if (costas_out)
            memcpy(costas_out + 2 * dpos, ba, sizeof(float) * 2 * (size_t)len);
        int ns = xo_mm_work(c->mm, ba, bb, (int)len);
        if (total + ns > cap)
            ns = (int)(cap - total);
        memcpy(sym_out + 2 * total, bb, sizeof(float) * 2 * (size_t)ns);
        total += ns;
        done += n;
        dpos += len;
    }
    return total;
}

int64_t xo_chain_process(xo_chain *c, const float *iq, int64_t n_complex, float *sym_out, int64_t cap)
{
    return xo_chain_process_tap(c, iq, n_complex, sym_out, cap, NULL, NULL, NULL, NULL);
}

/* SymbolManager::process (SymbolManager.cpp:43-46) */
void xo_soft_i8(const float *sym_cf32, int64_t n, int8_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        float f = sym_cf32[2 * i] * 127;
        f = f > 127 ? 127 : f;
        f = f < -128 ? -128 : f;
        out[i] = (int8_t)(f);
    }
}

/* onSamplesAvailable (demodulator.cpp:57-70) */
void xo_convert_s16(const int16_t *in, int64_t n_complex, float *out)
{
    for (int64_t i = 0; i < 2 * n_complex; i++)
        out[i] = in[i] / 32768.f;
}

void xo_convert_s8(const int8_t *in, int64_t n_complex, float *out)
{
    for (int64_t i = 0; i < 2 * n_complex; i++)
        out[i] = in[i] / 128.f;
}

/* SpyServer u8 samples (SpyServerFrontend.cpp:404-407): (v - 128) / 128.f */
void xo_convert_u8(const uint8_t *in, int64_t n_complex, float *out)
{
    for (int64_t i = 0; i < 2 * n_complex; i++)
        out[i] = (in[i] - 128) / 128.f;
}

/* RtlFrontend (RtlFrontend.cpp:27,57,104-116): lut[v] = (v - 128) * (1.f / 127.f), then the DC
 * blocker.  The reference tests `i % 1` (always 0), so every float -- I and Q -- goes through the
 * one running average iavg; kept as is.  *avg carries the state across calls. */
float xo_rtl_alpha(uint32_t sample_rate) { return (float)(1.f - exp(-1.0 / (sample_rate * 0.05f))); }
void xo_convert_rtl_u8(const uint8_t *in, int64_t n_complex, float alpha, float *avg, float *out)
{
    float iavg = *avg;
    for (int64_t i = 0; i < 2 * n_complex; i++) {
        float v = (in[i] - 128) * (1.f / 127.f);
        iavg += alpha * (v - iavg);
        v -= iavg;
        out[i] = v;
    }
    *avg = iavg;
}

/* DiagManager::threadLoop (DiagManager.cpp:35-42): val * 128, clamp to [-128, 127], static_cast<char> */
void xo_diag_i8(const float *v, int64_t n, int8_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        float val = v[i];
        val *= 128.f;
        val = val > 127 ? 127 : val;
        val = val < -128 ? -128 : val;
        out[i] = (int8_t)(val);
    }
}

/* ========================================================================= */
/* Decoder front half (reference decoder/src/newdecoder.cpp:212-290)          */
/* ========================================================================= */
/*
 * The decoder's first steps on the soft-symbol byte stream this path emits: sync-word correlation
 * (:218-247), phase fix (:268-270), r = 1/2 k = 7 Viterbi (:281-290), NRZ-M for HRIT.  The classes
 * behind those calls (SatHelper::Correlator, PacketFixer, Viterbi27 on top of libcorrect) are
 * un-vendored like the DSP classes, so this is a restatement too -- but the reference pins the
 * code itself: the four sync-word constants of newdecoder.cpp:21-24 are the convolutionally
 * encoded attached sync marker 0x1ACFFC1D, and only one convention reproduces them (polynomials
 * 0x4F then 0x6D on a register that shifts the newest bit in at the LSB, start state 0, coded bit
 * c stored inverted, i.e. coded 0 = positive symbol; HRIT words additionally NRZ-M encoded from 0
 * and from 1).  tests/test_decoder_oracle.py checks all four.
 */
#define XO_FRAMEBITS 8192
#define XO_CODEDFRAME 16384
#define XO_LASTBITS 64

static inline int xo_parity8(unsigned v)
{
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    return (int)(v & 1);
}

/* r = 1/2, k = 7: per input bit two coded bits, poly 0x4F first, then 0x6D; sr = ((sr << 1) | bit) & 0x7F */
void xo_conv_encode(const uint8_t *bits, int64_t n, unsigned *state, uint8_t *coded /* 2 n */)
{
    unsigned sr = *state & 0x7F;
    for (int64_t i = 0; i < n; i++) {
        sr = ((sr << 1) | (bits[i] & 1)) & 0x7F;
        coded[2 * i] = (uint8_t)xo_parity8(sr & 0x4F);
        coded[2 * i + 1] = (uint8_t)xo_parity8(sr & 0x6D);
    }
    *state = sr;
}

/* NRZ-M: out = out_prev ^ in (encode), in = out ^ out_prev (decode) */
void xo_nrzm_encode(const uint8_t *bits, int64_t n, uint8_t *last, uint8_t *out)
{
    uint8_t l = *last & 1;
    for (int64_t i = 0; i < n; i++) {
        l ^= bits[i] & 1;
        out[i] = l;
    }
    *last = l;
}

/* DifferentialEncoding::nrzmDecode on packed bytes (MSB first), newdecoder.cpp:284,289: every bit is replaced by
 * its xor with the bit before it; the bit before the first one is 0 */
void xo_nrzm_decode_bytes(uint8_t *data, int64_t n)
{
    uint8_t mask, last = 0;
    for (int64_t i = 0; i < n; i++) {
        mask = (uint8_t)((data[i] >> 1) & 0x7F);
        mask |= (uint8_t)(last << 7);
        last = data[i] & 1;
        data[i] ^= mask;
    }
}

/* Correlator::correlate over `length` soft bytes with the given 64-bit words (MSB = first symbol): per position the
 * number of symbols whose hard decision agrees with the word, a byte >= 127 reading as word bit 0 and a byte < 127 as
 * word bit 1; the first position (and, there, the first word) with the strictly highest count wins. */
void xo_correlate(const uint8_t *data, uint32_t length, const uint64_t *words, int n_words, uint32_t *highest,
                  uint32_t *position, uint32_t *word)
{
    uint32_t best = 0, bpos = 0, bword = 0;
    if (length > 64) {
        const uint32_t max_search = length - 64;
        for (uint32_t i = 0; i < max_search; i++) {
            for (int n = 0; n < n_words; n++) {
                uint32_t c = 0;
                for (int k = 0; k < 64; k++) {
                    const int wbit = (int)((words[n] >> (63 - k)) & 1);
                    const uint8_t d = data[i + k];
                    c += (uint32_t)(((d >= 127) & (wbit == 0)) | ((d < 127) & (wbit == 1)));
                }
                if (c > best) {
                    best = c;
                    bpos = i;
                    bword = (uint32_t)n;
                }
            }
        }
    }
    *highest = best;
    *position = bpos;
    *word = bword;
}

/* PacketFixer::fixPacket(data, len, DEG_180, false): a 180 degree BPSK ambiguity inverts every soft byte */
void xo_fix_packet_180(uint8_t *data, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        data[i] ^= 0xFF;
}

/*
 * Viterbi27::decode: maximum-likelihood decoding of n_bits information bits from 2 n_bits soft bytes with the linear
 * metric |u - 255 c| per coded bit, u the soft byte as the decoder library reads it (0 = surely coded 0 ... 255 = surely
 * coded 1).  soft_mode 0 (the reference call chain, literally): u is the raw byte -- newdecoder.cpp:215-216,281 hand the
 * int8 symbols of SymbolManager to Viterbi27::decode untouched, so a weak positive symbol (+1 -> 1) counts as a surer 0
 * than a strong one (+127 -> 127); the sign, and with it every hard decision, is right, only the confidence within each
 * half is mirrored.  soft_mode 1: u = 127 - v for the signed symbol v (+127 -> 0, -128 -> 255), the metric a decoder fed
 * offset-binary symbols would use -- in case libSatHelper converts internally, which could not be checked here.
 * All 64 start states are equally likely (metric 0), the survivor with the smallest final metric is traced back
 * (lowest state on ties), a state keeps the predecessor with the older bit 0 on equal metrics.  Output packed MSB first.
 * Returns the number of coded bits whose hard decision (raw byte >> 7) differs from the re-encoded output:
 * Viterbi27::GetBER.
 */
int xo_viterbi27_decode(const uint8_t *soft, int n_bits, int soft_mode, uint8_t *out_bytes)
{
    uint32_t *metric = (uint32_t *)malloc(sizeof(uint32_t) * 64 * 2);
    uint64_t *dec = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n_bits);
    uint8_t out_a[128], out_b[128];      /* coded bits of the transition into 7-bit register value r */
    for (unsigned r = 0; r < 128; r++) {
        out_a[r] = (uint8_t)xo_parity8(r & 0x4F);
        out_b[r] = (uint8_t)xo_parity8(r & 0x6D);
    }
    uint32_t *cur = metric, *nxt = metric + 64;
    for (int s = 0; s < 64; s++)
        cur[s] = 0;
    for (int t = 0; t < n_bits; t++) {
        const int y0 = soft_mode ? ((127 - soft[2 * t]) & 0xFF) : soft[2 * t];
        const int y1 = soft_mode ? ((127 - soft[2 * t + 1]) & 0xFF) : soft[2 * t + 1];
        const uint32_t m0[2] = {(uint32_t)y0, (uint32_t)(255 - y0)};   /* |y - 255 c| for c = 0, 1 */
        const uint32_t m1[2] = {(uint32_t)y1, (uint32_t)(255 - y1)};
        uint64_t d = 0;
        for (unsigned ns = 0; ns < 64; ns++) {
            /* new state ns = ((old << 1) | bit) & 63; old is (ns >> 1) or (ns >> 1) | 32; register = (old << 1) | bit */
            const unsigned p0 = ns >> 1, p1 = (ns >> 1) | 32;
            const unsigned r0 = (p0 << 1) | (ns & 1), r1 = (p1 << 1) | (ns & 1);
            const uint32_t a = cur[p0] + m0[out_a[r0]] + m1[out_b[r0]];
            const uint32_t b = cur[p1] + m0[out_a[r1]] + m1[out_b[r1]];
            if (a <= b) {
                nxt[ns] = a;
            } else {
                nxt[ns] = b;
                d |= 1ull << ns;
            }
        }
        dec[t] = d;
        uint32_t *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    unsigned s = 0;
    for (unsigned k = 1; k < 64; k++)
        if (cur[k] < cur[s])
            s = k;
    memset(out_bytes, 0, (size_t)((n_bits + 7) / 8));
    for (int t = n_bits - 1; t >= 0; t--) {
        const unsigned bit = s & 1;
        if (bit)
            out_bytes[t >> 3] |= (uint8_t)(0x80 >> (t & 7));
        s = (s >> 1) | ((unsigned)((dec[t] >> s) & 1) << 5);
    }
    /* re-encode from the state the trace-back ended in and count disagreements with the hard decisions */
    unsigned sr = s << 0;   /* s now holds the 6 bits before the first decoded bit */
    int errors = 0;
    for (int t = 0; t < n_bits; t++) {
        const unsigned bit = (out_bytes[t >> 3] >> (7 - (t & 7))) & 1;
        sr = ((sr << 1) | bit) & 0x7F;
        errors += (xo_parity8(sr & 0x4F) != (soft[2 * t] >> 7)) + (xo_parity8(sr & 0x6D) != (soft[2 * t + 1] >> 7));
    }
    free(metric);
    free(dec);
    return errors;
}

/*
 * The loop body of newdecoder.cpp:212-300 over a buffered byte stream, without the flywheel shortcut (:222-236, which
 * only skips work when the sync word sits where it is expected): take 16384 bytes, correlate, drop the chunk when the
 * best correlation is below MINCORRELATIONBITS (46, :244-247), otherwise re-align on the sync word by pulling the
 * missing bytes from the stream (:250-264), undo a 180 degree phase (LRIT only, :268-270), prepend the last 64 soft
 * bytes of the previous frame (USE_LAST_FRAME_DATA, :273-275,296), decode, NRZ-M decode for HRIT (:283-285), drop the 4
 * warm-up bytes (:293).  frames: 1024 bytes each; meta: 4 ints per frame {stream offset of the frame, correlation,
 * word, Viterbi bit errors}; last_end: 64 bytes carried between calls (128s at start, :141-145).  Returns the number of
 * frames; *consumed = bytes of the stream that are done with.
 */
int64_t xo_decoder_front(const uint8_t *stream, int64_t n, int lrit, int soft_mode, uint8_t *last_end, uint8_t *frames,
                         int32_t *meta, int64_t cap, int64_t *consumed)
{
    static const uint64_t HRIT_UW[2] = {0xfc4ef4fd0cc2df89ull, 0x25010b02f33d2076ull};
    static const uint64_t LRIT_UW[2] = {0xfca2b63db00d9794ull, 0x035d49c24ff2686bull};
    const uint64_t *words = lrit ? LRIT_UW : HRIT_UW;
    uint8_t *vit = (uint8_t *)malloc(XO_CODEDFRAME + XO_LASTBITS);
    uint8_t dec[(XO_FRAMEBITS + XO_LASTBITS / 2) / 8];
    int64_t s = 0, nf = 0;
    while (s + XO_CODEDFRAME <= n && nf < cap) {
        uint32_t corr, pos, word;
        xo_correlate(stream + s, XO_CODEDFRAME, words, 2, &corr, &pos, &word);
        if (corr < 46) {
            s += XO_CODEDFRAME;
            continue;
        }
        const int64_t f = s + pos;
        if (f + XO_CODEDFRAME > n)
            break;   /* the rest of the frame has not arrived yet */
        memcpy(vit, last_end, XO_LASTBITS);
        memcpy(vit + XO_LASTBITS, stream + f, XO_CODEDFRAME);
        if (lrit && word == 1)
            xo_fix_packet_180(vit + XO_LASTBITS, XO_CODEDFRAME);
        const int ber = xo_viterbi27_decode(vit, XO_FRAMEBITS + XO_LASTBITS / 2, soft_mode, dec);
        if (!lrit)
            xo_nrzm_decode_bytes(dec, sizeof dec);
        memcpy(frames + 1024 * nf, dec + 4, 1024);
        memcpy(last_end, vit + XO_CODEDFRAME, XO_LASTBITS);
        meta[4 * nf] = (int32_t)f;
        meta[4 * nf + 1] = (int32_t)corr;
        meta[4 * nf + 2] = (int32_t)word;
        meta[4 * nf + 3] = ber;
        nf++;
        s = f + XO_CODEDFRAME;
    }
    free(vit);
    *consumed = s;
    return nf;
}
