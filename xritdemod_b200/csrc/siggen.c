/*
 * siggen.c -- deterministic synthetic LRIT/HRIT-shaped BPSK IQ source (host side).
 *
 * Stands in for the reference's only offline input, CFileFrontend (raw interleaved
 * complex<float>, reference demodulator/src/CFileFrontend.cpp:33-62): it produces
 * the same cf32 layout the sample callback carries (FrontendDevice.h:11-13,37).
 * Signal model: random +-1 bits, RRC pulse shaping evaluated at
 * the exact non-integer samples-per-symbol, timing offset, carrier offset + phase,
 * linear amplitude ramp, complex AWGN at a given Es/N0.  Every sample is a pure
 * function of (seed, absolute sample index), so any chunking yields the same stream.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    double sample_rate;
    double symbol_rate;
    double rrc_alpha;
    double timing_offset; /* symbols */
    double carrier_hz;
    double phase0;        /* rad */
    double amp_start, amp_end;
    uint64_t ramp_len;    /* samples over which the amplitude ramps, then holds */
    double esn0_db;
    int32_t noise;        /* 0 = noise-free */
    int32_t reserved;
    uint64_t seed;
} xrd_sig_params;

#define SPAN 10
#define OSR 2048
#define TBL (2 * SPAN * OSR + 2)

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline double sig_bit(uint64_t seed, int64_t k)
{
    return (splitmix64(seed * 0x2545F4914F6CDD1Dull + (uint64_t)k) >> 63) ? 1.0 : -1.0;
}

/* continuous-time unit-energy RRC impulse response, t in symbols */
static double rrc_pulse(double t, double a)
{
    if (fabs(t) < 1e-9)
        return 1.0 - a + 4.0 * a / M_PI;
    if (a > 0 && fabs(fabs(t) - 1.0 / (4.0 * a)) < 1e-9)
        return (a / sqrt(2.0)) * ((1.0 + 2.0 / M_PI) * sin(M_PI / (4.0 * a)) + (1.0 - 2.0 / M_PI) * cos(M_PI / (4.0 * a)));
    double num = sin(M_PI * t * (1.0 - a)) + 4.0 * a * t * cos(M_PI * t * (1.0 + a));
    double den = M_PI * t * (1.0 - 16.0 * a * a * t * t);
    return num / den;
}

/* worker threads of the generator (launchers such as torchrun export OMP_NUM_THREADS=1, which would make
 * every rank generate its stream on one core); 0 = leave the OpenMP default */
void xrd_siggen_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void xrd_siggen_bits(uint64_t seed, int64_t k_start, int64_t n, int8_t *out)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = (int8_t)sig_bit(seed, k_start + i);
}

/* writes n complex samples (interleaved float I,Q) for absolute indices [start, start+n) */
void xrd_siggen_cf32(const xrd_sig_params *p, uint64_t start, uint64_t n, float *out)
{
    double *tbl = (double *)malloc(sizeof(double) * TBL);
    for (int i = 0; i < TBL; i++)
        tbl[i] = rrc_pulse((double)i / OSR - SPAN, p->rrc_alpha);

    const double sps = p->sample_rate / p->symbol_rate;
    const double inv_sps = 1.0 / sps;
    const double w = 2.0 * M_PI * p->carrier_hz / p->sample_rate;
    /* unit symbol amplitude => unit sample power; sigma^2 = sps / (Es/N0) (complex) */
    const double sigma = p->noise ? sqrt(sps / pow(10.0, p->esn0_db / 10.0) / 2.0) : 0.0;
    const uint64_t nseed = splitmix64(p->seed ^ 0xA5A5A5A55A5A5A5Aull);
    const uint64_t BLK = 1 << 14;
    const int64_t nblk = (int64_t)((n + BLK - 1) / BLK);

#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < nblk; b++) {
        uint64_t i0 = (uint64_t)b * BLK, i1 = i0 + BLK < n ? i0 + BLK : n;
        double bits[2 * SPAN];
        int64_t cur_k0 = INT64_MIN;
        for (uint64_t i = i0; i < i1; i++) {
            uint64_t idx = start + i;
            double t = (double)idx * inv_sps + p->timing_offset;
            double fk = floor(t);
            int64_t k0 = (int64_t)fk;
            double frac = t - fk;
            if (k0 != cur_k0) {
                for (int j = 0; j < 2 * SPAN; j++)
                    bits[j] = sig_bit(p->seed, k0 - SPAN + 1 + j);
                cur_k0 = k0;
            }
            /* s = sum_j b[k0+j] g(frac - j), j = -SPAN+1..SPAN */
            double s = 0.0;
            for (int j = 0; j < 2 * SPAN; j++) {
                double tau = frac - (double)(j - SPAN + 1); /* in (-SPAN, SPAN) */
                double pos = (tau + SPAN) * OSR;
                int ip = (int)pos;
                double fr = pos - ip;
                s += bits[j] * (tbl[ip] + fr * (tbl[ip + 1] - tbl[ip]));
            }
            double th = w * (double)idx + p->phase0;
            double re = s * cos(th), im = s * sin(th);
            if (p->noise) {
                uint64_t r1 = splitmix64(nseed + 2 * idx), r2 = splitmix64(nseed + 2 * idx + 1);
                double u1 = ((double)(r1 >> 11) + 1.0) * (1.0 / 9007199254740993.0);
                double u2 = (double)(r2 >> 11) * (1.0 / 9007199254740992.0);
                double rad = sigma * sqrt(-2.0 * log(u1));
                re += rad * cos(2.0 * M_PI * u2);
                im += rad * sin(2.0 * M_PI * u2);
            }
            double amp = p->amp_end;
            if (p->ramp_len > 0 && idx < p->ramp_len)
                amp = p->amp_start + (p->amp_end - p->amp_start) * ((double)idx / (double)p->ramp_len);
            out[2 * i] = (float)(amp * re);
            out[2 * i + 1] = (float)(amp * im);
        }
    }
    free(tbl);
}

/* cf32 -> interleaved s16 / s8 in the scaling onSamplesAvailable undoes (demodulator.cpp:57-70) */
void xrd_cf32_to_s16(const float *in, uint64_t n_complex, int16_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(2 * n_complex); i++) {
        float v = rintf(in[i] * 32768.f);
        v = v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v);
        out[i] = (int16_t)v;
    }
}

/* cf32 -> the SpyServer / RTL-SDR u8 format (offset binary around 128) */
void xrd_cf32_to_u8(const float *in, uint64_t n_complex, uint8_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(2 * n_complex); i++) {
        float v = rintf(in[i] * 128.f) + 128.f;
        v = v > 255.f ? 255.f : (v < 0.f ? 0.f : v);
        out[i] = (uint8_t)v;
    }
}

void xrd_cf32_to_s8(const float *in, uint64_t n_complex, int8_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(2 * n_complex); i++) {
        float v = rintf(in[i] * 128.f);
        v = v > 127.f ? 127.f : (v < -128.f ? -128.f : v);
        out[i] = (int8_t)v;
    }
}
