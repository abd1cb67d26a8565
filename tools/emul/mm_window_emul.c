/*
 * mm_window_emul.c -- CPU emulation of the sliding-window fixed-point M&M chain
 * (the algorithm of mm_chain_kernel in xritdemod_b200/csrc), checked against the oracle.
 * Development aid: validates exactness and measures the advance per iteration.
 *   gcc -O2 -ffp-contract=off -I../../oracle mm_window_emul.c ../../oracle/xrit_oracle.c -lm -o /tmp/mm_emul
 *   /tmp/mm_emul costas.cf32 NT
 */
#include "xrit_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float table[129 * 8];
static float omega_mid, omega_lim, gain_omega, gain_mu;

static inline float clip_bl(float x, float c)
{
    float x1 = fabsf(x + c), x2 = fabsf(x - c);
    x1 -= x2;
    return 0.5f * x1;
}

static void interp(const float *x, float mu, float *out)
{
    int k = (int)rintf(mu * 128);
    const float *row = table + k * 8;
    float ar[4], ai[4];
    for (int l = 0; l < 4; l++) {
        float t0 = row[7 - l], t1 = row[7 - (l + 4)];
        ar[l] = fmaf(t1, x[2 * (l + 4)], t0 * x[2 * l]);
        ai[l] = fmaf(t1, x[2 * (l + 4) + 1], t0 * x[2 * l + 1]);
    }
    out[0] = (ar[0] + ar[1]) + (ar[2] + ar[3]);
    out[1] = (ai[0] + ai[1]) + (ai[2] + ai[3]);
}

static void update(const float *p0, const float *p1, const float *p2, float *mu, float *omega, long long *ii)
{
    float c0r = p0[0] > 0, c0i = p0[1] > 0, c1r = p1[0] > 0, c1i = p1[1] > 0, c2r = p2[0] > 0, c2i = p2[1] > 0;
    float ar = c0r - c2r, ai = c0i - c2i;
    float xr = ar * p1[0] + ai * p1[1];
    float br = p0[0] - p2[0], bi = p0[1] - p2[1];
    float yr = br * c1r + bi * c1i;
    float mm = clip_bl(yr - xr, 1.0f);
    float om = *omega + gain_omega * mm;
    om = omega_mid + clip_bl(om - omega_mid, omega_lim);
    float m = *mu + om + gain_mu * mm;
    float fl = floorf(m);
    *ii += (long long)(int)fl;
    *mu = m - fl;
    *omega = om;
}

#define FIX 4294967296.0f
#define UNFIX 2.3283064365386963e-10f

int main(int argc, char **argv)
{
    const char *path = argv[1];
    int NT = argc > 2 ? atoi(argv[2]) : 1024;
    int newton = argc > 3 ? atoi(argv[3]) : 0;
    FILE *f = fopen(path, "rb");
    fseek(f, 0, SEEK_END);
    long n = ftell(f) / 8;
    fseek(f, 0, SEEK_SET);
    float *x = malloc(8 * (n + 64));
    if (fread(x, 8, n, f) != (size_t)n) return 1;
    fclose(f);
    float sps = 2500000.f / 927000.f;
    float gm = 0.0037f, go = gm * gm / 4.0f;
    xo_mm *m = xo_mm_new(sps, go, 0.5f, gm, 0.005f);
    float *ref = malloc(8 * (n / 2 + 64));
    int nref = xo_mm_work(m, x, ref, (int)n);
    xo_mmse_table(table);
    omega_mid = sps; omega_lim = 0.005f * sps; gain_omega = go; gain_mu = gm;

    long long *T = malloc(8 * NT), *W = malloc(8 * NT), *nT = malloc(8 * NT), *nW = malloc(8 * NT);
    long long *dT = malloc(8 * NT), *dW = malloc(8 * NT);
    float *p = malloc(8 * NT);
    float *out = malloc(8 * (n / 2 + 64));
    long long Tb = (long long)(0.5f * FIX), Wb = (long long)(sps * FIX);
    float P1[2] = {0, 0}, P2[2] = {0, 0};
    int tb = 0;
    long count = 0, iters = 0;
    const long long last_ok = n - 8;
    for (int r = 0; r < NT; r++) { T[r] = Tb + r * Wb; W[r] = Wb; }
    double slope = 0.0;   /* running estimate of d(dT)/dT for the quasi-Newton correction */
    for (;;) {
        iters++;
        /* evaluate all lanes at their believed states */
        for (int t = 0; t < NT; t++) {
            long long ii = T[t] >> 32;
            float mu = (float)(unsigned)(T[t] & 0xffffffffLL) * UNFIX;
            long long iic = ii < 0 ? 0 : (ii > last_ok ? last_ok : ii);
            interp(x + 2 * iic, mu, p + 2 * t);
        }
        for (int t = 0; t < NT; t++) {
            int r = (t - tb) & (NT - 1);
            const float *p1 = r >= 1 ? p + 2 * ((t - 1) & (NT - 1)) : P1;
            const float *p2 = r >= 2 ? p + 2 * ((t - 2) & (NT - 1)) : (r == 1 ? P1 : P2);
            long long ii = T[t] >> 32, ii2 = ii;
            float mu = (float)(unsigned)(T[t] & 0xffffffffLL) * UNFIX, mu2 = mu;
            float om = (float)W[t] * UNFIX, om2 = om;
            update(p + 2 * t, p1, p2, &mu2, &om2, &ii2);
            dT[t] = (ii2 - ii) * 4294967296LL + ((long long)(mu2 * FIX) - (T[t] & 0xffffffffLL));
            dW[t] = (long long)(om2 * FIX) - W[t];
        }
        /* rotated exclusive prefix from the base lane */
        long long aT = Tb, aW = Wb;
        int first_changed = NT;
        double corr = 0.0;
        for (int r = 0; r < NT; r++) {
            int t = (tb + r) & (NT - 1);
            nT[t] = aT; nW[t] = aW;
            if ((nT[t] != T[t] || nW[t] != W[t]) && first_changed == NT) first_changed = r;
            aT += dT[t]; aW += dW[t];
        }
        int A = first_changed;
        /* stop lane: first exact lane (r <= A) that is not computable */
        int stop = -1;
        for (int r = 0; r <= A && r < NT; r++) {
            int t = (tb + r) & (NT - 1);
            if ((nT[t] >> 32) > last_ok) { stop = r; break; }
        }
        int emit = (stop >= 0) ? stop : A;
        for (int r = 0; r < emit; r++) {
            int t = (tb + r) & (NT - 1);
            out[2 * count] = p[2 * t]; out[2 * count + 1] = p[2 * t + 1];
            count++;
        }
        if (stop >= 0) break;
        /* carry P1, P2 */
        if (A >= 2) { int t1 = (tb + A - 1) & (NT - 1), t2 = (tb + A - 2) & (NT - 1); P2[0] = p[2*t2]; P2[1] = p[2*t2+1]; P1[0] = p[2*t1]; P1[1] = p[2*t1+1]; }
        else if (A == 1) { int t1 = tb; P2[0] = P1[0]; P2[1] = P1[1]; P1[0] = p[2*t1]; P1[1] = p[2*t1+1]; }
        /* new base */
        long long endT = aT, endW = aW;   /* state after the last lane */
        if (A < NT) { int t = (tb + A) & (NT - 1); Tb = nT[t]; Wb = nW[t]; } else { Tb = endT; Wb = endW; }
        /* kept lanes take their new states; optional quasi-Newton damping of the correction */
        for (int r = A; r < NT; r++) {
            int t = (tb + r) & (NT - 1);
            if (newton) {
                /* correction delta = nT - T; feedback model: later lanes' increments respond with gain -slope per
                   unit of accumulated correction: delta_eff_r = delta_r - sum_{j<r} a*delta_eff_j  */
                double d = (double)(nT[t] - T[t]);
                double de = d - corr;
                corr += slope * de;
                T[t] = T[t] + (long long)llrint(de / 1024.0) * 1024;
                W[t] = nW[t];
            } else { T[t] = nT[t]; W[t] = nW[t]; }
        }
        /* recycled lanes: linear extrapolation from the end state */
        for (int j = 0; j < A; j++) {
            int t = (tb + j) & (NT - 1);
            T[t] = endT + (long long)j * endW; W[t] = endW;
        }
        tb = (tb + A) & (NT - 1);
        (void)slope;
        slope = newton ? (double)newton * 1e-4 : 0.0;
    }
    int bad = 0;
    long cmp = count < nref ? count : nref;
    for (long i = 0; i < cmp; i++) if (out[2*i] != ref[2*i] || out[2*i+1] != ref[2*i+1]) { if (!bad) printf("first diff at symbol %ld\n", i); bad++; }
    printf("NT=%d newton=%d: symbols %ld (oracle %d) mismatches %d; iterations %ld -> advance %.1f symbols/iteration, %.2f lane-evals/symbol\n",
           NT, newton, count, nref, bad, iters, (double)count / iters, (double)iters * NT / count);
    return 0;
}
