// xrd_mmwn.cuh -- Mueller & Mueller clock recovery (ClockRecovery::Work) as a window-Newton chain.
//
// Same idea as xrd_wn.cuh, on the symbol-rate loop: one CTA advances one chain, NT = 32*WPC*K symbols
// at a time.  Slot s = t*K + k of thread t holds a BELIEVED loop state (ii, mu, omega) for its symbol;
// every iteration every slot
//   1. interpolates its symbol p0 at the believed (ii, mu)   (skipped while (ii, k) has not moved),
//   2. applies the literal loop update with the interpolants of the two preceding slots -> the state
//      the next slot should have,
//   3. the exact state differences (timing as (ii' - ii) + (mu' - mu), omega' - omega, both exact in
//      FP64 inside one window) are prefix-summed from the exact base: the next believed states.
// Acceptance is literal: slot r is exact iff slot r-1 is exact and valid and believed[r] equals the
// state slot r-1 produced, bit for bit.  The leading exact run leaves the window, the freed slots
// re-enter at the far end with linearly extrapolated states.  Compared with mm_chain32_kernel (one
// slot per thread, 32-bit fixed point, "unchanged" acceptance) a thread carries K slots, so the
// scans, barriers and control are paid once per K symbols.
#pragma once
#include "xrd_kernels.cuh"

namespace xrd {

struct MmWnLayout {
    // shared memory carve-up (bytes) for NT slots, WPC warps, ring of R samples
    size_t tab, p, oi, om, tot, bex, red, misc, ring, total;
    __host__ __device__ MmWnLayout(int NT, int WPC, int R)
    {
        size_t o = 0;
        tab = o;  o += sizeof(float) * MM_TAB_PAD;
        p = o;    o += sizeof(float2) * 2 * NT;          // [2][NT] interpolants
        om = o;   o += sizeof(float2) * 2 * NT;          // [2][NT] produced (mu, omega)
        tot = o;  o += sizeof(double) * 2 * WPC * 2;     // [2][WPC][2] warp totals
        bex = o;  o += sizeof(double) * 2 * 2;           // [2][2] base thread's in-warp exclusive prefix
        ring = o; o += sizeof(float2) * (size_t)R;
        oi = o;   o += sizeof(int) * 2 * NT;             // [2][NT] produced ii
        red = o;  o += sizeof(int) * 2 * 4;              // [2][4] first bad / stop / entry / checkpoint rank (atomicMin)
        misc = o; o += 16;
        total = o;
    }
};

__device__ __forceinline__ void mm_update_i(const MmParams &p, float2 p0, float2 p1, float2 p2, float &mu, float &omega, int &ii)
{
    long long t = ii;
    mm_update(p, p0, p1, p2, mu, omega, t);
    ii = (int)t;
}

template <int K, int WPC>
__global__ void __launch_bounds__(32 * WPC)
mm_wn_kernel(const float2 *__restrict__ in, float2 *__restrict__ stage, int n, int L, int W, int nseg, int cap_seg,
             MmState *__restrict__ entry, MmState *__restrict__ exit_, const MmState *__restrict__ carried,
             const unsigned char *__restrict__ redo, MmSegOut *__restrict__ segout, const float *__restrict__ table,
             MmParams prm, int mode, long long in_ch_stride, long long stage_ch_stride, int R, MmCk *__restrict__ ckpt,
             int ncp, int C)
{
    constexpr int T = 32 * WPC, NT = T * K, BIG = 1 << 30, INVALID = (int)0x80000000;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const MmWnLayout lay(NT, WPC, R);
    float *s_tab = reinterpret_cast<float *>(s_raw + lay.tab);
    float2 *s_p = reinterpret_cast<float2 *>(s_raw + lay.p);
    float2 *s_om = reinterpret_cast<float2 *>(s_raw + lay.om);
    double *s_tot = reinterpret_cast<double *>(s_raw + lay.tot);
    double *s_bex = reinterpret_cast<double *>(s_raw + lay.bex);
    float2 *s_x = reinterpret_cast<float2 *>(s_raw + lay.ring);
    int *s_oi = reinterpret_cast<int *>(s_raw + lay.oi);
    int *s_red = reinterpret_cast<int *>(s_raw + lay.red);
    int *s_misc = reinterpret_cast<int *>(s_raw + lay.misc);

    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int j = blockIdx.x, ch = blockIdx.y;
    in += (size_t)ch * in_ch_stride;
    stage += (size_t)ch * stage_ch_stride + (size_t)j * cap_seg;
    entry += (size_t)ch * nseg;
    exit_ += (size_t)ch * nseg;
    segout += (size_t)ch * nseg;
    if (mode == 1 && !redo[(size_t)ch * nseg + j]) return;
    for (int i = t; i < 129 * 8; i += T) {
        const int k = i >> 3, tp = i & 7;
        s_tab[tp * 129 + k] = table[i];
    }
    if (t == 0) s_misc[0] = 0;
    if (t < 8) s_red[t] = NT;
    MmCk *ck = ckpt ? ckpt + ((size_t)ch * nseg + j) * ncp : nullptr;
    int next_ck = j * L + C, ck_idx = 0;
    bool merged = false;

    const int seg0 = (j == 0) ? -BIG : j * L;
    const int seg1 = (j == nseg - 1) ? BIG : (j + 1) * L;
    const int last_ok = n - MM_NTAPS;
    MmState st;
    bool have_entry;
    if (mode == 0) {
        const long long begin = (long long)j * L - W;
        if (j == 0 || begin <= 0) {
            st = carried[ch];
        } else {
            st.ii = begin;
            st.mu = 0.5f;
            st.omega = prm.omega_mid;
            st.p0 = make_float2(0.f, 0.f);
            st.p1 = make_float2(0.f, 0.f);
        }
        have_entry = false;
    } else {
        st = entry[j];
        have_entry = true;
    }
    // exact base state (uniform)
    int ii_b = (int)st.ii;
    float mu_b = st.mu, om_b = st.omega;
    float2 P1 = st.p0, P2 = st.p1;
    int tbt = 0, count = 0, overflow = 0, iters = 0, par = 0;
    const int RM = R - 1;
    const int lo_min = -MM_TAIL;

    // believed states: linear extrapolation with zero timing error
    int bii[K], c_ii[K], c_k[K];
    float bmu[K], bom[K];
    float2 p0[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double Tr = (double)mu_b + (double)(t * K + k) * (double)om_b;
        const double fl = floor(Tr);
        bii[k] = ii_b + (int)fl;
        bmu[k] = (float)(Tr - fl);
        bom[k] = om_b;
        c_ii[k] = -BIG;
        c_k[k] = -1;
        p0[k] = make_float2(0.f, 0.f);
    }
    // ring: sample i sits in s_x[i & RM]; [x_ready - R, x_ready) resident and visible, [x_ready, x_fill) in flight
    int x_fill = (ii_b < lo_min ? lo_min : ii_b) & ~31;
    {
        const int target = x_fill + R;
        for (int i = x_fill + t; i < target; i += T)
            if (i >= lo_min && i < n) cp_async8(&s_x[i & RM], in + i);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        x_fill = target;
    }
    cp_async_wait_all();
    int x_ready = x_fill;
    __syncthreads();

    for (;;) {
        iters++;
        float2 *sp = s_p + par * NT;
        float2 *som = s_om + par * NT;
        int *soi = s_oi + par * NT;
        // ---- 1. interpolate every slot at its believed state
        bool valid[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int ii = bii[k];
            const int kk = (int)rintf(bmu[k] * (float)MM_NSTEPS);
            valid[k] = (ii >= x_fill - R) && (ii >= lo_min) && (ii + MM_NTAPS <= x_ready) && (ii <= last_ok);
            if (!valid[k]) {
                p0[k] = make_float2(0.f, 0.f);
                c_k[k] = -1;
            } else if (ii != c_ii[k] || kk != c_k[k]) {
                c_ii[k] = ii;
                c_k[k] = kk;
                const int b = ii & RM;
                float ar[4], ai[4];
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const float t0 = s_tab[(7 - l) * 129 + kk];
                    const float t1 = s_tab[(3 - l) * 129 + kk];
                    const float2 a = s_x[(b + l) & RM], bb = s_x[(b + l + 4) & RM];
                    ar[l] = fmaf(t1, bb.x, t0 * a.x);
                    ai[l] = fmaf(t1, bb.y, t0 * a.y);
                }
                p0[k] = make_float2((ar[0] + ar[1]) + (ar[2] + ar[3]), (ai[0] + ai[1]) + (ai[2] + ai[3]));
            }
            sp[t * K + k] = p0[k];
        }
        __syncthreads();   // B1: interpolants visible
        // ---- 2. literal update of every slot, exact differences, in-warp scan
        const int lr = ((t - tbt) & (T - 1)) * K;   // rank of this thread's slot 0
        int oii[K];
        float omu[K], oom[K];
        double liT[K], liW[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            float2 p1 = (k >= 1) ? p0[k - 1] : sp[(t * K + NT - 1) & (NT - 1)];
            float2 p2 = (k >= 2) ? p0[k - 2] : sp[(t * K + k + NT - 2) & (NT - 1)];
            if (lr + k == 0) { p1 = P1; p2 = P2; }
            if (lr + k == 1) { p2 = P1; }
            oii[k] = bii[k];
            omu[k] = bmu[k];
            oom[k] = bom[k];
            mm_update_i(prm, p0[k], p1, p2, omu[k], oom[k], oii[k]);
            const double dT = (double)(oii[k] - bii[k]) + ((double)omu[k] - (double)bmu[k]);
            const double dW = (double)oom[k] - (double)bom[k];
            liT[k] = k ? liT[k - 1] + dT : dT;
            liW[k] = k ? liW[k - 1] + dW : dW;
            soi[t * K + k] = valid[k] ? oii[k] : INVALID;   // nothing may be accepted on top of an invalid slot
            som[t * K + k] = make_float2(omu[k], oom[k]);
        }
        double weT, weW;
        {
            double vT = liT[K - 1], vW = liW[K - 1];
#pragma unroll
            for (int ofs = 1; ofs < 32; ofs <<= 1) {
                const double a = __shfl_up_sync(0xffffffffu, vT, ofs);
                const double b = __shfl_up_sync(0xffffffffu, vW, ofs);
                if (lane >= ofs) { vT += a; vW += b; }
            }
            weT = vT - liT[K - 1];
            weW = vW - liW[K - 1];
            double *tt = s_tot + (par * WPC + wid) * 2;
            if (lane == 31) { tt[0] = vT; tt[1] = vW; }
            if (t == tbt) { s_bex[par * 2 + 0] = weT; s_bex[par * 2 + 1] = weW; }
        }
        // the ring refill issued in the previous iteration must have landed before anyone reads it
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();   // B2: produced states, warp totals, ring visible
        x_ready = x_fill;
        // ---- 3. prefix sums across warps, acceptance and boundary candidates
        double totT, totW, baseT, baseW;   // base*: state before this thread's slot 0, relative to the base
        {
            // every warp scans the (<= 32) warp totals itself: lane q holds warp q's
            double vT = (lane < WPC) ? s_tot[(par * WPC + lane) * 2 + 0] : 0.0;
            double vW = (lane < WPC) ? s_tot[(par * WPC + lane) * 2 + 1] : 0.0;
            const double oT = vT, oW = vW;
#pragma unroll
            for (int ofs = 1; ofs < WPC; ofs <<= 1) {
                const double a = __shfl_up_sync(0xffffffffu, vT, ofs);
                const double b = __shfl_up_sync(0xffffffffu, vW, ofs);
                if (lane >= ofs) { vT += a; vW += b; }
            }
            const int wb = tbt >> 5;
            const double preT = __shfl_sync(0xffffffffu, vT - oT, wid), preW = __shfl_sync(0xffffffffu, vW - oW, wid);
            const double pbT = __shfl_sync(0xffffffffu, vT - oT, wb), pbW = __shfl_sync(0xffffffffu, vW - oW, wb);
            totT = __shfl_sync(0xffffffffu, vT, WPC - 1);
            totW = __shfl_sync(0xffffffffu, vW, WPC - 1);
            const double eT = (preT + weT) - (pbT + s_bex[par * 2 + 0]);
            const double eW = (preW + weW) - (pbW + s_bex[par * 2 + 1]);
            baseT = (t < tbt) ? eT + totT : eT;
            baseW = (t < tbt) ? eW + totW : eW;
        }
        int firstbad = NT, stopr = NT, entr = NT, ckr = NT;
#pragma unroll
        for (int k = K - 1; k >= 0; k--) {
            const int r = lr + k;
            int pii;
            float2 pmo;
            if (k >= 1) {
                pii = valid[k - 1] ? oii[k - 1] : INVALID;
                pmo = make_float2(omu[k - 1], oom[k - 1]);
            } else {
                pii = soi[(t * K + NT - 1) & (NT - 1)];
                pmo = som[(t * K + NT - 1) & (NT - 1)];
            }
            const bool ok = (r == 0) || ((bii[k] == pii) & (bmu[k] == pmo.x) & (bom[k] == pmo.y));
            if (!ok || !valid[k]) firstbad = r;
            if (bii[k] > last_ok || bii[k] >= seg1) stopr = r;
            if (bii[k] >= seg0) entr = r;
            if (bii[k] >= next_ck) ckr = r;
        }
        firstbad = __reduce_min_sync(0xffffffffu, firstbad);
        stopr = __reduce_min_sync(0xffffffffu, stopr);
        entr = __reduce_min_sync(0xffffffffu, entr);
        ckr = __reduce_min_sync(0xffffffffu, ckr);
        if (lane == 0) {
            int *rr = s_red + par * 4;
            if (firstbad < NT) atomicMin(rr + 0, firstbad);
            if (stopr < NT) atomicMin(rr + 1, stopr);
            if (entr < NT) atomicMin(rr + 2, entr);
            if (ckr < NT) atomicMin(rr + 3, ckr);
        }
        __syncthreads();   // B3
        // ---- 4. uniform control: every thread derives it from shared memory
        int A = s_red[par * 4 + 0], r_stop = s_red[par * 4 + 1], r_ent = s_red[par * 4 + 2], r_ck = s_red[par * 4 + 3];
        if (t < 4) s_red[(par ^ 1) * 4 + t] = NT;   // for the next iteration (ordered by its barriers)
        // exact states: ranks 0..A (rank A's is what rank A-1 produced; rank 0's is the base)
        int aii = ii_b;
        float amu = mu_b, aom = om_b;
        if (A >= 1) {
            const int sa = (tbt * K + A - 1) & (NT - 1);
            aii = soi[sa];   // valid: rank A-1 was accepted
            amu = som[sa].x;
            aom = som[sa].y;
        }
        // believed states of ranks >= A are not exact: their candidates do not count; rank A's own state does
        if (r_stop >= A) r_stop = (aii > last_ok || aii >= seg1) ? A : NT;
        if (r_ent >= A) r_ent = (aii >= seg0) ? A : NT;
        if (r_ck >= A) r_ck = (aii >= next_ck) ? A : NT;
        const bool stop = r_stop <= A;
        const int Ap = A & ~(K - 1);
        const int fin = stop ? r_stop : Ap;   // ranks [0, fin) are final; the state of rank fin is exact
        // state of rank r (r <= A) with the interpolants of the two symbols before it
        auto state_of = [&](int r) {
            MmState s;
            if (r == 0) {
                s.ii = ii_b; s.mu = mu_b; s.omega = om_b;
            } else {
                const int sr = (tbt * K + r - 1) & (NT - 1);
                s.ii = soi[sr]; s.mu = som[sr].x; s.omega = som[sr].y;
            }
            s.p0 = (r >= 1) ? sp[(tbt * K + r - 1) & (NT - 1)] : P1;
            s.p1 = (r >= 2) ? sp[(tbt * K + r - 2) & (NT - 1)] : ((r == 1) ? P1 : P2);
            return s;
        };
        int lo = 0;
        if (!have_entry) {
            if (r_ent > r_stop) r_ent = r_stop;   // empty segment: the stop symbol is also the entry
            if (r_ent <= fin) {
                lo = r_ent;
                if (t == 0) entry[j] = state_of(r_ent);
                have_entry = true;
            } else {
                lo = fin;   // still warming up
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int r = lr + k;
            if (r >= lo && r < fin) {
                const int pos = count + (r - lo);
                if (pos < cap_seg) stage[pos] = p0[k];
                else overflow = 1;
            }
        }
        if (ck && have_entry && r_ck <= fin && r_ck >= lo && next_ck < seg1 && ck_idx < ncp) {
            MmCk c;
            c.st = state_of(r_ck);
            c.count = count + (r_ck - lo);
            c.pad = 0;
            if (mode == 1) {
                // every thread reads the same words; the write waits until all have (next barrier)
                const MmCk old = ck[ck_idx];
                merged = mm_same(old.st, c.st) && old.count == c.count;
                __syncthreads();
            }
            if (!merged && t == 0) ck[ck_idx] = c;
            next_ck += C;
            ck_idx++;
        }
        count += (fin > lo) ? (fin - lo) : 0;
        if (merged) break;
        if (stop) {
            if (t == 0) exit_[j] = state_of(r_stop);
            break;
        }
        // ---- 5. slide by fin (a multiple of K): new base, next believed states
        if (fin >= 2) {
            P2 = sp[(tbt * K + fin - 2) & (NT - 1)];
            P1 = sp[(tbt * K + fin - 1) & (NT - 1)];
        } else if (fin == 1) {
            P2 = P1;
            P1 = sp[(tbt * K) & (NT - 1)];
        }
        // state after the whole window (for the freed slots), from the old base
        const double endT = (double)mu_b + totT, endW = (double)om_b + totW;
        const bool freed = lr < fin;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int r = lr + k;
            if (freed) {
                const double Tr = endT + (double)r * endW;
                const double fl = floor(Tr);
                bii[k] = ii_b + (int)fl;
                bmu[k] = (float)(Tr - fl);
                bom[k] = (float)endW;
            } else if (r > A) {
                const double Tr = (double)mu_b + baseT + (k ? liT[k - 1] : 0.0);
                const double fl = floor(Tr);
                bii[k] = ii_b + (int)fl;
                bmu[k] = (float)(Tr - fl);
                bom[k] = (float)((double)om_b + baseW + (k ? liW[k - 1] : 0.0));
            } else if (r == A) {
                bii[k] = aii;
                bmu[k] = amu;
                bom[k] = aom;
            }   // r < A: verified, keep
        }
        if (fin >= 1) {
            const int sr = (tbt * K + fin - 1) & (NT - 1);
            ii_b = soi[sr];
            mu_b = som[sr].x;
            om_b = som[sr].y;
        }
        tbt = (tbt + fin / K) & (T - 1);
        par ^= 1;
        {
            const int target = ((ii_b < lo_min ? lo_min : ii_b) & ~31) + R;
            for (int i = x_fill + t; i < target; i += T)
                if (i >= lo_min && i < n) cp_async8(&s_x[i & RM], in + i);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            if (target > x_fill) x_fill = target;
        }
    }
    overflow = __syncthreads_or(overflow);
    if (t == 0) {
        MmSegOut so;
        if (merged) {
            so = segout[j];
            so.overflow |= overflow;
            so.iters += iters;
            so.windows += iters;
        } else {
            so.n_sym = count;
            so.overflow = overflow;
            so.iters = iters;
            so.windows = iters;
        }
        segout[j] = so;
    }
}

}  // namespace xrd
