// xrd_wn.cuh -- "window Newton" runner for the sample-rate feedback loops (AGC, Costas).
//
// A feedback loop is a literal scalar recurrence s[i+1] = F(s[i], x[i]) in FP32 (LOOP::step, the
// same function the one-thread-per-segment kernel uses).  One warp advances ONE such chain, but
// NT = 32*K samples at a time:
//
//   * every slot of a ring of NT consecutive samples holds a BELIEVED state for its sample;
//   * each iteration all slots apply the literal step to their believed state: o = F(s, x);
//   * the state differences o - s are taken exactly (FP64 differences of FP32 states) and prefix-summed in
//     sample order from the exact state of the base sample; the sums are the next believed states
//     (if every believed state up to a slot was true, the sum telescopes to the true state of the
//     next slot exactly -- the truth is a fixed point of this map, and because F contracts and FP32
//     rounding snaps nearby states together, the iteration converges to it from a linear
//     extrapolation in a handful of rounds, ~30 samples per round at NT = 128; a certified re-run
//     starts from the trajectory the pass before recorded instead, see wn_run_cta);
//   * acceptance is literal: slot r is exact iff slot r-1 is exact and believed[r] == o[r-1]
//     bit for bit.  The leading exact run of A >= 1 samples is emitted, the base moves to sample A
//     with the literal state o[A-1], and the freed slots re-enter at the far end of the window.
//
// Proposals may be arbitrarily wrong (overflow, wrong 2*pi branch, unrepresentable tiny values):
// that only costs iterations, never correctness, because nothing is accepted without the
// literal check.  The result is the sequential trajectory, exactly.
#pragma once
#include "xrd_kernels.cuh"

namespace xrd {

// ---- exact wide views of the loop states: doubles (differences of two FP32 states and their
// running sums are exact in FP64 for every value the loops normally take; B200 runs DADD at half
// the FP32 rate, and an inexact sum only costs an iteration) ----
struct AgcWn {
    static constexpr int NS = 1;
    static constexpr bool GUIDE = false;   // AGC hand-offs almost always certify at once: nothing to guide
    __device__ static __forceinline__ void widen(const AgcState &s, double *f) { f[0] = (double)s.gain; }
    __device__ static __forceinline__ AgcState narrow(const double *f)
    {
        AgcState s;
        s.gain = (float)f[0];
        s.pad = 0.f;
        return s;
    }
    __device__ static __forceinline__ bool in_range(const AgcState &) { return true; }
    __device__ static __forceinline__ void normalise(double *) {}
    __device__ static __forceinline__ void extrapolate(const double *e, int, double *o) { o[0] = e[0]; }
};

struct CostasWn {
    static constexpr int NS = 2;
    static constexpr bool GUIDE = true;
    __device__ static __forceinline__ void widen(const CostasState &s, double *f)
    {
        f[0] = (double)s.phase;
        f[1] = (double)s.freq;
    }
    __device__ static __forceinline__ CostasState narrow(const double *f)
    {
        CostasState s;
        s.phase = (float)f[0];
        s.freq = (float)f[1];
        return s;
    }
    // a believed phase outside the literal loop's range (a sum across a turn that is not settled yet, an
    // extrapolation past +-2*pi) is brought back by whole turns, as the loop itself would have
    __device__ static __forceinline__ bool in_range(const CostasState &s) { return fabsf(s.phase) <= 6.2831855f; }
    __device__ static __forceinline__ void normalise(double *f)
    {
        const double T = 6.283185307179586;
        double p = f[0];
        p -= T * rint(p * (0.5 / T) * 0.999999);   // whole turns towards zero-ish; |p| <= 2*pi afterwards for sane p
        p = (p > T) ? p - T : p;
        p = (p < -T) ? p + T : p;
        // a NaN phase (non-finite input) stays NaN: that IS the literal loop's state from then on
        f[0] = (p != p) ? p : ((p > T || p < -T) ? 0.0 : p);
        f[1] = (f[1] != f[1]) ? f[1] : fmin(fmax(f[1], -1.0), 1.0);
    }
    __device__ static __forceinline__ void extrapolate(const double *e, int r, double *o)
    {
        o[0] = fma((double)r, e[1], e[0]);
        o[1] = e[1];
    }
};

template <class LOOP> struct WnOf;
template <> struct WnOf<AgcLoop> { typedef AgcWn type; };
template <> struct WnOf<CostasLoopK> { typedef CostasWn type; };

constexpr int WN_WARPS = 4;   // independent chains per CTA

template <int K> __host__ __device__ constexpr int wn_ring() { return 32 * K * 4; }   // samples per warp ring (power of two)
template <int K> __host__ __device__ constexpr size_t wn_smem_bytes() { return sizeof(float2) * wn_ring<K>() * WN_WARPS; }

// checkpoint use of a run
enum { WN_CK_NONE = 0, WN_CK_RECORD = 1, WN_CK_COMPARE = 2 };

template <class State> __device__ __forceinline__ State wn_shfl(const State &v, int src)
{
    // State is a pair of 32-bit words in both loops
    float2 t = *reinterpret_cast<const float2 *>(&v);
    t.x = __shfl_sync(0xffffffffu, t.x, src);
    t.y = __shfl_sync(0xffffffffu, t.y, src);
    return *reinterpret_cast<State *>(&t);
}

// Runs samples [s_begin, s_end) of x (segment-relative indices, s_begin a multiple of K;
// x[s_begin..s_end) addressable) from the exact state `st`, returns the exact state after sample
// s_end-1.  WRITE: outputs go to y (s_begin >= 0).  CK: exact states before samples that are
// multiples of C (a power of two >= 32*K; 0 < i < s_end) are recorded in / compared with ck[i / C];
// in compare mode the run stops at the first checkpoint that equals the stored one (the
// trajectories have merged: everything after it is already in place) and *merged is set.
// Warp-collective.
//
// Slot lane*K + k of the ring holds one sample; the base (oldest unfinished sample) always sits in
// slot 0 of a lane (the window slides by multiples of K; up to K-1 accepted samples stay in the
// window one more round), so a slot's rank is ((lane - tbl) & 31) * K + k.
// cp.async of one raw sample (8 bytes cf32, 4 bytes s16 IQ)
template <class RAW> __device__ __forceinline__ void cp_async_raw(RAW *smem_dst, const RAW *gmem_src)
{
    static_assert(sizeof(RAW) == 8 || sizeof(RAW) == 4, "ring samples are staged with 8- or 4-byte cp.async");
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(RAW) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}

// IN: ingest format of x (InF32, or InS16 when the AGC is the first kernel of the chain and converts as it loads)
template <class LOOP, int K, bool WRITE, int CK, class IN = InF32>
__device__ __forceinline__ typename LOOP::State wn_run(const typename IN::raw *__restrict__ x, float2 *__restrict__ y, int s_begin,
                                                       int s_end, typename LOOP::State st, const typename LOOP::Params &prm,
                                                       typename IN::raw *ring, typename LOOP::State *ck, int C, bool *merged,
                                                       unsigned long long *iters_out, typename LOOP::State *__restrict__ tr = nullptr)
{
    // tr (optional, WRITE runs): trajectory record -- tr[i] receives the exact state before sample i, which guides the
    // proposals of later certified re-runs of this segment (wn_run_cta)
    typedef typename LOOP::State State;
    typedef typename WnOf<LOOP>::type WN;
    constexpr int NS = WN::NS;
    constexpr int NT = 32 * K;
    constexpr int RS = wn_ring<K>();
    constexpr int MARGIN = 2 * NT;
    const int lane = threadIdx.x & 31;

    int base = s_begin, tbl = 0;
    State bs[K];          // believed state before the sample of slot lane*K + k
    float2 xs[K];         // that sample
    double basew[NS];
    WN::widen(st, basew);
    State base_state = st;

    // ring: sample i lives in ring[i & (RS-1)]; [fill - RS, fill) resident or in flight
    int fill = s_begin;
    {
        const int target = min(s_begin + NT + MARGIN, s_end);
        for (int i = fill + lane; i < target; i += 32) cp_async_raw(&ring[i & (RS - 1)], x + i);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        fill = max(fill, target);
        cp_async_wait_all();
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int r = lane * K + k;
        double f[NS];
        WN::extrapolate(basew, r, f);
        WN::normalise(f);
        bs[k] = (r == 0) ? st : WN::narrow(f);
        const int i = base + r;
        xs[k] = (i < s_end) ? IN::cvt(ring[i & (RS - 1)]) : make_float2(0.f, 0.f);
    }
    unsigned long long iters = 0;
    bool done_merged = false;

    while (base < s_end) {
        iters++;
        // ---- 1. literal step of every slot, exact state differences summed within the lane
        State o[K];
        float2 yo[K];
        double li[K][NS];
#pragma unroll
        for (int k = 0; k < K; k++) {
            o[k] = bs[k];
            yo[k] = LOOP::step_sel(o[k], prm, xs[k]);
            double fo[NS], fs[NS];
            WN::widen(o[k], fo);
            WN::widen(bs[k], fs);
#pragma unroll
            for (int c = 0; c < NS; c++) li[k][c] = k ? li[k - 1][c] + (fo[c] - fs[c]) : (fo[c] - fs[c]);
        }
        // ---- 2. warp scan of the lane totals (slot order); state before this lane's slot 0 if all before it is true
        double total[NS], lanebase[NS];
#pragma unroll
        for (int c = 0; c < NS; c++) {
            double v = li[K - 1][c];
#pragma unroll
            for (int ofs = 1; ofs < 32; ofs <<= 1) {
                const double a = __shfl_up_sync(0xffffffffu, v, ofs);
                if (lane >= ofs) v += a;
            }
            const double we = v - li[K - 1][c];
            total[c] = __shfl_sync(0xffffffffu, v, 31);
            const double b = __shfl_sync(0xffffffffu, we, tbl);
            lanebase[c] = basew[c] + ((lane < tbl) ? (we - b) + total[c] : (we - b));
        }
        // ---- 3. acceptance: believed[r] == o[r-1], literally, for a leading run of ranks
        State prev[K];
        prev[0] = wn_shfl(o[K - 1], (lane + 31) & 31);
#pragma unroll
        for (int k = 1; k < K; k++) prev[k] = o[k - 1];
        const int lr = ((lane - tbl) & 31) * K;   // rank of this lane's slot 0
        unsigned mybad = NT;
#pragma unroll
        for (int k = K - 1; k >= 1; k--)
            if (!LOOP::same(bs[k], prev[k])) mybad = lr + k;
        // rank 0 is the exact base whatever precedes it in the ring
        if (lr != 0 && !LOOP::same(bs[0], prev[0])) mybad = lr;
        const int A = min((int)__reduce_min_sync(0xffffffffu, mybad), s_end - base);   // >= 1 exact samples
        if (base + A >= s_end) {
            // ---- the run ends inside the window: flush, pick the state after the last sample
#pragma unroll
            for (int k = 0; k < K; k++)
                if (WRITE && lr + k < A) {
                    y[base + lr + k] = yo[k];
                    if (tr) tr[base + lr + k] = bs[k];
                }
            const int ra = A - 1;
            State sel = o[0];
#pragma unroll
            for (int k = 1; k < K; k++)
                if (k == (ra % K)) sel = o[k];
            base_state = wn_shfl(sel, (tbl + ra / K) & 31);
            // a checkpoint inside this last window (only the ragged last segment of a call has one so close to its
            // end) must describe what is in place now: a later re-run that found an older run's state there would
            // declare itself merged with a trajectory this run has just overwritten
            if (CK != WN_CK_NONE) {
                const int i0 = (s_end - 1) & ~(C - 1);
                if (i0 > base && i0 > 0) {
                    const int rc = i0 - base - 1;   // rank whose literal result is the state before sample i0 (< A)
                    State selc = o[0];
#pragma unroll
                    for (int k = 1; k < K; k++)
                        if (k == (rc % K)) selc = o[k];
                    const State cs = wn_shfl(selc, (tbl + rc / K) & 31);
                    if (lane == 0) ck[i0 >> (31 - __clz(C))] = cs;
                }
            }
            base += A;
            break;
        }
        const int Ap = A & ~(K - 1);   // slide by whole lanes
        const bool freed = lr < Ap;
        if (Ap) {
            // ---- 4. emit the samples that leave the window; exact state after them = o of rank Ap-1
            if (WRITE && freed) {
#pragma unroll
                for (int k = 0; k < K; k++) y[base + lr + k] = yo[k];
                if (tr) {
#pragma unroll
                    for (int k = 0; k < K; k++) tr[base + lr + k] = bs[k];
                }
            }
            const State nb = wn_shfl(o[K - 1], (tbl + Ap / K - 1) & 31);
            // checkpoint: the multiple of C in (base, base+Ap], if any (at most one: C >= NT)
            if (CK != WN_CK_NONE) {
                const int i0 = (base + Ap) & ~(C - 1);
                if (i0 > base && i0 > 0 && i0 < s_end) {
                    const int r0 = i0 - base;   // K .. Ap, a multiple of K
                    const State cs = wn_shfl(prev[0], (tbl + r0 / K) & 31);   // state before rank r0 (== nb when r0 == Ap)
                    bool mrg = false;
                    if (lane == 0) {
                        State *slot = ck + (i0 >> (31 - __clz(C)));
                        if (CK == WN_CK_COMPARE && LOOP::same(*slot, cs)) mrg = true;
                        else *slot = cs;
                    }
                    if (__any_sync(0xffffffffu, mrg)) {
                        done_merged = true;
                        break;
                    }
                }
            }
            base_state = nb;
            // the group issued two iterations ago covers every sample the freed slots need now
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
            __syncwarp();
        }
        // ---- 5. next believed states: freed slots are extrapolated from the state after the window, accepted
        // slots that stay keep their (verified) state, the first unaccepted slot takes the literal o[A-1], the
        // rest take the prefix sums
        // per lane: state before slot 0 (pb) and what the slots before slot k add to it (li[k-1])
        double pb[NS];
#pragma unroll
        for (int c = 0; c < NS; c++) pb[c] = lanebase[c];
        if (freed) {
            double endw[NS];
#pragma unroll
            for (int c = 0; c < NS; c++) endw[c] = basew[c] + total[c];
            WN::extrapolate(endw, lr, pb);
#pragma unroll
            for (int k = 0; k < K - 1; k++) {
                double z[NS], e[NS];
#pragma unroll
                for (int c = 0; c < NS; c++) z[c] = (c == NS - 1 && NS > 1) ? endw[c] : 0.0;   // frequency component drives the phase
                if (NS > 1) {
                    WN::extrapolate(z, k + 1, e);   // e[0] = (k+1) * endw[1], e[1] = endw[1]
                    li[k][0] = e[0];
                    li[k][NS - 1] = 0.0;
                } else {
                    li[k][0] = 0.0;
                }
            }
            // the freed slots' samples: base + NT + lr .. +K-1 (contiguous in the ring: lr, base and RS are multiples of K)
            const int i0 = base + NT + lr;
#pragma unroll
            for (int k = 0; k < K; k++) xs[k] = (i0 + k < s_end) ? IN::cvt(ring[(i0 & (RS - 1)) + k]) : make_float2(0.f, 0.f);
        }
        bool odd = false;
        State nbs[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            double f[NS];
#pragma unroll
            for (int c = 0; c < NS; c++) f[c] = k ? pb[c] + li[k - 1][c] : pb[c];
            nbs[k] = WN::narrow(f);
            odd |= !WN::in_range(nbs[k]);
        }
        if (odd) {
            // rare: a believed phase of this lane left the loop's range; bring those back by whole turns
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (!WN::in_range(nbs[k])) {
                    double f[NS];
#pragma unroll
                    for (int c = 0; c < NS; c++) f[c] = k ? pb[c] + li[k - 1][c] : pb[c];
                    WN::normalise(f);
                    nbs[k] = WN::narrow(f);
                }
            }
        }
        if (!freed && lr <= A) {
            // the one lane that holds the end of the accepted run: verified slots keep their state, the
            // first unaccepted slot takes the literal o[A-1].  Applied LAST: whatever the literal value is (a NaN
            // after a non-finite sample is "out of range" too) it must survive, or the run would never advance.
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (lr + k < A) nbs[k] = bs[k];
                else if (lr + k == A) nbs[k] = prev[k];
            }
        }
        // a whole window was accepted (the slide is by all NT slots): its first slot is the new base again and holds the
        // exact state literally, not as the (telescoped or record-guided) proposal that equals it in all but
        // pathological cases
        if (Ap == NT && lr == 0) nbs[0] = base_state;
#pragma unroll
        for (int k = 0; k < K; k++) bs[k] = nbs[k];
        if (Ap) {
            WN::widen(base_state, basew);
            tbl = (tbl + Ap / K) & 31;
            base += Ap;
            const int target = min(base + NT + MARGIN, s_end);
            for (int i = fill + lane; i < target; i += 32) cp_async_raw(&ring[i & (RS - 1)], x + i);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            fill = max(fill, target);
        }
    }
    cp_async_wait_all();
    __syncwarp();
    if (merged) *merged = done_merged;
    if (iters_out) *iters_out += iters;
    return base_state;
}

// Segment-parallel loop kernel on window-Newton chains: one warp per work item, same contract as
// seg_loop_kernel (mode 0: speculative warm-up + segment; mode 1: re-run of the listed segments from
// entry[g]), plus checkpoints of the exact state every C samples so that a re-run stops as soon as
// it has merged with the trajectory already in place, and the first pass in two halves -- mode 2:
// warm-ups only (entry states), mode 3: every segment from entry[g] -- so that the host can put
// the Costas entries on one carrier-phase branch in between (costas_resolve_kernel).
template <class LOOP, int K, class IN = InF32>
__global__ void __launch_bounds__(WN_WARPS * 32)
wn_loop_kernel(const typename IN::raw *__restrict__ in, float2 *__restrict__ out, long long n, int L, int W, int nseg, int n_work,
               typename LOOP::State *__restrict__ entry, typename LOOP::State *__restrict__ exit_,
               const typename LOOP::State *__restrict__ carried, const int *__restrict__ list,
               typename LOOP::State *__restrict__ ckpt, int ncp, int C, unsigned long long *__restrict__ iters_total,
               typename LOOP::Params prm, int mode, long long in_ch_stride, long long out_ch_stride, long long hist,
               typename LOOP::State *__restrict__ pre, int pre_len, typename LOOP::State *__restrict__ traj)
{
    extern __shared__ __align__(16) unsigned char wn_smem[];
    typedef typename LOOP::State State;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    typename IN::raw *ring = reinterpret_cast<typename IN::raw *>(wn_smem) + (size_t)wid * wn_ring<K>();
    const int w = blockIdx.x * WN_WARPS + wid;
    if (w >= n_work) return;
    const int g = (mode == 1) ? list[w] : w;
    const int ch = g / nseg, j = g - ch * nseg;
    const long long seg0 = (long long)j * L;
    const int len = (int)min((long long)L, n - seg0);
    const typename IN::raw *x = in + (size_t)ch * in_ch_stride + seg0;
    float2 *y = out + (size_t)ch * out_ch_stride + seg0;
    State *tr = traj ? traj + (size_t)ch * out_ch_stride + seg0 : nullptr;   // trajectory record, indexed like the output
    State *ck = ckpt + (size_t)g * ncp;
    unsigned long long iters = 0;
    State st;
    if (mode == 0 || mode == 2) {
        // speculative warm-up: the state at the segment start (mode 2 stops there)
        // `hist` samples of this stream before in[0] are still in place (earlier piece of the same call): warm-ups may use them
        const bool from_carried = (j == 0 || seg0 - W + hist <= 0);
        int s_begin;
        if (from_carried) {
            st = carried[ch];
            s_begin = (int)-seg0;
        } else {
            s_begin = -W;
            float2 first[16];
#pragma unroll
            for (int i = 0; i < 16; i++) first[i] = IN::cvt(__ldg(x + s_begin + i));
            st = LOOP::guess(prm, first, 16);
        }
        if (s_begin < 0) st = wn_run<LOOP, K, false, WN_CK_NONE, IN>(x, y, s_begin, 0, st, prm, ring, nullptr, C, nullptr, &iters);
        if (lane == 0) entry[g] = st;
        // mode 2, first segment of a channel: it has no warm-up (its entry is the carried state, which right after a
        // reset is not locked yet), so run its first pre_len samples here, next to the others' warm-ups, and leave the
        // state reached -- the TRUE trajectory's, acquisition included -- for the branch resolution to refer to
        if (mode == 2 && pre && j == 0 && pre_len > 0 && pre_len <= len) {
            const State p = wn_run<LOOP, K, false, WN_CK_NONE, IN>(x, y, 0, pre_len, st, prm, ring, nullptr, C, nullptr, &iters);
            if (lane == 0) pre[ch] = p;
        }
    }
    if (mode == 3) st = entry[g];
    if (mode == 0 || mode == 3) {
        st = wn_run<LOOP, K, true, WN_CK_RECORD, IN>(x, y, 0, len, st, prm, ring, ck, C, nullptr, &iters, tr);
        if (lane == 0) exit_[g] = st;
    } else if (mode == 1) {
        st = entry[g];
        bool merged = false;
        st = wn_run<LOOP, K, true, WN_CK_COMPARE, IN>(x, y, 0, len, st, prm, ring, ck, C, &merged, &iters, tr);
        if (lane == 0 && !merged) exit_[g] = st;
    }
    if (lane == 0 && iters_total) atomicAdd(iters_total, iters);
}


// ---------------------------------------------------------------------------------------
// The same chain run by a whole CTA: WPC warps share one window of NT = 32 * WPC * K samples
// (thread t holds slots t*K .. t*K+K-1).  A chain advances NT-lane iterations whose length is set by
// the dependent latency of one literal step plus a scan, so spreading the window over more warps
// (K = 1) makes the single chain ~3x faster -- which is what bounds the certified re-run rounds
// (few chains, each as long as its merge time) -- and puts 4x the warps on an SM for the same
// number of segments.  Two CTA barriers per iteration; everything exchanged between threads goes
// through double-buffered shared memory and every thread derives the (uniform) control state itself.
// ---------------------------------------------------------------------------------------
template <class LOOP, int K, int WPC, class IN = InF32> struct WnCta {
    typedef typename LOOP::State State;
    static constexpr int T = 32 * WPC;        // threads
    static constexpr int NT = T * K;          // slots
    static constexpr int RS = 4 * NT;         // ring samples
    static constexpr int NS = WnOf<LOOP>::type::NS;
    static constexpr bool GUIDE = WnOf<LOOP>::type::GUIDE;   // re-runs take their proposals from the trajectory record
    struct Shared {
        typename IN::raw ring[RS];
        State trr[GUIDE ? RS : 1]; // ring of the trajectory record in place (guided re-runs), same indexing as `ring`
        State o[2][NT];            // literal step results of every slot
        double tot[2][WPC][NS];    // warp totals of the differences
        double bex[2][NS];         // in-warp exclusive prefix of the base thread
        int bad[2][WPC];           // first unaccepted rank seen by each warp
    };
};

template <class LOOP, int K, int WPC, bool WRITE, int CK, class IN = InF32>
__device__ __forceinline__ typename LOOP::State
wn_run_cta(const typename IN::raw *__restrict__ x, float2 *__restrict__ y, int s_begin, int s_end, typename LOOP::State st,
           const typename LOOP::Params &prm, typename WnCta<LOOP, K, WPC, IN>::Shared &sh, typename LOOP::State *ck, int C,
           bool *merged, unsigned long long *iters_out, typename LOOP::State *__restrict__ tr = nullptr)
{
    // tr (optional): trajectory record of this segment, tr[i] = state before sample i as the pass before left it.
    // WRITE runs keep it up to date.  A re-run (CK == WN_CK_COMPARE) of a loop with WN::GUIDE also READS it: the
    // trajectory in place started from a warm-up and is, after a short transient, the true one shifted by a few ulps
    // (typically: same phase, frequency word a few ulps off, for tens of thousands of samples), so "record + offset at
    // the base" is a far better proposal for the slots entering the window than a linear extrapolation -- whole
    // windows are accepted per iteration instead of ~1/6 of one.  Acceptance stays literal, so the record only
    // decides how far an iteration gets.
    typedef typename LOOP::State State;
    typedef typename WnOf<LOOP>::type WN;
    typedef WnCta<LOOP, K, WPC, IN> G;
    constexpr int NS = G::NS, T = G::T, NT = G::NT, RS = G::RS, MARGIN = 2 * NT;
    constexpr bool GUIDED = G::GUIDE && CK == WN_CK_COMPARE;
    const bool guided = GUIDED && tr != nullptr;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;

    int base = s_begin, tbt = 0, par = 0;   // tbt: thread that holds the base slot (its slot 0)
    State bs[K];
    float2 xs[K];
    double basew[NS];
    WN::widen(st, basew);
    State base_state = st;
    State *pend_slot = nullptr;   // checkpoint write deferred past the next barrier (thread 0)
    State pend_val = st;

    int fill = s_begin;
    {
        const int target = min(s_begin + NT + MARGIN, s_end);
        for (int i = fill + t; i < target; i += T) {
            cp_async_raw(&sh.ring[i & (RS - 1)], x + i);
            if (GUIDED && guided) cp_async_raw(&sh.trr[i & (G::GUIDE ? RS - 1 : 0)], tr + i);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        fill = max(fill, target);
        cp_async_wait_all();
        __syncthreads();
    }
    double goff[NS];   // guided: exact base state minus the recorded state at the base
#pragma unroll
    for (int c = 0; c < NS; c++) goff[c] = 0.0;
    if (GUIDED && guided && base < s_end) {
        double fb[NS];
        WN::widen(sh.trr[base & (G::GUIDE ? RS - 1 : 0)], fb);
#pragma unroll
        for (int c = 0; c < NS; c++) goff[c] = basew[c] - fb[c];
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int r = t * K + k;
        const int i = base + r;
        double f[NS];
        if (GUIDED && guided && i < s_end) {
            WN::widen(sh.trr[i & (G::GUIDE ? RS - 1 : 0)], f);
#pragma unroll
            for (int c = 0; c < NS; c++) f[c] += goff[c];
        } else {
            WN::extrapolate(basew, r, f);
        }
        WN::normalise(f);
        bs[k] = (r == 0) ? st : WN::narrow(f);
        xs[k] = (i < s_end) ? IN::cvt(sh.ring[i & (RS - 1)]) : make_float2(0.f, 0.f);
    }
    unsigned long long iters = 0;
    bool done_merged = false;

    while (base < s_end) {
        iters++;
        // ---- 1. literal step of every slot, exact differences, in-warp scan
        State o[K];
        float2 yo[K];
        double li[K][NS];
#pragma unroll
        for (int k = 0; k < K; k++) {
            o[k] = bs[k];
            yo[k] = LOOP::step_sel(o[k], prm, xs[k]);
            sh.o[par][t * K + k] = o[k];
            double fo[NS], fs[NS];
            WN::widen(o[k], fo);
            WN::widen(bs[k], fs);
#pragma unroll
            for (int c = 0; c < NS; c++) li[k][c] = k ? li[k - 1][c] + (fo[c] - fs[c]) : (fo[c] - fs[c]);
        }
        double we[NS];
#pragma unroll
        for (int c = 0; c < NS; c++) {
            double v = li[K - 1][c];
#pragma unroll
            for (int ofs = 1; ofs < 32; ofs <<= 1) {
                const double a = __shfl_up_sync(0xffffffffu, v, ofs);
                if (lane >= ofs) v += a;
            }
            we[c] = v - li[K - 1][c];
            if (lane == 31) sh.tot[par][wid][c] = v;
            if (t == tbt) sh.bex[par][c] = we[c];
        }
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // ring samples the freed slots will need (see wn_run)
        __syncthreads();   // B1
        if (t == 0 && pend_slot) {
            *pend_slot = pend_val;
            pend_slot = nullptr;
        }
        // ---- 2. state before this thread's slot 0 if all before it is true; acceptance
        double total[NS], lanebase[NS];
        {
            const int wb = tbt >> 5;
#pragma unroll
            for (int c = 0; c < NS; c++) {
                double pre = 0.0, preb = 0.0, tt = 0.0;
#pragma unroll
                for (int q = 0; q < WPC; q++) {
                    const double v = sh.tot[par][q][c];
                    if (q < wid) pre += v;
                    if (q < wb) preb += v;
                    tt += v;
                }
                total[c] = tt;
                const double e = (pre + we[c]) - (preb + sh.bex[par][c]);
                lanebase[c] = basew[c] + ((t < tbt) ? e + tt : e);
            }
        }
        State prev[K];
        prev[0] = sh.o[par][(t * K + NT - 1) & (NT - 1)];
#pragma unroll
        for (int k = 1; k < K; k++) prev[k] = o[k - 1];
        const int lr = ((t - tbt) & (T - 1)) * K;
        unsigned mybad = NT;
#pragma unroll
        for (int k = K - 1; k >= 1; k--)
            if (!LOOP::same(bs[k], prev[k])) mybad = lr + k;
        if (lr != 0 && !LOOP::same(bs[0], prev[0])) mybad = lr;
        mybad = __reduce_min_sync(0xffffffffu, mybad);
        if (lane == 0) sh.bad[par][wid] = (int)mybad;
        __syncthreads();   // B2
        int A = s_end - base;
#pragma unroll
        for (int q = 0; q < WPC; q++) A = min(A, sh.bad[par][q]);
        if (base + A >= s_end) {
#pragma unroll
            for (int k = 0; k < K; k++)
                if (WRITE && lr + k < A) {
                    y[base + lr + k] = yo[k];
                    if (tr) tr[base + lr + k] = bs[k];
                }
            base_state = sh.o[par][(tbt * K + A - 1) & (NT - 1)];
            if (CK != WN_CK_NONE) {   // checkpoint inside the last window: see wn_run
                const int i0 = (s_end - 1) & ~(C - 1);
                if (i0 > base && i0 > 0 && t == 0) ck[i0 >> (31 - __clz(C))] = sh.o[par][(tbt * K + (i0 - base) - 1) & (NT - 1)];
            }
            base += A;
            break;
        }
        const int Ap = A & ~(K - 1);
        const bool freed = lr < Ap;
        if (Ap) {
            if (WRITE && freed) {
#pragma unroll
                for (int k = 0; k < K; k++) y[base + lr + k] = yo[k];
                if (tr) {
#pragma unroll
                    for (int k = 0; k < K; k++) tr[base + lr + k] = bs[k];
                }
            }
            const State nb = sh.o[par][(tbt * K + Ap - 1) & (NT - 1)];
            if (CK != WN_CK_NONE) {
                const int i0 = (base + Ap) & ~(C - 1);
                if (i0 > base && i0 > 0 && i0 < s_end) {
                    const State cs = sh.o[par][(tbt * K + (i0 - base) - 1) & (NT - 1)];   // state before sample i0
                    State *slot = ck + (i0 >> (31 - __clz(C)));
                    if (CK == WN_CK_COMPARE && LOOP::same(*slot, cs)) {   // every thread reads the same words
                        done_merged = true;
                        break;
                    }
                    if (t == 0) {
                        pend_slot = slot;   // written after the next barrier: nobody may still be reading it
                        pend_val = cs;
                    }
                }
            }
            base_state = nb;
        }
        // ---- 3. next believed states
        double pb[NS];
#pragma unroll
        for (int c = 0; c < NS; c++) pb[c] = lanebase[c];
        if (freed) {
            double endw[NS];
#pragma unroll
            for (int c = 0; c < NS; c++) endw[c] = basew[c] + total[c];
            WN::extrapolate(endw, lr, pb);
#pragma unroll
            for (int k = 0; k < K - 1; k++) {
                if (NS > 1) {
                    double z[NS], e[NS];
#pragma unroll
                    for (int c = 0; c < NS; c++) z[c] = (c == NS - 1) ? endw[c] : 0.0;
                    WN::extrapolate(z, k + 1, e);
                    li[k][0] = e[0];
                    li[k][NS - 1] = 0.0;
                } else {
                    li[k][0] = 0.0;
                }
            }
            const int i0 = base + NT + lr;
#pragma unroll
            for (int k = 0; k < K; k++) xs[k] = (i0 + k < s_end) ? IN::cvt(sh.ring[(i0 & (RS - 1)) + k]) : make_float2(0.f, 0.f);
            if (GUIDED && guided) {
                // the slots that enter the window believe the record shifted by the offset measured at the new base
                // (expressed through pb / li so that the code below is the same for both kinds of proposal)
                double fb[NS], f0[NS];
                WN::widen(sh.trr[(base + Ap) & (G::GUIDE ? RS - 1 : 0)], fb);
                double nbw[NS];
                WN::widen(base_state, nbw);
#pragma unroll
                for (int c = 0; c < NS; c++) goff[c] = nbw[c] - fb[c];
                if (i0 < s_end) {
                    WN::widen(sh.trr[i0 & (G::GUIDE ? RS - 1 : 0)], f0);
#pragma unroll
                    for (int c = 0; c < NS; c++) pb[c] = f0[c] + goff[c];
#pragma unroll
                    for (int k = 0; k < K - 1; k++) {
                        double fk[NS];
                        WN::widen(sh.trr[((i0 & (RS - 1)) + k + 1) & (G::GUIDE ? RS - 1 : 0)], fk);
#pragma unroll
                        for (int c = 0; c < NS; c++) li[k][c] = (i0 + k + 1 < s_end) ? fk[c] - f0[c] : 0.0;
                    }
                }
            }
        }
        bool odd = false;
        State nbs[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            double f[NS];
#pragma unroll
            for (int c = 0; c < NS; c++) f[c] = k ? pb[c] + li[k - 1][c] : pb[c];
            nbs[k] = WN::narrow(f);
            odd |= !WN::in_range(nbs[k]);
        }
        if (odd) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (!WN::in_range(nbs[k])) {
                    double f[NS];
#pragma unroll
                    for (int c = 0; c < NS; c++) f[c] = k ? pb[c] + li[k - 1][c] : pb[c];
                    WN::normalise(f);
                    nbs[k] = WN::narrow(f);
                }
            }
        }
        if (!freed && lr <= A) {   // last, so that the literal value survives whatever it is (see wn_run)
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (lr + k < A) nbs[k] = bs[k];
                else if (lr + k == A) nbs[k] = prev[k];
            }
        }
        // a whole window was accepted (the slide is by all NT slots): its first slot is the new base again and holds the
        // exact state literally, not as the (telescoped or record-guided) proposal that equals it in all but
        // pathological cases
        if (Ap == NT && lr == 0) nbs[0] = base_state;
#pragma unroll
        for (int k = 0; k < K; k++) bs[k] = nbs[k];
        if (Ap) {
            WN::widen(base_state, basew);
            tbt = (tbt + Ap / K) & (T - 1);
            base += Ap;
            const int target = min(base + NT + MARGIN, s_end);
            for (int i = fill + t; i < target; i += T) {
                cp_async_raw(&sh.ring[i & (RS - 1)], x + i);
                if (GUIDED && guided) cp_async_raw(&sh.trr[i & (G::GUIDE ? RS - 1 : 0)], tr + i);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            fill = max(fill, target);
        }
        par ^= 1;
    }
    cp_async_wait_all();
    __syncthreads();
    if (t == 0 && pend_slot) *pend_slot = pend_val;
    if (merged) *merged = done_merged;
    if (iters_out) *iters_out += iters;
    return base_state;
}

// one CTA per work item; contract and modes of wn_loop_kernel
template <class LOOP, int K, int WPC, class IN = InF32>
__global__ void __launch_bounds__(32 * WPC)
wn_cta_kernel(const typename IN::raw *__restrict__ in, float2 *__restrict__ out, long long n, int L, int W, int nseg, int n_work,
              typename LOOP::State *__restrict__ entry, typename LOOP::State *__restrict__ exit_,
              const typename LOOP::State *__restrict__ carried, const int *__restrict__ list,
              typename LOOP::State *__restrict__ ckpt, int ncp, int C, unsigned long long *__restrict__ iters_total,
              typename LOOP::Params prm, int mode, long long in_ch_stride, long long out_ch_stride, long long hist,
              const unsigned char *__restrict__ redo, typename LOOP::State *__restrict__ traj)
{
    typedef typename LOOP::State State;
    __shared__ __align__(16) typename WnCta<LOOP, K, WPC, IN>::Shared sh;
    const int w = blockIdx.x;
    if (w >= n_work) return;
    const int g = (mode == 1) ? list[w] : w;
    const int ch = g / nseg, j = g - ch * nseg;
    const long long seg0 = (long long)j * L;
    const int len = (int)min((long long)L, n - seg0);
    const typename IN::raw *x = in + (size_t)ch * in_ch_stride + seg0;
    float2 *y = out + (size_t)ch * out_ch_stride + seg0;
    State *tr = traj ? traj + (size_t)ch * out_ch_stride + seg0 : nullptr;   // trajectory record, indexed like the output
    State *ck = ckpt + (size_t)g * ncp;
    unsigned long long iters = 0;
    State st;
    if (mode == 0 || mode == 2) {
        const bool from_carried = (j == 0 || seg0 - W + hist <= 0);
        int s_begin;
        if (from_carried) {
            st = carried[ch];
            s_begin = (int)-seg0;
        } else {
            s_begin = -W;
            float2 first[16];
#pragma unroll
            for (int i = 0; i < 16; i++) first[i] = IN::cvt(__ldg(x + s_begin + i));
            st = LOOP::guess(prm, first, 16);
        }
        if (s_begin < 0)
            st = wn_run_cta<LOOP, K, WPC, false, WN_CK_NONE, IN>(x, y, s_begin, 0, st, prm, sh, nullptr, C, nullptr, &iters);
        if (threadIdx.x == 0) entry[g] = st;
    }
    if (mode == 3) st = entry[g];
    if (mode == 0 || mode == 3) {
        st = wn_run_cta<LOOP, K, WPC, true, WN_CK_RECORD, IN>(x, y, 0, len, st, prm, sh, ck, C, nullptr, &iters, tr);
        if (threadIdx.x == 0) exit_[g] = st;
    } else if (mode == 1) {
        st = entry[g];
        bool merged = false;
        st = wn_run_cta<LOOP, K, WPC, true, WN_CK_COMPARE, IN>(x, y, 0, len, st, prm, sh, ck, C, &merged, &iters, tr);
        // A re-run that reaches the end of its segment without having merged holds the exact state there, so it
        // simply keeps going into the next segment -- its exit IS that segment's true entry -- until it merges with the
        // trajectory in place, as long as nobody else is re-running that segment in this round (redo flag clear: its
        // own hand-off had been certified against the exit this chain has just replaced).  Without this every such
        // miss costs one more host round that lasts as long as its slowest chain; slow-merging streams (carrier
        // offset near zero) had 4-7 rounds.
        int gg = g, jj = j;
        while (!merged && redo && jj + 1 < nseg && !redo[gg + 1]) {
            __syncthreads();
            if (threadIdx.x == 0) {
                exit_[gg] = st;
                entry[gg + 1] = st;
            }
            gg++;
            jj++;
            x += L;
            y += L;
            if (tr) tr += L;
            ck += ncp;
            const int len2 = (int)min((long long)L, n - (long long)jj * L);
            st = wn_run_cta<LOOP, K, WPC, true, WN_CK_COMPARE, IN>(x, y, 0, len2, st, prm, sh, ck, C, &merged, &iters, tr);
        }
        if (threadIdx.x == 0 && !merged) exit_[gg] = st;
    }
    if (threadIdx.x == 0 && iters_total) atomicAdd(iters_total, iters);
}

// ---------------------------------------------------------------------------------------
// Costas: which of the two stable lock points (carrier phase, carrier phase + pi) a cold warm-up
// ends on is a coin flip, and a segment run on the wrong one has to be re-run in full.  The
// carrier phase itself is observable without the loop: arg(sum x^2) over a short block is twice
// the carrier phase, and the wrapped differences of consecutive blocks sum to its continuous
// advance across a segment.  That predicts the branch a segment's entry state must be on given its
// predecessor's, so the entries can be put on one branch before the segments are run.  A wrong
// prediction (no lock, very low SNR) only costs the re-run it would have cost anyway.
// ---------------------------------------------------------------------------------------
constexpr int CPB = 32;   // samples per carrier-phase block

// psi[b] = arg(sum of x^2 over block b); one warp per 32 blocks, coalesced 16-byte loads
__global__ void __launch_bounds__(256)
costas_block_phase_kernel(const float2 *__restrict__ in, float *__restrict__ psi, long long nblk, long long in_ch_stride,
                          long long psi_ch_stride)
{
    const int lane = threadIdx.x & 31;
    const long long wtile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // 32 blocks = 1024 samples
    const int ch = blockIdx.y;
    const long long b0 = wtile * 32;
    if (b0 >= nblk) return;
    const float2 *p2 = in + (size_t)ch * in_ch_stride + b0 * CPB;
    const float4 *p = reinterpret_cast<const float4 *>(p2);
    const bool vec = (reinterpret_cast<unsigned long long>(p2) & 15) == 0;
    const int nb = (int)min(32LL, nblk - b0);
    float mr = 0.f, mi = 0.f;
#pragma unroll 4
    for (int it = 0; it < 16; it++) {
        // float4 number it*32 + lane of the tile holds samples 2*(it*32+lane), +1: block it*2 + lane/16
        float sr = 0.f, si = 0.f;
        if (it * 2 + (lane >> 4) < nb) {
            float4 q;
            if (vec) q = __ldg(p + it * 32 + lane);
            else {
                const float2 a = __ldg(p2 + 2 * (it * 32 + lane)), b = __ldg(p2 + 2 * (it * 32 + lane) + 1);
                q = make_float4(a.x, a.y, b.x, b.y);
            }
            sr = (q.x * q.x - q.y * q.y) + (q.z * q.z - q.w * q.w);
            si = 2.f * (q.x * q.y + q.z * q.w);
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            si += __shfl_xor_sync(0xffffffffu, si, o);
        }
        // lanes 0-15 hold block 2*it, lanes 16-31 block 2*it+1; park them in lanes 2*it, 2*it+1
        const float r1 = __shfl_sync(0xffffffffu, sr, 16), i1 = __shfl_sync(0xffffffffu, si, 16);
        const float r0 = __shfl_sync(0xffffffffu, sr, 0), i0 = __shfl_sync(0xffffffffu, si, 0);
        if (lane == 2 * it) { mr = r0; mi = i0; }
        if (lane == 2 * it + 1) { mr = r1; mi = i1; }
    }
    if (lane < nb) psi[(size_t)ch * psi_ch_stride + b0 + lane] = atan2f(mi, mr);
}

// adv[g] = continuous carrier-phase advance from the start of segment g to the start of segment g+1
__global__ void __launch_bounds__(256)
costas_seg_advance_kernel(const float *__restrict__ psi, float *__restrict__ adv, int nseg, int blk_per_seg, long long nblk,
                          long long psi_ch_stride, float *__restrict__ adv0_tail, int pre_blk)
{
    // adv0_tail[ch] (optional) = the part of segment 0's advance from block pre_blk on, i.e. from where the pre-run of
    // segment 0 stopped (wn_loop_kernel, mode 2) to the start of segment 1
    __shared__ double s_part[2][8];
    const int j = blockIdx.x, ch = blockIdx.y;
    const float *ps = psi + (size_t)ch * psi_ch_stride;
    const long long b0 = (long long)j * blk_per_seg;
    double acc = 0.0, tail = 0.0;
    for (long long b = b0 + threadIdx.x; b < b0 + blk_per_seg && b + 1 < nblk; b += blockDim.x) {
        float d = ps[b + 1] - ps[b];
        d = (d > 3.14159265f) ? d - 6.28318531f : ((d < -3.14159265f) ? d + 6.28318531f : d);
        acc += (double)d;
        if (b - b0 >= pre_blk) tail += (double)d;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        tail += __shfl_xor_sync(0xffffffffu, tail, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_part[0][threadIdx.x >> 5] = acc;
        s_part[1][threadIdx.x >> 5] = tail;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, u = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            t += s_part[0][w];
            u += s_part[1][w];
        }
        adv[(size_t)ch * nseg + j] = (float)(0.5 * t);
        if (j == 0 && adv0_tail) adv0_tail[ch] = (float)(0.5 * u);
    }
}

// puts the warm-up entry states of a channel on the branch of segment 0 (which starts from the exact state)
__global__ void costas_resolve_kernel(int nseg, int L, long long W /* warm-up minus addressable history */,
                                      CostasState *__restrict__ entry, const float *__restrict__ adv,
                                      int *__restrict__ n_flipped, const CostasState *__restrict__ pre,
                                      const float *__restrict__ adv0_tail)
{
    // pre (optional): the state the TRUE trajectory reached inside segment 0 (pre-run of wn_loop_kernel, mode 2), with
    // adv0_tail the carrier advance from there to the start of segment 1: segment 1 is then referred to that instead of
    // to segment 0's entry, which right after a reset is the unlocked initial state and says nothing about the branch
    // the loop acquires on
    // flip[j] = entry j lies on the other branch than (entry j-1 advanced by adv[j-1]) -- independent of what
    // happens to j-1, because flipping j-1 by pi flips the prediction by pi too: so the branch of j relative to
    // segment 0 is the running parity of the flips, a prefix XOR done here by warp ballots
    extern __shared__ unsigned s_par[];   // parity of each 32-segment word
    const int ch = blockIdx.x;
    CostasState *e = entry + (size_t)ch * nseg;
    const float *a = adv + (size_t)ch * nseg;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nwords = (nseg + 31) / 32;
    for (int w = wid; w < nwords; w += nw) {
        const int j = w * 32 + lane;
        bool f = false;
        if (j >= 1 && j < nseg) {
            const bool exact = ((long long)j * L - W <= 0);          // ran from the carried state: on the true branch
            const float want = (j == 1 && pre) ? pre[ch].phase + adv0_tail[ch] : e[j - 1].phase + a[j - 1];
            f = !exact && (__cosf(e[j].phase - want) < 0.f);
        }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_par[w] = m;
    }
    __syncthreads();
    // exclusive parity prefix over the words (few: nseg / 32), serial in every thread's own loop below
    for (int w = wid; w < nwords; w += nw) {
        unsigned par = 0;
        for (int q = 0; q < w; q++) par ^= __popc(s_par[q]) & 1u;
        const int j = w * 32 + lane;
        const unsigned m = s_par[w];
        const unsigned upto = (lane == 31) ? m : (m & ((2u << lane) - 1u));
        const bool flipped = ((par ^ (__popc(upto) & 1u)) & 1u) != 0;
        const bool exact = (j < nseg) && ((long long)j * L - W <= 0);
        if (j >= 1 && j < nseg && flipped && !exact) {
            const float ph = e[j].phase;
            e[j].phase = (ph > 0.f) ? ph - 3.14159265358979f : ph + 3.14159265358979f;
        }
    }
    (void)n_flipped;
}

}  // namespace xrd
