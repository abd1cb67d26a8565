// xrd_kernels.cuh -- sm_100a device code of the xritdemod sample-stream hot path.
//
// Path (reference demodulator/src/demodulator.cpp:135-157): decimating FIR -> AGC -> RRC FIR
// -> Costas loop -> Mueller&Mueller clock recovery.  Every floating-point operation below is a
// single IEEE round-to-nearest mul/add/fma/sqrt in a fixed order (this TU is compiled with
// -fmad=false; fused operations are explicit fmaf), so a stage started from the same state on
// the same input reproduces the sequential CPU definition bit for bit.  That is a design
// requirement, not a nicety: M&M selects its interpolator row with rint(mu*128), so a 1-ulp
// upstream difference decorrelates the timing trajectory at the 1e-5-sample level and shows up
// as ~1e-4 RMS on the soft symbols (DESIGN.md "Why bit-exact").
//
// Parallelisation of the three feedback loops is in time: the stream is cut into segments,
// every segment is run by its own thread (AGC, Costas) or warp (M&M) from a speculative state
// after a warm-up, and hand-offs are certified bitwise (entry state of segment j == exit state
// of segment j-1); segments that fail are re-run from the exact predecessor state until the
// chain is closed.  The result is the sequential trajectory, exactly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xrd {

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float clip_bl(float x, float c)
{
    // branchless clip used by the Costas and M&M loops: 0.5*(|x+c| - |x-c|)
    float x1 = fabsf(x + c);
    float x2 = fabsf(x - c);
    x1 -= x2;
    return 0.5f * x1;
}

// 8-byte asynchronous global -> shared copy (LDGSTS)
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// NCO sine/cosine for |x| <= 2*pi (+slack): Cody-Waite by pi/2, Cephes polynomials.
// Operation order is part of the contract (see header comment).
__device__ __forceinline__ void nco_sincos(float x, float &sn, float &cs)
{
    const float q = rintf(x * 0.636619772367581343f);
    float r = fmaf(-q, 1.5703125f, x);
    r = fmaf(-q, 4.837512969970703125e-4f, r);
    r = fmaf(-q, 7.54978995489188216e-8f, r);
    const float z = r * r;
    float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float s = fmaf(ps * z, r, r);
    float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    const float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    const int n = (int)q & 3;
    float so = (n & 1) ? c : s;
    float co = (n & 1) ? s : c;
    if (n & 2) so = -so;
    if ((n + 1) & 2) co = -co;
    sn = so;
    cs = co;
}

// ---------------------------------------------------------------------------------------
// ingest: S16IQ / S8IQ -> cf32 (reference demodulator.cpp:57-70: v / 32768.f, v / 128.f) and the SpyServer
// u8 format ((v - 128) / 128.f, SpyServerFrontend.cpp:406)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void convert_kernel(const T *__restrict__ in, float *__restrict__ out, size_t n_floats, float scale, float bias)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_floats; i += stride)
        out[i] = ((float)in[i] - bias) * scale;   // small integers and a power-of-two scale: identical to the division
}

// The same conversions as loads: the first kernel of the chain (the decimator when decimation > 1, else the AGC)
// reads the raw samples and converts them in registers, so integer input never takes a pass of its own.
struct InF32 {
    typedef float2 raw;
    __device__ static __forceinline__ float2 cvt(float2 v) { return v; }
};
struct InS16 {
    typedef short2 raw;
    // v / 32768.f (demodulator.cpp:60-61); the scale is a power of two: identical to the division
    __device__ static __forceinline__ float2 cvt(short2 v) { return make_float2((float)v.x * (1.0f / 32768.f), (float)v.y * (1.0f / 32768.f)); }
};
struct InS8 {
    typedef char2 raw;
    __device__ static __forceinline__ float2 cvt(char2 v) { return make_float2((float)v.x * (1.0f / 128.f), (float)v.y * (1.0f / 128.f)); }
};
struct InU8 {
    typedef uchar2 raw;
    __device__ static __forceinline__ float2 cvt(uchar2 v)
    {
        return make_float2(((float)v.x - 128.f) * (1.0f / 128.f), ((float)v.y - 128.f) * (1.0f / 128.f));
    }
};

// ---------------------------------------------------------------------------------------
// FIR (FirFilter::Work): out[i] = sum_k taps[k] * x[i*D - k], accumulated k = 0..T-1 by fmaf
// from 0.  `in` points at x[0]; x[-1..-(T-1)] (history) must be addressable before it.
// ---------------------------------------------------------------------------------------
constexpr int FIR_THREADS = 256;
constexpr int FIR_R = 9;                       // outputs per thread (odd: conflict-free smem stride)
constexpr int FIR_TILE = FIR_THREADS * FIR_R;  // outputs per CTA

// D == 1: register-blocked sliding window.  Thread t owns outputs [t*R, t*R+R); the R-wide input
// window lives in registers and rotates by one slot per tap (indices are compile-time after
// unrolling, so the rotation costs no moves).
__global__ void __launch_bounds__(FIR_THREADS)
fir1_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, const float *__restrict__ taps, int ntaps,
            long long n_out, long long in_ch_stride, long long out_ch_stride)
{
    extern __shared__ float s_mem[];
    const int H = ntaps - 1;
    float *s_taps = s_mem;                                        // [ntaps rounded up to even]
    // [1 pad + FIR_TILE + H]: the pad absorbs the (unused) window refill after the last tap
    float2 *s_x = reinterpret_cast<float2 *>(s_mem + ((ntaps + 1) & ~1)) + 1;
    const int ch = blockIdx.y;
    in += (size_t)ch * in_ch_stride;
    out += (size_t)ch * out_ch_stride;
    const long long tile0 = (long long)blockIdx.x * FIR_TILE;
    const int tile_n = (int)min((long long)FIR_TILE, n_out - tile0);
    for (int i = threadIdx.x; i < ntaps; i += FIR_THREADS) s_taps[i] = taps[i];
    // stage x[tile0 - H, tile0 + tile_n): coalesced 8-byte loads
    for (int i = threadIdx.x; i < tile_n + H; i += FIR_THREADS) s_x[i] = __ldg(in + tile0 - H + i);
    if (threadIdx.x == 0) s_x[-1] = make_float2(0.f, 0.f);
    __syncthreads();

    const int o0 = threadIdx.x * FIR_R;           // first output of this thread (tile-relative)
    if (o0 >= tile_n) return;
    // s_x index of x[tile0 + m] is m + H
    float2 w[FIR_R], acc[FIR_R];
#pragma unroll
    for (int r = 0; r < FIR_R; r++) {
        const int m = o0 + r;
        w[r] = (m < tile_n) ? s_x[m + H] : make_float2(0.f, 0.f);
        acc[r] = make_float2(0.f, 0.f);
    }
    const float2 *xb = s_x + o0 + H;              // xb[-k-1] = next sample entering the window
    int kb = 0;
    for (; kb + FIR_R <= ntaps; kb += FIR_R) {
#pragma unroll
        for (int q = 0; q < FIR_R; q++) {
            const float h = s_taps[kb + q];
            const float2 h2 = make_float2(h, h);
#pragma unroll
            for (int r = 0; r < FIR_R; r++) {
                const int s = (r - q + FIR_R) % FIR_R;
                acc[r] = __ffma2_rn(h2, w[s], acc[r]);   // two IEEE fmaf in one FFMA2 (I and Q share the tap)
            }
            // relative index -(k+1) enters the slot that (R-1-k) leaves
            w[(FIR_R - 1 - q) % FIR_R] = xb[-(kb + q) - 1];
        }
    }
#pragma unroll
    for (int q = 0; q < FIR_R; q++) {
        if (kb + q < ntaps) {
            const float h = s_taps[kb + q];
            const float2 h2 = make_float2(h, h);
#pragma unroll
            for (int r = 0; r < FIR_R; r++) {
                const int s = (r - q + FIR_R) % FIR_R;
                acc[r] = __ffma2_rn(h2, w[s], acc[r]);   // two IEEE fmaf in one FFMA2 (I and Q share the tap)
            }
            w[(FIR_R - 1 - q) % FIR_R] = xb[-(kb + q) - 1];
        }
    }
#pragma unroll
    for (int r = 0; r < FIR_R; r++)
        if (o0 + r < tile_n) out[tile0 + o0 + r] = acc[r];
}

// generic decimation D >= 1: one output per thread-iteration straight from the staged tile
constexpr int FIRD_TILE = 1024;  // outputs per CTA
// `in` holds the samples of this call in their ingest format IN (x[0] first); the H = ntaps - 1 samples before x[0]
// come from `hist` (cf32, hist[H + g] = x[g] for g < 0), which the host carries from call to call.
template <class IN>
__global__ void __launch_bounds__(FIR_THREADS)
fird_kernel(const typename IN::raw *__restrict__ in, const float2 *__restrict__ hist, float2 *__restrict__ out,
            const float *__restrict__ taps, int ntaps, int D, long long n_out, long long in_ch_stride,
            long long out_ch_stride)
{
    extern __shared__ float s_mem[];
    const int H = ntaps - 1;
    float *s_taps = s_mem;
    float2 *s_x = reinterpret_cast<float2 *>(s_mem + ((ntaps + 1) & ~1));  // [(FIRD_TILE-1)*D + 1 + H]
    const int ch = blockIdx.y;
    in += (size_t)ch * in_ch_stride;
    hist += (size_t)ch * H;
    out += (size_t)ch * out_ch_stride;
    const long long tile0 = (long long)blockIdx.x * FIRD_TILE;
    const int tile_n = (int)min((long long)FIRD_TILE, n_out - tile0);
    const int span = (tile_n - 1) * D + 1 + H;
    for (int i = threadIdx.x; i < ntaps; i += FIR_THREADS) s_taps[i] = taps[i];
    for (int i = threadIdx.x; i < span; i += FIR_THREADS) {
        const long long g = tile0 * D - H + i;
        s_x[i] = (g >= 0) ? IN::cvt(__ldg(in + g)) : hist[H + g];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < tile_n; o += FIR_THREADS) {
        const float2 *x = s_x + o * D + H;
        float ar = 0.f, ai = 0.f;
        for (int k = 0; k < ntaps; k++) {
            const float h = s_taps[k];
            const float2 v = x[-k];
            ar = fmaf(h, v.x, ar);
            ai = fmaf(h, v.y, ai);
        }
        out[tile0 + o] = make_float2(ar, ai);
    }
}

// Taps as a kernel parameter (constant bank), already duplicated into the (h, h) pairs FFMA2 wants: a uniform constant
// load delivers the pair without a shared-memory wavefront or a register move, and FFMA2 takes it as a uniform operand.
constexpr int FT_MAX_TAPS = 256;
struct FtTaps {
    float2 h2[FT_MAX_TAPS + 8];
};

// Decimating FIR, polyphase form: out[i] = sum_m sum_p h[m*D + p] * x_p[i - m] with x_p[i] = x[i*D - p].
// Every phase is a stride-1 FIR, so the register sliding window of fir1_kernel applies per phase: a thread
// owns R consecutive outputs and D windows of R samples; taps are still applied in the order k = 0..T-1
// (m outer, p inner), so the result is the oracle's, bit for bit.  The tile is staged in shared memory
// de-interleaved by phase.
constexpr int FP_THREADS = 256;
template <int D> struct FirPoly {
    static constexpr int R = (D <= 4) ? 9 : ((D <= 6) ? 7 : 5);   // odd: conflict-free shared-memory stride
    static constexpr int TILE = FP_THREADS * R;
    __host__ __device__ static int blocks(int ntaps) { return (ntaps + D - 1) / D; }
    __host__ __device__ static size_t smem_bytes(int ntaps)
    {
        const int M = blocks(ntaps);
        return sizeof(float2) * (size_t)D * (TILE + M);
    }
};

template <int D, class IN>
__global__ void __launch_bounds__(FP_THREADS)
fird_poly_kernel(const typename IN::raw *__restrict__ in, const float2 *__restrict__ hist, float2 *__restrict__ out,
                 const __grid_constant__ FtTaps taps, int ntaps, long long n_out, long long in_ch_stride,
                 long long out_ch_stride)
{
    constexpr int R = FirPoly<D>::R, TILE = FirPoly<D>::TILE;
    extern __shared__ float s_mem[];
    const int M = FirPoly<D>::blocks(ntaps);
    const int XL = TILE + M;                                   // samples per phase: x_p[tile0 - M .. tile0 + TILE)
    float2 *s_xp = reinterpret_cast<float2 *>(s_mem);
    const int ch = blockIdx.y;
    in += (size_t)ch * in_ch_stride;
    hist += (size_t)ch * (ntaps - 1);
    out += (size_t)ch * out_ch_stride;
    const long long tile0 = (long long)blockIdx.x * TILE;
    const int tile_n = (int)min((long long)TILE, n_out - tile0);
    // stage the contiguous input span; sample g = i'*D - p goes to phase p, slot i' - (tile0 - M)
    const long long g_lo = (tile0 - M) * D - (D - 1);
    const long long g_hi = (tile0 + tile_n - 1) * D;            // last sample any output of the tile uses
    const long long g_min = tile0 * D - (ntaps - 1);            // first one (older ones only meet taps >= ntaps)
    const int span = (int)(g_hi - g_lo + 1);
    for (int s = threadIdx.x; s < span; s += FP_THREADS) {
        const long long g = g_lo + s;
        // i' = ceil(g / D) for any sign of g
        long long ip = (g >= 0) ? (g + D - 1) / D : -((-g) / D);
        const int p = (int)(ip * D - g);
        const int slot = (int)(ip - (tile0 - M));
        // samples before x[0] come from the carried history (hist[ntaps - 1 + g] = x[g])
        s_xp[p * XL + slot] = (g >= g_min) ? ((g >= 0) ? IN::cvt(__ldg(in + g)) : hist[ntaps - 1 + g]) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int o0 = threadIdx.x * R;
    if (o0 >= tile_n) return;
    float2 w[D][R], acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < D; p++) w[p][r] = s_xp[p * XL + o0 + r + M];
    }
    // whole blocks of R tap rows (every k = m * D + p below ntaps), then the ragged rest.  The trip count goes through a
    // shuffle so that the compiler knows it is warp-uniform and keeps counter and taps in uniform registers.
    const int nfull = __shfl_sync(__activemask(), (ntaps / D) / R, 0);
    int mb = 0;
    for (int ib = 0; ib < nfull; ib++, mb += R) {
#pragma unroll
        for (int q = 0; q < R; q++) {
            const int m = mb + q;
#pragma unroll
            for (int p = 0; p < D; p++) {
                const float2 h2 = taps.h2[m * D + p];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int sl = (r - q + R) % R;
                    acc[r] = __ffma2_rn(h2, w[p][sl], acc[r]);   // two IEEE fmaf in one FFMA2
                }
            }
            // x_p[o0 - m - 1] enters the slot that x_p[o0 + R - 1 - m] leaves
#pragma unroll
            for (int p = 0; p < D; p++) w[p][(R - 1 - q) % R] = s_xp[p * XL + o0 - m - 1 + M];
        }
    }
#pragma unroll
    for (int q = 0; q < R; q++) {
        const int m = mb + q;
        if (m < M) {
#pragma unroll
            for (int p = 0; p < D; p++) {
                const int k = m * D + p;
                if (k < ntaps) {
                    const float2 h2 = taps.h2[k];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int sl = (r - q + R) % R;
                        acc[r] = __ffma2_rn(h2, w[p][sl], acc[r]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < D; p++) w[p][(R - 1 - q) % R] = s_xp[p * XL + o0 - m - 1 + M];
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++)
        if (o0 + r < tile_n) out[tile0 + o0 + r] = acc[r];
}

// next call's decimator history: the last H samples of [hist | x[0..n)], converted (H <= 1024 per CTA pass)
template <class IN>
__global__ void fir_hist_carry_kernel(const typename IN::raw *__restrict__ in, float2 *__restrict__ hist, int H, long long n,
                                      long long in_ch_stride)
{
    extern __shared__ float2 s_keep[];
    in += (size_t)blockIdx.x * in_ch_stride;
    hist += (size_t)blockIdx.x * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        const long long g = n - H + i;   // index into x of the sample that becomes hist[i]
        s_keep[i] = (g >= 0) ? IN::cvt(__ldg(in + g)) : hist[H + g];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) hist[i] = s_keep[i];
}

// ---------------------------------------------------------------------------------------
// Segment-parallel feedback loops (AGC, Costas): one thread per segment.
// ---------------------------------------------------------------------------------------
struct AgcParams { float rate, ref, max_gain; };
struct AgcState { float gain; float pad; };

struct AgcLoop {
    typedef AgcParams Params;
    typedef AgcState State;
    __device__ static __forceinline__ float2 step(State &s, const Params &p, float2 x)
    {
        // AGC::Work: y = x*g; g += rate*(ref - |y|); clamp to max_gain
        float2 y;
        y.x = x.x * s.gain;
        y.y = x.y * s.gain;
        float g = s.gain + p.rate * (p.ref - sqrtf(y.x * y.x + y.y * y.y));
        if (p.max_gain > 0.0f && g > p.max_gain) g = p.max_gain;
        s.gain = g;
        return y;
    }
    __device__ static __forceinline__ float2 step_sel(State &s, const Params &p, float2 x) { return step(s, p, x); }
    // bitwise: certification is bit for bit (+0 / -0 differ), and a NaN state (non-finite input) equals itself, so
    // the window kernels keep advancing on it instead of never accepting a slot
    __device__ static __forceinline__ bool same(const State &a, const State &b)
    {
        return __float_as_uint(a.gain) == __float_as_uint(b.gain);
    }
    // speculative start: the gain that puts the mean level of the first samples on the reference
    __device__ static __forceinline__ State guess(const Params &p, const float2 *x, int m)
    {
        float acc = 0.f;
        for (int i = 0; i < m; i++) acc += sqrtf(x[i].x * x[i].x + x[i].y * x[i].y);
        State s;
        float g = (acc > 0.f) ? p.ref * (float)m / acc : 1.0f;
        if (p.max_gain > 0.0f && g > p.max_gain) g = p.max_gain;
        s.gain = g;
        s.pad = 0.f;
        return s;
    }
};

struct CostasParams { float alpha, beta, max_freq, min_freq; };
struct CostasState { float phase, freq; };

struct CostasLoopK {
    typedef CostasParams Params;
    typedef CostasState State;
    __device__ static __forceinline__ float2 step(State &s, const Params &p, float2 x)
    {
        // CostasLoop::Work, order 2: y = x*e^{-j phase}; err = clip(Re y * Im y); advance loop;
        // wrap phase to +-2pi; limit frequency
        float sn, cs;
        nco_sincos(-s.phase, sn, cs);
        float2 y;
        y.x = x.x * cs - x.y * sn;
        y.y = x.x * sn + x.y * cs;
        float err = clip_bl(y.x * y.y, 1.0f);
        float freq = s.freq + p.beta * err;
        float phase = s.phase + freq + p.alpha * err;
        // phase_wrap compares against the double 2*pi; 0x40C90FDA < 2*pi < 0x40C90FDB, so in float
        // "phase > 2*pi" is "phase > 6.28318500518798828125f".  The subtraction stays in double.
        while (phase > 6.28318500518798828125f)
            phase = (float)((double)phase - 6.283185307179586);
        while (phase < -6.28318500518798828125f)
            phase = (float)((double)phase + 6.283185307179586);
        if (freq > p.max_freq) freq = p.max_freq;
        else if (freq < p.min_freq) freq = p.min_freq;
        s.phase = phase;
        s.freq = freq;
        return y;
    }
    // The same step without branches, for the window kernel.  |freq| <= 1 after the clamp and
    // alpha < 0.83 for every loop bandwidth, so one conditional turn is exactly what the while loops
    // above do for any state inside +-2*pi (the host checks alpha + max|freq| < 2*pi before using it).
    __device__ static __forceinline__ float2 step_sel(State &s, const Params &p, float2 x)
    {
        float sn, cs;
        nco_sincos(-s.phase, sn, cs);
        float2 y;
        y.x = x.x * cs - x.y * sn;
        y.y = x.x * sn + x.y * cs;
        const float err = clip_bl(y.x * y.y, 1.0f);
        float freq = s.freq + p.beta * err;
        float phase = s.phase + freq + p.alpha * err;
        const float T = 6.28318500518798828125f;
        const double turn = (phase > T) ? -6.283185307179586 : 6.283185307179586;
        const float wrapped = (float)((double)phase + turn);
        phase = (phase > T || phase < -T) ? wrapped : phase;
        freq = (freq > p.max_freq) ? p.max_freq : ((freq < p.min_freq) ? p.min_freq : freq);
        s.phase = phase;
        s.freq = freq;
        return y;
    }
    __device__ static __forceinline__ bool same(const State &a, const State &b)
    {
        // bitwise (see AgcLoop::same)
        return (__float_as_uint(a.phase) == __float_as_uint(b.phase)) & (__float_as_uint(a.freq) == __float_as_uint(b.freq));
    }
    __device__ static __forceinline__ State guess(const Params &, const float2 *, int)
    {
        State s;
        s.phase = 0.f;
        s.freq = 0.f;
        return s;
    }
};

// Segment-parallel loop kernel: one thread per work item (segment).  A thread streams its own
// samples in 128-byte tiles (TS = 16 samples = 8 x LDG.128 into registers, issued one tile ahead so
// the HBM latency hides behind the 16 serial steps of the current tile) and writes its outputs the
// same way (8 x STG.128): every access is a full line of a private row, the serial dependence never
// leaves the register file, and there is no shared memory or barrier on the path.  Interior tiles
// run a branch-free unrolled body; only tiles that straddle a stream/segment boundary take the
// guarded path.
//
// mode 0: first pass -- work item w is segment g = w (g = ch * nseg + j).  Segment j warms up over
//         [j*L - W, j*L) from a speculative state (or starts at sample 0 from the carried exact
//         state when the warm-up would reach the stream start), records its entry state at j*L,
//         runs its segment writing outputs, records its exit state.
// mode 1: fix-up pass -- work item w is segment list[w]; it restarts at j*L from entry[g] (the
//         predecessor's exit, stored there by the verify kernel) and overwrites outputs and exit.
template <int TS> struct SegTile {
    float2 v[TS];
};

template <int TS>
__device__ __forceinline__ void seg_tile_load(SegTile<TS> &tl, const float2 *__restrict__ x, int s0, int lo, int len, bool vec)
{
    if (s0 >= lo && s0 + TS <= len) {
        if (vec) {
            const float4 *p = reinterpret_cast<const float4 *>(x + s0);
#pragma unroll
            for (int i = 0; i < TS / 2; i++) {
                const float4 q = __ldg(p + i);
                tl.v[2 * i] = make_float2(q.x, q.y);
                tl.v[2 * i + 1] = make_float2(q.z, q.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < TS; i++) tl.v[i] = __ldg(x + s0 + i);
        }
    } else {
#pragma unroll
        for (int i = 0; i < TS; i++) {
            const int sr = s0 + i;
            tl.v[i] = (sr >= lo && sr < len) ? __ldg(x + sr) : make_float2(0.f, 0.f);
        }
    }
}

template <class LOOP, int TS>
__global__ void __launch_bounds__(64)
seg_loop_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n, int L, int W, int nseg, int n_work,
                typename LOOP::State *__restrict__ entry, typename LOOP::State *__restrict__ exit_,
                const typename LOOP::State *__restrict__ carried, const int *__restrict__ list,
                typename LOOP::Params prm, int mode, long long in_ch_stride, long long out_ch_stride)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_work) return;
    const int g = (mode == 0) ? w : list[w];
    const int ch = g / nseg, j = g - ch * nseg;
    const long long seg0 = (long long)j * L;
    const int len = (int)min((long long)L, n - seg0);
    const int s_begin = (mode == 0) ? -W : 0;
    // segments whose warm-up would start before the stream take the carried exact state at sample 0
    const bool from_carried = (mode == 0) && (j == 0 || seg0 - W <= 0);
    const int lo = (mode == 0) ? (from_carried ? (int)-seg0 : -W) : 0;
    const float2 *x = in + (size_t)ch * in_ch_stride + seg0;
    float2 *y = out + (size_t)ch * out_ch_stride + seg0;
    const bool vin = ((reinterpret_cast<unsigned long long>(x) & 15) == 0);
    const bool vout = ((reinterpret_cast<unsigned long long>(y) & 15) == 0);
    typename LOOP::State st;
    if (mode == 1) st = entry[g];
    else if (from_carried) st = carried[ch];

    SegTile<TS> cur, nxt;
    seg_tile_load<TS>(cur, x, s_begin, lo, len, vin);
    if (mode == 0 && !from_carried) st = LOOP::guess(prm, cur.v, TS);
    for (int s0 = s_begin; s0 < L; s0 += TS) {
        if (s0 + TS < L) seg_tile_load<TS>(nxt, x, s0 + TS, lo, len, vin);
        if (s0 >= lo && s0 + TS <= len) {
            // interior tile: branch-free
            if (mode == 0 && s0 == 0) entry[g] = st;
#pragma unroll
            for (int e = 0; e < TS; e++) cur.v[e] = LOOP::step(st, prm, cur.v[e]);
            if (s0 >= 0) {
                if (vout) {
                    float4 *p = reinterpret_cast<float4 *>(y + s0);
#pragma unroll
                    for (int i = 0; i < TS / 2; i++)
                        p[i] = make_float4(cur.v[2 * i].x, cur.v[2 * i].y, cur.v[2 * i + 1].x, cur.v[2 * i + 1].y);
                } else {
#pragma unroll
                    for (int i = 0; i < TS; i++) y[s0 + i] = cur.v[i];
                }
            }
        } else if (s0 + TS > lo && s0 < len) {
            // boundary tile (unrolled as well: no dynamic indexing of the register tile)
#pragma unroll
            for (int e = 0; e < TS; e++) {
                const int sr = s0 + e;
                if (sr >= lo && sr < len) {
                    if (mode == 0 && sr == 0) entry[g] = st;
                    const float2 o = LOOP::step(st, prm, cur.v[e]);
                    if (sr >= 0) y[sr] = o;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < TS; e++) cur.v[e] = nxt.v[e];
    }
    exit_[g] = st;
}

// verify hand-offs: redo[j] = entry[j] != exit[j-1]; on mismatch entry[j] := exit[j-1].
// Costas only: `mirror` carries the BPSK pi-ambiguity of first-pass warm-ups (segment j's
// trajectory may be the true one rotated by pi); the relative rotation of neighbouring
// segments is observable, its prefix product gives each segment's absolute rotation, and a
// rotated predecessor exit is de-rotated before it seeds a re-run (the loop equations are
// invariant under phase+pi, y -> -y up to rounding).
template <class LOOP>
__global__ void seg_verify_kernel(int nseg, typename LOOP::State *__restrict__ entry,
                                  const typename LOOP::State *__restrict__ exit_, unsigned char *__restrict__ redo,
                                  unsigned char *__restrict__ mirror, int *__restrict__ n_redo, int *__restrict__ list,
                                  int first_round)
{
    // one CTA per channel; nseg is small (<= a few 10k): serial prefix in thread 0 for the mirror
    const int ch = blockIdx.x;
    entry += (size_t)ch * nseg;
    exit_ += (size_t)ch * nseg;
    redo += (size_t)ch * nseg;
    if (mirror) mirror += (size_t)ch * nseg;
    (void)first_round;
    if (mirror) {
        // Every round, not only the first: after a stretch without a lockable signal (dropout, burst of noise) the
        // exact trajectory re-acquires on one of the two branches and everything the first pass did behind it may sit
        // on the other one.  The re-run that crosses the stretch brings the true branch; from then on the relative
        // rotation below flags ALL segments behind it at once and they are re-run in parallel from de-rotated states,
        // instead of being discovered one per round.
        // relative rotation r_j between entry[j] and exit[j-1]; mirror[j] = xor prefix
        for (int j = threadIdx.x; j < nseg; j += blockDim.x) {
            unsigned char r = 0;
            if (j > 0) {
                float d = ((const float *)&entry[j])[0] - ((const float *)&exit_[j - 1])[0];
                r = (cosf(d) < 0.f) ? 1 : 0;
            }
            redo[j] = r;   // scratch
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned char m = 0;
            for (int j = 0; j < nseg; j++) {
                m ^= redo[j];
                mirror[j] = m;
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < nseg; j += blockDim.x) {
        unsigned char r = 0;
        if (j > 0) {
            typename LOOP::State want = exit_[j - 1];
            if (mirror && mirror[j - 1]) {
                // de-rotate by pi, staying inside (-2pi, 2pi)
                float *ph = (float *)&want;
                ph[0] = (ph[0] > 0.f) ? ph[0] - 3.14159265358979f : ph[0] + 3.14159265358979f;
            }
            if (!LOOP::same(entry[j], want)) {
                r = 1;
                entry[j] = want;
            }
        }
        redo[j] = r;
        if (r) list[atomicAdd(n_redo, 1)] = ch * nseg + j;   // work list of the fix-up pass
    }
    __syncthreads();
    if (mirror) {
        // after this round every segment either matched a de-rotated state or is re-run from one
        for (int j = threadIdx.x; j < nseg; j += blockDim.x) mirror[j] = 0;
    }
}

// ---------------------------------------------------------------------------------------
// Mueller & Mueller clock recovery (ClockRecovery::Work).
//
// The loop state is (ii, mu, omega) plus the two previous interpolants.  mu and omega stay on
// a fixed binary grid (mu+omega never leaves one binade), so the state advances by *integer*
// increments that depend on the state only through the symbol's timing-error value.  A window
// of consecutive symbols is therefore solved as a fixed point: lane i holds a believed state for
// symbol i, applies the literal sequential transition to it, the per-lane increments are
// prefix-summed exactly (fixed point, 2^-32 sample units) into new believed states, and lanes are
// accepted when their believed state did not change.  At the fixed point every lane applied the
// true transition to the true state, so the result is the sequential trajectory bit for bit;
// lane m is exact after m iterations, so termination is guaranteed for finite samples (a
// non-finite sample stops the advance of the loop itself; the chain kernels then stop at their
// iteration cap and report overflow).
// ---------------------------------------------------------------------------------------
constexpr int MM_NTAPS = 8;
constexpr int MM_NSTEPS = 128;
constexpr int MM_TAIL = 16;   // samples addressable before index 0 of the Costas output buffer

struct MmParams {
    float omega_mid, omega_lim, gain_omega, gain_mu;
};
struct MmState {
    long long ii;      // next interpolation base, relative to the current chunk (may be < 0: tail)
    float mu, omega;
    float2 p0, p1;     // interpolants of the two previous symbols (c0, c1 are their slicer values)
};

__device__ __forceinline__ bool mm_same(const MmState &a, const MmState &b)
{
    return a.ii == b.ii && __float_as_uint(a.mu) == __float_as_uint(b.mu) &&
           __float_as_uint(a.omega) == __float_as_uint(b.omega) && __float_as_uint(a.p0.x) == __float_as_uint(b.p0.x) &&
           __float_as_uint(a.p0.y) == __float_as_uint(b.p0.y) && __float_as_uint(a.p1.x) == __float_as_uint(b.p1.x) &&
           __float_as_uint(a.p1.y) == __float_as_uint(b.p1.y);
}

// s_tab: transposed MMSE table, s_tab[j * 129 + k] = taps[k][j]
__device__ __forceinline__ float2 mm_interp(const float2 *__restrict__ x, const float *__restrict__ s_tab, float mu)
{
    const int k = (int)rintf(mu * (float)MM_NSTEPS);
    float ar[4], ai[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        const float t0 = s_tab[(7 - l) * 129 + k];
        const float t1 = s_tab[(3 - l) * 129 + k];
        const float2 a = __ldg(x + l), b = __ldg(x + l + 4);
        ar[l] = fmaf(t1, b.x, t0 * a.x);
        ai[l] = fmaf(t1, b.y, t0 * a.y);
    }
    return make_float2((ar[0] + ar[1]) + (ar[2] + ar[3]), (ai[0] + ai[1]) + (ai[2] + ai[3]));
}

// literal loop update after the symbol p0 with predecessors p1, p2
__device__ __forceinline__ void mm_update(const MmParams &p, float2 p0, float2 p1, float2 p2, float &mu, float &omega,
                                          long long &ii)
{
    const float c0r = p0.x > 0.f ? 1.f : 0.f, c0i = p0.y > 0.f ? 1.f : 0.f;
    const float c1r = p1.x > 0.f ? 1.f : 0.f, c1i = p1.y > 0.f ? 1.f : 0.f;
    const float c2r = p2.x > 0.f ? 1.f : 0.f, c2i = p2.y > 0.f ? 1.f : 0.f;
    const float ar = c0r - c2r, ai = c0i - c2i;
    const float xr = ar * p1.x + ai * p1.y;
    const float br = p0.x - p2.x, bi = p0.y - p2.y;
    const float yr = br * c1r + bi * c1i;
    float mm = clip_bl(yr - xr, 1.0f);
    float om = omega + p.gain_omega * mm;
    om = p.omega_mid + clip_bl(om - p.omega_mid, p.omega_lim);
    float m = mu + om + p.gain_mu * mm;
    const float fl = floorf(m);
    ii += (long long)(int)fl;
    mu = m - fl;
    omega = om;
}

constexpr float MM_FIX = 4294967296.0f;          // 2^32
constexpr float MM_UNFIX = 2.3283064365386963e-10f;  // 2^-32

struct MmSegOut {
    int n_sym;       // symbols this segment emitted into its staging slot
    int overflow;    // staging capacity exceeded
    int iters;       // fixed-point iterations spent (diagnostic)
    int windows;
};

// checkpoint of a chain: the exact state before the first symbol at or after a sample boundary, and how many
// symbols the segment had emitted by then.  A re-run that reaches a checkpoint with both unchanged has merged
// with the trajectory already in place and stops there.
struct MmCk {
    MmState st;
    int count;
    int pad;
};

// per-symbol record of the trajectory in place: the state BEFORE every symbol a segment emitted (same indexing as the
// staging slots), in the fixed point of mm_chain32_kernel.  mm_delta_kernel re-runs a segment relative to it.
struct MmTraj {
    int4 *rec;      // x: interpolation base (sample index in the chunk), y: mu * 2^32, z: (omega - omega_mid) * 2^32,
                    // w: interpolator row rint(mu * 128)
};

// ---------------------------------------------------------------------------------------
// M&M as a CTA-wide sliding-window chain: one CTA per segment, NT lanes = NT consecutive symbols.
//
// Same fixed point as above, but the window slides: symbol s lives in lane s mod NT (a ring), every
// iteration all NT lanes apply the literal transition to their believed states, the increments are
// prefix-summed in symbol order from the exact base lane, and the leading run of lanes whose
// believed state did not change is exact and is emitted (lane r's new state is exact when lanes
// 0..r-1 were).  The freed lanes re-enter at the far end with linearly extrapolated states, so
// every lane has been refined several times by the time the exact front reaches it: the front
// advances ~200 symbols per iteration at NT = 1024 (tools/emul/mm_window_emul.c) instead of one
// symbol per ~150 cycles of a serial thread.  Samples are staged in a shared-memory ring filled by
// cp.async ahead of the lanes.
// ---------------------------------------------------------------------------------------

constexpr int MM_TAB_PAD = 1040;   // 8 * 129 floats, padded

__host__ __device__ inline size_t mm_chain_smem_bytes(int NT, int R)
{
    return sizeof(float) * MM_TAB_PAD + sizeof(float2) * NT + sizeof(long long) * (32 + 32 + 8) + sizeof(unsigned) * 8 +
           sizeof(float2) * (size_t)R;
}

template <int NT>
__global__ void __launch_bounds__(NT)
mm_chain_kernel(const float2 *__restrict__ in /* index 0 = first new sample; MM_TAIL before it valid */,
                float2 *__restrict__ stage, long long n, long long L, long long W, int nseg, long long cap_seg,
                MmState *__restrict__ entry, MmState *__restrict__ exit_, const MmState *__restrict__ carried,
                const unsigned char *__restrict__ redo, MmSegOut *__restrict__ segout, const float *__restrict__ table,
                MmParams prm, int mode, long long in_ch_stride, long long stage_ch_stride, int R, int *__restrict__ stalled)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    float *s_tab = reinterpret_cast<float *>(s_raw);
    float2 *s_p = reinterpret_cast<float2 *>(s_tab + MM_TAB_PAD);
    long long *s_wT = reinterpret_cast<long long *>(s_p + NT);
    long long *s_wW = s_wT + 32;
    long long *s_misc = s_wW + 32;   // 0,1: within-warp exclusive of the base lane; 2,3: new base; 4,5: end state; 6,7: totals
    unsigned *s_min = reinterpret_cast<unsigned *>(s_misc + 8);   // [parity*4 + {first changed, stop, entry}]
    float2 *s_x = reinterpret_cast<float2 *>(s_min + 8);
    constexpr int NW = NT / 32;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int j = blockIdx.x, ch = blockIdx.y;
    in += (size_t)ch * in_ch_stride;
    stage += (size_t)ch * stage_ch_stride + (size_t)j * cap_seg;
    entry += (size_t)ch * nseg;
    exit_ += (size_t)ch * nseg;
    segout += (size_t)ch * nseg;
    if (mode == 1 && !redo[(size_t)ch * nseg + j]) return;
    for (int i = t; i < 129 * 8; i += NT) {
        const int k = i >> 3, tp = i & 7;
        s_tab[tp * 129 + k] = table[i];
    }
    if (t < 8) s_min[t] = NT;

    const long long seg0 = (j == 0) ? -(1LL << 62) : (long long)j * L;
    const long long seg1 = (j == nseg - 1) ? (1LL << 62) : (long long)(j + 1) * L;
    const long long last_ok = n - MM_NTAPS;
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
    long long Tb = st.ii * 4294967296LL + (long long)(st.mu * MM_FIX);
    long long Wb = (long long)(st.omega * MM_FIX);
    float2 P1 = st.p0, P2 = st.p1;
    int tb = 0, count = 0, overflow = 0, iters = 0, par = 0;
    // believed state of this lane: linear extrapolation with zero timing error
    long long T = Tb + (long long)t * Wb, Wv = Wb;
    long long c_ii = -(1LL << 62);
    int c_k = -1;
    float2 p0 = make_float2(0.f, 0.f);
    const int RM = R - 1;
    const long long lo_min = -(long long)MM_TAIL;
    // ring: sample i sits in s_x[i & RM]; samples [x_fill - R, x_fill) are resident
    long long x_fill = ((st.ii < lo_min ? lo_min : st.ii) & ~31LL);
    {
        const long long target = x_fill + R;
        for (long long i = x_fill + t; i < target; i += NT)
            if (i >= lo_min && i < n) cp_async8(&s_x[i & RM], in + i);
        x_fill = target;
    }

    // every iteration accepts at least one symbol, a symbol advances the base by at least one sample and the segment
    // holds at most cap_seg symbols, so a finite stream needs fewer iterations than this; a non-finite sample stalls the
    // loop itself (the advance becomes NaN), and the chain then stops here, reports overflow and raises *stalled (the
    // host stops its re-run rounds on it)
    const long long iter_cap = cap_seg + W + 1024;
    for (;;) {
        iters++;
        if (iters > iter_cap) {
            overflow = 1;
            if (t == 0) atomicExch(stalled, 1);
            break;
        }
        cp_async_wait_all();
        __syncthreads();   // S0: ring visible, previous iteration's shared scratch consumed
        // ---- 1. interpolate at the believed state (skipped when (ii, k) did not move)
        const long long ii = T >> 32;
        const unsigned lo32 = (unsigned)(T & 0xffffffffLL);
        const float mu = (float)lo32 * MM_UNFIX;
        const float om = (float)Wv * MM_UNFIX;
        const int k = (int)rintf(mu * (float)MM_NSTEPS);
        const bool inrange = (ii >= x_fill - R) && (ii >= lo_min) && (ii + MM_NTAPS <= x_fill) && (ii <= last_ok);
        if (!inrange) {
            p0 = make_float2(0.f, 0.f);
            c_k = -1;   // never trust a cached interpolant across an out-of-ring visit
        } else if (ii != c_ii || k != c_k) {
            c_ii = ii;
            c_k = k;
            {
                const int b = (int)(ii & RM);
                float ar[4], ai[4];
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const float t0 = s_tab[(7 - l) * 129 + k];
                    const float t1 = s_tab[(3 - l) * 129 + k];
                    const float2 a = s_x[(b + l) & RM], bb = s_x[(b + l + 4) & RM];
                    ar[l] = fmaf(t1, bb.x, t0 * a.x);
                    ai[l] = fmaf(t1, bb.y, t0 * a.y);
                }
                p0 = make_float2((ar[0] + ar[1]) + (ar[2] + ar[3]), (ai[0] + ai[1]) + (ai[2] + ai[3]));
            }
        }
        s_p[t] = p0;
        __syncthreads();   // S1
        // ---- 2. literal loop update with the two preceding symbols' interpolants
        const int r = (t - tb) & (NT - 1);
        float2 p1 = s_p[(t - 1) & (NT - 1)], p2 = s_p[(t - 2) & (NT - 1)];
        if (r == 0) { p1 = P1; p2 = P2; }
        if (r == 1) { p2 = P1; }
        float mu2 = mu, om2 = om;
        long long ii2 = ii;
        mm_update(prm, p0, p1, p2, mu2, om2, ii2);
        const long long dT = (ii2 - ii) * 4294967296LL + ((long long)(mu2 * MM_FIX) - (long long)lo32);
        const long long dW = (long long)(om2 * MM_FIX) - Wv;
        // ---- 3. inclusive scan in thread order
        long long iT = dT, iW = dW;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long a = __shfl_up_sync(0xffffffffu, iT, o);
            const long long b = __shfl_up_sync(0xffffffffu, iW, o);
            if (lane >= o) { iT += a; iW += b; }
        }
        if (lane == 31) { s_wT[wid] = iT; s_wW[wid] = iW; }
        if (t == tb) { s_misc[0] = iT - dT; s_misc[1] = iW - dW; }
        __syncthreads();   // S2
        if (wid == 0) {
            long long a = (lane < NW) ? s_wT[lane] : 0, b = (lane < NW) ? s_wW[lane] : 0;
            const long long a0 = a, b0 = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long x = __shfl_up_sync(0xffffffffu, a, o);
                const long long y = __shfl_up_sync(0xffffffffu, b, o);
                if (lane >= o) { a += x; b += y; }
            }
            if (lane < NW) { s_wT[lane] = a - a0; s_wW[lane] = b - b0; }
            if (lane == 31) { s_misc[6] = a; s_misc[7] = b; }
        }
        __syncthreads();   // S3
        // ---- 4. states implied by the increments, in symbol order from the base lane
        const long long eT = s_wT[wid] + (iT - dT), eW = s_wW[wid] + (iW - dW);
        const long long bT = s_wT[tb >> 5] + s_misc[0], bW = s_wW[tb >> 5] + s_misc[1];
        const long long rotT = (t >= tb) ? eT - bT : s_misc[6] - bT + eT;
        const long long rotW = (t >= tb) ? eW - bW : s_misc[7] - bW + eW;
        const long long nT = Tb + rotT, nW = Wb + rotW;
        const long long nii = nT >> 32;
        const bool changed = (nT != T) || (nW != Wv) || !inrange;
        const bool stopc = (nii > last_ok) || (nii >= seg1);
        const bool entc = !have_entry && (nii >= seg0);
        unsigned m0 = __reduce_min_sync(0xffffffffu, changed ? (unsigned)r : (unsigned)NT);
        unsigned m1 = __reduce_min_sync(0xffffffffu, stopc ? (unsigned)r : (unsigned)NT);
        unsigned m2 = __reduce_min_sync(0xffffffffu, entc ? (unsigned)r : (unsigned)NT);
        if (lane == 0) {
            if (m0 < NT) atomicMin(&s_min[par * 4 + 0], m0);
            if (m1 < NT) atomicMin(&s_min[par * 4 + 1], m1);
            if (m2 < NT) atomicMin(&s_min[par * 4 + 2], m2);
        }
        __syncthreads();   // S4
        const int A = (int)s_min[par * 4 + 0];        // lanes r < A hold exact states and interpolants; lane A's new state is exact
        const int r_stop = (int)s_min[par * 4 + 1];
        int r_ent = (int)s_min[par * 4 + 2];
        const bool stop = (r_stop < NT) && (r_stop <= A);
        const int hi = stop ? r_stop : A;             // symbols r < hi are final
        int lo = 0;
        if (!have_entry) {
            if (r_ent > r_stop) r_ent = r_stop;       // empty segment: the stop lane is also the entry
            if (r_ent <= hi && r_ent < NT) {
                lo = r_ent;
                if (r == r_ent) {
                    float2 q1 = s_p[(t - 1) & (NT - 1)], q2 = s_p[(t - 2) & (NT - 1)];
                    if (r == 0) { q1 = P1; q2 = P2; }
                    if (r == 1) { q2 = P1; }
                    MmState s;
                    s.ii = nii;
                    s.mu = (float)(unsigned)(nT & 0xffffffffLL) * MM_UNFIX;
                    s.omega = (float)nW * MM_UNFIX;
                    s.p0 = q1;
                    s.p1 = q2;
                    entry[j] = s;
                }
                have_entry = true;
            } else {
                lo = hi;   // still warming up: nothing to emit
            }
        }
        if (r >= lo && r < hi) {
            const long long pos = (long long)count + (r - lo);
            if (pos < cap_seg) stage[pos] = p0;
            else overflow = 1;
        }
        count += (hi > lo) ? (hi - lo) : 0;
        if (stop) {
            if (r == r_stop) {
                float2 q1 = s_p[(t - 1) & (NT - 1)], q2 = s_p[(t - 2) & (NT - 1)];
                if (r == 0) { q1 = P1; q2 = P2; }
                if (r == 1) { q2 = P1; }
                MmState s;
                s.ii = nii;
                s.mu = (float)(unsigned)(nT & 0xffffffffLL) * MM_UNFIX;
                s.omega = (float)nW * MM_UNFIX;
                s.p0 = q1;
                s.p1 = q2;
                exit_[j] = s;
            }
            break;
        }
        // ---- 5. slide: the base moves to lane A, freed lanes re-enter at the far end
        if (A >= 2) {
            P2 = s_p[(tb + A - 2) & (NT - 1)];
            P1 = s_p[(tb + A - 1) & (NT - 1)];
        } else {   // A == 1 (lane 0 never changes, so A >= 1)
            P2 = P1;
            P1 = s_p[tb];
        }
        if (r == A) { s_misc[2] = nT; s_misc[3] = nW; }
        if (r == NT - 1) { s_misc[4] = nT + dT; s_misc[5] = nW + dW; }
        if (t < 4) s_min[(par ^ 1) * 4 + t] = NT;
        __syncthreads();   // S5
        const long long endT = s_misc[4], endW = s_misc[5];
        if (A < NT) { Tb = s_misc[2]; Wb = s_misc[3]; }
        else { Tb = endT; Wb = endW; }
        if (r >= A) { T = nT; Wv = nW; }
        else { T = endT + (long long)r * endW; Wv = endW; }
        tb = (tb + A) & (NT - 1);
        par ^= 1;
        // refill the ring behind the new base
        {
            const long long bii = Tb >> 32;
            const long long target = ((bii < lo_min ? lo_min : bii) & ~31LL) + R;
            for (long long i = x_fill + t; i < target; i += NT)
                if (i >= lo_min && i < n) cp_async8(&s_x[i & RM], in + i);
            if (target > x_fill) x_fill = target;
        }
    }
    overflow = __syncthreads_or(overflow);
    if (t == 0) {
        MmSegOut so;
        so.n_sym = count;
        so.overflow = overflow;
        so.iters = iters;
        so.windows = iters;
        segout[j] = so;
    }
}

// ---------------------------------------------------------------------------------------
// 32-bit variant of the chain kernel (the one normally used): identical algorithm, but the state
// is held as (int sample index, 2^-32 fraction, 2^-32 offset of omega from omega_mid) and the
// per-lane increments are scanned as 32-bit deviations from the base lane's omega; only the
// cross-warp prefix is 64-bit.  Valid when every per-symbol deviation is far below half a sample
// and n < 2^30 (checked on the host, which otherwise launches mm_chain_kernel).
// ---------------------------------------------------------------------------------------
__host__ __device__ inline size_t mm_chain32_smem_bytes(int NT, int R)
{
    return sizeof(float) * MM_TAB_PAD + sizeof(float2) * 2 * NT + sizeof(long long) * NT + sizeof(int) * (NT + 32 + 32 + 8) +
           sizeof(unsigned) * 8 + sizeof(float2) * (size_t)R;
}

template <int NT>
__global__ void __launch_bounds__(NT)
mm_chain32_kernel(const float2 *__restrict__ in, float2 *__restrict__ stage, int n, int L, int W, int nseg, int cap_seg,
                  MmState *__restrict__ entry, MmState *__restrict__ exit_, const MmState *__restrict__ carried,
                  const unsigned char *__restrict__ redo, MmSegOut *__restrict__ segout, const float *__restrict__ table,
                  MmParams prm, int mode, long long in_ch_stride, long long stage_ch_stride, int R, MmCk *__restrict__ ckpt,
                  int ncp, int C, MmTraj tr, int *__restrict__ stalled)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    float *s_tab = reinterpret_cast<float *>(s_raw);
    float2 *s_p = reinterpret_cast<float2 *>(s_tab + MM_TAB_PAD);   // [2][NT] interpolants, double buffered
    long long *s_nT = reinterpret_cast<long long *>(s_p + 2 * NT);  // [NT] new (ii, fraction) of every lane
    int *s_nW = reinterpret_cast<int *>(s_nT + NT);                 // [NT] new omega offset
    int *s_wD = s_nW + NT;                                          // [32] warp totals of dev
    int *s_wW = s_wD + 32;                                          // [32] warp totals of dw
    int *s_m32 = s_wW + 32;                                         // 0,1: base lane's in-warp exclusive (dev, dw)
    unsigned *s_min = reinterpret_cast<unsigned *>(s_m32 + 8);      // [2][4]
    float2 *s_x = reinterpret_cast<float2 *>(s_min + 8);            // [R] sample ring
    constexpr int NW = NT / 32;
    constexpr int BIG = 1 << 30;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int j = blockIdx.x, ch = blockIdx.y;
    in += (size_t)ch * in_ch_stride;
    stage += (size_t)ch * stage_ch_stride + (size_t)j * cap_seg;
    if (tr.rec) tr.rec += (size_t)ch * stage_ch_stride + (size_t)j * cap_seg;
    entry += (size_t)ch * nseg;
    exit_ += (size_t)ch * nseg;
    segout += (size_t)ch * nseg;
    if (mode == 1 && !redo[(size_t)ch * nseg + j]) return;
    for (int i = t; i < 129 * 8; i += NT) {
        const int k = i >> 3, tp = i & 7;
        s_tab[tp * 129 + k] = table[i];
    }
    if (t < 8) s_min[t] = NT;
    __shared__ int s_merged;
    if (t == 0) s_merged = 0;
    MmCk *ck = ckpt ? ckpt + ((size_t)ch * nseg + j) * ncp : nullptr;
    int next_ck = j * L + C, ck_idx = 0;   // sample boundary of the next checkpoint
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
    const float omid = prm.omega_mid;
    const long long OM = (long long)(omid * MM_FIX);   // exact: omega_mid has 24 significant bits
    // base lane (uniform): sample index, fraction, omega offset
    int ii_b = (int)st.ii;
    unsigned fr_b = (unsigned)(st.mu * MM_FIX);
    int wr_b = (int)((st.omega - omid) * MM_FIX);
    float2 P1 = st.p0, P2 = st.p1;
    int tb = 0, count = 0, overflow = 0, iters = 0, par = 0;
    // believed state of this lane
    int ii, wr = wr_b;
    unsigned fr;
    {
        const long long T = ((long long)ii_b << 32) + fr_b + (long long)t * (OM + wr_b);
        ii = (int)(T >> 32);
        fr = (unsigned)T;
    }
    int c_ii = -BIG, c_k = -1;
    float2 p0 = make_float2(0.f, 0.f);
    const int RM = R - 1;
    const int lo_min = -MM_TAIL;
    // ring: sample i sits in s_x[i & RM].  [x_ready - R, x_ready) is resident and visible to every thread;
    // [x_ready, x_fill) is in flight (cp.async group of the previous iteration).
    int x_fill = (ii_b < lo_min ? lo_min : ii_b) & ~31;
    {
        const int target = x_fill + R;
        for (int i = x_fill + t; i < target; i += NT)
            if (i >= lo_min && i < n) cp_async8(&s_x[i & RM], in + i);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        x_fill = target;
    }
    cp_async_wait_all();
    int x_ready = x_fill;
    __syncthreads();

    // iteration cap: see mm_chain_kernel (a non-finite sample stalls the loop; stop and report overflow)
    const long long iter_cap = (long long)cap_seg + W + 1024;
    for (;;) {
        iters++;
        if (iters > iter_cap) {
            overflow = 1;
            if (t == 0) atomicExch(stalled, 1);
            break;
        }
        float2 *sp = s_p + par * NT;
        // ---- 1. interpolate at the believed state (skipped when (ii, k) did not move)
        const float mu = (float)fr * MM_UNFIX;
        const float om = fmaf((float)wr, MM_UNFIX, omid);   // exact
        const int k = (int)rintf(mu * (float)MM_NSTEPS);
        const bool inrange = (ii >= x_fill - R) && (ii >= lo_min) && (ii + MM_NTAPS <= x_ready) && (ii <= last_ok);
        if (!inrange) {
            p0 = make_float2(0.f, 0.f);
            c_k = -1;
        } else if (ii != c_ii || k != c_k) {
            c_ii = ii;
            c_k = k;
            const int b = ii & RM;
            float ar[4], ai[4];
#pragma unroll
            for (int l = 0; l < 4; l++) {
                const float t0 = s_tab[(7 - l) * 129 + k];
                const float t1 = s_tab[(3 - l) * 129 + k];
                const float2 a = s_x[(b + l) & RM], bb = s_x[(b + l + 4) & RM];
                ar[l] = fmaf(t1, bb.x, t0 * a.x);
                ai[l] = fmaf(t1, bb.y, t0 * a.y);
            }
            p0 = make_float2((ar[0] + ar[1]) + (ar[2] + ar[3]), (ai[0] + ai[1]) + (ai[2] + ai[3]));
        }
        sp[t] = p0;
        __syncthreads();   // S1
        // ---- 2. literal loop update
        const int r = (t - tb) & (NT - 1);
        float2 p1 = sp[(t - 1) & (NT - 1)], p2 = sp[(t - 2) & (NT - 1)];
        if (r == 0) { p1 = P1; p2 = P2; }
        if (r == 1) { p2 = P1; }
        float mu2 = mu, om2 = om;
        long long ii2l = 0;
        mm_update(prm, p0, p1, p2, mu2, om2, ii2l);          // ii2l = floor(mu + omega' + gain_mu * mm): unused here
        const long long Wb = OM + wr_b;
        const unsigned fr2 = (unsigned)(mu2 * MM_FIX);
        // deviation of this symbol's advance from the base omega in 2^-32 samples: the advance is
        // ii2l * 2^32 + fr2 - fr, the base omega is Wb, and their difference is far below 2^31 in
        // magnitude, so its low 32 bits are the whole value
        const int dev = (int)(fr2 - fr - (unsigned)Wb);
        const int dw = (int)((om2 - omid) * MM_FIX) - wr;
        // ---- 3. inclusive scans in thread order
        int iD = dev, iWd = dw;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, iD, o);
            const int b = __shfl_up_sync(0xffffffffu, iWd, o);
            if (lane >= o) { iD += a; iWd += b; }
        }
        if (lane == 31) { s_wD[wid] = iD; s_wW[wid] = iWd; }
        if (t == tb) { s_m32[0] = iD - dev; s_m32[1] = iWd - dw; }
        __syncthreads();   // S2
        // ---- 4. cross-warp prefixes by warp-wide reductions of the warp totals (no serial second level):
        // lane l holds warp l's totals; a 32-bit total is summed as two 16-bit halves so the 64-bit sum is exact
        const int wtD = (lane < NW) ? s_wD[lane] : 0;
        const int wtW = (lane < NW) ? s_wW[lane] : 0;
        const int wb = tb >> 5;
        const int selD = (lane < wid) ? wtD : 0, selB = (lane < wb) ? wtD : 0;
        const long long preD = ((long long)__reduce_add_sync(0xffffffffu, selD >> 16) << 16) +
                               (long long)__reduce_add_sync(0xffffffffu, (unsigned)(selD & 0xffff));
        const long long preB = ((long long)__reduce_add_sync(0xffffffffu, selB >> 16) << 16) +
                               (long long)__reduce_add_sync(0xffffffffu, (unsigned)(selB & 0xffff));
        const long long totD = ((long long)__reduce_add_sync(0xffffffffu, wtD >> 16) << 16) +
                               (long long)__reduce_add_sync(0xffffffffu, (unsigned)(wtD & 0xffff));
        const int preW = __reduce_add_sync(0xffffffffu, (lane < wid) ? wtW : 0);
        const int preWB = __reduce_add_sync(0xffffffffu, (lane < wb) ? wtW : 0);
        const int totW = __reduce_add_sync(0xffffffffu, wtW);
        const long long eD = preD + (long long)(iD - dev);
        const int eW = preW + (iWd - dw);
        const long long bD = preB + (long long)s_m32[0];
        const int bW = preWB + s_m32[1];
        const long long rotD = (t >= tb) ? eD - bD : totD - bD + eD;
        const int rotW = (t >= tb) ? eW - bW : totW - bW + eW;
        const long long Tb = ((long long)ii_b << 32) + fr_b;
        const long long nT = Tb + (long long)r * Wb + rotD;
        const int nii = (int)(nT >> 32);
        const unsigned nfr = (unsigned)nT;
        const int nwr = wr_b + rotW;
        s_nT[t] = nT;
        s_nW[t] = nwr;
        const bool changed = (nii != ii) || (nfr != fr) || (nwr != wr) || !inrange;
        const bool stopc = (nii > last_ok) || (nii >= seg1);
        const unsigned m0 = __reduce_min_sync(0xffffffffu, changed ? (unsigned)r : (unsigned)NT);
        const unsigned m1 = __reduce_min_sync(0xffffffffu, stopc ? (unsigned)r : (unsigned)NT);
        if (lane == 0) {
            if (m0 < NT) atomicMin(&s_min[par * 4 + 0], m0);
            if (m1 < NT) atomicMin(&s_min[par * 4 + 1], m1);
        }
        if (!have_entry) {
            const bool entc = (nii >= seg0);
            const unsigned m2 = __reduce_min_sync(0xffffffffu, entc ? (unsigned)r : (unsigned)NT);
            if (lane == 0 && m2 < NT) atomicMin(&s_min[par * 4 + 2], m2);
        }
        if (ck) {
            const unsigned m3 = __reduce_min_sync(0xffffffffu, (nii >= next_ck) ? (unsigned)r : (unsigned)NT);
            if (lane == 0 && m3 < NT) atomicMin(&s_min[par * 4 + 3], m3);
        }
        // the ring refill issued in the previous iteration must have landed before anyone reads it
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();   // S4
        x_ready = x_fill;
        const int A = (int)s_min[par * 4 + 0];
        const int r_stop = (int)s_min[par * 4 + 1];
        int r_ent = (int)s_min[par * 4 + 2];
        const bool stop = (r_stop < NT) && (r_stop <= A);
        const int hi = stop ? r_stop : A;
        int lo = 0;
        if (!have_entry) {
            if (r_ent > r_stop) r_ent = r_stop;
            if (r_ent <= hi && r_ent < NT) {
                lo = r_ent;
                if (r == r_ent) {
                    float2 q1 = sp[(t - 1) & (NT - 1)], q2 = sp[(t - 2) & (NT - 1)];
                    if (r == 0) { q1 = P1; q2 = P2; }
                    if (r == 1) { q2 = P1; }
                    MmState s;
                    s.ii = nii;
                    s.mu = (float)nfr * MM_UNFIX;
                    s.omega = fmaf((float)nwr, MM_UNFIX, omid);
                    s.p0 = q1;
                    s.p1 = q2;
                    entry[j] = s;
                }
                have_entry = true;
            } else {
                lo = hi;
            }
        }
        if (r >= lo && r < hi) {
            const int pos = count + (r - lo);
            if (pos < cap_seg) {
                stage[pos] = p0;
                // accepted lanes did not change: (ii, fr, wr) is the exact state before this symbol
                if (tr.rec) tr.rec[pos] = make_int4(ii, (int)fr, wr, k);
            } else overflow = 1;
        }
        if (ck && have_entry) {
            // checkpoint: first symbol at or after sample next_ck, if its state is exact already (r_ck <= hi)
            const int r_ck = (int)s_min[par * 4 + 3];
            if (r_ck < NT && r_ck <= hi && r_ck >= lo && next_ck < seg1 && ck_idx < ncp) {
                if (r == r_ck) {
                    float2 q1 = sp[(t - 1) & (NT - 1)], q2 = sp[(t - 2) & (NT - 1)];
                    if (r == 0) { q1 = P1; q2 = P2; }
                    if (r == 1) { q2 = P1; }
                    MmCk c;
                    c.st.ii = nii;
                    c.st.mu = (float)nfr * MM_UNFIX;
                    c.st.omega = fmaf((float)nwr, MM_UNFIX, omid);
                    c.st.p0 = q1;
                    c.st.p1 = q2;
                    c.count = count + (r_ck - lo);
                    c.pad = 0;
                    if (mode == 1 && mm_same(ck[ck_idx].st, c.st) && ck[ck_idx].count == c.count) s_merged = 1;
                    else ck[ck_idx] = c;
                }
                if (mode == 1) {
                    __syncthreads();
                    merged = (s_merged != 0);
                }
                next_ck += C;
                ck_idx++;
            }
        }
        count += (hi > lo) ? (hi - lo) : 0;
        if (merged) break;
        if (stop) {
            if (r == r_stop) {
                float2 q1 = sp[(t - 1) & (NT - 1)], q2 = sp[(t - 2) & (NT - 1)];
                if (r == 0) { q1 = P1; q2 = P2; }
                if (r == 1) { q2 = P1; }
                MmState s;
                s.ii = nii;
                s.mu = (float)nfr * MM_UNFIX;
                s.omega = fmaf((float)nwr, MM_UNFIX, omid);
                s.p0 = q1;
                s.p1 = q2;
                exit_[j] = s;
            }
            break;
        }
        // ---- 5. slide: every thread derives the new base and the end state itself (no further barrier)
        if (A >= 2) {
            P2 = sp[(tb + A - 2) & (NT - 1)];
            P1 = sp[(tb + A - 1) & (NT - 1)];
        } else {
            P2 = P1;
            P1 = sp[tb];
        }
        // state after the last lane = base + NT base advances + all deviations
        const long long endT = Tb + (long long)NT * Wb + totD;
        const int endW = wr_b + totW;
        long long Tn;
        if (A < NT) {
            const int ta = (tb + A) & (NT - 1);
            Tn = s_nT[ta];
            wr_b = s_nW[ta];
        } else {
            Tn = endT;
            wr_b = endW;
        }
        ii_b = (int)(Tn >> 32);
        fr_b = (unsigned)Tn;
        if (r >= A) {
            ii = nii; fr = nfr; wr = nwr;
        } else {
            const long long T = endT + (long long)r * (OM + endW);
            ii = (int)(T >> 32);
            fr = (unsigned)T;
            wr = endW;
        }
        if (t < 4) s_min[(par ^ 1) * 4 + t] = NT;
        tb = (tb + A) & (NT - 1);
        par ^= 1;
        {
            const int target = ((ii_b < lo_min ? lo_min : ii_b) & ~31) + R;
            for (int i = x_fill + t; i < target; i += NT)
                if (i >= lo_min && i < n) cp_async8(&s_x[i & RM], in + i);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            if (target > x_fill) x_fill = target;
        }
    }
    overflow = __syncthreads_or(overflow);
    if (t == 0) {
        MmSegOut so;
        if (merged) {
            // the rest of the segment is in place from the earlier pass: keep its totals
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

// ---------------------------------------------------------------------------------------
// Certified re-run of a segment RELATIVE to the trajectory already in place (mm_delta_kernel).
//
// A segment whose hand-off failed ran from a warm-up state B that is a few grid units away from the
// true state A.  Two such trajectories see the same interpolator rows (ii, k) and therefore the same
// timing errors for hundreds of symbols at a time, so A's states are B's recorded states (MmTraj)
// shifted by an offset that drifts linearly: T_A[s] = T_B[s] + dT + (s - m) * dW, w_A[s] = w_B[s] + dW.
// One CTA walks the segment in windows of NT symbols: lane r believes that shifted state for symbol
// m + r, reuses B's interpolant when (ii, k) agree (fresh interpolation from the sample ring
// otherwise), applies the LITERAL loop update and compares the result, bit for bit, with what lane
// r + 1 believes.  The window is accepted up to the first lane whose belief is not its predecessor's
// result -- literal acceptance from an exact base, as in the window-Newton kernels, so the result is
// the sequential trajectory exactly; the hypothesis only decides how far one iteration gets.  At a
// break the offsets are re-fitted from the jumps measured at the next three lanes (one interpolant
// enters three consecutive timing errors).  The walk stops when A is bitwise B (state and the two
// interpolants of the history): the rest of the segment is in place already.  Staging slots and
// trajectory records are patched in place.  A walk that advances too slowly (B was not close: cold
// start, cycle slip, symbol indices out of step) gives up and leaves redo[j] set; the host then runs
// mm_chain32_kernel (mode 1) on what is left flagged.
//
// Records, interpolants in place and input samples are streamed into shared-memory rings with
// cp.async, three groups in flight (DRAM latency spans several iterations).
// ---------------------------------------------------------------------------------------
constexpr long long MM_NOSTATE = (long long)0x8000000000000000ULL;
constexpr int MM_DELTA_RING = 8;   // ring capacity of mm_delta_kernel in windows (>= 6: three groups in flight)
// RX: capacity of the sample ring (power of two, >= 5 windows of samples: see the in-flight accounting in the kernel)
__host__ __device__ inline size_t mm_delta_smem_bytes(int NT, int RX)
{
    return (size_t)MM_DELTA_RING * NT * (sizeof(int4) + sizeof(float2)) + sizeof(float2) * (size_t)RX;
}

template <int NT>
__global__ void __launch_bounds__(NT)
mm_delta_kernel(const float2 *__restrict__ in, float2 *__restrict__ stage, MmTraj tr, int n, int L, int nseg, int cap_seg,
                const MmState *__restrict__ entry, MmState *__restrict__ exit_, unsigned char *__restrict__ redo,
                MmSegOut *__restrict__ segout, const float *__restrict__ table, MmParams prm, long long in_ch_stride,
                long long stage_ch_stride, MmCk *__restrict__ ckpt, int ncp, int *__restrict__ n_bail, int RX)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    constexpr int R = MM_DELTA_RING * NT, RM = R - 1;
    int4 *s_tr = reinterpret_cast<int4 *>(s_dyn);        // [R] ring of trajectory records, symbol s in slot s & RM
    float2 *s_pb = reinterpret_cast<float2 *>(s_tr + R);  // [R] ring of the interpolants in place
    float2 *s_x = s_pb + R;                               // [RX] ring of input samples, sample i in slot i & RXM
    const int RXM = RX - 1;
    __shared__ float s_tab[MM_TAB_PAD];
    __shared__ float2 s_p[2][NT];              // interpolants, double buffered
    __shared__ long long s_nT[NT], s_jT[NT];   // update result of lane r / its jump against the belief of lane r + 1
    __shared__ int s_nW[NT], s_jW[NT];
    __shared__ unsigned char s_same[NT];       // interpolant bitwise equal to the one in place
    __shared__ unsigned s_min[2][2];
    __shared__ int s_fresh;                    // interpolations that could not be taken from the trajectory (diagnostic)
    constexpr int BIG = 1 << 30;
    const int t = threadIdx.x;
    const int j = blockIdx.x, ch = blockIdx.y;
    redo += (size_t)ch * nseg;
    if (!redo[j]) return;
    in += (size_t)ch * in_ch_stride;
    {
        const size_t o = (size_t)ch * stage_ch_stride + (size_t)j * cap_seg;
        stage += o;
        tr.rec += o;
    }
    entry += (size_t)ch * nseg;
    exit_ += (size_t)ch * nseg;
    segout += (size_t)ch * nseg;
    for (int i = t; i < 129 * 8; i += NT) {
        const int k = i >> 3, tp = i & 7;
        s_tab[tp * 129 + k] = table[i];
    }
    if (t < 4) (&s_min[0][0])[t] = NT;
    if (t == 0) s_fresh = 0;
    const int seg1 = (j == nseg - 1) ? BIG : (j + 1) * L;
    const int last_ok = n - MM_NTAPS;
    const int lo_min = -MM_TAIL;
    const float omid = prm.omega_mid;
    const int countB = min(segout[j].n_sym, cap_seg);
    const MmState st = entry[j];
    // exact base (uniform): state before symbol m, interpolants of symbols m-1 and m-2
    long long Tb = ((long long)(int)st.ii << 32) + (unsigned)(st.mu * MM_FIX);
    int wb = (int)((st.omega - omid) * MM_FIX);
    float2 P1 = st.p0, P2 = st.p1;
    bool same1 = false, same2 = false;
    int m = 0, par = 0, iters = 0, overflow = 0, fresh = 0;
    // Rings.  xf[0] = samples requested so far, xf[4] = samples every thread may read in this iteration (requested
    // five slides ago: three groups stay in flight, and a thread's wait is published by the barriers of the iteration
    // after it).  Records are requested R symbols ahead of the base, which covers the same lag (R = 8 NT).
    int fill = 0;
    int xf[5];
    xf[0] = ((int)st.ii < lo_min ? lo_min : (int)st.ii) & ~31;
    // one block of NT records (a slide frees at most NT slots) and whole blocks of NT samples per call
    auto refill = [&](int base_sym, int ii_base) {
        if (fill + NT <= base_sym + R) {
            const int s = fill + t;
            if (s < countB) {
                cp_async16(&s_tr[s & RM], tr.rec + s);
                cp_async8(&s_pb[s & RM], stage + s);
            }
            fill += NT;
        }
        const int xt = (ii_base & ~31) + RX;
        while (xf[0] + NT <= xt) {
            const int i = xf[0] + t;
            if (i >= lo_min && i < n) cp_async8(&s_x[i & RXM], in + i);
            xf[0] += NT;
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    {
        const int x0 = xf[0];
        for (int q = 0; q < MM_DELTA_RING; q++) refill(0, x0);   // one block of records per call until fill == R
    }
    xf[1] = xf[2] = xf[3] = xf[4] = xf[0];
    cp_async_wait_all();
    __syncthreads();
    // offsets of the belief: lane r >= 2 believes (TB + dT + r * dW, wB + dW), lane 1 (TB + d1T, wB + d1W), lane 0 is
    // the exact base
    long long dT = 0, d1T = 0;
    int dW = 0, d1W = 0;
    if (countB > 0) {
        const int4 b0 = s_tr[0];
        dT = Tb - (((long long)b0.x << 32) + (unsigned)b0.y);
        dW = wb - b0.z;
        d1T = dT + dW;
        d1W = dW;
    }
    bool bail = false, done = false;

    for (;;) {
        // merged with the trajectory in place: same state before symbol m, same history
        if (m >= 2 && same1 && same2 && m < countB) {
            const int4 b = s_tr[m & RM];
            if ((((long long)b.x << 32) + (unsigned)b.y) == Tb && b.z == wb) break;
        }
        if (iters > 64 + (m >> 3)) { bail = true; break; }
        iters++;
        float2 *sp = s_p[par];
        // ---- 1. believed state and its interpolant
        const int s = m + t;
        const bool hasB = s < countB;
        const int4 b = s_tr[s & RM];
        const float2 p0B = s_pb[s & RM];
        const long long TB = ((long long)b.x << 32) + (unsigned)b.y;
        long long T = TB + ((t == 1) ? d1T : dT + (long long)t * dW);
        int w = b.z + ((t == 1) ? d1W : dW);
        if (t == 0) { T = Tb; w = wb; }
        bool valid = hasB || (t == 0);
        const int ii = (int)(T >> 32);
        const unsigned fr = (unsigned)T;
        const float mu = (float)fr * MM_UNFIX;
        const float om = fmaf((float)w, MM_UNFIX, omid);   // exact
        const int k = (int)rintf(mu * (float)MM_NSTEPS);
        const bool stopc = valid && ((ii > last_ok) || (ii >= seg1));
        const bool match = hasB && ii == b.x && k == b.w;
        float2 p0 = p0B;
        if (!match) {
            if (valid && !stopc && ii >= lo_min && ii >= xf[0] - RX && ii + MM_NTAPS <= xf[4]) {
                float ar[4], ai[4];
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const float t0 = s_tab[(7 - l) * 129 + k];
                    const float t1 = s_tab[(3 - l) * 129 + k];
                    const float2 a = s_x[(ii + l) & RXM], bb = s_x[(ii + l + 4) & RXM];
                    ar[l] = fmaf(t1, bb.x, t0 * a.x);
                    ai[l] = fmaf(t1, bb.y, t0 * a.y);
                }
                p0 = make_float2((ar[0] + ar[1]) + (ar[2] + ar[3]), (ai[0] + ai[1]) + (ai[2] + ai[3]));
                fresh++;
            } else {
                p0 = make_float2(0.f, 0.f);
                if (!stopc) valid = false;
            }
        }
        sp[t] = p0;
        __syncthreads();   // S1
        // ---- 2. literal loop update, compared with the belief of the next lane
        float2 p1 = P1, p2 = P2;
        if (t >= 1) { p1 = sp[t - 1]; p2 = P1; }
        if (t >= 2) p2 = sp[t - 2];
        float mu2 = mu, om2 = om;
        long long adv = 0;
        mm_update(prm, p0, p1, p2, mu2, om2, adv);
        const long long nT = ((long long)(ii + (int)adv) << 32) + (unsigned)(mu2 * MM_FIX);
        const int nw = (int)((om2 - omid) * MM_FIX);
        long long jT = MM_NOSTATE;
        int jW = 0;
        bool next_ok = false;
        if (valid && s + 1 < countB) {
            const int4 bn = s_tr[(s + 1) & RM];
            const long long TBn = ((long long)bn.x << 32) + (unsigned)bn.y;
            const long long Tn = TBn + ((t == 0) ? d1T : dT + (long long)(t + 1) * dW);
            const int wn = bn.z + ((t == 0) ? d1W : dW);
            jT = nT - Tn;
            jW = nw - wn;
            next_ok = (jT == 0) && (jW == 0);
        }
        s_nT[t] = nT;
        s_nW[t] = nw;
        s_jT[t] = jT;
        s_jW[t] = jW;
        s_same[t] = hasB && __float_as_uint(p0.x) == __float_as_uint(p0B.x) && __float_as_uint(p0.y) == __float_as_uint(p0B.y);
        // first lane that is not exact: lane r + 1 if its belief is not this lane's result, lane r if it could not evaluate
        const unsigned c0 = !valid ? (unsigned)t : (!next_ok ? (unsigned)(t + 1) : (unsigned)NT);
        const unsigned m0 = __reduce_min_sync(0xffffffffu, c0);
        const unsigned m1 = __reduce_min_sync(0xffffffffu, stopc ? (unsigned)t : (unsigned)NT);
        if ((t & 31) == 0) {
            if (m0 < NT) atomicMin(&s_min[par][0], m0);
            if (m1 < NT) atomicMin(&s_min[par][1], m1);
        }
        __syncthreads();   // S2
        const int fb = (int)s_min[par][0];        // first lane that is not exact (>= 1: the base always is)
        const int r_stop = (int)s_min[par][1];
        if (fb == 0) { bail = true; break; }      // the base could not be evaluated (cannot happen for a state in range)
        const bool stop = r_stop < fb;            // an exact lane is past the segment
        const int hi = stop ? r_stop : fb;
        if (t < hi) {
            if (s < cap_seg) {
                if (!match) stage[s] = p0;
                if (!hasB || T != TB || w != b.z) tr.rec[s] = make_int4(ii, (int)fr, w, k);
            } else overflow = 1;
        }
        if (stop) {
            if (t == r_stop) {
                MmState so;
                so.ii = ii;
                so.mu = mu;
                so.omega = om;
                so.p0 = p1;   // interpolants of the two symbols before this lane's
                so.p1 = p2;
                exit_[j] = so;
            }
            m += hi;
            done = true;
            break;
        }
        // ---- 3. slide by fb: lane fb - 1's result is the new exact base
        Tb = s_nT[fb - 1];
        wb = s_nW[fb - 1];
        if (fb >= 2) {
            P2 = sp[fb - 2];
            P1 = sp[fb - 1];
            same2 = s_same[fb - 2] != 0;
            same1 = s_same[fb - 1] != 0;
        } else {
            P2 = P1;
            P1 = sp[0];
            same2 = same1;
            same1 = s_same[0] != 0;
        }
        // re-fit the offsets from the jumps measured at lanes fb, fb + 1, fb + 2 (published by the lane before each)
        {
            long long a0 = s_jT[fb - 1], a1 = 0, a2 = 0;
            int b0 = s_jW[fb - 1], b1 = 0, b2 = 0;
            if (a0 == MM_NOSTATE) { a0 = 0; b0 = 0; }
            if (fb < NT && s_jT[fb] != MM_NOSTATE) { a1 = s_jT[fb]; b1 = s_jW[fb]; }
            if (fb + 1 < NT && s_jT[fb + 1] != MM_NOSTATE) { a2 = s_jT[fb + 1]; b2 = s_jW[fb + 1]; }
            // (jumps are measured against the beliefs actually held, and lanes fb + 1.. held the general one, so these
            // sums are right whichever offsets lane fb itself had)
            d1T = dT + (long long)(fb + 1) * dW + a0 + a1 + b0;
            d1W = dW + b0 + b1;
            dT = dT + (long long)fb * dW + a0 + a1 + a2 - ((long long)b1 + 2LL * (long long)b2);
            dW = dW + b0 + b1 + b2;
        }
        m += fb;
        if (t < 2) s_min[par ^ 1][t] = NT;
        par ^= 1;
        // slots of symbols below the new base are free: stream in up to m + R.  Every thread then waits for all of its
        // groups but the three newest; the barriers of the next iteration publish them, so the window after that reads
        // nothing above (base four slides ago) + R >= its own base + 2 NT (a slide is at most NT; R = 8 NT).
        xf[4] = xf[3]; xf[3] = xf[2]; xf[2] = xf[1]; xf[1] = xf[0];
        refill(m, (int)(Tb >> 32));
        asm volatile("cp.async.wait_group 3;\n" ::: "memory");
    }
    cp_async_wait_all();
    if (fresh) atomicAdd(&s_fresh, fresh);
    overflow = __syncthreads_or(overflow);
    if (bail) {
        if (t == 0) {
            atomicAdd(n_bail, 1);
            segout[j].iters += iters;
        }
        return;   // redo[j] stays set: the chain kernel re-runs this segment from entry[j]
    }
    // checkpoints recorded on the old trajectory inside the patched part no longer describe what is in place
    if (ckpt) {
        MmCk *ck = ckpt + ((size_t)ch * nseg + j) * ncp;
        for (int i = t; i < ncp; i += NT)
            if (done || ck[i].count < m) ck[i].count = -1;
    }
    if (t == 0) {
        MmSegOut so = segout[j];
        if (done) so.n_sym = m;
        so.overflow |= overflow;
        so.iters += iters;
        so.windows += s_fresh;
        segout[j] = so;
        redo[j] = 0;
    }
}

// hand-off check for M&M: redo[j] = entry[j] != exit[j-1] (then entry[j] := exit[j-1])
__global__ void mm_verify_kernel(int nseg, MmState *__restrict__ entry, const MmState *__restrict__ exit_,
                                 unsigned char *__restrict__ redo, int *__restrict__ n_redo)
{
    const int ch = blockIdx.y;
    entry += (size_t)ch * nseg;
    exit_ += (size_t)ch * nseg;
    redo += (size_t)ch * nseg;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nseg) return;
    unsigned char r = 0;
    if (j > 0) {
        const MmState want = exit_[j - 1];
        if (!mm_same(entry[j], want)) {
            r = 1;
            entry[j] = want;
        }
    }
    redo[j] = r;
    if (r) atomicAdd(n_redo, 1);
}

// int8 soft symbol (SymbolManager::process, reference SymbolManager.cpp:43-46): f = Re(s)*127, clamp to [-128, 127],
// C cast (truncation toward zero)
__device__ __forceinline__ signed char soft_i8(float re)
{
    float f = re * 127.f;
    f = f > 127.f ? 127.f : f;
    f = f < -128.f ? -128.f : f;
    return (signed char)(int)f;
}

// DiagManager's byte rule (DiagManager.cpp:37-41): val * 128, clamp to [-128, 127], C cast
__device__ __forceinline__ signed char diag_i8(float v)
{
    float f = v * 128.f;
    f = f > 127.f ? 127.f : f;
    f = f < -128.f ? -128.f : f;
    return (signed char)(int)f;
}

// Per-channel diagnostics of one call, produced by the compaction pass (it touches every symbol anyway):
//  * frame: what the reference hands to DiagManager::addSamples after every chunk -- the first min(symbols, 1024)
//    FLOATS of the interleaved complex symbol buffer (demodulator.cpp:161-163) -- already in the int8 form
//    DiagManager's thread puts on its UDP socket (DiagManager.cpp:31-47);
//  * sums over all symbols of the call for a lock / SNR estimate (the reference's GNU Radio prototype shows an
//    RMS-ratio SNR, demod_tcp_qt.py:263-298; the C++ demodulator has none).
struct MmDiag {
    double sum_abs_i, sum_sq_i, sum_sq_q;
    unsigned long long n;
    int n_frame;
    int pad;
    signed char frame[1024];
};

// gather the per-segment staging slots into the contiguous symbol stream (cf32 and / or int8 soft symbols)
__global__ void mm_compact_kernel(const float2 *__restrict__ stage, float2 *__restrict__ out, int nseg, long long cap_seg,
                                  const MmSegOut *__restrict__ segout, long long *__restrict__ offsets /* [nseg+1] */,
                                  long long out_cap, long long stage_ch_stride, long long out_ch_stride,
                                  signed char *__restrict__ out_i8, MmDiag *__restrict__ diag)
{
    const int ch = blockIdx.y;
    stage += (size_t)ch * stage_ch_stride;
    if (out) out += (size_t)ch * out_ch_stride;
    if (out_i8) out_i8 += (size_t)ch * out_ch_stride;
    segout += (size_t)ch * nseg;
    offsets += (size_t)ch * (nseg + 1);
    // blockIdx.x = segment * parts + part: every segment is copied by `parts` CTAs so that the grid fills the GPU
    const int parts = gridDim.x / nseg;
    const int j = blockIdx.x / parts, part = blockIdx.x - j * parts;
    if (j >= nseg) return;
    const long long o = offsets[j];
    const int c = segout[j].n_sym;
    const float2 *src = stage + (size_t)j * cap_seg;
    double s_abs = 0.0, s_i = 0.0, s_q = 0.0;
    unsigned cnt = 0;
    for (long long i = (long long)part * blockDim.x + threadIdx.x; i < c; i += (long long)parts * blockDim.x)
        if (o + i < out_cap) {
            const float2 v = src[i];
            if (out) out[o + i] = v;
            if (out_i8) out_i8[o + i] = soft_i8(v.x);   // the byte SymbolManager::process sends (SymbolManager.cpp:43-46)
            if (diag) {
                s_abs += (double)fabsf(v.x);
                s_i += (double)v.x * (double)v.x;
                s_q += (double)v.y * (double)v.y;
                cnt++;
                if (o + i < 512) {
                    diag[ch].frame[2 * (o + i)] = diag_i8(v.x);
                    diag[ch].frame[2 * (o + i) + 1] = diag_i8(v.y);
                }
            }
        }
    if (diag) {
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
            s_abs += __shfl_xor_sync(0xffffffffu, s_abs, w);
            s_i += __shfl_xor_sync(0xffffffffu, s_i, w);
            s_q += __shfl_xor_sync(0xffffffffu, s_q, w);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, w);
        }
        if ((threadIdx.x & 31) == 0 && cnt) {
            atomicAdd(&diag[ch].sum_abs_i, s_abs);
            atomicAdd(&diag[ch].sum_sq_i, s_i);
            atomicAdd(&diag[ch].sum_sq_q, s_q);
            atomicAdd(&diag[ch].n, (unsigned long long)cnt);
        }
    }
}

__global__ void mm_offsets_kernel(int nseg, const MmSegOut *__restrict__ segout, long long *__restrict__ offsets,
                                  int *__restrict__ overflow, MmDiag *__restrict__ diag)
{
    const int ch = blockIdx.x;
    segout += (size_t)ch * nseg;
    offsets += (size_t)ch * (nseg + 1);
    if (threadIdx.x == 0) {
        long long o = 0;
        int ov = 0;
        for (int j = 0; j < nseg; j++) {
            offsets[j] = o;
            o += segout[j].n_sym;
            ov |= segout[j].overflow;
        }
        offsets[nseg] = o;
        if (diag) diag[ch].n_frame = (int)(o < 1024 ? o : 1024);   // demodulator.cpp:162: symbols < 1024 ? symbols : 1024 floats
        if (ov) atomicExch(overflow, 1);
    }
}

// int8 soft symbols (SymbolManager::process, reference SymbolManager.cpp:43-46):
// f = Re(s)*127, clamp to [-128, 127], C cast (truncation toward zero)
__global__ void soft_i8_kernel(const float2 *__restrict__ sym, signed char *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = soft_i8(sym[i].x);
}

}  // namespace xrd
