// xrd_api.cu -- host side of the B200-native xritdemod hot path: stage runners, the chain
// (processSamples(), reference demodulator/src/demodulator.cpp:100-168) and the C ABI of
// include/xrd.h.  Compiled with -fmad=false (see xrd_kernels.cuh).
#include "../../include/xrd.h"
#include "xrd_kernels.cuh"
#include "xrd_wn.cuh"
#include "xrd_fir_tma.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <string>
#include <vector>

namespace xrd {

// Last error message of the calling thread (xrd_last_error / xrd_stage_last_error).  Per thread, like errno: the
// reference wiring calls xrd_add_samples on the frontend thread and xrd_process on the symbol-loop thread
// (demodulator.cpp:434,475), and neither may clobber a message the other is reading.
static thread_local std::string g_error;

struct CudaError {
    std::string msg;
    int code;
};

#define XRD_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            char b__[512];                                                                               \
            snprintf(b__, sizeof b__, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            throw CudaError{b__, (e__ == cudaErrorMemoryAllocation) ? XRD_E_NOMEM : XRD_E_CUDA};          \
        }                                                                                                \
    } while (0)

struct Counters {
    uint64_t launches = 0;
};

#define XRD_LAUNCH(ctr, kernel, grid, block, smem, stream, ...)  \
    do {                                                         \
        kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);      \
        (ctr).launches++;                                        \
        XRD_CUDA(cudaGetLastError());                            \
    } while (0)

// grow-only device buffer (the reference's checkAndResizeBuffers, demodulator.cpp:76-92)
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    void ensure(size_t need)
    {
        if (need <= bytes) return;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        // whole 1 KB rows: the TMA tensor map of a sample buffer (xrd_fir_tma.cuh) covers the allocation exactly
        size_t want = (need + need / 8 + 256 + 1023) & ~(size_t)1023;
        XRD_CUDA(cudaMalloc(&p, want));
        bytes = want;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    void init()
    {
        if (!a) {
            XRD_CUDA(cudaEventCreate(&a));
            XRD_CUDA(cudaEventCreate(&b));
        }
    }
    void start(cudaStream_t s) { init(); XRD_CUDA(cudaEventRecord(a, s)); }
    void stop(cudaStream_t s) { XRD_CUDA(cudaEventRecord(b, s)); }
    float ms()
    {
        float t = 0.f;
        if (a && cudaEventSynchronize(b) == cudaSuccess) cudaEventElapsedTime(&t, a, b);
        return t;
    }
    ~Timer()
    {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

// ---------------------------------------------------------------------------------------
// tap designers (host, FP64 then cast, as the blocks the reference names do)
// ---------------------------------------------------------------------------------------
static const double kPi = 3.14159265358979323846;

// Filters::RRC == firdes.root_raised_cosine (demodulator.cpp:443, demod_tcp_qt.py:95-96)
static int design_rrc(double gain, double fs, double rs, double alpha, int ntaps, std::vector<float> &out)
{
    ntaps |= 1;
    out.assign(ntaps, 0.f);
    const double spb = fs / rs;
    const int mid = ntaps / 2;
    double sum = 0.0;
    for (int i = 0; i < ntaps; i++) {
        const double xi = (double)(i - mid);
        const double x1 = kPi * xi / spb;
        const double x2 = 4.0 * alpha * xi / spb;
        double x3 = x2 * x2 - 1.0;
        double num, den;
        if (std::fabs(x3) >= 0.000001) {
            num = (i != mid) ? std::cos((1 + alpha) * x1) + std::sin((1 - alpha) * x1) / (4 * alpha * xi / spb)
                             : std::cos((1 + alpha) * x1) + (1 - alpha) * kPi / (4 * alpha);
            den = x3 * kPi;
        } else {
            if (alpha == 1) {
                out[i] = -1.f;
                sum += out[i];
                continue;
            }
            const double a3 = (1 - alpha) * x1, a2 = (1 + alpha) * x1;
            num = std::sin(a2) * (1 + alpha) * kPi - std::cos(a3) * ((1 - alpha) * kPi * spb) / (4 * alpha * xi) +
                  std::sin(a3) * spb * spb / (4 * alpha * xi * xi);
            den = -32 * kPi * alpha * alpha * xi / spb;
        }
        out[i] = (float)(4 * alpha * num / den);
        sum += out[i];
    }
    for (auto &t : out) t = (float)(t * gain / sum);
    return ntaps;
}

static int lowpass_ntaps(double fs, double tw)
{
    int n = (int)(53.0 * fs / (22.0 * tw));   // Hamming: 53 dB
    return (n & 1) ? n : n + 1;
}

// Filters::lowPass(..., HAMMING) == firdes.low_pass (demodulator.cpp:444, demod_tcp_qt.py:261-262)
static int design_lowpass(double gain, double fs, double fc, double tw, std::vector<float> &out)
{
    const int ntaps = lowpass_ntaps(fs, tw);
    const int M = (ntaps - 1) / 2;
    const double wc = 2 * kPi * fc / fs;
    out.assign(ntaps, 0.f);
    for (int n = -M; n <= M; n++) {
        const float w = (float)(0.54 - 0.46 * std::cos((2 * kPi * (n + M)) / (ntaps - 1)));
        out[n + M] = (n == 0) ? (float)(wc / kPi * w) : (float)(std::sin(n * wc) / (n * kPi) * w);
    }
    double fmax = out[M];
    for (int n = 1; n <= M; n++) fmax += 2 * out[n + M];
    const double g = gain / fmax;
    for (auto &t : out) t = (float)(t * g);
    return ntaps;
}

// mmse_fir_interpolator taps: least-squares 8-tap fractional delay over |f| <= 1/4, 128 steps,
// rounded to the 6 significant digits the upstream table is printed with.  Solved here by
// Cholesky factorisation of the (symmetric positive definite) sinc Gram matrix.
static void mmse_table(float *tab)
{
    auto sinc = [](double x) { return std::fabs(x) < 1e-12 ? 1.0 : std::sin(kPi * x) / (kPi * x); };
    const int N = MM_NTAPS;
    const double B = 0.25;
    double Lm[8][8] = {{0}};
    for (int i = 0; i < N; i++)
        for (int j = 0; j <= i; j++) {
            double s = sinc(2 * B * (double)(i - j));
            for (int k = 0; k < j; k++) s -= Lm[i][k] * Lm[j][k];
            Lm[i][j] = (i == j) ? std::sqrt(s) : s / Lm[j][j];
        }
    for (int k = 0; k <= MM_NSTEPS; k++) {
        const double mu = (double)k / MM_NSTEPS;
        double y[8], h[8];
        for (int i = 0; i < N; i++) {
            double s = sinc(2 * B * ((double)(i - 4) + mu));
            for (int j = 0; j < i; j++) s -= Lm[i][j] * y[j];
            y[i] = s / Lm[i][i];
        }
        for (int i = N - 1; i >= 0; i--) {
            double s = y[i];
            for (int j = i + 1; j < N; j++) s -= Lm[j][i] * h[j];
            h[i] = s / Lm[i][i];
        }
        for (int j = 0; j < N; j++) {
            char buf[64];
            snprintf(buf, sizeof buf, "%.5e", h[j]);
            tab[k * N + j] = (float)strtod(buf, nullptr);
        }
    }
    for (int j = 0; j < N; j++) {
        tab[j] = (j == 4) ? 1.f : 0.f;
        tab[MM_NSTEPS * N + j] = (j == 3) ? 1.f : 0.f;
    }
}

static void costas_gains(float bw, float &alpha, float &beta)
{
    // control_loop::update_gains with damping sqrt(2)/2, all FP32
    const float damping = sqrtf(2.0f) / 2.0f;
    const float denom = (1.0f + 2.0f * damping * bw + bw * bw);
    alpha = (4 * damping * bw) / denom;
    beta = (4 * bw * bw) / denom;
}

// ---------------------------------------------------------------------------------------
// FIR stage
// ---------------------------------------------------------------------------------------
struct FirStage {
    int D = 1, ntaps = 0;
    DevBuf d_taps;
    void init(unsigned decim, const float *taps, int n)
    {
        D = decim ? (int)decim : 1;
        ntaps = n;
        d_taps.ensure(sizeof(float) * n);
        XRD_CUDA(cudaMemcpy(d_taps.p, taps, sizeof(float) * n, cudaMemcpyHostToDevice));
        memset(&tma_taps, 0, sizeof tma_taps);
        for (int k = 0; k < n && k < FT_MAX_TAPS; k++) tma_taps.h2[k] = make_float2(taps[k], taps[k]);
    }
    int hist() const { return ntaps - 1; }
    // TMA path of the stride-1 filter: tensor map of the allocation the input lives in, re-encoded when it moves
    bool use_tma = true;
    CUtensorMap tmap;
    FtTaps tma_taps;
    const void *tmap_base = nullptr;
    size_t tmap_bytes = 0;
    int tma_ctas_per_sm = 0, sm_count = 0;
    bool run_tma(Counters &c, cudaStream_t st, const float2 *in, float2 *out, long long n_out, int nch, long long in_stride,
                 long long out_stride, const DevBuf *src)
    {
        if (!use_tma || !src || !src->p || ft_rows(ntaps) > 256 || ntaps > FT_MAX_TAPS) return false;
        const size_t smem = ft_smem_bytes(ntaps);
        if (smem > 200 * 1024) return false;
        const float2 *base = src->as<float2>();
        if (in - (ntaps - 1) < base || n_out > (1LL << 40)) return false;
        if (tmap_base != src->p || tmap_bytes != src->bytes) {
            if (!ft_make_map(&tmap, src->p, src->bytes, ntaps)) return false;
            tmap_base = src->p;
            tmap_bytes = src->bytes;
        }
        if (!tma_ctas_per_sm) {
            XRD_CUDA(cudaFuncSetAttribute(fir_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int dev = 0;
            XRD_CUDA(cudaGetDevice(&dev));
            XRD_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
            XRD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tma_ctas_per_sm, fir_tma_kernel, FT_THREADS, smem));
            if (tma_ctas_per_sm < 1) tma_ctas_per_sm = 1;
        }
        const long long tiles_per_ch = (n_out + FT_TILE - 1) / FT_TILE;
        const long long n_tiles = tiles_per_ch * nch;
        if (n_tiles > 0x7fffffffLL) return false;
        const int grid = (int)std::min<long long>(n_tiles, (long long)sm_count * tma_ctas_per_sm);
        XRD_LAUNCH(c, fir_tma_kernel, grid, FT_THREADS, smem, st, tmap, tma_taps, out, ntaps, n_out,
                   (long long)(in - base), in_stride, out_stride, (int)tiles_per_ch, (int)n_tiles);
        return true;
    }
    // in: x[0] of this call (history before it); n_out outputs per channel.  src: the allocation `in` points into
    // (enables the TMA-staged kernel for decimation 1)
    void run(Counters &c, cudaStream_t st, const float2 *in, float2 *out, long long n_out, int nch, long long in_stride,
             long long out_stride, const DevBuf *src = nullptr)
    {
        if (n_out <= 0) return;
        const int tp = (ntaps + 1) & ~1;
        if (D == 1) {
            if (run_tma(c, st, in, out, n_out, nch, in_stride, out_stride, src)) return;
            const size_t smem = sizeof(float) * tp + sizeof(float2) * (1 + FIR_TILE + ntaps - 1);
            if (smem > 48 * 1024)
                XRD_CUDA(cudaFuncSetAttribute(fir1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid((unsigned)((n_out + FIR_TILE - 1) / FIR_TILE), nch);
            XRD_LAUNCH(c, fir1_kernel, grid, FIR_THREADS, smem, st, in, out, d_taps.as<float>(), ntaps, n_out, in_stride,
                       out_stride);
        } else {
            // decimating filters read their history from a separate buffer: run_decim
            throw CudaError{"FirStage::run: decimating filters go through run_decim", XRD_E_ARG};
        }
    }
    bool force_generic = false;

    // Decimating filter (D > 1): `in` holds n_out * D samples of ingest format `type` per channel (x[0] first, channel
    // stride in_stride samples), `hist` the ntaps - 1 cf32 samples before x[0] of every channel ([nch][ntaps - 1]); the
    // history is advanced to the end of this call afterwards.  S16 input converts as it loads (polyphase kernels);
    // S8 / U8 go through the generic kernel.
    template <class IN>
    void launch_decim(Counters &c, cudaStream_t st, const void *in_any, float2 *hist, float2 *out, long long n_out, int nch,
                      long long in_stride, long long out_stride, bool poly_ok)
    {
        const typename IN::raw *in = static_cast<const typename IN::raw *>(in_any);
        bool done = false;
        if (poly_ok && D >= 2 && D <= 5 && !force_generic && ntaps <= FT_MAX_TAPS) {
#define XRD_FIR_POLY(DV)                                                                                                    \
    do {                                                                                                                    \
        const size_t smem = FirPoly<DV>::smem_bytes(ntaps);                                                                 \
        if (smem > 200 * 1024) break;                                                                                       \
        XRD_CUDA(cudaFuncSetAttribute((fird_poly_kernel<DV, IN>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dim3 grid((unsigned)((n_out + FirPoly<DV>::TILE - 1) / FirPoly<DV>::TILE), nch);                                    \
        XRD_LAUNCH(c, (fird_poly_kernel<DV, IN>), grid, FP_THREADS, smem, st, in, hist, out, tma_taps, ntaps, n_out,        \
                   in_stride, out_stride);                                                                                  \
        done = true;                                                                                                        \
    } while (0)
            if (D == 2) XRD_FIR_POLY(2);
            else if (D == 3) XRD_FIR_POLY(3);
            else if (D == 4) XRD_FIR_POLY(4);
            else XRD_FIR_POLY(5);
#undef XRD_FIR_POLY
        }
        if (!done) {
            const int tp = (ntaps + 1) & ~1;
            const size_t smem = sizeof(float) * tp + sizeof(float2) * ((size_t)(FIRD_TILE - 1) * D + 1 + ntaps - 1);
            if (smem > 48 * 1024)
                XRD_CUDA(cudaFuncSetAttribute(fird_kernel<IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid((unsigned)((n_out + FIRD_TILE - 1) / FIRD_TILE), nch);
            XRD_LAUNCH(c, fird_kernel<IN>, grid, FIR_THREADS, smem, st, in, hist, out, d_taps.as<float>(), ntaps, D, n_out,
                       in_stride, out_stride);
        }
        if (ntaps > 1)
            XRD_LAUNCH(c, fir_hist_carry_kernel<IN>, nch, 256, sizeof(float2) * (ntaps - 1), st, in, hist, ntaps - 1, n_out * D,
                       in_stride);
    }
    void run_decim(Counters &c, cudaStream_t st, const void *in, int type, float2 *hist, float2 *out, long long n_out, int nch,
                   long long in_stride, long long out_stride)
    {
        if (n_out <= 0) return;
        if (sizeof(float2) * (size_t)(ntaps - 1) > 200 * 1024) throw CudaError{"FirFilter: too many taps", XRD_E_ARG};
        switch (type) {
        case XRD_FLOATIQ: launch_decim<InF32>(c, st, in, hist, out, n_out, nch, in_stride, out_stride, true); break;
        case XRD_S16IQ: launch_decim<InS16>(c, st, in, hist, out, n_out, nch, in_stride, out_stride, true); break;
        case XRD_S8IQ: launch_decim<InS8>(c, st, in, hist, out, n_out, nch, in_stride, out_stride, false); break;
        case XRD_U8IQ: launch_decim<InU8>(c, st, in, hist, out, n_out, nch, in_stride, out_stride, false); break;
        default: throw CudaError{"FirFilter: unknown sample type", XRD_E_ARG};
        }
    }
};

// keep the last `keep` samples of [prefix | n] as the next call's prefix (FIR history, M&M tail)
__global__ void carry_prefix_kernel(float2 *buf /* start of the prefix */, int keep, long long n, long long ch_stride)
{
    extern __shared__ float2 s_keep[];
    buf += (size_t)blockIdx.x * ch_stride;
    for (int i = threadIdx.x; i < keep; i += blockDim.x) s_keep[i] = buf[n + i];
    __syncthreads();
    for (int i = threadIdx.x; i < keep; i += blockDim.x) buf[i] = s_keep[i];
}

// Round counters go back to the host through mapped pinned memory, written by a one-thread kernel, not through the
// copy engine: a 4-byte cudaMemcpyAsync queues behind whatever the device-to-host engine is doing, and with several
// calls in flight that is another handle's symbol buffer (0.37 GB: 7 ms on a 55 GB/s link, 20 ms on the 19 GB/s links
// of an 8-GPU box) -- every certified round of every stage would wait that long with its SMs idle.
__global__ void publish_kernel(volatile int *host_dst, const int *a, const int *b)
{
    host_dst[0] = *a;
    if (b) host_dst[1] = *b;
    __threadfence_system();
}

template <class S> __global__ void take_last_kernel(S *carried, const S *exit_, int nseg, int nch)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < nch) carried[ch] = exit_[(size_t)ch * nseg + (nseg - 1)];
}

// ---------------------------------------------------------------------------------------
// AGC / Costas: segment-parallel loop stage
// ---------------------------------------------------------------------------------------
constexpr int SEG_NTH = 64;   // segments (threads) per CTA of the segment loop kernel
constexpr int SEG_TS = 16;    // samples per register tile (one 128-byte line)

constexpr int WN_K = 4;       // samples per lane of a window-Newton chain (window = 128 samples per warp)
constexpr int WN_CKPT = 2048; // samples between state checkpoints (power of two >= 32 * WN_K)

template <class LOOP> struct SegStage {
    typedef typename LOOP::State State;
    typename LOOP::Params prm;
    // thread-per-segment kernel (seg_loop_kernel): short segments, one thread each
    int L = 4096, W = 32768;
    // window-Newton kernel (wn_loop_kernel): one warp per segment, long segments.  Lw == 0: as many
    // segments as the device holds chains (chains_per_sm warps per SM), at least Lw_min samples each
    int Lw = 0, Ww = 16384, Lw_min = 16384, chains_per_sm = 8, chains_per_sm_max = 16;
    bool use_wn = true;
    int wn_variant = 2;   // 2: one warp per chain (K = 4); 3..7: one CTA per chain, (K, warps) = (1,4) (2,4) (1,2) (2,2) (2,8)
    int redo_variant = 4; // kernel of the certified re-runs: few chains, so the widest window (fastest single chain) wins
    bool use_mirror = false;
    bool guided = true;       // the first pass records its trajectory and the CTA-chain re-runs take their proposals from it
    bool chase = true;        // CTA-chain re-runs that do not merge inside their segment keep going into the next one
    int nch = 1, sm_count = 148;
    DevBuf d_carried, d_entry, d_exit, d_redo, d_mirror, d_nredo, d_list, d_ckpt, d_iters, d_psi, d_adv, d_pre, d_adv0;
    int pre_len = 0;          // Costas: samples of segment 0 run ahead of the branch resolution (0: none)
    int *h_nredo = nullptr;   // pinned, mapped
    int *h_nredo_dev = nullptr; // its device-side address (publish_kernel)
    uint64_t rounds = 0, redone = 0, escalations = 0;
    State s_init;

    void reset()
    {
        std::vector<State> v(nch, s_init);
        XRD_CUDA(cudaMemcpy(d_carried.p, v.data(), sizeof(State) * nch, cudaMemcpyHostToDevice));
    }
    void init(int nch_, const State &s0)
    {
        nch = nch_;
        s_init = s0;
        d_carried.ensure(sizeof(State) * nch);
        reset();
        d_nredo.ensure(sizeof(int));
        d_iters.ensure(sizeof(unsigned long long));
        XRD_CUDA(cudaMemset(d_iters.p, 0, sizeof(unsigned long long)));
        if (!h_nredo) {
            XRD_CUDA(cudaHostAlloc(&h_nredo, 2 * sizeof(int), cudaHostAllocMapped));
            XRD_CUDA(cudaHostGetDevicePointer(&h_nredo_dev, h_nredo, 0));
        }
        int dev = 0;
        XRD_CUDA(cudaGetDevice(&dev));
        XRD_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    ~SegStage()
    {
        if (h_nredo) cudaFreeHost(h_nredo);
    }
    State get(int ch)
    {
        State s;
        XRD_CUDA(cudaMemcpy(&s, d_carried.as<State>() + ch, sizeof(State), cudaMemcpyDeviceToHost));
        return s;
    }
    uint64_t wn_iters()
    {
        unsigned long long v = 0;
        if (d_iters.p) cudaMemcpy(&v, d_iters.p, sizeof v, cudaMemcpyDeviceToHost);
        return v;
    }

    void resolve(Counters &c, cudaStream_t st, int nseg, int Ls, long long Ws) { resolve_impl(c, st, nseg, Ls, Ws, d_entry.as<State>()); }
    void resolve_impl(Counters &c, cudaStream_t st, int nseg, int Ls, long long Ws, CostasState *e)
    {
        XRD_LAUNCH(c, costas_resolve_kernel, nch, 256, sizeof(unsigned) * ((nseg + 31) / 32), st, nseg, Ls, Ws, e,
                   d_adv.as<float>(), (int *)nullptr, pre_len > 0 ? d_pre.as<CostasState>() : (const CostasState *)nullptr,
                   pre_len > 0 ? d_adv0.as<float>() : (const float *)nullptr);
    }
    void resolve_impl(Counters &, cudaStream_t, int, int, long long, AgcState *) {}

    long long hist = 0;   // samples of the same stream addressable before `in` (set by run)
    bool in_s16 = false;  // this call's `in` holds S16 IQ samples (AGC as the first kernel of the chain; set by run)
    State *traj = nullptr; // this call's trajectory record (set by run; null: none, re-runs are not guided)
    int first_variant = 2; // kernel of this call's first pass (set by run): wn_variant, or the CTA chains of the re-runs when
                           // the call has too few segments to fill the device with warp chains (FIFO-sized calls)
    // the fused S16 ingest exists for the default kernel shapes of the AGC only
    bool can_fuse_s16() const { return std::is_same<LOOP, AgcLoop>::value && use_wn && wn_variant == 2 && redo_variant == 4; }
    void launch(Counters &c, cudaStream_t st, bool wn, const void *in_any, float2 *out, long long n, int Ls, int Ws, int nseg,
                int n_work, int ncp, int mode, long long in_stride, long long out_stride)
    {
        const float2 *in = static_cast<const float2 *>(in_any);
        const int variant = (mode == 1) ? redo_variant : first_variant;
        const unsigned char *redo_flags = (mode == 1 && chase) ? d_redo.as<unsigned char>() : (const unsigned char *)nullptr;
        if constexpr (std::is_same<LOOP, AgcLoop>::value) {
            if (in_s16) {
                // converts as it loads: wn_loop_kernel / wn_cta_kernel on short2 samples (can_fuse_s16() was checked)
                const short2 *in16 = static_cast<const short2 *>(in_any);
                if (variant == 4) {
                    XRD_LAUNCH(c, (wn_cta_kernel<LOOP, 2, 4, InS16>), n_work, 32 * 4, 0, st, in16, out, n, Ls, Ws, nseg, n_work,
                               d_entry.template as<State>(), d_exit.template as<State>(), d_carried.template as<State>(),
                               d_list.template as<int>(), d_ckpt.template as<State>(), ncp, WN_CKPT,
                               d_iters.template as<unsigned long long>(), prm, mode, in_stride, out_stride, hist, redo_flags,
                               traj);
                } else {
                    const int grid = (n_work + WN_WARPS - 1) / WN_WARPS;
                    XRD_LAUNCH(c, (wn_loop_kernel<LOOP, WN_K, InS16>), grid, WN_WARPS * 32, wn_smem_bytes<WN_K>(), st, in16, out, n,
                               Ls, Ws, nseg, n_work, d_entry.template as<State>(), d_exit.template as<State>(),
                               d_carried.template as<State>(), d_list.template as<int>(), d_ckpt.template as<State>(), ncp,
                               WN_CKPT, d_iters.template as<unsigned long long>(), prm, mode, in_stride, out_stride, hist,
                               (State *)nullptr, 0, traj);
                }
                return;
            }
        }
        if (wn && variant >= 3) {
            // one CTA of WPC warps per chain (wn_cta_kernel<LOOP, K, WPC>)
#define XRD_WN_CTA(KV, WV)                                                                                              \
    XRD_LAUNCH(c, (wn_cta_kernel<LOOP, KV, WV>), n_work, 32 * WV, 0, st, in, out, n, Ls, Ws, nseg, n_work,              \
               d_entry.as<State>(), d_exit.as<State>(), d_carried.as<State>(), d_list.as<int>(), d_ckpt.as<State>(), ncp, \
               WN_CKPT, d_iters.as<unsigned long long>(), prm, mode, in_stride, out_stride, hist, redo_flags, traj)
            if (variant == 3) XRD_WN_CTA(1, 4);
            else if (variant == 4) XRD_WN_CTA(2, 4);
            else if (variant == 5) XRD_WN_CTA(1, 2);
            else if (variant == 6) XRD_WN_CTA(2, 2);
            else XRD_WN_CTA(2, 8);
#undef XRD_WN_CTA
        } else if (wn) {
            const int grid = (n_work + WN_WARPS - 1) / WN_WARPS;
            XRD_LAUNCH(c, (wn_loop_kernel<LOOP, WN_K>), grid, WN_WARPS * 32, wn_smem_bytes<WN_K>(), st, in, out, n, Ls, Ws, nseg,
                       n_work, d_entry.as<State>(), d_exit.as<State>(), d_carried.as<State>(), d_list.as<int>(),
                       d_ckpt.as<State>(), ncp, WN_CKPT, d_iters.as<unsigned long long>(), prm, mode, in_stride, out_stride,
                       hist, (mode == 2 && pre_len > 0) ? d_pre.as<State>() : (State *)nullptr, pre_len, traj);
        } else {
            const int grid = (n_work + SEG_NTH - 1) / SEG_NTH;
            XRD_LAUNCH(c, (seg_loop_kernel<LOOP, SEG_TS>), grid, SEG_NTH, 0, st, in, out, n, Ls, Ws, nseg, n_work,
                       d_entry.as<State>(), d_exit.as<State>(), d_carried.as<State>(), d_list.as<int>(), prm, mode, in_stride,
                       out_stride);
        }
    }

    // hist_avail: samples of the same stream that are still in place right before `in` (an earlier piece of the
    // same call); speculative warm-ups of the window kernel may start there instead of at in[0]
    // s16: `in` holds S16 IQ samples instead of cf32 (caller checked can_fuse_s16())
    // traj_buf: trajectory record of the call, indexed like `out` (same channel stride), or null
    void run(Counters &c, cudaStream_t st, const void *in_any, float2 *out, long long n, long long in_stride,
             long long out_stride, long long hist_avail = 0, bool s16 = false, State *traj_buf = nullptr)
    {
        if (n <= 0) return;
        const float2 *in = static_cast<const float2 *>(in_any);
        in_s16 = s16;
        const bool wn = use_wn;
        traj = (wn && guided) ? traj_buf : nullptr;
        hist = wn ? hist_avail : 0;
        int Ls, Ws;
        if (wn) {
            long long l = Lw;
            if (l <= 0) {
                // more chains per SM when the call is large enough to keep them long (many channels)
                int cps = chains_per_sm;
                if (chains_per_sm_max > cps)
                    cps = (int)std::min<long long>(chains_per_sm_max, std::max<long long>(cps, n * nch / 100000 / sm_count));
                const int per_ch = std::max(1, sm_count * cps / nch);
                l = std::max<long long>(Lw_min, (n + per_ch - 1) / per_ch);
            }
            l = std::min<long long>(std::max<long long>(l, 32), 1 << 30);
            Ls = (int)((l + WN_CKPT - 1) / WN_CKPT * WN_CKPT);   // checkpoints and carrier-phase blocks tile the segment
            Ws = (std::max(Ww, 0) + 31) / 32 * 32;
        } else {
            Ls = std::max(L, SEG_TS);
            Ws = ((std::max(W, 0) + SEG_TS - 1) / SEG_TS) * SEG_TS;
        }
        for (int attempt = 0;; attempt++) {
            const int nseg = (int)((n + Ls - 1) / Ls);
            const size_t tot = (size_t)nseg * nch;
            const int ncp = Ls / WN_CKPT + 2;
            first_variant = wn_variant;
            if (wn && wn_variant == 2 && redo_variant >= 3 && tot <= (size_t)2 * sm_count) first_variant = redo_variant;
            d_entry.ensure(sizeof(State) * tot);
            d_exit.ensure(sizeof(State) * tot);
            d_redo.ensure(tot);
            d_list.ensure(sizeof(int) * tot);
            if (use_mirror) d_mirror.ensure(tot);
            if (wn) d_ckpt.ensure(sizeof(State) * tot * ncp);
            if (wn && use_mirror && nseg > 1 && n >= 2 * CPB) {
                // Costas: warm-ups, then put the entry states on one carrier-phase branch, then the segments.
                // Segment 0 has no warm-up; the window kernel runs its first pre_len samples beside the others'
                // warm-ups so that the resolution can refer to the branch the true trajectory acquires on.
                pre_len = (first_variant == 2 && hist == 0) ? std::min(Ws, Ls) / CPB * CPB : 0;
                if (pre_len > 0) {
                    d_pre.ensure(sizeof(State) * nch);
                    d_adv0.ensure(sizeof(float) * nch);
                }
                launch(c, st, wn, in_any, out, n, Ls, Ws, nseg, (int)tot, ncp, 2, in_stride, out_stride);
                const long long nblk = n / CPB;
                d_psi.ensure(sizeof(float) * (size_t)nblk * nch);
                d_adv.ensure(sizeof(float) * tot);
                dim3 g1((unsigned)((nblk + 255) / 256), nch);   // 8 warps x 32 blocks per CTA
                XRD_LAUNCH(c, costas_block_phase_kernel, g1, 256, 0, st, in, d_psi.as<float>(), nblk, in_stride, nblk);
                dim3 g2(nseg, nch);
                XRD_LAUNCH(c, costas_seg_advance_kernel, g2, 256, 0, st, d_psi.as<float>(), d_adv.as<float>(), nseg, Ls / CPB,
                           nblk, nblk, pre_len > 0 ? d_adv0.as<float>() : (float *)nullptr, pre_len / CPB);
                resolve(c, st, nseg, Ls, (long long)Ws - hist);
                launch(c, st, wn, in_any, out, n, Ls, Ws, nseg, (int)tot, ncp, 3, in_stride, out_stride);
            } else {
                launch(c, st, wn, in_any, out, n, Ls, Ws, nseg, (int)tot, ncp, 0, in_stride, out_stride);
            }
            bool escalate = false;
            for (int round = 0; nseg > 1 && round < nseg; round++) {
                XRD_CUDA(cudaMemsetAsync(d_nredo.p, 0, sizeof(int), st));
                XRD_LAUNCH(c, (seg_verify_kernel<LOOP>), nch, 256, 0, st, nseg, d_entry.as<State>(), d_exit.as<State>(),
                           d_redo.as<unsigned char>(), use_mirror ? d_mirror.as<unsigned char>() : nullptr,
                           d_nredo.as<int>(), d_list.as<int>(), round == 0 ? 1 : 0);
                XRD_LAUNCH(c, publish_kernel, 1, 1, 0, st, h_nredo_dev, d_nredo.as<int>(), (const int *)nullptr);
                XRD_CUDA(cudaStreamSynchronize(st));
                const int nr = *(volatile int *)h_nredo;
                if (nr == 0) break;
                // speculation that mostly fails (weak signal: slow AGC; no lock) would need a fix-up round per
                // segment: redo the pass with longer warm-ups and segments instead (bounded).  Re-runs of the
                // window kernel stop as soon as they merge with the trajectory in place, so it never needs this.
                if (!wn && round == 0 && (size_t)nr * 100 > tot * 95 && tot >= 64 && attempt < 3) {
                    escalate = true;
                    break;
                }
                rounds++;
                redone += (uint64_t)nr;
                launch(c, st, wn, in_any, out, n, Ls, Ws, nseg, nr, ncp, 1, in_stride, out_stride);
            }
            if (!escalate) {
                XRD_LAUNCH(c, (take_last_kernel<State>), (nch + 127) / 128, 128, 0, st, d_carried.as<State>(),
                           d_exit.as<State>(), nseg, nch);
                return;
            }
            escalations++;
            Ls *= 4;
            Ws *= 4;
        }
    }
};

// the window kernel holds the AGC gain in 2^-40 fixed point: needs a finite clamp well inside 2^23
static bool agc_wn_ok(float max_gain) { return max_gain > 0.f && max_gain < 4.0e6f; }
// the window kernel's Costas step removes at most one turn per sample (CostasLoopK::step_sel)
static bool costas_wn_ok(const CostasParams &p)
{
    return std::fabs(p.alpha) + std::max(std::fabs(p.max_freq), std::fabs(p.min_freq)) < 6.0f;
}

// ---------------------------------------------------------------------------------------
// M&M stage
// ---------------------------------------------------------------------------------------
__global__ void mm_rebase_kernel(MmState *carried, const MmState *exit_, int nseg, int nch, long long n)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= nch) return;
    MmState s = exit_[(size_t)ch * nseg + (nseg - 1)];
    s.ii -= n;   // relative to the next chunk's first sample
    carried[ch] = s;
}

static int next_pow2(long long v)
{
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

struct MmStage {
    MmParams prm;
    long long L = 0;                // 0: one segment per SM (set per call)
    long long W_user = 0;           // speculative warm-up in samples; 0: 80 k when re-runs can walk relative to the recorded
                                    // trajectory (which only needs the warm-up to land within a few grid units of the
                                    // truth), 800 k when they are full chain re-runs (which need a bitwise merge)
    long long W = 800000;           // the warm-up of the current call
    long long Lmin = 262144;
    long long single_max = 600000;  // FIFO-sized calls (<= 512 Ki samples, Parameters.h:57) run as ONE exact chain: no
                                    // speculation, no walks
    int nt = 0;                     // lanes per chain of mm_chain32_kernel (0 = auto)
    bool force64 = false;           // tests: always use the generic 64-bit chain kernel
    int sm_count = 148;
    int nch = 1;
    DevBuf d_table, d_carried, d_entry, d_exit, d_redo, d_nredo, d_segout, d_offsets, d_stage, d_overflow, d_ckpt;
    DevBuf d_traj;                  // per-symbol trajectory record (MmTraj), same indexing as d_stage
    bool use_delta = true;          // certified re-runs walk relative to the trajectory in place (mm_delta_kernel)
    int delta_nt = 512;             // lanes of mm_delta_kernel (halved until its rings fit in shared memory)
    bool traj_on = false;           // this call records the trajectory (more than one segment, 32-bit chain kernel)
    uint64_t bails = 0;             // delta re-runs that gave up and fell back to the chain kernel
    int ck_spacing = 65536;         // samples between chain checkpoints
    int *h_nredo = nullptr, *h_nredo_dev = nullptr;   // pinned + mapped; [1] overflow, [2] walks that gave up, [3] stall
                                                        // flag, [4..5] written by publish_kernel every round
    signed char *out_i8 = nullptr;  // when set, the compaction also (out != null) or only (out == null) emits int8 soft symbols
    std::vector<long long> h_offsets;
    DevBuf d_diag;                  // per-channel diagnostics of the last call (MmDiag), filled by the compaction pass
    std::vector<MmDiag> h_diag;
    uint64_t rounds = 0, redone = 0, windows = 0, iters = 0;

    MmState s_init;
    void reset()
    {
        std::vector<MmState> v(nch, s_init);
        XRD_CUDA(cudaMemcpy(d_carried.p, v.data(), sizeof(MmState) * nch, cudaMemcpyHostToDevice));
    }
    void init(int nch_, float omega, float gain_omega, float mu, float gain_mu, float omega_rel_limit)
    {
        nch = nch_;
        prm.omega_mid = omega;
        prm.omega_lim = omega_rel_limit * omega;
        prm.gain_omega = gain_omega;
        prm.gain_mu = gain_mu;
        std::vector<float> tab(129 * 8);
        mmse_table(tab.data());
        d_table.ensure(sizeof(float) * tab.size());
        XRD_CUDA(cudaMemcpy(d_table.p, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice));
        memset(&s_init, 0, sizeof s_init);
        s_init.mu = mu;
        s_init.omega = omega;
        d_carried.ensure(sizeof(MmState) * nch);
        reset();
        d_nredo.ensure(4 * sizeof(int));   // [0] segments flagged by the verify pass, [1] delta re-runs that gave up,
                                           // [2] a chain hit its iteration cap (non-finite samples stalled the loop)
        d_overflow.ensure(sizeof(int));
        d_diag.ensure(sizeof(MmDiag) * nch);
        if (!h_nredo) {
            XRD_CUDA(cudaHostAlloc(&h_nredo, 8 * sizeof(int), cudaHostAllocMapped));
            XRD_CUDA(cudaHostGetDevicePointer(&h_nredo_dev, h_nredo, 0));
            memset(h_nredo, 0, 8 * sizeof(int));
        }
        int dev = 0;
        XRD_CUDA(cudaGetDevice(&dev));
        XRD_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    MmTraj traj()
    {
        MmTraj tr{nullptr};
        if (traj_on) tr.rec = d_traj.as<int4>();
        return tr;
    }
    // the 32-bit chain kernel needs small per-symbol deviations (see its header comment) and 31-bit indices
    bool fast32(int NT, long long n, long long Lseg, long long cap_seg) const
    {
        const double dev_max = (double)NT * prm.gain_omega + prm.gain_mu;
        return !force64 && n < (1LL << 30) && Lseg < (1LL << 30) && W < (1LL << 30) && cap_seg < (1LL << 30) &&
               32.0 * dev_max < 0.45 && 2.0 * prm.omega_lim + prm.gain_mu < 0.45 && prm.omega_mid < 1024.f;
    }
    // certified re-run of the flagged segments relative to the recorded trajectory; segments it gives up on stay flagged
    void launch_delta(Counters &c, cudaStream_t st, dim3 grid, const float2 *in, long long n, long long Lseg, int nseg,
                      long long cap_seg, long long in_stride, long long stage_stride)
    {
        const int ncp = (int)(Lseg / ck_spacing) + 2;
        const double adv = (double)prm.omega_mid + (double)prm.omega_lim + (double)prm.gain_mu + 0.01;
        if (delta_nt != 128 && delta_nt != 256 && delta_nt != 512) delta_nt = 512;
        int NTd = delta_nt, RX = 0;
        for (;; NTd >>= 1) {
            // samples of five windows (the one being read and the four slides its request may lag behind) plus the
            // block granularity of the refill
            RX = next_pow2((long long)(5 * NTd * adv) + NTd + 96);
            if (mm_delta_smem_bytes(NTd, RX) <= 190 * 1024 || NTd <= 128) break;
        }
        if (mm_delta_smem_bytes(NTd, RX) > 190 * 1024) return;   // symbols too long: the chain kernel takes all of it
#define XRD_MM_DELTA(NTV)                                                                                               \
    do {                                                                                                                \
        const size_t smem = mm_delta_smem_bytes(NTV, RX);                                                               \
        XRD_CUDA(cudaFuncSetAttribute(mm_delta_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        XRD_LAUNCH(c, mm_delta_kernel<NTV>, grid, NTV, smem, st, in, d_stage.as<float2>(), traj(), (int)n, (int)Lseg,    \
                   nseg, (int)cap_seg, d_entry.as<MmState>(), d_exit.as<MmState>(), d_redo.as<unsigned char>(),         \
                   d_segout.as<MmSegOut>(), d_table.as<float>(), prm, in_stride, stage_stride, d_ckpt.as<MmCk>(), ncp,  \
                   d_nredo.as<int>() + 1, RX);                                                                          \
    } while (0)
        if (NTd == 128) XRD_MM_DELTA(128);
        else if (NTd == 256) XRD_MM_DELTA(256);
        else XRD_MM_DELTA(512);
#undef XRD_MM_DELTA
    }
    // launch the chain kernel with the widest window whose sample ring fits in shared memory
    void launch_chain(Counters &c, cudaStream_t st, dim3 grid, const float2 *in, long long n, long long Lseg, int nseg,
                      long long cap_seg, int mode, long long in_stride, long long stage_stride)
    {
        const double adv = (double)prm.omega_mid + (double)prm.omega_lim + (double)prm.gain_mu + 0.01;
        int NT = nt ? nt : 1024;
        int R = 0;
        for (;; NT >>= 1) {
            // the lanes span NT * adv samples; the refill lags one iteration, so keep twice that resident
            R = next_pow2((long long)(2 * NT * adv) + 160);
            if (mm_chain32_smem_bytes(NT, R) <= 200 * 1024 || NT <= 128) break;
        }
        const bool fast = fast32(NT, n, Lseg, cap_seg);
        if (fast) {
            const size_t smem32 = mm_chain32_smem_bytes(NT, R);
            const int ncp = (int)(Lseg / ck_spacing) + 2;
            d_ckpt.ensure(sizeof(MmCk) * (size_t)grid.x * grid.y * ncp);
#define XRD_MM_CHAIN32(NTV)                                                                                              \
    do {                                                                                                                 \
        XRD_CUDA(cudaFuncSetAttribute(mm_chain32_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32)); \
        XRD_LAUNCH(c, mm_chain32_kernel<NTV>, grid, NTV, smem32, st, in, d_stage.as<float2>(), (int)n, (int)Lseg, (int)W, \
                   nseg, (int)cap_seg, d_entry.as<MmState>(), d_exit.as<MmState>(), d_carried.as<MmState>(),             \
                   d_redo.as<unsigned char>(), d_segout.as<MmSegOut>(), d_table.as<float>(), prm, mode, in_stride,       \
                   stage_stride, R, d_ckpt.as<MmCk>(), ncp, ck_spacing, traj(), d_nredo.as<int>() + 2);                  \
    } while (0)
            if (NT >= 1024) XRD_MM_CHAIN32(1024);
            else if (NT == 512) XRD_MM_CHAIN32(512);
            else if (NT == 256) XRD_MM_CHAIN32(256);
            else XRD_MM_CHAIN32(128);
#undef XRD_MM_CHAIN32
            return;
        }
        const size_t smem = mm_chain_smem_bytes(NT, R);
        if (smem > 227 * 1024) throw CudaError{"ClockRecovery: samples per symbol too large for the chain kernel", XRD_E_ARG};
#define XRD_MM_CHAIN(NTV)                                                                                              \
    do {                                                                                                               \
        XRD_CUDA(cudaFuncSetAttribute(mm_chain_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        XRD_LAUNCH(c, mm_chain_kernel<NTV>, grid, NTV, smem, st, in, d_stage.as<float2>(), n, Lseg, W, nseg, cap_seg,  \
                   d_entry.as<MmState>(), d_exit.as<MmState>(), d_carried.as<MmState>(), d_redo.as<unsigned char>(),   \
                   d_segout.as<MmSegOut>(), d_table.as<float>(), prm, mode, in_stride, stage_stride, R,                \
                   d_nredo.as<int>() + 2);                                                                             \
    } while (0)
        if (NT >= 1024) XRD_MM_CHAIN(1024);
        else if (NT == 512) XRD_MM_CHAIN(512);
        else if (NT == 256) XRD_MM_CHAIN(256);
        else XRD_MM_CHAIN(128);
#undef XRD_MM_CHAIN
    }
    ~MmStage()
    {
        if (h_nredo) cudaFreeHost(h_nredo);
    }
    MmState get(int ch)
    {
        MmState s;
        XRD_CUDA(cudaMemcpy(&s, d_carried.as<MmState>() + ch, sizeof(MmState), cudaMemcpyDeviceToHost));
        return s;
    }
    long long max_symbols(long long n) const
    {
        // every symbol advances the base by floor(mu + omega + gain_mu*mm) >= floor(omega_min - gain_mu)
        double adv = std::floor((double)prm.omega_mid - (double)prm.omega_lim - (double)prm.gain_mu);
        if (adv < 1.0) adv = 1.0;
        return (long long)((double)(n + MM_TAIL) / adv) + 8;
    }

    // in: index 0 = first new sample (MM_TAIL samples before it addressable); out: [nch][out_cap]
    // n_sym (host) receives per-channel counts.  Returns XRD_OK or XRD_E_OVERFLOW.
    int run(Counters &c, cudaStream_t st, const float2 *in, float2 *out, long long n, long long out_cap,
            long long in_stride, long long out_stride, int64_t *n_sym)
    {
        // segment plan: one chain per SM unless the caller pinned a segment length
        long long Ls = L;
        if (Ls <= 0) {
            const int per_ch = std::max(1, sm_count / nch);
            Ls = std::max<long long>(Lmin, (n + per_ch - 1) / per_ch);
            if (n <= single_max) Ls = std::max<long long>(n, 64);
        }
        const int nseg = (int)std::max<long long>(1, (n + Ls - 1) / Ls);
        const size_t tot = (size_t)nseg * nch;
        const long long cap_seg = max_symbols(std::min<long long>(Ls, std::max<long long>(n, 1))) + 64;
        d_entry.ensure(sizeof(MmState) * tot);
        d_exit.ensure(sizeof(MmState) * tot);
        d_redo.ensure(tot);
        d_segout.ensure(sizeof(MmSegOut) * tot);
        d_offsets.ensure(sizeof(long long) * (size_t)(nseg + 1) * nch);
        d_stage.ensure(sizeof(float2) * (size_t)cap_seg * tot);
        const long long stage_stride = cap_seg * nseg;
        dim3 grid(nseg, nch);
        // record the trajectory when there are hand-offs to certify and the kernel that records it will run
        W = W_user > 0 ? W_user : 80000;
        traj_on = use_delta && nseg > 1 && fast32(nt ? nt : 1024, n, Ls, cap_seg);
        if (!traj_on && W_user <= 0) W = 800000;
        if (traj_on) {
            d_traj.ensure(sizeof(int4) * (size_t)cap_seg * tot);
        }
        XRD_CUDA(cudaMemsetAsync(d_nredo.p, 0, 4 * sizeof(int), st));
        launch_chain(c, st, grid, in, n, Ls, nseg, cap_seg, 0, in_stride, stage_stride);
        for (int round = 0; nseg > 1 && round < nseg; round++) {
            XRD_CUDA(cudaMemsetAsync(d_nredo.p, 0, sizeof(int), st));
            dim3 vg((nseg + 127) / 128, nch);
            XRD_LAUNCH(c, mm_verify_kernel, vg, 128, 0, st, nseg, d_entry.as<MmState>(), d_exit.as<MmState>(),
                       d_redo.as<unsigned char>(), d_nredo.as<int>());
            // [4] segments flagged, [5] a chain stalled -- through mapped memory, see publish_kernel
            XRD_LAUNCH(c, publish_kernel, 1, 1, 0, st, h_nredo_dev + 4, d_nredo.as<int>(), d_nredo.as<int>() + 2);
            XRD_CUDA(cudaStreamSynchronize(st));
            const int nr = ((volatile int *)h_nredo)[4], stalled_now = ((volatile int *)h_nredo)[5];
            h_nredo[3] = stalled_now;
            if (nr == 0 || stalled_now) break;   // closed, or a chain stalled (reported as overflow below)
            rounds++;
            redone += (uint64_t)nr;
            // the delta kernel clears the flag of every segment it finishes; the chain kernel takes what is left
            if (traj_on) launch_delta(c, st, grid, in, n, Ls, nseg, cap_seg, in_stride, stage_stride);
            launch_chain(c, st, grid, in, n, Ls, nseg, cap_seg, 1, in_stride, stage_stride);
        }
        XRD_CUDA(cudaMemsetAsync(d_overflow.p, 0, sizeof(int), st));
        XRD_CUDA(cudaMemsetAsync(d_diag.p, 0, sizeof(MmDiag) * nch, st));
        XRD_LAUNCH(c, mm_offsets_kernel, nch, 32, 0, st, nseg, d_segout.as<MmSegOut>(), d_offsets.as<long long>(),
                   d_overflow.as<int>(), d_diag.as<MmDiag>());
        const int parts = std::max(1, std::min(64, (sm_count * 16) / std::max(1, nseg * nch)));
        dim3 cg((unsigned)nseg * parts, nch);
        XRD_LAUNCH(c, mm_compact_kernel, cg, 256, 0, st, d_stage.as<float2>(), out, nseg, cap_seg,
                   d_segout.as<MmSegOut>(), d_offsets.as<long long>(), out_cap, stage_stride, out_stride, out_i8,
                   d_diag.as<MmDiag>());
        XRD_LAUNCH(c, mm_rebase_kernel, (nch + 127) / 128, 128, 0, st, d_carried.as<MmState>(), d_exit.as<MmState>(),
                   nseg, nch, n);
        h_offsets.resize((size_t)(nseg + 1) * nch);
        std::vector<MmSegOut> so(tot);
        XRD_CUDA(cudaMemcpyAsync(h_offsets.data(), d_offsets.p, sizeof(long long) * h_offsets.size(),
                                 cudaMemcpyDeviceToHost, st));
        XRD_CUDA(cudaMemcpyAsync(so.data(), d_segout.p, sizeof(MmSegOut) * tot, cudaMemcpyDeviceToHost, st));
        h_diag.resize(nch);
        XRD_CUDA(cudaMemcpyAsync(h_diag.data(), d_diag.p, sizeof(MmDiag) * nch, cudaMemcpyDeviceToHost, st));
        XRD_CUDA(cudaMemcpyAsync(h_nredo + 1, d_overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        XRD_CUDA(cudaMemcpyAsync(h_nredo + 2, d_nredo.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        XRD_CUDA(cudaStreamSynchronize(st));
        bails += (uint64_t)h_nredo[2];
        int rc = XRD_OK;
        for (int ch = 0; ch < nch; ch++) {
            const long long cnt = h_offsets[(size_t)ch * (nseg + 1) + nseg];
            n_sym[ch] = cnt;
            if (cnt > out_cap) rc = XRD_E_OVERFLOW;
        }
        for (auto &s : so) {
            windows += (uint64_t)s.windows;
            iters += (uint64_t)s.iters;
        }
        if (h_nredo[1]) rc = XRD_E_OVERFLOW;
        return rc;
    }
};

}  // namespace xrd

using namespace xrd;

// ---------------------------------------------------------------------------------------
// Host -> device input copies of different demodulator handles on one device run one after the other, in
// submission order, each at the full speed of the link, instead of side by side at a share of it: with several calls
// in flight (one handle per host thread) the first call then starts computing while the second one
// copies, and the link never waits for a convoy of calls that all finish copying -- and all start computing -- together.
// ---------------------------------------------------------------------------------------
struct H2dGate {
    std::mutex mu;
    cudaEvent_t last[64] = {nullptr};
    static H2dGate &get()
    {
        static H2dGate g;
        return g;
    }
};
static int g_h2d_serialize = 1;
// host-input calls in flight in this process (all handles): when there are others, their kernels already fill the time
// this call's input copy takes, and cutting the copy in pieces only shortens the loop segments (measured, 4 calls in
// flight: 22.9 ms per 125 M-sample step with one piece, 24.9 with two; one call at a time: 46.4 against 42.8)
static std::atomic<int> g_host_calls{0};
struct HostCallScope {
    int others;
    HostCallScope() : others(g_host_calls.fetch_add(1)) {}
    ~HostCallScope() { g_host_calls.fetch_sub(1); }
};

// ---------------------------------------------------------------------------------------
// the chain
// ---------------------------------------------------------------------------------------
static size_t type_bytes(int type)
{
    // device-side ingest formats (XRD_RTLU8IQ carries a serial DC blocker and exists on the FIFO seam only)
    return type == XRD_FLOATIQ ? 8 : (type == XRD_S16IQ ? 4 : ((type == XRD_S8IQ || type == XRD_U8IQ) ? 2 : 0));
}

static const size_t XRD_FIFO_FLOATS = 1024 * 1024;   // FIFO_SIZE, Parameters.h:57
struct HostFifo {
    std::vector<float> buf;   // interleaved floats, ring (CircularBuffer<float>, demodulator.cpp:38)
    size_t head = 0, count = 0;
    float rtl_alpha = 0.f, rtl_avg = 0.f;   // DC blocker of the RTL u8 ingest (RtlFrontend.cpp:57-59,104-116)
};

struct xrd_demod {
    xrd_config cfg;
    int nch = 1, D = 1;
    float sps = 0.f;
    cudaStream_t stream = nullptr;
    Counters ctr;
    FirStage dec, rrc;
    SegStage<AgcLoop> agc;
    SegStage<CostasLoopK> costas;
    MmStage mm;
    // chunk buffers; per-channel stride = prefix + capacity
    DevBuf b_dec, b_agc, b_rrc, b_cos, b_sym, b_raw, b_i8;
    DevBuf b_ctraj;            // Costas trajectory record (state before every sample), laid out like b_cos
    DevBuf d_dec_hist;         // decimator history: the last ntaps_lpf - 1 input samples of every channel, cf32 (D > 1)
    long long cap_n = 0;       // input samples per channel the buffers are sized for
    bool hist_zeroed = false;
    std::vector<uint64_t> n_in, n_sym;
    Timer t_dec, t_agc, t_rrc, t_cos, t_mm;
    float ms[5] = {0, 0, 0, 0, 0};
    // host FIFO per channel (xrd_add_samples / xrd_process)
    std::mutex fifo_mu;
    std::vector<HostFifo> fifo;
    void *h_pin = nullptr;
    size_t h_pin_bytes = 0;
    std::vector<float> h_sym;
    std::vector<int64_t> h_cnt;

    ~xrd_demod()
    {
        if (h_pin) cudaFreeHost(h_pin);
        for (auto e : piece_ev) cudaEventDestroy(e);
        if (h2d_done) {
            // nobody may be left pointing at this handle's event
            std::lock_guard<std::mutex> gate(H2dGate::get().mu);
            for (auto &e : H2dGate::get().last)
                if (e == h2d_done) e = nullptr;
            cudaEventDestroy(h2d_done);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }
    long long agc_stride() const { return cap_n / D + rrc.hist(); }
    long long cos_stride() const { return cap_n / D + MM_TAIL; }
    long long nd_cap() const { return cap_n / D; }

    void ensure(long long n)
    {
        if (n <= cap_n) return;
        // carried prefixes (RRC history, M&M tail) must survive a regrow
        const long long old_cap = cap_n;
        const long long new_cap = n + n / 8;
        std::vector<float2> keep_agc, keep_cos;
        const int Hr = rrc.hist();
        if (old_cap > 0) {
            XRD_CUDA(cudaStreamSynchronize(stream));
            keep_agc.resize((size_t)nch * Hr);
            keep_cos.resize((size_t)nch * MM_TAIL);
            for (int ch = 0; ch < nch; ch++) {
                XRD_CUDA(cudaMemcpy(keep_agc.data() + (size_t)ch * Hr,
                                    b_agc.as<float2>() + (size_t)ch * (old_cap / D + Hr), sizeof(float2) * Hr,
                                    cudaMemcpyDeviceToHost));
                XRD_CUDA(cudaMemcpy(keep_cos.data() + (size_t)ch * MM_TAIL,
                                    b_cos.as<float2>() + (size_t)ch * (old_cap / D + MM_TAIL), sizeof(float2) * MM_TAIL,
                                    cudaMemcpyDeviceToHost));
            }
        }
        cap_n = new_cap - new_cap % D;
        const long long nd = cap_n / D;
        if (D > 1) b_dec.ensure(sizeof(float2) * (size_t)nd * nch);
        b_agc.ensure(sizeof(float2) * (size_t)(nd + Hr) * nch);
        b_rrc.ensure(sizeof(float2) * (size_t)nd * nch);
        b_cos.ensure(sizeof(float2) * (size_t)(nd + MM_TAIL) * nch);
        for (int ch = 0; ch < nch; ch++) {
            float2 *pa = b_agc.as<float2>() + (size_t)ch * (nd + Hr);
            float2 *pc = b_cos.as<float2>() + (size_t)ch * (nd + MM_TAIL);
            if (old_cap > 0) {
                XRD_CUDA(cudaMemcpy(pa, keep_agc.data() + (size_t)ch * Hr, sizeof(float2) * Hr, cudaMemcpyHostToDevice));
                XRD_CUDA(cudaMemcpy(pc, keep_cos.data() + (size_t)ch * MM_TAIL, sizeof(float2) * MM_TAIL,
                                    cudaMemcpyHostToDevice));
            } else {
                XRD_CUDA(cudaMemset(pa, 0, sizeof(float2) * Hr));
                XRD_CUDA(cudaMemset(pc, 0, sizeof(float2) * MM_TAIL));
            }
        }
    }

    // back to the just-created state: loop states, FIR histories, M&M tail, counters
    void reset()
    {
        XRD_CUDA(cudaStreamSynchronize(stream));
        agc.reset();
        costas.reset();
        mm.reset();
        const int Hd = (D > 1) ? dec.hist() : 0, Hr = rrc.hist();
        if (Hd) XRD_CUDA(cudaMemset(d_dec_hist.p, 0, sizeof(float2) * (size_t)Hd * nch));
        if (cap_n > 0) {
            for (int ch = 0; ch < nch; ch++) {
                XRD_CUDA(cudaMemset(b_agc.as<float2>() + (size_t)ch * agc_stride(), 0, sizeof(float2) * Hr));
                XRD_CUDA(cudaMemset(b_cos.as<float2>() + (size_t)ch * cos_stride(), 0, sizeof(float2) * MM_TAIL));
            }
        }
        std::fill(n_in.begin(), n_in.end(), 0);
        std::fill(n_sym.begin(), n_sym.end(), 0);
        std::lock_guard<std::mutex> lk(fifo_mu);
        for (auto &f : fifo) {
            f.head = f.count = 0;
            f.rtl_avg = 0.f;
        }
    }

    // Front half of the chain (everything at the sample rate: ingest, decimator, AGC, RRC, Costas) over input
    // samples [off, off + m) of a call whose raw input (n_total samples per channel) is iq_dev.  Pieces of one call
    // are processed in order; off > 0 is only used with one FLOATIQ channel and no decimation, where the earlier
    // pieces stay in place in front of this one (warm-ups reach back into them).
    void run_front(const void *iq_dev, long long n_total, long long off, long long m, int type)
    {
        const long long nd = m / D, nd_off = off / D;
        const int Hr = rrc.hist();
        const void *x = nullptr;   // AGC input, [nch] with stride xs (samples)
        long long xs = 0;
        bool x_s16 = false;
        if (D > 1) {
            // the decimator is the first kernel of the chain: it reads the raw samples in their ingest format
            t_dec.start(stream);
            dec.run_decim(ctr, stream, iq_dev, type, d_dec_hist.as<float2>(), b_dec.as<float2>(), nd, nch, n_total, nd_cap());
            t_dec.stop(stream);
            x = b_dec.as<float2>();
            xs = nd_cap();
        } else if (type == XRD_FLOATIQ) {
            x = (const float2 *)iq_dev + off;
            xs = n_total;
        } else if (type == XRD_S16IQ && agc.can_fuse_s16()) {
            // the AGC is the first kernel: it converts as it loads
            x = (const short2 *)iq_dev + off;
            xs = n_total;
            x_s16 = true;
        } else {
            // S8 / U8 (and S16 with non-default AGC kernels): one conversion pass into b_rrc, which the AGC has
            // consumed by the time the RRC filter writes it
            float2 *dst = b_rrc.as<float2>();
            const long long ds = nd_cap();
            for (int ch = 0; ch < nch; ch++) {
                float *o = reinterpret_cast<float *>(dst + (size_t)ch * ds);
                const size_t nf = (size_t)m * 2;
                const int blocks = (int)std::min<size_t>((nf + 1023) / 1024, 148 * 16);
                if (type == XRD_S16IQ) {
                    XRD_LAUNCH(ctr, (convert_kernel<short>), blocks, 256, 0, stream, (const short *)iq_dev + (size_t)ch * nf,
                               o, nf, 1.0f / 32768.f, 0.f);
                } else if (type == XRD_S8IQ) {
                    XRD_LAUNCH(ctr, (convert_kernel<signed char>), blocks, 256, 0, stream,
                               (const signed char *)iq_dev + (size_t)ch * nf, o, nf, 1.0f / 128.f, 0.f);
                } else {
                    XRD_LAUNCH(ctr, (convert_kernel<unsigned char>), blocks, 256, 0, stream,
                               (const unsigned char *)iq_dev + (size_t)ch * nf, o, nf, 1.0f / 128.f, 128.f);
                }
            }
            x = dst;
            xs = ds;
        }
        float2 *agc_out = b_agc.as<float2>() + Hr + nd_off;
        t_agc.start(stream);
        agc.run(ctr, stream, x, agc_out, nd, xs, agc_stride(), nd_off, x_s16);
        t_agc.stop(stream);
        t_rrc.start(stream);
        // (when D == 1 and the input was converted into b_rrc, AGC has consumed it by now)
        rrc.run(ctr, stream, agc_out, b_rrc.as<float2>() + nd_off, nd, nch, agc_stride(), nd_cap(), &b_agc);
        t_rrc.stop(stream);
        float2 *cos_out = b_cos.as<float2>() + MM_TAIL + nd_off;
        t_cos.start(stream);
        if (costas.guided) b_ctraj.ensure(sizeof(CostasState) * (size_t)cos_stride() * nch);
        costas.run(ctr, stream, b_rrc.as<float2>() + nd_off, cos_out, nd, nd_cap(), cos_stride(), nd_off, false,
                   costas.guided ? b_ctraj.as<CostasState>() + MM_TAIL + nd_off : nullptr);
        t_cos.stop(stream);
        ms[0] += (D > 1) ? t_dec.ms() : 0.f;
        ms[1] += t_agc.ms();
        ms[2] += t_rrc.ms();
        ms[3] += t_cos.ms();
    }

    // Back half: M&M over the Costas output of the whole call, history carries, totals
    int run_back(long long n, float2 *sym_dev, long long cap, int64_t *counts)
    {
        const long long nd = n / D;
        const int Hr = rrc.hist();
        XRD_LAUNCH(ctr, carry_prefix_kernel, nch, 256, sizeof(float2) * Hr, stream, b_agc.as<float2>(), Hr, nd,
                   agc_stride());
        float2 *cos_out = b_cos.as<float2>() + MM_TAIL;
        t_mm.start(stream);
        int rc = mm.run(ctr, stream, cos_out, sym_dev, nd, cap, cos_stride(), cap, counts);
        XRD_LAUNCH(ctr, carry_prefix_kernel, nch, 32, sizeof(float2) * MM_TAIL, stream, b_cos.as<float2>(), MM_TAIL, nd,
                   cos_stride());
        t_mm.stop(stream);
        XRD_CUDA(cudaStreamSynchronize(stream));
        ms[4] = t_mm.ms();
        for (int ch = 0; ch < nch; ch++) {
            n_in[ch] += (uint64_t)n;
            n_sym[ch] += (uint64_t)std::min<long long>(counts[ch], cap);
        }
        if (rc == XRD_E_OVERFLOW)
            g_error = "symbol output capacity too small (or non-finite samples stalled the timing loop)";
        return rc;
    }

    bool check_len(long long n, int64_t *counts, int &rc)
    {
        if (n % D) {
            g_error = "n_complex must be a multiple of the decimation";
            rc = XRD_E_ARG;
            return false;
        }
        if (n == 0) {
            for (int ch = 0; ch < nch; ch++) counts[ch] = 0;
            rc = XRD_OK;
            return false;
        }
        return true;
    }

    // iq_dev: [nch][n] samples of `type` on the device.  sym_dev: [nch][cap].
    int run_device(const void *iq_dev, long long n, int type, float2 *sym_dev, long long cap, int64_t *counts)
    {
        int rc = XRD_OK;
        if (!check_len(n, counts, rc)) return rc;
        ensure(n);
        for (float &v : ms) v = 0.f;
        run_front(iq_dev, n, 0, n, type);
        return run_back(n, sym_dev, cap, counts);
    }

    // Host input: the H2D copy is cut into pieces and the front half of the chain runs on every piece as it lands
    // (copy engine and SMs overlap); M&M runs once over the whole call.  Pieces are only used where earlier pieces
    // stay in place for the warm-ups of later ones (one FLOATIQ channel, no decimation) and the call is long.
    int run_host(const void *iq, long long n, int type, float2 *sym_dev, long long cap, int64_t *counts)
    {
        int rc = XRD_OK;
        if (!check_len(n, counts, rc)) return rc;
        const size_t sb = type_bytes(type);
        b_raw.ensure(sb * (size_t)n * nch);
        ensure(n);
        for (float &v : ms) v = 0.f;
        int pieces = 1;
        HostCallScope in_flight;
        if ((type == XRD_FLOATIQ || (type == XRD_S16IQ && agc.can_fuse_s16())) && D == 1 && nch == 1 && piece_min > 0 &&
            (in_flight.others == 0 || pieces_forced))
            pieces = (int)std::min<long long>(max_pieces, n / piece_min);
        if (!h2d_done) XRD_CUDA(cudaEventCreateWithFlags(&h2d_done, cudaEventDisableTiming));
        const int dev = cfg.device_ordinal & 63;
        if (pieces <= 1) {
            {
                std::unique_lock<std::mutex> gate(H2dGate::get().mu, std::defer_lock);
                if (g_h2d_serialize) {
                    gate.lock();
                    if (H2dGate::get().last[dev]) XRD_CUDA(cudaStreamWaitEvent(stream, H2dGate::get().last[dev], 0));
                }
                XRD_CUDA(cudaMemcpyAsync(b_raw.p, iq, sb * (size_t)n * nch, cudaMemcpyHostToDevice, stream));
                if (g_h2d_serialize) {
                    XRD_CUDA(cudaEventRecord(h2d_done, stream));
                    H2dGate::get().last[dev] = h2d_done;
                }
            }
            run_front(b_raw.p, n, 0, n, type);
            return run_back(n, sym_dev, cap, counts);
        }
        if (!copy_stream) XRD_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        while ((int)piece_ev.size() < pieces) {
            cudaEvent_t e;
            XRD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            piece_ev.push_back(e);
        }
        // piece boundaries on checkpoint multiples so that segments tile every piece the same way
        const long long step = ((n + pieces - 1) / pieces + WN_CKPT - 1) / WN_CKPT * WN_CKPT;
        // the copies must not start before earlier work on the compute stream is done with b_raw
        XRD_CUDA(cudaEventRecord(piece_ev[0], stream));
        XRD_CUDA(cudaStreamWaitEvent(copy_stream, piece_ev[0], 0));
        int np = 0;
        {
            std::unique_lock<std::mutex> gate(H2dGate::get().mu, std::defer_lock);
            if (g_h2d_serialize) {
                gate.lock();
                if (H2dGate::get().last[dev]) XRD_CUDA(cudaStreamWaitEvent(copy_stream, H2dGate::get().last[dev], 0));
            }
            for (long long off = 0; off < n; off += step, np++) {
                const long long m = std::min(step, n - off);
                XRD_CUDA(cudaMemcpyAsync((char *)b_raw.p + sb * off, (const char *)iq + sb * off, sb * (size_t)m,
                                         cudaMemcpyHostToDevice, copy_stream));
                XRD_CUDA(cudaEventRecord(piece_ev[np], copy_stream));
            }
            if (g_h2d_serialize) {
                XRD_CUDA(cudaEventRecord(h2d_done, copy_stream));
                H2dGate::get().last[dev] = h2d_done;
            }
        }
        np = 0;
        for (long long off = 0; off < n; off += step, np++) {
            const long long m = std::min(step, n - off);
            XRD_CUDA(cudaStreamWaitEvent(stream, piece_ev[np], 0));
            run_front(b_raw.p, n, off, m, type);
        }
        return run_back(n, sym_dev, cap, counts);
    }
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t h2d_done = nullptr;   // this handle's latest input copy (the next handle's copy queues behind it)
    std::vector<cudaEvent_t> piece_ev;
    long long piece_min = 16000000;   // samples; 0 disables the pieces
    bool pieces_forced = false;     // set_tuning(h2d_pieces): use the pieces whatever else is in flight (tests)
    int max_pieces = 2;             // more pieces shorten the Costas segments and cost more re-run rounds than the overlap wins
};

template <class F> static int guarded(F &&f)
{
    try {
        return f();
    } catch (const CudaError &e) {
        g_error = e.msg;
        return e.code;
    } catch (const std::bad_alloc &) {
        g_error = "host allocation failed";
        return XRD_E_NOMEM;
    }
}

static int select_device(int dev)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return XRD_E_CUDA;
    }
    if (dev < 0 || dev >= n) {
        g_error = "device ordinal out of range";
        return XRD_E_ARG;
    }
    e = cudaSetDevice(dev);
    if (e != cudaSuccess) {
        g_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return XRD_E_CUDA;
    }
    return XRD_OK;
}

extern "C" {

const char *xrd_version(void) { return "xritdemod_b200 0.1 (sm_100a)"; }

int xrd_device_check(int device, char *name, int name_cap, int *sm_count, int *cc_major, int *cc_minor)
{
    int rc = select_device(device);
    if (rc) return rc;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return XRD_E_CUDA;
    if (name && name_cap > 0) snprintf(name, name_cap, "%s", p.name);
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return (p.major == 10) ? XRD_OK : XRD_E_CUDA;
}

void xrd_config_defaults(xrd_config *cfg, int hrit)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->sample_rate = hrit ? 2500000u : 1250000u;
    cfg->symbol_rate = hrit ? 927000u : 293883u;   // Parameters.h:18,23
    cfg->rrc_alpha = hrit ? 0.3f : 0.5f;           // Parameters.h:19,24
    cfg->decimation = 1;                           // DEFAULT_DECIMATION
    cfg->rrc_taps = 63;                            // RRC_TAPS
    cfg->loop_order = 2;                           // LOOP_ORDER
    cfg->pll_alpha = 0.0037f;                      // demodulator.cpp:220: CLOCK_ALPHA, not PLL_ALPHA
    cfg->clock_alpha = 0.0037f;
    cfg->clock_mu = 0.5f;
    cfg->clock_omega_limit = 0.005f;
    cfg->agc_rate = 0.01f;
    cfg->agc_ref = 0.5f;
    cfg->agc_gain = 1.0f;
    cfg->agc_max_gain = 4000.f;
    cfg->device_ordinal = 0;
    cfg->n_channels = 1;
}

int xrd_create(const xrd_config *cfg, xrd_demod **out)
{
    if (!cfg || !out) return XRD_E_ARG;
    if (const char *e = getenv("XRD_H2D_SERIALIZE")) g_h2d_serialize = atoi(e);   // (measurements: 0 = copies side by side)
    *out = nullptr;
    if (cfg->n_channels < 1 || cfg->symbol_rate == 0 || cfg->sample_rate == 0 || cfg->rrc_taps < 1) {
        g_error = "bad config";
        return XRD_E_ARG;
    }
    if (cfg->loop_order != 2) {
        g_error = "only loop_order 2 (BPSK) is implemented";
        return XRD_E_ARG;
    }
    int rc = select_device(cfg->device_ordinal);
    if (rc) return rc;
    xrd_demod *d = new xrd_demod();
    rc = guarded([&]() {
        d->cfg = *cfg;
        d->nch = cfg->n_channels;
        d->D = cfg->decimation ? (int)cfg->decimation : 1;
        XRD_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        // demodulator.cpp:436-437 (float arithmetic)
        const float circuit = (float)cfg->sample_rate / (float)d->D;
        d->sps = circuit / (float)cfg->symbol_rate;
        std::vector<float> taps;
        design_rrc(1, circuit, cfg->symbol_rate, cfg->rrc_alpha, (int)cfg->rrc_taps, taps);   // :443
        d->rrc.init(1, taps.data(), (int)taps.size());                                        // :450
        if (d->D > 1) {
            design_lowpass(1, (double)cfg->sample_rate, circuit / 2, 100e3, taps);            // :444
            d->dec.init(d->D, taps.data(), (int)taps.size());                                 // :446
            d->d_dec_hist.ensure(sizeof(float2) * (size_t)d->dec.hist() * d->nch);
            XRD_CUDA(cudaMemset(d->d_dec_hist.p, 0, sizeof(float2) * (size_t)d->dec.hist() * d->nch));
        }
        d->agc.prm = AgcParams{cfg->agc_rate, cfg->agc_ref, cfg->agc_max_gain};               // :447
        d->agc.L = 2048;
        d->agc.W = 16384;
        d->agc.Ww = 8192;            // the AGC contracts fast: short warm-ups, twice the chains
        d->agc.chains_per_sm = 16;
        d->agc.use_wn = agc_wn_ok(cfg->agc_max_gain);
        AgcState a0{cfg->agc_gain, 0.f};
        d->agc.init(d->nch, a0);
        float ca, cb;
        costas_gains(cfg->pll_alpha, ca, cb);                                                 // :448
        d->costas.prm = CostasParams{ca, cb, 1.0f, -1.0f};
        d->costas.use_mirror = true;
        d->costas.use_wn = costas_wn_ok(d->costas.prm);
        d->costas.L = 4096;
        d->costas.W = 32768;
        d->costas.chains_per_sm_max = 20;   // many channels: as many chains as the kernel's registers let an SM hold
        d->agc.chains_per_sm_max = 24;
        CostasState c0{0.f, 0.f};
        d->costas.init(d->nch, c0);
        const float gain_omega = (cfg->clock_alpha * cfg->clock_alpha) / 4.0f;                // Parameters.h:33
        d->mm.init(d->nch, d->sps, gain_omega, cfg->clock_mu, cfg->clock_alpha, cfg->clock_omega_limit);  // :449
        d->n_in.assign(d->nch, 0);
        d->n_sym.assign(d->nch, 0);
        if (const char *e = getenv("XRD_H2D_PIECES")) d->max_pieces = std::max(1, atoi(e));   // (measurements)
        d->fifo.resize(d->nch);
        for (auto &f : d->fifo) f.rtl_alpha = (float)(1.f - exp(-1.0 / (cfg->sample_rate * 0.05f)));   // RtlFrontend.cpp:57
        return (int)XRD_OK;
    });
    if (rc) {
        delete d;
        return rc;
    }
    *out = d;
    return XRD_OK;
}

void xrd_destroy(xrd_demod *d)
{
    if (!d) return;
    cudaSetDevice(d->cfg.device_ordinal);
    delete d;
}

const char *xrd_last_error(const xrd_demod *) { return g_error.c_str(); }

int xrd_demod_device(xrd_demod *d, const void *iq_dev, size_t n_complex, int type, float *sym_dev, size_t cap,
                     int64_t *n_sym)
{
    if (!d || !iq_dev || !sym_dev || !n_sym || !type_bytes(type)) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        return d->run_device(iq_dev, (long long)n_complex, type, (float2 *)sym_dev, (long long)cap, n_sym);
    });
}

int xrd_demod_batch(xrd_demod *d, const void *iq, size_t n_complex, int type, float *sym_out, size_t cap,
                    int64_t *n_sym)
{
    if (!d || !iq || !sym_out || !n_sym || !type_bytes(type)) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        d->b_sym.ensure(sizeof(float2) * cap * d->nch);
        int rc = d->run_host(iq, (long long)n_complex, type, d->b_sym.as<float2>(), (long long)cap, n_sym);
        if (rc != XRD_OK && rc != XRD_E_OVERFLOW) return rc;
        for (int ch = 0; ch < d->nch; ch++) {
            const size_t cnt = (size_t)std::min<long long>(n_sym[ch], (long long)cap);
            XRD_CUDA(cudaMemcpyAsync(sym_out + 2 * cap * ch, d->b_sym.as<float2>() + cap * ch, sizeof(float2) * cnt,
                                     cudaMemcpyDeviceToHost, d->stream));
        }
        XRD_CUDA(cudaStreamSynchronize(d->stream));
        return rc;
    });
}

int xrd_demod_batch_i8(xrd_demod *d, const void *iq, size_t n_complex, int type, int8_t *soft_out, size_t cap,
                       int64_t *n_sym)
{
    if (!d || !iq || !soft_out || !n_sym || !type_bytes(type)) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        d->b_i8.ensure(cap * d->nch);
        // the M&M compaction writes the bytes itself; no cf32 symbol stream is materialised
        d->mm.out_i8 = d->b_i8.as<signed char>();
        int rc;
        try {
            rc = d->run_host(iq, (long long)n_complex, type, nullptr, (long long)cap, n_sym);
        } catch (...) {
            d->mm.out_i8 = nullptr;
            throw;
        }
        d->mm.out_i8 = nullptr;
        if (rc != XRD_OK && rc != XRD_E_OVERFLOW) return rc;
        for (int ch = 0; ch < d->nch; ch++) {
            const size_t cnt = (size_t)std::min<long long>(n_sym[ch], (long long)cap);
            XRD_CUDA(cudaMemcpyAsync(soft_out + cap * ch, d->b_i8.as<signed char>() + cap * ch, cnt, cudaMemcpyDeviceToHost,
                                     d->stream));
        }
        XRD_CUDA(cudaStreamSynchronize(d->stream));
        return rc;
    });
}

// onSamplesAvailable's conversions (demodulator.cpp:56-70) plus the two u8 formats the reference's front ends convert
// themselves before calling it: SpyServer (SpyServerFrontend.cpp:404-407) and RTL-SDR (RtlFrontend.cpp:27,104-116)
static void convert_to_fifo(float *dst, const void *data, size_t first, size_t count, int type, HostFifo &f)
{
    switch (type) {
    case XRD_FLOATIQ:
        memcpy(dst, (const float *)data + first, sizeof(float) * count);
        break;
    case XRD_S16IQ: {
        const int16_t *p = (const int16_t *)data + first;
        for (size_t i = 0; i < count; i++) dst[i] = p[i] / 32768.f;   // demodulator.cpp:60-61
        break;
    }
    case XRD_S8IQ: {
        const int8_t *p = (const int8_t *)data + first;
        for (size_t i = 0; i < count; i++) dst[i] = p[i] / 128.f;     // demodulator.cpp:67-68
        break;
    }
    case XRD_U8IQ: {
        const uint8_t *p = (const uint8_t *)data + first;
        for (size_t i = 0; i < count; i++) dst[i] = (p[i] - 128) / 128.f;   // SpyServerFrontend.cpp:406
        break;
    }
    default: {   // XRD_RTLU8IQ
        // RtlFrontend::internalCallback: lut[v] = (v - 128) * (1.f / 127.f), then a one-pole DC blocker.  The
        // reference tests `i % 1`, which is never true, so EVERY float (I and Q alike) runs through the one
        // average `iavg`; that is reproduced as is.
        const uint8_t *p = (const uint8_t *)data + first;
        float avg = f.rtl_avg;
        const float alpha = f.rtl_alpha;
        for (size_t i = 0; i < count; i++) {
            float v = ((int)p[i] - 128) * (1.f / 127.f);
            avg += alpha * (v - avg);
            v -= avg;
            dst[i] = v;
        }
        f.rtl_avg = avg;
        break;
    }
    }
}

int xrd_add_samples(xrd_demod *d, int channel, const void *data, int n_complex, int type)
{
    if (!d || channel < 0 || channel >= d->nch || n_complex < 0 || (!data && n_complex)) return XRD_E_ARG;
    if (type < XRD_FLOATIQ || type > XRD_RTLU8IQ) {
        g_error = "Unknown sample type";   // demodulator.cpp:71-73
        return XRD_E_ARG;
    }
    std::lock_guard<std::mutex> lk(d->fifo_mu);
    HostFifo &f = d->fifo[channel];
    if (f.buf.empty()) f.buf.resize(XRD_FIFO_FLOATS);
    const size_t nf = (size_t)n_complex * 2;
    if (f.count + nf > XRD_FIFO_FLOATS) {
        g_error = "Input Samples Fifo is overflowing!";   // demodulator.cpp:104-106
        return XRD_E_OVERFLOW;
    }
    // two spans of the ring, converted in bulk (the reference pushes float by float under its mutex)
    const size_t w = (f.head + f.count) % XRD_FIFO_FLOATS;
    const size_t first = std::min(nf, XRD_FIFO_FLOATS - w);
    convert_to_fifo(f.buf.data() + w, data, 0, first, type, f);
    if (nf > first) convert_to_fifo(f.buf.data(), data, first, nf - first, type, f);
    f.count += nf;
    return XRD_OK;
}

int64_t xrd_process(xrd_demod *d, int64_t min_samples, xrd_symbols_cb cb, void *user)
{
    if (!d) return XRD_E_ARG;
    return guarded([&]() -> int {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        size_t n = 0;
        {
            std::lock_guard<std::mutex> lk(d->fifo_mu);
            n = (size_t)-1;
            for (auto &f : d->fifo) n = std::min(n, f.count / 2);
            n -= n % d->D;   // keep the remainder queued (the reference drops it, demodulator.cpp:137)
            if (n == 0 || (int64_t)n < min_samples) return 0;
            const size_t bytes = sizeof(float) * 2 * n * d->nch;
            if (bytes > d->h_pin_bytes) {
                if (d->h_pin) cudaFreeHost(d->h_pin);
                d->h_pin = nullptr;
                d->h_pin_bytes = 0;
                XRD_CUDA(cudaMallocHost(&d->h_pin, bytes));
                d->h_pin_bytes = bytes;
            }
            // copy out, but leave the samples queued until the chain has succeeded on them
            float *dst = (float *)d->h_pin;
            for (int ch = 0; ch < d->nch; ch++) {
                const HostFifo &f = d->fifo[ch];
                const size_t first = std::min(2 * n, XRD_FIFO_FLOATS - f.head);
                memcpy(dst + (size_t)ch * 2 * n, f.buf.data() + f.head, sizeof(float) * first);
                if (2 * n > first) memcpy(dst + (size_t)ch * 2 * n + first, f.buf.data(), sizeof(float) * (2 * n - first));
            }
        }
        const size_t cap = (size_t)d->mm.max_symbols((long long)(n / d->D));
        d->h_sym.resize(2 * cap * d->nch);
        d->h_cnt.assign(d->nch, 0);
        int rc = xrd_demod_batch(d, d->h_pin, n, XRD_FLOATIQ, d->h_sym.data(), cap, d->h_cnt.data());
        if (rc) return rc;   // nothing was dequeued; the loop state may have advanced: xrd_reset before retrying
        {
            std::lock_guard<std::mutex> lk(d->fifo_mu);
            for (auto &f : d->fifo) {
                f.head = (f.head + 2 * n) % XRD_FIFO_FLOATS;
                f.count -= 2 * n;
            }
        }
        if (cb)
            for (int ch = 0; ch < d->nch; ch++) cb(user, ch, d->h_sym.data() + 2 * cap * ch, (int)d->h_cnt[ch]);
        return (int)n;
    });
}

int64_t xrd_symbol_capacity(const xrd_demod *d, size_t n_complex)
{
    if (!d) return XRD_E_ARG;
    return (int64_t)d->mm.max_symbols((long long)(n_complex / (size_t)d->D));
}

int xrd_soft_i8(xrd_demod *d, const float *sym, size_t n, int8_t *out)
{
    if (!d || (!sym && n) || (!out && n)) return XRD_E_ARG;
    if (!n) return XRD_OK;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        d->b_sym.ensure(sizeof(float2) * n);
        d->b_i8.ensure(n);
        XRD_CUDA(cudaMemcpyAsync(d->b_sym.p, sym, sizeof(float2) * n, cudaMemcpyHostToDevice, d->stream));
        const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
        XRD_LAUNCH(d->ctr, soft_i8_kernel, blocks, 256, 0, d->stream, d->b_sym.as<float2>(), d->b_i8.as<signed char>(),
                   (long long)n);
        XRD_CUDA(cudaMemcpyAsync(out, d->b_i8.p, n, cudaMemcpyDeviceToHost, d->stream));
        XRD_CUDA(cudaStreamSynchronize(d->stream));
        return (int)XRD_OK;
    });
}

int xrd_reset(xrd_demod *d)
{
    if (!d) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        d->reset();
        return (int)XRD_OK;
    });
}

void *xrd_stream(xrd_demod *d) { return d ? (void *)d->stream : nullptr; }

int xrd_get_state(xrd_demod *d, int channel, xrd_loop_state *st)
{
    if (!d || !st || channel < 0 || channel >= d->nch) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        XRD_CUDA(cudaStreamSynchronize(d->stream));
        AgcState a = d->agc.get(channel);
        CostasState c = d->costas.get(channel);
        MmState m = d->mm.get(channel);
        st->agc_gain = a.gain;
        st->costas_phase = c.phase;
        st->costas_freq = c.freq;
        st->mm_mu = m.mu;
        st->mm_omega = m.omega;
        st->mm_p0[0] = m.p0.x;
        st->mm_p0[1] = m.p0.y;
        st->mm_p1[0] = m.p1.x;
        st->mm_p1[1] = m.p1.y;
        st->mm_next = m.ii;
        st->n_in = d->n_in[channel];
        st->n_sym = d->n_sym[channel];
        return (int)XRD_OK;
    });
}

int xrd_set_state(xrd_demod *d, int channel, const xrd_loop_state *st)
{
    if (!d || !st || channel < 0 || channel >= d->nch) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        XRD_CUDA(cudaStreamSynchronize(d->stream));
        const AgcState a{st->agc_gain, 0.f};
        const CostasState c{st->costas_phase, st->costas_freq};
        MmState m;
        m.ii = st->mm_next;
        m.mu = st->mm_mu;
        m.omega = st->mm_omega;
        m.p0 = make_float2(st->mm_p0[0], st->mm_p0[1]);
        m.p1 = make_float2(st->mm_p1[0], st->mm_p1[1]);
        XRD_CUDA(cudaMemcpy(d->agc.d_carried.as<AgcState>() + channel, &a, sizeof a, cudaMemcpyHostToDevice));
        XRD_CUDA(cudaMemcpy(d->costas.d_carried.as<CostasState>() + channel, &c, sizeof c, cudaMemcpyHostToDevice));
        XRD_CUDA(cudaMemcpy(d->mm.d_carried.as<MmState>() + channel, &m, sizeof m, cudaMemcpyHostToDevice));
        d->n_in[channel] = st->n_in;
        d->n_sym[channel] = st->n_sym;
        return (int)XRD_OK;
    });
}

// ---- checkpoint: everything the five operators hold between calls, all channels ----
namespace {
struct CkptHeader {
    uint32_t magic, version, nch, D, Hd, Hr, tail, reserved;
    xrd_config cfg;
};
const uint32_t CKPT_MAGIC = 0x43445258u;   // "XRDC"
size_t ckpt_bytes(const xrd_demod *d)
{
    const size_t Hd = (d->D > 1) ? (size_t)d->dec.hist() : 0, Hr = (size_t)d->rrc.hist();
    return sizeof(CkptHeader) + (size_t)d->nch * (sizeof(xrd_loop_state) + sizeof(float) + sizeof(float2) * (Hd + Hr + MM_TAIL));
}
}  // namespace

size_t xrd_checkpoint_size(const xrd_demod *d) { return d ? ckpt_bytes(d) : 0; }

int xrd_checkpoint_save(xrd_demod *d, void *blob, size_t cap)
{
    if (!d || !blob) return XRD_E_ARG;
    if (cap < ckpt_bytes(d)) return XRD_E_OVERFLOW;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        XRD_CUDA(cudaStreamSynchronize(d->stream));
        const int Hd = (d->D > 1) ? d->dec.hist() : 0, Hr = d->rrc.hist();
        CkptHeader h;
        memset(&h, 0, sizeof h);
        h.magic = CKPT_MAGIC;
        h.version = 1;
        h.nch = (uint32_t)d->nch;
        h.D = (uint32_t)d->D;
        h.Hd = (uint32_t)Hd;
        h.Hr = (uint32_t)Hr;
        h.tail = MM_TAIL;
        h.cfg = d->cfg;
        char *p = (char *)blob;
        memcpy(p, &h, sizeof h);
        p += sizeof h;
        for (int ch = 0; ch < d->nch; ch++) {
            xrd_loop_state st;
            int rc = xrd_get_state(d, ch, &st);
            if (rc) return rc;
            memcpy(p, &st, sizeof st);
            p += sizeof st;
            float avg;
            {
                std::lock_guard<std::mutex> lk(d->fifo_mu);
                avg = d->fifo[ch].rtl_avg;
            }
            memcpy(p, &avg, sizeof avg);
            p += sizeof avg;
            // histories live in the prefixes of the chunk buffers; all zero before the first call
            auto grab = [&](const DevBuf &b, long long stride, int count) {
                if (count <= 0) return;
                if (d->cap_n > 0) XRD_CUDA(cudaMemcpy(p, b.as<float2>() + (size_t)ch * stride, sizeof(float2) * count, cudaMemcpyDeviceToHost));
                else memset(p, 0, sizeof(float2) * count);
                p += sizeof(float2) * count;
            };
            if (Hd) {   // the decimator history has its own buffer, valid from creation on
                XRD_CUDA(cudaMemcpy(p, d->d_dec_hist.as<float2>() + (size_t)ch * Hd, sizeof(float2) * Hd, cudaMemcpyDeviceToHost));
                p += sizeof(float2) * Hd;
            }
            grab(d->b_agc, d->agc_stride(), Hr);
            grab(d->b_cos, d->cos_stride(), MM_TAIL);
        }
        return (int)XRD_OK;
    });
}

int xrd_checkpoint_load(xrd_demod *d, const void *blob, size_t bytes)
{
    if (!d || !blob || bytes < sizeof(CkptHeader)) return XRD_E_ARG;
    CkptHeader h;
    memcpy(&h, blob, sizeof h);
    const int Hd = (d->D > 1) ? d->dec.hist() : 0, Hr = d->rrc.hist();
    xrd_config a = h.cfg, b = d->cfg;
    a.device_ordinal = b.device_ordinal = 0;   // a checkpoint may move between devices
    if (h.magic != CKPT_MAGIC || h.version != 1 || (int)h.nch != d->nch || (int)h.D != d->D || (int)h.Hd != Hd ||
        (int)h.Hr != Hr || h.tail != MM_TAIL || memcmp(&a, &b, sizeof a) != 0 || bytes < ckpt_bytes(d)) {
        g_error = "checkpoint does not belong to a demodulator with this configuration";
        return XRD_E_STATE;
    }
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(d->cfg.device_ordinal));
        XRD_CUDA(cudaStreamSynchronize(d->stream));
        d->ensure(std::max<long long>(d->cap_n, 1024LL * d->D));
        const char *p = (const char *)blob + sizeof h;
        for (int ch = 0; ch < d->nch; ch++) {
            xrd_loop_state st;
            memcpy(&st, p, sizeof st);
            p += sizeof st;
            int rc = xrd_set_state(d, ch, &st);
            if (rc) return rc;
            float avg;
            memcpy(&avg, p, sizeof avg);
            p += sizeof avg;
            {
                std::lock_guard<std::mutex> lk(d->fifo_mu);
                d->fifo[ch].rtl_avg = avg;
            }
            auto put = [&](DevBuf &buf, long long stride, int count) {
                if (count <= 0) return;
                XRD_CUDA(cudaMemcpy(buf.as<float2>() + (size_t)ch * stride, p, sizeof(float2) * count, cudaMemcpyHostToDevice));
                p += sizeof(float2) * count;
            };
            put(d->d_dec_hist, Hd, Hd);
            put(d->b_agc, d->agc_stride(), Hr);
            put(d->b_cos, d->cos_stride(), MM_TAIL);
        }
        return (int)XRD_OK;
    });
}

int xrd_get_diag(xrd_demod *d, int channel, xrd_diag *out)
{
    if (!d || !out || channel < 0 || channel >= d->nch) return XRD_E_ARG;
    memset(out, 0, sizeof *out);
    if ((int)d->mm.h_diag.size() <= channel) return XRD_OK;   // no call yet
    const MmDiag &g = d->mm.h_diag[channel];
    out->n_frame = g.n_frame;
    memcpy(out->frame, g.frame, sizeof out->frame);
    out->n_symbols = g.n;
    if (g.n) {
        const double m1 = g.sum_abs_i / (double)g.n, m2 = g.sum_sq_i / (double)g.n, q2 = g.sum_sq_q / (double)g.n;
        out->mean_abs_i = m1;
        out->mean_sq_i = m2;
        out->mean_sq_q = q2;
        const double noise = m2 - m1 * m1;
        out->snr_db = (noise > 0 && m1 > 0) ? (float)(10.0 * log10(m1 * m1 / noise)) : 0.f;
        out->lock = (m2 + q2 > 0) ? (float)(m2 / (m2 + q2)) : 0.f;
    }
    return XRD_OK;
}

int xrd_set_tuning(xrd_demod *d, const xrd_tuning *t)
{
    if (!d || !t) return XRD_E_ARG;
    if (t->agc_seg < 0 || t->agc_warm < 0 || t->costas_seg < 0 || t->costas_warm < 0 || t->mm_seg < 0 || t->mm_warm < 0)
        return XRD_E_ARG;
    auto lanes_ok = [](int v, int lo) { return v == 0 || (v >= lo && v <= 1024 && (v & (v - 1)) == 0); };
    if (!lanes_ok(t->mm_lanes, 128) || !lanes_ok(t->mm_walk_lanes, 128) || t->mm_walk_lanes > 512) return XRD_E_ARG;
    if (t->mm_kernel < 0 || t->mm_kernel > 2 || t->mm_rerun < 0 || t->mm_rerun > 2) return XRD_E_ARG;
    if (t->loop_kernel < 0 || t->loop_kernel > 7 || t->rerun_kernel < 0 || t->rerun_kernel == 1 || t->rerun_kernel > 7)
        return XRD_E_ARG;
    if (t->h2d_pieces < 0 || t->h2d_pieces > 16 || t->h2d_piece_min_ki < 0) return XRD_E_ARG;
    if (t->costas_chains_per_sm < 0 || t->costas_chains_per_sm > 64 || t->agc_chains_per_sm < 0 || t->agc_chains_per_sm > 64)
        return XRD_E_ARG;
    if (t->agc_seg) d->agc.L = d->agc.Lw = t->agc_seg;
    if (t->agc_warm) d->agc.W = d->agc.Ww = t->agc_warm;
    if (t->costas_seg) d->costas.L = d->costas.Lw = t->costas_seg;
    if (t->costas_warm) d->costas.W = d->costas.Ww = t->costas_warm;
    if (t->loop_kernel == 1) d->agc.use_wn = d->costas.use_wn = false;
    else if (t->loop_kernel) {
        d->agc.use_wn = agc_wn_ok(d->agc.prm.max_gain);
        d->costas.use_wn = costas_wn_ok(d->costas.prm);
        d->agc.wn_variant = d->costas.wn_variant = t->loop_kernel;
        d->agc.redo_variant = d->costas.redo_variant = t->loop_kernel;   // unless rerun_kernel says otherwise below
    }
    if (t->rerun_kernel) d->agc.redo_variant = d->costas.redo_variant = t->rerun_kernel;
    if (t->agc_kernel < 0 || t->agc_kernel > 7) return XRD_E_ARG;
    if (t->agc_kernel == 1) d->agc.use_wn = false;
    else if (t->agc_kernel) {
        d->agc.use_wn = agc_wn_ok(d->agc.prm.max_gain);
        d->agc.wn_variant = t->agc_kernel;
    }
    if (t->chase) d->agc.chase = d->costas.chase = (t->chase == 1);
    if (t->guided < 0 || t->guided > 2) return XRD_E_ARG;
    if (t->guided) d->costas.guided = (t->guided == 1);
    if (t->costas_chains_per_sm) d->costas.chains_per_sm = t->costas_chains_per_sm;
    if (t->agc_chains_per_sm) d->agc.chains_per_sm = t->agc_chains_per_sm;
    if (t->mm_seg) d->mm.L = std::max<long long>(t->mm_seg, 64);
    if (t->mm_lanes) d->mm.nt = t->mm_lanes;
    if (t->mm_kernel) d->mm.force64 = (t->mm_kernel == 2);
    if (t->mm_rerun) d->mm.use_delta = (t->mm_rerun == 1);
    if (t->mm_walk_lanes) d->mm.delta_nt = t->mm_walk_lanes;
    if (t->mm_warm) d->mm.W_user = t->mm_warm;
    if (t->h2d_pieces) {
        d->max_pieces = t->h2d_pieces;
        d->pieces_forced = true;
    }
    if (t->h2d_piece_min_ki) d->piece_min = (long long)t->h2d_piece_min_ki * 1024;
    return XRD_OK;
}

int xrd_get_stats(xrd_demod *d, xrd_stats *s)
{
    if (!d || !s) return XRD_E_ARG;
    memset(s, 0, sizeof *s);
    s->kernel_launches = d->ctr.launches;
    s->agc_rounds = d->agc.rounds;
    s->costas_rounds = d->costas.rounds;
    s->mm_rounds = d->mm.rounds;
    s->agc_redo = d->agc.redone;
    s->costas_redo = d->costas.redone;
    s->mm_redo = d->mm.redone;
    s->mm_windows = d->mm.windows;
    s->mm_iters = d->mm.iters;
    s->mm_bail = d->mm.bails;
    s->agc_iters = d->agc.wn_iters();
    s->costas_iters = d->costas.wn_iters();
    s->ms_fir_dec = d->ms[0];
    s->ms_agc = d->ms[1];
    s->ms_fir_rrc = d->ms[2];
    s->ms_costas = d->ms[3];
    s->ms_mm = d->ms[4];
    return XRD_OK;
}

// ---- designers ----
int xrd_design_rrc(double gain, double fs, double rs, double alpha, int ntaps, float *taps, int cap)
{
    if (!taps || ntaps < 1 || fs <= 0 || rs <= 0) return XRD_E_ARG;
    std::vector<float> t;
    int n = design_rrc(gain, fs, rs, alpha, ntaps, t);
    if (n > cap) return -n;
    memcpy(taps, t.data(), sizeof(float) * n);
    return n;
}

int xrd_design_lowpass(double gain, double fs, double fc, double tw, float *taps, int cap)
{
    if (!taps || fs <= 0 || tw <= 0) return XRD_E_ARG;
    std::vector<float> t;
    int n = design_lowpass(gain, fs, fc, tw, t);
    if (n > cap) return -n;
    memcpy(taps, t.data(), sizeof(float) * n);
    return n;
}

void xrd_mmse_table(float *t) { mmse_table(t); }
void xrd_costas_gains(float bw, float *a, float *b) { costas_gains(bw, *a, *b); }

}  // extern "C"

// ---------------------------------------------------------------------------------------
// stage operators (SatHelper::FirFilter / AGC / CostasLoop / ClockRecovery on host buffers)
// ---------------------------------------------------------------------------------------
struct xrd_stage {
    enum Kind { FIR, AGC, COSTAS, MM } kind;
    int device = 0;
    cudaStream_t stream = nullptr;
    Counters ctr;
    FirStage fir;
    SegStage<AgcLoop> agc;
    SegStage<CostasLoopK> costas;
    MmStage mm;
    DevBuf b_in, b_out, b_traj;
    long long cap = 0;
    int prefix = 0;
    ~xrd_stage()
    {
        if (stream) cudaStreamDestroy(stream);
    }
    void ensure(long long n_in, long long n_out)
    {
        if (n_in > cap) {
            std::vector<float2> keep(prefix);
            if (cap > 0 && prefix) XRD_CUDA(cudaMemcpy(keep.data(), b_in.p, sizeof(float2) * prefix, cudaMemcpyDeviceToHost));
            const bool had = cap > 0;
            cap = n_in + n_in / 4;
            b_in.ensure(sizeof(float2) * (size_t)(cap + prefix));
            if (prefix) {
                if (had) XRD_CUDA(cudaMemcpy(b_in.p, keep.data(), sizeof(float2) * prefix, cudaMemcpyHostToDevice));
                else XRD_CUDA(cudaMemset(b_in.p, 0, sizeof(float2) * prefix));
            }
        }
        b_out.ensure(sizeof(float2) * (size_t)std::max<long long>(n_out, 1));
    }
};

static int stage_new(int device, xrd_stage::Kind k, xrd_stage **out, xrd_stage *&s)
{
    if (!out) return XRD_E_ARG;
    *out = nullptr;
    int rc = select_device(device);
    if (rc) return rc;
    s = new xrd_stage();
    s->kind = k;
    s->device = device;
    return XRD_OK;
}

extern "C" {

int xrd_fir_create(int device, unsigned decimation, const float *taps, int ntaps, xrd_stage **out)
{
    if (!taps || ntaps < 1) return XRD_E_ARG;
    xrd_stage *s = nullptr;
    int rc = stage_new(device, xrd_stage::FIR, out, s);
    if (rc) return rc;
    rc = guarded([&]() {
        XRD_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->fir.init(decimation, taps, ntaps);
        s->prefix = ntaps - 1;
        return (int)XRD_OK;
    });
    if (rc) { delete s; return rc; }
    *out = s;
    return XRD_OK;
}

int xrd_agc_create(int device, float rate, float reference, float gain, float max_gain, xrd_stage **out)
{
    xrd_stage *s = nullptr;
    int rc = stage_new(device, xrd_stage::AGC, out, s);
    if (rc) return rc;
    rc = guarded([&]() {
        XRD_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->agc.prm = AgcParams{rate, reference, max_gain};
        s->agc.L = 2048;
        s->agc.W = 16384;
        s->agc.Ww = 8192;
        s->agc.chains_per_sm = 16;
        s->agc.use_wn = agc_wn_ok(max_gain);
        s->agc.init(1, AgcState{gain, 0.f});
        return (int)XRD_OK;
    });
    if (rc) { delete s; return rc; }
    *out = s;
    return XRD_OK;
}

int xrd_costas_create(int device, float loop_bw, int order, xrd_stage **out)
{
    if (order != 2) {
        g_error = "only order 2 (BPSK) is implemented";
        return XRD_E_ARG;
    }
    xrd_stage *s = nullptr;
    int rc = stage_new(device, xrd_stage::COSTAS, out, s);
    if (rc) return rc;
    rc = guarded([&]() {
        XRD_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        float a, b;
        costas_gains(loop_bw, a, b);
        s->costas.prm = CostasParams{a, b, 1.0f, -1.0f};
        s->costas.use_mirror = true;
        s->costas.use_wn = costas_wn_ok(s->costas.prm);
        s->costas.L = 4096;
        s->costas.W = 32768;
        s->costas.init(1, CostasState{0.f, 0.f});
        return (int)XRD_OK;
    });
    if (rc) { delete s; return rc; }
    *out = s;
    return XRD_OK;
}

int xrd_clock_recovery_create(int device, float omega, float gain_omega, float mu, float gain_mu, float omega_rel_limit,
                              xrd_stage **out)
{
    xrd_stage *s = nullptr;
    int rc = stage_new(device, xrd_stage::MM, out, s);
    if (rc) return rc;
    rc = guarded([&]() {
        XRD_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->mm.init(1, omega, gain_omega, mu, gain_mu, omega_rel_limit);
        s->prefix = MM_TAIL;
        return (int)XRD_OK;
    });
    if (rc) { delete s; return rc; }
    *out = s;
    return XRD_OK;
}

int xrd_stage_work(xrd_stage *s, const float *in, float *out, int length)
{
    if (!s || length < 0 || (length && (!in || !out))) return XRD_E_ARG;
    if (length == 0) return 0;
    return guarded([&]() -> int {
        XRD_CUDA(cudaSetDevice(s->device));
        const long long n_out = length;
        const long long n_in = (s->kind == xrd_stage::FIR) ? n_out * s->fir.D : n_out;
        const long long out_cap = (s->kind == xrd_stage::MM) ? s->mm.max_symbols(n_in) : n_out;
        s->ensure(n_in, out_cap);
        float2 *x = s->b_in.as<float2>() + s->prefix;
        XRD_CUDA(cudaMemcpyAsync(x, in, sizeof(float2) * n_in, cudaMemcpyHostToDevice, s->stream));
        int ret = 0;
        long long n_copy = n_out;
        switch (s->kind) {
        case xrd_stage::FIR:
            if (s->fir.D > 1)   // history = the prefix of b_in, right before x
                s->fir.run_decim(s->ctr, s->stream, x, XRD_FLOATIQ, s->b_in.as<float2>(), s->b_out.as<float2>(), n_out, 1, 0, 0);
            else
                s->fir.run(s->ctr, s->stream, x, s->b_out.as<float2>(), n_out, 1, 0, 0, &s->b_in);
            break;
        case xrd_stage::AGC:
            s->agc.run(s->ctr, s->stream, x, s->b_out.as<float2>(), n_in, 0, 0);
            break;
        case xrd_stage::COSTAS:
            if (s->costas.guided) s->b_traj.ensure(sizeof(CostasState) * (size_t)std::max<long long>(n_in, 1));
            s->costas.run(s->ctr, s->stream, x, s->b_out.as<float2>(), n_in, 0, 0, 0, false,
                          s->costas.guided ? s->b_traj.as<CostasState>() : nullptr);
            break;
        case xrd_stage::MM: {
            int64_t cnt = 0;
            int rc = s->mm.run(s->ctr, s->stream, x, s->b_out.as<float2>(), n_in, out_cap, 0, 0, &cnt);
            if (rc) {
                g_error = "symbol staging overflow";
                return rc;
            }
            ret = (int)cnt;
            n_copy = cnt;
            break;
        }
        }
        if (s->prefix && !(s->kind == xrd_stage::FIR && s->fir.D > 1))   // (run_decim has advanced its history itself)
            XRD_LAUNCH(s->ctr, carry_prefix_kernel, 1, 256, sizeof(float2) * s->prefix, s->stream, s->b_in.as<float2>(),
                       s->prefix, n_in, 0);
        XRD_CUDA(cudaMemcpyAsync(out, s->b_out.p, sizeof(float2) * n_copy, cudaMemcpyDeviceToHost, s->stream));
        XRD_CUDA(cudaStreamSynchronize(s->stream));
        return ret;
    });
}

int xrd_stage_set_tuning(xrd_stage *s, int64_t seg, int64_t warm)
{
    if (!s || seg < 0 || warm < 0) return XRD_E_ARG;
    switch (s->kind) {
    case xrd_stage::AGC:
        if (seg) s->agc.L = s->agc.Lw = (int)seg;
        if (warm) s->agc.W = s->agc.Ww = (int)warm;
        break;
    case xrd_stage::COSTAS:
        if (seg) s->costas.L = s->costas.Lw = (int)seg;
        if (warm) s->costas.W = s->costas.Ww = (int)warm;
        break;
    case xrd_stage::MM:
        if (seg) s->mm.L = std::max<long long>(seg, 64);
        if (warm) s->mm.W_user = warm;
        break;
    default:
        break;
    }
    return XRD_OK;
}

int xrd_stage_set_loop_kernel(xrd_stage *s, int kernel)
{
    if (!s || kernel < 0 || kernel > 7) return XRD_E_ARG;
    const bool wn = kernel != 1;
    if (s->kind == xrd_stage::AGC) s->agc.use_wn = wn && agc_wn_ok(s->agc.prm.max_gain);
    if (s->kind == xrd_stage::COSTAS) s->costas.use_wn = wn && costas_wn_ok(s->costas.prm);
    if (kernel >= 2) s->agc.wn_variant = s->costas.wn_variant = s->agc.redo_variant = s->costas.redo_variant = kernel;
    return XRD_OK;
}

void xrd_stage_destroy(xrd_stage *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    delete s;
}

const char *xrd_stage_last_error(const xrd_stage *) { return g_error.c_str(); }

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Decoder front half: sync-word correlation, frame alignment, phase fix, Viterbi, NRZ-M
// (reference decoder/src/newdecoder.cpp:212-290) on the soft-symbol byte stream xrd_demod_batch_i8 emits
// ---------------------------------------------------------------------------------------
#include "xrd_decoder.cuh"

struct xrd_decoder_front {
    int device = 0, lrit = 1, soft_mode = 0;
    cudaStream_t stream = nullptr;
    Counters ctr;
    DevBuf d_soft, d_bits, d_blockmax, d_frames, d_out, d_err, d_scalars, d_last[2], d_key;
    int cur = 0;
    unsigned long long words[2] = {0, 0};
    ~xrd_decoder_front()
    {
        if (stream) cudaStreamDestroy(stream);
    }
    void pack(const uint8_t *soft_dev, long long n)
    {
        const long long n_words = (n + 31) / 32;
        d_bits.ensure(sizeof(unsigned) * (size_t)(n_words + 4));
        XRD_CUDA(cudaMemsetAsync(d_bits.p, 0, sizeof(unsigned) * (size_t)(n_words + 4), stream));
        if (n_words > 0)
            XRD_LAUNCH(ctr, df_pack_kernel, (unsigned)((n_words + 255) / 256), 256, 0, stream, soft_dev, n, d_bits.as<unsigned>(),
                       n_words);
    }
};

extern "C" {

int xrd_decoder_front_create(int device, int lrit, int soft_mode, xrd_decoder_front **out)
{
    if (!out || soft_mode < 0 || soft_mode > 1) return XRD_E_ARG;
    *out = nullptr;
    int rc = select_device(device);
    if (rc) return rc;
    xrd_decoder_front *f = new xrd_decoder_front();
    f->device = device;
    f->lrit = lrit ? 1 : 0;
    f->soft_mode = soft_mode;
    // newdecoder.cpp:21-24,147-153: the encoded sync marker for 0 and 180 degrees
    f->words[0] = lrit ? 0xfca2b63db00d9794ull : 0xfc4ef4fd0cc2df89ull;
    f->words[1] = lrit ? 0x035d49c24ff2686bull : 0x25010b02f33d2076ull;
    rc = guarded([&]() {
        XRD_CUDA(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            f->d_last[i].ensure(DF_LAST);
            XRD_CUDA(cudaMemset(f->d_last[i].p, 128, DF_LAST));   // lastFrameEnd[i] = 128, newdecoder.cpp:141-145
        }
        f->d_scalars.ensure(64);
        f->d_key.ensure(sizeof(unsigned long long));
        XRD_CUDA(cudaFuncSetAttribute(df_viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)df_viterbi_smem()));
        return (int)XRD_OK;
    });
    if (rc) {
        delete f;
        return rc;
    }
    *out = f;
    return XRD_OK;
}

void xrd_decoder_front_destroy(xrd_decoder_front *f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    delete f;
}

int xrd_decoder_front_reset(xrd_decoder_front *f)
{
    if (!f) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(f->device));
        XRD_CUDA(cudaStreamSynchronize(f->stream));
        for (int i = 0; i < 2; i++) XRD_CUDA(cudaMemset(f->d_last[i].p, 128, DF_LAST));
        return (int)XRD_OK;
    });
}

int xrd_correlate(xrd_decoder_front *f, const uint8_t *data, uint32_t length, uint32_t *highest, uint32_t *position,
                  uint32_t *word)
{
    if (!f || (!data && length) || !highest || !position || !word) return XRD_E_ARG;
    *highest = *position = *word = 0;
    if (length <= 64) return XRD_OK;   // Correlator::correlate: nothing to search
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(f->device));
        f->d_soft.ensure(length);
        XRD_CUDA(cudaMemcpyAsync(f->d_soft.p, data, length, cudaMemcpyHostToDevice, f->stream));
        f->pack(f->d_soft.as<uint8_t>(), length);
        XRD_LAUNCH(f->ctr, df_correlate_kernel, 1, 1024, 0, f->stream, f->d_bits.as<unsigned>(), (long long)length - 64,
                   f->words[0], f->words[1], 2, f->d_key.as<unsigned long long>());
        unsigned long long k = 0;
        XRD_CUDA(cudaMemcpyAsync(&k, f->d_key.p, sizeof k, cudaMemcpyDeviceToHost, f->stream));
        XRD_CUDA(cudaStreamSynchronize(f->stream));
        // (all counts zero: the reference reports correlation 0 at position 0 with word 0)
        if ((k >> 41) > 0) {
            *highest = (uint32_t)(k >> 41);
            *position = (uint32_t)(0xFFFFFFFFFFull - ((k >> 1) & 0xFFFFFFFFFFull));
            *word = 1u - (uint32_t)(k & 1);
        }
        return (int)XRD_OK;
    });
}

int xrd_decoder_front_run(xrd_decoder_front *f, const int8_t *soft, size_t n, uint8_t *frames_out, xrd_frame_meta *meta_out,
                          size_t cap, size_t *n_frames, size_t *consumed)
{
    if (!f || (!soft && n) || !n_frames || !consumed || (cap && (!frames_out || !meta_out))) return XRD_E_ARG;
    *n_frames = 0;
    *consumed = 0;
    if (n < (size_t)DF_FRAME || cap == 0) return XRD_OK;
    if (n >= (1ull << 38) || cap > 0x7fffffffull) return XRD_E_ARG;
    return guarded([&]() {
        XRD_CUDA(cudaSetDevice(f->device));
        const long long N = (long long)n;
        f->d_soft.ensure(n);
        XRD_CUDA(cudaMemcpyAsync(f->d_soft.p, soft, n, cudaMemcpyHostToDevice, f->stream));
        f->pack(f->d_soft.as<uint8_t>(), N);
        const long long n_pos = N - 64;
        const long long n_blk = (n_pos + DF_BLK - 1) / DF_BLK;
        f->d_blockmax.ensure(sizeof(unsigned long long) * (size_t)n_blk);
        XRD_LAUNCH(f->ctr, df_blockmax_kernel, (unsigned)n_blk, DF_BLK, 0, f->stream, f->d_bits.as<unsigned>(), n_pos, f->words[0],
                   f->words[1], f->d_blockmax.as<unsigned long long>());
        const int max_frames = (int)std::min<size_t>(cap, n / DF_FRAME);
        f->d_frames.ensure(sizeof(DfFrame) * (size_t)max_frames);
        int *d_nf = f->d_scalars.as<int>();
        long long *d_cons = reinterpret_cast<long long *>(f->d_scalars.as<char>() + 8);
        XRD_LAUNCH(f->ctr, df_walk_kernel, 1, 32, 0, f->stream, f->d_bits.as<unsigned>(), f->d_blockmax.as<unsigned long long>(), N,
                   f->words[0], f->words[1], f->d_frames.as<DfFrame>(), max_frames, d_nf, d_cons);
        int nf = 0;
        long long cons = 0;
        XRD_CUDA(cudaMemcpyAsync(&nf, d_nf, sizeof nf, cudaMemcpyDeviceToHost, f->stream));
        XRD_CUDA(cudaMemcpyAsync(&cons, d_cons, sizeof cons, cudaMemcpyDeviceToHost, f->stream));
        XRD_CUDA(cudaStreamSynchronize(f->stream));
        *consumed = (size_t)cons;
        if (nf <= 0) return (int)XRD_OK;
        f->d_out.ensure((size_t)nf * (DF_BITS / 8));
        f->d_err.ensure(sizeof(int) * (size_t)nf);
        XRD_LAUNCH(f->ctr, df_viterbi_kernel, nf, 32, df_viterbi_smem(), f->stream, f->d_soft.as<uint8_t>(), f->d_frames.as<DfFrame>(),
                   nf, f->lrit, f->soft_mode, f->d_last[f->cur].as<uint8_t>(), f->d_last[f->cur ^ 1].as<uint8_t>(), f->d_out.as<uint8_t>(),
                   f->d_err.as<int>());
        f->cur ^= 1;
        std::vector<DfFrame> fr((size_t)nf);
        std::vector<int> err((size_t)nf);
        XRD_CUDA(cudaMemcpyAsync(frames_out, f->d_out.p, (size_t)nf * (DF_BITS / 8), cudaMemcpyDeviceToHost, f->stream));
        XRD_CUDA(cudaMemcpyAsync(fr.data(), f->d_frames.p, sizeof(DfFrame) * (size_t)nf, cudaMemcpyDeviceToHost, f->stream));
        XRD_CUDA(cudaMemcpyAsync(err.data(), f->d_err.p, sizeof(int) * (size_t)nf, cudaMemcpyDeviceToHost, f->stream));
        XRD_CUDA(cudaStreamSynchronize(f->stream));
        for (int i = 0; i < nf; i++) {
            meta_out[i].offset = fr[i].offset;
            meta_out[i].correlation = fr[i].corr;
            meta_out[i].word = fr[i].word;
            meta_out[i].bit_errors = err[i];
        }
        *n_frames = (size_t)nf;
        return (int)XRD_OK;
    });
}

}  // extern "C"
