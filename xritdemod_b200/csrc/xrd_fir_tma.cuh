// xrd_fir_tma.cuh -- the stride-1 FIR (FirFilter::Work with decimation 1: the RRC matched filter,
// reference demodulator.cpp:148,450) as a persistent kernel whose input tiles are staged by the
// TMA unit (cp.async.bulk.tensor, UTMALDG in SASS) into a double-buffered shared-memory ring
// guarded by mbarriers.
//
// The sample buffer is described to the TMA as a 2-D tensor of floats, 256 floats (128 cf32
// samples, 1 KB) per row; a tile of FT_TILE outputs needs the FT_TILE + ntaps - 1 samples that
// end at its last output, i.e. at most ft_rows(ntaps) whole rows starting at the row that holds
// its first history sample.  A producer warp arms the stage's `full` mbarrier with the byte count
// and issues ONE bulk tensor copy per tile, up to three tiles ahead; eight consumer warps run the
// same register sliding window as fir1_kernel out of shared memory (thread t owns 9 consecutive
// outputs; I and Q accumulate in one packed FFMA2, taps k = 0..T-1 by fmaf from zero -- the
// oracle's order, so the result is bit-identical) and hand the stage back through its `empty`
// mbarrier warp by warp: there is no CTA-wide barrier in the loop.  Taps are a kernel parameter
// (constant bank), already duplicated into the (h, h) pairs FFMA2 wants.  Rows past the end of the
// allocation are zero-filled by the TMA and only ever feed outputs that are not stored.
//
// Persistent: grid = CTAs resident on the device; CTA b takes tiles b, b + grid, b + 2 grid, ...
// (of all channels).
#pragma once
#include <cuda.h>

#include "xrd_kernels.cuh"

namespace xrd {

constexpr int FT_WARPS = 8;                  // consumer warps per CTA
constexpr int FT_CONSUMERS = FT_WARPS * 32;
constexpr int FT_THREADS = FT_CONSUMERS + 32; // + one producer warp (one lane issues the copies)
constexpr int FT_R = 9;                      // outputs per thread (odd: conflict-free 8-byte shared loads)
constexpr int FT_TILE = FT_CONSUMERS * FT_R; // outputs per tile
constexpr int FT_ROW = 128;                  // samples per tensor row (1 KB)
#ifndef XRD_FT_STAGES
#define XRD_FT_STAGES 2
#endif
constexpr int FT_STAGES = XRD_FT_STAGES;
constexpr int FT_HEAD = 128;                 // bytes before the first stage: the mbarriers, and slot -1 of stage 0
__host__ __device__ inline int ft_rows(int ntaps)
{
    // a tile starts up to FT_ROW - 1 samples into its first row and spans FT_TILE + ntaps - 1 samples
    return (FT_ROW - 1 + FT_TILE + ntaps - 1 + FT_ROW - 1) / FT_ROW;
}
constexpr int FT_WTILE = 32 * FT_R;          // outputs per consumer warp and tile (2304 bytes)
__host__ __device__ inline size_t ft_smem_bytes(int ntaps)
{
    // mbarriers | input ring | one output staging slab per consumer warp
    return FT_HEAD + (size_t)FT_STAGES * ft_rows(ntaps) * FT_ROW * sizeof(float2) + (size_t)FT_WARPS * FT_WTILE * sizeof(float2);
}

__device__ __forceinline__ void ft_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ft_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void ft_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    // bounded: a copy that never completes (a bad tensor map) must end in a launch failure, not in a hung device
    for (unsigned spins = 0;; spins++) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();
    }
}
// one bulk tensor copy: the box of `tmap` at (column 0, row) -> shared memory, completion counted on `bar`
__device__ __forceinline__ void ft_tma_load_rows(void *smem_dst, const CUtensorMap *tmap, int row, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(0), "r"(row), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

// Warp-specialised: the producer warp runs FT_STAGES tiles ahead (it waits for a stage's `empty` barrier, which the
// eight consumer warps arrive on one by one as they finish with it, then arms `full` and issues the copy); a consumer
// warp waits for `full`, computes its 288 outputs of the tile and moves on -- no CTA-wide barrier in the loop.
__global__ void __launch_bounds__(FT_THREADS)
fir_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FtTaps taps, float2 *__restrict__ out, int ntaps,
               long long n_out, long long in_off /* samples from the tensor base to x[0] of channel 0 */,
               long long in_ch_stride, long long out_ch_stride, int tiles_per_ch, int n_tiles)
{
    extern __shared__ __align__(128) unsigned char ft_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(ft_smem);          // [FT_STAGES]
    uint64_t *empty = full + FT_STAGES;                              // [FT_STAGES]
    const int H = ntaps - 1;
    const int rows = ft_rows(ntaps);
    const int stage_elems = rows * FT_ROW;
    float2 *buf0 = reinterpret_cast<float2 *>(ft_smem + FT_HEAD);
    float2 *s_out = buf0 + (size_t)FT_STAGES * stage_elems;   // [FT_WARPS][FT_WTILE]: outputs on their way to HBM
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < FT_STAGES; s++) {
            ft_mbar_init(&full[s], 1);
            ft_mbar_init(&empty[s], FT_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // first sample a tile needs, relative to the tensor base
    auto tile_origin = [&](int tile, int &ch, long long &tile0) -> long long {
        ch = tile / tiles_per_ch;
        tile0 = (long long)(tile - ch * tiles_per_ch) * FT_TILE;
        return in_off + (long long)ch * in_ch_stride + tile0 - H;
    };

    if (tid >= FT_CONSUMERS) {
        // ---- producer warp
        if (tid == FT_CONSUMERS) {
            for (int it = 0;; it++) {
                const int tile = blockIdx.x + it * gridDim.x;
                if (tile >= n_tiles) break;
                const int s = it % FT_STAGES, use = it / FT_STAGES;
                if (use > 0) ft_mbar_wait(&empty[s], (unsigned)((use - 1) & 1));   // every consumer warp is done with it
                int ch;
                long long tile0;
                const long long g0 = tile_origin(tile, ch, tile0);
                ft_mbar_expect_tx(&full[s], (unsigned)(stage_elems * sizeof(float2)));
                ft_tma_load_rows(buf0 + (size_t)s * stage_elems, &tmap, (int)(g0 / FT_ROW), &full[s]);
            }
        }
        return;
    }

    // ---- consumer warps
    const int o0 = tid * FT_R;   // first output of this thread (tile-relative)
    const int lane = tid & 31, wid = tid >> 5;
    float2 *stg = s_out + (size_t)wid * FT_WTILE;   // this warp's staging slab
    for (int it = 0;; it++) {
        const int tile = blockIdx.x + it * gridDim.x;
        if (tile >= n_tiles) break;
        const int s = it % FT_STAGES, use = it / FT_STAGES;
        int ch;
        long long tile0;
        const long long g0 = tile_origin(tile, ch, tile0);
        const int shift = (int)(g0 % FT_ROW);
        const int tile_n = (int)min((long long)FT_TILE, n_out - tile0);
        ft_mbar_wait(&full[s], (unsigned)(use & 1));
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the slab's last copy has read it
        __syncwarp();
        float2 acc[FT_R];
#pragma unroll
        for (int r = 0; r < FT_R; r++) acc[r] = make_float2(0.f, 0.f);
        {
            // every lane computes, whether its outputs exist or not (the stage holds the samples either way; what does
            // not exist is not stored): the loop below then sits in warp-uniform control flow and the taps stay in uniform
            // registers.  s_x[i] = x[tile0 - H + i]
            const float2 *s_x = buf0 + (size_t)s * stage_elems + shift;
            float2 w[FT_R];
#pragma unroll
            for (int r = 0; r < FT_R; r++) {
                const int m = o0 + r;
                w[r] = s_x[m + H];
            }
            const float2 *xb = s_x + o0 + H;   // xb[-k-1] = next sample entering the window
            int kb = 0;
            // (the trip count goes through a shuffle so that the compiler KNOWS it is warp-uniform: it then keeps the loop
            // counter and the taps in uniform registers and issues FFMA2 with a uniform operand -- with three vector
            // operands the same loop ran at 50 instead of 64 TFLOP/s)
            const int nfull = __shfl_sync(0xffffffffu, ntaps / FT_R, 0);
            for (int ib = 0; ib < nfull; ib++, kb += FT_R) {
#pragma unroll
                for (int q = 0; q < FT_R; q++) {
                    const float2 h2 = taps.h2[kb + q];
#pragma unroll
                    for (int r = 0; r < FT_R; r++) {
                        const int sl = (r - q + FT_R) % FT_R;
                        acc[r] = __ffma2_rn(h2, w[sl], acc[r]);   // two IEEE fmaf in one FFMA2 (I and Q share the tap)
                    }
                    // relative index -(k+1) enters the slot that (R-1-k) leaves (the last one reads slot -1: unused)
                    w[(FT_R - 1 - q) % FT_R] = xb[-(kb + q) - 1];
                }
            }
#pragma unroll
            for (int q = 0; q < FT_R; q++) {
                if (kb + q < ntaps) {
                    const float2 h2 = taps.h2[kb + q];
#pragma unroll
                    for (int r = 0; r < FT_R; r++) {
                        const int sl = (r - q + FT_R) % FT_R;
                        acc[r] = __ffma2_rn(h2, w[sl], acc[r]);
                    }
                    w[(FT_R - 1 - q) % FT_R] = xb[-(kb + q) - 1];
                }
            }
            // Outputs leave through shared memory: a thread's 9 consecutive outputs are 72 bytes, so storing them
            // directly costs 32 partial sectors per store instruction (the kernel was L1-bound on that at <= 31 taps).
            // Staged, a full warp slab (288 outputs, 2304 contiguous bytes) goes out as ONE bulk copy issued by lane 0
            // (cp.async.bulk shared -> global); ragged or unaligned slabs take coalesced 8-byte stores.
#pragma unroll
            for (int r = 0; r < FT_R; r++) stg[lane * FT_R + r] = acc[r];
        }
        {
            float2 *o = out + (size_t)ch * out_ch_stride + tile0 + (size_t)wid * FT_WTILE;
            const int wn = tile_n - wid * FT_WTILE;   // outputs of this warp that exist (warp-uniform)
            if (wn >= FT_WTILE && ((reinterpret_cast<unsigned long long>(o) & 15) == 0)) {
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(o),
                                 "r"((unsigned)__cvta_generic_to_shared(stg)), "n"(FT_WTILE * (int)sizeof(float2))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                }
            } else if (wn > 0) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < FT_R; i++) {
                    const int idx = lane + 32 * i;
                    if (idx < wn) o[idx] = stg[idx];
                }
            }
        }
        __syncwarp();
        if (lane == 0) ft_mbar_arrive(&empty[s]);   // this warp has read everything it needs from the stage
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // before the slab's memory goes away
}

// ---- host side: the tensor map of a sample buffer (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so that libxrd.so does not link libcuda) ----
typedef CUresult (*ft_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline ft_encode_fn ft_encoder()
{
    static ft_encode_fn fn = []() -> ft_encode_fn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<ft_encode_fn>(p);
    }();
    return fn;
}

// base: start of the allocation (cudaMalloc: 256-byte aligned), bytes: its size.  false: no TMA path (caller falls back)
inline bool ft_make_map(CUtensorMap *map, void *base, size_t bytes, int ntaps)
{
    ft_encode_fn enc = ft_encoder();
    if (!enc || !base || (reinterpret_cast<unsigned long long>(base) & 15)) return false;
    const cuuint64_t rows = (cuuint64_t)(bytes / (FT_ROW * sizeof(float2)));   // whole rows only
    if (rows == 0 || rows > 0xffffffffull) return false;
    const cuuint64_t dims[2] = {2 * FT_ROW, rows};
    const cuuint64_t strides[1] = {FT_ROW * sizeof(float2)};
    const cuuint32_t box[2] = {2 * FT_ROW, (cuuint32_t)ft_rows(ntaps)};
    const cuuint32_t estr[2] = {1, 1};
    if (box[1] > 256) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace xrd
