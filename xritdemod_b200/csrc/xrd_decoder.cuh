// xrd_decoder.cuh -- the decoder's front half on the soft-symbol byte stream this path emits
// (reference decoder/src/newdecoder.cpp:212-290): sync-word correlation (SatHelper::Correlator, :218-247), frame
// alignment (:250-264), 180 degree phase fix (PacketFixer, :268-270), r = 1/2 k = 7 Viterbi with the 64 soft bytes of the
// previous frame in front (Viterbi27, :273-296), NRZ-M decoding for HRIT (:283-285).  SURVEY.md 8(f) row 3.
//
// All integer / byte work, bit-exact against the CPU restatement under oracle/ (its xo_decoder_front).  The convolutional code itself is
// pinned by the reference: its four sync-word constants (newdecoder.cpp:21-24) are the encoded attached sync marker, and
// only polynomials 0x4F, 0x6D on a register fed at the LSB with coded 0 = positive symbol reproduce them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xrd {

constexpr int DF_FRAME = 16384;     // CODEDFRAMESIZE, decoder/src/parameters.h:30
constexpr int DF_BITS = 8192;       // FRAMEBITS
constexpr int DF_LAST = 64;         // LASTFRAMEDATABITS: soft bytes of the previous frame decoded in front of every frame
constexpr int DF_MINCORR = 46;      // MINCORRELATIONBITS
constexpr int DF_SEARCH = DF_FRAME - 64;   // positions Correlator::correlate looks at in one chunk
constexpr int DF_BLK = 256;         // positions per block maximum

struct DfFrame {
    long long offset;   // of the frame's first soft byte in the stream
    int corr, word;
};

// hard decision as the correlator takes it (byte < 127 reads as word bit 1), 32 positions per word, first position in
// the MSB
__global__ void df_pack_kernel(const uint8_t *__restrict__ soft, long long n, unsigned *__restrict__ bits, long long n_words)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    unsigned v = 0;
    const long long p0 = w * 32;
    if (p0 + 32 <= n && ((reinterpret_cast<unsigned long long>(soft + p0) & 15) == 0)) {
        const uint4 a = *reinterpret_cast<const uint4 *>(soft + p0), b = *reinterpret_cast<const uint4 *>(soft + p0 + 16);
        const unsigned q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) v |= (unsigned)(((q[k] >> (8 * j)) & 0xff) < 127u) << (31 - (4 * k + j));
    } else {
        for (int i = 0; i < 32; i++)
            if (p0 + i < n) v |= (unsigned)(soft[p0 + i] < 127u) << (31 - i);
    }
    bits[w] = v;
}

// agreement counts of the 64 symbols at position p with the two sync words; key orders candidates as
// Correlator::correlate does: highest count, then lowest position, then word 0 before word 1
__device__ __forceinline__ unsigned long long df_key(const unsigned *__restrict__ bits, long long p, unsigned long long w0,
                                                     unsigned long long w1)
{
    const long long w = p >> 5;
    const int s = (int)(p & 31);
    const unsigned b0 = bits[w], b1 = bits[w + 1], b2 = bits[w + 2];
    const unsigned hi = __funnelshift_l(b1, b0, s), lo = __funnelshift_l(b2, b1, s);
    const unsigned long long win = ((unsigned long long)hi << 32) | lo;
    const int c0 = 64 - __popcll(win ^ w0), c1 = 64 - __popcll(win ^ w1);
    const int c = c1 > c0 ? c1 : c0;
    const unsigned word = c1 > c0 ? 1u : 0u;
    // [count : 8][~position : 40][~word : 1] -- larger is better
    return ((unsigned long long)c << 41) | ((0xFFFFFFFFFFull - (unsigned long long)p) << 1) | (1u - word);
}
__device__ __forceinline__ void df_unkey(unsigned long long k, int &corr, long long &pos, int &word)
{
    corr = (int)(k >> 41);
    pos = (long long)(0xFFFFFFFFFFull - ((k >> 1) & 0xFFFFFFFFFFull));
    word = 1 - (int)(k & 1);
}

// best candidate of every block of DF_BLK positions (positions >= n_pos do not exist)
__global__ void __launch_bounds__(DF_BLK)
df_blockmax_kernel(const unsigned *__restrict__ bits, long long n_pos, unsigned long long w0, unsigned long long w1,
                   unsigned long long *__restrict__ blockmax)
{
    __shared__ unsigned long long s_k[DF_BLK / 32];
    const long long p = (long long)blockIdx.x * DF_BLK + threadIdx.x;
    unsigned long long k = (p < n_pos) ? df_key(bits, p, w0, w1) : 0ull;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long q = __shfl_xor_sync(0xffffffffu, k, o);
        k = q > k ? q : k;
    }
    if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = s_k[0];
        for (int i = 1; i < DF_BLK / 32; i++) m = s_k[i] > m ? s_k[i] : m;
        blockmax[blockIdx.x] = m;
    }
}

// The frame walk of newdecoder.cpp:212-264 (one warp; every step depends on the one before): take the 16384 bytes at
// s, find the best sync candidate among its first DF_SEARCH positions, drop the chunk when the count is below
// MINCORRELATIONBITS, else the frame starts at that position and the next chunk right behind it.  Stops when the
// stream cannot hold the next chunk or frame.
__global__ void __launch_bounds__(32)
df_walk_kernel(const unsigned *__restrict__ bits, const unsigned long long *__restrict__ blockmax, long long n,
               unsigned long long w0, unsigned long long w1, DfFrame *__restrict__ frames, int cap, int *__restrict__ n_frames,
               long long *__restrict__ consumed)
{
    const int lane = threadIdx.x;
    long long s = 0;
    int nf = 0;
    while (s + DF_FRAME <= n && nf < cap) {
        const long long lo = s, hi = s + DF_SEARCH;            // candidates [lo, hi)
        const long long b_lo = (lo + DF_BLK - 1) / DF_BLK, b_hi = hi / DF_BLK;   // whole blocks [b_lo, b_hi)
        unsigned long long k = 0;
        if (b_lo <= b_hi) {
            for (long long b = b_lo + lane; b < b_hi; b += 32) {
                const unsigned long long q = blockmax[b];
                k = q > k ? q : k;
            }
            for (long long p = lo + lane; p < b_lo * DF_BLK; p += 32) {
                const unsigned long long q = df_key(bits, p, w0, w1);
                k = q > k ? q : k;
            }
            for (long long p = b_hi * DF_BLK + lane; p < hi; p += 32) {
                const unsigned long long q = df_key(bits, p, w0, w1);
                k = q > k ? q : k;
            }
        } else {
            for (long long p = lo + lane; p < hi; p += 32) {
                const unsigned long long q = df_key(bits, p, w0, w1);
                k = q > k ? q : k;
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long q = __shfl_xor_sync(0xffffffffu, k, o);
            k = q > k ? q : k;
        }
        int corr, word;
        long long pos;
        df_unkey(k, corr, pos, word);
        if (corr < DF_MINCORR) {
            s += DF_FRAME;
            continue;
        }
        if (pos + DF_FRAME > n) break;   // the rest of the frame has not arrived yet
        if (lane == 0) {
            frames[nf].offset = pos;
            frames[nf].corr = corr;
            frames[nf].word = word;
        }
        nf++;
        s = pos + DF_FRAME;
    }
    if (lane == 0) {
        *n_frames = nf;
        *consumed = s;
    }
}

// Correlator::correlate on one buffer: best candidate among the first length - 64 positions
__global__ void __launch_bounds__(1024)
df_correlate_kernel(const unsigned *__restrict__ bits, long long n_pos, unsigned long long w0, unsigned long long w1,
                    int n_words, unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long s_k[32];
    unsigned long long k = 0;
    for (long long p = threadIdx.x; p < n_pos; p += blockDim.x) {
        const unsigned long long q = df_key(bits, p, w0, n_words > 1 ? w1 : w0);
        k = q > k ? q : k;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long q = __shfl_xor_sync(0xffffffffu, k, o);
        k = q > k ? q : k;
    }
    if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) m = s_k[i] > m ? s_k[i] : m;
        *out = m;
    }
}

// ---------------------------------------------------------------------------------------
// Viterbi27::decode, one warp per frame.  Lane l holds the path metrics of states l and l + 32 (state = the last six
// information bits, newest at the LSB): both predecessors of the new states 2l and 2l + 1, so the add-compare-select is
// lane-local, and two shuffles put the new metrics back in place.  Metrics are 16-bit (re-based on the smallest one every
// 32 steps: the spread never exceeds 6 * 510), decisions go to shared memory as two ballots per step, one thread traces
// back.  Metric |u - 255 c| on the soft byte u (raw, or 127 - v: soft_mode), ties keep the predecessor with the older bit 0, the best final state is
// traced (lowest on ties) -- the oracle's rules, so the output is identical.
// ---------------------------------------------------------------------------------------
constexpr int DF_VBITS = DF_BITS + DF_LAST / 2;          // 8224 decoded bits per frame
constexpr int DF_VBYTES = DF_FRAME + DF_LAST;            // 16448 soft bytes
__host__ __device__ inline size_t df_viterbi_smem() { return (size_t)DF_VBITS * 8 + DF_VBYTES + DF_VBITS / 8 + 16; }

__device__ __forceinline__ int df_parity(unsigned v) { return __popc(v) & 1; }

__global__ void __launch_bounds__(32)
df_viterbi_kernel(const uint8_t *__restrict__ soft, const DfFrame *__restrict__ frames, int n_frames, int lrit, int soft_mode,
                  const uint8_t *__restrict__ last_end_in /* 64 bytes carried from the call before */,
                  uint8_t *__restrict__ last_end_out, uint8_t *__restrict__ out /* 1024 bytes per frame */,
                  int *__restrict__ bit_errors)
{
    extern __shared__ __align__(16) unsigned char df_smem[];
    uint2 *s_dec = reinterpret_cast<uint2 *>(df_smem);                     // [DF_VBITS] ballots (even, odd new states)
    uint8_t *s_y = df_smem + (size_t)DF_VBITS * 8;                         // [DF_VBYTES] soft bytes, phase fixed
    uint8_t *s_out = s_y + DF_VBYTES;                                      // [DF_VBITS / 8] decoded bytes
    const int f = blockIdx.x, lane = threadIdx.x;
    if (f >= n_frames) return;
    const DfFrame fr = frames[f];
    const uint8_t fix = (lrit && fr.word == 1) ? 0xFF : 0x00;             // PacketFixer, DEG_180: every byte inverted
    // the 64 soft bytes in front: the end of the previous frame as IT was decoded (its own phase fix applied)
    if (f == 0) {
        for (int i = lane; i < DF_LAST; i += 32) s_y[i] = last_end_in[i];
    } else {
        const DfFrame pr = frames[f - 1];
        const uint8_t pfix = (lrit && pr.word == 1) ? 0xFF : 0x00;
        for (int i = lane; i < DF_LAST; i += 32) s_y[i] = soft[pr.offset + DF_FRAME - DF_LAST + i] ^ pfix;
    }
    for (int i = lane; i < DF_FRAME; i += 32) s_y[DF_LAST + i] = soft[fr.offset + i] ^ fix;
    __syncwarp();
    if (f == n_frames - 1)
        for (int i = lane; i < DF_LAST; i += 32) last_end_out[i] = s_y[DF_FRAME + i];

    // coded bits of the four transitions this lane evaluates: register r = (old << 1) | bit, old = l or l + 32
    unsigned oa = 0, ob = 0;   // bit (2 * hi + bit): poly 0x4F / 0x6D output
#pragma unroll
    for (int hi = 0; hi < 2; hi++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const unsigned r = (((unsigned)lane + 32u * hi) << 1) | (unsigned)b;
            oa |= (unsigned)df_parity(r & 0x4F) << (2 * hi + b);
            ob |= (unsigned)df_parity(r & 0x6D) << (2 * hi + b);
        }
    unsigned m_lo = 0, m_hi = 0;   // metrics of states lane, lane + 32
    for (int t = 0; t < DF_VBITS; t++) {
        // the byte as the decoder library reads it: raw (soft_mode 0, the reference call chain), or 127 - v for the signed
        // symbol v (soft_mode 1) -- see the restatement of Viterbi27::decode under oracle/
        const unsigned y0 = soft_mode ? ((127u - s_y[2 * t]) & 0xFFu) : s_y[2 * t];
        const unsigned y1 = soft_mode ? ((127u - s_y[2 * t + 1]) & 0xFFu) : s_y[2 * t + 1];
        const unsigned c0[2] = {y0, 255u - y0}, c1[2] = {y1, 255u - y1};
        // new state 2l + b from old l (a) or old l + 32 (c)
        unsigned nv[2], dbit[2];
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const unsigned a = m_lo + c0[(oa >> b) & 1] + c1[(ob >> b) & 1];
            const unsigned c = m_hi + c0[(oa >> (2 + b)) & 1] + c1[(ob >> (2 + b)) & 1];
            dbit[b] = a <= c ? 0u : 1u;
            nv[b] = a <= c ? a : c;
        }
        const unsigned d_even = __ballot_sync(0xffffffffu, dbit[0]);   // bit l: decision of new state 2l
        const unsigned d_odd = __ballot_sync(0xffffffffu, dbit[1]);    // bit l: decision of new state 2l + 1
        if (lane == 0) s_dec[t] = make_uint2(d_even, d_odd);
        // new state j (and j + 32) lives in lane j: it was produced by lane j >> 1 (lane 16 + (j >> 1)) as its b = j & 1
        const unsigned packed = nv[0] | (nv[1] << 16);
        const unsigned from_lo = __shfl_sync(0xffffffffu, packed, lane >> 1);
        const unsigned from_hi = __shfl_sync(0xffffffffu, packed, 16 + (lane >> 1));
        m_lo = (lane & 1) ? (from_lo >> 16) : (from_lo & 0xffff);
        m_hi = (lane & 1) ? (from_hi >> 16) : (from_hi & 0xffff);
        if ((t & 31) == 31) {
            const unsigned mn = __reduce_min_sync(0xffffffffu, m_lo < m_hi ? m_lo : m_hi);
            m_lo -= mn;
            m_hi -= mn;
        }
    }
    // best final state, lowest on ties
    const unsigned klo = (m_lo << 6) | (unsigned)lane, khi = (m_hi << 6) | (unsigned)(lane + 32);
    unsigned s = __reduce_min_sync(0xffffffffu, klo < khi ? klo : khi) & 63u;
    __syncwarp();
    for (int i = lane; i < DF_VBITS / 8; i += 32) s_out[i] = 0;
    __syncwarp();
    if (lane == 0) {
        for (int t = DF_VBITS - 1; t >= 0; t--) {
            const unsigned bit = s & 1u;
            if (bit) s_out[t >> 3] |= (uint8_t)(0x80 >> (t & 7));
            const uint2 d = s_dec[t];
            const unsigned dec = ((bit ? d.y : d.x) >> (s >> 1)) & 1u;
            s = (s >> 1) | (dec << 5);
        }
        // Viterbi27::GetBER: re-encode from the state the trace-back ended in, count disagreements with the hard bits
        unsigned sr = s;
        int err = 0;
        for (int t = 0; t < DF_VBITS; t++) {
            const unsigned bit = (s_out[t >> 3] >> (7 - (t & 7))) & 1u;
            sr = ((sr << 1) | bit) & 0x7F;
            err += (df_parity(sr & 0x4F) != (int)(s_y[2 * t] >> 7)) + (df_parity(sr & 0x6D) != (int)(s_y[2 * t + 1] >> 7));
        }
        bit_errors[f] = err;
        if (!lrit) {
            // DifferentialEncoding::nrzmDecode over the decoded bytes (newdecoder.cpp:283-285)
            uint8_t last = 0;
            for (int i = 0; i < DF_VBITS / 8; i++) {
                const uint8_t v = s_out[i];
                const uint8_t mask = (uint8_t)(((v >> 1) & 0x7F) | (last << 7));
                last = v & 1;
                s_out[i] = v ^ mask;
            }
        }
    }
    __syncwarp();
    // drop the 4 warm-up bytes (newdecoder.cpp:293)
    for (int i = lane; i < DF_BITS / 8; i += 32) out[(size_t)f * (DF_BITS / 8) + i] = s_out[DF_LAST / 16 + i];
}

}  // namespace xrd
