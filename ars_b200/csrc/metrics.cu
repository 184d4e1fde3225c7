// Result metrics (calculate_audio_metrics, rs.py:674-711): sample peak and RMS come from the
// running sums of the epilogue; this file adds the K-weighted, gated integrated loudness that
// the reference obtains from pyloudnorm (restated in SURVEY.md App. B -- PARITY UNPINNED, the
// package is not installable here).
//
// The two K-weighting biquads are IIR recurrences over the whole signal.  They are evaluated
// as a parallel scan of affine maps in float64: every thread filters a 32-sample chunk from a
// zero state, the chunk end-states are combined with precomputed powers of the 2x2 state
// matrix (block scan, then a short serial scan over block aggregates), and a second sweep
// re-filters every chunk from its true start state.  As in pyloudnorm, the stage output is
// rounded to float32 before the next stage reads it.
#include "metrics.cuh"
#include "final_math.cuh"

#include <cmath>
#include <math_constants.h>

namespace ars {

constexpr int CH = 32;                // samples per thread
constexpr int NTB = 256;              // threads per block
constexpr int BS = CH * NTB;          // samples per block

struct Biquad {                       // normalised (a0 = 1)
    double b0, b1, b2, a1, a2;
};
struct Mat2 { double m00, m01, m10, m11; };
struct ScanCoef {
    Biquad q;
    Mat2 pw[9];                       // A^(CH * 2^k), k = 0..8  (pw[8] = A^BS)
    double2 ends[CH];                 // A^(CH-1-j) B: weight of sample j of a chunk in the chunk's zero-start end state
};

static Mat2 mat_mul(const Mat2& a, const Mat2& b) {
    return {a.m00 * b.m00 + a.m01 * b.m10, a.m00 * b.m01 + a.m01 * b.m11,
            a.m10 * b.m00 + a.m11 * b.m10, a.m10 * b.m01 + a.m11 * b.m11};
}

static ScanCoef make_coef(const Biquad& q) {
    ScanCoef c;
    c.q = q;
    Mat2 a = {-q.a1, 1.0, -q.a2, 0.0};      // direct form II transposed state matrix
    Mat2 p = a;
    for (int i = 1; i < CH; i <<= 1) p = mat_mul(p, p);     // A^CH (CH is a power of two)
    c.pw[0] = p;
    for (int k = 1; k < 9; ++k) c.pw[k] = mat_mul(c.pw[k - 1], c.pw[k - 1]);
    // z' = A z + B x with B = (b1 - a1 b0, b2 - a2 b0): the state after a chunk from a zero state is linear in its samples
    double2 v = make_double2(q.b1 - q.a1 * q.b0, q.b2 - q.a2 * q.b0);
    for (int j = CH - 1; j >= 0; --j) {
        c.ends[j] = v;
        v = make_double2(a.m00 * v.x + a.m01 * v.y, a.m10 * v.x + a.m11 * v.y);
    }
    return c;
}

__device__ __forceinline__ double2 mat_vec(const Mat2& m, double2 v) {
    return make_double2(m.m00 * v.x + m.m01 * v.y, m.m10 * v.x + m.m11 * v.y);
}

// scipy.signal.lfilter's recurrence (direct form II transposed)
__device__ __forceinline__ double df2t(const Biquad& q, double x, double2& z) {
    const double y = q.b0 * x + z.x;
    z.x = q.b1 * x - q.a1 * y + z.y;
    z.y = q.b2 * x - q.a2 * y;
    return y;
}

// PHASE 0: per-block aggregate end state (zero start).  PHASE 1: real output, block start states given.
template <int PHASE>
__global__ void __launch_bounds__(NTB) biquad_kernel(const float* __restrict__ x, i64 N, ScanCoef cf,
                                                     double2* __restrict__ block_state, float* __restrict__ y) {
    __shared__ float sx[NTB * (CH + 1)];
    __shared__ double2 sv[NTB];
    const i64 base = (i64)blockIdx.x * BS;
    const int t = threadIdx.x;
    for (int i = t; i < BS; i += NTB) {
        const i64 g = base + i;
        sx[(i / CH) * (CH + 1) + (i % CH)] = g < N ? x[g] : 0.f;
    }
    __syncthreads();
    float* mine = sx + t * (CH + 1);
    double2 z = make_double2(0.0, 0.0);
    #pragma unroll 8
    for (int j = 0; j < CH; ++j) df2t(cf.q, (double)mine[j], z);
    // inclusive scan of w_t = P w_{t-1} + g_t with P = A^CH; the block's start state enters at t = 0
    double2 v = z;
    if (PHASE == 1 && t == 0) {
        const double2 s0 = block_state[blockIdx.x];
        const double2 ps = mat_vec(cf.pw[0], s0);
        v.x += ps.x; v.y += ps.y;
    }
    sv[t] = v;
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int off = 1 << k;
        double2 add = make_double2(0.0, 0.0);
        if (t >= off) add = mat_vec(cf.pw[k], sv[t - off]);
        __syncthreads();
        v.x += add.x; v.y += add.y;
        sv[t] = v;
        __syncthreads();
    }
    if (PHASE == 0) {
        if (t == NTB - 1) block_state[blockIdx.x] = v;       // aggregate: end state from a zero start
        return;
    }
    double2 s = (t == 0) ? block_state[blockIdx.x] : sv[t - 1];
    #pragma unroll 8
    for (int j = 0; j < CH; ++j) mine[j] = (float)df2t(cf.q, (double)mine[j], s);   // float32 store, as pyloudnorm
    __syncthreads();
    for (int i = t; i < BS; i += NTB) {
        const i64 g = base + i;
        if (g < N) y[g] = sx[(i / CH) * (CH + 1) + (i % CH)];
    }
}

// exclusive scan over block aggregates: start[b+1] = A^BS start[b] + agg[b].  One CTA of 1024 threads,
// Hillis-Steele over tiles of 1024 aggregates with the powers (A^BS)^(2^k) and a running carry.
struct BlockPow { Mat2 p[10]; };
__global__ void __launch_bounds__(1024) block_scan_kernel(double2* state, int nblocks, BlockPow bp) {
    __shared__ double2 sv[1024];
    __shared__ double2 carry;
    const int t = threadIdx.x;
    if (t == 0) carry = make_double2(0.0, 0.0);
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int b = base + t;
        double2 v = b < nblocks ? state[b] : make_double2(0.0, 0.0);
        if (t == 0) {
            const double2 pc = mat_vec(bp.p[0], carry);
            v.x += pc.x; v.y += pc.y;
        }
        sv[t] = v;
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < 10; ++k) {
            const int off = 1 << k;
            double2 add = make_double2(0.0, 0.0);
            if (t >= off) add = mat_vec(bp.p[k], sv[t - off]);
            __syncthreads();
            v.x += add.x; v.y += add.y;
            sv[t] = v;
            __syncthreads();
        }
        const double2 start = (t == 0) ? carry : sv[t - 1];
        if (b < nblocks) state[b] = start;
        __syncthreads();
        if (t == 1023) carry = v;
        __syncthreads();
    }
}

// z_j = sum(y[l_j:u_j]^2) / (T_g * rate) over the 400 ms gating blocks (75 % overlap)
__global__ void __launch_bounds__(256) gate_energy_kernel(const float* __restrict__ y, i64 N, double rate, int nblocks,
                                                          double* __restrict__ z) {
    const int j = blockIdx.x;
    if (j >= nblocks) return;
    const double Tg = 0.4, step = 0.25;
    // int(T_g * (j * step) * rate), int(T_g * (j * step + 1) * rate)  -- same operation order as pyloudnorm
    i64 lo = (i64)__dmul_rn(__dmul_rn(Tg, __dmul_rn((double)j, step)), rate);
    i64 hi = (i64)__dmul_rn(__dmul_rn(Tg, __dadd_rn(__dmul_rn((double)j, step), 1.0)), rate);
    lo = max((i64)0, min(lo, N));
    hi = max(lo, min(hi, N));
    double acc = 0.0;
    for (i64 i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float v = __ldg(y + i);
        acc += (double)__fmul_rn(v, v);
    }
    __shared__ double s[8];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += s[w];
        z[j] = __dmul_rn(1.0 / __dmul_rn(Tg, rate), tot);
    }
}

// ---- hop arithmetic of the one-pass meter (hop >= 4096 samples, i.e. rate >= 40 960 Hz) ----------------------------
// The 400 ms gating blocks are made of 100 ms hops [lo_h, lo_{h+1}) (hi_j == lo_{j+4} exactly: both are
// int(0.4 * (0.25 j + 1) * rate) with 0.25 j + 1 exact), so the squares go straight into hop sums; only the order in
// which the squares of a gating block are added differs from the stage-wise chain.
__device__ __forceinline__ i64 hop_lo(i64 h, double rate) {        // int(T_g * (h * step) * rate), pyloudnorm's order
    return (i64)__dmul_rn(__dmul_rn(0.4, __dmul_rn((double)h, 0.25)), rate);
}
__device__ __forceinline__ i64 hop_of(i64 i, double rate) {         // largest h with hop_lo(h) <= i
    i64 h = (i64)((double)i / (0.1 * rate));
    while (h > 0 && hop_lo(h, rate) > i) --h;
    while (hop_lo(h + 1, rate) <= i) ++h;
    return h;
}

// ---- one-pass loudness meter ------------------------------------------------------------------------------------
// One kernel reads the meter's input ONCE and leaves the hop energies: per block of BS samples it filters stage 1 and
// stage 2 (each: chunk end states from a zero state, block scan, true start states, second sweep), rounding stage 1's
// output to float32 in shared memory as pyloudnorm does in its array, and squares stage 2's output straight into the
// 100 ms hops.  Blocks are handed out in ticket order; a block publishes its zero-start aggregate of a stage BEFORE it
// waits for the aggregates of the few blocks in front of it (the K-weighting filters forget fast: |A^BS| ~ 1e-17 at
// 48 kHz, so `depth` terms of the look-back suffice), so no block ever waits on a block behind it and there is no serial
// chain through the blocks.  Same arithmetic per sample as the stage-wise chain; the filtered signals never leave the SM.
struct SrcMono {                       // a materialised feed (ars_metrics, block-sharded renders)
    const float* x;
    __device__ __forceinline__ bool prepare() { return true; }
    __device__ __forceinline__ float2 load(i64 i) const { return make_float2(__ldg(x + i), 0.f); }
    template <int G> __device__ __forceinline__ float value(float2 raw) const { return raw.x; }
};
struct SrcStage {                      // mean(ch0, ch1) of the final frame (rs.py:687-688), recomputed from the convolution stage's
    const float2* y;                   // output exactly as final_kernel forms it -- the meter then needs no feed array and can
    TailSpec ts;                       // run next to the final pass instead of behind it
    const RenderState* st;
    Guard g1, g2, g3;
    __device__ __forceinline__ bool prepare() {             // -> every peak guard idle (the per-sample form without them)
        g1 = make_guard(st->max_stereo);
        g2 = make_guard(st->max_pan);
        g3 = make_guard(ts.layout == LAYOUT_STEREO ? st->max_map : 0u);
        return g1.mode == 0 && g2.mode == 0 && g3.mode == 0;
    }
    __device__ __forceinline__ float2 load(i64 i) const { return __ldg(y + (i - ts.y0)); }
    template <int G> __device__ __forceinline__ float value(float2 raw) const {
        FrameIn f;
        f.v = raw;
        f.w = make_float2(0.f, 0.f);
        float o[8];
        frame_math<G, G>(f, -1, ts, g1, g2, o);             // (frame index -1: the delayed pair is not needed)
        return __fmul_rn(__fadd_rn(guardT<G>(o[0], g3), guardT<G>(o[1], g3)), 0.5f);
    }
};

struct LoudArgs {
    i64 N;
    i64 g_base = 0;                    // absolute sample index of block 0 (a multiple of BS); > 0: one rank's part of a render
    i64 src_lo = 0;                    // samples before it do not exist in this launch's source (read as zero)
    i64 e_lo = 0, e_hi = 0;            // squares count for samples in [e_lo, e_hi) only (e_hi = 0: all)
    double rate;
    ScanCoef c1, c2;
    int dep1, dep2;
    double2* agg1;                     // per block: zero-start end state of stage 1 / stage 2
    double2* agg2;
    int* flag1;                        // per block: aggregate published (zeroed before the launch)
    int* flag2;
    unsigned* ticket;                  // block counter (zeroed before the launch)
    double* part;                      // [4 b + k]: energy of block b inside hop hop_of(b BS) + k
    unsigned* mono_max;                // bits of max |feed| (null: the caller has it already)
};

__device__ __forceinline__ double2 shfl_up2(double2 v, int d) {
    return make_double2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}

// True start state of this thread's chunk for one stage: zero-state sweep, scan, publish, look back, propagate.
// The scan is two-level: inside a warp by shuffles (five steps with the tabulated A^(CH 2^k), no barrier), across the
// eight warps by one thread; a chunk's start state is then  (zero-start state of the chunks before it in its warp)
// + A^(CH lane) (state at the warp's first chunk), the power again by the tabulated squarings.  Two barriers per stage.
__device__ __forceinline__ double2 chunk_start_state(const float* mine, const ScanCoef& cf, double2* __restrict__ agg,
                                                     int* __restrict__ flag, int depth, int b, double2* sw, double2* sc) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // zero-start end state of the chunk as a weighted sum of its samples (two independent accumulators per component:
    // no recurrence, two multiply-adds per sample instead of the filter's five)
    double2 za = make_double2(0.0, 0.0), zb = make_double2(0.0, 0.0);
    #pragma unroll
    for (int j = 0; j < CH; j += 2) {
        const double x0 = (double)mine[j], x1 = (double)mine[j + 1];
        za.x = fma(cf.ends[j].x, x0, za.x);
        za.y = fma(cf.ends[j].y, x0, za.y);
        zb.x = fma(cf.ends[j + 1].x, x1, zb.x);
        zb.y = fma(cf.ends[j + 1].y, x1, zb.y);
    }
    double2 v = make_double2(za.x + zb.x, za.y + zb.y);    // inclusive scan inside the warp (zero start at its first chunk)
    #pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double2 up = shfl_up2(v, 1 << k);
        if (lane >= (1 << k)) { const double2 add = mat_vec(cf.pw[k], up); v.x += add.x; v.y += add.y; }
    }
    if (lane == 31) sw[warp] = v;
    __syncthreads();
    if (t == 0) {
        // zero-start aggregate of the block: agg = sum_w A^(32 CH (7 - w)) sw[w]; publish it before looking back
        double2 a = make_double2(0.0, 0.0);
        #pragma unroll
        for (int w = 0; w < NTB / 32; ++w) { const double2 m = mat_vec(cf.pw[5], a); a = make_double2(m.x + sw[w].x, m.y + sw[w].y); }
        agg[b] = a;
        __threadfence();
        *reinterpret_cast<volatile int*>(flag + b) = 1;
        // block start state s_b = agg[b-1] + M (agg[b-2] + M (...)), M = A^BS
        double2 s = make_double2(0.0, 0.0);
        const int d0 = b < depth ? b : depth;
        for (int d = d0; d >= 1; --d) {
            while (*reinterpret_cast<volatile int*>(flag + (b - d)) == 0) __nanosleep(40);
            __threadfence();
            const double2 p = __ldcg(agg + (b - d));
            const double2 ms = mat_vec(cf.pw[8], s);
            s = make_double2(p.x + ms.x, p.y + ms.y);
        }
        // state at every warp's first chunk
        #pragma unroll
        for (int w = 0; w < NTB / 32; ++w) {
            sc[w] = s;
            const double2 m = mat_vec(cf.pw[5], s);
            s = make_double2(m.x + sw[w].x, m.y + sw[w].y);
        }
    }
    __syncthreads();
    double2 w0 = sc[warp];                                 // A^(CH lane) (warp start state)
    #pragma unroll
    for (int k = 0; k < 5; ++k)
        if ((lane >> k) & 1) w0 = mat_vec(cf.pw[k], w0);
    const double2 prev = shfl_up2(v, 1);
    if (lane > 0) { w0.x += prev.x; w0.y += prev.y; }
    return w0;
}

// the block's samples into shared memory, eight loads in flight per thread (the feed's arithmetic would otherwise wait
// for every load in turn)
template <class SRC, int G>
__device__ __forceinline__ unsigned loud_feed(const SRC& src, float* sx, i64 base, i64 N, i64 src_lo) {
    const int t = threadIdx.x;
    unsigned mm = 0;
    #pragma unroll 1
    for (int i0 = 0; i0 < CH; i0 += 8) {
        float2 raw[8];
        #pragma unroll
        for (int u = 0; u < 8; ++u) {
            const i64 g = base + (i64)(i0 + u) * NTB + t;
            raw[u] = (g < N && g >= src_lo) ? src.load(g) : make_float2(0.f, 0.f);
        }
        #pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = (i0 + u) * NTB + t;
            const float v = src.template value<G>(raw[u]);
            sx[(i / CH) * (CH + 1) + (i % CH)] = v;
            mm = max(mm, abs_bits(v));
        }
    }
    return mm;
}

// shared memory of a meter CTA and the filter part of one block (the feed has filled sh.sx and the barrier after it is
// the caller's): both K-weighting stages of the thread's chunk and the block's hop energies into a.part
struct LoudShared {
    float sx[NTB * (CH + 1)];
    double2 sw[2][NTB / 32], sc[2][NTB / 32];                 // per stage: warp aggregates, warp start states
    i64 s_lo[3];
    unsigned s_b;
    double se[4][NTB / 32];
};
__device__ __forceinline__ void loud_block(LoudShared& sh, const LoudArgs& a, int b, i64 base) {
    const int t = threadIdx.x;
    float* mine = sh.sx + t * (CH + 1);
    // stage 1 (high shelf): float32 store, as pyloudnorm
    double2 s = chunk_start_state(mine, a.c1, a.agg1, a.flag1, a.dep1, b, sh.sw[0], sh.sc[0]);
    {
        const Biquad q1 = a.c1.q;
        #pragma unroll 8
        for (int j = 0; j < CH; ++j) mine[j] = (float)df2t(q1, (double)mine[j], s);
    }
    // stage 2 (high pass) on the thread's own chunk of stage 1's output, squared into the hops
    s = chunk_start_state(mine, a.c2, a.agg2, a.flag2, a.dep2, b, sh.sw[1], sh.sc[1]);
    // the block starts in hop hb and (hop >= 4096 samples) reaches at most hop hb + 2: three boundaries, found once
    const i64 g0 = base + (i64)t * CH;
    const int k0 = (g0 >= sh.s_lo[0] ? 1 : 0) + (g0 >= sh.s_lo[1] ? 1 : 0) + (g0 >= sh.s_lo[2] ? 1 : 0);
    const i64 next = sh.s_lo[k0 < 3 ? k0 : 2];
    // samples of the chunk that count: [ja, jb) (inside the signal and this launch's range), of which [ja, jn) fall
    // into the hop the chunk starts in -- small integers found once, not 64-bit compares per sample
    const i64 hi = a.e_hi == 0 ? a.N : (a.e_hi < a.N ? a.e_hi : a.N);
    const int ja = (int)max((i64)0, min((i64)CH, a.e_lo - g0));
    const int jb = (int)max((i64)ja, min((i64)CH, hi - g0));
    const int jn = (int)max((i64)ja, min((i64)jb, next - g0));
    double e0 = 0.0, e1 = 0.0;
    const Biquad q2 = a.c2.q;
    if (ja == 0 && jb == CH && (jn == CH || jn == 0)) {           // the whole chunk in one hop: all but ~1 % of the chunks
        double e = 0.0;
        #pragma unroll 8
        for (int j = 0; j < CH; ++j) {
            const float o = (float)df2t(q2, (double)mine[j], s);
            e += (double)__fmul_rn(o, o);
        }
        if (jn == CH) e0 = e; else e1 = e;
    } else {
        #pragma unroll 8
        for (int j = 0; j < CH; ++j) {
            const float o = (float)df2t(q2, (double)mine[j], s);
            const double sq = (double)__fmul_rn(o, o);
            if (j >= ja && j < jb) { if (j < jn) e0 += sq; else e1 += sq; }
        }
    }
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        double c = (k0 == k ? e0 : 0.0) + (k0 + 1 == k ? e1 : 0.0);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((t & 31) == 0) sh.se[k][t >> 5] = c;
    }
    __syncthreads();
    if (t < 4) {
        double tot = 0.0;
        for (int w = 0; w < NTB / 32; ++w) tot += sh.se[t][w];
        a.part[(i64)b * 4 + t] = tot;
    }
}

// Persistent CTAs: each takes the next block in ticket order until none is left (the coefficient tables in the
// kernel's parameter space are then fetched once per CTA, not once per block).
template <class SRC>
__global__ void __launch_bounds__(NTB) loudness_kernel(SRC src, LoudArgs a, int nblocks) {
    __shared__ LoudShared sh;
    // (the stages' coefficient tables are read from the kernel's parameter space: a copy in shared memory was measured --
    // fewer constant-cache misses, but its loads compete with the samples' for the shared-memory pipe: 85 against 77 us)
    const int t = threadIdx.x;
    const bool idle = src.prepare();
    unsigned mm = 0;
    for (;;) {
        if (t == 0) sh.s_b = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const int b = (int)sh.s_b;
        if (b >= nblocks) break;
        const i64 base = a.g_base + (i64)b * BS;
        if (t < 3) sh.s_lo[t] = hop_lo(hop_of(base, a.rate) + 1 + t, a.rate);
        const unsigned m1 = idle ? loud_feed<SRC, 0>(src, sh.sx, base, a.N, a.src_lo)
                                 : loud_feed<SRC, 1>(src, sh.sx, base, a.N, a.src_lo);
        if (a.e_hi == 0 || (base + BS > a.e_lo && base < a.e_hi)) mm = max(mm, m1);     // (warm-up blocks do not count)
        __syncthreads();
        loud_block(sh, a, b, base);
        __syncthreads();                                   // the block's shared arrays (and s_b) are rewritten next
    }
    if (a.mono_max) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) mm = max(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        if ((t & 31) == 0 && mm > *reinterpret_cast<volatile unsigned*>(a.mono_max)) atomicMax(a.mono_max, mm);
    }
}

// ---- the final pass inside the meter's feed ------------------------------------------------------------------------------
// The final pass (epilogue.cu) is bound by instruction issue and load latency at ~70 % of the issue slots, the meter by its
// barriers and the float64 recurrences at ~40 %; one after the other they took 109 + 78 us of the 300 s render.  Here a
// meter CTA produces its block's 8192 frames ITSELF -- guards, pan, map, PCM, sums, exactly the final pass's frame math
// (final_math.cuh) -- leaves the loudness feed in shared memory where the filters read it, and stores the frames on the
// way: one kernel whose CTAs are in different phases, so that one CTA's barriers and recurrences run under another's frame
// arithmetic; the feed never travels through memory.  Same per-sample arithmetic as the two kernels, same block order.
#ifndef ARS_FL_D
#define ARS_FL_D 4
#endif
#ifndef ARS_FL_MINB
#define ARS_FL_MINB 4
#endif
struct FinalArgs {
    const float2* y;
    float* out;
    short* pcm;
    RenderState* st;
};

// MODE 0 / 2: the float32 form of the 5.1-based layouts with the stereo guard idle / dividing (pan guard idle); 1: any
template <int C, int LAY, int MODE>
__device__ __forceinline__ unsigned final_feed(const FinalArgs& f, const TailSpec* __restrict__ tsp, const Guard& g1, const Guard& g2,
                                              const Guard& g3, float* sx, i64 base, LeanAcc& acc) {
    const TailSpec& ts = *tsp;
    const int t = threadIdx.x;
    const i64 N = ts.N, dl = ts.delay > 0 ? ts.delay : 0;
    unsigned mm = 0;
    constexpr int D = ARS_FL_D;                            // frames in flight per thread
    #pragma unroll 1
    for (int i0 = 0; i0 < CH; i0 += D) {
        float2 v[D], w[D];
        #pragma unroll
        for (int u = 0; u < D; ++u) {
            const i64 g = base + (i64)(i0 + u) * NTB + t;
            v[u] = make_float2(0.f, 0.f);
            w[u] = make_float2(0.f, 0.f);
            if (g < N) {
                v[u] = __ldcs(f.y + g);
                if (ts.layout >= LAYOUT_7_1 && g >= ts.delay) w[u] = __ldg(f.y + (g - dl));
            }
        }
        #pragma unroll
        for (int u = 0; u < D; ++u) {
            const int i = (i0 + u) * NTB + t;
            const i64 g = base + i;
            float mv = 0.f;
            if (g < N) {
                float o[8];
                bool literal = false;
                if constexpr (MODE != 1) literal = lean_math_split<C, LAY, MODE>(v[u], w[u], ts, g1, o);
                if (MODE != 1 && !literal) {
                    split_emit<C, 0>(o, (unsigned)g, f.out, f.pcm, nullptr, acc);
                    mv = __fmul_rn(__fadd_rn(o[0], o[1]), 0.5f);
                } else if (MODE != 1) {
                    const float4 r = slow_frame<C>(f.y, g, tsp, f.st->max_stereo, f.out, f.pcm, nullptr);
                    acc.pkf = fmaxf(acc.pkf, r.x);
                    acc.ss += (double)r.y;
                    mv = r.z;
                } else {
                    FrameIn fr;
                    fr.v = v[u];
                    fr.w = w[u];
                    frame_math<1, 1>(fr, g, ts, g1, g2, o);
                    float fs = 0.f;
                    #pragma unroll
                    for (int c = 0; c < C; ++c) {
                        o[c] = guard1(o[c], g3);
                        acc.pkf = fmaxf(acc.pkf, fabsf(o[c]));
                        fs = __fmaf_rn(o[c], o[c], fs);
                    }
                    acc.ss += (double)fs;                  // (a NaN sample makes this sum NaN for good: read back as "NaN seen")
                    if (f.out) {
                        float2* p = reinterpret_cast<float2*>(f.out) + g * (C / 2);
                        #pragma unroll
                        for (int c = 0; c < C; c += 2) p[c >> 1] = make_float2(o[c], o[c + 1]);
                    }
                    if (f.pcm) {
                        unsigned* p = reinterpret_cast<unsigned*>(f.pcm) + g * (C / 2);
                        #pragma unroll
                        for (int c = 0; c < C; c += 2) __stcs(p + (c >> 1), pcm_pair(o[c], o[c + 1]));
                    }
                    mv = __fmul_rn(__fadd_rn(o[0], o[1]), 0.5f);
                }
            }
            sx[(i / CH) * (CH + 1) + (i % CH)] = mv;
            mm = max(mm, abs_bits(mv));
        }
    }
    return mm;
}

template <int C, int LAY>
__global__ void __launch_bounds__(NTB, ARS_FL_MINB) final_loud_kernel(FinalArgs f, const __grid_constant__ TailSpec ts, LoudArgs a, int nblocks) {
    __shared__ LoudShared sh;
    const int t = threadIdx.x;
    const Guard g1 = make_guard(f.st->max_stereo), g2 = make_guard(f.st->max_pan);
    const Guard g3 = make_guard(ts.layout == LAYOUT_STEREO ? f.st->max_map : 0u);
    int mode = 1;
    if (LAY >= 1 && ts.split_ok && g2.mode == 0) {
        if (g1.mode == 0) mode = 0;
        else if (g1.mode == 1 && g1.r != 0.f) mode = 2;
    }
    LeanAcc acc = {0.f, 0u, 0.0};
    unsigned mm = 0;
    for (;;) {
        if (t == 0) sh.s_b = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const int b = (int)sh.s_b;
        if (b >= nblocks) break;
        const i64 base = (i64)b * BS;
        if (t < 3) sh.s_lo[t] = hop_lo(hop_of(base, a.rate) + 1 + t, a.rate);
        unsigned m1;
        if constexpr (LAY >= 1) {
            if (mode == 0) m1 = final_feed<C, LAY, 0>(f, &ts, g1, g2, g3, sh.sx, base, acc);
            else if (mode == 2) m1 = final_feed<C, LAY, 2>(f, &ts, g1, g2, g3, sh.sx, base, acc);
            else m1 = final_feed<C, LAY, 1>(f, &ts, g1, g2, g3, sh.sx, base, acc);
        } else {
            m1 = final_feed<C, 0, 1>(f, &ts, g1, g2, g3, sh.sx, base, acc);
        }
        mm = max(mm, m1);
        __syncthreads();
        loud_block(sh, a, b, base);
        __syncthreads();
    }
    const unsigned pk = (acc.ss != acc.ss) ? 0x7fc00000u : __float_as_uint(acc.pkf);
    block_atomic_max(pk, &f.st->peak_final);
    block_atomic_max(mm, &f.st->mono_max);
    block_atomic_add(acc.ss, &f.st->sumsq);
}

// z_j = (E_j + E_{j+1} + E_{j+2} + E_{j+3}) / (T_g * rate), E_h gathered from the per-block partial sums in block order
__device__ __forceinline__ double hop_block_energy(const double* __restrict__ part, i64 N, double rate, int j) {
    double tot = 0.0;
    for (int q = 0; q < 4; ++q) {
        const i64 h = j + q;
        const i64 lo = min(hop_lo(h, rate), N), hi = min(hop_lo(h + 1, rate), N);
        if (hi <= lo) continue;
        for (i64 b = lo / BS; b <= (hi - 1) / BS; ++b) {
            const i64 k = h - hop_of(b * BS, rate);
            if (k >= 0 && k < 4) tot += part[b * 4 + k];
        }
    }
    return __dmul_rn(1.0 / __dmul_rn(0.4, rate), tot);
}
// (one thread per gating block over many CTAs: gathered inside the single-CTA gate kernel this took 16 us longer)
__global__ void __launch_bounds__(256) hop_combine_kernel(const double* __restrict__ part, i64 N, double rate, int nblocks,
                                                          double* __restrict__ z) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nblocks) z[j] = hop_block_energy(part, N, rate, j);
}

static int g_lufs_fused = 1;
static int g_lufs_ctas_per_sm = 3;
void loudness_set_fused(int on) { g_lufs_fused = on ? 1 : 0; }
void loudness_set_ctas_per_sm(int n) { g_lufs_ctas_per_sm = n < 1 ? 1 : (n > 8 ? 8 : n); }

static void k_weighting(double rate, Biquad out[2]) {
    // pyloudnorm IIRfilter coefficients: high shelf +4 dB @1500 Hz Q=1/sqrt2, high pass 38 Hz Q=0.5
    const double PI = 3.14159265358979323846;
    {
        const double G = 4.0, Q = 1.0 / std::sqrt(2.0), fc = 1500.0;
        const double A = std::pow(10.0, G / 40.0);
        const double w0 = 2.0 * PI * (fc / rate);
        const double al = std::sin(w0) / (2.0 * Q);
        const double cw = std::cos(w0), sA = std::sqrt(A);
        const double b0 = A * ((A + 1) + (A - 1) * cw + 2 * sA * al);
        const double b1 = -2 * A * ((A - 1) + (A + 1) * cw);
        const double b2 = A * ((A + 1) + (A - 1) * cw - 2 * sA * al);
        const double a0 = (A + 1) - (A - 1) * cw + 2 * sA * al;
        const double a1 = 2 * ((A - 1) - (A + 1) * cw);
        const double a2 = (A + 1) - (A - 1) * cw - 2 * sA * al;
        out[0] = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
    }
    {
        const double Q = 0.5, fc = 38.0;
        const double w0 = 2.0 * PI * (fc / rate);
        const double al = std::sin(w0) / (2.0 * Q);
        const double cw = std::cos(w0);
        const double b0 = (1 + cw) / 2, b1 = -(1 + cw), b2 = (1 + cw) / 2;
        const double a0 = 1 + al, a1 = -2 * cw, a2 = 1 - al;
        out[1] = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
    }
}

static void run_biquad(const float* d_in, float* d_out, i64 N, const Biquad& q) {
    Ctx& c = ctx();
    const int nblocks = (int)((N + BS - 1) / BS);
    ScanCoef cf = make_coef(q);
    double2* st = c.buf("lufs.state", sizeof(double2) * (size_t)nblocks).as<double2>();
    biquad_kernel<0><<<nblocks, NTB, 0, c.stream>>>(d_in, N, cf, st, nullptr);
    ARS_LAUNCH_CHECK();
    BlockPow bp;
    bp.p[0] = cf.pw[8];
    for (int k = 1; k < 10; ++k) bp.p[k] = mat_mul(bp.p[k - 1], bp.p[k - 1]);
    block_scan_kernel<<<1, 1024, 0, c.stream>>>(st, nblocks, bp);
    ARS_LAUNCH_CHECK();
    biquad_kernel<1><<<nblocks, NTB, 0, c.stream>>>(d_in, N, cf, st, d_out);
    ARS_LAUNCH_CHECK();
    count_launch(3);
}

int loudness_blocks(i64 N, double rate) {
    // pyloudnorm: numBlocks = int(np.round(((T - T_g) / (T_g * step))) + 1); np.round = half to even
    const double T = (double)N / rate;
    const double v = (T - 0.4) / (0.4 * 0.25);
    return (int)(std::nearbyint(v) + 1);
}

__device__ __forceinline__ void block_sum_dc(double& v, int& n) {
    __shared__ double sv[32];
    __shared__ int sn[32];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; sn[threadIdx.x >> 5] = n; }
    __syncthreads();
    v = 0.0; n = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { v += sv[w]; n += sn[w]; }     // same order in every thread
    __syncthreads();
}

// BS.1770-4 gating exactly as pyloudnorm implements it (mono => channel gain 1), one CTA over the block
// energies: absolute gate at -70 LUFS, relative gate 10 LU under the abs-gated mean, loudness of the survivors.
// Also applies the reference's silence test (rs.py:689): peak of the mono feed < 1e-6 -> -inf.
__global__ void __launch_bounds__(1024) gate_kernel(const double* __restrict__ z, int nb, const unsigned* __restrict__ mono_max,
                                                    double* lufs_out) {
    if (__uint_as_float(*mono_max) < 1e-6f) {
        if (threadIdx.x == 0) *lufs_out = -CUDART_INF;
        return;
    }
    // l_j = -0.691 + 10 log10(z_j) is monotone in z_j, so both gates compare z_j itself: l_j >= -70 <=> z_j >= 10^-6.9309,
    // and l_j > Gamma_r = -0.691 + 10 log10(mean z) - 10 <=> z_j > mean z / 10 -- no logarithm per block
    const double z_abs = 1.1724653045822981e-07;          // 10^((-70 + 0.691) / 10)
    double s = 0.0;
    int n = 0;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        if (z[j] >= z_abs) { s += z[j]; ++n; }
    }
    block_sum_dc(s, n);
    const double z_rel = n > 0 ? (s / (double)n) * 0.1 : CUDART_NAN;
    double s2 = 0.0;
    int n2 = 0;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        if (z[j] > z_rel && z[j] > z_abs) { s2 += z[j]; ++n2; }
    }
    block_sum_dc(s2, n2);
    if (threadIdx.x == 0) *lufs_out = -0.691 + 10.0 * log10(n2 > 0 ? s2 / (double)n2 : 0.0);   // log10(0) = -inf, as numpy
}

template <class SRC>
static int loudness_run(SRC src, i64 N, double rate, unsigned* d_mono_max_out, const float* d_mono_stagewise,
                        const unsigned* d_mono_max, double* d_lufs) {
    Ctx& c = ctx();
    if (!((double)N >= 0.4 * rate)) return 1;            // "Audio must have length greater than the block size"
    const int nb = loudness_blocks(N, rate);
    if (nb <= 0) return 1;
    Biquad q[2];
    k_weighting(rate, q);
    double* dz = c.buf("lufs.z", sizeof(double) * (size_t)nb).as<double>();
    const ScanCoef c1 = make_coef(q[0]), c2 = make_coef(q[1]);
    // look-back depth at which the dropped start-state terms are below 1e-25 of the kept ones (0: use the block scan)
    auto depth_of = [](const Mat2& M) {
        const double nrm = std::sqrt(M.m00 * M.m00 + M.m01 * M.m01 + M.m10 * M.m10 + M.m11 * M.m11);
        if (!(nrm < 0.05)) return 0;
        return std::max(2, (int)std::ceil(-25.0 * std::log(10.0) / std::log(nrm)));
    };
    const int dep1 = depth_of(c1.pw[8]), dep2 = depth_of(c2.pw[8]);
    if (g_lufs_fused && 0.1 * rate >= 4096.0 && dep1 > 0 && dep2 > 0) {
        const int nblocks = (int)((N + BS - 1) / BS);
        // [agg1 | agg2 | part | flag1 | flag2 | ticket]: the flags and the ticket are zeroed per launch
        const size_t off_flags = sizeof(double2) * 2 * (size_t)nblocks + sizeof(double) * 4 * (size_t)nblocks;
        const size_t bytes_flags = sizeof(int) * (2 * (size_t)nblocks + 4);
        char* ws = c.buf("lufs.onepass", off_flags + bytes_flags).as<char>();
        LoudArgs a;
        a.N = N;
        a.rate = rate;
        a.c1 = c1; a.c2 = c2;
        a.dep1 = dep1; a.dep2 = dep2;
        a.agg1 = reinterpret_cast<double2*>(ws);
        a.agg2 = a.agg1 + nblocks;
        a.part = reinterpret_cast<double*>(a.agg2 + nblocks);
        a.flag1 = reinterpret_cast<int*>(ws + off_flags);
        a.flag2 = a.flag1 + nblocks;
        a.ticket = reinterpret_cast<unsigned*>(a.flag2 + nblocks);
        a.mono_max = d_mono_max_out;
        ARS_CUDA(cudaMemsetAsync(ws + off_flags, 0, bytes_flags, c.stream));
        {
            KernelScope prof("loudness_kernel (K-weighting stages + hop energies, one pass)", (double)N * (d_mono_max_out ? 8.0 : 4.0));
            // next to the final pass the meter takes a few CTAs per SM only, so that both kernels are resident together
            static const int feed_per_sm = getenv("ARS_LUFS_FEED_CTAS") ? std::max(1, std::min(8, atoi(getenv("ARS_LUFS_FEED_CTAS")))) : 6;   // (experiments)
            const int per_sm = d_mono_max_out ? g_lufs_ctas_per_sm : feed_per_sm;
            const int grid = std::max(1, std::min(nblocks, c.sm_count * per_sm));
            loudness_kernel<SRC><<<grid, NTB, 0, c.stream>>>(src, a, nblocks);
        }
        KernelScope prof("loudness gating (hop_combine + gate)", 0.0);
        hop_combine_kernel<<<ceil_div(nb, 256), 256, 0, c.stream>>>(a.part, N, rate, nb, dz);
        ARS_LAUNCH_CHECK();
        count_launch(2);
    } else {
        ARS_CHECK(d_mono_stagewise != nullptr, "loudness: the stage-wise chain needs a materialised feed");
        float* y1 = c.buf("lufs.y1", sizeof(float) * (size_t)N).as<float>();
        float* y2 = c.buf("lufs.y2", sizeof(float) * (size_t)N).as<float>();
        run_biquad(d_mono_stagewise, y1, N, q[0]);
        run_biquad(y1, y2, N, q[1]);
        gate_energy_kernel<<<nb, 256, 0, c.stream>>>(y2, N, rate, nb, dz);
        ARS_LAUNCH_CHECK();
        count_launch();
    }
    KernelScope prof("loudness gating (hop_combine + gate)", 0.0);
    gate_kernel<<<1, 1024, 0, c.stream>>>(dz, nb, d_mono_max, d_lufs);
    ARS_LAUNCH_CHECK();
    count_launch();
    return 0;
}

// Enqueues the whole loudness measurement; *d_lufs receives the value.  Returns 1 (and enqueues nothing)
// when the signal is shorter than one 400 ms block (pyloudnorm raises -> the reference reports None).
int integrated_loudness_async(const float* d_mono, i64 N, double rate, const unsigned* d_mono_max, double* d_lufs) {
    SrcMono src;
    src.x = d_mono;
    return loudness_run(src, N, rate, nullptr, d_mono, d_mono_max, d_lufs);
}

bool loudness_from_stage_possible(double rate) { return g_lufs_fused && 0.1 * rate >= 4096.0; }

// The same measurement fed straight from the convolution stage's output (mean of the first two final channels is
// recomputed per sample): no feed array, so the meter can run next to the final pass.  Needs the guards' maxima in
// *d_state (max_stereo, max_pan, and max_map for the Stereo layout); writes d_state->mono_max and d_state->lufs.
int integrated_loudness_from_stage(const float2* d_y, const TailSpec& ts, double rate, RenderState* d_state) {
    ARS_CHECK(loudness_from_stage_possible(rate), "loudness: the one-pass meter needs rate >= 40 960 Hz");
    SrcStage src;
    src.y = d_y;
    src.ts = ts;
    src.st = d_state;
    return loudness_run(src, ts.N, rate, &d_state->mono_max, nullptr, &d_state->mono_max, &d_state->lufs);
}

static int g_final_in_meter = 0;       // (measured: 214 us against 109 + 77 us for the two kernels -- a CTA in its filter phase takes warps from the frame arithmetic)
void loudness_set_final_in_meter(int on) { g_final_in_meter = on ? 1 : 0; }

// The whole tail of a render in one kernel: final pass (frames to d_out / d_pcm, peak and sum of squares into *d_state)
// inside the one-pass meter's feed, then the gating; the loudness lands in d_state->lufs.  -> false (nothing enqueued) when
// this form does not apply: the caller runs tail_final and the meter one after the other.
bool final_with_loudness(const float2* d_y, const TailSpec& ts_in, double rate, RenderState* d_state, float* d_out,
                         short* d_pcm) {
    Ctx& c = ctx();
    const i64 N = ts_in.N;
    const bool whole = ts_in.i_lo == 0 && (ts_in.i_hi < 0 || ts_in.i_hi == N) && ts_in.y0 == 0 && ts_in.out0 == 0;
    if (!g_final_in_meter || !whole || N <= 0 || N >= ((i64)1 << 28) || !loudness_from_stage_possible(rate)) return false;
    if (!((double)N >= 0.4 * rate)) return false;
    const int nb = loudness_blocks(N, rate);
    if (nb <= 0) return false;
    Biquad q[2];
    k_weighting(rate, q);
    const ScanCoef c1 = make_coef(q[0]), c2 = make_coef(q[1]);
    auto depth_of = [](const Mat2& M) {
        const double nrm = std::sqrt(M.m00 * M.m00 + M.m01 * M.m01 + M.m10 * M.m10 + M.m11 * M.m11);
        if (!(nrm < 0.05)) return 0;
        return std::max(2, (int)std::ceil(-25.0 * std::log(10.0) / std::log(nrm)));
    };
    const int dep1 = depth_of(c1.pw[8]), dep2 = depth_of(c2.pw[8]);
    if (dep1 <= 0 || dep2 <= 0) return false;
    TailSpec ts = ts_in;
    ts.i_hi = N;
    tail_prepare(ts);
    double* dz = c.buf("lufs.z", sizeof(double) * (size_t)nb).as<double>();
    const int nblocks = (int)((N + BS - 1) / BS);
    const size_t off_flags = sizeof(double2) * 2 * (size_t)nblocks + sizeof(double) * 4 * (size_t)nblocks;
    const size_t bytes_flags = sizeof(int) * (2 * (size_t)nblocks + 4);
    char* ws = c.buf("lufs.onepass", off_flags + bytes_flags).as<char>();
    LoudArgs a;
    a.N = N;
    a.rate = rate;
    a.c1 = c1; a.c2 = c2;
    a.dep1 = dep1; a.dep2 = dep2;
    a.agg1 = reinterpret_cast<double2*>(ws);
    a.agg2 = a.agg1 + nblocks;
    a.part = reinterpret_cast<double*>(a.agg2 + nblocks);
    a.flag1 = reinterpret_cast<int*>(ws + off_flags);
    a.flag2 = a.flag1 + nblocks;
    a.ticket = reinterpret_cast<unsigned*>(a.flag2 + nblocks);
    a.mono_max = &d_state->mono_max;
    FinalArgs f;
    f.y = d_y; f.out = d_out; f.pcm = d_pcm; f.st = d_state;
    ARS_CUDA(cudaMemsetAsync(ws + off_flags, 0, bytes_flags, c.stream));
    {
        KernelScope prof("final + loudness kernel (guards, pan, map, clip, PCM16, sums; K-weighting stages + hop energies)",
                         (double)N * (8.0 + (d_pcm ? 2.0 * ts.C : 0.0) + (d_out ? 4.0 * ts.C : 0.0)));
        const int grid = std::max(1, std::min(nblocks, c.sm_count * ARS_FL_MINB));
        if (ts.C == 2) final_loud_kernel<2, 0><<<grid, NTB, 0, c.stream>>>(f, ts, a, nblocks);
        else if (ts.C == 6) final_loud_kernel<6, 1><<<grid, NTB, 0, c.stream>>>(f, ts, a, nblocks);
        else if (ts.layout == LAYOUT_7_1) final_loud_kernel<8, 2><<<grid, NTB, 0, c.stream>>>(f, ts, a, nblocks);
        else final_loud_kernel<8, 3><<<grid, NTB, 0, c.stream>>>(f, ts, a, nblocks);
        ARS_LAUNCH_CHECK();
    }
    KernelScope prof("loudness gating (hop_combine + gate)", 0.0);
    hop_combine_kernel<<<ceil_div(nb, 256), 256, 0, c.stream>>>(a.part, N, rate, nb, dz);
    gate_kernel<<<1, 1024, 0, c.stream>>>(dz, nb, &d_state->mono_max, &d_state->lufs);
    ARS_LAUNCH_CHECK();
    count_launch(3);
    return true;
}

// ---- the meter split over the ranks of a block-sharded render (sharding.py) -------------------------------------
// A rank filters its own stretch of the signal (after a warm-up of a few blocks that the previous rank's tail of the
// stage output provides: the K-weighting filters forget 1e-17 per block) and leaves the energies of its samples per
// 100 ms hop, indexed by ABSOLUTE hop; the ranks add those vectors (one small all-reduce) and every rank gates.
int loudness_hop_count(i64 N, double rate) {
    const int nb = loudness_blocks(N, rate);
    return nb > 0 ? nb + 3 : 0;
}

__global__ void __launch_bounds__(256) hop_energy_kernel(const double* __restrict__ part, int nblocks, i64 g_base, i64 N,
                                                         double rate, i64 e_lo, i64 e_hi, double* __restrict__ hops, int n_hops) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hops) return;
    i64 lo = min(hop_lo(h, rate), N), hi = min(hop_lo(h + 1, rate), N);
    lo = max(lo, e_lo);
    hi = min(hi, e_hi);
    double tot = 0.0;
    if (hi > lo) {
        for (i64 b = lo / BS; b <= (hi - 1) / BS; ++b) {
            const i64 bl = b - g_base / BS;
            const i64 k = (i64)h - hop_of(b * BS, rate);
            if (bl >= 0 && bl < nblocks && k >= 0 && k < 4) tot += part[bl * 4 + k];
        }
    }
    hops[h] = tot;
}

__global__ void __launch_bounds__(256) hops_to_z_kernel(const double* __restrict__ hops, int nb, double rate, double* __restrict__ z) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    const double tot = ((hops[j] + hops[j + 1]) + hops[j + 2]) + hops[j + 3];
    z[j] = __dmul_rn(1.0 / __dmul_rn(0.4, rate), tot);
}

int loudness_hops_from_stage(const float2* d_y, const TailSpec& ts, double rate, RenderState* d_state, i64 e_lo, i64 e_hi,
                             double* d_hops, int n_hops) {
    Ctx& c = ctx();
    const i64 N = ts.N;
    ARS_CHECK(loudness_from_stage_possible(rate), "loudness: the one-pass meter needs rate >= 40 960 Hz");
    ARS_CHECK(n_hops == loudness_hop_count(N, rate) && n_hops > 0, "loudness: wrong hop count");
    ARS_CHECK(e_lo >= 0 && e_lo <= e_hi && e_hi <= N, "loudness: bad sample range");
    if (e_hi <= e_lo) {
        ARS_CUDA(cudaMemsetAsync(d_hops, 0, sizeof(double) * (size_t)n_hops, c.stream));
        return 0;
    }
    Biquad q[2];
    k_weighting(rate, q);
    const ScanCoef c1 = make_coef(q[0]), c2 = make_coef(q[1]);
    auto depth_of = [](const Mat2& M) {
        const double nrm = std::sqrt(M.m00 * M.m00 + M.m01 * M.m01 + M.m10 * M.m10 + M.m11 * M.m11);
        return std::max(2, (int)std::ceil(-25.0 * std::log(10.0) / std::log(nrm)));
    };
    const i64 g_base = std::max<i64>(0, (e_lo - 2 * (i64)BS) / BS * BS);
    const int nblocks = (int)((e_hi - g_base + BS - 1) / BS);
    const size_t off_flags = sizeof(double2) * 2 * (size_t)nblocks + sizeof(double) * 4 * (size_t)nblocks;
    const size_t bytes_flags = sizeof(int) * (2 * (size_t)nblocks + 4);
    char* ws = c.buf("lufs.onepass", off_flags + bytes_flags).as<char>();
    LoudArgs a;
    a.N = N;
    a.g_base = g_base;
    a.src_lo = ts.y0;
    a.e_lo = e_lo;
    a.e_hi = e_hi;
    a.rate = rate;
    a.c1 = c1; a.c2 = c2;
    a.dep1 = depth_of(c1.pw[8]); a.dep2 = depth_of(c2.pw[8]);
    a.agg1 = reinterpret_cast<double2*>(ws);
    a.agg2 = a.agg1 + nblocks;
    a.part = reinterpret_cast<double*>(a.agg2 + nblocks);
    a.flag1 = reinterpret_cast<int*>(ws + off_flags);
    a.flag2 = a.flag1 + nblocks;
    a.ticket = reinterpret_cast<unsigned*>(a.flag2 + nblocks);
    a.mono_max = &d_state->mono_max;
    SrcStage src;
    src.y = d_y;
    src.ts = ts;
    src.st = d_state;
    ARS_CUDA(cudaMemsetAsync(ws + off_flags, 0, bytes_flags, c.stream));
    {
        KernelScope prof("loudness_kernel (K-weighting stages + hop energies, one pass)", 8.0 * (double)(e_hi - g_base));
        const int grid = std::max(1, std::min(nblocks, c.sm_count * 6));
        loudness_kernel<SrcStage><<<grid, NTB, 0, c.stream>>>(src, a, nblocks);
    }
    hop_energy_kernel<<<ceil_div(n_hops, 256), 256, 0, c.stream>>>(a.part, nblocks, g_base, N, rate, e_lo, e_hi, d_hops, n_hops);
    ARS_LAUNCH_CHECK();
    count_launch(2);
    return 0;
}

int loudness_gate_from_hops(const double* d_hops, int n_hops, i64 N, double rate, RenderState* d_state) {
    Ctx& c = ctx();
    const int nb = loudness_blocks(N, rate);
    if (!((double)N >= 0.4 * rate) || nb <= 0) return 1;
    ARS_CHECK(n_hops == nb + 3, "loudness: wrong hop count");
    double* dz = c.buf("lufs.z", sizeof(double) * (size_t)nb).as<double>();
    hops_to_z_kernel<<<ceil_div(nb, 256), 256, 0, c.stream>>>(d_hops, nb, rate, dz);
    gate_kernel<<<1, 1024, 0, c.stream>>>(dz, nb, &d_state->mono_max, &d_state->lufs);
    ARS_LAUNCH_CHECK();
    count_launch(2);
    return 0;
}

// ---- 4x-oversampled true peak (ITU-R BS.1770-4 Annex 2) -- an ADD-ON: the reference's "true_peak_dbfs" is the plain
// sample peak (rs.py:695-697) and stays that; BASELINE's north star names the oversampled figure, so it is reported
// next to it.  Interpolator: the standard's 48-tap FIR as four 12-tap phases (coefficients restated from Annex 2; every
// value is a multiple of 2^-13).  The peak is taken over the four phases of every sample of every signal, the filter's
// tail past the last sample included.
__constant__ float TP_COEF[4][12] = {
    {0.0017089843750f, 0.0109863281250f, -0.0196533203125f, 0.0332031250000f, -0.0594482421875f, 0.1373291015625f,
     0.9721679687500f, -0.1022949218750f, 0.0476074218750f, -0.0266113281250f, 0.0148925781250f, -0.0083007812500f},
    {-0.0291748046875f, 0.0292968750000f, -0.0517578125000f, 0.0891113281250f, -0.1665039062500f, 0.4650878906250f,
     0.7797851562500f, -0.2003173828125f, 0.1015625000000f, -0.0582275390625f, 0.0330810546875f, -0.0189208984375f},
    {-0.0189208984375f, 0.0330810546875f, -0.0582275390625f, 0.1015625000000f, -0.2003173828125f, 0.7797851562500f,
     0.4650878906250f, -0.1665039062500f, 0.0891113281250f, -0.0517578125000f, 0.0292968750000f, -0.0291748046875f},
    {-0.0083007812500f, 0.0148925781250f, -0.0266113281250f, 0.0476074218750f, -0.1022949218750f, 0.9721679687500f,
     0.1373291015625f, -0.0594482421875f, 0.0332031250000f, -0.0196533203125f, 0.0109863281250f, 0.0017089843750f}};

constexpr int TP_TILE = 1024;          // frames per tile (256 threads x 4 frames)
constexpr int TP_HIST = 11;            // taps - 1

struct TpStage {                       // signals s_i = a_i L + b_i R + c_i f32(f32(L + R) * 0.707) of the stage output
    const float2* y;
    i64 y0;
    float a[3], b[3], c[3];
    __device__ __forceinline__ void get(i64 i, float (&v)[3]) const {
        const float2 f = __ldg(y + (i - y0));
        const float m = __fmul_rn(__fadd_rn(f.x, f.y), 0.707f);
        #pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = a[k] * f.x + b[k] * f.y + c[k] * m;
    }
};
struct TpArray {                       // three channels [c0, c0 + 3) of an interleaved (n, ch) array (absent ones: zero)
    const float* x;
    int ch, c0;
    __device__ __forceinline__ void get(i64 i, float (&v)[3]) const {
        #pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = (c0 + k < ch) ? __ldg(x + i * ch + c0 + k) : 0.f;
    }
};

template <class SRC>
__global__ void __launch_bounds__(256) true_peak_kernel(SRC src, i64 N, unsigned* __restrict__ out_bits) {
    __shared__ float s[3][TP_TILE + TP_HIST + 1];
    const int t = threadIdx.x;
    float pk[3] = {0.f, 0.f, 0.f};
    const i64 ntiles = (N + TP_HIST + TP_TILE - 1) / TP_TILE;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 t0 = tile * TP_TILE;
        for (int i = t; i < TP_TILE + TP_HIST; i += 256) {
            const i64 g = t0 - TP_HIST + i;
            float v[3] = {0.f, 0.f, 0.f};
            if (g >= 0 && g < N) src.get(g, v);
            s[0][i] = v[0]; s[1][i] = v[1]; s[2][i] = v[2];
        }
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < 3; ++k) {
            float w[4 + TP_HIST];
            #pragma unroll
            for (int q = 0; q < 4 + TP_HIST; ++q) w[q] = s[k][4 * t + q];      // w[q] = sample (t0 + 4 t - 11 + q)
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                #pragma unroll
                for (int p = 0; p < 4; ++p) {
                    float acc = 0.f;
                    #pragma unroll
                    for (int h = 0; h < 12; ++h) acc = fmaf(TP_COEF[p][h], w[j + TP_HIST - h], acc);
                    pk[k] = fmaxf(pk[k], fabsf(acc));
                }
            }
        }
        __syncthreads();
    }
    #pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned m = __float_as_uint(pk[k]);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((t & 31) == 0 && m > *reinterpret_cast<volatile unsigned*>(out_bits + k)) atomicMax(out_bits + k, m);
    }
}

// Every output channel of a render is one of (at most) three signals times a non-negative gain: L, R and the mono mix
// for the 5.1-based layouts (rs.py:484-494; the delayed side / height pairs are attenuated copies of the rear pair),
// two L / R / mono combinations for the Stereo down-mix (rs.py:533-535) -- and oversampling is linear.  So the filter
// runs over those signals of the stage output and the host applies the gains and the peak guards' scales
// (true_peak_weights); the float32 roundings of the per-sample products (<= 6e-8 relative) are not replayed.
void true_peak_from_stage(const float2* d_y, const TailSpec& ts, RenderState* d_state) {
    Ctx& c = ctx();
    TpStage src;
    src.y = d_y;
    src.y0 = ts.y0;
    for (int k = 0; k < 3; ++k) src.a[k] = src.b[k] = src.c[k] = 0.f;
    if (ts.layout == LAYOUT_STEREO) {
        src.a[0] = (float)(ts.g_fl + 0.5 * ts.g_rl); src.c[0] = (float)(0.707 * ts.g_c);
        src.b[1] = (float)(ts.g_fr + 0.5 * ts.g_rr); src.c[1] = (float)(0.707 * ts.g_c);
    } else {
        src.a[0] = 1.f; src.b[1] = 1.f; src.c[2] = 1.f;
    }
    ARS_CUDA(cudaMemsetAsync(d_state->tp_bits, 0, sizeof(unsigned) * 4, c.stream));
    KernelScope prof("true_peak_kernel (4x polyphase FIR, add-on)", 8.0 * (double)ts.N);
    const i64 ntiles = (ts.N + TP_HIST + TP_TILE - 1) / TP_TILE;
    true_peak_kernel<TpStage><<<(int)std::min<i64>(ntiles, (i64)c.sm_count * 8), 256, 0, c.stream>>>(src, ts.N, d_state->tp_bits);
    ARS_LAUNCH_CHECK();
    count_launch();
}

double true_peak_linear(const RenderState& h, const TailSpec& ts) {
    auto f = [](unsigned u) { float v; memcpy(&v, &u, 4); return (double)v; };
    auto guard_scale = [&](unsigned bits) {
        const double m = f(bits);
        return m > 1.0 ? 1.0 / m : ((m > 0.0 && m < 1e-9) ? 0.0 : 1.0);
    };
    double scale = guard_scale(h.max_stereo) * guard_scale(h.max_pan);
    double w[3];
    if (ts.layout == LAYOUT_STEREO) {
        scale *= guard_scale(h.max_map);
        w[0] = w[1] = 1.0;
        w[2] = 0.0;
    } else {
        w[0] = std::max(std::fabs(ts.g_fl), std::fabs(ts.g_rl));
        w[1] = std::max(std::fabs(ts.g_fr), std::fabs(ts.g_rr));
        w[2] = std::max(std::fabs(ts.g_c), (double)ts.g_lfe);
    }
    double tp = 0.0;
    for (int k = 0; k < 3; ++k) tp = std::max(tp, w[k] * f(h.tp_bits[k]));
    return tp * scale;
}

// the same over the channels of a materialised (n, ch) array -> linear peak (synchronises)
double true_peak_of_array(const float* d_x, i64 n, int ch) {
    Ctx& c = ctx();
    unsigned* bits = c.buf("tp.bits", sizeof(unsigned) * 4).as<unsigned>();
    double tp = 0.0;
    for (int c0 = 0; c0 < ch; c0 += 3) {
        TpArray src;
        src.x = d_x;
        src.ch = ch;
        src.c0 = c0;
        ARS_CUDA(cudaMemsetAsync(bits, 0, sizeof(unsigned) * 4, c.stream));
        const i64 ntiles = (n + TP_HIST + TP_TILE - 1) / TP_TILE;
        true_peak_kernel<TpArray><<<(int)std::min<i64>(ntiles, (i64)c.sm_count * 8), 256, 0, c.stream>>>(src, n, bits);
        ARS_LAUNCH_CHECK();
        count_launch();
        unsigned h[4];
        ARS_CUDA(cudaMemcpyAsync(h, bits, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
        ARS_CUDA(cudaStreamSynchronize(c.stream));
        for (int k = 0; k < 3; ++k) { float v; memcpy(&v, &h[k], 4); tp = std::max(tp, (double)v); }
    }
    return tp;
}

// ---- spectrogram of the visualiser (rs.py:626-634): scipy.signal.spectrogram(x, fs, window='hann', nperseg,
// noverlap = nperseg // 2) with scipy's defaults -- detrend='constant' (each segment's mean removed), one-sided power
// spectral density: Sxx[k, j] = c_k |FFT(w (x_j - mean x_j))[k]|^2 / (fs sum w^2), c_k = 2 except at DC and Nyquist.
// One CTA per segment: mean, periodic Hann window, an in-shared-memory Stockham radix-2 FFT in natural order (this is
// plotting support, far off the render path: the pass kernels' permuted order would need an extra reordering pass).
__global__ void __launch_bounds__(256) spectrogram_kernel(const float* __restrict__ x, i64 n, int stride, int logN, int hop,
                                                          int nseg, float scale, float* __restrict__ out) {
    extern __shared__ float2 sbuf[];
    const int N = 1 << logN, t = threadIdx.x, j = blockIdx.x;
    float2* a = sbuf;
    float2* b = sbuf + N;
    const float* seg = x + (i64)j * hop * stride;
    float part = 0.f;
    for (int i = t; i < N; i += blockDim.x) {
        const float v = seg[(i64)i * stride];
        a[i] = make_float2(v, 0.f);
        part += v;
    }
    __shared__ float red[8];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((t & 31) == 0) red[t >> 5] = part;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mean += red[w];
    mean /= (float)N;
    for (int i = t; i < N; i += blockDim.x) {
        float s, c;
        sincospif(2.0f * (float)i / (float)N, &s, &c);
        a[i].x = (a[i].x - mean) * (0.5f - 0.5f * c);                 // scipy get_window('hann', N): periodic
    }
    __syncthreads();
    // Stockham autosort, decimation in frequency: after stage q (L = N >> q) the output is in natural order at the end
    for (int L = N, m = 1; L > 1; L >>= 1, m <<= 1) {
        const int half = L >> 1;
        for (int i = t; i < (N >> 1); i += blockDim.x) {
            const int p = i / m, q = i % m;                            // p < half
            float s, c;
            sincospif(-2.0f * (float)p / (float)L, &s, &c);
            const float2 u = a[q + m * p], v = a[q + m * (p + half)];
            b[q + m * (2 * p)] = make_float2(u.x + v.x, u.y + v.y);
            const float2 d = make_float2(u.x - v.x, u.y - v.y);
            b[q + m * (2 * p + 1)] = make_float2(d.x * c - d.y * s, d.x * s + d.y * c);
        }
        __syncthreads();
        float2* tmp = a; a = b; b = tmp;
    }
    const int nf = (N >> 1) + 1;
    for (int k = t; k < nf; k += blockDim.x) {
        const float2 X = a[k];
        const float onesided = (k == 0 || k == (N >> 1)) ? 1.f : 2.f;
        out[(i64)k * nseg + j] = (X.x * X.x + X.y * X.y) * scale * onesided;
    }
}

void spectrogram_psd(const float* d_x, i64 n, int stride, double rate, int nperseg, float* d_out, int* nseg_out) {
    Ctx& c = ctx();
    int logN = 0;
    while ((1 << logN) < nperseg) ++logN;
    ARS_CHECK(nperseg >= 2 && (1 << logN) == nperseg && nperseg <= 8192, "spectrogram: nperseg must be a power of two in 2..8192");
    ARS_CHECK(n >= nperseg, "spectrogram: signal shorter than one segment");
    const int hop = nperseg - nperseg / 2;
    const int nseg = (int)((n - nperseg / 2) / hop);                  // scipy: (n - noverlap) // (nperseg - noverlap)
    double wsum = 0.0;
    for (int i = 0; i < nperseg; ++i) {
        const float w = 0.5f - 0.5f * (float)std::cos(2.0 * 3.14159265358979323846 * i / nperseg);
        wsum += (double)w * (double)w;
    }
    const float scale = (float)(1.0 / (rate * wsum));
    const size_t smem = sizeof(float2) * 2 * (size_t)nperseg;
    static unsigned long long attr_gen = 0;
    if (attr_gen != ctx_generation()) {
        ARS_CUDA(cudaFuncSetAttribute(spectrogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float2) * 2 * 8192)));
        attr_gen = ctx_generation();
    }
    spectrogram_kernel<<<nseg, 256, smem, c.stream>>>(d_x, n, stride, logN, hop, nseg, scale, d_out);
    ARS_LAUNCH_CHECK();
    count_launch();
    if (nseg_out) *nseg_out = nseg;
}

}  // namespace ars
