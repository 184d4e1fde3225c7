// Device code of the final pass that more than one kernel uses (epilogue.cu: final_kernel; metrics.cu: the loudness
// meter that carries the final pass in its feed): block reductions, int16 packing, the lean frame math of the 5.1-based
// layouts -- packed guard division, products with the float64 gains from float32 pieces -- and the out-of-line literal frame.
#pragma once
#include "tail_math.cuh"

namespace ars {

__device__ __forceinline__ unsigned warp_max(unsigned m) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
}
__device__ __forceinline__ void block_atomic_max(unsigned m, unsigned* dst) {
    __shared__ unsigned s_m[32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x + 31) / 32) ? s_m[threadIdx.x] : 0u;
        m = warp_max(m);
        if (threadIdx.x == 0 && m > *reinterpret_cast<volatile unsigned*>(dst)) atomicMax(dst, m);
    }
    __syncthreads();      // the staging array is reused by the next reduction of the same block
}
__device__ __forceinline__ void block_atomic_add(double v, double* dst) {
    __shared__ double s_v[32];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x + 31) / 32) ? s_v[threadIdx.x] : 0.0;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(dst, v);
    }
    __syncthreads();
}

__device__ __forceinline__ short pcm_of(float v) {
    // np.clip(+-0.9999) in float32, NaN -> 0, then lrintf(x * 32767.0f)   (rs.py:1082-1084).
    // Evaluated as clamp(rint(x * 32767), +-32764): rounding is monotone and rint(0.9999f * 32767) = 32764, so
    // clamping after the conversion gives the same integers; cvt.rni maps NaN to 0 and +-inf to the clamp.
    const int q = __float2int_rn(__fmul_rn(v, 32767.0f));
    return (short)min(max(q, -32764), 32764);
}
// two samples at once: saturating pack to int16 x 2, then one two-lane max and one two-lane min for the +-32764 clamp
__device__ __forceinline__ unsigned pcm_pair(float v0, float v1) {
    const int q0 = __float2int_rn(__fmul_rn(v0, 32767.0f)), q1 = __float2int_rn(__fmul_rn(v1, 32767.0f));
    unsigned p;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(p) : "r"(q1), "r"(q0));      // q1 -> upper half, q0 -> lower half
    return __vmins2(__vmaxs2(p, 0x80048004u), 0x7ffc7ffcu);
}

// The frame loop of the final pass; A1 / A2 / A3 say which peak guards are active (checked once per thread, so
// idle guards cost nothing per sample).
// ---- lean form of the frame loop for the 5.1-based layouts when only the stereo guard can be active (the pan guard
// and the map guard idle: every render whose pan gains keep the six channels <= 1).  Same per-sample arithmetic as
// final_body, fewer issue slots: the stereo guard's division runs on the loaded (L, R) pairs with packed FP32x2
// instructions behind ONE range test per frame (the element-wise form tests and branches per value), 32-bit frame
// offsets inside the window, the delayed pair's presence as a loop split instead of a per-frame test, packed PCM scaling.
__device__ __forceinline__ float2 guard_div2(float2 v, const Guard& g) {
    const float2 r2 = make_float2(g.r, g.r), nm = make_float2(-g.m, -g.m);
    float2 q = __fmul2_rn(v, r2);                                  // the two corrections of guard_div, two lanes at once
    q = __ffma2_rn(__ffma2_rn(q, nm, v), r2, q);
    return __ffma2_rn(__ffma2_rn(q, nm, v), r2, q);
}
// every one of the four values zero or inside guard_div's plain range [1e-25, 1e30)
__device__ __forceinline__ bool plain4(float2 v, float2 w) {
    constexpr unsigned LO = 0x15f79688u /* 1e-25f */, HI = 0x7149f2cau /* 1e30f */;
    const unsigned a = abs_bits(v.x), b = abs_bits(v.y), c = abs_bits(w.x), d = abs_bits(w.y);
    const unsigned hi = max(max(a, b), max(c, d));
    const unsigned lo = min(min(a - 1u, b - 1u), min(c - 1u, d - 1u));      // (zero -> 0xffffffff: no lower bound)
    return hi < HI && lo >= LO - 1u;
}
static __device__ __noinline__ float4 guard_slow4(float4 q, float m, float r) {       // (values in and out in registers)
    Guard g;
    g.mode = 1; g.m = m; g.r = r;
    return make_float4(guard_div(q.x, g), guard_div(q.y, g), guard_div(q.z, g), guard_div(q.w, g));
}

// Threads per CTA of the lean loops.  Small CTAs: the warps of a CTA drift apart over its ~50 frames per thread and the
// CTA keeps its registers until the last one is through the closing reduction (ncu, 256 threads: 12 % of the warp time
// waiting at that barrier); 64 registers per thread either way, i.e. 1024 resident threads per SM.
#ifndef ARS_FINAL_NT
#define ARS_FINAL_NT 128
#endif
constexpr int FINAL_NT = ARS_FINAL_NT;
struct LeanAcc { float pkf; unsigned mm; double ss; };

// one frame's channels: LAY = 1 (5.1), 2 (7.1), 3 (5.1.2); A1 = 0 (stereo guard idle), 2 (divides, reciprocal form valid)
// or 3 (divides, element-wise form)
template <int C, int LAY, int A1>
__device__ __forceinline__ void lean_math(float2 v, float2 w, const TailSpec& ts, const Guard& g1, float (&o)[8]) {
    if constexpr (A1 == 2) {
        if (plain4(v, w)) {
            v = guard_div2(v, g1);
            if constexpr (LAY >= 2) w = guard_div2(w, g1);
        } else {
            const float4 q = guard_slow4(make_float4(v.x, v.y, w.x, w.y), g1.m, g1.r);
            v = make_float2(q.x, q.y);
            w = make_float2(q.z, q.w);
        }
    } else if constexpr (A1 == 3) {       // (the redo of a frame the float32 form could not decide)
        const float4 q = guard_slow4(make_float4(v.x, v.y, w.x, w.y), g1.m, g1.r);
        v = make_float2(q.x, q.y);
        w = make_float2(q.z, q.w);
    }
    const float mn = __fmul_rn(__fadd_rn(v.x, v.y), 0.707f);
    const double dl = (double)v.x, dr = (double)v.y;
    o[0] = __double2float_rn(__dmul_rn(dl, ts.g_fl));
    o[1] = __double2float_rn(__dmul_rn(dr, ts.g_fr));
    o[2] = __double2float_rn(__dmul_rn((double)mn, ts.g_c));
    o[3] = __fmul_rn(mn, ts.g_lfe);
    o[4] = __double2float_rn(__dmul_rn(dl, ts.g_rl));
    o[5] = __double2float_rn(__dmul_rn(dr, ts.g_rr));
    if constexpr (LAY >= 2) {
        // (a frame without a delayed partner carries w = 0: its pair is 0 * gain = 0, as the zero-prepended copy)
        const float rl = __double2float_rn(__dmul_rn((double)w.x, ts.g_rl));
        const float rr = __double2float_rn(__dmul_rn((double)w.y, ts.g_rr));
        if constexpr (LAY == 2) {
            o[6] = __fmul_rn(rl, 0.7f);
            o[7] = __fmul_rn(rr, 0.7f);
        } else {
            o[6] = __double2float_rn(__dmul_rn((double)rl, ts.height_gain));
            o[7] = __double2float_rn(__dmul_rn((double)rr, ts.height_gain));
        }
    }
}

// ---- RN32(RN64(x * g)) without float64 conversions --------------------------------------------------------------------
// numpy forms `audio * gain` with an np.float64 gain in float64 and rounds it into the float32 array (rs.py:475-494,
// 550-553).  Evaluated literally that is two conversions and a float64 multiply per product -- 17 conversions per 5.1.2
// frame, and the conversion pipe is what bounded the final pass (ncu: 109 us with 125 instructions per frame, the same
// as with 200).  The same value from float32 pieces, two lanes per instruction:
//     p = RN(x gh)             gh = gain rounded toward zero, gl = RN32(gain - gh) >= 0  (host: split_gain)
//     q = fma(x, -gh, p)       = -(x gh - p), exact
//     e = fma(x, gl, -q)       = (x gh - p) + x gl, one rounding; |e| < 3 ulp(p)
//     r = RN(p + e)            the candidate;  rho = (p - r) + e is the exact residual of that sum
// x g = p + e + eta with |eta| < 2^-21 ulp(p) (rounding of e, the bits of g below gl, the float64 rounding), so r is the
// wanted value unless p + e sits that close to a rounding boundary -- tested as fma(rho, 1 + 2^-16, r) != r, which fires
// for 1.5e-5 of all products; such a frame (and any frame with an operand outside the plain range, NaN and infinities
// included) is redone through float64.  The signs of zero products come out as numpy's (that is what rounding the split
// toward zero is for).  tests/host_emul/prod_emul.c replays this on the CPU against the float64 evaluation (3e8 random
// and 5e7 adversarial operands: no unflagged difference).
__device__ __forceinline__ float2 prod2(float2 x, float2 gh, float2 gl, bool& bad) {
    const float2 p = __fmul2_rn(x, gh);
    const float2 q = __ffma2_rn(x, make_float2(-gh.x, -gh.y), p);
    const float2 e = __ffma2_rn(x, gl, make_float2(-q.x, -q.y));
    // (sums of a packed product go through fma(a, 1, b): ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2,
    // which the scalar .rn forms never are -- seen on the PCM scaling below, where it cost the second rounding)
    const float2 one = make_float2(1.f, 1.f);
    const float2 r = __ffma2_rn(p, one, e);
    const float2 rho = __fadd2_rn(__ffma2_rn(r, make_float2(-1.f, -1.f), p), e);
    const float2 c = __ffma2_rn(rho, make_float2(1.0000152587890625f, 1.0000152587890625f), r);
    bad |= (c.x != r.x) | (c.y != r.y);
    return r;
}
__device__ __forceinline__ float prod1(float x, float gh, float gl, bool& bad) {
    const float p = __fmul_rn(x, gh);
    const float q = __fmaf_rn(x, -gh, p);
    const float e = __fmaf_rn(x, gl, -q);
    const float r = __fadd_rn(p, e);
    const float rho = __fadd_rn(__fsub_rn(p, r), e);
    bad |= (__fmaf_rn(rho, 1.0000152587890625f, r) != r);
    return r;
}
// every one of the four values zero or inside [2^-26, 2^40): guard_div2's plain range, and after a division by a maximum
// in (1, 1e10) still >= 2^-60, where every product of the frame and its error term stay normal numbers
__device__ __forceinline__ bool plain4e(float2 v, float2 w) {
    constexpr unsigned LO = 0x32800000u /* 2^-26 */, HI = 0x53800000u /* 2^40 */;
    const unsigned a = abs_bits(v.x), b = abs_bits(v.y), c = abs_bits(w.x), d = abs_bits(w.y);
    const unsigned hi = max(max(a, b), max(c, d));
    const unsigned lo = min(min(a - 1u, b - 1u), min(c - 1u, d - 1u));
    return hi < HI && lo >= LO - 1u;
}

// one frame's channels from float32 pieces of the gains; -> true when the frame has to be redone literally
template <int C, int LAY, int A1>
__device__ __forceinline__ bool lean_math_split(float2 v, float2 w, const TailSpec& ts, const Guard& g1, float (&o)[8]) {
    bool bad = !plain4e(v, w);
    float2 gv = v, gw = w;
    if constexpr (A1 == 2) {
        gv = guard_div2(v, g1);
        if constexpr (LAY >= 2) gw = guard_div2(w, g1);
    }
    const float mn = __fmul_rn(__fadd_rn(gv.x, gv.y), 0.707f);
    bad |= (abs_bits(mn) - 1u) < (0x21800000u /* 2^-60 */ - 1u);       // (L + R may cancel to something tiny)
    const float2 f = prod2(gv, make_float2(ts.g_hi[0], ts.g_hi[1]), make_float2(ts.g_lo[0], ts.g_lo[1]), bad);
    const float2 r = prod2(gv, make_float2(ts.g_hi[3], ts.g_hi[4]), make_float2(ts.g_lo[3], ts.g_lo[4]), bad);
    o[0] = f.x; o[1] = f.y;
    o[2] = prod1(mn, ts.g_hi[2], ts.g_lo[2], bad);
    o[3] = __fmul_rn(mn, ts.g_lfe);
    o[4] = r.x; o[5] = r.y;
    if constexpr (LAY >= 2) {
        const float2 d = prod2(gw, make_float2(ts.g_hi[3], ts.g_hi[4]), make_float2(ts.g_lo[3], ts.g_lo[4]), bad);
        float2 h;
        if constexpr (LAY == 2) h = __fmul2_rn(d, make_float2(0.7f, 0.7f));
        else h = prod2(d, make_float2(ts.g_hi[5], ts.g_hi[5]), make_float2(ts.g_lo[5], ts.g_lo[5]), bad);
        o[6] = h.x; o[7] = h.y;
    }
    return bad;
}

// The literal frame for the few the float32 form cannot decide (and every frame holding a NaN): the general code --
// loads, float64 products, stores and all -- out of line, so that the loop around it stays small.
// -> (max |channel|, the frame's sum of squares, its loudness-feed value)
template <int C>
__device__ __noinline__ float4 slow_frame(const float2* __restrict__ y, i64 i, const TailSpec* __restrict__ tsp, unsigned stereo_bits,
                                          float* __restrict__ out, short* __restrict__ pcm, float* __restrict__ mono) {
    const TailSpec& ts = *tsp;
    const Guard g1 = make_guard(stereo_bits), idle = make_guard(0u);
    float o[8];
    frame_out<1, 0>(y, i, ts, g1, idle, o);
    float pk = 0.f, fs = 0.f;
    #pragma unroll
    for (int c = 0; c < C; ++c) {
        pk = fmaxf(pk, fabsf(o[c]));
        fs = __fmaf_rn(o[c], o[c], fs);
    }
    if (out) {
        float* p = out + (i - ts.out0) * C;
        #pragma unroll
        for (int c = 0; c < C; c += 2) reinterpret_cast<float2*>(p)[c >> 1] = make_float2(o[c], o[c + 1]);
    }
    if (pcm) {
        short* p = pcm + (i - ts.out0) * C;
        #pragma unroll
        for (int c = 0; c < C; c += 2) reinterpret_cast<unsigned*>(p)[c >> 1] = pcm_pair(o[c], o[c + 1]);
    }
    const float mv = __fmul_rn(__fadd_rn(o[0], o[1]), 0.5f);
    if (mono) mono[i - ts.out0] = mv;
    return make_float4(pk, fs, mv, 0.f);
}

// peak / squares / stores of a frame of the float32 form.  Its channels are finite and <= 1 in magnitude (the pan guard is
// idle and NaN frames went the other way), so round-to-nearest-even of x * 32767 is the low half of RN(x * 32767 + 1.5 * 2^23)
// -- a packed add and a byte permute per pair instead of two conversions; the +-32764 clamp (np.clip at +-0.9999) follows on
// the packed int16 pair as in pcm_pair.
template <int C, int IO>
__device__ __forceinline__ void split_emit(const float (&o)[8], unsigned k, float* __restrict__ out, short* __restrict__ pcm,
                                           float* __restrict__ mono, LeanAcc& acc) {
    float fs = 0.f;
    #pragma unroll
    for (int c = 0; c < C; ++c) {
        acc.pkf = fmaxf(acc.pkf, fabsf(o[c]));
        fs = __fmaf_rn(o[c], o[c], fs);
    }
    acc.ss += (double)fs;
    if (IO == 0 && out) {
        float2* p = reinterpret_cast<float2*>(out) + (size_t)k * (C / 2);
        #pragma unroll
        for (int c = 0; c < C; c += 2) p[c >> 1] = make_float2(o[c], o[c + 1]);
    }
    if (IO == 1 || pcm) {
        unsigned u[C / 2];
        #pragma unroll
        for (int c = 0; c < C; c += 2) {
            // (two roundings, as lrintf(x * 32767.0f) has: the sum as fma(v, 1, magic) keeps ptxas from contracting them)
            const float2 t = __ffma2_rn(__fmul2_rn(make_float2(o[c], o[c + 1]), make_float2(32767.0f, 32767.0f)),
                                        make_float2(1.f, 1.f), make_float2(12582912.0f, 12582912.0f));
            const unsigned pk = __byte_perm(__float_as_uint(t.x), __float_as_uint(t.y), 0x5410);
            u[c >> 1] = __vmins2(__vmaxs2(pk, 0x80048004u), 0x7ffc7ffcu);
        }
        if constexpr (C == 8) {
            __stcs(reinterpret_cast<uint4*>(pcm) + k, make_uint4(u[0], u[1], u[2], u[3]));
        } else {
            unsigned* p = reinterpret_cast<unsigned*>(pcm) + (size_t)k * (C / 2);
            #pragma unroll
            for (int c = 0; c < C / 2; ++c) __stcs(p + c, u[c]);
        }
    }
    if (IO == 1 || mono) {
        const float mv = __fmul_rn(__fadd_rn(o[0], o[1]), 0.5f);
        mono[k] = mv;
        acc.mm = max(acc.mm, abs_bits(mv));
    }
}

}  // namespace ars
