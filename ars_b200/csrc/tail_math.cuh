// Per-frame arithmetic of the render tail, shared by the epilogue kernels (epilogue.cu) and the loudness meter's feed
// (metrics.cu): peak guards, 3-D pan, layout map.  Every operation replays numpy's operation order and promotion
// (rs.py:402-404, 475-499, 532-560) with explicit round-to-nearest intrinsics.
#pragma once
#include "epilogue.cuh"

namespace ars {

// ----------------------------------------------------------- frame math ------
struct Guard {              // 0: leave, 1: divide by m, 2: flush to zero
    int mode;
    float m;
    float r;                // mode 1: correctly rounded 1 / m, when the short division below is valid (else 0)
};

__device__ __forceinline__ Guard make_guard(unsigned bits) {
    // rs.py:402-404 / 497-499 / 558-560: max > 1 -> x / max ; any(x) and max < 1e-9 -> zeros
    const float m = __uint_as_float(bits);
    Guard g;
    g.m = m;
    g.mode = (m > 1.0f) ? 1 : ((m > 0.f && m < 1e-9f) ? 2 : 0);
    // (a divisor whose mantissa is all ones is the one case the theorem below excludes; huge divisors could underflow)
    g.r = (g.mode == 1 && m < 1e10f && (bits & 0x007fffffu) != 0x007fffffu) ? __frcp_rn(m) : 0.f;
    return g;
}
// v / m, correctly rounded like numpy's float32 division.  Thousands of samples are divided by the same maximum, so
// the reciprocal is formed once: q = RN(v r), then two residual corrections q += RN(v - q m) r in fused arithmetic.
// With r the correctly rounded reciprocal and q within an ulp of the quotient the corrected value IS the correctly
// rounded quotient (Markstein); checked against exact rational arithmetic on 3e5 random and adversarial pairs
// (scratch-free restatement in tests/test_host_logic.py).  Five straight-line instructions instead of the generic
// division's reciprocal approximation, Newton steps, range check and branch.  Values outside the plain range take
// the generic division.
__device__ __forceinline__ float guard_div(float v, const Guard& g) {
    // plain range 1e-25 <= |v| < 1e30 (or zero) as one unsigned compare on the bit pattern
    constexpr unsigned LO = 0x15f79688u /* 1e-25f */, HI = 0x7149f2cau /* 1e30f */;
    const unsigned u = __float_as_uint(v) & 0x7fffffffu;
    if (g.r != 0.f && (((u - LO) < (HI - LO)) | (u == 0u))) {
        float q = __fmul_rn(v, g.r);
        q = __fmaf_rn(__fmaf_rn(-q, g.m, v), g.r, q);
        return __fmaf_rn(__fmaf_rn(-q, g.m, v), g.r, q);
    }
    return __fdiv_rn(v, g.m);
}
__device__ __forceinline__ float guard1(float v, const Guard& g) {
    return g.mode == 0 ? v : (g.mode == 1 ? guard_div(v, g) : 0.f);
}

__device__ __forceinline__ void pan6(float L, float R, const TailSpec& ts, float (&o)[6]) {
    // rs.py:484-494: the mono mix and the LFE use Python-float (weak) gains => float32 multiplies; the
    // position gains are np.float64 => float64 product, rounded on the store into the float32 array
    const float mono = __fmul_rn(__fadd_rn(L, R), 0.707f);
    o[0] = __double2float_rn(__dmul_rn((double)L, ts.g_fl));
    o[1] = __double2float_rn(__dmul_rn((double)R, ts.g_fr));
    o[2] = __double2float_rn(__dmul_rn((double)mono, ts.g_c));
    o[3] = __fmul_rn(mono, ts.g_lfe);
    o[4] = __double2float_rn(__dmul_rn((double)L, ts.g_rl));
    o[5] = __double2float_rn(__dmul_rn((double)R, ts.g_rr));
}

// out channels of one frame from its (guarded) six channels and the (guarded) rear pair d frames earlier
__device__ __forceinline__ void map_frame(const float (&s)[6], float rl_d, float rr_d, const TailSpec& ts,
                                          float (&o)[8]) {
    if (ts.layout == LAYOUT_STEREO) {           // rs.py:533-535
        o[0] = __fadd_rn(__fadd_rn(s[0], __fmul_rn(s[2], 0.707f)), __fmul_rn(s[4], 0.5f));
        o[1] = __fadd_rn(__fadd_rn(s[1], __fmul_rn(s[2], 0.707f)), __fmul_rn(s[5], 0.5f));
        return;
    }
    #pragma unroll
    for (int c = 0; c < 6; ++c) o[c] = s[c];
    if (ts.layout == LAYOUT_7_1) {              // rs.py:541-545
        o[6] = __fmul_rn(rl_d, 0.7f);
        o[7] = __fmul_rn(rr_d, 0.7f);
    } else if (ts.layout == LAYOUT_5_1_2) {     // rs.py:548-554: float64 product, rounded on store
        o[6] = __double2float_rn(__dmul_rn((double)rl_d, ts.height_gain));
        o[7] = __double2float_rn(__dmul_rn((double)rr_d, ts.height_gain));
    }
}

// A = 0 compiles the guard away (the caller has checked that its mode is 0); A = 2: the caller has checked that the guard
// divides (mode 1) -- no per-sample mode test; A = 1: any mode
template <int A> __device__ __forceinline__ float guardT(float v, const Guard& g) {
    if constexpr (A == 1) return guard1(v, g);
    else if constexpr (A == 2) return guard_div(v, g);
    else return v;
}

// the frame's two loads: the stereo frame itself and, for the layouts with a delayed pair, the frame `delay` earlier
struct FrameIn { float2 v, w; };
__device__ __forceinline__ FrameIn frame_load(const float2* __restrict__ y, i64 i, const TailSpec& ts) {
    FrameIn f;
    f.v = (ts.stream & 2) ? __ldcs(y + (i - ts.y0)) : __ldg(y + (i - ts.y0));
    f.w = make_float2(0.f, 0.f);
    // a delay <= 0 leaves the signal where it is (rs.py:510-511)
    if (ts.layout >= LAYOUT_7_1 && i >= ts.delay) f.w = __ldg(y + (i - (ts.delay > 0 ? ts.delay : 0) - ts.y0));
    return f;
}

template <int A1 = 1, int A2 = 1>
__device__ __forceinline__ void frame_math(const FrameIn& f, i64 i, const TailSpec& ts, const Guard& g1, const Guard& g2,
                                           float (&o)[8]) {
    float s[6];
    pan6(guardT<A1>(f.v.x, g1), guardT<A1>(f.v.y, g1), ts, s);
    #pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = guardT<A2>(s[c], g2);
    float rl_d = 0.f, rr_d = 0.f;
    if (ts.layout >= LAYOUT_7_1 && i >= ts.delay) {
        rl_d = guardT<A2>(__double2float_rn(__dmul_rn((double)guardT<A1>(f.w.x, g1), ts.g_rl)), g2);
        rr_d = guardT<A2>(__double2float_rn(__dmul_rn((double)guardT<A1>(f.w.y, g1), ts.g_rr)), g2);
    }
    map_frame(s, rl_d, rr_d, ts, o);
}

template <int A1 = 1, int A2 = 1>
__device__ __forceinline__ void frame_out(const float2* __restrict__ y, i64 i, const TailSpec& ts, const Guard& g1,
                                          const Guard& g2, float (&o)[8]) {
    frame_math<A1, A2>(frame_load(y, i, ts), i, ts, g1, g2, o);
}

}  // namespace ars
