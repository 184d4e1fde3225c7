// Per-frame arithmetic of the render tail, shared by the epilogue kernels (epilogue.cu) and the loudness meter's feed
// (metrics.cu): peak guards, 3-D pan, layout map.  Every operation replays numpy's operation order and promotion
// (rs.py:402-404, 475-499, 532-560) with explicit round-to-nearest intrinsics.
#pragma once
#include "epilogue.cuh"

namespace ars {

// ----------------------------------------------------------- frame math ------
struct Guard { int mode; float m; };      // 0: leave, 1: divide by m, 2: flush to zero

__device__ __forceinline__ Guard make_guard(unsigned bits) {
    // rs.py:402-404 / 497-499 / 558-560: max > 1 -> x / max ; any(x) and max < 1e-9 -> zeros
    const float m = __uint_as_float(bits);
    Guard g;
    g.m = m;
    g.mode = (m > 1.0f) ? 1 : ((m > 0.f && m < 1e-9f) ? 2 : 0);
    return g;
}
__device__ __forceinline__ float guard1(float v, const Guard& g) {
    return g.mode == 0 ? v : (g.mode == 1 ? __fdiv_rn(v, g.m) : 0.f);
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

// A = false compiles the guard away (the caller has checked that its mode is 0)
template <bool A> __device__ __forceinline__ float guardT(float v, const Guard& g) {
    if constexpr (A) return guard1(v, g);
    else return v;
}

// the frame's two loads: the stereo frame itself and, for the layouts with a delayed pair, the frame `delay` earlier
struct FrameIn { float2 v, w; };
__device__ __forceinline__ FrameIn frame_load(const float2* __restrict__ y, i64 i, const TailSpec& ts) {
    FrameIn f;
    f.v = __ldg(y + (i - ts.y0));
    f.w = make_float2(0.f, 0.f);
    // a delay <= 0 leaves the signal where it is (rs.py:510-511)
    if (ts.layout >= LAYOUT_7_1 && i >= ts.delay) f.w = __ldg(y + (i - (ts.delay > 0 ? ts.delay : 0) - ts.y0));
    return f;
}

template <bool A1 = true, bool A2 = true>
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

template <bool A1 = true, bool A2 = true>
__device__ __forceinline__ void frame_out(const float2* __restrict__ y, i64 i, const TailSpec& ts, const Guard& g1,
                                          const Guard& g2, float (&o)[8]) {
    frame_math<A1, A2>(frame_load(y, i, ts), i, ts, g1, g2, o);
}

}  // namespace ars
