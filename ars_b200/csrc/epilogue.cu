// Epilogue kernels (see epilogue.cuh).  All of them are one-touch HBM streams: grid-stride,
// grid = a multiple of the SM count, vector loads/stores, block reduce + one atomic per block.
#include "epilogue.cuh"
#include "final_math.cuh"

#include <cstdlib>

namespace ars {

static inline int stream_grid(i64 items, int per_block = 256) {
    const i64 need = (items + per_block - 1) / per_block;
    const i64 cap = (i64)ctx().sm_count * 8;
    return (int)std::max<i64>(1, std::min(need, cap));
}

// ------------------------------------------------------- fused tail kernels --
// max |six| before its guard.  When the stereo peak guard is idle (the common case) the maximum follows exactly
// from three maxima the last FFT pass already tracked -- every pan channel is a monotone function of |L|, |R| or
// |float32(L + R)| (non-negative gains, monotone rounding) -- so no pass over the signal is needed.
__global__ void __launch_bounds__(256) pan_max_kernel(const float2* __restrict__ y, TailSpec ts, RenderState* st) {
    const Guard g1 = make_guard(st->max_stereo);
    if (g1.mode == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            float s[6];
            const float l = __uint_as_float(st->max_l), r = __uint_as_float(st->max_r);
            const float mono = __fmul_rn(__uint_as_float(st->max_lr), 0.707f);
            s[0] = __double2float_rn(__dmul_rn((double)l, ts.g_fl));
            s[1] = __double2float_rn(__dmul_rn((double)r, ts.g_fr));
            s[2] = __double2float_rn(__dmul_rn((double)mono, ts.g_c));
            s[3] = __fmul_rn(mono, ts.g_lfe);
            s[4] = __double2float_rn(__dmul_rn((double)l, ts.g_rl));
            s[5] = __double2float_rn(__dmul_rn((double)r, ts.g_rr));
            unsigned m = 0;
            #pragma unroll
            for (int c = 0; c < 6; ++c) m = max(m, abs_bits(s[c]));
            st->max_pan = m;
        }
        return;
    }
    if (g1.mode == 1) {
        // Stereo guard active: |L'| and |R'| maxima still follow in closed form (x / m is monotone), only the mono
        // term max |float32(L' + R')| does not.  The pan maximum is used for nothing but its guard decision (> 1:
        // divide; < 1e-9: flush), so when an upper bound of the mono term keeps every channel <= 1 and an exact
        // L / R channel is >= 1e-9 the guard is provably idle and the pass over the signal is skipped; the exact
        // L / R maximum is stored as a stand-in inside (1e-9, 1].
        const float l = __fdiv_rn(__uint_as_float(st->max_l), g1.m), r = __fdiv_rn(__uint_as_float(st->max_r), g1.m);
        const float mono_ub = __fmul_rn(__fadd_rn(l, r), 0.707f);
        const float lr = fmaxf(fmaxf(__double2float_rn(__dmul_rn((double)l, ts.g_fl)), __double2float_rn(__dmul_rn((double)r, ts.g_fr))),
                               fmaxf(__double2float_rn(__dmul_rn((double)l, ts.g_rl)), __double2float_rn(__dmul_rn((double)r, ts.g_rr))));
        const float ub = fmaxf(fmaxf(fabsf(__double2float_rn(__dmul_rn((double)mono_ub, ts.g_c))), __fmul_rn(mono_ub, ts.g_lfe)), fabsf(lr));
        if (ub <= 1.0f && fabsf(lr) >= 1e-9f) {
            if (blockIdx.x == 0 && threadIdx.x == 0) st->max_pan = abs_bits(lr);
            return;
        }
    }
    unsigned m = 0;
    for (i64 i = ts.i_lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < ts.i_hi; i += (i64)gridDim.x * blockDim.x) {
        const float2 v = __ldg(y + (i - ts.y0));
        float s[6];
        pan6(guard1(v.x, g1), guard1(v.y, g1), ts, s);
        #pragma unroll
        for (int c = 0; c < 6; ++c) m = max(m, abs_bits(s[c]));
    }
    block_atomic_max(m, &st->max_pan);
}

__global__ void __launch_bounds__(256) map_max_kernel(const float2* __restrict__ y, TailSpec ts, RenderState* st) {
    const Guard g1 = make_guard(st->max_stereo), g2 = make_guard(st->max_pan);
    unsigned m = 0;
    for (i64 i = ts.i_lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < ts.i_hi; i += (i64)gridDim.x * blockDim.x) {
        float o[8];
        frame_out(y, i, ts, g1, g2, o);
        #pragma unroll
        for (int c = 0; c < 8; ++c) if (c < ts.C) m = max(m, abs_bits(o[c]));
    }
    block_atomic_max(m, &st->max_map);
}

template <int C, int A1, int A2, int A3>
__device__ __forceinline__ void final_body(const float2* __restrict__ y, const TailSpec& ts, const Guard& g1, const Guard& g2,
                                           const Guard& g3, float* __restrict__ out, short* __restrict__ pcm,
                                           float* __restrict__ mono, float& pkf, bool& nan_seen, unsigned& mm, double& ss) {
    // the loads of the next frame are issued before the arithmetic of this one: the pass is bound by load latency
    const i64 stride = (i64)gridDim.x * blockDim.x;
    i64 i = ts.i_lo + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    FrameIn cur = i < ts.i_hi ? frame_load(y, i, ts) : FrameIn();
    for (; i < ts.i_hi; i += stride) {
        const FrameIn f = cur;
        if (i + stride < ts.i_hi) cur = frame_load(y, i + stride, ts);
        float o[8];
        frame_math<A1, A2>(f, i, ts, g1, g2, o);
        float fs = 0.f;        // one frame's squares in float32 (numpy squares in float32 too), then one conversion
        #pragma unroll
        for (int c = 0; c < C; ++c) {
            o[c] = guardT<A3>(o[c], g3);
            pkf = fmaxf(pkf, fabsf(o[c]));
            fs = __fmaf_rn(o[c], o[c], fs);
        }
        nan_seen |= (fs != fs);                       // a NaN sample makes the frame's sum NaN (np.max would return NaN)
        ss += (double)fs;
        if (out) {
            float* p = out + (i - ts.out0) * C;
            if (C == 8) {
                reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
            } else {
                #pragma unroll
                for (int c = 0; c < C; c += 2) reinterpret_cast<float2*>(p)[c >> 1] = make_float2(o[c], o[c + 1]);
            }
        }
        if (pcm) {
            short* p = pcm + (i - ts.out0) * C;
            if (C == 8) {
                uint4 u;
                u.x = pcm_pair(o[0], o[1]);
                u.y = pcm_pair(o[2], o[3]);
                u.z = pcm_pair(o[4], o[5]);
                u.w = pcm_pair(o[6], o[7]);
                if (ts.stream & 1) __stcs(reinterpret_cast<uint4*>(p), u);
                else *reinterpret_cast<uint4*>(p) = u;
            } else {
                #pragma unroll
                for (int c = 0; c < C; c += 2) {
                    if (ts.stream & 1) __stcs(reinterpret_cast<unsigned*>(p) + (c >> 1), pcm_pair(o[c], o[c + 1]));
                    else reinterpret_cast<unsigned*>(p)[c >> 1] = pcm_pair(o[c], o[c + 1]);
                }
            }
        }
        if (mono) {
            const float mv = __fmul_rn(__fadd_rn(o[0], o[1]), 0.5f);    // np.mean(data[:, :2], axis=1) (rs.py:688); x / 2 == x * 0.5 exactly
            mono[i - ts.out0] = mv;
            mm = max(mm, abs_bits(mv));
        }
    }
}

// Frames [k_lo, k_hi) of the window (relative to ts.i_lo) through the float32 form.  A CTA takes tiles of 2 x FINAL_NT frames
// (a thread's two frames FINAL_NT apart: one address per array and step), the next step's loads issued before this step's
// arithmetic, the steps unrolled by two so that the register sets swap roles.  PAIR: every frame has its delayed partner.
template <int C, int LAY, int A1, int IO, bool PAIR>
__device__ __forceinline__ void split_range(const float2* __restrict__ yw, unsigned k_lo, unsigned k_hi, unsigned dl,
                                            const TailSpec* __restrict__ tsp, const Guard& g1, unsigned stereo_bits,
                                            const float2* __restrict__ y, float* __restrict__ out, short* __restrict__ pcm,
                                            float* __restrict__ mono, float* __restrict__ outw, short* __restrict__ pcmw,
                                            float* __restrict__ monow, LeanAcc& acc) {
    constexpr int U = 2;
    const TailSpec& ts = *tsp;
    const unsigned step = gridDim.x * ((unsigned)(FINAL_NT * U));
    unsigned k = k_lo + blockIdx.x * ((unsigned)(FINAL_NT * U)) + threadIdx.x;
    if (k >= k_hi) return;
    const float2* __restrict__ yd = yw - dl;
    float2 va[U], wa[U], vb[U], wb[U];
    auto fetch = [&](unsigned k0, float2 (&v)[U], float2 (&w)[U]) {
        const float2* pv = yw + k0;
        const float2* pw = yd + k0;
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u] = make_float2(0.f, 0.f);
            w[u] = make_float2(0.f, 0.f);
            if (k0 + (unsigned)(u * FINAL_NT) < k_hi) {
                v[u] = __ldcs(pv + u * FINAL_NT);
                if (LAY >= 2 && PAIR) w[u] = __ldg(pw + u * FINAL_NT);
            }
        }
    };
    auto work = [&](unsigned k0, const float2 (&v)[U], const float2 (&w)[U]) {
        float o[U][8];
        bool bad[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) bad[u] = lean_math_split<C, LAY, A1>(v[u], w[u], ts, g1, o[u]);
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned kk = k0 + (unsigned)(u * FINAL_NT);
            if (kk >= k_hi) continue;
            if (!bad[u]) {
                split_emit<C, IO>(o[u], kk, outw, pcmw, monow, acc);
            } else {
                const float4 r = slow_frame<C>(y, ts.i_lo + (i64)kk, tsp, stereo_bits, out, pcm, mono);
                acc.pkf = fmaxf(acc.pkf, r.x);
                acc.ss += (double)r.y;                       // (a NaN frame leaves its mark here)
                if (mono) acc.mm = max(acc.mm, abs_bits(r.z));
            }
        }
    };
    fetch(k, va, wa);
    for (;;) {
        const unsigned k1 = k + step;
        const bool more1 = k1 < k_hi && k1 > k;
        if (more1) fetch(k1, vb, wb);
        work(k, va, wa);
        if (!more1) break;
        const unsigned k2 = k1 + step;
        const bool more2 = k2 < k_hi && k2 > k1;
        if (more2) fetch(k2, va, wa);
        work(k1, vb, wb);
        if (!more2) break;
        k = k2;
    }
}

template <int C, int LAY, int A1, int IO>
__device__ __forceinline__ void final_split(const float2* __restrict__ y, const TailSpec* __restrict__ tsp, const Guard& g1,
                                            unsigned stereo_bits, float* __restrict__ out, short* __restrict__ pcm,
                                            float* __restrict__ mono, LeanAcc& acc) {
    const TailSpec& ts = *tsp;
    const float2* yw = y + (ts.i_lo - ts.y0);                      // frame 0 of the window
    const unsigned count = (unsigned)(ts.i_hi - ts.i_lo);
    float* outw = out ? out + (ts.i_lo - ts.out0) * C : nullptr;
    short* pcmw = pcm ? pcm + (ts.i_lo - ts.out0) * C : nullptr;
    float* monow = mono ? mono + (ts.i_lo - ts.out0) : nullptr;
    if constexpr (LAY >= 2) {
        const i64 dl = ts.delay > 0 ? ts.delay : 0;
        const i64 head = ts.delay - ts.i_lo;
        const unsigned k_head = head <= 0 ? 0u : (head >= (i64)count ? count : (unsigned)head);
        split_range<C, LAY, A1, IO, false>(yw, 0u, k_head, 0u, tsp, g1, stereo_bits, y, out, pcm, mono, outw, pcmw, monow, acc);
        split_range<C, LAY, A1, IO, true>(yw, k_head, count, (unsigned)dl, tsp, g1, stereo_bits, y, out, pcm, mono, outw, pcmw, monow, acc);
    } else {
        split_range<C, LAY, A1, IO, false>(yw, 0u, count, 0u, tsp, g1, stereo_bits, y, out, pcm, mono, outw, pcmw, monow, acc);
    }
}

// peak / squares / stores of frame k (byte offsets inside the window fit 32 bits: the caller checks the frame count).
// A NaN sample makes the frame's sum of squares NaN and with it the running double sum, for good: the caller reads
// "NaN seen" off that sum instead of testing every frame.
template <int C, int IO>          // IO = 1: PCM + loudness feed, no float frames (the render's default); 0: whatever is given
__device__ __forceinline__ void lean_emit(const float (&o)[8], unsigned k, bool valid, const TailSpec& ts, float* __restrict__ out,
                                          short* __restrict__ pcm, float* __restrict__ mono, LeanAcc& acc) {
    float fs = 0.f;
    #pragma unroll
    for (int c = 0; c < C; ++c) {
        acc.pkf = fmaxf(acc.pkf, fabsf(o[c]));
        fs = __fmaf_rn(o[c], o[c], fs);
    }
    acc.ss += (double)fs;
    if (!valid) return;           // (a frame past the end carries zeros: its sums change nothing, its stores are skipped)
    if (IO == 0 && out) {
        float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(out) + k * (unsigned)(4 * C));
        #pragma unroll
        for (int c = 0; c < C; c += 2) reinterpret_cast<float2*>(p)[c >> 1] = make_float2(o[c], o[c + 1]);
    }
    if (IO == 1 || pcm) {
        short* p = reinterpret_cast<short*>(reinterpret_cast<char*>(pcm) + k * (unsigned)(2 * C));
        const float2 sc = make_float2(32767.0f, 32767.0f);
        unsigned u[C / 2];
        #pragma unroll
        for (int c = 0; c < C; c += 2) {
            const float2 s2 = __fmul2_rn(make_float2(o[c], o[c + 1]), sc);
            const int q0 = __float2int_rn(s2.x), q1 = __float2int_rn(s2.y);
            unsigned pk;
            asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(pk) : "r"(q1), "r"(q0));
            u[c >> 1] = __vmins2(__vmaxs2(pk, 0x80048004u), 0x7ffc7ffcu);
        }
        if constexpr (C == 8) {
            __stcs(reinterpret_cast<uint4*>(p), make_uint4(u[0], u[1], u[2], u[3]));
        } else {
            #pragma unroll
            for (int c = 0; c < C / 2; ++c) __stcs(reinterpret_cast<unsigned*>(p) + c, u[c]);
        }
    }
    if (IO == 1 || mono) {
        const float mv = __fmul_rn(__fadd_rn(o[0], o[1]), 0.5f);
        *reinterpret_cast<float*>(reinterpret_cast<char*>(mono) + k * 4u) = mv;
        acc.mm = max(acc.mm, abs_bits(mv));
    }
}

// frames [k_lo, k_hi) of the window (k relative to ts.i_lo); PAIR: every frame has its delayed partner dl frames back.
// A thread takes frames k, k + stride, ...: U of them per step, the next step's loads issued before this step's
// arithmetic, the steps unrolled by two so that the two register sets swap roles instead of being copied.
template <int C, int LAY, int A1, int IO, int U, bool PIPE, bool PAIR>
__device__ __forceinline__ void lean_range(const float2* __restrict__ yw, unsigned k_lo, unsigned k_hi, unsigned dl,
                                           const TailSpec& ts, const Guard& g1, float* __restrict__ out,
                                           short* __restrict__ pcm, float* __restrict__ mono, LeanAcc& acc) {
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned k = k_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_hi) return;
    const char* __restrict__ yb = reinterpret_cast<const char*>(yw);
    const char* __restrict__ ydb = reinterpret_cast<const char*>(yw - dl);          // the partner `delay` frames earlier
    float2 va[U], wa[U], vb[U], wb[U];
    auto fetch = [&](unsigned k0, float2 (&v)[U], float2 (&w)[U]) {
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned kk = k0 + u * stride;
            v[u] = make_float2(0.f, 0.f);
            w[u] = make_float2(0.f, 0.f);
            if (kk < k_hi) {
                v[u] = __ldcs(reinterpret_cast<const float2*>(yb + kk * 8u));      // (touched once more, as a partner, through the L2)
                if (LAY >= 2 && PAIR) w[u] = __ldg(reinterpret_cast<const float2*>(ydb + kk * 8u));
            }
        }
    };
    auto work = [&](unsigned k0, const float2 (&v)[U], const float2 (&w)[U]) {
        float o[U][8];
        #pragma unroll
        for (int u = 0; u < U; ++u) lean_math<C, LAY, A1>(v[u], w[u], ts, g1, o[u]);
        #pragma unroll
        for (int u = 0; u < U; ++u) lean_emit<C, IO>(o[u], k0 + u * stride, k0 + u * stride < k_hi, ts, out, pcm, mono, acc);
    };
    if constexpr (!PIPE) {
        // U frames' loads in flight per thread, consumed in order (no second register set)
        for (; k < k_hi; k += U * stride) {
            fetch(k, va, wa);
            work(k, va, wa);
        }
        return;
    }
    fetch(k, va, wa);
    for (;;) {
        const unsigned k1 = k + U * stride;
        if (k1 < k_hi) fetch(k1, vb, wb);
        work(k, va, wa);
        if (k1 >= k_hi) break;
        const unsigned k2 = k1 + U * stride;
        if (k2 < k_hi) fetch(k2, va, wa);
        work(k1, vb, wb);
        if (k2 >= k_hi) break;
        k = k2;
    }
}

template <int C, int LAY, int A1, int IO>
__device__ __forceinline__ void final_lean(const float2* __restrict__ y, const TailSpec& ts, const Guard& g1,
                                           float* __restrict__ out, short* __restrict__ pcm, float* __restrict__ mono,
                                           LeanAcc& acc) {
    const float2* yw = y + (ts.i_lo - ts.y0);                      // frame 0 of the window
    const unsigned count = (unsigned)(ts.i_hi - ts.i_lo);
    float* outw = out ? out + (ts.i_lo - ts.out0) * C : nullptr;
    short* pcmw = pcm ? pcm + (ts.i_lo - ts.out0) * C : nullptr;
    float* monow = mono ? mono + (ts.i_lo - ts.out0) : nullptr;
    if constexpr (LAY >= 2) {
        // frames below `delay` have no partner (rs.py:507-515 prepends zeros); a delay <= 0 leaves the pair in place
        const i64 dl = ts.delay > 0 ? ts.delay : 0;
        const i64 head = ts.delay - ts.i_lo;
        const unsigned k_head = head <= 0 ? 0u : (head >= (i64)count ? count : (unsigned)head);
        lean_range<C, LAY, A1, IO, 1, false, false>(yw, 0u, k_head, 0u, ts, g1, outw, pcmw, monow, acc);
        lean_range<C, LAY, A1, IO, 2, true, true>(yw, k_head, count, (unsigned)dl, ts, g1, outw, pcmw, monow, acc);
    } else {
        lean_range<C, LAY, A1, IO, 2, true, false>(yw, 0u, count, 0u, ts, g1, outw, pcmw, monow, acc);
    }
}

// LAY = 0: any layout through the general loop; 1 / 2 / 3: 5.1 / 7.1 / 5.1.2 with the lean loop when it applies
template <int C, int LAY, bool SPLIT>
__global__ void __launch_bounds__(LAY ? FINAL_NT : 256, LAY ? 1024 / FINAL_NT : 4) final_kernel(const float2* __restrict__ y, const __grid_constant__ TailSpec ts, RenderState* st,
                                                    float* __restrict__ out, short* __restrict__ pcm,
                                                    float* __restrict__ mono) {
    const Guard g1 = make_guard(st->max_stereo), g2 = make_guard(st->max_pan);
    // 5.1 / 7.1 / 5.1.2 carry the six channels through unchanged and add only attenuated (x0.7, x<=0.6) delayed
    // copies, so their maximum is the guarded six-channel maximum (<= 1, or >= 1e-9): that guard never fires
    const Guard g3 = make_guard(ts.layout == LAYOUT_STEREO ? st->max_map : 0u);
    float pkf = 0.f;
    bool nan_seen = false;
    unsigned mm = 0;
    double ss = 0.0;
    bool done = false;
    if constexpr (LAY >= 1) {
        // (frame count below 2^28: byte offsets of every array fit 32 bits inside the window)
        if (g2.mode == 0 && (g1.mode == 0 || (g1.mode == 1 && g1.r != 0.f)) && ts.i_hi - ts.i_lo < ((i64)1 << 28)) {
            LeanAcc acc = {0.f, 0u, 0.0};
            const bool io1 = !out && pcm && mono;
            if (g1.mode == 0) {
                if (io1) { if constexpr (SPLIT) final_split<C, LAY, 0, 1>(y, &ts, g1, st->max_stereo, out, pcm, mono, acc); else final_lean<C, LAY, 0, 1>(y, ts, g1, out, pcm, mono, acc); }
                else { if constexpr (SPLIT) final_split<C, LAY, 0, 0>(y, &ts, g1, st->max_stereo, out, pcm, mono, acc); else final_lean<C, LAY, 0, 0>(y, ts, g1, out, pcm, mono, acc); }
            } else {
                if (io1) { if constexpr (SPLIT) final_split<C, LAY, 2, 1>(y, &ts, g1, st->max_stereo, out, pcm, mono, acc); else final_lean<C, LAY, 2, 1>(y, ts, g1, out, pcm, mono, acc); }
                else { if constexpr (SPLIT) final_split<C, LAY, 2, 0>(y, &ts, g1, st->max_stereo, out, pcm, mono, acc); else final_lean<C, LAY, 2, 0>(y, ts, g1, out, pcm, mono, acc); }
            }
            pkf = acc.pkf; nan_seen = (acc.ss != acc.ss); mm = acc.mm; ss = acc.ss;
            done = true;
        }
    }
    if (!done) {
        if (g2.mode == 0 && g3.mode == 0) {
            if (g1.mode == 0) final_body<C, 0, 0, 0>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
            else if (g1.mode == 1) final_body<C, 2, 0, 0>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
            else final_body<C, 1, 0, 0>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
        } else {
            final_body<C, 1, 1, 1>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
        }
    }
    const unsigned pk = nan_seen ? 0x7fc00000u : __float_as_uint(pkf);
    block_atomic_max(pk, &st->peak_final);
    block_atomic_max(mm, &st->mono_max);
    block_atomic_add(ss, &st->sumsq);
}

static int g_final_lean = 2;
void tail_set_lean(int v) { g_final_lean = v < 0 ? 0 : (v > 2 ? 2 : v); }

// hi = g rounded toward zero, lo = RN32(g - hi) >= 0; false when g is not +0 or inside [2^-16, 2^16] (prod2's conditions)
static bool split_gain(double g, float& hi, float& lo) {
    hi = lo = 0.f;
    if (g == 0.0) return !std::signbit(g);
    if (!(g >= 0x1p-16 && g <= 0x1p16)) return false;
    float f = (float)g;
    if ((double)f > g) f = std::nextafterf(f, 0.f);
    hi = f;
    lo = (float)(g - (double)f);
    return true;
}
void tail_prepare(TailSpec& ts) {
    const double g[6] = {ts.g_fl, ts.g_fr, ts.g_c, ts.g_rl, ts.g_rr, ts.layout == LAYOUT_5_1_2 ? ts.height_gain : 0.0};
    bool ok = true;
    for (int i = 0; i < 6; ++i) ok = split_gain(g[i], ts.g_hi[i], ts.g_lo[i]) && ok;
    ts.split_ok = ok ? 1 : 0;
}

static TailSpec with_window(const TailSpec& in) {
    TailSpec ts = in;
    if (ts.i_hi < 0) ts.i_hi = ts.N;
    return ts;
}

void tail_pan_max(const float2* d_y, const TailSpec& ts_in, RenderState* d_state) {
    const TailSpec ts = with_window(ts_in);
    if (ts.i_hi <= ts.i_lo) return;
    KernelScope prof("pan_max_kernel", 0.0);
    pan_max_kernel<<<stream_grid(ts.i_hi - ts.i_lo), 256, 0, ctx().stream>>>(d_y, ts, d_state);
    ARS_LAUNCH_CHECK();
    count_launch();
}

void tail_map_max(const float2* d_y, const TailSpec& ts_in, RenderState* d_state) {
    const TailSpec ts = with_window(ts_in);
    if (ts.i_hi <= ts.i_lo || ts.layout != LAYOUT_STEREO) return;
    KernelScope prof("map_max_kernel", 8.0 * (double)(ts.i_hi - ts.i_lo));
    map_max_kernel<<<stream_grid(ts.i_hi - ts.i_lo), 256, 0, ctx().stream>>>(d_y, ts, d_state);
    ARS_LAUNCH_CHECK();
    count_launch();
}

void tail_maxes(const float2* d_y, const TailSpec& ts_in, RenderState* d_state) {
    const TailSpec ts = with_window(ts_in);
    if (ts.N <= 0) return;
    Ctx& c = ctx();
    {
        KernelScope prof("pan_max_kernel", 0.0);
        pan_max_kernel<<<stream_grid(ts.i_hi - ts.i_lo), 256, 0, c.stream>>>(d_y, ts, d_state);
    }
    ARS_LAUNCH_CHECK();
    count_launch();
    if (ts.layout == LAYOUT_STEREO) {
        KernelScope prof("map_max_kernel", 8.0 * (double)(ts.i_hi - ts.i_lo));
        map_max_kernel<<<stream_grid(ts.i_hi - ts.i_lo), 256, 0, c.stream>>>(d_y, ts, d_state);
        ARS_LAUNCH_CHECK();
        count_launch();
    }
}

void tail_final(const float2* d_y, const TailSpec& ts_in, RenderState* d_state, float* d_out, short* d_pcm,
                float* d_mono) {
    TailSpec ts = with_window(ts_in);
    if (ts.N <= 0 || ts.i_hi <= ts.i_lo) return;
    tail_prepare(ts);
    Ctx& c = ctx();
    const int grid = stream_grid(ts.i_hi - ts.i_lo);
    static const int waves = [] { const char* e = getenv("ARS_FINAL_WAVES"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    const int grid_lean = (int)std::max<i64>(1, std::min<i64>((ts.i_hi - ts.i_lo + 2 * FINAL_NT - 1) / (2 * FINAL_NT),
                                                             (i64)c.sm_count * (1024 / FINAL_NT) * waves));
    KernelScope prof("final_kernel (guards, pan, map, clip, PCM16, sums)",
                     (double)(ts.i_hi - ts.i_lo) * (8.0 + (d_pcm ? 2.0 * ts.C : 0.0) + (d_out ? 4.0 * ts.C : 0.0) + (d_mono ? 4.0 : 0.0)));
    // 0: general loop; 1: lean loop, float64 products; 2: lean loop, products from float32 pieces (when the gains allow)
    const int lean = ts.layout != LAYOUT_STEREO ? g_final_lean : 0;
    #define ARS_FINAL(CC, LL, SS) final_kernel<CC, LL, SS><<<(LL) ? grid_lean : grid, (LL) ? FINAL_NT : 256, 0, c.stream>>>(d_y, ts, d_state, d_out, d_pcm, d_mono)
    #define ARS_FINAL_L(CC, LL) do { if (lean == 2 && ts.split_ok) ARS_FINAL(CC, LL, true); else if (lean) ARS_FINAL(CC, LL, false); \
                                     else ARS_FINAL(CC, 0, false); } while (0)
    if (ts.C == 2) ARS_FINAL(2, 0, false);
    else if (ts.C == 6) ARS_FINAL_L(6, 1);
    else if (ts.layout == LAYOUT_7_1) ARS_FINAL_L(8, 2);
    else ARS_FINAL_L(8, 3);
    #undef ARS_FINAL_L
    #undef ARS_FINAL
    ARS_LAUNCH_CHECK();
    count_launch();
}

// --------------------------------------------------------- stage kernels -----
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, i64 count, unsigned* dst) {
    unsigned m = 0;
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x, step = (i64)gridDim.x * blockDim.x;
    const i64 n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? count / 4 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (i64 i = tid; i < n4; i += step) {
        const float4 v = __ldg(x4 + i);
        m = max(max(m, abs_bits(v.x)), max(abs_bits(v.y), max(abs_bits(v.z), abs_bits(v.w))));
    }
    for (i64 i = n4 * 4 + tid; i < count; i += step) m = max(m, abs_bits(__ldg(x + i)));
    block_atomic_max(m, dst);
}
void absmax_f32(const float* d_x, i64 count, unsigned* d_maxbits) {
    if (count <= 0) return;
    absmax_kernel<<<stream_grid(count / 4 + 1), 256, 0, ctx().stream>>>(d_x, count, d_maxbits);
    ARS_LAUNCH_CHECK();
    count_launch();
}

__global__ void __launch_bounds__(256) guard_kernel(float* __restrict__ x, i64 count, const unsigned* maxbits) {
    const Guard g = make_guard(*maxbits);
    if (g.mode == 0) return;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x)
        x[i] = guard1(x[i], g);
}
void guard_apply(float* d_x, i64 count, const unsigned* d_maxbits) {
    if (count <= 0) return;
    guard_kernel<<<stream_grid(count), 256, 0, ctx().stream>>>(d_x, count, d_maxbits);
    ARS_LAUNCH_CHECK();
    count_launch();
}

// rs.py:113-121: (dmf*(1-dw)) * dry + dw * wet over the common length, the longer tail appended scaled;
// np.float64 scalars => float64 products and sum, rounded to float32 once.
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ dry, const float* __restrict__ wet, i64 count,
                                                  double dry_scale, double dw, float* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x)
        out[i] = __double2float_rn(__dadd_rn(__dmul_rn(dry_scale, (double)dry[i]), __dmul_rn(dw, (double)wet[i])));
}
__global__ void __launch_bounds__(256) mix_tail_kernel(const float* __restrict__ src, i64 lo, i64 hi, double s0, double s1,
                                                       int two, float* __restrict__ out) {
    for (i64 i = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (i64)gridDim.x * blockDim.x) {
        double v = __dmul_rn((double)src[i], s0);
        if (two) v = __dmul_rn(v, s1);
        out[i] = __double2float_rn(v);
    }
}
void mix_dry_wet(const float* d_dry, i64 n_dry, const float* d_wet, i64 n_wet, int ch, double dmf, double dw,
                 float* d_out) {
    // dmf = dry_mix_factor; the common part uses (dmf * (1 - dw)) as one float64 scalar (rs.py:113)
    const i64 c_dry = n_dry * ch, c_wet = n_wet * ch;
    const i64 common = std::min(c_dry, c_wet), total = std::max(c_dry, c_wet);
    if (total <= 0) return;
    Ctx& c = ctx();
    const double dry_scale = dmf * (1.0 - dw);
    if (common > 0) {
        mix_kernel<<<stream_grid(common), 256, 0, c.stream>>>(d_dry, d_wet, common, dry_scale, dw, d_out);
        ARS_LAUNCH_CHECK();
        count_launch();
    }
    if (c_dry > common) {          // rs.py:117: dry[min_len:] * dmf * (1.0 - dw)  (two successive float64 multiplies)
        mix_tail_kernel<<<stream_grid(c_dry - common), 256, 0, c.stream>>>(d_dry, common, c_dry, dmf, 1.0 - dw, 1, d_out);
        ARS_LAUNCH_CHECK();
        count_launch();
    } else if (c_wet > common) {   // rs.py:119: wet[min_len:] * dw
        mix_tail_kernel<<<stream_grid(c_wet - common), 256, 0, c.stream>>>(d_wet, common, c_wet, dw, 1.0, 0, d_out);
        ARS_LAUNCH_CHECK();
        count_launch();
    }
}

__global__ void __launch_bounds__(256) pan_kernel(const float2* __restrict__ s, TailSpec ts, float* __restrict__ six) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < ts.N; i += (i64)gridDim.x * blockDim.x) {
        const float2 v = __ldg(s + i);
        float o[6];
        pan6(v.x, v.y, ts, o);
        float2* p = reinterpret_cast<float2*>(six + i * 6);
        p[0] = make_float2(o[0], o[1]);
        p[1] = make_float2(o[2], o[3]);
        p[2] = make_float2(o[4], o[5]);
    }
}
void pan_stage(const float* d_stereo, i64 N, const TailSpec& ts_in, float* d_six) {
    if (N <= 0) return;
    TailSpec ts = ts_in;
    ts.N = N;
    pan_kernel<<<stream_grid(N), 256, 0, ctx().stream>>>(reinterpret_cast<const float2*>(d_stereo), ts, d_six);
    ARS_LAUNCH_CHECK();
    count_launch();
}

__global__ void __launch_bounds__(256) map_kernel(const float* __restrict__ six, TailSpec ts, float* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < ts.N; i += (i64)gridDim.x * blockDim.x) {
        float s[6], o[8];
        const float2* p = reinterpret_cast<const float2*>(six + i * 6);
        const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y; s[4] = c.x; s[5] = c.y;
        float rl_d = 0.f, rr_d = 0.f;
        if (ts.layout >= LAYOUT_7_1 && i >= ts.delay) {
            const float2 d = __ldg(reinterpret_cast<const float2*>(six + (i - (ts.delay > 0 ? ts.delay : 0)) * 6) + 2);
            rl_d = d.x; rr_d = d.y;
        }
        map_frame(s, rl_d, rr_d, ts, o);
        #pragma unroll
        for (int ch = 0; ch < 8; ++ch) if (ch < ts.C) out[i * ts.C + ch] = o[ch];
    }
}
void map_stage(const float* d_six, i64 N, const TailSpec& ts_in, float* d_out) {
    if (N <= 0) return;
    TailSpec ts = ts_in;
    ts.N = N;
    map_kernel<<<stream_grid(N), 256, 0, ctx().stream>>>(d_six, ts, d_out);
    ARS_LAUNCH_CHECK();
    count_launch();
}

__global__ void __launch_bounds__(256) delay_kernel(const float* __restrict__ in, i64 count, i64 shift, float* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x)
        out[i] = i >= shift ? in[i - shift] : 0.f;
}
void delay_stage(const float* d_in, i64 N, int ch, i64 delay, float* d_out) {
    if (N <= 0 || ch <= 0) return;
    delay_kernel<<<stream_grid(N * ch), 256, 0, ctx().stream>>>(d_in, N * ch, std::max<i64>(0, delay) * ch, d_out);
    ARS_LAUNCH_CHECK();
    count_launch();
}

__global__ void __launch_bounds__(256) pcm16_kernel(const float* __restrict__ x, i64 count, short* __restrict__ pcm) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x)
        pcm[i] = pcm_of(__ldg(x + i));
}
void pcm16_stage(const float* d_x, i64 count, short* d_pcm) {
    if (count <= 0) return;
    pcm16_kernel<<<stream_grid(count), 256, 0, ctx().stream>>>(d_x, count, d_pcm);
    ARS_LAUNCH_CHECK();
    count_launch();
}

__global__ void __launch_bounds__(256) sums_kernel(const float* __restrict__ x, i64 N, int C, RenderState* st,
                                                   float* __restrict__ mono) {
    unsigned pk = 0, mm = 0;
    double ss = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (i64)gridDim.x * blockDim.x) {
        const float* p = x + i * C;
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(p + c);
            pk = max(pk, abs_bits(v));
            ss += (double)__fmul_rn(v, v);
        }
        if (mono) {
            const float mv = C >= 2 ? __fdiv_rn(__fadd_rn(__ldg(p), __ldg(p + 1)), 2.0f) : __ldg(p);
            mono[i] = mv;
            mm = max(mm, abs_bits(mv));
        }
    }
    block_atomic_max(pk, &st->peak_final);
    block_atomic_max(mm, &st->mono_max);
    block_atomic_add(ss, &st->sumsq);
}
void sums_stage(const float* d_x, i64 N, int C, RenderState* d_state, float* d_mono) {
    if (N <= 0 || C <= 0) return;
    sums_kernel<<<stream_grid(N), 256, 0, ctx().stream>>>(d_x, N, C, d_state, d_mono);
    ARS_LAUNCH_CHECK();
    count_launch();
}

// per-channel sum of squares and the side signal's ((ch0 - ch1) * 0.5) sum of squares: the numerics of the
// reference's A/B report (rs.py:769-798); sums[c] for c < C, sums[C] = side
__global__ void __launch_bounds__(256) channel_sums_kernel(const float* __restrict__ x, i64 N, int C, double* sums) {
    double acc[9];
    #pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (i64)gridDim.x * blockDim.x) {
        const float* p = x + i * C;
        #pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < C) { const float v = __ldg(p + c); acc[c] += (double)__fmul_rn(v, v); }
        if (C >= 2) { const float sd = __fmul_rn(__fsub_rn(__ldg(p), __ldg(p + 1)), 0.5f); acc[8] += (double)__fmul_rn(sd, sd); }
    }
    #pragma unroll
    for (int c = 0; c < 9; ++c)
        if (c < C || c == 8) block_atomic_add(acc[c], sums + (c == 8 ? C : c));
}
void channel_sums(const float* d_x, i64 N, int C, double* d_sums) {
    if (N <= 0 || C <= 0) return;
    ARS_CHECK(C <= 8, "channel_sums: at most 8 channels");
    channel_sums_kernel<<<stream_grid(N), 256, 0, ctx().stream>>>(d_x, N, C, d_sums);
    ARS_LAUNCH_CHECK();
    count_launch();
}

__global__ void __launch_bounds__(256) stereo_from_kernel(const float* __restrict__ x, i64 n, int cin, float2* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const float l = x[i * cin];
        out[i] = make_float2(l, cin > 1 ? x[i * cin + 1] : l);
    }
}
void stereo_from(const float* d_x, i64 n, int cin, float* d_out) {
    if (n <= 0) return;
    stereo_from_kernel<<<stream_grid(n), 256, 0, ctx().stream>>>(d_x, n, cin, reinterpret_cast<float2*>(d_out));
    ARS_LAUNCH_CHECK();
    count_launch();
}

}  // namespace ars
