// Epilogue kernels (see epilogue.cuh).  All of them are one-touch HBM streams: grid-stride,
// grid = a multiple of the SM count, vector loads/stores, block reduce + one atomic per block.
#include "epilogue.cuh"
#include "tail_math.cuh"

namespace ars {

static inline int stream_grid(i64 items, int per_block = 256) {
    const i64 need = (items + per_block - 1) / per_block;
    const i64 cap = (i64)ctx().sm_count * 8;
    return (int)std::max<i64>(1, std::min(need, cap));
}

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

template <int C>
__global__ void __launch_bounds__(256) final_kernel(const float2* __restrict__ y, TailSpec ts, RenderState* st,
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
    if (g2.mode == 0 && g3.mode == 0) {
        if (g1.mode == 0) final_body<C, 0, 0, 0>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
        else if (g1.mode == 1) final_body<C, 2, 0, 0>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
        else final_body<C, 1, 0, 0>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
    } else {
        final_body<C, 1, 1, 1>(y, ts, g1, g2, g3, out, pcm, mono, pkf, nan_seen, mm, ss);
    }
    const unsigned pk = nan_seen ? 0x7fc00000u : __float_as_uint(pkf);
    block_atomic_max(pk, &st->peak_final);
    block_atomic_max(mm, &st->mono_max);
    block_atomic_add(ss, &st->sumsq);
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
    const TailSpec ts = with_window(ts_in);
    if (ts.N <= 0 || ts.i_hi <= ts.i_lo) return;
    Ctx& c = ctx();
    const int grid = stream_grid(ts.i_hi - ts.i_lo);
    KernelScope prof("final_kernel (guards, pan, map, clip, PCM16, sums)",
                     (double)(ts.i_hi - ts.i_lo) * (8.0 + (d_pcm ? 2.0 * ts.C : 0.0) + (d_out ? 4.0 * ts.C : 0.0) + (d_mono ? 4.0 : 0.0)));
    if (ts.C == 2) final_kernel<2><<<grid, 256, 0, c.stream>>>(d_y, ts, d_state, d_out, d_pcm, d_mono);
    else if (ts.C == 6) final_kernel<6><<<grid, 256, 0, c.stream>>>(d_y, ts, d_state, d_out, d_pcm, d_mono);
    else final_kernel<8><<<grid, 256, 0, c.stream>>>(d_y, ts, d_state, d_out, d_pcm, d_mono);
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
