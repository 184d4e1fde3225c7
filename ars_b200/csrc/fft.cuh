// Power-of-two complex64 FFT engine for sm_100a (hand-written; cuFFT is not used).
//
// A length-M transform (M = 2^m, m <= 30) is a short sequence of in-place "passes"
// over an M-element complex buffer in HBM.  Each pass moves a tile through shared
// memory exactly once (read 8 B + write 8 B per element), does a complete
// R-point sub-transform of every tile column with register-resident radix-8/16
// butterflies, and applies the inter-pass twiddle on the way out.
//
//   forward  = decimation in frequency: strided passes first (segment length
//              Lg = M, M/R1, ...), the contiguous pass last; output is left in a
//              digit-permuted order;
//   inverse  = the exact algebraic inverse (decimation in time, conjugate
//              twiddles, passes and stages in reverse order); it consumes the
//              permuted order and produces natural order.
//
// Because every use in this library is  IFFT(FFT(a) .* S)  with S produced by the same
// forward transform, no reordering pass is ever needed.  Loads of the first pass and
// stores of the last pass go through small functors (Ld / St) so chirp multiplies,
// zero padding, spectrum products, scaling and the abs-max reduction are fused into
// the passes instead of costing extra trips through HBM.
#pragma once
#include "ars_common.cuh"

#ifdef __CUDA_ARCH__
#define ARS_LDG(p) __ldg(p)
// streaming forms: data touched once gets the evict-first policy so that it does not push the work buffers out of the L2
#define ARS_LDCS(p) __ldcs(p)
#define ARS_STCS(p, v) __stcs(p, v)
#else
#define ARS_LDG(p) (*(p))
#define ARS_LDCS(p) (*(p))
#define ARS_STCS(p, v) (*(p) = (v))
#endif
#define ARS_HD __host__ __device__ __forceinline__

namespace ars {
namespace fft {

// ------------------------------------------------------------------ radices ---
template <int LOGR> struct Rad;
#define ARS_RAD(L, N, A, B, C, D)                                   \
    template <> struct Rad<L> {                                     \
        static constexpr int n = N;                                 \
        static constexpr int r0 = A, r1 = B, r2 = C, r3 = D;        \
        static __host__ __device__ constexpr int r(int s) { return s == 0 ? A : s == 1 ? B : s == 2 ? C : D; } \
    };
ARS_RAD(1, 1, 2, 1, 1, 1)
ARS_RAD(2, 1, 4, 1, 1, 1)
ARS_RAD(3, 1, 8, 1, 1, 1)
ARS_RAD(4, 1, 16, 1, 1, 1)
ARS_RAD(5, 2, 8, 4, 1, 1)
ARS_RAD(6, 2, 8, 8, 1, 1)
ARS_RAD(7, 2, 16, 8, 1, 1)
ARS_RAD(8, 2, 16, 16, 1, 1)
ARS_RAD(9, 3, 8, 8, 8, 1)
ARS_RAD(10, 3, 16, 8, 8, 1)
ARS_RAD(11, 3, 16, 16, 8, 1)
ARS_RAD(12, 3, 16, 16, 16, 1)
// (other four-stage splits of 13 measured on the overlap-save route: 16.16.16.2 -0.4 %, 16.16.8.4 +0.3 %, 8.8.8.16 +0.8 %,
// 2.16.16.16 +8 % -- the stage count and its barriers set the pace of this tile, not the radix mix)
ARS_RAD(13, 4, 16, 8, 8, 8)
#undef ARS_RAD

constexpr int BIG_LO_LOG = 15;         // two-level table for w_M^e

// ---------------------------------------------------------------- butterflies -
// Packed FP32x2 arithmetic (Blackwell FADD2 / FFMA2: one instruction per complex add) -- the butterflies are
// mostly complex additions, so this nearly halves their instruction count.  Results are bit-identical to the
// scalar forms (each component is one IEEE add; a - b is fma(b, -1, a)).  The host build uses the scalar forms.
ARS_HD float fadd_rn(float a, float b) {      // a float32 add that can never be contracted into an FMA
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
ARS_HD float2 padd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return cadd(a, b);
#endif
}
ARS_HD float2 psub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
#else
    return csub(a, b);
#endif
}
// t + (-/+ i) d  and  t - (-/+ i) d   (forward: -i, inverse: +i)
template <bool INV> ARS_HD void rot_addsub(float2 t, float2 d, float2& plus, float2& minus) {
#ifdef __CUDA_ARCH__
    const float2 sw = make_float2(d.y, d.x);
    const float2 cp = INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f);
    const float2 cm = INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f);
    plus = __ffma2_rn(sw, cp, t);
    minus = __ffma2_rn(sw, cm, t);
#else
    const float2 r = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    plus = cadd(t, r);
    minus = csub(t, r);
#endif
}

template <bool INV> ARS_HD float2 mul_mi(float2 a) {
    // forward: multiply by -i ; inverse: multiply by +i
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

constexpr float COS16[16] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                             0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                             -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
                             0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
constexpr float SIN16[16] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                             1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                             0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                             -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};

// a * w_16^E  (forward: w = exp(-2 pi i/16); inverse: conjugate)
template <int E, bool INV> ARS_HD float2 mulw16(float2 a) {
    constexpr int e = E & 15;
    if constexpr (e == 0) return a;
    else if constexpr (e == 4) return mul_mi<INV>(a);
    else if constexpr (e == 8) return make_float2(-a.x, -a.y);
    else if constexpr (e == 12) return mul_mi<!INV>(a);
    else {
        constexpr float c = COS16[e];
        constexpr float s = INV ? SIN16[e] : -SIN16[e];
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

template <bool INV> ARS_HD void bf4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = padd(a0, a2), t1 = psub(a0, a2), t2 = padd(a1, a3), d = psub(a1, a3);
    a0 = padd(t0, t2); a2 = psub(t0, t2);
    rot_addsub<INV>(t1, d, a1, a3);
}

// natural-order in, natural-order out R-point DFT held in registers
template <int R, bool INV> struct Dft;
template <bool INV> struct Dft<2, INV> {
    static ARS_HD void run(float2 (&v)[2]) {
        float2 t = v[0]; v[0] = padd(t, v[1]); v[1] = psub(t, v[1]);
    }
};
template <bool INV> struct Dft<4, INV> {
    static ARS_HD void run(float2 (&v)[4]) { bf4<INV>(v[0], v[1], v[2], v[3]); }
};
template <bool INV> struct Dft<8, INV> {
    static ARS_HD void run(float2 (&v)[8]) {
        bf4<INV>(v[0], v[2], v[4], v[6]);
        bf4<INV>(v[1], v[3], v[5], v[7]);
        float2 b1 = mulw16<2, INV>(v[3]), b3 = mulw16<6, INV>(v[7]);
        float2 o0 = padd(v[0], v[1]), o4 = psub(v[0], v[1]);
        float2 o1 = padd(v[2], b1), o5 = psub(v[2], b1);
        float2 o2, o6;
        rot_addsub<INV>(v[4], v[5], o2, o6);
        float2 o3 = padd(v[6], b3), o7 = psub(v[6], b3);
        v[0] = o0; v[1] = o1; v[2] = o2; v[3] = o3; v[4] = o4; v[5] = o5; v[6] = o6; v[7] = o7;
    }
};
template <bool INV> struct Dft<16, INV> {
    static ARS_HD void run(float2 (&v)[16]) {
        bf4<INV>(v[0], v[4], v[8], v[12]);
        bf4<INV>(v[1], v[5], v[9], v[13]);
        bf4<INV>(v[2], v[6], v[10], v[14]);
        bf4<INV>(v[3], v[7], v[11], v[15]);
        // v[n2 + 4*k1] *= w16^(n2*k1)
        v[5] = mulw16<1, INV>(v[5]);  v[6] = mulw16<2, INV>(v[6]);   v[7] = mulw16<3, INV>(v[7]);
        v[9] = mulw16<2, INV>(v[9]);  v[10] = mulw16<4, INV>(v[10]); v[11] = mulw16<6, INV>(v[11]);
        v[13] = mulw16<3, INV>(v[13]); v[14] = mulw16<6, INV>(v[14]); v[15] = mulw16<9, INV>(v[15]);
        bf4<INV>(v[0], v[1], v[2], v[3]);
        bf4<INV>(v[4], v[5], v[6], v[7]);
        bf4<INV>(v[8], v[9], v[10], v[11]);
        bf4<INV>(v[12], v[13], v[14], v[15]);
        // X[k1 + 4*k2] sits in v[4*k1 + k2]: transpose
        float2 t;
        t = v[1]; v[1] = v[4]; v[4] = t;   t = v[2]; v[2] = v[8]; v[8] = t;
        t = v[3]; v[3] = v[12]; v[12] = t; t = v[6]; v[6] = v[9]; v[9] = t;
        t = v[7]; v[7] = v[13]; v[13] = t; t = v[11]; v[11] = v[14]; v[14] = t;
    }
};

// ------------------------------------------------------------- Ld / St functors
enum LdMode { LD_PLAIN = 0, LD_MULSPEC, LD_CHIRP_X2, LD_CHIRP_XC, LD_CHIRP_PAIR, LD_CHIRP_B, LD_CHIRP_C,
              LD_REAL_PAIR, LD_OLS_X, LD_OLS_IR, LD_OLS_MAC, LD_OLS_CHIRPSIG, LD_OLS_IRC, LD_OLS_X2, LD_OLS_IR2,
              LD_OLSB_X, LD_TAPS };
enum StMode { ST_PLAIN = 0, ST_SCALE, ST_CHIRP, ST_FINAL, ST_OLS, ST_OLS_CHIRP, ST_OLS2, ST_OLSB };

struct Ld {
    int mode = LD_PLAIN;
    int stream = 0;                 // LD_OLSB_X: read the signal with the evict-first policy (it is read once)
    const float2* a = nullptr;      // complex source / work buffer
    const float2* b = nullptr;      // second complex operand (spectrum or chirp)
    const float* f0 = nullptr;      // real sources
    const float* f1 = nullptr;
    i64 nvalid = 0;                 // elements of the source that exist (rest is zero padding)
    i64 nvalid1 = 0;
    i64 N = 0, M = 0;
    int cin = 2;
    // overlap-save (upols.cu): segment s of 2^logF points = hop-B window of the signal / partition of the IR
    const float2* a2 = nullptr;     // LD_OLS_MAC: second delay line (spectra of the conjugated signal), or null
    const float2* b2 = nullptr;     // LD_OLS_MAC: its coefficient spectra
    const unsigned char* nz = nullptr;   // LD_OLS_MAC: per-partition "has non-zero taps" flags
    const int* plist = nullptr;          // LD_OLS_MAC: compacted non-zero partition indices + count at [P] (preferred)
    int P = 0;                      // partitions
    int logF = 13;                  // log2 of the segment (FFT) length; hop B = F / 2
    float c0 = 1.f, c1 = 0.f;       // LD_OLS_IR: tap = c0 * f0[i*cin] + c1 * f1[i*cin]; LD_OLS_X: c1 = -1 conjugates
    i64 seg0 = 0;                   // LD_OLS_X: absolute block index of the launch's segment 0
    i64 frame0 = 0;                 // LD_OLS_X: absolute frame index of f0[0] (a rank may hold a slice of the signal)
    i64 lookback = 0;               // LD_OLS_MAC: delay-line segments that exist before the launch's segment 0
    i64 adv = 0;                    // LD_OLS_X: the windows start `adv` frames later (IR with taps at negative times)
    i64 circ = 0;                   // LD_OLS_X: > 0: the signal is the circ-periodic extension of the zero-padded frames
    // LD_OLS_X2 / LD_OLS_IR2 (and ST_OLS2): the 2B-point overlap-save transform as a radix-2 stage folded into the
    // load (store) plus two B-point transforms -- sub-segment 0 of a segment holds the even bins, sub-segment 1 the odd
    // ones.  The two halves of an overlap-save window are exactly the operands of that stage: a = w[i] + w[i+B],
    // b = (w[i] - w[i+B]) w_2B^i; an IR partition has an empty second half: a = h[i], b = h[i] w_2B^i.
    const float2* tw2 = nullptr;    // w_2B^i, i < B
    // LD_OLSB_X (big-block overlap-save, upols.cu): transform j of 2^logF points is the window of the signal that
    // starts `skip` frames before output frame (seg0 + j) * hop; skip = taps - 1 aliased outputs are dropped, hop <= F - skip.
    // seg0 / frame0 / nvalid / cin / adv / circ / c1 as for LD_OLS_X.
    i64 hop = 0, skip = 0;
    // LD_TAPS: real taps c0 * f0[i * cin] + c1 * f1[i * cin] (+ delta at index delta_at: the dry path of the mix folded
    // into the impulse response, y = dg x + dw (x * h) = x * (dg delta + dw h))
    i64 delta_at = -1;
    float delta = 0.f;
    // MODE >= 0: compile-time access mode (fast kernels); MODE < 0: runtime switch on `mode` (generic kernels)
    template <int MODE> ARS_HD float2 get(i64 idx) const {
        if constexpr (MODE < 0) return (*this)(idx);
        else if constexpr (MODE == LD_PLAIN) return a[idx];
        else if constexpr (MODE == LD_MULSPEC) return cmul(a[idx], ARS_LDG(b + idx));
        else if constexpr (MODE == LD_CHIRP_X2) {      // interleaved stereo frames = complex samples L + iR
            return idx < nvalid ? cmul(ARS_LDG(reinterpret_cast<const float2*>(f0) + idx), ARS_LDG(b + idx))
                                : make_float2(0.f, 0.f);
        } else if constexpr (MODE == LD_CHIRP_XC) {    // (frames, cin) floats: cin == 1 duplicates, cin > 2 keeps the first two
            if (idx >= nvalid) return make_float2(0.f, 0.f);
            float2 v;
            if ((cin & 1) == 0) {                      // even channel count: the first two channels are one aligned 8-byte load
                v = ARS_LDG(reinterpret_cast<const float2*>(f0 + idx * cin));
            } else {
                v.x = ARS_LDG(f0 + idx * cin);
                v.y = cin > 1 ? ARS_LDG(f0 + idx * cin + 1) : v.x;
            }
            return cmul(v, ARS_LDG(b + idx));
        } else if constexpr (MODE == LD_CHIRP_PAIR) {  // two real arrays packed as re + i*im (each may be absent / shorter)
            const float l = (f0 && idx < nvalid) ? ARS_LDG(f0 + idx) : 0.f;
            const float r = (f1 && idx < nvalid1) ? ARS_LDG(f1 + idx) : 0.f;
            if (idx >= N) return make_float2(0.f, 0.f);
            return cmul(make_float2(l, r), ARS_LDG(b + idx));
        } else if constexpr (MODE == LD_CHIRP_B) {     // Bluestein kernel: conj(chirp) on (-N, N), wrapped modulo M
            if (idx < N) return cconj(ARS_LDG(b + idx));
            if (idx > M - N) return cconj(ARS_LDG(b + (M - idx)));
            return make_float2(0.f, 0.f);
        } else if constexpr (MODE == LD_CHIRP_C) {     // complex N-vector times chirp
            return idx < nvalid ? cmul(ARS_LDG(a + idx), ARS_LDG(b + idx)) : make_float2(0.f, 0.f);
        } else if constexpr (MODE == LD_OLS_X) {       // overlap-save window s: frames [(s-1)B, (s+1)B) of the zero-padded signal
            const i64 seg = seg0 + (idx >> logF);
            i64 fr = (seg - 1) * ((i64)1 << (logF - 1)) + (idx & (((i64)1 << logF) - 1)) + adv - frame0;
            if (circ > 0) { if (fr < 0) fr += circ; else if (fr >= circ) fr -= circ; }      // |fr| < 2 circ (upols.cu)
            if (fr < 0 || fr >= nvalid) return make_float2(0.f, 0.f);      // nvalid = frames held at f0
            float l, r;
            if ((cin & 1) == 0) {                      // even channel count: the first two channels are one aligned 8-byte load
                const float2 v = ARS_LDG(reinterpret_cast<const float2*>(f0 + fr * cin));
                l = v.x; r = v.y;
            } else {
                l = ARS_LDG(f0 + fr * cin);
                r = cin > 1 ? ARS_LDG(f0 + fr * cin + 1) : l;
            }
            return make_float2(l, c1 < 0.f ? -r : r);
        } else if constexpr (MODE == LD_OLS_IR) {      // IR partition p: taps [pB, (p+1)B), zero-padded to 2B, real
            const i64 seg = idx >> logF;
            const i64 t = idx & (((i64)1 << logF) - 1);
            const i64 i = (seg << (logF - 1)) + t;
            if (t >= ((i64)1 << (logF - 1))) return make_float2(0.f, 0.f);
            const float u = (f0 && i < nvalid) ? ARS_LDG(f0 + i * cin) : 0.f;
            const float v = (f1 && i < nvalid1) ? ARS_LDG(f1 + i * cin) : 0.f;
            return make_float2(c0 * u + c1 * v, 0.f);
        } else if constexpr (MODE == LD_OLS_CHIRPSIG) {    // overlap-save window over the shifted Bluestein kernel:
            // sig[f] = conj(chirp[|f - D|]) for 0 <= f < N + D, D = frame0 (short-IR spectrum path, spectral.cu)
            const i64 seg = seg0 + (idx >> logF);
            const i64 fr = ((seg - 1) << (logF - 1)) + (idx & (((i64)1 << logF) - 1));
            if (fr < 0 || fr >= N + frame0) return make_float2(0.f, 0.f);
            i64 m = fr - frame0;
            if (m < 0) m = -m;
            if (m >= N) return make_float2(0.f, 0.f);       // only ever multiplied by zero-padded taps
            return cconj(ARS_LDG(b + m));
        } else if constexpr (MODE == LD_OLS_IRC) {         // complex taps (f0 + i f1) * chirp, partition p, zero-padded
            const i64 seg = idx >> logF;
            const i64 t = idx & (((i64)1 << logF) - 1);
            const i64 i = (seg << (logF - 1)) + t;
            if (t >= ((i64)1 << (logF - 1)) || i >= N) return make_float2(0.f, 0.f);
            const float u = (f0 && i < nvalid) ? ARS_LDG(f0 + i * cin) : 0.f;
            const float v = (f1 && i < nvalid1) ? ARS_LDG(f1 + i * cin) : 0.f;
            return cmul(make_float2(u, v), ARS_LDG(b + i));
        } else if constexpr (MODE == LD_OLS_X2) {      // radix-2 stage over the window of segment s (see tw2)
            const i64 B = (i64)1 << (logF - 1);
            const i64 seg = seg0 + (idx >> logF);
            const i64 i = idx & (B - 1);
            const float2 x0 = frame_at((seg - 1) * B + i + adv - frame0), x1 = frame_at(seg * B + i + adv - frame0);
            if ((idx & B) == 0) return make_float2(x0.x + x1.x, x0.y + x1.y);
            return cmul(make_float2(x0.x - x1.x, x0.y - x1.y), ARS_LDG(tw2 + i));
        } else if constexpr (MODE == LD_OLS_IR2) {     // IR partition p, radix-2 stage over [taps | zeros]
            const i64 B = (i64)1 << (logF - 1);
            const i64 t = idx & (B - 1);
            const i64 i = ((idx >> logF) << (logF - 1)) + t;
            const float u = (f0 && i < nvalid) ? ARS_LDG(f0 + i * cin) : 0.f;
            const float v = (f1 && i < nvalid1) ? ARS_LDG(f1 + i * cin) : 0.f;
            const float h = c0 * u + c1 * v;
            if ((idx & B) == 0) return make_float2(h, 0.f);
            const float2 w = ARS_LDG(tw2 + t);
            return make_float2(h * w.x, h * w.y);
        } else if constexpr (MODE == LD_OLS_MAC) {     // Y_s = sum_p X_{s-p} H_p (+ Xc_{s-p} Hc_p)
            float2 acc[1];
            get_mac<1>(idx, 0, acc);
            return acc[0];
        } else if constexpr (MODE == LD_OLSB_X) {      // big-block overlap-save window (see hop / skip above)
            const i64 j = idx >> logF;
            const i64 t = idx & (((i64)1 << logF) - 1);
            return frame_at((seg0 + j) * hop - skip + t + adv - frame0);
        } else if constexpr (MODE == LD_TAPS) {        // real taps c0 * f0[i * cin] + c1 * f1[i * cin], zero-padded
            const float u = (f0 && idx < nvalid) ? ARS_LDG(f0 + idx * cin) : 0.f;
            const float v = (f1 && idx < nvalid1) ? ARS_LDG(f1 + idx * cin) : 0.f;
            return make_float2(c0 * u + c1 * v + (idx == delta_at ? delta : 0.f), 0.f);
        } else {                                       // LD_REAL_PAIR: plain zero-padded packing
            const float l = (f0 && idx < nvalid) ? ARS_LDG(f0 + idx) : 0.f;
            const float r = (f1 && idx < nvalid1) ? ARS_LDG(f1 + idx) : 0.f;
            return make_float2(l, r);
        }
    }
    // LD_OLS_X for the r elements of one first-stage butterfly (`step` apart, all inside one window): when the whole
    // stretch lies inside the signal -- no padding, no wrap: all but a handful of tiles -- the frame index, the bounds
    // and the byte address are formed once instead of per element (a third of the pass's instructions otherwise).
    template <int r> ARS_HD void get_x(i64 idx0, i64 step, float2 (&v)[r]) const {
        const i64 F = (i64)1 << logF;
        const i64 fr = ((seg0 + (idx0 >> logF)) - 1) * (F >> 1) + (idx0 & (F - 1)) + adv - frame0;
        const i64 last = fr + (r - 1) * step;
        if ((cin & 1) == 0 && fr >= 0 && last < nvalid && (circ <= 0 || last < circ)) {
            const float2* p = reinterpret_cast<const float2*>(f0 + fr * cin);
            const i64 ps = step * (cin >> 1);
            #pragma unroll
            for (int k = 0; k < r; ++k) v[k] = ARS_LDG(p + k * ps);
            if (c1 < 0.f) {
                #pragma unroll
                for (int k = 0; k < r; ++k) v[k].y = -v[k].y;
            }
        } else {
            #pragma unroll
            for (int k = 0; k < r; ++k) v[k] = get<LD_OLS_X>(idx0 + k * step);
        }
    }
    // LD_OLSB_X for the r elements of one first-stage butterfly (`step` apart, all inside one transform): same idea
    template <int r> ARS_HD void get_xb(i64 idx0, i64 step, float2 (&v)[r]) const {
        const i64 j = idx0 >> logF;
        const i64 t0 = idx0 & (((i64)1 << logF) - 1);
        const i64 fr = (seg0 + j) * hop - skip + t0 + adv - frame0;
        const i64 last = fr + (r - 1) * step;
        if (fr >= 0 && last < nvalid && (circ <= 0 || last < circ)) {
            if ((cin & 1) == 0) {
                const float2* p = reinterpret_cast<const float2*>(f0 + fr * cin);
                const i64 ps = step * (cin >> 1);
                if (stream) {
                    #pragma unroll
                    for (int k = 0; k < r; ++k) v[k] = ARS_LDCS(p + k * ps);
                } else {
                    #pragma unroll
                    for (int k = 0; k < r; ++k) v[k] = ARS_LDG(p + k * ps);
                }
            } else {
                const float* p = f0 + fr * cin;
                const i64 ps = step * cin;
                #pragma unroll
                for (int k = 0; k < r; ++k) {
                    const float l = ARS_LDG(p + k * ps);
                    v[k] = make_float2(l, cin > 1 ? ARS_LDG(p + k * ps + 1) : l);
                }
            }
            if (c1 < 0.f) {
                #pragma unroll
                for (int k = 0; k < r; ++k) v[k].y = -v[k].y;
            }
        } else {
            #pragma unroll
            for (int k = 0; k < r; ++k) v[k] = get<LD_OLSB_X>(idx0 + k * step);
        }
    }
    // one stereo frame of the (periodically extended, zero-padded) signal as L + iR; c1 < 0 conjugates
    ARS_HD float2 frame_at(i64 fr) const {
        if (circ > 0) { if (fr < 0) fr += circ; else if (fr >= circ) fr -= circ; }
        if (fr < 0 || fr >= nvalid) return make_float2(0.f, 0.f);
        float l, r;
        if ((cin & 1) == 0) {
            const float2 v = ARS_LDG(reinterpret_cast<const float2*>(f0 + fr * cin));
            l = v.x; r = v.y;
        } else {
            l = ARS_LDG(f0 + fr * cin);
            r = cin > 1 ? ARS_LDG(f0 + fr * cin + 1) : l;
        }
        return make_float2(l, c1 < 0.f ? -r : r);
    }
    // LD_OLS_X: ask the L2 for the new half (B frames) of the windows of `nsegs` segments starting at launch segment
    // `seg` -- the pass waits on its first loads (DRAM at a third of its bandwidth), so a hint a couple of waves
    // ahead turns DRAM latency into L2 latency.  Only a hint: a stretch that wraps around the period is clipped.
    ARS_HD void prefetch_x(i64 seg, int nsegs, int tid, int nthreads) const {
#ifdef __CUDA_ARCH__
        const i64 B = (i64)1 << (logF - 1);
        i64 lo = (seg0 + seg) * B + adv - frame0;
        if (circ > 0) { if (lo < 0) lo += circ; else if (lo >= circ) lo -= circ; }
        i64 hi = lo + (i64)nsegs * B;
        if (lo < 0) lo = 0;
        if (hi > nvalid) hi = nvalid;
        if (hi <= lo) return;
        const char* p0 = reinterpret_cast<const char*>(f0 + lo * cin);
        const char* p1 = reinterpret_cast<const char*>(f0 + hi * cin);
        for (const char* p = p0 + (i64)tid * 128; p < p1; p += (i64)nthreads * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
    }
    // The complex multiply-accumulate over the partitions, fused into the loads of the inverse transform:
    // r spectrum bins of one butterfly at a time, so 2r (4r) independent loads are in flight per partition.
    template <int r> ARS_HD void get_mac(i64 idx0, i64 step, float2 (&v)[r]) const {
        const i64 F = (i64)1 << logF;
        const i64 seg = idx0 >> logF;
        const i64 t0 = idx0 & (F - 1);
        #pragma unroll
        for (int k = 0; k < r; ++k) v[k] = make_float2(0.f, 0.f);
        const i64 reach = seg + lookback;
        const int pmax = (int)(reach < (i64)(P - 1) ? reach : (i64)(P - 1));
        // plist (optional): ascending indices of the non-zero partitions, plist[P] = their count
        const int np = plist ? plist[P] : P;
        for (int q = 0; q < np; ++q) {
            const int p = plist ? plist[q] : q;
            if (p > pmax) break;
            if (!plist && nz && !nz[p]) continue;
            const float2* x = a + ((seg - p) << logF) + t0;
            const float2* h = b + ((i64)p << logF) + t0;
            #pragma unroll
            for (int k = 0; k < r; ++k) {
                const float2 xv = ARS_LDG(x + k * step), hv = ARS_LDG(h + k * step);
                v[k].x = fmaf(xv.x, hv.x, v[k].x);
                v[k].x = fmaf(-xv.y, hv.y, v[k].x);
                v[k].y = fmaf(xv.x, hv.y, v[k].y);
                v[k].y = fmaf(xv.y, hv.x, v[k].y);
            }
            if (a2) {
                const float2* x2 = a2 + ((seg - p) << logF) + t0;
                const float2* h2 = b2 + ((i64)p << logF) + t0;
                #pragma unroll
                for (int k = 0; k < r; ++k) {
                    const float2 xv = ARS_LDG(x2 + k * step), hv = ARS_LDG(h2 + k * step);
                    v[k].x = fmaf(xv.x, hv.x, v[k].x);
                    v[k].x = fmaf(-xv.y, hv.y, v[k].x);
                    v[k].y = fmaf(xv.x, hv.y, v[k].y);
                    v[k].y = fmaf(xv.y, hv.x, v[k].y);
                }
            }
        }
    }
    ARS_HD float2 operator()(i64 idx) const {
        switch (mode) {
            case LD_PLAIN: return get<LD_PLAIN>(idx);
            case LD_MULSPEC: return get<LD_MULSPEC>(idx);
            case LD_CHIRP_X2: return get<LD_CHIRP_X2>(idx);
            case LD_CHIRP_XC: return get<LD_CHIRP_XC>(idx);
            case LD_CHIRP_PAIR: return get<LD_CHIRP_PAIR>(idx);
            case LD_CHIRP_B: return get<LD_CHIRP_B>(idx);
            case LD_CHIRP_C: return get<LD_CHIRP_C>(idx);
            case LD_REAL_PAIR: return get<LD_REAL_PAIR>(idx);
            case LD_OLS_X: return get<LD_OLS_X>(idx);
            case LD_OLS_IR: return get<LD_OLS_IR>(idx);
            case LD_OLS_MAC: return get<LD_OLS_MAC>(idx);
            case LD_OLS_CHIRPSIG: return get<LD_OLS_CHIRPSIG>(idx);
            case LD_OLS_IRC: return get<LD_OLS_IRC>(idx);
            case LD_OLS_X2: return get<LD_OLS_X2>(idx);
            case LD_OLS_IR2: return get<LD_OLS_IR2>(idx);
            case LD_OLSB_X: return get<LD_OLSB_X>(idx);
            case LD_TAPS: return get<LD_TAPS>(idx);
        }
        return make_float2(0.f, 0.f);
    }
};

struct St {
    int mode = ST_PLAIN;
    int stream = 0;                  // ST_OLSB: store the frames with the evict-first policy (read again only after the whole pass)
    float2* a = nullptr;
    const float2* chirp = nullptr;
    i64 N = 0;
    float scale = 1.f;
    // ST_OLS: keep the second half of every overlap-save segment, mix with the dry frame, track the maxima
    const float* dry = nullptr;      // (n, cin) input frames
    int cin = 2;
    i64 n = 0;                       // input frames (dry is zero beyond)
    int logF = 13;
    float dg = 0.f, dw = 1.f;        // y = dg * dry + dw * wet   (rs.py:113)
    i64 seg0 = 0;                    // absolute block index of the launch's segment 0
    i64 frame0 = 0, dry_frame0 = 0;  // absolute frame index of a[0] / dry[0]; N = absolute end frame (exclusive)
    unsigned* maxbits = nullptr;     // ST_FINAL: 4 words: bits of max |x|, max |x.re|, max |x.im|, max |f32(re + im)|
    unsigned local_max = 0, local_l = 0, local_r = 0, local_lr = 0;
    const float2* tw2 = nullptr;     // ST_OLS2: w_2B^i, i < B (see Ld::tw2)
    // ST_OLSB (big-block overlap-save): element t of transform j is output frame (seg0 + j) * hop + t - skip when
    // skip <= t < skip + hop
    i64 hop = 0, skip = 0;
    // the dry frame that goes with overlap-save output frame `fr` (zero outside the slice held at `dry`)
    ARS_HD float2 dry_at(i64 fr) const {
        const i64 df = fr - dry_frame0;
        if (fr >= N || df < 0 || df >= n) return make_float2(0.f, 0.f);     // n = dry frames held at `dry`
        if ((cin & 1) == 0) return ARS_LDG(reinterpret_cast<const float2*>(dry + df * cin));
        const float l = ARS_LDG(dry + df * cin);
        return make_float2(l, cin > 1 ? ARS_LDG(dry + df * cin + 1) : l);
    }
    // one overlap-save output frame: mix with the dry frame, store, track the maxima
    ARS_HD void ols_out(i64 fr, float2 v, float2 d) {
        if (fr >= N) return;
        const float2 y = make_float2(dg * d.x + dw * v.x, dg * d.y + dw * v.y);
        a[fr - frame0] = y;
        const unsigned m0 = abs_bits(y.x), m1 = abs_bits(y.y), m2 = abs_bits(fadd_rn(y.x, y.y));
        if (m0 > local_l) local_l = m0;
        if (m1 > local_r) local_r = m1;
        if (m2 > local_lr) local_lr = m2;
    }
    // ST_OLS for the r outputs of one last-stage butterfly (`step` apart inside one segment): when the segment's whole
    // second half is inside the output and has dry frames, frame index, bounds and addresses are formed once
    template <int r> ARS_HD void put_ols(i64 idx0, i64 step, const float2 (&v)[r]) {
        const i64 F = (i64)1 << logF, B = F >> 1;
        const i64 t0 = idx0 & (F - 1);
        const i64 fr0 = ((seg0 + (idx0 >> logF)) << (logF - 1)) + t0 - B;      // frame of element 0 (first half: < block start)
        const i64 blk = fr0 - t0 + B;                                          // first output frame of the segment
        if ((cin & 1) == 0 && blk + B <= N && blk >= dry_frame0 && blk + B - dry_frame0 <= n) {
            const float2* dp = reinterpret_cast<const float2*>(dry + (fr0 - dry_frame0) * cin);
            const i64 ds = step * (cin >> 1);
            float2* ap = a + (fr0 - frame0);
            #pragma unroll
            for (int k = 0; k < r; ++k) {
                if (t0 + k * step >= B) {
                    const float2 d = ARS_LDG(dp + k * ds);
                    const float2 y = make_float2(dg * d.x + dw * v[k].x, dg * d.y + dw * v[k].y);
                    ap[k * step] = y;
                    const unsigned m0 = abs_bits(y.x), m1 = abs_bits(y.y), m2 = abs_bits(fadd_rn(y.x, y.y));
                    if (m0 > local_l) local_l = m0;
                    if (m1 > local_r) local_r = m1;
                    if (m2 > local_lr) local_lr = m2;
                }
            }
        } else {
            #pragma unroll
            for (int k = 0; k < r; ++k) put<ST_OLS>(idx0 + k * step, v[k], make_float2(1.f, 0.f));
        }
    }
    // ST_OLSB for the r outputs of one last-stage butterfly (`step` apart inside one transform)
    template <int r> ARS_HD void put_olsb(i64 idx0, i64 step64, const float2 (&v)[r]) {
        // (offsets inside a transform fit 32 bits: F <= 2^22; one unsigned compare tells 0 <= o < hop)
        const i64 j = idx0 >> logF;
        const int t0 = (int)(idx0 & (((i64)1 << logF) - 1));
        const int step = (int)step64, hp = (int)hop;
        const int o0 = t0 - (int)skip;                              // output index inside the transform's hop
        const i64 blk = (seg0 + j) * hop;                           // first output frame of the transform
        const int olast = o0 + (r - 1) * step;
        const bool mix = dg != 0.f;                                 // (dg == 0: the dry path is part of the taps, Ld::delta)
        const bool dry_ok = !mix ? true : ((cin & 1) == 0 && blk + (o0 < 0 ? 0 : o0) >= dry_frame0 &&
                                           blk + olast - dry_frame0 < n);
        if (dry_ok && blk + (olast < hp ? olast : hp - 1) < N) {
            float2* ap = a + (blk + o0 - frame0);
            if (!mix) {
                #pragma unroll
                for (int k = 0; k < r; ++k) {
                    if ((unsigned)(o0 + k * step) < (unsigned)hp) {
                        const float2 y = make_float2(dw * v[k].x, dw * v[k].y);
                        if (stream) ARS_STCS(ap + k * step, y);
                        else ap[k * step] = y;
                        local_l = max(local_l, abs_bits(y.x));
                        local_r = max(local_r, abs_bits(y.y));
                        local_lr = max(local_lr, abs_bits(fadd_rn(y.x, y.y)));
                    }
                }
            } else {
                const float2* dp = reinterpret_cast<const float2*>(dry + (blk + o0 - dry_frame0) * cin);
                const i64 ds = (i64)step * (cin >> 1);
                #pragma unroll
                for (int k = 0; k < r; ++k) {
                    if ((unsigned)(o0 + k * step) < (unsigned)hp) {
                        const float2 d = ARS_LDG(dp + k * ds);
                        const float2 y = make_float2(dg * d.x + dw * v[k].x, dg * d.y + dw * v[k].y);
                        ap[k * step] = y;
                        local_l = max(local_l, abs_bits(y.x));
                        local_r = max(local_r, abs_bits(y.y));
                        local_lr = max(local_lr, abs_bits(fadd_rn(y.x, y.y)));
                    }
                }
            }
        } else {
            #pragma unroll
            for (int k = 0; k < r; ++k) put<ST_OLSB>(idx0 + k * step64, v[k], make_float2(1.f, 0.f));
        }
    }
    // ST_OLS2: the radix-2 stage that ends the 2B-point inverse, second half only: y[i + B] = ya[i] - conj(w^i) yb[i]
    ARS_HD void put_ols2(i64 seg, int i, float2 ya, float2 yb) {
        const float2 t = cmulc(yb, ARS_LDG(tw2 + i));
        const i64 fr = ((seg0 + seg) << (logF - 1)) + i;
        ols_out(fr, make_float2(ya.x - t.x, ya.y - t.y), dry_at(fr));
    }
    // chirp operand of the store, fetched early so its latency overlaps the butterfly
    template <int MODE> ARS_HD float2 pre(i64 idx) const {
        if constexpr (MODE < 0) {
            if ((mode == ST_CHIRP || mode == ST_FINAL) && idx < N) return ARS_LDG(chirp + idx);
            return make_float2(1.f, 0.f);
        } else if constexpr (MODE == ST_CHIRP || MODE == ST_FINAL) {
            return idx < N ? ARS_LDG(chirp + idx) : make_float2(1.f, 0.f);
        } else {
            return make_float2(1.f, 0.f);
        }
    }
    template <int MODE> ARS_HD void put(i64 idx, float2 v, float2 aux) {
        if constexpr (MODE < 0) {
            switch (mode) {
                case ST_PLAIN: put<ST_PLAIN>(idx, v, aux); break;
                case ST_SCALE: put<ST_SCALE>(idx, v, aux); break;
                case ST_CHIRP: put<ST_CHIRP>(idx, v, aux); break;
                case ST_FINAL: put<ST_FINAL>(idx, v, aux); break;
                case ST_OLS: put<ST_OLS>(idx, v, aux); break;
                case ST_OLS_CHIRP: put<ST_OLS_CHIRP>(idx, v, aux); break;
                case ST_OLS2: break;      // fast instantiations only
                case ST_OLSB: put<ST_OLSB>(idx, v, aux); break;
            }
        } else if constexpr (MODE == ST_PLAIN) {
            a[idx] = v;
        } else if constexpr (MODE == ST_SCALE) {
            a[idx] = cscale(v, scale);
        } else if constexpr (MODE == ST_CHIRP) {
            if (idx < N) a[idx] = cmul(v, aux);
        } else if constexpr (MODE == ST_OLS_CHIRP) {    // keep the valid half of each segment, bin k = frame - frame0
            const i64 F = (i64)1 << logF, B = F >> 1;
            const i64 t = idx & (F - 1);
            const i64 k = ((seg0 + (idx >> logF)) << (logF - 1)) + (t - B) - frame0;
            if (t >= B && k >= 0 && k < N) a[k] = cmul(v, ARS_LDG(chirp + k));
        } else if constexpr (MODE == ST_OLS) {
            const i64 F = (i64)1 << logF, B = F >> 1;
            const i64 t = idx & (F - 1);
            // (fetching the dry frame ahead of the butterfly through pre() was measured: 268 us against 181 us -- sixteen
            // more live registers in a 64-register kernel)
            if (t >= B) {
                const i64 fr = ((seg0 + (idx >> logF)) << (logF - 1)) + (t - B);               // absolute output frame
                ols_out(fr, v, dry_at(fr));
            }
        } else if constexpr (MODE == ST_OLSB) {
            const i64 j = idx >> logF;
            const i64 o = (idx & (((i64)1 << logF) - 1)) - skip;
            if (o >= 0 && o < hop) {
                const i64 fr = (seg0 + j) * hop + o;
                ols_out(fr, v, dg == 0.f ? make_float2(0.f, 0.f) : dry_at(fr));
            }
        } else if constexpr (MODE == ST_OLS2) {
            // (stored by run_tile through put_ols2 once both sub-segments are back in shared memory)
        } else {
            if (idx < N) {
                float2 y = cmul(v, aux);
                y = make_float2(y.x * scale, -y.y * scale);
                a[idx] = y;
                const unsigned m0 = abs_bits(y.x), m1 = abs_bits(y.y), m2 = abs_bits(fadd_rn(y.x, y.y));
                if (m0 > local_l) local_l = m0;
                if (m1 > local_r) local_r = m1;
                if (m2 > local_lr) local_lr = m2;
            }
        }
    }
    template <int MODE> ARS_HD void put(i64 idx, float2 v) { put<MODE>(idx, v, pre<MODE>(idx)); }
    ARS_HD void finish() {
        if ((mode == ST_FINAL || mode == ST_OLS || mode == ST_OLS2 || mode == ST_OLSB) && maxbits) {
            local_max = local_l > local_r ? local_l : local_r;
#ifdef __CUDA_ARCH__
            unsigned m[4] = {local_max, local_l, local_r, local_lr};
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) m[q] = max(m[q], __shfl_xor_sync(0xffffffffu, m[q], o));
                // thousands of tiles update the same four words: look first, only a new maximum pays for the atomic
                if ((threadIdx.x & 31) == 0 && m[q] > *reinterpret_cast<volatile unsigned*>(maxbits + q))
                    atomicMax(maxbits + q, m[q]);
            }
#else
            const unsigned m[4] = {local_max, local_l, local_r, local_lr};
            for (int q = 0; q < 4; ++q) if (m[q] > maxbits[q]) maxbits[q] = m[q];
#endif
        }
    }
};

// ------------------------------------------------------------------ twiddles --
struct Tw {
    const float2* stage;    // per-stage tables: for Ls = 2^l, entry [stage_off(l) + j*(Ls/2) + i] = w_Ls^(i * 2^j), j < 4
    const float2* lo;       // w_M^e, e < min(M, 2^15)
    const float2* hi;       // w_M^(e << 15), e < M >> 15 (null when M <= 2^15)
};
constexpr int STAGE_LOG_MAX = 13;
ARS_HD constexpr int stage_off(int l) { return 2 * ((1 << l) - 2); }
constexpr int STAGE_TABLE_ELEMS = 2 * ((1 << (STAGE_LOG_MAX + 1)) - 2);

template <bool INV> ARS_HD float2 tw_big(const Tw& tw, unsigned e) {
    float2 w = ARS_LDG(tw.lo + (e & ((1u << BIG_LO_LOG) - 1)));
    if (tw.hi) w = cmul(w, ARS_LDG(tw.hi + (e >> BIG_LO_LOG)));
    return INV ? cconj(w) : w;
}

// w[k] = w[1]^k for k < r from the exactly tabulated w[1], w[2], w[4], w[8] (at most three products deep)
template <int r> ARS_HD void expand_pow(float2 (&w)[r]) {
    if constexpr (r >= 4) w[3] = cmul(w[2], w[1]);
    if constexpr (r >= 8) { w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]); }
    if constexpr (r >= 16) {
        w[9] = cmul(w[8], w[1]);  w[10] = cmul(w[8], w[2]); w[11] = cmul(w[8], w[3]); w[12] = cmul(w[8], w[4]);
        w[13] = cmul(w[8], w[5]); w[14] = cmul(w[8], w[6]); w[15] = cmul(w[8], w[7]);
    }
}
constexpr int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// local stage twiddles w_Ls^(i*k), k < r: four coalesced table reads (lanes hold consecutive i) + products
template <int Ls, int r, bool INV> ARS_HD void stage_twiddles(const Tw& tw, int i, float2 (&w)[r]) {
    constexpr int l = ilog2(Ls);
    const float2* t = tw.stage + stage_off(l) + i;
    #pragma unroll
    for (int j = 0; (1 << j) < r; ++j) {
        float2 x = ARS_LDG(t + j * (Ls / 2));
        w[1 << j] = INV ? cconj(x) : x;
    }
    expand_pow<r>(w);
}

// -------------------------------------------------------------- tile layouts --
// Strided pass: tile = R rows (stride `stride` elements apart in HBM) x T adjacent columns.
template <int LOGR, int LOGT> struct StridedLayout {
    static constexpr int R = 1 << LOGR, T = 1 << LOGT, C = T;
    static constexpr int ROW_PITCH = T;                 // slots between consecutive rows of a column
    static constexpr bool PADDED = true;                // one spare slot per 16 elements
    static constexpr int SMEM_ELEMS = R * T + ((R * T) >> 4);
    static ARS_HD int bfly(int q) { return q >> LOGT; }
    static ARS_HD int col(int q) { return q & (T - 1); }
    static ARS_HD int sidx(int row, int c) { int i = (row << LOGT) + c; return i + (i >> 4); }
};
// Contiguous pass: tile = C whole R-point segments.
template <int LOGR, int LOGC> struct ContigLayout {
    static constexpr int R = 1 << LOGR, C = 1 << LOGC;
    static constexpr int ROW_PITCH = 1;
    static constexpr bool PADDED = true;
    static constexpr int SMEM_ELEMS = R * C + ((R * C) >> 4);
    static ARS_HD int sidx(int row, int c) { int i = (c << LOGR) + row; return i + (i >> 4); }
};

// Strided tile without padding: the lanes of a warp run along the columns of one row in every stage, so the rows can
// sit back to back -- which is the layout a bulk copy delivers (pass_last_pipe_kernel).
template <int LOGR, int LOGT> struct StridedFlat {
    static constexpr int R = 1 << LOGR, T = 1 << LOGT, C = T;
    static constexpr int ROW_PITCH = T;
    static constexpr bool PADDED = false;
    static constexpr int SMEM_ELEMS = R * T;
    static ARS_HD int bfly(int q) { return q >> LOGT; }
    static ARS_HD int col(int q) { return q & (T - 1); }
    static ARS_HD int sidx(int row, int c) { return (row << LOGT) + c; }
};

struct PassArgs {
    i64 total = 0;  // > 0: points of the whole launch -- a batch of total / M independent M-point transforms (0: one transform)
    i64 M;          // transform length
    int logM;
    int logLg;      // strided: segment length of this pass (Lg); contiguous: == LOGR
    int prefetch;   // > 0: each CTA prefetches into the L2 the tile `prefetch` tiles ahead (0: off)
    // strided passes: lane-dependent factors of the inter-pass twiddle, laid out so that lanes read adjacent
    // entries: [j*T + c] = w_Lg^(c * mul * 2^j) for j < 4, then [4*T + kb*T + c] = w_Lg^(c * kb) for kb < mul
    const float2* ptab;
    Tw tw;
};

template <int LOGR> struct Prod {   // product of the radices before stage s
    static __host__ __device__ constexpr int before(int s) {
        int p = 1;
        for (int i = 0; i < s; ++i) p *= Rad<LOGR>::r(i);
        return p;
    }
};

// digit-reversed output index of slot (b, k) in the last stage (see file header)
template <int LOGR> ARS_HD int kfull_of(int b, int k) {
    constexpr int n = Rad<LOGR>::n;
    constexpr int R = 1 << LOGR;
    constexpr int rl = Rad<LOGR>::r(n - 1);
    int kf = 0;
    int mul = 1;
    #pragma unroll
    for (int s = 0; s < n - 1; ++s) {
        const int rs = Rad<LOGR>::r(s);
        const int W = R / (Prod<LOGR>::before(s + 1) * rl);     // weight of digit k_s inside b
        int ks = (b / W) % rs;
        kf += ks * mul;
        mul *= rs;
    }
    return kf + k * mul;
}

// One DIF (forward) or DIT (inverse) stage S of the R-point column transforms of a tile.
//   GIDX_FIRST(row, c): HBM index of tile element (row, c) on the natural-order side
//   GIDX_LAST(b, k, c): HBM index of output k of last-stage butterfly b on the permuted side
template <int LOGR, int S, bool INV, bool STRIDED, int NT, class LAYOUT, int LDM, int STM, class LD, class ST, class GF,
          class GL>
ARS_HD void run_stage(float2* sm, LD& ld, ST& st, const PassArgs& pa, GF gfirst, GL glast,
                                          unsigned col0, int tid) {
    constexpr int R = 1 << LOGR;
    constexpr int n = Rad<LOGR>::n;
    constexpr int r = Rad<LOGR>::r(S);
    constexpr int Ls = R / Prod<LOGR>::before(S);
    constexpr int sub = Ls / r;
    constexpr bool first = (S == 0), last = (S == n - 1);
    constexpr int NB = R / r;                  // butterflies per column
    constexpr int TOTAL = NB * LAYOUT::C;
    static_assert(!STRIDED || NT % LAYOUT::C == 0, "a thread must stay in one tile column");

    // inter-pass twiddle w_Lg^(i * kf), kf = kb(b) + k * mul: per-thread powers P[k] = w^(i*mul*k) (the column
    // i is fixed per thread) and one per-butterfly factor Q = w^(i*kb); both from exact two-level table reads
    constexpr int mul = R / r;                 // = product of the radices before the last stage (when `last`)
    float2 P[r];
    unsigned icol = 0;
    const unsigned shift = (unsigned)(pa.logM - pa.logLg);
    const unsigned emask = (unsigned)(pa.M - 1);
    // w^(i * x) with i = col0 + c splits into a CTA-uniform factor w^(col0 * x) (broadcast table reads) and a
    // lane-dependent one w^(c * x) that comes from the pass table with adjacent lanes reading adjacent entries --
    // per-lane gathers into the big twiddle tables cost more L1 wavefronts than the tile's data did.
    const int cl = tid & (LAYOUT::C - 1);
    if constexpr (STRIDED && last) {
        icol = col0;
        const unsigned eP = (col0 * (unsigned)mul) << shift;
        #pragma unroll
        for (int j = 0; (1 << j) < r; ++j) {
            const float2 u = tw_big<INV>(pa.tw, (eP << j) & emask);
            float2 l = ARS_LDG(pa.ptab + j * LAYOUT::C + cl);
            if (INV) l = cconj(l);
            P[1 << j] = cmul(u, l);
        }
        expand_pow<r>(P);
    }

    // Index arithmetic is strength-reduced by hand: HBM indices advance by a per-stage 64-bit step, shared-memory
    // slots by a compile-time step whenever the padding term is linear in the row (row step * row pitch % 16 == 0).
    const i64 fstep = gfirst.step() * (i64)sub;            // natural side: rows are `sub` apart
    const i64 lstep = glast.step();                         // permuted side: consecutive outputs k
    constexpr int PITCH = LAYOUT::ROW_PITCH;
    constexpr bool LIN = !LAYOUT::PADDED || ((sub * PITCH) % 16) == 0;
    constexpr int SSTEP = sub * PITCH + (LAYOUT::PADDED ? (sub * PITCH) / 16 : 0);
#define ARS_SM(t_) sm[LIN ? (s0 + (t_) * SSTEP) : LAYOUT::sidx(row0 + (t_) * sub, c)]

    // the stage that reads HBM is unrolled over its butterflies so all of a thread's loads are in flight at once
    constexpr int UNR = (((!INV && first) || (INV && last)) && TOTAL >= NT) ? (TOTAL / NT) : 1;
    #pragma unroll UNR
    for (int q = tid; q < TOTAL; q += NT) {
        int b, c;
        if constexpr (STRIDED) { b = LAYOUT::bfly(q); c = LAYOUT::col(q); }
        else { b = q % NB; c = q / NB; }
        const int seg = b / sub, i = b % sub;
        const int row0 = seg * Ls + i;
        const int s0 = LAYOUT::sidx(row0, c);
        float2 v[r];
        if constexpr (!INV) {
            if constexpr (first) {
                i64 idx = gfirst(row0, c);
                if constexpr (LDM == LD_OLS_X) ld.template get_x<r>(idx, fstep, v);
                else if constexpr (LDM == LD_OLSB_X) ld.template get_xb<r>(idx, fstep, v);
                else {
                    #pragma unroll
                    for (int t = 0; t < r; ++t) { v[t] = ld.template get<LDM>(idx); idx += fstep; }
                }
            } else {
                #pragma unroll
                for (int t = 0; t < r; ++t) v[t] = ARS_SM(t);
            }
            if constexpr (!last) {
                float2 w[r];
                stage_twiddles<Ls, r, false>(pa.tw, i, w);
                Dft<r, false>::run(v);
                #pragma unroll
                for (int k = 1; k < r; ++k) v[k] = cmul(v[k], w[k]);
                #pragma unroll
                for (int k = 0; k < r; ++k) ARS_SM(k) = v[k];
            } else {
                float2 Q = make_float2(1.f, 0.f);
                if constexpr (STRIDED) {
                    const unsigned kb = (unsigned)kfull_of<LOGR>(b, 0);
                    Q = cmul(tw_big<false>(pa.tw, (icol * kb) << shift), ARS_LDG(pa.ptab + (4 + kb) * LAYOUT::C + cl));
                }
                Dft<r, false>::run(v);
                i64 idx = glast(b, c);
                #pragma unroll
                for (int k = 0; k < r; ++k) {
                    if constexpr (STRIDED) v[k] = cmul(v[k], k ? cmul(Q, P[k]) : Q);
                    st.template put<STM>(idx, v[k]);
                    idx += lstep;
                }
            }
        } else {
            if constexpr (last) {
                float2 Q = make_float2(1.f, 0.f);
                if constexpr (STRIDED) {
                    const unsigned kb = (unsigned)kfull_of<LOGR>(b, 0);
                    Q = cmul(tw_big<true>(pa.tw, (icol * kb) << shift),
                             cconj(ARS_LDG(pa.ptab + (4 + kb) * LAYOUT::C + cl)));
                }
                i64 idx = glast(b, c);
                if constexpr (LDM == LD_OLS_MAC) ld.template get_mac<r>(idx, lstep, v);
                else {
                    #pragma unroll
                    for (int k = 0; k < r; ++k) { v[k] = ld.template get<LDM>(idx); idx += lstep; }
                }
                if constexpr (STRIDED) {
                    #pragma unroll
                    for (int k = 0; k < r; ++k) v[k] = cmul(v[k], k ? cmul(Q, P[k]) : Q);
                }
            } else {
                float2 w[r];
                stage_twiddles<Ls, r, true>(pa.tw, i, w);
                #pragma unroll
                for (int k = 0; k < r; ++k) v[k] = ARS_SM(k);
                #pragma unroll
                for (int k = 1; k < r; ++k) v[k] = cmul(v[k], w[k]);
            }
            if constexpr (first && STM == ST_OLS2) {
                Dft<r, true>::run(v);
                #pragma unroll
                for (int t = 0; t < r; ++t) ARS_SM(t) = v[t];
            } else if constexpr (first && STM == ST_OLS) {
                Dft<r, true>::run(v);
                st.template put_ols<r>(gfirst(row0, c), fstep, v);
            } else if constexpr (first && STM == ST_OLSB) {
                Dft<r, true>::run(v);
                st.template put_olsb<r>(gfirst(row0, c), fstep, v);
            } else if constexpr (first) {
                float2 aux[r];
                const i64 idx0 = gfirst(row0, c);
                #pragma unroll
                for (int t = 0; t < r; ++t) aux[t] = st.template pre<STM>(idx0 + t * fstep);
                Dft<r, true>::run(v);
                i64 idx = idx0;
                #pragma unroll
                for (int t = 0; t < r; ++t) { st.template put<STM>(idx, v[t], aux[t]); idx += fstep; }
            } else {
                Dft<r, true>::run(v);
                #pragma unroll
                for (int t = 0; t < r; ++t) ARS_SM(t) = v[t];
            }
        }
    }
#undef ARS_SM
}

#define ARS_STAGE(S_) run_stage<LOGR, S_, INV, STRIDED, NT, LAYOUT, LDM, STM>(sm, ld, st, pa, gfirst, glast, col0, tid)
template <int LOGR, bool INV, bool STRIDED, int NT, class LAYOUT, int LDM, int STM, class LD, class ST, class GF, class GL>
__device__ __forceinline__ void run_tile(float2* sm, LD& ld, ST& st, const PassArgs& pa, GF gfirst, GL glast,
                                         unsigned col0) {
    constexpr int n = Rad<LOGR>::n;
    const int tid = (int)threadIdx.x;
    if constexpr (!INV) {
        ARS_STAGE(0);
        if constexpr (n > 1) { __syncthreads(); ARS_STAGE(1); }
        if constexpr (n > 2) { __syncthreads(); ARS_STAGE(2); }
        if constexpr (n > 3) { __syncthreads(); ARS_STAGE(3); }
    } else {
        if constexpr (n > 3) { ARS_STAGE(3); __syncthreads(); }
        if constexpr (n > 2) { ARS_STAGE(2); __syncthreads(); }
        if constexpr (n > 1) { ARS_STAGE(1); __syncthreads(); }
        ARS_STAGE(0);
        if constexpr (STM == ST_OLS2) {            // closing radix-2 stage across the tile's two sub-segments
            static_assert(STM != ST_OLS2 || (LAYOUT::C == 2 && !STRIDED), "ST_OLS2: a tile is the two halves of one segment");
            __syncthreads();
            const i64 seg = gfirst(0, 0) >> (LOGR + 1);
            for (int i = tid; i < (1 << LOGR); i += NT) st.put_ols2(seg, i, sm[LAYOUT::sidx(i, 0)], sm[LAYOUT::sidx(i, 1)]);
        }
    }
    st.finish();
}

// L2 prefetch of a tile a full wave ahead: the CTA that will process tile `blockIdx.x + dist` finds its loads in
// the L2 instead of waiting on HBM, so the memory pipeline is as deep as the L2 allows rather than as deep as the
// resident CTAs' registers.  Same element-to-thread map as the first executed stage of run_tile.
template <int LOGR, bool INV, bool STRIDED, int NT, class LAYOUT, class GF, class GL>
__device__ __forceinline__ void prefetch_tile(const float2* a, const float2* b, GF gfirst, GL glast) {
    constexpr int n = Rad<LOGR>::n;
    constexpr int S = INV ? n - 1 : 0;                    // the stage that reads HBM
    constexpr int R = 1 << LOGR;
    constexpr int r = Rad<LOGR>::r(S);
    constexpr int Ls = R / Prod<LOGR>::before(S);
    constexpr int sub = Ls / r;
    constexpr int NB = R / r;
    constexpr int TOTAL = NB * LAYOUT::C;
    const i64 step = INV ? glast.step() : gfirst.step() * (i64)sub;
    for (int q = (int)threadIdx.x; q < TOTAL; q += NT) {
        int bf, c;
        if constexpr (STRIDED) { bf = LAYOUT::bfly(q); c = LAYOUT::col(q); }
        else { bf = q % NB; c = q / NB; }
        i64 idx = INV ? glast(bf, c) : gfirst((bf / sub) * Ls + (bf % sub), c);
        #pragma unroll
        for (int t = 0; t < r; ++t) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a + idx));
            if (b) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + idx));
            idx += step;
        }
    }
}

// Host emulation of one tile (tests/host_emul): the same stage code, "threads" run one after
// another, a barrier is simply the end of the loop over threads.
template <int LOGR, bool INV, bool STRIDED, int NT, class LAYOUT, int LDM, int STM, class LD, class ST, class GF, class GL>
inline void emulate_tile(float2* sm, LD& ld, ST& st, const PassArgs& pa, GF gfirst, GL glast, unsigned col0) {
    constexpr int n = Rad<LOGR>::n;
#define ARS_ALL(S_) for (int tid = 0; tid < NT; ++tid) ARS_STAGE(S_)
    if constexpr (!INV) {
        ARS_ALL(0);
        if constexpr (n > 1) { ARS_ALL(1); }
        if constexpr (n > 2) { ARS_ALL(2); }
        if constexpr (n > 3) { ARS_ALL(3); }
    } else {
        if constexpr (n > 3) { ARS_ALL(3); }
        if constexpr (n > 2) { ARS_ALL(2); }
        if constexpr (n > 1) { ARS_ALL(1); }
        ARS_ALL(0);
        if constexpr (STM == ST_OLS2) {
            const i64 seg = gfirst(0, 0) >> (LOGR + 1);
            for (int i = 0; i < (1 << LOGR); ++i) st.put_ols2(seg, i, sm[LAYOUT::sidx(i, 0)], sm[LAYOUT::sidx(i, 1)]);
        }
    }
#undef ARS_ALL
    st.finish();
}

// ------------------------------------------------------------------- kernels --
// HBM index maps of a tile.  "first" = natural-order side, "last" = permuted side.
template <int LOGR> struct StridedFirst {
    i64 base; int logStride;
    ARS_HD i64 operator()(int row, int c) const { return base + ((i64)row << logStride) + c; }
    ARS_HD i64 step() const { return (i64)1 << logStride; }          // HBM index step per row
};
template <int LOGR> struct StridedLast {
    i64 base; int logStride;
    ARS_HD i64 operator()(int b, int c) const {                       // output k = 0 of last-stage butterfly b
        constexpr int rl = Rad<LOGR>::r(Rad<LOGR>::n - 1);
        return base + ((i64)(b * rl) << logStride) + c;
    }
    ARS_HD i64 step() const { return (i64)1 << logStride; }          // per output k
};
template <int LOGR> struct ContigFirst {
    i64 base;
    ARS_HD i64 operator()(int row, int c) const { return base + ((i64)c << LOGR) + row; }
    ARS_HD constexpr i64 step() const { return 1; }
};
// permuted side of the contiguous pass: output k of last-stage butterfly b is parked at
// k*(R/rl) + b (not b*rl + k) so that adjacent lanes touch adjacent addresses
template <int LOGR> struct ContigLast {
    i64 base;
    ARS_HD i64 operator()(int b, int c) const { return base + ((i64)c << LOGR) + b; }
    ARS_HD constexpr i64 step() const { return (1 << LOGR) / Rad<LOGR>::r(Rad<LOGR>::n - 1); }
};

// tile (g, cb) of a strided pass over segments of length Lg = 2^logLg covers rows
// t = 0..R-1 at  g*Lg + t*(Lg/R) + cb*T + c
template <int LOGR, int LOGT> struct StridedTile {
    i64 base; unsigned col0; int logStride;
    ARS_HD StridedTile(i64 tile, const PassArgs& pa) {
        logStride = pa.logLg - LOGR;
        const int sh = logStride - LOGT;
        const i64 g = tile >> sh;
        const unsigned cb = (unsigned)(tile & (((i64)1 << sh) - 1));
        col0 = cb << LOGT;
        base = (g << pa.logLg) + col0;
    }
};

template <int LOGR, int LOGT, bool INV, int NT, int LDM, int STM>
__global__ void __launch_bounds__(NT, NT >= 512 ? 2 : 1) pass_strided_kernel(Ld ld, St st, PassArgs pa) {
    extern __shared__ float2 sm[];
    if constexpr (LDM == LD_PLAIN) {
        if (pa.prefetch > 0 && (i64)blockIdx.x + pa.prefetch < (i64)gridDim.x) {
            const StridedTile<LOGR, LOGT> t2((i64)blockIdx.x + pa.prefetch, pa);
            prefetch_tile<LOGR, INV, true, NT, StridedLayout<LOGR, LOGT>>(
                ld.a, nullptr, StridedFirst<LOGR>{t2.base, t2.logStride}, StridedLast<LOGR>{t2.base, t2.logStride});
        }
    }
    const StridedTile<LOGR, LOGT> t((i64)blockIdx.x, pa);
    run_tile<LOGR, INV, true, NT, StridedLayout<LOGR, LOGT>, LDM, STM>(
        sm, ld, st, pa, StridedFirst<LOGR>{t.base, t.logStride}, StridedLast<LOGR>{t.base, t.logStride}, t.col0);
}

// Contiguous pass: tile = C whole segments of R adjacent elements.
template <int LOGR, int LOGC, bool INV, int NT, int LDM, int STM>
__global__ void __launch_bounds__(NT, NT >= 512 ? 2 : (NT == 256 ? 3 : 1)) pass_contig_kernel(Ld ld, St st, PassArgs pa) {
    extern __shared__ float2 sm[];
    if constexpr (LDM == LD_PLAIN || LDM == LD_MULSPEC) {
        if (pa.prefetch > 0 && (i64)blockIdx.x + pa.prefetch < (i64)gridDim.x) {
            const i64 base2 = ((i64)blockIdx.x + pa.prefetch) << (LOGR + LOGC);
            prefetch_tile<LOGR, INV, false, NT, ContigLayout<LOGR, LOGC>>(
                ld.a, LDM == LD_MULSPEC ? ld.b : nullptr, ContigFirst<LOGR>{base2}, ContigLast<LOGR>{base2});
        }
    }
    if constexpr (LDM == LD_OLS_X) {
        if (pa.prefetch > 0 && (i64)blockIdx.x + pa.prefetch < (i64)gridDim.x)
            ld.prefetch_x(((i64)blockIdx.x + pa.prefetch) << LOGC, 1 << LOGC, threadIdx.x, NT);
    }
    const i64 base = (i64)blockIdx.x << (LOGR + LOGC);
    run_tile<LOGR, INV, false, NT, ContigLayout<LOGR, LOGC>, LDM, STM>(sm, ld, st, pa, ContigFirst<LOGR>{base},
                                                             ContigLast<LOGR>{base}, 0u);
}


// ------------------------------------------------- fused middle pass (big-block overlap-save, upols.cu) ---
// The last forward pass and the first inverse pass of an M-point transform are both the contiguous pass over the
// same tile, with the spectrum product in between -- so they run as ONE kernel: forward stages, product with the IR
// spectrum in registers (the forward's last butterfly leaves exactly the values the inverse's first butterfly takes),
// inverse stages; one trip through shared memory instead of two through HBM, and no launch in between.
//   plain  : Y = Z H                                   (one real IR for both channels of Z = FFT(L + iR))
//   MIRROR : Y = Z A + conj(Z[F - k]) Bc               (stereo IR: A = (H_L + H_R)/2, Bc = (H_L - H_R)/2, upols.cu)
// The mirror bin F - k of (first-pass digit k1, second-pass digit k2) is (R1 - k1, R2 - 1 - k2) for k1 != 0: it lives in
// the segment of digit R1 - k1, and complementing every digit of k2 reverses the position inside the segment, whatever the
// parking order.  A MIRROR tile therefore holds the two segments of a digit pair (k1, R1 - k1); the pair (0, R1/2) is its
// own mirror: segment R1/2 reverses in itself, segment 0 maps k2 -> (R2 - k2) mod R2 through the digit maps below.
struct MidArgs {
    const float2* h0 = nullptr;     // H (plain) or A (MIRROR): permuted order, pre-scaled by 1/F, F points
    const float2* h1 = nullptr;     // MIRROR: Bc
    i64 fmask = 0;                  // F - 1
    const int* rho = nullptr;       // MIRROR: row (segment) that holds first-pass digit k1, R1 entries
    int logR1 = 0;                  // MIRROR: log2 of the first (strided) pass's length
};

// slot (row inside the tile column, in-place order b * rl + k) of second-pass output digit kf: inverse of kfull_of
template <int LOGR> ARS_HD int inplace_row_of(int kf) {
    constexpr int n = Rad<LOGR>::n;
    constexpr int R = 1 << LOGR;
    constexpr int rl = Rad<LOGR>::r(n - 1);
    int b = 0;
    #pragma unroll
    for (int s = 0; s < n - 1; ++s) {
        const int rs = Rad<LOGR>::r(s);
        const int W = R / (Prod<LOGR>::before(s + 1) * rl);
        b += (kf % rs) * W;
        kf /= rs;
    }
    return b * rl + kf;
}

template <int LOGR> struct PairFirst {          // a tile made of two arbitrary segments
    i64 b0, b1;
    ARS_HD i64 operator()(int row, int c) const { return (c ? b1 : b0) + row; }
    ARS_HD constexpr i64 step() const { return 1; }
};
template <int LOGR> struct PairLast {
    i64 b0, b1;
    ARS_HD i64 operator()(int b, int c) const { return (c ? b1 : b0) + b; }
    ARS_HD constexpr i64 step() const { return (1 << LOGR) / Rad<LOGR>::r(Rad<LOGR>::n - 1); }
};

template <int LOGR, int LOGC, bool MIRROR> struct MidStage {
    using LAYOUT = ContigLayout<LOGR, LOGC>;
    static constexpr int n = Rad<LOGR>::n;
    static constexpr int R = 1 << LOGR;
    static constexpr int r = Rad<LOGR>::r(n - 1);
    static constexpr int NB = R / r;
    static constexpr int TOTAL = NB * LAYOUT::C;
    // forward last stage of butterfly q; MIRROR: the spectrum goes back to the butterfly's own rows for the partners to read
    static ARS_HD void fwd(float2* sm, int q, float2 (&v)[r]) {
        const int b = q % NB, c = q / NB;
        #pragma unroll
        for (int t = 0; t < r; ++t) v[t] = sm[LAYOUT::sidx(b * r + t, c)];
        Dft<r, false>::run(v);
        if constexpr (MIRROR) {
            #pragma unroll
            for (int k = 0; k < r; ++k) sm[LAYOUT::sidx(b * r + k, c)] = v[k];
        }
    }
    // spectrum product (v <- Y)
    template <class GL> static ARS_HD void mul(const float2* sm, const MidArgs& ma, GL glast, int q, bool selfm, float2 (&v)[r]) {
        const int b = q % NB, c = q / NB;
        i64 idx = glast(b, c);
        #pragma unroll
        for (int k = 0; k < r; ++k) {
            const i64 h = idx & ma.fmask;
            if constexpr (!MIRROR) {
                v[k] = cmul(v[k], ARS_LDG(ma.h0 + h));
            } else {
                int prow = R - 1 - (b * r + k), pc = 1 - c;
                if (selfm) {
                    pc = c;
                    if (c == 0) prow = inplace_row_of<LOGR>((R - kfull_of<LOGR>(b, k)) & (R - 1));
                }
                const float2 zm = sm[LAYOUT::sidx(prow, pc)];
                const float2 t0 = cmul(v[k], ARS_LDG(ma.h0 + h));
                const float2 t1 = cmul(cconj(zm), ARS_LDG(ma.h1 + h));
                v[k] = make_float2(t0.x + t1.x, t0.y + t1.y);
            }
            idx += glast.step();
        }
    }
    // inverse last stage (no twiddle: its sub-transforms have length r)
    static ARS_HD void inv(float2* sm, int q, float2 (&v)[r]) {
        const int b = q % NB, c = q / NB;
        Dft<r, true>::run(v);
        #pragma unroll
        for (int t = 0; t < r; ++t) sm[LAYOUT::sidx(b * r + t, c)] = v[t];
    }
};

// tile -> segment bases of the middle pass
template <int LOGR, int LOGC, bool MIRROR> struct MidTile {
    i64 b0, b1;
    bool selfm;
    ARS_HD MidTile(i64 tile, const MidArgs& ma) {
        if constexpr (!MIRROR) {
            b0 = tile << (LOGR + LOGC);
            b1 = b0 + ((i64)1 << LOGR);
            selfm = false;
        } else {
            static_assert(!MIRROR || LOGC == 1, "a mirror tile is a pair of segments");
            const int half = 1 << (ma.logR1 - 1);
            const int tau = (int)(tile & (half - 1));
            const i64 j = tile >> (ma.logR1 - 1);
            selfm = tau == 0;
            const int sa = ma.rho[tau], sb = ma.rho[selfm ? half : 2 * half - tau];
            b0 = ((j << ma.logR1) + sa) << LOGR;
            b1 = ((j << ma.logR1) + sb) << LOGR;
        }
    }
};

#define ARS_MID_STAGE(S_, INV_, LDM_, STM_) \
    run_stage<LOGR, S_, INV_, false, NT, LAYOUT, LDM_, STM_>(sm, ld, st, pa, gfirst, glast, 0u, tid)
template <int LOGR, int LOGC, int NT, bool MIRROR>
__global__ void __launch_bounds__(NT, NT >= 512 ? 2 : 4) pass_mid_kernel(Ld ld, St st, PassArgs pa, MidArgs ma) {
    extern __shared__ float2 sm[];
    using LAYOUT = ContigLayout<LOGR, LOGC>;
    using MS = MidStage<LOGR, LOGC, MIRROR>;
    constexpr int n = Rad<LOGR>::n;
    static_assert(n == 3 && MS::TOTAL == NT, "middle pass: three-stage tile, one last-stage butterfly per thread");
    const int tid = (int)threadIdx.x;
    const MidTile<LOGR, LOGC, MIRROR> mt((i64)blockIdx.x, ma);
    const PairFirst<LOGR> gfirst{mt.b0, mt.b1};
    const PairLast<LOGR> glast{mt.b0, mt.b1};
    ARS_MID_STAGE(0, false, LD_PLAIN, ST_PLAIN);
    __syncthreads();
    ARS_MID_STAGE(1, false, LD_PLAIN, ST_PLAIN);
    __syncthreads();
    float2 v[MS::r];
    MS::fwd(sm, tid, v);
    if constexpr (MIRROR) __syncthreads();
    MS::mul(sm, ma, glast, tid, mt.selfm, v);
    if constexpr (MIRROR) __syncthreads();
    MS::inv(sm, tid, v);
    __syncthreads();
    ARS_MID_STAGE(1, true, LD_PLAIN, ST_PLAIN);
    __syncthreads();
    ARS_MID_STAGE(0, true, LD_PLAIN, ST_PLAIN);
}

// ---- persistent, prefetching form of the plain middle pass ------------------------------------------------------
// One CTA keeps a padded work tile plus an unpadded landing buffer in shared memory and walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...  The next tile's segment (one contiguous 2^LOGR-point run of W) is fetched by a single
// bulk asynchronous copy (cp.async.bulk, completion counted on an mbarrier) issued as soon as the first stage of the
// current tile has drained the landing buffer, so the fetch overlaps the remaining five stages and the store of the
// current tile; no thread ever waits on a global load.
struct LdLanding {
    const float2* s;                // landing buffer (shared memory), element i = point i of the tile's segment
    template <int MODE> ARS_HD float2 get(i64 idx) const { return s[(int)idx]; }
};
struct TileFirst {
    ARS_HD i64 operator()(int row, int) const { return row; }
    ARS_HD i64 step() const { return 1; }
};
#ifdef __CUDACC__
namespace bulk {
__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arm the barrier with the byte count of the copy that follows, then start the copy (one thread)
__device__ __forceinline__ void fetch(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy reads of dst are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(saddr(bar)), "r"(parity) : "memory");
}
}  // namespace bulk

// copy without arming (the caller armed the barrier with the sum of the rows' bytes)
__device__ __forceinline__ void bulk_row(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(bulk::saddr(dst)), "l"(src), "r"(bytes), "r"(bulk::saddr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_arm(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bulk::saddr(bar)), "r"(bytes) : "memory");
}

// ---- persistent, prefetching form of the last (strided inverse) pass of the big-block overlap-save transforms ----
// A CTA owns two unpadded tiles of R rows x T columns and walks tiles blockIdx.x, + gridDim.x, ...  One warp fetches a
// tile as R bulk copies of one row each (T * 8 bytes, counted on the buffer's mbarrier) two tiles ahead, so the whole
// load of tile i + 2 overlaps the arithmetic and the stores of tiles i and i + 1.  The first executed stage reads its
// butterflies from the landing rows and writes them back in place.
template <int LOGR, int LOGT> struct FlatLast {
    ARS_HD i64 operator()(int b, int c) const {
        constexpr int rl = Rad<LOGR>::r(Rad<LOGR>::n - 1);
        return (i64)(((b * rl) << LOGT) + c);
    }
    ARS_HD i64 step() const { return (i64)1 << LOGT; }
};
template <int LOGR, int LOGT, int NT, int PER_SM>
__global__ void __launch_bounds__(NT, PER_SM) pass_last_pipe_kernel(Ld ld, St st, PassArgs pa, int tiles) {
    extern __shared__ __align__(128) float2 sm[];
    __shared__ __align__(8) unsigned long long bar[2];
    using LAYOUT = StridedFlat<LOGR, LOGT>;
    constexpr int R = 1 << LOGR, T = 1 << LOGT, ELEMS = R * T;
    constexpr unsigned ROW_BYTES = (unsigned)(T * sizeof(float2));
    static_assert(Rad<LOGR>::n == 2, "pipelined last pass: two-stage column transforms");
    static_assert(T >= 32 && ROW_BYTES % 16 == 0, "a warp stays inside one row");
    const int tid = (int)threadIdx.x;
    const int G = (int)gridDim.x;
    auto fetch = [&](int tile, int p) {           // warp 0
        const StridedTile<LOGR, LOGT> t((i64)tile, pa);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tid == 0) bulk_arm(&bar[p], ROW_BYTES * (unsigned)R);
        __syncwarp();
        for (int row = tid; row < R; row += 32)
            bulk_row(sm + p * ELEMS + row * T, ld.a + t.base + ((i64)row << t.logStride), ROW_BYTES, &bar[p]);
    };
    if (tid == 0) {
        bulk::bar_init(&bar[0], 1);
        bulk::bar_init(&bar[1], 1);
    }
    __syncthreads();
    if (tid < 32) {
        if ((int)blockIdx.x < tiles) fetch((int)blockIdx.x, 0);
        if ((int)blockIdx.x + G < tiles) fetch((int)blockIdx.x + G, 1);
    }
    unsigned phase = 0;
    int p = 0;
    const FlatLast<LOGR, LOGT> flast{};
    for (int tile = (int)blockIdx.x; tile < tiles; tile += G, p ^= 1) {
        float2* buf = sm + p * ELEMS;
        const StridedTile<LOGR, LOGT> t((i64)tile, pa);
        const StridedFirst<LOGR> gfirst{t.base, t.logStride};
        LdLanding ll{buf};
        bulk::wait(&bar[p], (phase >> p) & 1u);
        phase ^= 1u << p;
        run_stage<LOGR, 1, true, true, NT, LAYOUT, LD_PLAIN, ST_OLSB>(buf, ll, st, pa, gfirst, flast, t.col0, tid);
        __syncthreads();
        run_stage<LOGR, 0, true, true, NT, LAYOUT, LD_PLAIN, ST_OLSB>(buf, ll, st, pa, gfirst, flast, t.col0, tid);
        __syncthreads();
        if (tid < 32 && tile + 2 * G < tiles) fetch(tile + 2 * G, p);
    }
    st.finish();
}

#define ARS_MIDP_STAGE(S_, INV_, LD_, GF_) \
    run_stage<LOGR, S_, INV_, false, NT, LAYOUT, LD_PLAIN, ST_PLAIN>(sm, LD_, st, pa, GF_, glast, 0u, tid)
template <int LOGR, int NT, int PER_SM>
__global__ void __launch_bounds__(NT, PER_SM) pass_mid_pipe_kernel(Ld ld, St st, PassArgs pa, MidArgs ma, int tiles) {
    extern __shared__ __align__(128) float2 sm[];
    __shared__ __align__(8) unsigned long long bar;
    using LAYOUT = ContigLayout<LOGR, 0>;
    using MS = MidStage<LOGR, 0, false>;
    constexpr int R = 1 << LOGR;
    constexpr unsigned BYTES = (unsigned)(R * sizeof(float2));
    static_assert(Rad<LOGR>::n == 3 && MS::TOTAL == NT, "middle pass: three-stage tile, one last-stage butterfly per thread");
    static_assert((LAYOUT::SMEM_ELEMS * sizeof(float2)) % 128 == 0, "landing buffer alignment");
    float2* land = sm + LAYOUT::SMEM_ELEMS;
    const int tid = (int)threadIdx.x;
    int tile = (int)blockIdx.x;
    if (tid == 0) {
        bulk::bar_init(&bar, 1);
        if (tile < tiles) bulk::fetch(land, ld.a + ((i64)tile << LOGR), BYTES, &bar);
    }
    __syncthreads();
    LdLanding ll{land};
    const TileFirst tfirst{};
    unsigned parity = 0;
    for (; tile < tiles; tile += (int)gridDim.x) {
        const i64 b0 = (i64)tile << LOGR;
        const PairFirst<LOGR> gfirst{b0, b0};
        const PairLast<LOGR> glast{b0, b0};
        bulk::wait(&bar, parity);
        parity ^= 1u;
        ARS_MIDP_STAGE(0, false, ll, tfirst);
        __syncthreads();
        const int nxt = tile + (int)gridDim.x;
        if (tid == 0 && nxt < tiles) bulk::fetch(land, ld.a + ((i64)nxt << LOGR), BYTES, &bar);
        ARS_MIDP_STAGE(1, false, ld, gfirst);
        __syncthreads();
        float2 v[MS::r];
        MS::fwd(sm, tid, v);
        MS::mul(sm, ma, glast, tid, false, v);
        MS::inv(sm, tid, v);
        __syncthreads();
        ARS_MIDP_STAGE(1, true, ld, gfirst);
        __syncthreads();
        ARS_MIDP_STAGE(0, true, ld, gfirst);
        __syncthreads();            // the work tile is rewritten by the next tile's first stage
    }
}
#undef ARS_MIDP_STAGE
#endif

// host emulation of one middle-pass tile (tests/host_emul): per-"thread" registers live in `regs` across the barriers
template <int LOGR, int LOGC, int NT, bool MIRROR>
inline void emulate_mid_tile(float2* sm, float2* regs, Ld& ld, St& st, const PassArgs& pa, const MidArgs& ma, i64 tile) {
    using LAYOUT = ContigLayout<LOGR, LOGC>;
    using MS = MidStage<LOGR, LOGC, MIRROR>;
    const MidTile<LOGR, LOGC, MIRROR> mt(tile, ma);
    const PairFirst<LOGR> gfirst{mt.b0, mt.b1};
    const PairLast<LOGR> glast{mt.b0, mt.b1};
    typedef float2 (&VR)[MS::r];
    for (int tid = 0; tid < NT; ++tid) ARS_MID_STAGE(0, false, LD_PLAIN, ST_PLAIN);
    for (int tid = 0; tid < NT; ++tid) ARS_MID_STAGE(1, false, LD_PLAIN, ST_PLAIN);
    for (int q = 0; q < MS::TOTAL; ++q) MS::fwd(sm, q, reinterpret_cast<VR>(*(regs + (size_t)q * MS::r)));
    for (int q = 0; q < MS::TOTAL; ++q) MS::mul(sm, ma, glast, q, mt.selfm, reinterpret_cast<VR>(*(regs + (size_t)q * MS::r)));
    for (int q = 0; q < MS::TOTAL; ++q) MS::inv(sm, q, reinterpret_cast<VR>(*(regs + (size_t)q * MS::r)));
    for (int tid = 0; tid < NT; ++tid) ARS_MID_STAGE(1, true, LD_PLAIN, ST_PLAIN);
    for (int tid = 0; tid < NT; ++tid) ARS_MID_STAGE(0, true, LD_PLAIN, ST_PLAIN);
}
#undef ARS_MID_STAGE

// row (segment) in which a strided first pass of 2^logR1 points leaves its output digit k1 (StridedLast: b * rl + k)
inline int strided_row_of(int logR1, int k1) {
    switch (logR1) {
#define ARS_ROW(L) case L: return inplace_row_of<L>(k1);
        ARS_ROW(1) ARS_ROW(2) ARS_ROW(3) ARS_ROW(4) ARS_ROW(5) ARS_ROW(6) ARS_ROW(7) ARS_ROW(8) ARS_ROW(9) ARS_ROW(10)
        ARS_ROW(11) ARS_ROW(12) ARS_ROW(13)
#undef ARS_ROW
    }
    return k1;
}

}  // namespace fft

// instantiated (logR, logT|logC) pass variants; the launcher and the host emulator share the list
#define ARS_STRIDED_CASES(X) X(6, 7) X(7, 6) X(8, 5) X(9, 4) X(10, 3) X(11, 2) X(12, 1) X(12, 2)
#define ARS_CONTIG_CASES(X)                                                                         \
    X(1, 0) X(2, 0) X(3, 0) X(4, 0) X(5, 0) X(6, 0) X(7, 0) X(8, 0) X(9, 0) X(10, 0) X(11, 0) X(12, 0) \
    X(6, 7) X(7, 6) X(8, 5) X(9, 4) X(10, 3) X(11, 2) X(12, 1) X(13, 0)

// variants that also get compile-time-mode ("fast") instantiations: the ones big transforms are planned with
#define ARS_FAST_STRIDED(X) X(6, 7) X(7, 6) X(8, 5) X(9, 4)
#define ARS_FAST_CONTIG(X) X(12, 1) X(13, 0)

// ------------------------------------------------------------------ host API --
struct FftPass {
    bool strided;
    int logR;
    int logT;       // strided: log2 columns per tile; contiguous: log2 segments per tile
    int logLg;
    const float2* ptab = nullptr;   // strided: PassArgs::ptab (device), built with the plan
};
// last radix of the stage split of an R = 2^logR point column transform
inline int last_radix(int logR) {
    switch (logR) {
#define ARS_LR(L) case L: return fft::Rad<L>::r(fft::Rad<L>::n - 1);
        ARS_LR(1) ARS_LR(2) ARS_LR(3) ARS_LR(4) ARS_LR(5) ARS_LR(6) ARS_LR(7) ARS_LR(8) ARS_LR(9) ARS_LR(10)
        ARS_LR(11) ARS_LR(12) ARS_LR(13)
#undef ARS_LR
    }
    return 2;
}
// entries of a strided pass table (PassArgs::ptab): (4 + mul) * T with mul = R / last radix
inline int pass_table_mul(int logR) { return (1 << logR) / last_radix(logR); }
inline int pass_table_elems(int logR, int logT) { return (4 + pass_table_mul(logR)) << logT; }

struct FftPlan {
    int logM = 0;
    i64 M = 0;
    std::vector<FftPass> passes;   // forward order
    DevBuf tw_lo, tw_hi;
    std::vector<DevBuf> pass_tabs;  // per strided pass: PassArgs::ptab
    DevBuf rho;                     // big-block overlap-save, stereo IR: MidArgs::rho (built on first use)
    DevBuf pipe_tab;                // pipelined last pass: PassArgs::ptab for its tile width (built on first use)
    fft::Tw tw{};
};

std::vector<FftPass> fft_decompose(int logM);
FftPlan* get_fft_plan(int logM);

// Forward transform: first pass reads through `ld`, everything else works in `work`
// (M complex), last pass writes through `st` (permuted order).
void fft_forward(FftPlan* p, const fft::Ld& ld, float2* work, const fft::St& st);
// Inverse transform: first pass reads through `ld` (permuted order), last pass writes natural
// order through `st`.  Unnormalised.
void fft_inverse(FftPlan* p, const fft::Ld& ld, float2* work, const fft::St& st);
// `nseg` independent 8192-point overlap-save transforms as a radix-2 stage folded into the load / store plus two
// 4096-point transforms per segment (LD_OLS_X2 | LD_OLS_IR2 forward, ST_OLS2 inverse; fills in Ld::tw2 / St::tw2)
void fft_segments_r2(i64 nseg, fft::Ld ld, fft::St st, bool inverse);
// Big-block overlap-save (upols.cu): `nbatch` M-point transforms back to back in `work`; first = the strided forward
// pass (reads through ld), mid = contiguous forward pass x spectrum (h1: the mirror form for stereo IRs) x contiguous
// inverse pass in place, last = the strided inverse pass (writes through st).
void fft_batch_first(FftPlan* p, i64 nbatch, const fft::Ld& ld, const fft::St& st);
void fft_batch_mid(FftPlan* p, i64 nbatch, float2* work, const float2* h0, const float2* h1);
void fft_batch_last(FftPlan* p, i64 nbatch, const fft::Ld& ld, const fft::St& st);
void fft_touch_tables();      // builds the shared stage table on the current stream if it does not exist yet
// One contiguous pass over `nseg` independent 2^logF-point segments (logF = 12 or 13; nseg a multiple of
// fft_segment_tile(logF)): the block transforms of the overlap-save convolution.
int fft_segment_tile(int logF);
void fft_segments(int logF, i64 nseg, const fft::Ld& ld, const fft::St& st, bool inverse);

int next_pow2_log(i64 n);

}  // namespace ars
