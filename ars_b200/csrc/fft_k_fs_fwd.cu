// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh).
#include "fft_launch.cuh"

namespace ars {
namespace fftk {

// forward first pass: LD_CHIRP_* -> ST_PLAIN ; forward other passes: PLAIN -> PLAIN
bool fast_strided_fwd(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa) {
    const int lm = ld.mode, sm = st.mode;
    if (sm != ST_PLAIN) return false;
#define F_CASE(R, T)                                                                                                  \
    if (ps.logR == R && ps.logT == T) {                                                                               \
        if (lm == LD_PLAIN) { launch_strided<R, T, false, LD_PLAIN, ST_PLAIN>(ld, st, pa); return true; }           \
        if (lm == LD_CHIRP_X2) { launch_strided<R, T, false, LD_CHIRP_X2, ST_PLAIN>(ld, st, pa); return true; }     \
        if (lm == LD_CHIRP_XC) { launch_strided<R, T, false, LD_CHIRP_XC, ST_PLAIN>(ld, st, pa); return true; }     \
        if (lm == LD_CHIRP_PAIR) { launch_strided<R, T, false, LD_CHIRP_PAIR, ST_PLAIN>(ld, st, pa); return true; } \
        if (lm == LD_CHIRP_C) { launch_strided<R, T, false, LD_CHIRP_C, ST_PLAIN>(ld, st, pa); return true; }       \
        if (lm == LD_OLSB_X) { launch_strided<R, T, false, LD_OLSB_X, ST_PLAIN>(ld, st, pa); return true; }         \
        if (lm == LD_TAPS) { launch_strided<R, T, false, LD_TAPS, ST_PLAIN>(ld, st, pa); return true; }             \
    }
    ARS_FAST_STRIDED(F_CASE)
#undef F_CASE
    // 2^22-point big-block overlap-save transforms: 2^10 x 2^12
    if (ps.logR == 10 && ps.logT == 3 && lm == LD_OLSB_X) { launch_strided<10, 3, false, LD_OLSB_X, ST_PLAIN>(ld, st, pa); return true; }
    if (ps.logR == 10 && ps.logT == 3 && lm == LD_TAPS) { launch_strided<10, 3, false, LD_TAPS, ST_PLAIN>(ld, st, pa); return true; }
    return false;
}

}  // namespace fftk
}  // namespace ars
