// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh).
#include "fft_launch.cuh"

namespace ars {
namespace fftk {

// inverse passes: PLAIN -> PLAIN ; inverse last pass: PLAIN -> ST_CHIRP | ST_FINAL
bool fast_strided_inv(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa) {
    const int lm = ld.mode, sm = st.mode;
    if (lm != LD_PLAIN) return false;
#define F_CASE(R, T)                                                                                             \
    if (ps.logR == R && ps.logT == T) {                                                                          \
        if (sm == ST_PLAIN) { launch_strided<R, T, true, LD_PLAIN, ST_PLAIN>(ld, st, pa); return true; }       \
        if (sm == ST_CHIRP) { launch_strided<R, T, true, LD_PLAIN, ST_CHIRP>(ld, st, pa); return true; }       \
        if (sm == ST_FINAL) { launch_strided<R, T, true, LD_PLAIN, ST_FINAL>(ld, st, pa); return true; }       \
        if (sm == ST_OLSB) { launch_strided<R, T, true, LD_PLAIN, ST_OLSB>(ld, st, pa); return true; }         \
    }
    ARS_FAST_STRIDED(F_CASE)
#undef F_CASE
    if (ps.logR == 10 && ps.logT == 3 && sm == ST_OLSB) { launch_strided<10, 3, true, LD_PLAIN, ST_OLSB>(ld, st, pa); return true; }
    return false;
}

}  // namespace fftk
}  // namespace ars
