// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh).
#include "fft_launch.cuh"

namespace ars {
namespace fftk {

bool fast_contig_fwd(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa) {
    const int lm = ld.mode, sm = st.mode;
    if (sm == ST_SCALE && lm == LD_OLS_IR && ps.logR == 13 && ps.logT == 0) {      // IR partition spectra (upols.cu K2)
        launch_contig<13, 0, false, LD_OLS_IR, ST_SCALE>(ld, st, pa);
        return true;
    }
    if (ps.logR == 12 && ps.logT == 1) {          // radix-2-folded overlap-save transforms (fft_segments_r2)
        if (lm == LD_OLS_X2 && sm == ST_PLAIN) { launch_contig<12, 1, false, LD_OLS_X2, ST_PLAIN>(ld, st, pa); return true; }
        if (lm == LD_OLS_IR2 && sm == ST_SCALE) { launch_contig<12, 1, false, LD_OLS_IR2, ST_SCALE>(ld, st, pa); return true; }
    }
    if (sm == ST_SCALE && lm == LD_PLAIN && ps.logR == 12 && ps.logT == 1) {       // IR spectrum of the big-block route
        launch_contig<12, 1, false, LD_PLAIN, ST_SCALE>(ld, st, pa);
        return true;
    }
    if (sm != ST_PLAIN) return false;
#define F_CASE(R, C)                                                                                          \
    if (ps.logR == R && ps.logT == C) {                                                                       \
        if (lm == LD_PLAIN) { launch_contig<R, C, false, LD_PLAIN, ST_PLAIN>(ld, st, pa); return true; }    \
        if (lm == LD_OLS_X && ols_threads() == 256) { launch_contig<R, C, false, LD_OLS_X, ST_PLAIN, 256>(ld, st, pa); return true; } \
        if (lm == LD_OLS_X) { launch_contig<R, C, false, LD_OLS_X, ST_PLAIN>(ld, st, pa); return true; }    \
    }
    ARS_FAST_CONTIG(F_CASE)
#undef F_CASE
    return false;
}

}  // namespace fftk
}  // namespace ars
