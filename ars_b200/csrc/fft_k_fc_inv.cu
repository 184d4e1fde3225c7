// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh).
#include "fft_launch.cuh"

namespace ars {
namespace fftk {

// inverse first pass: LD_MULSPEC -> ST_PLAIN ; overlap-save: fused MAC (or the tiled MAC kernel's output) -> ST_OLS | ST_OLS_CHIRP
bool fast_contig_inv(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa) {
 const int lm = ld.mode, sm = st.mode;
    if (ps.logR == 12 && ps.logT == 1 && sm == ST_OLS2) {      // radix-2-folded overlap-save inverse (fft_segments_r2)
        if (lm == LD_OLS_MAC) { launch_contig<12, 1, true, LD_OLS_MAC, ST_OLS2>(ld, st, pa); return true; }
        if (lm == LD_PLAIN) { launch_contig<12, 1, true, LD_PLAIN, ST_OLS2>(ld, st, pa); return true; }
    }
#define F_CASE(R, C)                                                                                                            \
    if (ps.logR == R && ps.logT == C) {                                                                                         \
        if (lm == LD_PLAIN && sm == ST_PLAIN) { launch_contig<R, C, true, LD_PLAIN, ST_PLAIN>(ld, st, pa); return true; }     \
        if (lm == LD_MULSPEC && sm == ST_PLAIN) { launch_contig<R, C, true, LD_MULSPEC, ST_PLAIN>(ld, st, pa); return true; } \
        if (lm == LD_PLAIN && sm == ST_OLS && ols_threads() == 256) { launch_contig<R, C, true, LD_PLAIN, ST_OLS, 256>(ld, st, pa); return true; } \
        if (lm == LD_OLS_MAC && sm == ST_OLS && ols_threads() == 256) { launch_contig<R, C, true, LD_OLS_MAC, ST_OLS, 256>(ld, st, pa); return true; } \
        if (lm == LD_PLAIN && sm == ST_OLS) { launch_contig<R, C, true, LD_PLAIN, ST_OLS>(ld, st, pa); return true; }           \
        if (lm == LD_OLS_MAC && sm == ST_OLS) { launch_contig<R, C, true, LD_OLS_MAC, ST_OLS>(ld, st, pa); return true; }     \
        if (lm == LD_OLS_MAC && sm == ST_OLS_CHIRP) { launch_contig<R, C, true, LD_OLS_MAC, ST_OLS_CHIRP>(ld, st, pa); return true; } \
    }
    ARS_FAST_CONTIG(F_CASE)
#undef F_CASE
    return false;
}

}  // namespace fftk
}  // namespace ars
