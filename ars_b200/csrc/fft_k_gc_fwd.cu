// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh).
#include "fft_launch.cuh"

namespace ars {
namespace fftk {

// runtime-switch (generic) contiguous kernels, every tile shape
bool generic_contig_fwd(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa) {
#define C_CASE(R, C) if (ps.logR == R && ps.logT == C) { launch_contig<R, C, false, -1, -1>(ld, st, pa); return true; }
    ARS_CONTIG_CASES(C_CASE)
#undef C_CASE
    return false;
}

}  // namespace fftk
}  // namespace ars
