// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh).
#include "fft_launch.cuh"

namespace ars {
namespace fftk {

// runtime-switch (generic) strided kernels, every tile shape
bool generic_strided_inv(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa) {
#define S_CASE(R, T) if (ps.logR == R && ps.logT == T) { launch_strided<R, T, true, -1, -1>(ld, st, pa); return true; }
    ARS_STRIDED_CASES(S_CASE)
#undef S_CASE
    return false;
}

}  // namespace fftk
}  // namespace ars
