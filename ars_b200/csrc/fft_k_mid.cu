// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh): the fused middle pass of the big-block
// overlap-save transforms (contiguous forward pass, spectrum product, contiguous inverse pass in one kernel).
#include "fft_launch.cuh"

#include <cstdlib>

namespace ars {
namespace fftk {

void mid_pass(bool mirror, const Ld& ld, const St& st, const PassArgs& pa, const MidArgs& ma) {
    // plain form: one 4096-point segment per 256-thread CTA, four CTAs per SM (ARS_MID_NT=512: two segments per 512-thread
    // CTA, two per SM) -- the same threads and registers per SM, but four independent tiles whose load and arithmetic
    // phases interleave instead of two (measured on the 300 s render: 0.565 against 0.582 ms with one launch per pass,
    // 0.547 against 0.550 ms with four lanes)
    static const int nt = getenv("ARS_MID_NT") ? atoi(getenv("ARS_MID_NT")) : 256;
    // ARS_MID_PIPE (default 3): the persistent form whose next tile arrives by bulk asynchronous copy while the current one
    // is transformed, 3 | 2 CTAs per SM; 0: one tile per CTA with plain loads.  Measured on the 300 s render: the same
    // 84.4 us for the pass alone (it is bound by shared-memory traffic and issue slots, not by its loads), 0.539 against
    // 0.543 ms for the render.
    static const int pipe = getenv("ARS_MID_PIPE") ? atoi(getenv("ARS_MID_PIPE")) : 3;
    if (mirror) launch_mid<12, 1, true>(ld, st, pa, ma);
    else if (pipe == 3) launch_mid_pipe<12, 256, 3>(ld, st, pa, ma);
    else if (pipe == 2) launch_mid_pipe<12, 256, 2>(ld, st, pa, ma);
    else if (nt == 256) launch_mid<12, 0, false, 256>(ld, st, pa, ma);
    else launch_mid<12, 1, false>(ld, st, pa, ma);
}

// pipelined last pass (pass_last_pipe_kernel) for 64-row column transforms, i.e. 2^18-point blocks: tiles of 64 columns,
// two landing tiles per 256-thread CTA, three CTAs per SM.  ARS_LAST_PIPE: 0 off, 2 two CTAs per SM, 6 tiles of 32 columns
// x six 128-thread CTAs.  Measured on the 300 s render (last pass alone / whole render): off 86.3..90.5 us / 0.539 ms,
// default 70.0 us / 0.530 ms, two CTAs 78.2 us, 32-column tiles 69.8 us / 0.530 ms.
static int last_pipe_variant() {
    static const int v = getenv("ARS_LAST_PIPE") ? atoi(getenv("ARS_LAST_PIPE")) : 3;
    return v;
}
bool last_pass_pipe(int logR, int* logT) {
    const int v = last_pipe_variant();
    // (128- and 256-row column transforms with 32-column tiles, i.e. 256-byte rows: 82.3 against 80.3 us and 146 against
    // 78 us on the same render with 2^19- / 2^20-point blocks -- only the 64 x 64 tile pays)
    if (!v || logR != 6) return false;
    *logT = v == 6 ? 5 : 6;
    return true;
}
void last_pass_pipe_launch(int logR, const Ld& ld, const St& st, const PassArgs& pa) {
    ARS_CHECK(logR == 6, "no pipelined last pass for this column length");
    switch (last_pipe_variant()) {
        case 2: launch_last_pipe<6, 6, 256, 2>(ld, st, pa); break;
        case 6: launch_last_pipe<6, 5, 128, 6>(ld, st, pa); break;
        default: launch_last_pipe<6, 6, 256, 3>(ld, st, pa); break;
    }
}

}  // namespace fftk
}  // namespace ars
