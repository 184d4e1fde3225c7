// Launchers of the FFT pass kernels, split over several translation units so that the ~130 kernel
// instantiations compile in parallel (fft_k_*.cu); fft_plan.cu only dispatches.
#pragma once
#include "fft.cuh"

namespace ars {
namespace fftk {

using namespace fft;

constexpr int NT = 512;

template <int LOGR, int LOGT, bool INV, int LDM, int STM>
static void launch_strided(const Ld& ld, const St& st, const PassArgs& pa) {
    using L = StridedLayout<LOGR, LOGT>;
    static unsigned long long attr_gen = 0;          // (function attributes belong to the context: set again after a re-init)
    const size_t smem = (Rad<LOGR>::n > 1) ? sizeof(float2) * L::SMEM_ELEMS : 0;
    auto k = pass_strided_kernel<LOGR, LOGT, INV, NT, LDM, STM>;
    if (attr_gen != ctx_generation() && smem > 48 * 1024) {
        ARS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_gen = ctx_generation();
    }
    const i64 tiles = (pa.total > 0 ? pa.total : pa.M) >> (LOGR + LOGT);
    k<<<(unsigned)tiles, NT, smem, ctx().stream>>>(ld, st, pa);
    ARS_LAUNCH_CHECK();
    count_launch();
}

// NTC = 256: three 8192-point tiles (and 768 threads) per SM instead of two -- the overlap-save block transforms wait
// on their first loads, and a third tile in flight hides more of that latency than the fourth warp per scheduler
template <int LOGR, int LOGC, bool INV, int LDM, int STM, int NTC = NT>
static void launch_contig(const Ld& ld, const St& st, const PassArgs& pa) {
    using L = ContigLayout<LOGR, LOGC>;
    static unsigned long long attr_gen = 0;
    const size_t smem = (Rad<LOGR>::n > 1) ? sizeof(float2) * L::SMEM_ELEMS : 0;
    auto k = pass_contig_kernel<LOGR, LOGC, INV, NTC, LDM, STM>;
    if (attr_gen != ctx_generation() && smem > 48 * 1024) {
        ARS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_gen = ctx_generation();
    }
    const i64 tiles = (pa.total > 0 ? pa.total : pa.M) >> (LOGR + LOGC);
    k<<<(unsigned)tiles, NTC, smem, ctx().stream>>>(ld, st, pa);
    ARS_LAUNCH_CHECK();
    count_launch();
}

// fused middle pass of the big-block overlap-save transforms (fft.cuh: pass_mid_kernel)
template <int LOGR, int LOGC, bool MIRROR, int NTM = NT>
static void launch_mid(const Ld& ld, const St& st, const PassArgs& pa, const MidArgs& ma) {
    using L = ContigLayout<LOGR, LOGC>;
    static unsigned long long attr_gen = 0;
    const size_t smem = sizeof(float2) * L::SMEM_ELEMS;
    auto k = pass_mid_kernel<LOGR, LOGC, NTM, MIRROR>;
    if (attr_gen != ctx_generation() && smem > 48 * 1024) {
        ARS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_gen = ctx_generation();
    }
    const i64 tiles = (pa.total > 0 ? pa.total : pa.M) >> (LOGR + LOGC);
    k<<<(unsigned)tiles, NTM, smem, ctx().stream>>>(ld, st, pa, ma);
    ARS_LAUNCH_CHECK();
    count_launch();
}
// persistent middle pass with bulk-copy prefetch (fft.cuh: pass_mid_pipe_kernel); PER_SM CTAs of NTM threads per SM
template <int LOGR, int NTM, int PER_SM>
static void launch_mid_pipe(const Ld& ld, const St& st, const PassArgs& pa, const MidArgs& ma) {
    using L = ContigLayout<LOGR, 0>;
    static unsigned long long attr_gen = 0;
    const size_t smem = sizeof(float2) * (L::SMEM_ELEMS + ((size_t)1 << LOGR));
    auto k = pass_mid_pipe_kernel<LOGR, NTM, PER_SM>;
    if (attr_gen != ctx_generation()) {
        ARS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_gen = ctx_generation();
    }
    const i64 tiles = (pa.total > 0 ? pa.total : pa.M) >> LOGR;
    const i64 grid = std::min<i64>(tiles, (i64)ctx().sm_count * PER_SM);
    k<<<(unsigned)grid, NTM, smem, ctx().stream>>>(ld, st, pa, ma, (int)tiles);
    ARS_LAUNCH_CHECK();
    count_launch();
}
// persistent last pass with bulk-copy prefetch (fft.cuh: pass_last_pipe_kernel); pa.ptab must be the table for 2^LOGT columns
template <int LOGR, int LOGT, int NTM, int PER_SM>
static void launch_last_pipe(const Ld& ld, const St& st, const PassArgs& pa) {
    using L = StridedFlat<LOGR, LOGT>;
    static unsigned long long attr_gen = 0;
    const size_t smem = sizeof(float2) * 2 * L::SMEM_ELEMS;
    auto k = pass_last_pipe_kernel<LOGR, LOGT, NTM, PER_SM>;
    if (attr_gen != ctx_generation()) {
        ARS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_gen = ctx_generation();
    }
    const i64 tiles = (pa.total > 0 ? pa.total : pa.M) >> (LOGR + LOGT);
    const i64 grid = std::min<i64>(tiles, (i64)ctx().sm_count * PER_SM);
    k<<<(unsigned)grid, NTM, smem, ctx().stream>>>(ld, st, pa, (int)tiles);
    ARS_LAUNCH_CHECK();
    count_launch();
}
bool last_pass_pipe(int logR, int* logT);                                       // fft_k_mid.cu: is there a pipelined last pass?
void last_pass_pipe_launch(int logR, const Ld& ld, const St& st, const PassArgs& pa);
void mid_pass(bool mirror, const Ld& ld, const St& st, const PassArgs& pa, const MidArgs& ma);    // fft_k_mid.cu

int ols_threads();     // 512 | 256: CTA size of the overlap-save block transforms (ARS_OLS_NT, fft_plan.cu)

// each returns false when it has no instantiation for the request (defined in fft_k_*.cu)
bool fast_strided_fwd(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);
bool fast_strided_inv(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);
bool fast_contig_fwd(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);
bool fast_contig_inv(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);
bool generic_strided_fwd(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);
bool generic_strided_inv(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);
bool generic_contig_fwd(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);
bool generic_contig_inv(const FftPass& ps, const Ld& ld, const St& st, const PassArgs& pa);

}  // namespace fftk
}  // namespace ars
