// Procedural impulse-response synthesis (rs.py:249-305) on the device.
//
// The host replays the reference's random draws (tap delays, base strengths, raw tail
// noise) and the scalar shaping (decay, amplitude, smoothing width) and hands them in;
// the kernels do the array work in float64 exactly where numpy does, rounding to
// float32 where numpy stores float32:
//   early: sequential scatter-accumulate of <= a few dozen taps, peak 0.9 over early[1:]
//   late : boxcar-`width` smoothing ('same' window), std re-scaling, amp * decay^i
//          envelope, peak 0.7
#include "ir_synth.cuh"

#include <algorithm>
#include <cmath>

namespace ars {

struct IrStats {
    double sum_noise, sum_box;      // for the means
    double dev_noise, dev_box;      // sums of squared deviations
    unsigned max_late;              // bits of max |late| (float32, before normalisation)
    unsigned pad;
};

__device__ __forceinline__ double box_at(const double* __restrict__ noise, i64 n, int width, i64 i) {
    // np.convolve(noise, ones(width)/width, 'same')[i] = sum_j noise[i + (width-1)/2 - j] / width
    const i64 top = i + (width - 1) / 2;
    const double w = 1.0 / (double)width;
    double acc = 0.0;
    for (int j = width - 1; j >= 0; --j) {       // ascending sample index
        const i64 idx = top - j;
        if (idx >= 0 && idx < n) acc += noise[idx] * w;
    }
    return acc;
}

__device__ __forceinline__ void block_add2(double a, double b, double* da, double* db) {
    __shared__ double sa[32], sb[32];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) / 32;
        a = threadIdx.x < nw ? sa[threadIdx.x] : 0.0;
        b = threadIdx.x < nw ? sb[threadIdx.x] : 0.0;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (threadIdx.x == 0) { atomicAdd(da, a); atomicAdd(db, b); }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) ir_box_kernel(const double* __restrict__ noise, i64 n, int width,
                                                     double* __restrict__ box, IrStats* st) {
    double a = 0.0, b = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double v = box_at(noise, n, width, i);
        box[i] = v;
        a += noise[i];
        b += v;
    }
    block_add2(a, b, &st->sum_noise, &st->sum_box);
}

__global__ void __launch_bounds__(256) ir_dev_kernel(const double* __restrict__ noise, const double* __restrict__ box, i64 n,
                                                     IrStats* st) {
    const double mn = st->sum_noise / (double)n, mb = st->sum_box / (double)n;
    double a = 0.0, b = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double x = noise[i] - mn, y = box[i] - mb;
        a += x * x;
        b += y * y;
    }
    block_add2(a, b, &st->dev_noise, &st->dev_box);
}

__global__ void __launch_bounds__(256) ir_tail_kernel(const double* __restrict__ noise, const double* __restrict__ box, i64 n,
                                                      int smoothed, double amp, double decay, float* __restrict__ late_tail,
                                                      IrStats* st) {
    double s_raw = 0.0, s_box = 0.0;
    bool use_box = false;
    if (smoothed) {
        s_raw = sqrt(st->dev_noise / (double)n);
        s_box = sqrt(st->dev_box / (double)n);
        use_box = s_box > 1e-6;                       // rs.py:291-292
    }
    unsigned m = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double v = use_box ? __dmul_rn(__ddiv_rn(box[i], s_box), s_raw) : noise[i];
        v = __dmul_rn(__dmul_rn(v, amp), pow(decay, (double)i));   // rs.py:295-296
        const float f = __double2float_rn(v);
        late_tail[i] = f;
        m = max(m, abs_bits(f));
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&st->max_late, m);
}

__global__ void __launch_bounds__(256) ir_late_norm_kernel(float* __restrict__ late_tail, i64 n, const IrStats* st) {
    const float pk = __uint_as_float(st->max_late);
    if (!(pk > 1e-6f)) return;                        // rs.py:302-303
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        late_tail[i] = __fmul_rn(__fdiv_rn(late_tail[i], pk), 0.7f);
}

// The early part has at most a few dozen taps: the host merges colliding taps in the reference's draw order and
// normalises them (ir_early_taps below, same float64 -> float32 roundings as numpy); the device only scatters.
__global__ void ir_scatter_kernel(float* __restrict__ early, i64 length, const i64* __restrict__ pos,
                                  const double* __restrict__ val, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && pos[j] >= 0 && pos[j] < length) early[pos[j]] = (float)val[j];
}

int ir_early_taps(const i64* delay, const double* strength, int ntaps, i64 length, std::vector<i64>& pos,
                  std::vector<double>& val) {
    // rs.py:268: early[d] += strength  (float32 element + float64 scalar, rounded back to float32), in draw order
    pos.clear();
    val.clear();
    std::vector<float> acc;
    for (int j = 0; j < ntaps; ++j) {
        const i64 d = delay[j];
        if (d < 0 || d >= length) continue;
        size_t k = 0;
        while (k < pos.size() && pos[k] != d) ++k;
        if (k == pos.size()) { pos.push_back(d); acc.push_back(0.f); }
        acc[k] = (float)((double)acc[k] + strength[j]);
    }
    // rs.py:299-300: early[1:] = (early[1:] / max) * 0.9 when max > 1e-6 (float32 arithmetic)
    if (length > 1) {
        float pk = 0.f;
        for (size_t k = 0; k < pos.size(); ++k)
            if (pos[k] >= 1) pk = std::max(pk, std::fabs(acc[k]));
        if (pk > 1e-6f)
            for (size_t k = 0; k < pos.size(); ++k)
                if (pos[k] >= 1) { const float q = acc[k] / pk; acc[k] = q * 0.9f; }
    }
    val.assign(acc.begin(), acc.end());
    return (int)pos.size();
}

void ir_synth(const IrSpec& sp, const i64* d_delay, const double* d_strength, const double* d_noise, float* d_early,
              float* d_late) {
    Ctx& c = ctx();
    ARS_CHECK(sp.length >= 1 && sp.split >= 0 && sp.split <= sp.length, "ir_synth: bad geometry");
    KernelScope prof("ir synthesis chain (taps, smoothed tail, envelope, normalisation)", 16.0 * (double)sp.length);
    ARS_CUDA(cudaMemsetAsync(d_early, 0, sizeof(float) * (size_t)sp.length, c.stream));
    ARS_CUDA(cudaMemsetAsync(d_late, 0, sizeof(float) * (size_t)sp.length, c.stream));
    if (sp.ntaps > 0) {       // d_delay / d_strength hold the merged, normalised taps (ir_early_taps)
        ir_scatter_kernel<<<ceil_div(sp.ntaps, 128), 128, 0, c.stream>>>(d_early, sp.length, d_delay, d_strength, sp.ntaps);
        ARS_LAUNCH_CHECK();
        count_launch();
    }
    const i64 n = sp.length - sp.split;
    if (n <= 0) return;
    IrStats* st = c.buf("ir.stats", sizeof(IrStats)).as<IrStats>();
    ARS_CUDA(cudaMemsetAsync(st, 0, sizeof(IrStats), c.stream));
    const int grid = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)c.sm_count * 4));
    const int smoothed = (sp.width > 1 && n >= sp.width) ? 1 : 0;     // rs.py:286
    double* box = nullptr;
    if (smoothed) {
        box = c.buf("ir.box", sizeof(double) * (size_t)n).as<double>();
        ir_box_kernel<<<grid, 256, 0, c.stream>>>(d_noise, n, sp.width, box, st);
        ARS_LAUNCH_CHECK();
        ir_dev_kernel<<<grid, 256, 0, c.stream>>>(d_noise, box, n, st);
        ARS_LAUNCH_CHECK();
        count_launch(2);
    }
    float* tail = d_late + sp.split;
    ir_tail_kernel<<<grid, 256, 0, c.stream>>>(d_noise, box ? box : d_noise, n, smoothed, sp.amp, sp.decay, tail, st);
    ARS_LAUNCH_CHECK();
    ir_late_norm_kernel<<<grid, 256, 0, c.stream>>>(tail, n, st);
    ARS_LAUNCH_CHECK();
    count_launch(2);
}

}  // namespace ars
