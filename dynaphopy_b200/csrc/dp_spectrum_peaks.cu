// Hand-off of the finished spectra to the fitting stage: degenerate-set average and initial guess.
// Reference: dynaphopy/analysis/fitting/__init__.py:17-39 (degenerate_sets / average_phonon: the spectrum of a mode
// is replaced by np.average over the modes of its degenerate set) and :60-61 (guess_height = np.max,
// guess_position = frequency[np.argmax]) -- done there per mode in Python for one q at a time
// (phonon_fitting_analysis, called from the q loop of Quasiparticle.get_commensurate_points_data,
// dynaphopy/__init__.py:1130-1174).  Here: all series of the mesh in one launch; Lorentzian fitting itself stays in
// the reference's scipy code and receives (averaged spectra, peak bins, peak heights).
#include "dp_common.cuh"

namespace dp {

// np.max / np.argmax order: NaN beats everything, the first of equal values wins
__device__ __forceinline__ bool peak_better(double va, int ia, double vb, int ib) {
    const bool na = va != va, nb = vb != vb;
    if (na || nb) return na && (!nb || ia < ib);
    return va > vb || (va == vb && ia < ib);
}

// one warp per series; lanes stride over the frequencies
__global__ void __launch_bounds__(256)
spectra_peaks_kernel(const double* __restrict__ spectra, int F, int S, const int* __restrict__ set_ptr,
                     const int* __restrict__ set_members, const int* __restrict__ series_set,
                     double* __restrict__ averaged, int* __restrict__ peak_bin, double* __restrict__ peak_height) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < S; s += warps) {
        int m0 = s, m1 = s + 1;
        if (series_set) { const int g = series_set[s]; m0 = set_ptr[g]; m1 = set_ptr[g + 1]; }
        const int n = m1 - m0;
        double best = 0.0;
        int best_i = 0x7fffffff;
        for (int f = lane; f < F; f += 32) {
            const double* row = spectra + (size_t)f * S;
            double v;
            if (!series_set) v = row[s];
            else {
                // np.average over the stacked member spectra: sequential sum along axis 0, then one division
                v = row[set_members[m0]];
                for (int j = m0 + 1; j < m1; j++) v = __dadd_rn(v, row[set_members[j]]);
                if (n > 1) v = __ddiv_rn(v, (double)n);
            }
            if (averaged) averaged[(size_t)f * S + s] = v;
            if (best_i == 0x7fffffff || peak_better(v, f, best, best_i)) { best = v; best_i = f; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (oi != 0x7fffffff && (best_i == 0x7fffffff || peak_better(ov, oi, best, best_i))) { best = ov; best_i = oi; }
        }
        if (lane == 0) { peak_bin[s] = best_i == 0x7fffffff ? 0 : best_i; peak_height[s] = best; }
    }
}

}  // namespace dp

extern "C" {

int dp_spectra_peaks(const double* spectra, int F, int S, const int* set_ptr, const int* set_members,
                     const int* series_set, int n_sets, double* averaged, int* peak_bin, double* peak_height,
                     void* stream) {
    DP_REQUIRE(F > 0 && S > 0, DP_ERR_INVALID, "dp_spectra_peaks: bad sizes F=%d S=%d", F, S);
    DP_REQUIRE(spectra && peak_bin && peak_height, DP_ERR_INVALID, "dp_spectra_peaks: null pointer");
    DP_REQUIRE((series_set == nullptr) == (set_ptr == nullptr) && (series_set == nullptr) == (set_members == nullptr),
               DP_ERR_INVALID, "dp_spectra_peaks: set_ptr, set_members and series_set go together");
    DP_REQUIRE(series_set == nullptr || n_sets > 0, DP_ERR_INVALID, "dp_spectra_peaks: n_sets = %d", n_sets);
    DP_REQUIRE(averaged != spectra, DP_ERR_INVALID, "dp_spectra_peaks: the average cannot be taken in place");
    auto k = dp::spectra_peaks_kernel;
    const int64_t blocks = ((int64_t)S * 32 + 255) / 256;
    const int grid = (int)dp::imin<int64_t>(blocks, (int64_t)dp::imax(dp::sm_count(), 1) * 16);
    DP_LAUNCH(k, dim3(grid), dim3(256), 0, (cudaStream_t)stream, spectra, F, S, set_ptr, set_members, series_set,
              averaged, peak_bin, peak_height);
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

}  // extern "C"
