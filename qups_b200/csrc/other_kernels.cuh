// other_kernels.cuh — launchers for the non-DAS entry points of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "../../include/qups_b200.h"

namespace qups {
// return 0, a cudaError_t (> 0), -3 (unsupported) or -4 (allocation)
int launch_wsinterpd2(const qups_ws2_params &p, void *y, const void *w, const void *x, const void *t1, const void *t2,
                      cudaStream_t st);
const char *last_ws2_kernel_name();   // "ws2_tiled" | "wsinterpd2" | "none": what the last launch_wsinterpd2 dispatched to
int launch_greens(const qups_greens_params &p, void *y, const void *Pi, const void *a, const void *Pr, const void *Pv,
                  const void *kern, cudaStream_t st);
int launch_convd(const qups_convd_params &p, void *z, const void *x, const void *y, cudaStream_t st);

// ChannelData pre-processing (chd_prep.cu)
enum { PREP_REAL_F32 = 0, PREP_CPLX_F32 = 1, PREP_REAL_I16 = 2, PREP_REAL_F64 = 3 };

struct PrepArgs {
    const void *in;
    void *out;
    const float *t0;   // n_t0 start times (device) or nullptr (t0 = 0)
    uint64_t T, K, B, A, L;   // L = B + T + A
    uint64_t traces_per_t0, n_t0;
    int in_dtype, out_half, hilbert, downmix;
    float fs, cmix;    // cmix = fl32(-2*pi*fmix)
    uint32_t nfft, log2n; // power-of-two FFT size (L itself, or >= 2L-1 for Bluestein)
    int bluestein;
};

// returns 0, a cudaError_t (> 0) or -1000 (hilbert length not supported)
int launch_chd_prep(PrepArgs a, cudaStream_t st);

// aperture-domain post-processing (aperture.cu); lags is a HOST array of p.nlags entries
int launch_aperture(const qups_aperture_params &p, void *out, void *out2, const void *b, const uint32_t *lags, cudaStream_t st);

// pair-wise windowed cross-correlation (xcorr.cu); lags_dev is a DEVICE int32[L]
size_t xcorr_smem_bytes(uint32_t W, int dbl);
int launch_pwznxcorr(int dbl, void *y, const void *x, const void *x0, const void *w, const int32_t *lags_dev, uint32_t T, uint32_t P,
                     uint32_t N, uint32_t F, uint32_t L, uint32_t W, int ref, int zero, int norm, int x_complex, uint32_t S,
                     uint32_t x0N, uint32_t x0F, cudaStream_t st);

// REFoCUS decode (refocus.cu); dt_host: V doubles t0(v) - min(t0).  returns 0, a cudaError_t (> 0) or -3 (T not a power of two <= 8192)
int launch_refocus(void *y, const void *x, const void *Hi, const double *dt_host, uint64_t T, uint64_t N, uint64_t V, uint64_t E,
                   double fs, cudaStream_t st);
} // namespace qups
