// das_args.cuh — kernel argument block shared by the DAS kernels and launchers.
#pragma once
#include <stdint.h>
#include "common.cuh"
#include "apod_fused.cuh"

namespace qups {

constexpr int MAX_APOD = 8;

// One launch = one frame. Mirrors the reference feval argument list
// (kern/das_spec.m:372) plus its __constant__ sizes (src/sizes.cu).
template <typename R> struct DasArgs {
    uint64_t I1, I2, I3, I, N, M, T;
    int S, interp;
    int keep_rx, keep_tx, tpose, VS, DV, apod_real, accumulate;
    R fs;
    const R *Pi;   // 3 x I
    const R *Pr;   // 3 x N
    const R *Pv4;  // 4 x M (row 4 = t0)
    const R *Nv;   // 3 x M
    const R *cinv; // broadcast, strides cstride[0..4]
    const void *apod;
    const void *x;
    void *y;
    uint64_t cstride[6];
    uint64_t astride[MAX_APOD][6];
    double pitch_hint[2], c_hint; // optional launcher hints (pixel pitch along I1 / I2, sound speed); 0 = unknown
    int fused;      // 0 = no closed-form apodization; otherwise `fa` is valid (fp32 paths only)
    FusedApod fa;
    // table-driven delays (staged kernel only; set by launch_wsinterpd2 for the canonical bfDAS form, else nullptr):
    // xq = 1 + (lut_tm[m * I + i] + lut_tn[n * I + i]), y = w * sum; Pi / Pr / Pv4 / Nv / cinv are then unused
    const float *lut_tn = nullptr, *lut_tm = nullptr, *lut_w = nullptr;
    int lut_wcplx = 0;
    // coherence mode (staged kernel only, qups_das_cohfac): y = sum over both apertures, cf[i] = |sum_n b_n|^2 / sum_n |b_n|^2 / N with
    // b_n the per-receive sums (kern/cohfac.m on DAS(..., 'keep_rx') output) — no I x N cube is materialised
    int cohfac = 0;
    float *cf = nullptr;
};

// launch entry points implemented in das_generic.cu / das_tiled.cu
// (return cudaError_t as int)
template <typename DIN, typename DA, typename DOUT, typename R>
int launch_das_generic(const DasArgs<R> &a, cudaStream_t st);
template <typename R> int launch_delays(const DasArgs<R> &a, R *tau, cudaStream_t st);
template <typename DIN, typename DOUT, typename R>
int launch_modulate(DOUT *xo, const DIN *x, const R *t0, int t0_stride, uint64_t T, uint64_t N, uint64_t M, int tpose,
                    R fs, double fmod, cudaStream_t st);

// element-wise precision conversion of a complex array (n complex elements)
int launch_half2_to_float2(float2 *dst, const __half2 *src, uint64_t n, cudaStream_t st);
int launch_float2_to_half2(__half2 *dst, const float2 *src, uint64_t n, cudaStream_t st);
int launch_half_to_float(float *dst, const __half *src, uint64_t n, cudaStream_t st);   // n real elements
int launch_float_to_half(__half *dst, const float *src, uint64_t n, cudaStream_t st);

// tiled fast path (fp32 data, sum over both apertures, scalar cinv, no apodization arrays)
struct TiledPlan {
    int eligible;      // 1 if the tiled kernel can take this call
    const char *why;   // reason when not eligible
};
TiledPlan das_tiled_plan(const DasArgs<float> &a, int dtype_in, int dtype_out);
int launch_das_tiled(const DasArgs<float> &a, cudaStream_t st);

// dense image of a closed-form apodization (apod_gen.cu): which = 0 -> receive weights I x N, 1 -> transmit weights I x M;
// out is real fp32, or interleaved complex fp32 (imag = 0) when as_complex
int launch_apod_generate(const FusedApod &fa, int which, float *out, int as_complex, const float *Pi, const float *Pr,
                         uint64_t I1, uint64_t I2, uint64_t I3, uint64_t NM, cudaStream_t st);

void count_launch(uint64_t n = 1);

// Stream-ordered scratch from the library's OWN memory pool (one per device, release threshold = keep): the default pool
// returns memory to the driver at every synchronisation point, so a per-call cudaMallocAsync would re-map it each time.
cudaError_t ws_alloc(void **p, size_t bytes, cudaStream_t st);
cudaError_t ws_free(void *p, cudaStream_t st);

} // namespace qups
