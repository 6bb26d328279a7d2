// chd_prep.cu — ChannelData pre-processing fused into one pass over the cube.
//
// The step before DAS in every reference script is a chain of ChannelData methods, each a full pass (or three,
// for the FFTs) over the ~1 GB cube on the host or through gpuArray temporaries:
//     chd = singleT(chd); chd = zeropad(chd, B, A); chd = hilbert(chd); chd = downmix(chd, fc);   (example_.m:261-269)
//   zeropad   src/ChannelData.m:1153-1183   B zeros before / A after each trace, t0 -= B/fs
//   hilbert   src/ChannelData.m:935-966     analytic signal along time: fft, weights [1 2..2 1 0..0], ifft
//   downmix   src/ChannelData.m:757-807     x .* exp(-2i*pi*fc*time), time = t0 + (0:T-1)/fs  (:1667)
//   singleT / halfT  :452-483               storage-type casts
// Here one kernel reads each input trace once (real fp32 / complex fp32 / real int16 / real fp64), does
// pad -> hilbert -> downmix entirely in shared memory and writes the complex fp32 or half2 trace DAS consumes:
// 1 read + 1 write of the cube, any subset of the steps.
//
// FFT: hand-written shared-memory transform, three radix-2 stages fused per pass in registers (radix-8 passes, twiddles
// from sincospif on exact dyadic fractions), in place and never permuting.  For power-of-two lengths the first pass reads
// the trace from global memory and the last one writes the finished (scaled, mixed, cast) samples back, so the trace
// crosses shared memory 6 times at T = 2048.  Lengths that are not a power of two go through Bluestein's chirp-z identity
//     X[k] = c[k] * sum_n (x[n] c[n]) conj(c[k-n]),  c[n] = exp(-i*pi*n^2/L)   (n^2 mod 2L in integers)
// with power-of-two FFTs of size >= 2L-1; the chirp spectrum is computed once per CTA (persistent CTAs loop over traces).
// Numerics: fp32 throughout; the downmix phase follows the reference's single-precision sequence
//     t = fl(t0' + fl(j/fs)),  theta = fl(fl(-2*pi*fc) * t)
// so it is bit-identical to the fp32 oracle up to the accuracy of sincosf; the FFT is tolerance-level (~1e-6).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "other_kernels.cuh"
#include "fft_smem.cuh"

namespace qups {

void count_launch(uint64_t n);
// sample j of the padded input trace k (real part only under hilbert: MATLAB's hilbert ignores the imaginary part)
__device__ __forceinline__ float2 prep_load(const PrepArgs &a, uint64_t k, uint64_t j) {
    float2 v = make_float2(0.f, 0.f);
    if (j >= a.B && j < a.B + a.T) {
        const uint64_t e = k * a.T + (j - a.B);
        if (a.in_dtype == PREP_REAL_F32) v.x = __ldg(reinterpret_cast<const float *>(a.in) + e);
        else if (a.in_dtype == PREP_REAL_I16) v.x = (float)__ldg(reinterpret_cast<const short *>(a.in) + e);
        else if (a.in_dtype == PREP_REAL_F64) v.x = (float)__ldg(reinterpret_cast<const double *>(a.in) + e);
        else { v = __ldg(reinterpret_cast<const float2 *>(a.in) + e); if (a.hilbert) v.y = 0.f; }
    }
    return v;
}
// scale, downmix (t0p = t0 - B/fs, src/ChannelData.m:1182), cast and store output sample j of trace k
__device__ __forceinline__ void prep_store(const PrepArgs &a, uint64_t k, uint64_t j, float2 v, float t0p, float sc) {
    v = make_float2(v.x * sc, v.y * sc);
    if (a.downmix) {
        const float t = add_rn(t0p, div_rn((float)j, a.fs));
        const float th = mul_rn(a.cmix, t);
        float sn, cs;
        sincosf(th, &sn, &cs);
        v = make_float2(sub_rn(mul_rn(v.x, cs), mul_rn(v.y, sn)), add_rn(mul_rn(v.x, sn), mul_rn(v.y, cs)));
    }
    if (a.out_half) reinterpret_cast<__half2 *>(a.out)[k * a.L + j] = __floats2half2_rn(v.x, v.y);
    else reinterpret_cast<float2 *>(a.out)[k * a.L + j] = v;
}

// element-wise passes (zeropad / cast / downmix without hilbert): one CTA per trace, four samples per thread in flight
__global__ void __launch_bounds__(256) chd_cast_kernel(const PrepArgs a) {
    for (uint64_t k = blockIdx.x; k < a.K; k += gridDim.x) {
        float t0 = 0.f;
        if (a.t0) t0 = __ldg(a.t0 + (k / a.traces_per_t0) % a.n_t0);
        const float t0p = sub_rn(t0, div_rn((float)a.B, a.fs));
        for (uint64_t j0 = threadIdx.x; j0 < a.L; j0 += 4 * 256) {
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (j0 + u * 256 < a.L) ? prep_load(a, k, j0 + u * 256) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (j0 + u * 256 < a.L) prep_store(a, k, j0 + u * 256, v[u], t0p, 1.0f);
        }
    }
}

__global__ void __launch_bounds__(1024) chd_prep_kernel(const PrepArgs a) {
    extern __shared__ __align__(16) unsigned char prep_smem[];
    float2 *s = reinterpret_cast<float2 *>(prep_smem);                 // padded work array (padded_len(nfft))
    float2 *tw = s + (a.hilbert ? padded_len(a.nfft) : 0);             // per-stage twiddle tables, nfft - 1 (+1 pad) entries
    float2 *cb = tw + a.nfft;                                          // Bluestein: FFT of the wrapped conjugate chirp (nfft)
    float2 *ch = cb + a.nfft;                                          // Bluestein: chirp c[n], n < L
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint64_t L = a.L;
    const uint32_t n = a.nfft, log2n = a.log2n, rb = bottom_bits(log2n);

    if (a.hilbert) fft_twiddles(tw, n);
    if (a.hilbert && a.bluestein) {
        for (uint32_t i = tid; i < L; i += nt) ch[i] = chirp(i, L);
        __syncthreads();
        for (uint32_t i = tid; i < n; i += nt) {
            float2 v = make_float2(0.f, 0.f);
            if (i < L) { v = ch[i]; v.y = -v.y; }
            else if (n - i < L) { v = ch[n - i]; v.y = -v.y; }
            s[padi(i)] = v;
        }
        __syncthreads();
        fft_fwd_dif(s, tw, n, log2n);
        for (uint32_t i = tid; i < n; i += nt) cb[i] = s[padi(i)];
        __syncthreads();
    }

    for (uint64_t k = blockIdx.x; k < a.K; k += gridDim.x) {
        float t0 = 0.f;
        if (a.t0) t0 = __ldg(a.t0 + (k / a.traces_per_t0) % a.n_t0);
        const float t0p = sub_rn(t0, div_rn((float)a.B, a.fs)); // zeropad: t0 - B/fs  (src/ChannelData.m:1182)
        const float sc = a.hilbert ? 1.0f / (float)L : 1.0f;
        auto load_in = [&](uint64_t j) -> float2 { return prep_load(a, k, j); };
        auto store_out = [&](uint64_t j, float2 v) { prep_store(a, k, j, v, t0p, sc); };
        // analytic-signal weights [1, 2 x (Nd2-1), 1+mod(L,2), 0 ...] for frequency kf (src/ChannelData.m:960-964)
        const uint64_t nd2 = L / 2;
        auto hweight = [&](uint64_t kf) -> float {
            if (L == 1 || kf == 0) return 1.f;
            if (kf < nd2) return 2.f;
            if (kf == nd2) return (L & 1) ? 2.f : 1.f;
            return 0.f;
        };
        if (a.bluestein) {
            for (uint64_t j = tid; j < L; j += nt) s[padi((uint32_t)j)] = load_in(j);
            __syncthreads();
            dft_bluestein(s, tw, cb, ch, L, n, log2n, false);
            for (uint64_t j = tid; j < L; j += nt) { const float w = hweight(j); const float2 v = s[padi((uint32_t)j)]; s[padi((uint32_t)j)] = make_float2(v.x * w, v.y * w); }
            __syncthreads();
            dft_bluestein(s, tw, cb, ch, L, n, log2n, true);
            for (uint64_t j = tid; j < L; j += nt) store_out(j, s[padi((uint32_t)j)]);
            __syncthreads();
        } else {
            // power-of-two length: global -> [top forward group] -> shared ... -> [bottom group: forward stages, weights,
            // inverse stages in registers] -> ... shared -> [top inverse group] -> global.  The trace crosses shared memory
            // 2 * (groups - 1) times (n = 2048: 6 passes; the radix-2 version: 22 + load + weights + store)
            auto ldg = [&](uint32_t i) { return load_in(i); };
            auto stg = [&](uint32_t i, float2 v) { store_out(i, v); };
            // analytic-signal weights in the bit-reversed layout: position i holds frequency bitrev(i), so the positive
            // frequencies (k < n/2) are the EVEN positions, k = 0 is position 0 and k = n/2 is position 1
            auto mid = [&](uint32_t i, float2 v) {
                const float w = (i < 2 || L == 1) ? 1.f : ((i & 1) ? 0.f : 2.f);
                return make_float2(v.x * w, v.y * w);
            };
            switch (log2n) {
            case 0: if (tid == 0) store_out(0, load_in(0)); break;
#define QUPS_FFT_CASE(LG_) case LG_: fft_roundtrip_c<LG_>(s, tw, ldg, stg, mid); break;
            QUPS_FFT_CASE(1) QUPS_FFT_CASE(2) QUPS_FFT_CASE(3) QUPS_FFT_CASE(4) QUPS_FFT_CASE(5) QUPS_FFT_CASE(6) QUPS_FFT_CASE(7)
            QUPS_FFT_CASE(8) QUPS_FFT_CASE(9) QUPS_FFT_CASE(10) QUPS_FFT_CASE(11) QUPS_FFT_CASE(12) QUPS_FFT_CASE(13)
#undef QUPS_FFT_CASE
            default: break; // longer traces do not fit shared memory: rejected by the launcher
            }
        }
    }
}

// returns cudaError_t as int; -1000 = unsupported length
int launch_chd_prep(PrepArgs a, cudaStream_t st) {
    a.L = a.B + a.T + a.A;
    if (a.K == 0 || a.L == 0) return 0;
    size_t smem = 0;
    a.nfft = 0; a.log2n = 0; a.bluestein = 0;
    if (a.hilbert) {
        uint32_t n = 1, lg = 0;
        const bool pow2 = (a.L & (a.L - 1)) == 0;
        const uint64_t need = pow2 ? a.L : 2 * a.L - 1;
        while (n < need) { n <<= 1; ++lg; if (lg > 20) return -1000; }
        a.nfft = n; a.log2n = lg; a.bluestein = !pow2;
        smem = sizeof(float2) * (padded_len(n) + n + (a.bluestein ? (size_t)n + a.L : 0));
        if (smem > 227 * 1024) return -1000;
    }
    cudaError_t e = cudaFuncSetAttribute(chd_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024));
    if (e != cudaSuccess) return (int)e;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // FFT traces: persistent CTAs (tables are built once per CTA), as many as shared memory allows, wide CTAs when one
    // trace fills the SM; plain element-wise passes: one CTA per trace, no barriers
    unsigned threads = 256;
    uint64_t grid = a.K;
    if (a.hilbert) {
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        threads = per_sm >= 4 ? 256 : (per_sm >= 2 ? 512 : 1024);
        const uint64_t cap = (uint64_t)sms * per_sm;
        grid = a.K < cap ? a.K : cap;
    } else if (grid > 0x7fffffffull) grid = 0x7fffffffull;
    if (!a.hilbert) chd_cast_kernel<<<(unsigned)grid, 256, 0, st>>>(a);
    else chd_prep_kernel<<<(unsigned)grid, threads, smem, st>>>(a);
    count_launch(1);
    return (int)cudaGetLastError();
}

} // namespace qups
