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
// FFT: hand-written shared-memory radix-2 (bit-reversal + log2(L) butterfly passes, twiddles from sincospif on exact
// dyadic fractions).  Lengths that are not a power of two go through Bluestein's chirp-z identity
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

namespace qups {

void count_launch(uint64_t n);
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-place radix-2 FFT of s[0..n) in shared memory, all threads of the CTA.  Twiddles: one contiguous table PER STAGE,
// tw[half - 1 + k] = exp(-i*pi*k/half) for half = 1, 2, 4 .. n/2 (n - 1 entries, built once per CTA): a single table of
// exp(-2*pi*i*j/n) read at stride n/(2 half) put 16 lanes on one bank in the middle stages (ncu: 63 % of the shared
// wavefronts were conflicts).  inverse = conjugate twiddles (unscaled)
__device__ void fft_twiddles(float2 *tw, uint32_t n) {
    for (uint32_t j = threadIdx.x; j + 1 < n; j += blockDim.x) {
        const uint32_t half = 1u << (31 - __clz(j + 1)), k = j + 1 - half;
        float sn, cs;
        sincospif(-(float)k / (float)half, &sn, &cs); // exact dyadic argument
        tw[j] = make_float2(cs, sn);
    }
    __syncthreads();
}
// The transform pair never permutes: the FORWARD transform is decimation-in-frequency (natural order in, bit-reversed
// order out), the INVERSE is decimation-in-time on bit-reversed input (natural order out).  Everything in between
// (hilbert weights, Bluestein's spectrum product) is element-wise and simply indexes by the bit-reversed position.
// (A bit-reversal pass put consecutive lanes n/2 elements apart: 32-way bank conflicts, most of the shared wavefronts.)
__device__ void fft_fwd_dif(float2 *s, const float2 *tw, uint32_t n, uint32_t log2n) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    for (uint32_t st = log2n; st >= 1; --st) {
        const uint32_t half = 1u << (st - 1);
        const float2 *tws = tw + (half - 1);
        for (uint32_t b = tid; b < (n >> 1); b += nt) {
            const uint32_t k = b & (half - 1);
            const uint32_t i0 = ((b >> (st - 1)) << st) + k, i1 = i0 + half;
            const float2 a = s[i0], c = s[i1], w = tws[k];
            s[i0] = make_float2(a.x + c.x, a.y + c.y);
            s[i1] = cmulf(make_float2(a.x - c.x, a.y - c.y), w);
        }
        __syncthreads();
    }
}
__device__ void fft_inv_dit(float2 *s, const float2 *tw, uint32_t n, uint32_t log2n) { // unscaled
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    for (uint32_t st = 1; st <= log2n; ++st) {
        const uint32_t half = 1u << (st - 1);
        const float2 *tws = tw + (half - 1);
        for (uint32_t b = tid; b < (n >> 1); b += nt) {
            const uint32_t k = b & (half - 1);
            const uint32_t i0 = ((b >> (st - 1)) << st) + k, i1 = i0 + half;
            float2 w = tws[k];
            w.y = -w.y;
            const float2 a = s[i0], t = cmulf(s[i1], w);
            s[i0] = make_float2(a.x + t.x, a.y + t.y);
            s[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
}

// chirp c[n] = exp(-i*pi*n^2/L) with n^2 reduced mod 2L in integers
__device__ __forceinline__ float2 chirp(uint64_t n, uint64_t L) {
    const uint64_t r = (n * n) % (2 * L);
    float sn, cs;
    sincospif(-(float)((double)r / (double)L), &sn, &cs);
    return make_float2(cs, sn);
}

// DFT of length L (arbitrary) of s[0..L) via Bluestein; work arrays s (nfft) and cb (nfft, precomputed FFT of conj chirp)
__device__ void dft_bluestein(float2 *s, const float2 *tw, const float2 *cb, const float2 *ch, uint64_t L, uint32_t nfft, uint32_t log2n, bool inverse) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    // inverse DFT = conj(DFT(conj(x)))
    for (uint32_t i = tid; i < nfft; i += nt) {
        float2 v = make_float2(0.f, 0.f);
        if (i < L) {
            v = s[i];
            if (inverse) v.y = -v.y;
            v = cmulf(v, ch[i]);
        }
        s[i] = v;
    }
    __syncthreads();
    fft_fwd_dif(s, tw, nfft, log2n);
    for (uint32_t i = tid; i < nfft; i += nt) s[i] = cmulf(s[i], cb[i]); // both spectra in bit-reversed order
    __syncthreads();
    fft_inv_dit(s, tw, nfft, log2n);
    const float sc = 1.0f / (float)nfft;
    for (uint32_t i = tid; i < L; i += nt) {
        float2 v = cmulf(make_float2(s[i].x * sc, s[i].y * sc), ch[i]);
        if (inverse) v.y = -v.y;
        s[i] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) chd_prep_kernel(const PrepArgs a) {
    extern __shared__ __align__(16) unsigned char prep_smem[];
    float2 *s = reinterpret_cast<float2 *>(prep_smem);
    float2 *tw = s + (a.hilbert ? a.nfft : 0);        // per-stage twiddle tables, nfft - 1 (+1 pad) entries
    float2 *cb = tw + a.nfft;                         // Bluestein: FFT of the wrapped conjugate chirp (nfft)
    float2 *ch = cb + a.nfft;                         // Bluestein: chirp c[n], n < L
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint64_t L = a.L;

    if (a.hilbert) fft_twiddles(tw, a.nfft);
    if (a.hilbert && a.bluestein) {
        for (uint32_t i = tid; i < L; i += nt) ch[i] = chirp(i, L);
        __syncthreads();
        for (uint32_t i = tid; i < a.nfft; i += nt) {
            float2 v = make_float2(0.f, 0.f);
            if (i < L) { v = ch[i]; v.y = -v.y; }
            else if (a.nfft - i < L) { v = ch[a.nfft - i]; v.y = -v.y; }
            cb[i] = v;
        }
        __syncthreads();
        fft_fwd_dif(cb, tw, a.nfft, a.log2n);
    }

    for (uint64_t k = blockIdx.x; k < a.K; k += gridDim.x) {
        // ---- load + zero-pad -------------------------------------------------------------------------
        if (a.hilbert) {
            for (uint64_t j = tid; j < (a.bluestein ? L : (uint64_t)a.nfft); j += nt) {
                float v = 0.f;
                if (j >= a.B && j < a.B + a.T) {
                    const uint64_t e = k * a.T + (j - a.B);
                    if (a.in_dtype == PREP_REAL_F32) v = __ldg(reinterpret_cast<const float *>(a.in) + e);
                    else if (a.in_dtype == PREP_REAL_I16) v = (float)__ldg(reinterpret_cast<const short *>(a.in) + e);
                    else if (a.in_dtype == PREP_REAL_F64) v = (float)__ldg(reinterpret_cast<const double *>(a.in) + e);
                    else v = __ldg(reinterpret_cast<const float2 *>(a.in) + e).x; // MATLAB hilbert ignores the imaginary part
                }
                s[j] = make_float2(v, 0.f);
            }
            __syncthreads();
            // ---- analytic signal: fft, weights [1, 2 x (Nd2-1), 1+mod(L,2), 0 ...], ifft (src/ChannelData.m:960-964) ----
            if (a.bluestein) dft_bluestein(s, tw, cb, ch, L, a.nfft, a.log2n, false);
            else fft_fwd_dif(s, tw, a.nfft, a.log2n);
            const uint64_t nd2 = L / 2;
            for (uint64_t j = tid; j < L; j += nt) {
                // frequency index held at position j: natural order after Bluestein, bit-reversed after the in-place DIF
                const uint64_t kf = (a.bluestein || a.log2n == 0) ? j : (uint64_t)(__brev((uint32_t)j) >> (32 - a.log2n));
                float w;
                if (kf == 0) w = 1.f;
                else if (kf < nd2) w = 2.f;
                else if (kf == nd2) w = (L & 1) ? 2.f : 1.f;
                else w = 0.f;
                if (L == 1) w = 1.f;
                s[j] = make_float2(s[j].x * w, s[j].y * w);
            }
            __syncthreads();
            if (a.bluestein) dft_bluestein(s, tw, cb, ch, L, a.nfft, a.log2n, true);
            else fft_inv_dit(s, tw, a.nfft, a.log2n);
        }
        // ---- downmix + cast + store ----------------------------------------------------------------------
        float t0 = 0.f;
        if (a.t0) t0 = __ldg(a.t0 + (k / a.traces_per_t0) % a.n_t0);
        const float t0p = sub_rn(t0, div_rn((float)a.B, a.fs)); // zeropad: t0 - B/fs  (src/ChannelData.m:1182)
        const float sc = a.hilbert ? 1.0f / (float)L : 1.0f;
        for (uint64_t j = tid; j < L; j += nt) {
            float2 v;
            if (a.hilbert) {
                v = make_float2(s[j].x * sc, s[j].y * sc);
            } else {
                v = make_float2(0.f, 0.f);
                if (j >= a.B && j < a.B + a.T) {
                    const uint64_t e = k * a.T + (j - a.B);
                    if (a.in_dtype == PREP_REAL_F32) v.x = __ldg(reinterpret_cast<const float *>(a.in) + e);
                    else if (a.in_dtype == PREP_REAL_I16) v.x = (float)__ldg(reinterpret_cast<const short *>(a.in) + e);
                    else if (a.in_dtype == PREP_REAL_F64) v.x = (float)__ldg(reinterpret_cast<const double *>(a.in) + e);
                    else v = __ldg(reinterpret_cast<const float2 *>(a.in) + e);
                }
            }
            if (a.downmix) {
                const float t = add_rn(t0p, div_rn((float)j, a.fs));
                const float th = mul_rn(a.cmix, t);
                float sn, cs;
                sincosf(th, &sn, &cs);
                v = make_float2(sub_rn(mul_rn(v.x, cs), mul_rn(v.y, sn)), add_rn(mul_rn(v.x, sn), mul_rn(v.y, cs)));
            }
            if (a.out_half) reinterpret_cast<__half2 *>(a.out)[k * L + j] = __floats2half2_rn(v.x, v.y);
            else reinterpret_cast<float2 *>(a.out)[k * L + j] = v;
        }
        if (a.hilbert) __syncthreads();
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
        smem = sizeof(float2) * ((size_t)n + n + (a.bluestein ? (size_t)n + a.L : 0));
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
    chd_prep_kernel<<<(unsigned)grid, threads, smem, st>>>(a);
    count_launch(1);
    return (int)cudaGetLastError();
}

} // namespace qups
