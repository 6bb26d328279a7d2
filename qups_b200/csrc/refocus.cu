// refocus.cu — REFoCUS decoding of a transmit sequence back to full-synthetic-aperture data (SURVEY.md §8f-3;
// src/UltrasoundSystem.m:3729-3757: fft along time, time-alignment phase, per-frequency decode, phase, ifft).
//
// The reference applies the decoder with a MATLAB loop over transmit elements, each iteration a broadcast multiply of the
// whole T x N x V spectrum by one page of Hi and a sum over V (`y{v} = sum(sub(Hi,v,D+1) .* x, mdim)`, :3746-3750): E passes
// over the cube.  Here it is three kernels and two scratch cubes:
//   1. refocus_fwd_kernel   one CTA per group of 4 receive traces of one pulse: radix-8 shared-memory FFT (fft_smem.cuh,
//      first pass straight from global memory), time-alignment phase, spectrum written FREQUENCY-MAJOR
//      Xt[n + N (v + V i)] — 4 consecutive n per frequency = full 32-byte sectors — and left in the bit-reversed order
//      the in-place transform produces (position i holds frequency bitrev(i); nothing is ever permuted);
//   2. refocus_gemm_kernel  per frequency position a complex N x V by V x E product with that frequency's decoder page,
//      64 x 64 output tiles, 16-deep shared-memory stages, 4 x 4 outputs per thread (fp32 SIMT: the decoder's entries
//      span ~5 decades, tf32/bf16 tensor-core inputs would cost the 1e-5 parity);
//   3. refocus_inv_kernel   reads Yt[n + N (e + E i)] the same way, inverse transform (bit-reversed in, natural out),
//      1/T scale, coalesced store of the decoded traces.
// The decoder Hi (E x V x T, natural frequency order) is an input: it depends only on the sequence, the reference builds
// it with pagenorm / pagemldivide on gathered host arrays (:3702-3727), and so does the host-side mirror (in float64).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "other_kernels.cuh"
#include "fft_smem.cuh"

namespace qups {
void count_launch(uint64_t n);
cudaError_t ws_alloc(void **p, size_t bytes, cudaStream_t st);
cudaError_t ws_free(void *p, cudaStream_t st);

namespace {
constexpr int kG = 4; // traces per CTA of the transform kernels

struct RefocusArgs {
    const float2 *x;      // T x N x V
    const float2 *Hi;     // E x V x T
    float2 *y;            // T x N x E
    float2 *Xt, *Yt;      // scratch, frequency-major
    const float2 *phase;  // T x V time-alignment phasors (natural frequency order) or nullptr
    uint32_t T, log2T, N, V, E;
};

// exp(-2i*pi*f_k*dt_v), f_k = k fs / T (src/ChannelData.m:1491), dt_v = t0(v) - min(t0): the two phase factors of
// :3733 and :3756 commute with the (linear) decode and are applied as one.  float64 phase (f dt reaches hundreds of cycles)
__global__ void refocus_phase_kernel(float2 *ph, const double *dt, uint32_t T, uint32_t V, double fs) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)T * V) return;
    const uint32_t k = (uint32_t)(idx % T), v = (uint32_t)(idx / T);
    const double cyc = ((double)k * fs / (double)T) * dt[v];
    double sn, cs;
    sincospi(-2.0 * (cyc - floor(cyc)), &sn, &cs);
    ph[idx] = make_float2((float)cs, (float)sn);
}

__global__ void __launch_bounds__(256) refocus_fwd_kernel(const RefocusArgs a) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    float2 *tw = reinterpret_cast<float2 *>(rf_smem);
    float2 *sg = tw + a.T; // kG padded arrays
    const uint32_t T = a.T, lg = a.log2T, rb = bottom_bits(lg), plen = (uint32_t)padded_len(T);
    fft_twiddles(tw, T);
    const uint32_t groups_n = (a.N + kG - 1) / kG;
    for (uint64_t grp = blockIdx.x; grp < (uint64_t)groups_n * a.V; grp += gridDim.x) {
        const uint32_t n0 = (uint32_t)(grp % groups_n) * kG, v = (uint32_t)(grp / groups_n);
        for (int g = 0; g < kG; ++g) {
            float2 *s = sg + (size_t)g * plen;
            const uint32_t n = n0 + g;
            const float2 *src = a.x + ((uint64_t)v * a.N + (n < a.N ? n : a.N - 1)) * T;
            auto ldg = [&](uint32_t i) { return __ldg(src + i); };
            auto lds = [&](uint32_t i) { return s[padi(i)]; };
            auto sts = [&](uint32_t i, float2 val) { s[padi(i)] = val; };
            if (lg == 0) { if (threadIdx.x == 0) s[0] = __ldg(src); }
            else if (lg == rb) {
                if (rb == 3) fft_group<3, false>(T, 3, tw, ldg, sts); else if (rb == 2) fft_group<2, false>(T, 2, tw, ldg, sts); else fft_group<1, false>(T, 1, tw, ldg, sts);
            } else {
                fft_group<3, false>(T, lg, tw, ldg, sts);
                __syncthreads();
                for (uint32_t st = lg - 3; st > rb; st -= 3) { fft_group<3, false>(T, st, tw, lds, sts); __syncthreads(); }
                if (rb == 3) fft_group<3, false>(T, 3, tw, lds, sts); else if (rb == 2) fft_group<2, false>(T, 2, tw, lds, sts); else fft_group<1, false>(T, 1, tw, lds, sts);
            }
        }
        __syncthreads();
        // frequency-major store: position i (frequency bitrev(i)), kG consecutive receives
        for (uint32_t idx = threadIdx.x; idx < T * kG; idx += blockDim.x) {
            const uint32_t g = idx % kG, i = idx / kG, n = n0 + g;
            if (n >= a.N) continue;
            float2 val = sg[(size_t)g * plen + padi(i)];
            if (a.phase) {
                const uint32_t kf = lg ? (__brev(i) >> (32 - lg)) : 0u;
                val = cmulf(val, __ldg(a.phase + (uint64_t)v * T + kf));
            }
            a.Xt[n + (uint64_t)a.N * (v + (uint64_t)a.V * i)] = val;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) refocus_inv_kernel(const RefocusArgs a) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    float2 *tw = reinterpret_cast<float2 *>(rf_smem);
    float2 *sg = tw + a.T;
    const uint32_t T = a.T, lg = a.log2T, rb = bottom_bits(lg), plen = (uint32_t)padded_len(T);
    const float sc = 1.0f / (float)T;
    fft_twiddles(tw, T);
    const uint32_t groups_n = (a.N + kG - 1) / kG;
    for (uint64_t grp = blockIdx.x; grp < (uint64_t)groups_n * a.E; grp += gridDim.x) {
        const uint32_t n0 = (uint32_t)(grp % groups_n) * kG, e = (uint32_t)(grp / groups_n);
        for (uint32_t idx = threadIdx.x; idx < T * kG; idx += blockDim.x) {
            const uint32_t g = idx % kG, i = idx / kG, n = n0 + g;
            sg[(size_t)g * plen + padi(i)] = n < a.N ? __ldg(a.Yt + n + (uint64_t)a.N * (e + (uint64_t)a.E * i)) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        for (int g = 0; g < kG; ++g) {
            float2 *s = sg + (size_t)g * plen;
            const uint32_t n = n0 + g;
            if (n >= a.N) break; // CTA-uniform
            float2 *dst = a.y + ((uint64_t)e * a.N + n) * T;
            auto lds = [&](uint32_t i) { return s[padi(i)]; };
            auto sts = [&](uint32_t i, float2 val) { s[padi(i)] = val; };
            auto stg = [&](uint32_t i, float2 val) { dst[i] = make_float2(val.x * sc, val.y * sc); };
            if (lg == 0) { if (threadIdx.x == 0) dst[0] = s[0]; }
            else if (lg == rb) {
                if (rb == 3) fft_group<3, true>(T, 3, tw, lds, stg); else if (rb == 2) fft_group<2, true>(T, 2, tw, lds, stg); else fft_group<1, true>(T, 1, tw, lds, stg);
            } else {
                if (rb == 3) fft_group<3, true>(T, 3, tw, lds, sts); else if (rb == 2) fft_group<2, true>(T, 2, tw, lds, sts); else fft_group<1, true>(T, 1, tw, lds, sts);
                __syncthreads();
                for (uint32_t st = rb + 3; st < lg; st += 3) { fft_group<3, true>(T, st, tw, lds, sts); __syncthreads(); }
                fft_group<3, true>(T, lg, tw, lds, stg);
            }
        }
        __syncthreads();
    }
}

// C[n, e] = sum_v A[n, v] * B[e, v] per frequency position (grid.z): A = Xt page (N x V), B = Hi page of frequency
// bitrev(i) (E x V), C = Yt page (N x E), all column-major complex fp32
constexpr int kTM = 64, kTN = 64, kTK = 16;
__global__ void __launch_bounds__(256) refocus_gemm_kernel(const RefocusArgs a) {
    __shared__ float2 As[kTK][kTM + 1], Bs[kTK][kTN + 1];
    const uint32_t i = blockIdx.z, kf = a.log2T ? (__brev(i) >> (32 - a.log2T)) : 0u;
    const float2 *A = a.Xt + (uint64_t)i * a.N * a.V, *B = a.Hi + (uint64_t)kf * a.E * a.V;
    float2 *C = a.Yt + (uint64_t)i * a.N * a.E;
    const uint32_t n0 = blockIdx.x * kTM, e0 = blockIdx.y * kTN;
    const uint32_t tx = threadIdx.x % 16, ty = threadIdx.x / 16; // thread computes rows n0 + tx + 16 r, columns e0 + ty + 16 c
    float2 acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = make_float2(0.f, 0.f);
    for (uint32_t v0 = 0; v0 < a.V; v0 += kTK) {
        for (uint32_t idx = threadIdx.x; idx < kTK * kTM; idx += 256) {
            const uint32_t m = idx % kTM, k = idx / kTM;
            As[k][m] = (n0 + m < a.N && v0 + k < a.V) ? __ldg(A + (n0 + m) + (uint64_t)a.N * (v0 + k)) : make_float2(0.f, 0.f);
            Bs[k][m] = (e0 + m < a.E && v0 + k < a.V) ? __ldg(B + (e0 + m) + (uint64_t)a.E * (v0 + k)) : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTK; ++k) {
            float2 av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) av[r] = As[k][tx + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; ++c) bv[c] = Bs[k][ty + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[r][c].x = fmaf(av[r].x, bv[c].x, fmaf(-av[r].y, bv[c].y, acc[r][c].x));
                    acc[r][c].y = fmaf(av[r].x, bv[c].y, fmaf(av[r].y, bv[c].x, acc[r][c].y));
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t n = n0 + tx + 16 * r, e = e0 + ty + 16 * c;
            if (n < a.N && e < a.E) C[n + (uint64_t)a.N * e] = acc[r][c];
        }
}
} // namespace

// dt_host: V doubles t0(v) - min(t0) (all zero -> no phase pass).  returns 0, a cudaError_t (> 0), -3 (unsupported T)
int launch_refocus(void *y, const void *x, const void *Hi, const double *dt_host, uint64_t T, uint64_t N, uint64_t V, uint64_t E,
                   double fs, cudaStream_t st) {
    if (T == 0 || N == 0 || V == 0 || E == 0) return 0;
    if ((T & (T - 1)) != 0 || T > 8192) return -3;
    RefocusArgs a{};
    a.x = (const float2 *)x; a.Hi = (const float2 *)Hi; a.y = (float2 *)y;
    a.T = (uint32_t)T; a.N = (uint32_t)N; a.V = (uint32_t)V; a.E = (uint32_t)E;
    uint32_t lg = 0;
    while ((1ull << lg) < T) ++lg;
    a.log2T = lg;
    bool any = false;
    for (uint64_t v = 0; v < V; ++v) any = any || dt_host[v] != 0.0;
    double *ddt = nullptr;
    float2 *ph = nullptr;
    cudaError_t e = ws_alloc((void **)&a.Xt, sizeof(float2) * T * N * V, st);
    if (e == cudaSuccess) e = ws_alloc((void **)&a.Yt, sizeof(float2) * T * N * E, st);
    if (e == cudaSuccess && any) {
        e = ws_alloc((void **)&ddt, sizeof(double) * V, st);
        if (e == cudaSuccess) e = ws_alloc((void **)&ph, sizeof(float2) * T * V, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ddt, dt_host, sizeof(double) * V, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            refocus_phase_kernel<<<(unsigned)((T * V + 255) / 256), 256, 0, st>>>(ph, ddt, a.T, a.V, fs);
            count_launch(1);
            e = cudaGetLastError();
        }
        a.phase = ph;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(float2) * (T + kG * padded_len(T));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(refocus_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(refocus_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024));
    if (e == cudaSuccess) {
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        const uint64_t gn = (N + kG - 1) / kG, cap = (uint64_t)sms * per_sm;
        const uint64_t g1 = gn * V, g3 = gn * E;
        refocus_fwd_kernel<<<(unsigned)(g1 < cap ? g1 : cap), 256, smem, st>>>(a);
        count_launch(1);
        e = cudaGetLastError();
        if (e == cudaSuccess) {
            dim3 grid((unsigned)((N + kTM - 1) / kTM), (unsigned)((E + kTN - 1) / kTN), (unsigned)T);
            refocus_gemm_kernel<<<grid, 256, 0, st>>>(a);
            count_launch(1);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
            refocus_inv_kernel<<<(unsigned)(g3 < cap ? g3 : cap), 256, smem, st>>>(a);
            count_launch(1);
            e = cudaGetLastError();
        }
    }
    if (ph) ws_free(ph, st);
    if (ddt) ws_free(ddt, st);
    if (a.Yt) ws_free(a.Yt, st);
    if (a.Xt) ws_free(a.Xt, st);
    return (int)e;
}

} // namespace qups
