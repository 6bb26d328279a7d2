// greens.cu — Green's-function point-scatterer simulator for sm_100a.
//
// Replaces src/greens.cu:8-122 (greens_temp / greens / greensf / greensh) and
// the host-side windowing around it (src/UltrasoundSystem.m:678-714).
// Semantics follow the reference's CPU path (src/UltrasoundSystem.m:778-851):
//
//   x(t,n,m) = sum_i sum_{em,en} (att_i / fsr) * interp1(kern, 1 + fsr*(t - tau_tx - t0) - fsr*tau_rx, interp, 0)
//   att_i    = amp_i / (max(r_rx,R0) * max(r_tx,R0))     (R0 ~= 0, else amp_i)
//
// with t the integer output sample index.  One CTA owns one (n,m) trace: the
// per-scatterer delays / weights are computed once per trace into shared
// memory (the reference recomputes two sqrt per (sample, scatterer)), then each
// thread owns output samples and walks the scatterers IN ORDER with a cheap
// window test — deterministic, no atomics, and the per-sample sum order equals
// the oracle's.  Compiled with -fmad=false.
#include "common.cuh"
#include "other_kernels.cuh"

namespace qups {
void count_launch(uint64_t n);

constexpr int kGThreads = 256;

template <typename R> struct GreensDev {
    uint64_t I, S, T, N, M, E;
    long long n0;
    int interp;
    R t0s;   // wv.t0 * fs
    R fs, fsr, c0, R0, kspan;
};

template <typename DK, typename DOUT, typename R>
__global__ void __launch_bounds__(kGThreads) greens_kernel(const GreensDev<R> p, DOUT *y, const R *Pi, const R *amp,
                                                           const R *Pr, const R *Pv, const DK *kern) {
    constexpr int kGChunk = (sizeof(R) == 4) ? 2048 : 1024; // scatterer-subelement entries staged per pass (32 KB)
    __shared__ R s_ttx[kGChunk], s_trx[kGChunk], s_w[kGChunk], s_c[kGChunk];
    const uint64_t n = blockIdx.x, m = blockIdx.y;
    const uint64_t EE = p.E * p.E, total = p.I * EE;
    DOUT *yt = y + p.S * (n + p.N * m);
    // each thread owns samples s = threadIdx.x + k*kGThreads within a block of kGThreads*8 samples
    for (uint64_t sb = 0; sb < p.S; sb += (uint64_t)kGThreads * 8) {
        cplx<R> acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = {R(0), R(0)};
        const R blk_lo = (R)(p.n0 + (long long)sb), blk_hi = (R)(p.n0 + (long long)sb + kGThreads * 8);
        for (uint64_t e0 = 0; e0 < total; e0 += kGChunk) {
            const int cnt = (int)((total - e0 < (uint64_t)kGChunk) ? (total - e0) : kGChunk);
            __syncthreads();
            for (int q = threadIdx.x; q < cnt; q += kGThreads) {
                // entry order = for s, for em, for en   (src/UltrasoundSystem.m:785-790)
                const uint64_t e = e0 + q, i = e / EE, em = (e % EE) / p.E, en = e % p.E;
                const R sx = Pi[3 * i], sy = Pi[3 * i + 1], sz = Pi[3 * i + 2];
                const R *pr = Pr + 3 * (n + p.N * en), *pv = Pv + 3 * (m + p.M * em);
                const R r_rx = rx_dist(sx, sy, sz, pr[0], pr[1], pr[2]);
                const R r_tx = rx_dist(sx, sy, sz, pv[0], pv[1], pv[2]);
                R att;
                if (p.R0 != R(0)) {
                    const R a = (r_rx > p.R0) ? r_rx : p.R0, b = (r_tx > p.R0) ? r_tx : p.R0;
                    att = div_rn(amp[i], mul_rn(a, b));
                } else att = amp[i];
                const R trx = mul_rn(div_rn(r_rx, p.c0), p.fs), ttx = mul_rn(div_rn(r_tx, p.c0), p.fs);
                s_ttx[q] = ttx;
                s_trx[q] = mul_rn(-p.fsr, trx); // t2 = -fsr * tau_rx
                s_w[q] = div_rn(att, p.fsr);
                s_c[q] = add_rn(add_rn(ttx, trx), p.t0s); // arrival sample (window test only)
            }
            __syncthreads();
            for (int q = 0; q < cnt; ++q) {
                const R c = s_c[q];
                // block-uniform reject: no sample of this block can see the scatterer
                if (c > blk_hi + R(2) || c + p.kspan < blk_lo - R(2)) continue;
                const R ttx = s_ttx[q], t2 = s_trx[q], wg = s_w[q];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t s = sb + threadIdx.x + (uint64_t)k * kGThreads;
                    const R tv = (R)(p.n0 + (long long)s);
                    const R d = tv - c;
                    if (d >= R(-2) && d <= p.kspan + R(2)) {
                        R t1 = sub_rn(tv, ttx);
                        t1 = sub_rn(t1, p.t0s);
                        t1 = mul_rn(p.fsr, t1);
                        const R xq = add_rn(R(1), add_rn(t1, t2));
                        const cplx<R> v = interp1<DK>(kern, (long)p.T, xq, p.interp);
                        acc[k].re = add_rn(acc[k].re, mul_rn(wg, v.re));
                        acc[k].im = add_rn(acc[k].im, mul_rn(wg, v.im));
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint64_t s = sb + threadIdx.x + (uint64_t)k * kGThreads;
            if (s < p.S)
                data_traits<DOUT>::store(yt, s, {(typename data_traits<DOUT>::real)acc[k].re,
                                                 (typename data_traits<DOUT>::real)acc[k].im});
        }
    }
}

int launch_greens(const qups_greens_params &p, void *y, const void *Pi, const void *a, const void *Pr, const void *Pv,
                  const void *kern, cudaStream_t st) {
    if (p.S == 0 || p.N == 0 || p.M == 0) return 0;
    if (p.N > 0x7fffffffull || p.M > 65535) return -3;
    const uint64_t E = p.E ? p.E : 1;
    dim3 grid((unsigned)p.N, (unsigned)p.M), block(kGThreads);
    if (p.dtype == QUPS_F32) {
        GreensDev<float> d{p.I, p.S, p.T, p.N, p.M, E, (long long)p.n0, p.interp,
                           (float)p.t0x * (float)p.fs, (float)p.fs, (float)p.fsr, (float)p.c0, (float)p.R0,
                           (float)((double)p.T / p.fsr)};
        greens_kernel<float2, float2, float><<<grid, block, 0, st>>>(d, (float2 *)y, (const float *)Pi, (const float *)a,
                                                                     (const float *)Pr, (const float *)Pv, (const float2 *)kern);
    } else if (p.dtype == QUPS_F64) {
        GreensDev<double> d{p.I, p.S, p.T, p.N, p.M, E, (long long)p.n0, p.interp,
                            p.t0x * p.fs, p.fs, p.fsr, p.c0, p.R0, (double)p.T / p.fsr};
        greens_kernel<double2, double2, double><<<grid, block, 0, st>>>(d, (double2 *)y, (const double *)Pi, (const double *)a,
                                                                        (const double *)Pr, (const double *)Pv, (const double2 *)kern);
    } else
        return -3;
    count_launch(1);
    return (int)cudaGetLastError();
}

} // namespace qups
