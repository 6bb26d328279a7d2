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
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "das_args.cuh"
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

// ---------------------------------------------------------------------------------------------------------
// Binned variant (the default): per (trace, 2048-sample block, scatterer chunk) the staged entries are bucketed
// by arrival sample with a STABLE counting sort (warp-ordered ranks via match_any), so an output sample visits only
// the ~(kspan + bin)/S fraction of scatterers that can reach it instead of testing all of them.  Deterministic (no
// atomics, fixed order: by bucket, then by scatterer index); the per-sample sum order differs from the oracle's
// plain scatterer order, hence tolerance-level (not bit-level) parity.  greens_kernel above stays as the exact-order
// variant (QUPS_B200_GREENS=simple) and as the fallback for very long waveforms.
constexpr int kGBW = 32;      // bucket width in output samples
constexpr int kGNB = 128;     // max buckets per block: (2048 + kspan + 4)/32 + 1 must fit

template <typename DK, typename DOUT, typename R>
__global__ void __launch_bounds__(kGThreads) greens_binned_kernel(const GreensDev<R> p, DOUT *y, const R *Pi, const R *amp,
                                                                  const R *Pr, const R *Pv, const DK *kern) {
    constexpr int kGChunk = (sizeof(R) == 4) ? 2048 : 1024;
    constexpr int kW = kGThreads / 32;
    __shared__ R s_ttx[kGChunk], s_trx[kGChunk], s_w[kGChunk];
    __shared__ float s_c[kGChunk];
    __shared__ unsigned short s_perm[kGChunk];
    __shared__ int s_cnt[kW][kGNB];
    __shared__ int s_start[kGNB + 1];
    const uint64_t n = blockIdx.x, m = blockIdx.y;
    const uint64_t EE = p.E * p.E, total = p.I * EE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DOUT *yt = y + p.S * (n + p.N * m);
    const float kspan = (float)p.kspan;
    const int nbk = min(kGNB, (int)((kGThreads * 8 + kspan + 4.f) / kGBW) + 2);
    for (uint64_t sb = 0; sb < p.S; sb += (uint64_t)kGThreads * 8) {
        cplx<R> acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = {R(0), R(0)};
        // bucket origin: arrivals c in [org, org + nbk*BW) can reach this block
        const float org = (float)(p.n0 + (long long)sb) - kspan - 2.f;
        for (uint64_t e0 = 0; e0 < total; e0 += kGChunk) {
            const int cnt = (int)((total - e0 < (uint64_t)kGChunk) ? (total - e0) : kGChunk);
            __syncthreads();
            for (int q = threadIdx.x; q < kW * kGNB; q += kGThreads) (&s_cnt[0][0])[q] = 0;
            for (int q = threadIdx.x; q < cnt; q += kGThreads) {
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
                s_trx[q] = mul_rn(-p.fsr, trx);
                s_w[q] = div_rn(att, p.fsr);
                s_c[q] = (float)add_rn(add_rn(ttx, trx), p.t0s);
            }
            __syncthreads();
            // ---- stable counting sort of the chunk by bucket: warp w owns a contiguous, ascending range ----
            const int per = ((cnt + kW * 32 - 1) / (kW * 32)) * 32;
            const int q0 = warp * per, q1 = min(cnt, q0 + per);
            for (int qb = q0; qb < q1; qb += 32) {
                const int q = qb + lane;
                int b = -1;
                if (q < q1) {
                    const float rel = (s_c[q] - org) * (1.f / kGBW);
                    if (rel >= 0.f && rel < (float)nbk) b = (int)rel;
                }
                const unsigned mk = __match_any_sync(0xffffffffu, b);
                if (b >= 0 && lane == (__ffs((int)mk) - 1)) s_cnt[warp][b] += __popc(mk);
                __syncwarp();
            }
            __syncthreads();
            // bucket totals (one thread per bucket) -> exclusive scan (one warp) -> per-warp write cursors
            if ((int)threadIdx.x < nbk) {
                int tot = 0;
                for (int w = 0; w < kW; ++w) tot += s_cnt[w][threadIdx.x];
                s_start[threadIdx.x + 1] = tot;
            }
            __syncthreads();
            if (warp == 0) {
                int carry = 0;
                for (int b0 = 0; b0 < nbk; b0 += 32) {
                    const int b = b0 + lane;
                    int v = (b < nbk) ? s_start[b + 1] : 0, incl = v;
                    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
                    if (b < nbk) s_start[b + 1] = carry + incl;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) s_start[0] = 0;
            }
            __syncthreads();
            if ((int)threadIdx.x < nbk) {
                int run = s_start[threadIdx.x];
                for (int w = 0; w < kW; ++w) { const int c = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = run; run += c; }
            }
            __syncthreads();
            for (int qb = q0; qb < q1; qb += 32) {
                const int q = qb + lane;
                int b = -1;
                if (q < q1) {
                    const float rel = (s_c[q] - org) * (1.f / kGBW);
                    if (rel >= 0.f && rel < (float)nbk) b = (int)rel;
                }
                const unsigned mk = __match_any_sync(0xffffffffu, b);
                if (b >= 0) {
                    const int pos = s_cnt[warp][b] + __popc(mk & ((1u << lane) - 1u));
                    s_perm[pos] = (unsigned short)q;
                }
                __syncwarp();
                if (b >= 0 && lane == (__ffs((int)mk) - 1)) s_cnt[warp][b] += __popc(mk);
                __syncwarp();
            }
            __syncthreads();
            // ---- evaluation: each owned sample walks only the buckets that can reach it ----
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int sr = threadIdx.x + k * kGThreads;                 // sample inside the block
                const uint64_t s = sb + (uint64_t)sr;
                if (s >= p.S) continue;
                const R tv = (R)(p.n0 + (long long)s);
                // arrivals c with tv - c in [-2, kspan + 2]  <=>  c - org in [sr, sr + kspan + 4]
                const int b0 = sr / kGBW, b1 = min(nbk - 1, (int)(((float)sr + kspan + 4.f) / kGBW));
                const int e_lo = s_start[b0], e_hi = s_start[b1 + 1];
                for (int e = e_lo; e < e_hi; ++e) {
                    const int q = s_perm[e];
                    const R d = tv - (R)s_c[q];
                    if (d >= R(-2) && d <= p.kspan + R(2)) {
                        R t1 = sub_rn(tv, s_ttx[q]);
                        t1 = sub_rn(t1, p.t0s);
                        t1 = mul_rn(p.fsr, t1);
                        const R xq = add_rn(R(1), add_rn(t1, s_trx[q]));
                        const cplx<R> v = interp1<DK>(kern, (long)p.T, xq, p.interp);
                        acc[k].re = add_rn(acc[k].re, mul_rn(s_w[q], v.re));
                        acc[k].im = add_rn(acc[k].im, mul_rn(s_w[q], v.im));
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

// ---------------------------------------------------------------------------------------------------------
// Convolution variant (default for fsr == 1, fp32, nearest / linear / cubic).
//
// With the kernel sampled at the output rate every scatterer contributes a SHIFTED copy of the same interpolated
// waveform: for output sample t = ci + p (ci = ceil(arrival), f = ci - arrival in [0,1)) the interpolation weights
// w_j(f) do not depend on p, so
//     x(t) = sum_j sum_{p=0}^{K-2} kx[p-1+j] * I_j[t-p]  +  kx[K-1] * I_z[t-(K-1)],
//     I_j[tau] = sum_{scatterers with ci = tau} att * w_j(f)      (I_z: those with f == 0, the closed end xq == K)
// where kx is the kernel extended by interp1's end padding (kx[-1] = 3k0-3k1+k2, kx[K] likewise).  The p-range is
// exactly the support of interp1(kern, xq, interp, 0) (xq in [1,K]), so the result is the oracle's sum, re-associated:
// I x K x 4 multiply-adds per trace become S x K x 4 (10 k scatterers, S ~ 2.5 k samples: 4x fewer, and regular).
// Per trace: (1) arrival / attenuation / weights per scatterer in chunks, (2) stable counting sort of the chunk by
// 32-sample bucket (warp-ordered ranks, no atomics), (3) each train position gathers the entries of its bucket with
// ci == tau in sorted order (deterministic), (4) the five short real x complex convolutions from shared memory.
// Arrival times are evaluated in fp64 from the fp32 geometry: at ~4000 samples an fp32 delay carries ~1e-4 samples
// of rounding noise (the fp32 oracle has it too, the reference's own CPU-vs-GPU bar is 1e-3, test/SimTest.m:327-357);
// with fp64 delays this kernel matches the fp64 oracle to ~2e-6 instead.  greens_kernel / greens_binned_kernel keep
// the oracle's fp32 sequence (QUPS_B200_GREENS=simple|binned).
#ifndef QUPS_GREENS_THREADS
#define QUPS_GREENS_THREADS 256
#endif
#ifndef QUPS_GREENS_CHUNK
#define QUPS_GREENS_CHUNK 1536
#endif
#ifndef QUPS_GREENS_NO
#define QUPS_GREENS_NO 11
#endif
#ifdef QUPS_GREENS_MINB
#define QUPS_GREENS_BOUNDS __launch_bounds__(QUPS_GREENS_THREADS, QUPS_GREENS_MINB)
#else
#define QUPS_GREENS_BOUNDS __launch_bounds__(QUPS_GREENS_THREADS)
#endif
constexpr int kCT = QUPS_GREENS_THREADS;   // threads per CTA of the convolution kernel
constexpr int kCChunk = QUPS_GREENS_CHUNK; // scatterer entries staged per pass
constexpr int kCBW = 32;        // bucket width (train positions)
constexpr int kCNB = 136;       // max buckets: (kCBlock + K) / 32 + 1
constexpr int kCBlock = 3072;   // output samples per train window
constexpr int kNO = QUPS_GREENS_NO;         // step (4): consecutive outputs per thread (odd: conflict-free at a lane stride of kNO words)
constexpr int kPB = 8;          // step (4): kernel taps per register block
constexpr int kConvPad = 12;    // zero samples after every train row (>= kNO - 1)

// Path-length tables (TAB): r_rx depends on (scatterer, receive element) and r_tx on (scatterer, transmit element) only, but a
// trace CTA needs them for ITS (n, m): N x M CTAs each recomputing two fp64 square roots and a division per scatterer was most
// of the per-entry work at C5 scale.  greens_dist_kernel evaluates {r, 1 / max(r, R0)} once per (element, scatterer) with the
// same fp64 operations; the tables ((N + M) E I x 16 B = 82 MB at C5) stay in L2 and a CTA reads its two rows coalesced.
__global__ void __launch_bounds__(256) greens_dist_kernel(double2 *out, const float *Pi, const float *Pe, uint64_t I, uint64_t cnt, double R0) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= I * cnt) return;
    const uint64_t i = g % I, k = g / I;
    const double sx = Pi[3 * i], sy = Pi[3 * i + 1], sz = Pi[3 * i + 2];
    const double ax = sx - Pe[3 * k], ay = sy - Pe[3 * k + 1], az = sz - Pe[3 * k + 2];
    const double r = sqrt(ax * ax + ay * ay + az * az);
    out[g] = make_double2(r, R0 != 0.0 ? 1.0 / fmax(r, R0) : 1.0);
}

template <typename DOUT, bool TAB>
__global__ void QUPS_GREENS_BOUNDS greens_conv_kernel(const GreensDev<float> p, DOUT *y, const float *Pi, const float *amp,
                                                                const float *Pr, const float *Pv, const float2 *kern, int blk,
                                                                double c0, double fs, double t0s, double R0,
                                                                const double2 *Trx, const double2 *Ttx) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int K = (int)p.T;
    const int W = blk + K - 1;                                   // train window: tau in [tau0, tau0 + W)
    const int Wp = (W + kConvPad + 3) & ~3;
    float2 *kx = reinterpret_cast<float2 *>(gsm);                // kx[q + 1], q = -1 .. K
    float *trains = reinterpret_cast<float *>(kx + ((K + 2 + 1) & ~1)); // 5 x Wp
    float *s_w5 = trains + 5 * Wp + ((5 * Wp) & 3 ? 4 - ((5 * Wp) & 3) : 0);   // [5][kCChunk]: att * w_j(f) (j = 0..3), att if f == 0 (j = 4)
    int *s_ci = reinterpret_cast<int *>(s_w5 + 5 * kCChunk);
    int *s_cnt = s_ci + kCChunk;                                 // [kW][kCNB]
    int *s_start = s_cnt + (kCT / 32) * kCNB;              // [kCNB + 1]
    unsigned short *s_perm = reinterpret_cast<unsigned short *>(s_start + kCNB + 1);
    constexpr int kW = kCT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t n = blockIdx.x, m = blockIdx.y;
    const uint64_t EE = p.E * p.E, total = p.I * EE;
    DOUT *yt = y + p.S * (n + p.N * m);

    // extended kernel (interp1's end padding for cubic: v(0) = 3v(1)-3v(2)+v(3), v(T+1) = 3v(T)-3v(T-1)+v(T-2))
    for (int q = tid; q < K; q += kCT) kx[q + 1] = __ldg(kern + q);
    if (tid == 0) {
        float2 lo = make_float2(0.f, 0.f), hi = lo;
        if (p.interp == 2 && K >= 3) {
            const float2 a = __ldg(kern), b = __ldg(kern + 1), c = __ldg(kern + 2);
            lo = make_float2(3.f * a.x - 3.f * b.x + c.x, 3.f * a.y - 3.f * b.y + c.y);
            const float2 d = __ldg(kern + K - 1), e = __ldg(kern + K - 2), f = __ldg(kern + K - 3);
            hi = make_float2(3.f * d.x - 3.f * e.x + f.x, 3.f * d.y - 3.f * e.y + f.y);
        }
        kx[0] = lo; kx[K + 1] = hi;
    }
    const int nbk = min(kCNB, (W + kCBW - 1) / kCBW);
    const double fs_c0 = fs / c0;   // one fp64 division per thread instead of one per (scatterer, trace) entry

    for (uint64_t sb = 0; sb < p.S; sb += (uint64_t)blk) {
        const long long tau0 = p.n0 + (long long)sb - (K - 1);
        for (int r = tid; r < 5 * Wp; r += kCT) trains[r] = 0.f;
        for (int q = tid; q < kW * kCNB; q += kCT) s_cnt[q] = 0;
        for (uint64_t e0 = 0; e0 < total; e0 += kCChunk) {
            const int cnt = (int)((total - e0 < (uint64_t)kCChunk) ? (total - e0) : kCChunk);
            __syncthreads();   // the previous chunk's walk (and its reset of s_cnt) is complete
            // Warp w owns the contiguous, ascending entry range [q0, q1) of the chunk — for step (1) AND for the counting sort,
            // so the bucket of an entry never leaves the registers of the thread that computed it and the two steps need no
            // barrier between them.
            constexpr int kIt = (kCChunk + kW * 32 - 1) / (kW * 32);
            const int per = ((cnt + kW * 32 - 1) / (kW * 32)) * 32;
            const int q0 = warp * per, q1 = min(cnt, q0 + per);
            int bsv[kIt];
            unsigned mkv[kIt];
            // ---- (1) per-entry arrival, attenuation, interpolation weights (planar in shared memory: s_w5[j][q]) -----------
#pragma unroll
            for (int k = 0; k < kIt; ++k) {
                const int q = q0 + k * 32 + lane;
                bsv[k] = -1;
                if (q >= q1) continue;
                const uint64_t e = e0 + q;
                uint64_t i = e, em = 0, en = 0;
                if (EE != 1) { i = e / EE; em = (e % EE) / p.E; en = e % p.E; }   // (E = 1 skips the 64-bit divisions)
                double r_rx, r_tx, att = (double)amp[i];
                if constexpr (TAB) {
                    const double2 a = Trx[(n + p.N * en) * p.I + i], b = Ttx[(m + p.M * em) * p.I + i];
                    r_rx = a.x; r_tx = b.x;
                    att *= a.y * b.y;   // amp / (max(r_rx, R0) max(r_tx, R0)) to 2 ulp of fp64
                } else {
                    const double sx = Pi[3 * i], sy = Pi[3 * i + 1], sz = Pi[3 * i + 2];
                    const float *pr = Pr + 3 * (n + p.N * en), *pv = Pv + 3 * (m + p.M * em);
                    const double ax = sx - pr[0], ay = sy - pr[1], az = sz - pr[2];
                    const double bx = sx - pv[0], by = sy - pv[1], bz = sz - pv[2];
                    r_rx = sqrt(ax * ax + ay * ay + az * az); r_tx = sqrt(bx * bx + by * by + bz * bz);
                    if (R0 != 0.0) att /= (fmax(r_rx, R0) * fmax(r_tx, R0));
                }
                const double c = fma(r_rx + r_tx, fs_c0, t0s);      // arrival (r_rx + r_tx) / c0 * fs + t0s: kernel position of sample t is d = t - c
                const double cc = ceil(c);
                const float f = (float)(cc - c), a = (float)att;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.interp == 2) {            // Keys a = -1/2
                    const float u = f, u2 = u * u;
                    w.x = fmaf(fmaf(-0.5f, u, 1.0f), u, -0.5f) * u;
                    w.y = fmaf(fmaf(1.5f, u, -2.5f), u2, 1.0f);
                    w.z = fmaf(fmaf(-1.5f, u, 2.0f), u, 0.5f) * u;
                    w.w = fmaf(0.5f, u, -0.5f) * u2;
                } else if (p.interp == 1) {     // v(k) + u (v(k+1) - v(k))
                    w.y = 1.f - f; w.z = f;
                } else {                        // round half away from zero
                    if (f < 0.5f) w.y = 1.f; else w.z = 1.f;
                }
                s_w5[q] = a * w.x; s_w5[kCChunk + q] = a * w.y; s_w5[2 * kCChunk + q] = a * w.z; s_w5[3 * kCChunk + q] = a * w.w;
                s_w5[4 * kCChunk + q] = (f == 0.f) ? a : 0.f;
                const double rel = cc - (double)tau0;
                const int ci = (rel >= 0.0 && rel < (double)W && c == c) ? (int)rel : -1;
                s_ci[q] = ci;
                if (ci >= 0) bsv[k] = min(ci / kCBW, nbk - 1);
            }
            // ---- (2) stable counting sort of the chunk by bucket (warp-ordered ranks, no atomics) ------------------------
            // one match_any per 32 entries: its mask serves the count pass here and the scatter pass below
#pragma unroll
            for (int k = 0; k < kIt; ++k) {
                const int b = bsv[k];
                const unsigned mk = __match_any_sync(0xffffffffu, b);
                mkv[k] = mk;
                if (b >= 0 && lane == (__ffs((int)mk) - 1)) s_cnt[warp * kCNB + b] += __popc(mk);
                __syncwarp();
            }
            __syncthreads();
            // bucket starts + per-warp offsets, one warp: lane l owns buckets 5 l .. 5 l + 4 (register prefix), one warp scan
            if (warp == 0) {
                constexpr int kPer = (kCNB + 31) / 32;
                int tot[kPer], sum = 0;
#pragma unroll
                for (int k = 0; k < kPer; ++k) {
                    const int b = lane * kPer + k;
                    int t = 0;
                    if (b < nbk)
                        for (int w = 0; w < kW; ++w) t += s_cnt[w * kCNB + b];
                    tot[k] = t;
                    sum += t;
                }
                int incl = sum;
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
                int run = incl - sum;
#pragma unroll
                for (int k = 0; k < kPer; ++k) {
                    const int b = lane * kPer + k;
                    if (b < nbk) {
                        s_start[b] = run;
                        int r2 = run;
                        for (int w = 0; w < kW; ++w) { const int c = s_cnt[w * kCNB + b]; s_cnt[w * kCNB + b] = r2; r2 += c; }
                    }
                    run += tot[k];
                }
                if (lane == 31) s_start[nbk] = incl;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kIt; ++k) {
                const int b = bsv[k];
                const unsigned mk = mkv[k];
                if (b >= 0) s_perm[s_cnt[warp * kCNB + b] + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)(q0 + k * 32 + lane);
                __syncwarp();
                if (b >= 0 && lane == (__ffs((int)mk) - 1)) s_cnt[warp * kCNB + b] += __popc(mk);
                __syncwarp();
            }
            __syncthreads();
            for (int q = tid; q < kW * kCNB; q += kCT) s_cnt[q] = 0;   // for the next chunk (the walk does not use it)
            // ---- (3) walk the buckets in sorted order and add each entry into its train position --------------------------
            // O(entries) instead of O(positions x bucket size) (round 1: every train position scanned its whole bucket for
            // the entries arriving exactly there — 35 % of the kernel's instructions).  Work item = (train j, bucket): a train
            // position is owned by one thread and its entries arrive in stable (scatterer-index) order, so the sum per position is
            // deterministic; all 8 warps walk (one thread per bucket adding into all five trains left 4 of them idle: 24 % of the
            // stall samples sat on the barrier after it); the next entry is fetched before the current read-modify-write, so the
            // walk costs one shared-memory round trip per entry instead of four dependent ones.
            for (int item = tid; item < 5 * nbk; item += kCT) {
                const int j = item / nbk, bk = item - j * nbk;
                const float *wsrc = s_w5 + j * kCChunk;
                float *tr = trains + j * Wp;
                int e = s_start[bk];
                const int e_hi = s_start[bk + 1];
                if (e < e_hi) {
                    int q = s_perm[e], r = s_ci[q];
                    float v = wsrc[q];
                    for (++e;; ++e) {
                        const bool more = e < e_hi;
                        int rn = 0;
                        float vn = 0.f;
                        if (more) { const int qn = s_perm[e]; rn = s_ci[qn]; vn = wsrc[qn]; }
                        tr[r] += v;
                        if (!more) break;
                        r = rn; v = vn;
                    }
                }
            }
        }
        __syncthreads();
        // ---- (4) x(t) = sum_p sum_j kx[p-1+j] I_j[t-p] + kx[K-1] I_z[t-(K-1)] ----------------------------------
        // A thread owns kNO CONSECUTIVE outputs (odd count: lanes kNO words apart hit 32 distinct banks) and walks p in blocks
        // of kPB: per train the kNO + kPB - 1 samples the block touches are loaded once and feed kNO x kPB products —
        // 7 FMA per shared load instead of the 2 of a one-output-at-a-time loop.  The rows carry kConvPad zeros so the window
        // of the last outputs needs no range test.
        for (int task = tid; task * kNO < blk; task += kCT) {
            const int o0 = task * kNO;
            float2 acc[kNO];
#pragma unroll
            for (int i = 0; i < kNO; ++i) acc[i] = make_float2(0.f, 0.f);
            int pp0 = 0;
            for (; pp0 + kPB - 1 <= K - 2; pp0 += kPB) {
                float2 kk[kPB + 3];
#pragma unroll
                for (int q = 0; q < kPB + 3; ++q) kk[q] = kx[pp0 + q];   // kx[p - 1 + j] at index p + j <= K + 1
                const int base = o0 + (K - 1) - pp0 - (kPB - 1);          // >= 1: o0 >= 0, pp0 + kPB - 1 <= K - 2
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v[kNO + kPB - 1];
                    const float *tr = trains + j * Wp + base;
#pragma unroll
                    for (int m = 0; m < kNO + kPB - 1; ++m) v[m] = tr[m];
#pragma unroll
                    for (int d = 0; d < kPB; ++d)
#pragma unroll
                        for (int i = 0; i < kNO; ++i) {
                            acc[i].x = fmaf(kk[d + j].x, v[i + kPB - 1 - d], acc[i].x);
                            acc[i].y = fmaf(kk[d + j].y, v[i + kPB - 1 - d], acc[i].y);
                        }
                }
            }
            for (; pp0 <= K - 2; ++pp0) {                                 // remaining p (fewer than kPB)
                const float2 k0 = kx[pp0], k1 = kx[pp0 + 1], k2 = kx[pp0 + 2], k3 = kx[pp0 + 3];
#pragma unroll
                for (int i = 0; i < kNO; ++i) {
                    const int r = o0 + i + (K - 1) - pp0;
                    const float i0 = trains[r], i1 = trains[Wp + r], i2 = trains[2 * Wp + r], i3 = trains[3 * Wp + r];
                    acc[i].x = fmaf(k0.x, i0, acc[i].x); acc[i].y = fmaf(k0.y, i0, acc[i].y);
                    acc[i].x = fmaf(k1.x, i1, acc[i].x); acc[i].y = fmaf(k1.y, i1, acc[i].y);
                    acc[i].x = fmaf(k2.x, i2, acc[i].x); acc[i].y = fmaf(k2.y, i2, acc[i].y);
                    acc[i].x = fmaf(k3.x, i3, acc[i].x); acc[i].y = fmaf(k3.y, i3, acc[i].y);
                }
            }
            const float2 kl = kx[K]; // kernel sample K-1: the closed end of interp1's support (xq == K)
#pragma unroll
            for (int i = 0; i < kNO; ++i) {
                const float iz = trains[4 * Wp + o0 + i];
                acc[i].x = fmaf(kl.x, iz, acc[i].x); acc[i].y = fmaf(kl.y, iz, acc[i].y);
                const uint64_t sidx = sb + (uint64_t)(o0 + i);
                if (o0 + i < blk && sidx < p.S)
                    data_traits<DOUT>::store(yt, sidx, {acc[i].x, acc[i].y});
            }
        }
        __syncthreads();
    }
}

static size_t greens_conv_smem(int K, int blk) {
    const int W = blk + K - 1, Wp = (W + kConvPad + 3) & ~3;
    size_t b = sizeof(float2) * ((K + 2 + 1) & ~1);
    b += sizeof(float) * (5 * Wp + 4);
    b += sizeof(float4) * kCChunk + sizeof(float) * kCChunk + sizeof(int) * kCChunk;
    b += sizeof(int) * ((kCT / 32) * kCNB + kCNB + 1);
    b += sizeof(unsigned short) * kCChunk + 16;
    return b;
}

int launch_greens(const qups_greens_params &p, void *y, const void *Pi, const void *a, const void *Pr, const void *Pv,
                  const void *kern, cudaStream_t st) {
    if (p.S == 0 || p.N == 0 || p.M == 0) return 0;
    if (p.N > 0x7fffffffull || p.M > 65535) return -3;
    if (p.dtype == QUPS_F16) {
        // greensh (src/greens.cu:113-122): half2 waveform in, half2 traces out, fp32 geometry.  The waveform (T samples) is widened
        // once — exactly — and the fp32 kernels run as they are: fp32 delays and fp32 accumulation over the scatterers (the
        // reference accumulates in half2), one rounding to half2 per output sample at the end (or none with y_f32).
        float2 *k32 = nullptr, *y32 = nullptr;
        const uint64_t ny = p.S * p.N * p.M;
        cudaError_t e = ws_alloc((void **)&k32, sizeof(float2) * (p.T ? p.T : 1), st);
        if (e == cudaSuccess && !p.y_f32) e = ws_alloc((void **)&y32, sizeof(float2) * ny, st);
        int rc = e == cudaSuccess ? 0 : -4;
        if (rc == 0) rc = launch_half2_to_float2(k32, (const __half2 *)kern, p.T, st);
        if (rc == 0) {
            qups_greens_params q = p;
            q.dtype = QUPS_F32;
            rc = launch_greens(q, p.y_f32 ? y : (void *)y32, Pi, a, Pr, Pv, k32, st);
        }
        if (rc == 0 && !p.y_f32) rc = launch_float2_to_half2((__half2 *)y, y32, ny, st);
        if (k32) ws_free(k32, st);
        if (y32) ws_free(y32, st);
        return rc;
    }
    const uint64_t E = p.E ? p.E : 1;
    dim3 grid((unsigned)p.N, (unsigned)p.M), block(kGThreads);
    // bucketed kernel unless the waveform is so long that its reach exceeds the bucket table, or the exact
    // scatterer-order variant is requested (tests pin bit-exactness on it)
    bool binned = ((double)p.T / p.fsr + 4.0 + kGThreads * 8) / kGBW + 2 <= kGNB;
    if (const char *ev = getenv("QUPS_B200_GREENS")) { if (!strcmp(ev, "simple")) binned = false; }
    // convolution kernel: kernel sampled at the output rate, fp32, tap-linear interpolators, enough scatterers to pay for it
    bool conv = p.dtype == QUPS_F32 && p.fsr == 1.0 && p.interp >= 0 && p.interp <= 2 && p.T >= 3 && p.T <= 1024 && p.I * E * E >= 64;
    if (const char *ev = getenv("QUPS_B200_GREENS")) {
        if (!strcmp(ev, "simple") || !strcmp(ev, "binned")) conv = false;
        else if (strcmp(ev, "conv")) { /* unknown value: keep the automatic choice */ }
    }
    if (p.dtype == QUPS_F32) {
        GreensDev<float> d{p.I, p.S, p.T, p.N, p.M, E, (long long)p.n0, p.interp,
                           (float)p.t0x * (float)p.fs, (float)p.fs, (float)p.fsr, (float)p.c0, (float)p.R0,
                           (float)((double)p.T / p.fsr)};
        if (conv) {
            const int blk = (int)(p.S <= (uint64_t)kCBlock ? ((p.S + 255) / 256) * 256 : 2048);
            const size_t smem = greens_conv_smem((int)p.T, blk);
            cudaError_t e = cudaFuncSetAttribute(greens_conv_kernel<float2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            // path-length tables unless they would be unreasonably large (then every CTA computes its own, as before)
            const uint64_t trx = p.N * E * p.I, ttx = p.M * E * p.I;
            double2 *tab = nullptr;
            if (!getenv("QUPS_B200_GREENS_NOTAB") && (trx + ttx) * sizeof(double2) <= (4ull << 30) &&
                ws_alloc((void **)&tab, (trx + ttx) * sizeof(double2), st) != cudaSuccess) { tab = nullptr; (void)cudaGetLastError(); }
            if (tab) {
                greens_dist_kernel<<<(unsigned)((trx + 255) / 256), 256, 0, st>>>(tab, (const float *)Pi, (const float *)Pr, p.I, p.N * E, p.R0);
                greens_dist_kernel<<<(unsigned)((ttx + 255) / 256), 256, 0, st>>>(tab + trx, (const float *)Pi, (const float *)Pv, p.I, p.M * E, p.R0);
                count_launch(2);
                e = cudaFuncSetAttribute(greens_conv_kernel<float2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e == cudaSuccess)
                    greens_conv_kernel<float2, true><<<grid, kCT, smem, st>>>(d, (float2 *)y, (const float *)Pi, (const float *)a, (const float *)Pr,
                                                                                (const float *)Pv, (const float2 *)kern, blk, p.c0, p.fs, p.t0x * p.fs, p.R0, tab, tab + trx);
                ws_free(tab, st);
                if (e != cudaSuccess) return (int)e;
            } else
            greens_conv_kernel<float2, false><<<grid, kCT, smem, st>>>(d, (float2 *)y, (const float *)Pi, (const float *)a, (const float *)Pr,
                                                                         (const float *)Pv, (const float2 *)kern, blk, p.c0, p.fs, p.t0x * p.fs, p.R0, nullptr, nullptr);
        } else if (binned)
            greens_binned_kernel<float2, float2, float><<<grid, block, 0, st>>>(d, (float2 *)y, (const float *)Pi, (const float *)a,
                                                                                (const float *)Pr, (const float *)Pv, (const float2 *)kern);
        else
            greens_kernel<float2, float2, float><<<grid, block, 0, st>>>(d, (float2 *)y, (const float *)Pi, (const float *)a,
                                                                         (const float *)Pr, (const float *)Pv, (const float2 *)kern);
    } else if (p.dtype == QUPS_F64) {
        GreensDev<double> d{p.I, p.S, p.T, p.N, p.M, E, (long long)p.n0, p.interp,
                            p.t0x * p.fs, p.fs, p.fsr, p.c0, p.R0, (double)p.T / p.fsr};
        if (binned)
            greens_binned_kernel<double2, double2, double><<<grid, block, 0, st>>>(d, (double2 *)y, (const double *)Pi, (const double *)a,
                                                                                   (const double *)Pr, (const double *)Pv, (const double2 *)kern);
        else
            greens_kernel<double2, double2, double><<<grid, block, 0, st>>>(d, (double2 *)y, (const double *)Pi, (const double *)a,
                                                                            (const double *)Pr, (const double *)Pv, (const double2 *)kern);
    } else
        return -3;
    count_launch(1);
    return (int)cudaGetLastError();
}

} // namespace qups
