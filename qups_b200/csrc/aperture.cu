// aperture.cu — aperture-domain post-processing of the per-receive beamformed cube.
//
// Consumers of DAS(..., 'keep_rx', true) output (I x N complex): the reference evaluates them with whole-array MATLAB
// expressions, several temporaries of the size of the cube each:
//   cohfac  kern/cohfac.m   r = |sum_n b|^2 / sum_n |b|^2 / N
//   dmas    kern/dmas.m     b = sum_{lag in L} sum_n b(n) b(n+lag);  out = exp(1j angle(b)) sqrt(|b|)
//   pcf     kern/pcf.m      phase coherence factor: w = max(0, 1 - gamma/sg0 * min(std(phi), std(phi - pi sign(phi)))), sg0 = sqrt(pi/3)
//   slsc    kern/slsc.m     short-lag spatial coherence, "average" (per-sample normalised) and "ensemble" estimators,
//                           time-sample dimension kdim singleton
// Here each is ONE pass: the cube is viewed as C x A x S (A = aperture dimension `dim`, C / S = the dimensions before /
// after it), one thread per output element (c, s), coalesced along C, walking the aperture with stride C.  The pair sums
// of dmas / slsc re-read the thread's own column through L1/L2 (A <= ~1k elements x 8 B per thread).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "other_kernels.cuh"

namespace qups {
void count_launch(uint64_t n);
cudaError_t ws_alloc(void **p, size_t bytes, cudaStream_t st);
cudaError_t ws_free(void *p, cudaStream_t st);

template <typename R> struct c2 { using type = float2; };
template <> struct c2<double> { using type = double2; };

template <typename R> __device__ __forceinline__ R rsqrt_(R x) { return R(1) / sqrt(x); }
__device__ __forceinline__ float fast_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double fast_rsqrt(double x) { return 1.0 / sqrt(x); }
template <typename R> __device__ __forceinline__ R atan2_(R y, R x) { return atan2(y, x); }
__device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }

template <typename R>
__global__ void __launch_bounds__(256) aperture_kernel(int op, void *out, void *out2, const void *bin, const unsigned char *lagmask,
                                                       uint64_t C, uint64_t A, uint64_t S, uint32_t nlags, R gamma) {
    using V = typename c2<R>::type;
    const V *b = reinterpret_cast<const V *>(bin);
    const uint64_t total = C * S;
    const R PI = R(3.14159265358979323846);
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = e % C, s = e / C;
        const V *col = b + c + s * C * A; // element n at col[n * C]
        if (op == QUPS_APD_COHFAC) {
            R sr = 0, si = 0, p = 0;
            for (uint64_t n = 0; n < A; ++n) {
                const V v = col[n * C];
                sr += v.x; si += v.y;
                p += v.x * v.x + v.y * v.y;
            }
            reinterpret_cast<R *>(out)[e] = (sr * sr + si * si) / p / (R)A;
        } else if (op == QUPS_APD_DMAS) {
            R zr = 0, zi = 0;
            for (uint64_t lag = 1; lag < A; ++lag) {
                if (!lagmask[lag]) continue;
                R ar = 0, ai = 0;
                for (uint64_t n = 0; n + lag < A; ++n) { // sum(sub(bn,1:N-i) .* sub(bn,1+i:N)) — no conjugate
                    const V u = col[n * C], v = col[(n + lag) * C];
                    ar += u.x * v.x - u.y * v.y;
                    ai += u.x * v.y + u.y * v.x;
                }
                zr += ar; zi += ai;
            }
            const R mag = sqrt(sqrt(zr * zr + zi * zi)); // sqrt(abs(b))
            const R ph = atan2_(zi, zr);
            R sn, cs;
            sincos(ph, &sn, &cs);
            V o; o.x = mag * cs; o.y = mag * sn;
            reinterpret_cast<V *>(out)[e] = o;
        } else if (op == QUPS_APD_PCF) {
            // std(phi, 1, dim, "omitnan") two-pass, population normalisation; once for phi, once for phi - pi*sign(phi)
            R m0 = 0, m1 = 0; uint64_t cnt = 0;
            for (uint64_t n = 0; n < A; ++n) {
                const V v = col[n * C];
                const R ph = atan2_(v.y, v.x);
                if (ph == ph) { const R sg = ph > 0 ? R(1) : (ph < 0 ? R(-1) : R(0)); m0 += ph; m1 += ph - PI * sg; ++cnt; }
            }
            R s0 = 0, s1 = 0;
            if (cnt) {
                m0 /= (R)cnt; m1 /= (R)cnt;
                for (uint64_t n = 0; n < A; ++n) {
                    const V v = col[n * C];
                    const R ph = atan2_(v.y, v.x);
                    if (ph == ph) { const R sg = ph > 0 ? R(1) : (ph < 0 ? R(-1) : R(0)); const R d0 = ph - m0, d1 = ph - PI * sg - m1; s0 += d0 * d0; s1 += d1 * d1; }
                }
                s0 = sqrt(s0 / (R)cnt); s1 = sqrt(s1 / (R)cnt);
            } else { s0 = s1 = R(0) / R(0); }
            const R sf = fmin(s0, s1);
            const R sg0 = sqrt(PI / R(3));
            reinterpret_cast<R *>(out)[e] = fmax(R(0), R(1) - (gamma / sg0) * sf);
            if (out2) reinterpret_cast<R *>(out2)[e] = sf;
        } else { // SLSC
            const bool avg = op == QUPS_APD_SLSC_AVERAGE;
            R zr = 0, zi = 0, na = 0, nb = 0;
            for (uint64_t lag = 0; lag < A; ++lag) {
                if (!lagmask[lag]) continue;
                R ar = 0, pa = 0;
                if (lag == 0) { // ismember(H, lags) with a zero lag selects the diagonal: each element paired with itself once
                    for (uint64_t n = 0; n < A; ++n) {
                        const V u = col[n * C];
                        const R q = u.x * u.x + u.y * u.y;
                        ar += avg ? (q > 0 ? R(1) : R(0)) : q;
                        pa += q;
                    }
                    if (avg) zr += ar / (R)A / R(2) / (R)nlags;
                    else { zr += ar; na += pa; nb += pa; }
                    continue;
                }
                for (uint64_t n = 0; n + lag < A; ++n) {
                    V u = col[n * C], v = col[(n + lag) * C];
                    if (avg) { // x ./ vecnorm(x,2,kdim), nan2zero
                        const R mu = sqrt(u.x * u.x + u.y * u.y), mv = sqrt(v.x * v.x + v.y * v.y);
                        u.x = mu > 0 ? u.x / mu : 0; u.y = mu > 0 ? u.y / mu : 0;
                        v.x = mv > 0 ? v.x / mv : 0; v.y = mv > 0 ? v.y / mv : 0;
                    }
                    // both orders (i,j) and (j,i): conj(u) v + conj(v) u = 2 Re(conj(u) v)
                    ar += R(2) * (u.x * v.x + u.y * v.y);
                    pa += u.x * u.x + u.y * u.y + v.x * v.x + v.y * v.y;
                }
                if (avg) zr += ar / (R)(A - lag) / R(2) / (R)nlags; // W = S ./ (A - H) / 2 / L
                else { zr += ar; na += pa; nb += pa; }
            }
            V o;
            if (avg) { o.x = zr; o.y = zi; }
            else { const R sc = rsqrt_(na) * rsqrt_(nb); o.x = isfinite(sc) ? zr * sc : 0; o.y = 0; }
            reinterpret_cast<V *>(out)[e] = o;
        }
    }
}

// ---- pair-sum operators on a shared-memory tile -------------------------------------------------------------------------
// dmas / slsc sum products of aperture elements `lag` apart: sum_{lag in L} sum_n f(b(n), b(n + lag)).  The one-thread-per-
// output kernel above re-reads its column through L1/L2 once per lag (C2 keep_rx cube, L = 16: dmas 9.8 ms, slsc-average
// 20 ms for a 2.15 GB cube).  Here a CTA stages 32 adjacent columns x the whole aperture in shared memory ONCE (slsc
// "average": already normalised, x / |x|), its 8 warps share the (lag block, aperture chunk) work items, and every item
// is register-blocked 4 elements x 4 consecutive lags: 11 shared loads feed 16 pair products.  The per-pair power sums of
// the "ensemble" estimator collapse to prefix sums: sum over pairs of |u|^2 + |v|^2 = P(A - lag) + Ptot - P(lag).
// Every total is linear in the item sums, so the fixed item -> warp assignment and the in-order reduction over the warps
// keep the result deterministic.  pcf uses the same tile: one atan2 per element instead of two.
template <typename R> struct apd_acc { R zr, zi, na; };

template <typename R, int OP>
__global__ void __launch_bounds__(256) aperture_tile_kernel(void *out, void *out2, const void *bin, const unsigned char *lagmask,
                                                            uint64_t C, uint32_t A, uint64_t S, uint32_t nlags, R gamma) {
    using V = typename c2<R>::type;
    extern __shared__ __align__(16) unsigned char ap_smem[];
    constexpr uint32_t kPadRows = 12;                          // zero rows after the aperture: the register-blocked loads of the
                                                               // pair loop run past the last valid pair without any range test
    V *tile = reinterpret_cast<V *>(ap_smem);                  // [A + kPadRows][32]
    R *pre = reinterpret_cast<R *>(tile + (size_t)(A + kPadRows) * 32); // prefix powers [A + 1][32] (slsc ensemble), phases [A][32] (pcf)
    __shared__ R red[8][3][32];
    __shared__ R segtot[8][32];                                // per-warp segment totals of the prefix powers (slsc ensemble)
    __shared__ uint16_t s_act[4096];                           // lag blocks with at least one requested lag (A < 16384)
    __shared__ uint32_t s_nact;
    const V *b = reinterpret_cast<const V *>(bin);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t total = C * S, e = (uint64_t)blockIdx.x * 32 + lane;
    const bool valid = e < total;
    const uint64_t c = valid ? e % C : 0, sidx = valid ? e / C : 0;
    const V *col = b + c + sidx * C * A;
    const R PI = R(3.14159265358979323846);
    if (threadIdx.x == 0) s_nact = 0;
    for (uint32_t r = threadIdx.x; r < kPadRows * 32; r += 256) { V z; z.x = R(0); z.y = R(0); tile[(size_t)A * 32 + r] = z; }
    const V *src = col + (uint64_t)warp * C;                   // this warp's rows: n = warp, warp + 8, ...
    const uint64_t step = 8 * C;
    // 8 rows of this warp in flight per batch: the loads are issued before the first store (one exposed DRAM latency per
    // batch instead of one per row)
    for (uint32_t n = warp; n < A; n += 64, src += 8 * step) {
        V vb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            vb[u].x = R(0); vb[u].y = R(0);
            if (valid && n + 8 * u < A) vb[u] = src[(uint64_t)u * step];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (n + 8 * u >= A) break;
            V v = vb[u];
            if (OP == QUPS_APD_SLSC_AVERAGE) { // x ./ vecnorm(x, 2, kdim), nan2zero
                const R q2 = v.x * v.x + v.y * v.y;
                const R inv = q2 > 0 ? fast_rsqrt(q2) : R(0); // x * rsqrt(|x|^2): 2 ulp of x / |x| (fp32: MUFU.RSQ instead of sqrt + divide)
                v.x *= inv; v.y *= inv;
            }
            tile[(n + 8 * u) * 32 + lane] = v;
            if (OP == QUPS_APD_PCF) pre[(n + 8 * u) * 32 + lane] = atan2_(v.y, v.x);
        }
    }
    __syncthreads();
    R zr = 0, zi = 0, na = 0;
    if (OP == QUPS_APD_PCF) {
        // std(phi, 1, dim, "omitnan") two-pass over the staged phases; once for phi, once for phi - pi*sign(phi).  Warp w
        // sums elements n = w, w + 8, ...; the partial sums are combined in warp order
        R m0 = 0, m1 = 0, cnt = 0;
        for (uint32_t n = warp; n < A; n += 8) {
            const R ph = pre[n * 32 + lane];
            if (ph == ph) { const R sg = ph > 0 ? R(1) : (ph < 0 ? R(-1) : R(0)); m0 += ph; m1 += ph - PI * sg; cnt += 1; }
        }
        red[warp][0][lane] = m0; red[warp][1][lane] = m1; red[warp][2][lane] = cnt;
        __syncthreads();
        m0 = m1 = cnt = 0;
        for (int w = 0; w < 8; ++w) { m0 += red[w][0][lane]; m1 += red[w][1][lane]; cnt += red[w][2][lane]; }
        __syncthreads();
        R s0 = 0, s1 = 0;
        if (cnt > 0) {
            m0 /= cnt; m1 /= cnt;
            for (uint32_t n = warp; n < A; n += 8) {
                const R ph = pre[n * 32 + lane];
                if (ph == ph) { const R sg = ph > 0 ? R(1) : (ph < 0 ? R(-1) : R(0)); const R d0 = ph - m0, d1 = ph - PI * sg - m1; s0 += d0 * d0; s1 += d1 * d1; }
            }
        }
        red[warp][0][lane] = s0; red[warp][1][lane] = s1;
        __syncthreads();
        if (warp == 0 && valid) {
            s0 = s1 = 0;
            for (int w = 0; w < 8; ++w) { s0 += red[w][0][lane]; s1 += red[w][1][lane]; }
            if (cnt > 0) { s0 = sqrt(s0 / cnt); s1 = sqrt(s1 / cnt); } else { s0 = s1 = R(0) / R(0); }
            const R sf = fmin(s0, s1);
            const R sg0 = sqrt(PI / R(3));
            reinterpret_cast<R *>(out)[e] = fmax(R(0), R(1) - (gamma / sg0) * sf);
            if (out2) reinterpret_cast<R *>(out2)[e] = sf;
        }
        return;
    }
    const uint32_t seg = (A + 7) / 8;   // prefix powers: warp w scans rows [w seg, (w + 1) seg), the segment totals join at use
    auto P = [&](uint32_t k) -> R {     // P(k) = sum_{n < k} |x_n|^2 of this lane's column
        if (k == 0) return R(0);
        R v = pre[k * 32 + lane];
        for (uint32_t w = 0; w < (k - 1) / seg; ++w) v += segtot[w][lane];
        return v;
    };
    if (OP == QUPS_APD_SLSC_ENSEMBLE) {
        R p = 0;
        for (uint32_t n = warp * seg; n < min(A, (warp + 1) * seg); ++n) { const V v = tile[n * 32 + lane]; p += v.x * v.x + v.y * v.y; pre[(n + 1) * 32 + lane] = p; }
        segtot[warp][lane] = p;
        __syncthreads();
    }
    // work items: (block of 4 consecutive lags starting at 1 + 4 blk, chunk of 64 aperture elements); only the lag blocks
    // that hold a requested lag are listed (in ascending order: the item -> warp assignment stays a function of the
    // arguments alone, so the result is deterministic)
    const uint32_t NB = (A + 2) / 4, NK = (A + 63) / 64; // lags 1 .. A-1
    if (warp == 0) {
        for (uint32_t b0 = 0; b0 < NB; b0 += 32) {
            const uint32_t blk = b0 + lane, lag0 = 1 + 4 * blk;
            bool any = false;
            if (blk < NB)
                for (int j = 0; j < 4; ++j) any = any || ((lag0 + j < A) && lagmask[lag0 + j]);
            const unsigned bal = __ballot_sync(0xffffffffu, any);
            const uint32_t base = s_nact;
            if (any) s_act[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)blk;
            __syncwarp();
            if (lane == 0) s_nact = base + __popc(bal);
            __syncwarp();
        }
    }
    __syncthreads();
    const uint32_t nact = s_nact;
    for (uint32_t it = warp; it < nact * NK; it += 8) {
        const uint32_t blk = s_act[it / NK], k = it % NK, lag0 = 1 + 4 * blk;
        bool m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = (lag0 + j < A) && lagmask[lag0 + j];
        R ar[4] = {0, 0, 0, 0}, ai[4] = {0, 0, 0, 0};
        // pairs (n, n + lag) need n + lag < A: past n = A - lag0 every partner is a zero row; chunk ends are multiples of 4
        // (or the end of the aperture, followed by zero rows), so neither load needs a range test
        const uint32_t nend = min(min(A, k * 64 + 64), A > lag0 ? A - lag0 : 0u);
        const V *tu = tile + lane, *tv = tile + (size_t)lag0 * 32 + lane;
        for (uint32_t n0 = k * 64; n0 < nend; n0 += 4) {
            V u[4], v[7];
#pragma unroll
            for (int i = 0; i < 4; ++i) u[i] = tu[(n0 + i) * 32];
#pragma unroll
            for (int t = 0; t < 7; ++t) v[t] = tv[(n0 + t) * 32];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (OP == QUPS_APD_DMAS) { // b(n) b(n + lag), no conjugate
                        ar[j] += u[i].x * v[i + j].x - u[i].y * v[i + j].y;
                        ai[j] += u[i].x * v[i + j].y + u[i].y * v[i + j].x;
                    } else {                   // both orders (i,j) and (j,i): 2 Re(conj(u) v)
                        ar[j] += R(2) * (u[i].x * v[i + j].x + u[i].y * v[i + j].y);
                    }
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!m[j]) continue;
            const uint32_t lag = lag0 + j;
            if (OP == QUPS_APD_DMAS) { zr += ar[j]; zi += ai[j]; }
            else if (OP == QUPS_APD_SLSC_AVERAGE) zr += ar[j] / (R)(A - lag) / R(2) / (R)nlags; // W = S ./ (A - H) / 2 / L
            else { zr += ar[j]; if (k == 0) na += P(A - lag) + P(A) - P(lag); }
        }
    }
    if (OP != QUPS_APD_DMAS && warp == 0 && lagmask[0]) { // ismember(H, lags) with a zero lag: the diagonal, each element once
        R d = 0, pw = 0;
        for (uint32_t n = 0; n < A; ++n) { const V v = tile[n * 32 + lane]; const R q = v.x * v.x + v.y * v.y; d += (OP == QUPS_APD_SLSC_AVERAGE) ? (q > 0 ? R(1) : R(0)) : q; pw += q; }
        if (OP == QUPS_APD_SLSC_AVERAGE) zr += d / (R)A / R(2) / (R)nlags;
        else { zr += d; na += pw; }
    }
    red[warp][0][lane] = zr; red[warp][1][lane] = zi; red[warp][2][lane] = na;
    __syncthreads();
    if (warp == 0 && valid) {
        zr = zi = na = 0;
        for (int w = 0; w < 8; ++w) { zr += red[w][0][lane]; zi += red[w][1][lane]; na += red[w][2][lane]; }
        V o;
        if (OP == QUPS_APD_DMAS) {
            const R mag = sqrt(sqrt(zr * zr + zi * zi)); // sqrt(abs(b))
            const R ph = atan2_(zi, zr);
            R sn, cs;
            sincos(ph, &sn, &cs);
            o.x = mag * cs; o.y = mag * sn;
        } else if (OP == QUPS_APD_SLSC_AVERAGE) { o.x = zr; o.y = 0; }
        else { const R sc = rsqrt_(na) * rsqrt_(na); o.x = isfinite(sc) ? zr * sc : 0; o.y = 0; } // na == nb
        reinterpret_cast<V *>(out)[e] = o;
    }
}

int launch_aperture(const qups_aperture_params &p, void *out, void *out2, const void *b, const uint32_t *lags, cudaStream_t st) {
    const uint64_t total = p.C * p.S;
    if (total == 0 || p.A == 0) return 0;
    unsigned char *mask = nullptr;
    const bool need = p.op == QUPS_APD_DMAS || p.op == QUPS_APD_SLSC_AVERAGE || p.op == QUPS_APD_SLSC_ENSEMBLE;
    uint32_t nl = 0;
    if (need) {
        unsigned char *h = (unsigned char *)calloc(p.A + 1, 1);
        if (!h) return -4;
        for (uint32_t k = 0; k < p.nlags; ++k)
            if (lags[k] < p.A && (lags[k] >= 1 || p.op != QUPS_APD_DMAS) && !h[lags[k]]) { h[lags[k]] = 1; ++nl; } // dmas: intersect(1:N-1, L)
        cudaError_t e = ws_alloc((void **)&mask, p.A + 1, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(mask, h, p.A + 1, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st); // h is freed below
        free(h);
        if (e != cudaSuccess) { if (mask) ws_free(mask, st); return (int)e; }
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (total + 255) / 256, cap = (uint64_t)sms * 32;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    // pair-sum operators: shared-memory tile of 32 columns x the whole aperture when it fits (pcf measured slower on the tile —
    // 2.47 vs 1.7 ms on the C2 cube: its two atan2 per element hide behind the stream — and stays on the streaming kernel)
    const size_t esz = p.dtype == QUPS_F64 ? 16 : 8;
    const size_t tile_smem = (size_t)(p.A + 12) * 32 * esz + (p.op == QUPS_APD_SLSC_ENSEMBLE || p.op == QUPS_APD_PCF ? (size_t)(p.A + 1) * 32 * (esz / 2) : 0);
    if (p.op != QUPS_APD_COHFAC && p.op != QUPS_APD_PCF && tile_smem <= 200 * 1024 && p.A < 16384 && !getenv("QUPS_B200_APERTURE_SIMPLE")) {
        const uint64_t blocks = (total + 31) / 32;
        if (blocks <= 0x7fffffffull) {
            cudaError_t e2 = cudaSuccess;
#define QUPS_APD_LAUNCH(R_, OP_)                                                                                              \
    {                                                                                                                         \
        auto kfn = aperture_tile_kernel<R_, OP_>;                                                                             \
        e2 = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tile_smem > 48 * 1024 ? tile_smem : 48 * 1024)); \
        if (e2 == cudaSuccess) kfn<<<(unsigned)blocks, 256, tile_smem, st>>>(out, out2, b, mask, p.C, (uint32_t)p.A, p.S, p.nlags, (R_)p.gamma); \
    }
            if (p.dtype == QUPS_F64) {
                if (p.op == QUPS_APD_DMAS) QUPS_APD_LAUNCH(double, QUPS_APD_DMAS)
                else if (p.op == QUPS_APD_PCF) QUPS_APD_LAUNCH(double, QUPS_APD_PCF)
                else if (p.op == QUPS_APD_SLSC_AVERAGE) QUPS_APD_LAUNCH(double, QUPS_APD_SLSC_AVERAGE)
                else QUPS_APD_LAUNCH(double, QUPS_APD_SLSC_ENSEMBLE)
            } else {
                if (p.op == QUPS_APD_DMAS) QUPS_APD_LAUNCH(float, QUPS_APD_DMAS)
                else if (p.op == QUPS_APD_PCF) QUPS_APD_LAUNCH(float, QUPS_APD_PCF)
                else if (p.op == QUPS_APD_SLSC_AVERAGE) QUPS_APD_LAUNCH(float, QUPS_APD_SLSC_AVERAGE)
                else QUPS_APD_LAUNCH(float, QUPS_APD_SLSC_ENSEMBLE)
            }
#undef QUPS_APD_LAUNCH
            count_launch(1);
            if (e2 == cudaSuccess) e2 = cudaGetLastError();
            if (mask) ws_free(mask, st);
            return (int)e2;
        }
    }
    // slsc "average" normalises by the number of lags the caller asked for (L = numel(lags), kern/slsc.m)
    if (p.dtype == QUPS_F64) aperture_kernel<double><<<grid, 256, 0, st>>>(p.op, out, out2, b, mask, p.C, p.A, p.S, p.nlags, p.gamma);
    else aperture_kernel<float><<<grid, 256, 0, st>>>(p.op, out, out2, b, mask, p.C, p.A, p.S, p.nlags, (float)p.gamma);
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (mask) ws_free(mask, st);
    (void)nl;
    return (int)e;
}

} // namespace qups
