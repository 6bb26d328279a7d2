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

#include "common.cuh"
#include "other_kernels.cuh"

namespace qups {
void count_launch(uint64_t n);
cudaError_t ws_alloc(void **p, size_t bytes, cudaStream_t st);
cudaError_t ws_free(void *p, cudaStream_t st);

template <typename R> struct c2 { using type = float2; };
template <> struct c2<double> { using type = double2; };

template <typename R> __device__ __forceinline__ R rsqrt_(R x) { return R(1) / sqrt(x); }
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
