// xcorr.cu — pair-wise windowed zero-normalised cross-correlation (replaces the whole-array MATLAB
// expressions of kern/pwznxcorr.m:142-266, native branch: integer lags, U = 1, multi = false).
//
// The reference materialises ~8 temporaries of the size of the (padded) data per lag (shifted copy, its moving sum, the
// debiased copy, the product, its moving sum, the power, its moving sum, the quotient).  Here ONE CTA owns a tile of
// kTile output samples of one (channel pair, frame): it stages the two traces (tile + two window halos) in shared memory
// once, derives the debiased reference trace and its windowed power once, and then walks the lags — every intermediate
// lives in shared memory, the data are read once per tile and the output is written once.
//
// Arithmetic follows the oracle (oracle/xcorr_np.py) statement by statement in the data precision; the moving sums run in
// the oracle's tap order (k = 0 .. W-1), so the result differs from it only by FMA contraction.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "other_kernels.cuh"

namespace qups {
void count_launch(uint64_t n);

namespace {
constexpr int kTile = 512;
constexpr int kThreads = 256;

template <typename R> struct cx { using type = float2; };
template <> struct cx<double> { using type = double2; };

struct XcorrArgs {
    const void *x, *x0, *w;
    void *y;
    const int32_t *lags;
    uint32_t T, Tp, N, Nout, F, L, W;
    int ref, zero, norm, x_complex;
    uint32_t S, n0a, n0b;       // neighbor stride; centre channel(s) (0-based)
    uint32_t x0N, x0F;          // extents of x0 along channels / frames (1 = broadcast)
};

template <typename R, typename V> __device__ __forceinline__ V load_c(const void *p, uint64_t i, int is_complex) {
    if (is_complex) return reinterpret_cast<const V *>(p)[i];
    V v; v.x = reinterpret_cast<const R *>(p)[i]; v.y = R(0); return v;
}

// out[i] (i in [0, n)) = sum_k w[k] * in[i + h + c - k], in/out indexed relative to their own origins: in starts h samples
// before out.  `in` already holds zeros outside the valid time range (MATLAB 'same': zero padding).
template <typename R, typename V>
__device__ __forceinline__ V msum_c(const V *in, const R *w, int i, int W, int c) {
    V a; a.x = R(0); a.y = R(0);
    const V *p = in + i + c;
    for (int k = 0; k < W; ++k) { const V v = p[-k]; a.x += w[k] * v.x; a.y += w[k] * v.y; }
    return a;
}
template <typename R> __device__ __forceinline__ R msum_r(const R *in, const R *w, int i, int W, int c) {
    R a = R(0);
    const R *p = in + i + c;
    for (int k = 0; k < W; ++k) a += w[k] * p[-k];
    return a;
}

// Three consecutive outputs at once: out[j] = sum_k w[k] * in[i0 + j + c - k], j = 0..2.  The window slides one sample per tap, so
// every tap costs ONE new input load (+ the weight) for three outputs instead of two loads per output; the taps are still added
// in the oracle's order k = 0 .. W-1.  Three, not four: lanes 3 elements apart hit distinct banks (3 is coprime to 16 and 32),
// lanes 4 apart would be a 4-way conflict and give the saving back.
template <typename R, typename V>
__device__ __forceinline__ void msum3_c(const V *in, const R *w, int i0, int W, int c, V (&out)[3]) {
    const V *p = in + i0 + c;
    V a1 = p[1], a2 = p[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) { out[j].x = R(0); out[j].y = R(0); }
    for (int k = 0; k < W; ++k) {
        const R wk = w[k];
        const V a0 = p[-k];
        out[0].x += wk * a0.x; out[0].y += wk * a0.y;
        out[1].x += wk * a1.x; out[1].y += wk * a1.y;
        out[2].x += wk * a2.x; out[2].y += wk * a2.y;
        a2 = a1; a1 = a0;
    }
}
template <typename R> __device__ __forceinline__ void msum3_r(const R *in, const R *w, int i0, int W, int c, R (&out)[3]) {
    const R *p = in + i0 + c;
    R a1 = p[1], a2 = p[2];
    out[0] = out[1] = out[2] = R(0);
    for (int k = 0; k < W; ++k) {
        const R wk = w[k], a0 = p[-k];
        out[0] += wk * a0; out[1] += wk * a1; out[2] += wk * a2;
        a2 = a1; a1 = a0;
    }
}

template <typename R>
__global__ void __launch_bounds__(kThreads) pwznxcorr_kernel(const XcorrArgs a) {
    using V = typename cx<R>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const int W = (int)a.W, h = W - 1, c = W / 2;
    const int nA = kTile + 4 * h + 4, nB = kTile + 2 * h + 4;   // + 4: the 3-wide groups read up to 2 elements past the last window
    // layout: w[W] | xa[nA] (left trace, tile -2h .. +2h) | xb[nA] (shifted right trace) | xlz[nB] | xrz[nB] | q[nB] | pl[nB] pr[nB] | xln[kTile]
    R *sw = reinterpret_cast<R *>(smem);
    V *xa = reinterpret_cast<V *>(smem + (((size_t)W * sizeof(R) + 15) & ~(size_t)15));
    V *xb = xa + nA;
    V *xlz = xb + nA;
    V *xrz = xlz + nB;
    V *q = xrz + nB;
    R *pl = reinterpret_cast<R *>(q + nB);
    R *pr = pl + nB;
    R *xln = pr + nB;
    const int tid = threadIdx.x;
    const int t0 = (int)blockIdx.x * kTile;
    const uint32_t n = blockIdx.y, f = blockIdx.z;
    const int T = (int)a.T, Tp = (int)a.Tp;
    for (int k = tid; k < W; k += kThreads) sw[k] = reinterpret_cast<const R *>(a.w)[k];
    // left trace: channel n of x, zero beyond T (the appended padding) and outside [0, Tp)
    const uint64_t xoff = ((uint64_t)f * a.N + n) * a.T;
    for (int i = tid; i < nA; i += kThreads) {
        const int s = t0 - 2 * h + i;
        V v; v.x = R(0); v.y = R(0);
        if (s >= 0 && s < T) v = load_c<R, V>(a.x, xoff + s, a.x_complex);
        xa[i] = v;
    }
    __syncthreads();
    // xlz = xl - kernfun(xl) on [t0 - h, t0 + kTile + h), zero outside [0, Tp)   (kern/pwznxcorr.m:225)
    const int nBv = kTile + 2 * h;                       // valid entries of the nB-arrays
    for (int i0 = 3 * tid; i0 < nBv; i0 += 3 * kThreads) {
        V m[3];
        if (a.zero) msum3_c<R, V>(xa, sw, i0 + h, W, c, m);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = i0 + j, s = t0 - h + i;
            V v; v.x = R(0); v.y = R(0);
            if (s >= 0 && s < Tp) {
                v = xa[i + h];
                if (a.zero) { v.x -= m[j].x; v.y -= m[j].y; }
            }
            xlz[i] = v;                                   // (i < nB: the 4 pad entries absorb the last group's overrun)
            pl[i] = v.x * v.x + v.y * v.y;
        }
    }
    __syncthreads();
    if (a.norm)
        for (int i0 = 3 * tid; i0 < kTile; i0 += 3 * kThreads) { // :226
            R m[3];
            msum3_r<R>(pl, sw, i0 + h, W, c, m);
#pragma unroll
            for (int j = 0; j < 3; ++j) if (i0 + j < kTile) xln[i0 + j] = m[j];
        }
    // right trace source
    auto right = [&](int s) -> V { // xr[s], s in [0, Tp): zero in the appended padding
        V v; v.x = R(0); v.y = R(0);
        if (s >= T) return v;
        if (a.ref == 0) return load_c<R, V>(a.x, ((uint64_t)f * a.N + (n + a.S)) * a.T + s, a.x_complex);
        if (a.ref == 1) {
            const V p0 = load_c<R, V>(a.x, ((uint64_t)f * a.N + a.n0a) * a.T + s, a.x_complex);
            if (a.n0a == a.n0b) return p0;
            const V p1 = load_c<R, V>(a.x, ((uint64_t)f * a.N + a.n0b) * a.T + s, a.x_complex);
            v.x = (p0.x + p1.x) / R(2); v.y = (p0.y + p1.y) / R(2); // mean(xr, ndim) of the two median channels (:203-204)
            return v;
        }
        const uint64_t nn = a.x0N > 1 ? n : 0, ff = a.x0F > 1 ? f : 0;
        return load_c<R, V>(a.x0, (ff * a.x0N + nn) * a.T + s, 1);
    };
    for (uint32_t li = 0; li < a.L; ++li) {
        const int lag = a.lags[li];
        __syncthreads();
        // xr_l = conj(circshift(xr, -l)) over the padded length (:175, :233), zero outside [0, Tp)
        for (int i = tid; i < nA; i += kThreads) {
            const int s = t0 - 2 * h + i;
            V v; v.x = R(0); v.y = R(0);
            if (s >= 0 && s < Tp) {
                int sl = (s + lag) % Tp;
                if (sl < 0) sl += Tp;
                v = right(sl);
                v.y = -v.y;
            }
            xb[i] = v;
        }
        __syncthreads();
        for (int i0 = 3 * tid; i0 < nBv; i0 += 3 * kThreads) { // xrz_l, the product and the power (:240-246, :251)
            V m[3];
            if (a.zero) msum3_c<R, V>(xb, sw, i0 + h, W, c, m);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i = i0 + j, s = t0 - h + i;
                V v; v.x = R(0); v.y = R(0);
                if (s >= 0 && s < Tp) {
                    v = xb[i + h];
                    if (a.zero) { v.x -= m[j].x; v.y -= m[j].y; }
                }
                xrz[i] = v;
                const V u = xlz[i];
                V p; p.x = u.x * v.x - u.y * v.y; p.y = u.x * v.y + u.y * v.x;
                q[i] = p;
                pr[i] = v.x * v.x + v.y * v.y;
            }
        }
        __syncthreads();
        for (int i0 = 3 * tid; i0 < kTile; i0 += 3 * kThreads) {
            if (t0 + i0 >= T) break;
            V yv[3];
            R xrn[3] = {R(1), R(1), R(1)};
            msum3_c<R, V>(q, sw, i0 + h, W, c, yv);
            if (a.norm) msum3_r<R>(pr, sw, i0 + h, W, c, xrn);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i = i0 + j, t = t0 + i;
                if (t >= T || i >= kTile) break;
                if (a.norm) {
                    const R r = sqrt(xln[i]) * sqrt(xrn[j]); // .* sqrt(Wn), Wn = 1 (:258)
                    yv[j].x /= r; yv[j].y /= r;
                }
                reinterpret_cast<V *>(a.y)[(((uint64_t)li * a.F + f) * a.Nout + n) * a.T + t] = yv[j];
            }
        }
    }
}
} // namespace

size_t xcorr_smem_bytes(uint32_t W, int dbl) {
    const size_t R = dbl ? 8 : 4, V = 2 * R, h = W - 1;
    const size_t nA = kTile + 4 * h + 4, nB = kTile + 2 * h + 4;
    return (((size_t)W * R + 15) & ~(size_t)15) + 2 * nA * V + 3 * nB * V + 2 * nB * R + kTile * R;
}

int launch_pwznxcorr(int dbl, void *y, const void *x, const void *x0, const void *w, const int32_t *lags_dev, uint32_t T, uint32_t P,
                     uint32_t N, uint32_t F, uint32_t L, uint32_t W, int ref, int zero, int norm, int x_complex, uint32_t S,
                     uint32_t x0N, uint32_t x0F, cudaStream_t st) {
    XcorrArgs a{};
    a.x = x; a.x0 = x0; a.w = w; a.y = y; a.lags = lags_dev;
    a.T = T; a.Tp = T + P; a.N = N; a.F = F; a.L = L; a.W = W;
    a.ref = ref; a.zero = zero; a.norm = norm; a.x_complex = x_complex; a.S = S;
    a.Nout = ref == 0 ? N - S : N;
    // centre reference: mid = (N + 1)/2 (1-based), n = unique([floor(mid), ceil(mid)])  (kern/pwznxcorr.m:199-200)
    a.n0a = (N + 1) / 2 - 1; a.n0b = (N + 2) / 2 - 1;
    a.x0N = x0N; a.x0F = x0F;
    const size_t smem = xcorr_smem_bytes(W, dbl);
    auto kern = dbl ? pwznxcorr_kernel<double> : pwznxcorr_kernel<float>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((T + kTile - 1) / kTile, a.Nout, F);
    kern<<<grid, kThreads, smem, st>>>(a);
    count_launch(1);
    return (int)cudaGetLastError();
}

} // namespace qups
