// wsinterpd2.cu — weighted-sum interpolation with separable delay tables.
//
// Replaces src/interpd.cu:292-396 (wsinterpd_temp / wsinterpd2_temp) — the
// kernel behind bfDAS / bfDASLUT (src/UltrasoundSystem.m:4651 ->
// src/ChannelData.m:1445 -> kern/wsinterpd2.m) and focusTx
// (src/UltrasoundSystem.m:3498).
//
//   y(l) = sum_{dims with ystride == 0}  exp(1i*omega*t) * w(k) * interp1(x(:,v), 1 + t, interp, 0),  t = t1(r)+t2(u)
//
// The reference walks every (i,n,f) element and reduces with global atomicAdd
// (src/interpd.cu:339,393) — non-deterministic.  Here one thread owns one
// OUTPUT element and walks the summed dims itself (first summed dim fastest),
// so the result is deterministic.  Compiled with -fmad=false; the per-term
// operation order mirrors kern/wsinterpd2.m:290.
#include "common.cuh"
#include "other_kernels.cuh"

namespace qups {
void count_launch(uint64_t n);

struct Ws2Dev {
    uint64_t T;
    int D, interp, w_real, has_t2;
    uint64_t sizes[8];
    uint64_t sw[8], sy[8], s1[8], s2[8], sx[8];
    int kept[8], nkept, summed[8], nsummed;
    uint64_t nout, nsum;
};

template <typename DIN, typename DOUT, typename R>
__global__ void __launch_bounds__(128) wsinterpd2_kernel(const Ws2Dev p, DOUT *y, const void *w, const DIN *x, const R *t1,
                                                         const R *t2, R omega) {
    const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= p.nout) return;
    // decode the kept-dim sub-indices (column-major over kept dims)
    uint64_t bw = 0, by = 0, b1 = 0, b2 = 0, bx = 0, rem = o;
    for (int q = 0; q < p.nkept; ++q) {
        const int d = p.kept[q];
        const uint64_t j = rem % p.sizes[d];
        rem /= p.sizes[d];
        bw += j * p.sw[d]; by += j * p.sy[d]; b1 += j * p.s1[d]; b2 += j * p.s2[d]; bx += j * p.sx[d];
    }
    cplx<R> acc = {R(0), R(0)};
    for (uint64_t e = 0; e < p.nsum; ++e) {
        uint64_t kw = bw, k1 = b1, k2 = b2, kx = bx, r2 = e;
        for (int q = 0; q < p.nsummed; ++q) {
            const int d = p.summed[q];
            const uint64_t j = r2 % p.sizes[d];
            r2 /= p.sizes[d];
            kw += j * p.sw[d]; k1 += j * p.s1[d]; k2 += j * p.s2[d]; kx += j * p.sx[d];
        }
        R t = __ldg(t1 + k1);
        if (p.has_t2) t = add_rn(t, __ldg(t2 + k2));
        const R xq = add_rn(R(1), t);
        const cplx<R> v = interp1<DIN>(x + kx * p.T, (long)p.T, xq, p.interp);
        cplx<R> a;
        if (p.w_real) a = {__ldg(reinterpret_cast<const R *>(w) + kw), R(0)};
        else a = data_traits<DIN>::load(reinterpret_cast<const DIN *>(w), kw);
        if (omega != R(0)) { // exp(omega .* tau) .* amp
            const R th = mul_rn(omega, t);
            R s, c;
            if constexpr (sizeof(R) == 4) sincosf(th, &s, &c); else sincos(th, &s, &c);
            a = cmul(cplx<R>{c, s}, a);
        }
        const cplx<R> z = cmul(a, v);
        if (z.re == z.re && z.im == z.im) { // sum(..., 'omitnan')
            acc.re = add_rn(acc.re, z.re);
            acc.im = add_rn(acc.im, z.im);
        }
    }
    data_traits<DOUT>::store(y, by, {(typename data_traits<DOUT>::real)acc.re, (typename data_traits<DOUT>::real)acc.im});
}

int launch_wsinterpd2(const qups_ws2_params &p, void *y, const void *w, const void *x, const void *t1, const void *t2,
                      cudaStream_t st) {
    Ws2Dev d{};
    d.T = p.T; d.D = (int)p.D; d.interp = p.interp; d.w_real = p.w_real; d.has_t2 = t2 != nullptr;
    d.nout = 1; d.nsum = 1;
    for (int k = 0; k < d.D; ++k) {
        d.sizes[k] = p.sizes[k];
        d.sw[k] = p.dstride[0 + 5 * k]; d.sy[k] = p.dstride[1 + 5 * k]; d.s1[k] = p.dstride[2 + 5 * k];
        d.s2[k] = p.dstride[3 + 5 * k]; d.sx[k] = p.dstride[4 + 5 * k];
        if (p.sizes[k] == 0) return 0; // empty
        if (p.sizes[k] == 1) continue;
        if (d.sy[k] != 0) { d.kept[d.nkept++] = k; d.nout *= p.sizes[k]; }
        else { d.summed[d.nsummed++] = k; d.nsum *= p.sizes[k]; }
    }
    const unsigned grid = (unsigned)((d.nout + 127) / 128);
    if (p.dtype == QUPS_F32)
        wsinterpd2_kernel<float2, float2, float><<<grid, 128, 0, st>>>(d, (float2 *)y, w, (const float2 *)x, (const float *)t1, (const float *)t2, (float)p.omega);
    else if (p.dtype == QUPS_F64)
        wsinterpd2_kernel<double2, double2, double><<<grid, 128, 0, st>>>(d, (double2 *)y, w, (const double2 *)x, (const double *)t1, (const double *)t2, p.omega);
    else
        return -3;
    count_launch(1);
    return (int)cudaGetLastError();
}

} // namespace qups
