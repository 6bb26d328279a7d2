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
#include <stdlib.h>
#include "common.cuh"
#include "das_args.cuh"
#include "other_kernels.cuh"

namespace qups {
static const char *g_last_ws2 = "none";
const char *last_ws2_kernel_name() { return g_last_ws2; }

struct Ws2Dev {
    uint64_t T;
    int D, interp, w_real, has_t2;
    uint64_t sizes[8];
    uint64_t sw[8], sy[8], s1[8], s2[8], sx[8];
    int kept[8], nkept, summed[8], nsummed;
    uint64_t nout, nsum;
    // the summed dims compacted (first summed dim first): extents and strides at STATIC positions, so the odometer of the
    // term loop below reads them straight from the constant bank
    uint64_t zn[8], zw[8], z1[8], z2[8], zx[8];
};

// table / real-weight element -> R  (the reference's half kernels take half delay tables, src/interpd.cu:451-458)
template <typename R, typename TT> __device__ __forceinline__ R ldr(const TT *p, uint64_t i) { return (R)__ldg(p + i); }
template <> __device__ __forceinline__ float ldr<float, __half>(const __half *p, uint64_t i) { return __half2float(__ldg(p + i)); }

template <typename DIN, typename DOUT, typename R, typename TT>
__global__ void __launch_bounds__(128) wsinterpd2_kernel(const Ws2Dev p, DOUT *y, const void *w, const DIN *x, const TT *t1,
                                                         const TT *t2, R omega) {
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
    // the summed sub-indices advance like an odometer (first summed dim fastest — the same term order as a div / mod decode of
    // the linear term index, which cost two 64-bit divisions per summed dim and term)
    uint64_t kw = bw, k1 = b1, k2 = b2, kx = bx;
    uint64_t cnt[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cnt[q] = 0;
    for (uint64_t e = 0; e < p.nsum; ++e) {
        if (e != 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q >= p.nsummed) break;
                kw += p.zw[q]; k1 += p.z1[q]; k2 += p.z2[q]; kx += p.zx[q];
                if (++cnt[q] < p.zn[q]) break;
                cnt[q] = 0;
                kw -= p.zn[q] * p.zw[q]; k1 -= p.zn[q] * p.z1[q]; k2 -= p.zn[q] * p.z2[q]; kx -= p.zn[q] * p.zx[q];
            }
        }
        R t = ldr<R, TT>(t1, k1);
        if (p.has_t2) t = add_rn(t, ldr<R, TT>(t2, k2));
        const R xq = add_rn(R(1), t);
        const cplx<R> v = interp1<DIN>(x + kx * p.T, (long)p.T, xq, p.interp);
        cplx<R> a;
        if (p.w_real) a = {ldr<R, TT>(reinterpret_cast<const TT *>(w), kw), R(0)};
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
        else {
            d.zn[d.nsummed] = p.sizes[k]; d.zw[d.nsummed] = d.sw[k]; d.z1[d.nsummed] = d.s1[k]; d.z2[d.nsummed] = d.s2[k]; d.zx[d.nsummed] = d.sx[k];
            d.summed[d.nsummed++] = k; d.nsum *= p.sizes[k];
        }
    }
    // ---- fp16 call in the canonical look-up-table form: widen data / tables / weight once (exact) and take the staged fp32 path
    if (p.dtype == QUPS_F16 && t2 != nullptr && p.omega == 0.0 && p.interp >= 0 && p.interp <= 2 && !getenv("QUPS_B200_WS2_GENERIC")) {
        uint64_t n1 = 1, n2 = 1, nx = 1, ny = 1, nw = 1;
        for (int k = 0; k < d.D; ++k) {
            if (p.sizes[k] <= 1) continue;
            if (d.s1[k]) n1 *= p.sizes[k];
            if (d.s2[k]) n2 *= p.sizes[k];
            if (d.sx[k]) nx *= p.sizes[k];
            if (d.sy[k]) ny *= p.sizes[k];
            if (d.sw[k]) nw *= p.sizes[k];
        }
        if (nw == 1 && n1 > 1 && n2 > 1 && nx > 1) {   // (only worth it for the dense-table form; the F32 branch below re-checks the pattern)
            float *f1 = nullptr, *f2 = nullptr, *fw = nullptr;
            float2 *fx = nullptr, *fy = nullptr;
            cudaError_t e = ws_alloc((void **)&f1, sizeof(float) * n1, st);
            if (e == cudaSuccess) e = ws_alloc((void **)&f2, sizeof(float) * n2, st);
            if (e == cudaSuccess) e = ws_alloc((void **)&fx, sizeof(float2) * nx * p.T, st);
            if (e == cudaSuccess) e = ws_alloc((void **)&fw, sizeof(float) * 2, st);
            if (e == cudaSuccess && !p.y_f32) e = ws_alloc((void **)&fy, sizeof(float2) * ny, st);
            int rc = e == cudaSuccess ? 0 : -4;
            if (rc == 0) rc = launch_half_to_float(f1, (const __half *)t1, n1, st);
            if (rc == 0) rc = launch_half_to_float(f2, (const __half *)t2, n2, st);
            if (rc == 0) rc = launch_half_to_float(fw, (const __half *)w, p.w_real ? 1 : 2, st);
            if (rc == 0) rc = launch_half2_to_float2(fx, (const __half2 *)x, nx * p.T, st);
            if (rc == 0) {
                qups_ws2_params q = p;
                q.dtype = QUPS_F32; q.y_f32 = 0;
                rc = launch_wsinterpd2(q, p.y_f32 ? y : (void *)fy, fw, fx, f1, f2, st);
            }
            if (rc == 0 && !p.y_f32) rc = launch_float2_to_half2((__half2 *)y, fy, ny, st);
            for (void *q : {(void *)f1, (void *)f2, (void *)fx, (void *)fw, (void *)fy}) if (q) ws_free(q, st);
            return rc;
        }
    }
    // ---- canonical look-up-table delay-and-sum (bfDAS -> bfDASLUT -> sample2sep, src/UltrasoundSystem.m:4640-4656,
    // src/ChannelData.m:1428-1445): y(i) = w * sum_n sum_m interp1(x(:,n,m), 1 + t_a(i,m) + t_b(i,n)) with dense I x M and I x N
    // tables, fp32, no phasor.  Same data flow as DAS with the path lengths read from the tables => the staged kernel
    // (das_tiled.cu, LUT mode): windows of the traces staged in shared memory, tables read coalesced, no per-term index decode.
    if (p.dtype == QUPS_F32 && t2 != nullptr && p.omega == 0.0 && p.interp >= 0 && p.interp <= 2 && !getenv("QUPS_B200_WS2_GENERIC")) {
        int pix[8], npix = 0, dA = -1, dB = -1;   // dA: summed dim indexed by t1, dB: summed dim indexed by t2
        bool ok = true, wscalar = true;
        uint64_t I = 1;
        for (int k = 0; k < d.D && ok; ++k) {
            if (p.sizes[k] == 1) continue;
            if (d.sw[k] != 0) wscalar = false;
            if (d.sy[k] != 0) {   // kept dim = pixel dim: contiguous in y and in both tables, not a dim of x
                ok = d.sy[k] == I && d.s1[k] == I && d.s2[k] == I && d.sx[k] == 0 && npix < 3;
                pix[npix++] = k;
                I *= p.sizes[k];
            } else if (d.s1[k] != 0 && d.s2[k] == 0 && d.sx[k] != 0 && dA < 0) dA = k;
            else if (d.s2[k] != 0 && d.s1[k] == 0 && d.sx[k] != 0 && dB < 0) dB = k;
            else ok = false;
        }
        // (the pixel dims come first among the non-singleton dims only if the tables are I x ... : check the table strides)
        ok = ok && wscalar && npix >= 1 && dA >= 0 && dB >= 0 && d.s1[dA] == I && d.s2[dB] == I;
        if (ok) {
            const uint64_t nA = p.sizes[dA], nB = p.sizes[dB];
            int inner = -1;   // the summed dim whose traces are adjacent in x
            if (d.sx[dA] == 1 && d.sx[dB] == nA) inner = 0;
            else if (d.sx[dB] == 1 && d.sx[dA] == nB) inner = 1;
            if (inner >= 0) {
                DasArgs<float> a{};
                a.I1 = p.sizes[pix[0]]; a.I2 = npix > 1 ? p.sizes[pix[1]] : 1; a.I3 = npix > 2 ? p.sizes[pix[2]] : 1;
                a.I = I; a.T = p.T;
                a.N = inner == 0 ? nA : nB; a.M = inner == 0 ? nB : nA;
                a.S = 0; a.interp = p.interp; a.keep_rx = a.keep_tx = 0; a.tpose = 0; a.VS = 1; a.DV = 1; a.apod_real = 1; a.accumulate = 0;
                a.fs = 1.0f;
                a.Pi = a.Pr = a.Pv4 = a.Nv = a.cinv = nullptr; a.apod = nullptr;
                a.x = x; a.y = y;
                for (int q = 0; q < 6; ++q) a.cstride[q] = 0;
                for (int q = 0; q < MAX_APOD; ++q) for (int r = 0; r < 6; ++r) a.astride[q][r] = 0;
                a.pitch_hint[0] = a.pitch_hint[1] = a.c_hint = 0.0;
                a.fused = 0; a.fa = FusedApod{};
                a.lut_tn = (const float *)(inner == 0 ? t1 : t2);
                a.lut_tm = (const float *)(inner == 0 ? t2 : t1);
                a.lut_w = (const float *)w; a.lut_wcplx = p.w_real ? 0 : 1;
                // the staged kernel's own envelope (T range, smem for N + M, alignment of x); Pv4 is unused in LUT mode
                a.Pv4 = reinterpret_cast<const float *>(uintptr_t(16));
                const TiledPlan plan = das_tiled_plan(a, 0, 0);
                a.Pv4 = nullptr;
                if (plan.eligible) {
                    g_last_ws2 = "ws2_tiled";
                    return launch_das_tiled(a, st);
                }
            }
        }
    }
    g_last_ws2 = "wsinterpd2";
    const unsigned grid = (unsigned)((d.nout + 127) / 128);
    if (p.dtype == QUPS_F32)
        wsinterpd2_kernel<float2, float2, float, float><<<grid, 128, 0, st>>>(d, (float2 *)y, w, (const float2 *)x, (const float *)t1, (const float *)t2, (float)p.omega);
    else if (p.dtype == QUPS_F64)
        wsinterpd2_kernel<double2, double2, double, double><<<grid, 128, 0, st>>>(d, (double2 *)y, w, (const double2 *)x, (const double *)t1, (const double *)t2, p.omega);
    else if (p.dtype == QUPS_F16) {
        // wsinterpd2h / wsinterpdh (src/interpd.cu:422-429,451-458): half2 data and weights, HALF delay tables; sample positions,
        // interpolation and the sum in fp32 (the reference computes in half2), output half2 or float2 (y_f32)
        if (p.y_f32)
            wsinterpd2_kernel<__half2, float2, float, __half><<<grid, 128, 0, st>>>(d, (float2 *)y, w, (const __half2 *)x, (const __half *)t1, (const __half *)t2, (float)p.omega);
        else
            wsinterpd2_kernel<__half2, __half2, float, __half><<<grid, 128, 0, st>>>(d, (__half2 *)y, w, (const __half2 *)x, (const __half *)t1, (const __half *)t2, (float)p.omega);
    } else
        return -3;
    count_launch(1);
    return (int)cudaGetLastError();
}

} // namespace qups
