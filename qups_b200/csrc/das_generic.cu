// das_generic.cu — the complete (every option) DAS kernel family for sm_100a.
//
// Replaces src/bf.cu:49-172 (DAS_temp / DAS / DASf / DASh) and :209-298
// (delays) of the reference for EVERY argument combination of
// kern/das_spec.m: fun = DAS | SYN | MUL | BF | delays, nearest | linear |
// cubic | lanczos3, S broadcast apodization arrays, sound-speed maps,
// transposed data, fp32 / fp16 / fp64.  The hot configuration (fp32, both
// apertures summed) is served by the staged kernel in das_tiled.cu;
// this file is the general path and the semantics reference on the GPU:
// it performs the same individually rounded operations, in the same order, as
// the CPU branch (kern/das_spec.m:391-561), so its fp32 output is bit-exact
// against oracle/qups_oracle.c.
//
// Parallelisation: one thread per OUTPUT element (pixel x kept aperture
// element) — no atomics, deterministic.  Compiled with -fmad=false.
#include "das_args.cuh"

namespace qups {

template <typename R> __device__ __forceinline__ uint64_t bidx(const uint64_t *st, uint64_t i1, uint64_t i2, uint64_t i3, uint64_t n, uint64_t m) {
    return st[5] + i1 * st[0] + i2 * st[1] + i3 * st[2] + n * st[3] + m * st[4];
}

template <typename DA, typename R>
__device__ __forceinline__ cplx<R> load_apod(const void *base, int apod_real, uint64_t ix) {
    if (apod_real) {
        if constexpr (sizeof(DA) == 4) // half2 data -> half weights
            return {(R)__half2float(__ldg(reinterpret_cast<const __half *>(base) + ix)), R(0)};
        else
            return {__ldg(reinterpret_cast<const R *>(base) + ix), R(0)};
    }
    return data_traits<DA>::load(reinterpret_cast<const DA *>(base), ix);
}

// one (pixel, n, m) term: apodized sample, CPU-branch operation order
template <typename DIN, typename DA, typename R, bool BF>
__device__ __forceinline__ cplx<R> das_term(const DasArgs<R> &a, R px, R py, R pz, uint64_t i1, uint64_t i2, uint64_t i3,
                                            uint64_t n, uint64_t m, R dv, R dr, R t0m) {
    const R ci = __ldg(a.cinv + bidx<R>(a.cstride, i1, i2, i3, n, m));
    const R xq = sample_pos(dv, dr, ci, t0m, a.fs);
    const uint64_t nm = a.tpose ? (m + n * a.M) : (n + m * a.N);
    const DIN *tr = reinterpret_cast<const DIN *>(a.x) + nm * a.T;
    cplx<R> w = {R(1), R(0)};
    if (!BF && a.S > 0) {
        // a = asn{end}; for s = 1:S-1, a = a .* asn{s}      kern/das_spec.m:473
        w = load_apod<DA, R>(a.apod, a.apod_real, bidx<R>(a.astride[a.S - 1], i1, i2, i3, n, m));
        for (int s = 0; s < a.S - 1; ++s) {
            const cplx<R> b = load_apod<DA, R>(a.apod, a.apod_real, bidx<R>(a.astride[s], i1, i2, i3, n, m));
            if (a.apod_real) w.re = mul_rn(w.re, b.re); else w = cmul(w, b);
        }
        // masks (acceptance angle, f-number ...) cut the gather entirely, as src/bf.cu:121-126
        if (w.re == R(0) && w.im == R(0)) return {R(0), R(0)};
    }
    cplx<R> v = interp1<DIN>(tr, (long)a.T, xq, a.interp);
    if (!BF) {
        if (a.S > 0) {
            if (a.apod_real) { v.re = mul_rn(w.re, v.re); v.im = mul_rn(w.re, v.im); }
            else v = cmul(w, v);
        }
    } else {
        // for s = 1:S, y = y .* apod{s}                      kern/das_spec.m:558
        for (int s = 0; s < a.S; ++s) {
            const cplx<R> b = load_apod<DA, R>(a.apod, a.apod_real, bidx<R>(a.astride[s], i1, i2, i3, n, m));
            if (a.apod_real) { v.re = mul_rn(v.re, b.re); v.im = mul_rn(v.im, b.re); }
            else v = cmul(v, b);
        }
    }
    return v;
}

template <typename DIN, typename DA, typename DOUT, typename R>
__global__ void __launch_bounds__(128) das_generic_kernel(const DasArgs<R> a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.I) return;
    const uint64_t i1 = i % a.I1, i2 = (i / a.I1) % a.I2, i3 = i / (a.I1 * a.I2);
    const R px = __ldg(a.Pi + 3 * i), py = __ldg(a.Pi + 3 * i + 1), pz = __ldg(a.Pi + 3 * i + 2);
    const uint64_t On = a.keep_rx ? a.N : 1, Om = a.keep_tx ? a.M : 1;
    DOUT *y = reinterpret_cast<DOUT *>(a.y);
    const bool VS = a.VS, DV = a.DV;

    for (uint64_t o = blockIdx.y; o < On * Om; o += gridDim.y) {
        const uint64_t on = o % On, om = o / On;
        cplx<R> acc = {R(0), R(0)};
        if (a.keep_rx && a.keep_tx) { // BF  kern/das_spec.m:543-559
            const R *pv = a.Pv4 + 4 * om, *nv = a.Nv + 3 * om, *pr = a.Pr + 3 * on;
            const R dv = tx_dist(px, py, pz, pv[0], pv[1], pv[2], nv[0], nv[1], nv[2], VS, DV);
            const R dr = rx_dist(px, py, pz, pr[0], pr[1], pr[2]);
            acc = das_term<DIN, DA, R, true>(a, px, py, pz, i1, i2, i3, on, om, dv, dr, pv[3]);
        } else if (a.keep_rx) { // SYN: sum over transmits  :513-542
            const R *pr = a.Pr + 3 * on;
            const R dr = rx_dist(px, py, pz, pr[0], pr[1], pr[2]);
            for (uint64_t m = 0; m < a.M; ++m) {
                const R *pv = a.Pv4 + 4 * m, *nv = a.Nv + 3 * m;
                const R dv = tx_dist(px, py, pz, pv[0], pv[1], pv[2], nv[0], nv[1], nv[2], VS, DV);
                const cplx<R> v = das_term<DIN, DA, R, false>(a, px, py, pz, i1, i2, i3, on, m, dv, dr, pv[3]);
                acc.re = add_rn(acc.re, v.re);
                acc.im = add_rn(acc.im, v.im);
            }
        } else if (a.keep_tx) { // MUL: sum over receives  :483-512
            const R *pv = a.Pv4 + 4 * om, *nv = a.Nv + 3 * om;
            const R dv = tx_dist(px, py, pz, pv[0], pv[1], pv[2], nv[0], nv[1], nv[2], VS, DV);
            for (uint64_t n = 0; n < a.N; ++n) {
                const R *pr = a.Pr + 3 * n;
                const R dr = rx_dist(px, py, pz, pr[0], pr[1], pr[2]);
                const cplx<R> v = das_term<DIN, DA, R, false>(a, px, py, pz, i1, i2, i3, n, om, dv, dr, pv[3]);
                acc.re = add_rn(acc.re, v.re);
                acc.im = add_rn(acc.im, v.im);
            }
        } else { // DAS: for m { yn = sum_n ... ; y += yn }  :462-481
            for (uint64_t m = 0; m < a.M; ++m) {
                const R *pv = a.Pv4 + 4 * m, *nv = a.Nv + 3 * m;
                const R dv = tx_dist(px, py, pz, pv[0], pv[1], pv[2], nv[0], nv[1], nv[2], VS, DV);
                const R t0m = pv[3];
                cplx<R> yn = {R(0), R(0)};
                for (uint64_t n = 0; n < a.N; ++n) {
                    const R *pr = a.Pr + 3 * n;
                    const R dr = rx_dist(px, py, pz, pr[0], pr[1], pr[2]);
                    const cplx<R> v = das_term<DIN, DA, R, false>(a, px, py, pz, i1, i2, i3, n, m, dv, dr, t0m);
                    yn.re = add_rn(yn.re, v.re);
                    yn.im = add_rn(yn.im, v.im);
                }
                acc.re = add_rn(acc.re, yn.re);
                acc.im = add_rn(acc.im, yn.im);
            }
        }
        if (a.accumulate) { // y += result (transmit-chunked callers); not for BF (nothing is summed there)
            const auto old = data_traits<DOUT>::load(y, i + a.I * o);
            acc.re = add_rn((R)old.re, acc.re);
            acc.im = add_rn((R)old.im, acc.im);
        }
        data_traits<DOUT>::store(y, i + a.I * o, {(typename data_traits<DOUT>::real)acc.re,
                                                  (typename data_traits<DOUT>::real)acc.im});
    }
}

template <typename DIN, typename DA, typename DOUT, typename R>
int launch_das_generic(const DasArgs<R> &a, cudaStream_t st) {
    if (a.I == 0) return 0;
    const uint64_t On = a.keep_rx ? a.N : 1, Om = a.keep_tx ? a.M : 1;
    dim3 block(128), grid((unsigned)((a.I + 127) / 128), (unsigned)((On * Om < 65535) ? On * Om : 65535));
    das_generic_kernel<DIN, DA, DOUT, R><<<grid, block, 0, st>>>(a);
    count_launch();
    return (int)cudaGetLastError();
}

template int launch_das_generic<float2, float2, float2, float>(const DasArgs<float> &, cudaStream_t);
template int launch_das_generic<__half2, __half2, __half2, float>(const DasArgs<float> &, cudaStream_t);
template int launch_das_generic<__half2, __half2, float2, float>(const DasArgs<float> &, cudaStream_t);
template int launch_das_generic<float2, __half2, __half2, float>(const DasArgs<float> &, cudaStream_t);
template int launch_das_generic<float2, __half2, float2, float>(const DasArgs<float> &, cudaStream_t);
template int launch_das_generic<double2, double2, double2, double>(const DasArgs<double> &, cudaStream_t);

// ---- delays: tau(i,n,m) = cinv .* (dv + dr)        kern/das_spec.m:448-449 ----
template <typename R> __global__ void __launch_bounds__(128) delays_kernel(const DasArgs<R> a, R *tau) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.I) return;
    const uint64_t i1 = i % a.I1, i2 = (i / a.I1) % a.I2, i3 = i / (a.I1 * a.I2);
    const R px = __ldg(a.Pi + 3 * i), py = __ldg(a.Pi + 3 * i + 1), pz = __ldg(a.Pi + 3 * i + 2);
    for (uint64_t o = blockIdx.y; o < a.N * a.M; o += gridDim.y) {
        const uint64_t n = o % a.N, m = o / a.N;
        const R *pv = a.Pv4 + 4 * m, *nv = a.Nv + 3 * m, *pr = a.Pr + 3 * n;
        const R dv = tx_dist(px, py, pz, pv[0], pv[1], pv[2], nv[0], nv[1], nv[2], (bool)a.VS, (bool)a.DV);
        const R dr = rx_dist(px, py, pz, pr[0], pr[1], pr[2]);
        const R ci = __ldg(a.cinv + bidx<R>(a.cstride, i1, i2, i3, n, m));
        tau[i + a.I * o] = mul_rn(ci, add_rn(dv, dr));
    }
}
template <typename R> int launch_delays(const DasArgs<R> &a, R *tau, cudaStream_t st) {
    if (a.I == 0 || a.N * a.M == 0) return 0;
    const uint64_t O = a.N * a.M;
    dim3 block(128), grid((unsigned)((a.I + 127) / 128), (unsigned)(O < 65535 ? O : 65535));
    delays_kernel<R><<<grid, block, 0, st>>>(a, tau);
    count_launch();
    return (int)cudaGetLastError();
}
template int launch_delays<float>(const DasArgs<float> &, float *, cudaStream_t);
template int launch_delays<double>(const DasArgs<double> &, double *, cudaStream_t);

// ---- modulation pre-pass: x .* exp(2i*pi*fmod.*(t0 + (0:T-1)'/fs))   kern/das_spec.m:413-417 ----
template <typename DIN, typename DOUT, typename R>
__global__ void __launch_bounds__(256) modulate_kernel(DOUT *xo, const DIN *x, const R *t0, int t0_stride, uint64_t T,
                                                       uint64_t N, uint64_t M, int tpose, R fs, R w) {
    const uint64_t tr = blockIdx.x; // trace index n + N*m (or m + M*n)
    const uint64_t m = tpose ? (tr % M) : (tr / N);
    const R t0m = t0[m * (uint64_t)t0_stride];
    for (uint64_t j = threadIdx.x; j < T; j += blockDim.x) {
        R tj = div_rn((R)j, fs);
        tj = add_rn(t0m, tj);
        const R th = mul_rn(w, tj);
        R s, c;
        if constexpr (sizeof(R) == 4) sincosf(th, &s, &c); else sincos(th, &s, &c);
        const cplx<R> v = data_traits<DIN>::load(x, tr * T + j);
        cplx<R> o;
        o.re = sub_rn(mul_rn(v.re, c), mul_rn(v.im, s));
        o.im = add_rn(mul_rn(v.re, s), mul_rn(v.im, c));
        data_traits<DOUT>::store(xo, tr * T + j, {(typename data_traits<DOUT>::real)o.re,
                                                  (typename data_traits<DOUT>::real)o.im});
    }
}
template <typename DIN, typename DOUT, typename R>
int launch_modulate(DOUT *xo, const DIN *x, const R *t0, int t0_stride, uint64_t T, uint64_t N, uint64_t M, int tpose,
                    R fs, double fmod, cudaStream_t st) {
    if (T * N * M == 0) return 0;
    if (N * M > 2147483647ull) return (int)cudaErrorInvalidValue;
    modulate_kernel<DIN, DOUT, R><<<(unsigned)(N * M), 256, 0, st>>>(xo, x, t0, t0_stride, T, N, M, tpose, fs,
                                                                     (R)(2.0 * 3.14159265358979323846 * fmod));
    count_launch();
    return (int)cudaGetLastError();
}
template int launch_modulate<float2, float2, float>(float2 *, const float2 *, const float *, int, uint64_t, uint64_t,
                                                    uint64_t, int, float, double, cudaStream_t);
template int launch_modulate<__half2, float2, float>(float2 *, const __half2 *, const float *, int, uint64_t, uint64_t,
                                                     uint64_t, int, float, double, cudaStream_t);
template int launch_modulate<__half2, __half2, float>(__half2 *, const __half2 *, const float *, int, uint64_t,
                                                      uint64_t, uint64_t, int, float, double, cudaStream_t);
template int launch_modulate<double2, double2, double>(double2 *, const double2 *, const double *, int, uint64_t,
                                                       uint64_t, uint64_t, int, double, double, cudaStream_t);

// ---- half2 <-> float2 conversion (fp16 data through the staged fp32 kernel) ----
__global__ void __launch_bounds__(256) h2f_kernel(float2 *dst, const __half2 *src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = __half22float2(__ldg(src + i));
}
__global__ void __launch_bounds__(256) f2h_kernel(__half2 *dst, const float2 *src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 v = __ldg(src + i);
        dst[i] = __floats2half2_rn(v.x, v.y);
    }
}
__global__ void __launch_bounds__(256) h2f1_kernel(float *dst, const __half *src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = __half2float(__ldg(src + i));
}
__global__ void __launch_bounds__(256) f2h1_kernel(__half *dst, const float *src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = __float2half_rn(__ldg(src + i));
}
static unsigned conv_grid(uint64_t n) {
    const uint64_t g = (n + 255) / 256;
    return (unsigned)(g < 148ull * 16 ? (g ? g : 1) : 148ull * 16);
}
int launch_half2_to_float2(float2 *dst, const __half2 *src, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    h2f_kernel<<<conv_grid(n), 256, 0, st>>>(dst, src, n);
    count_launch();
    return (int)cudaGetLastError();
}
int launch_half_to_float(float *dst, const __half *src, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    h2f1_kernel<<<conv_grid(n), 256, 0, st>>>(dst, src, n);
    count_launch();
    return (int)cudaGetLastError();
}
int launch_float_to_half(__half *dst, const float *src, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    f2h1_kernel<<<conv_grid(n), 256, 0, st>>>(dst, src, n);
    count_launch();
    return (int)cudaGetLastError();
}
int launch_float2_to_half2(__half2 *dst, const float2 *src, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    f2h_kernel<<<conv_grid(n), 256, 0, st>>>(dst, src, n);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace qups
