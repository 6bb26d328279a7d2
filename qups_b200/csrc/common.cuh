// common.cuh — shared device helpers for libqups_b200 (sm_100a only).
//
// Numerics contract ("canonical fp32 sequence", SURVEY.md §8c / DESIGN.md §4):
// every geometry / delay operation is one individually rounded IEEE operation
// (explicit __fadd_rn/__fmul_rn/__fsqrt_rn so the compiler can never contract
// them into FMAs), in the same order as the CPU branch of kern/das_spec.m:
//   rv = Pi - Pv ; q = rx*rx ; q += ry*ry ; q += rz*rz ; d = sqrt(q)     (:427-436)
//   tau = cinv*(dv+dr) ; tau -= t0(m) ; xq = tau*fs ; xq = 1 + xq          (:469,:477)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qups {

// ---- exactly rounded scalar ops (never contracted) -------------------------
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ float floor_(float a) { return floorf(a); }
__device__ __forceinline__ double floor_(double a) { return floor(a); }
__device__ __forceinline__ float round_(float a) { return roundf(a); }
__device__ __forceinline__ double round_(double a) { return round(a); }

// ---- complex sample types ---------------------------------------------------
template <typename R> struct cplx { R re, im; };

template <typename D> struct data_traits;
template <> struct data_traits<float2> {
    using real = float;
    static __device__ __forceinline__ cplx<float> load(const float2 *p, size_t i) {
        float2 v = __ldg(p + i);
        return {v.x, v.y};
    }
    static __device__ __forceinline__ void store(float2 *p, size_t i, cplx<float> v) { p[i] = make_float2(v.re, v.im); }
};
template <> struct data_traits<double2> {
    using real = double;
    static __device__ __forceinline__ cplx<double> load(const double2 *p, size_t i) {
        double2 v = __ldg(p + i);
        return {v.x, v.y};
    }
    static __device__ __forceinline__ void store(double2 *p, size_t i, cplx<double> v) { p[i] = make_double2(v.re, v.im); }
};
template <> struct data_traits<__half2> {
    using real = float; // geometry and accumulation in fp32
    static __device__ __forceinline__ cplx<float> load(const __half2 *p, size_t i) {
        float2 v = __half22float2(__ldg(p + i));
        return {v.x, v.y};
    }
    static __device__ __forceinline__ void store(__half2 *p, size_t i, cplx<float> v) {
        p[i] = __floats2half2_rn(v.re, v.im);
    }
};

// ---- geometry: kern/das_spec.m:427-436 --------------------------------------
template <typename R> __device__ __forceinline__ R norm3(R x, R y, R z) {
    R q = mul_rn(x, x);
    q = add_rn(q, mul_rn(y, y));
    q = add_rn(q, mul_rn(z, z));
    return sqrt_rn(q);
}
template <typename R> __device__ __forceinline__ R dot3(R ax, R ay, R az, R bx, R by, R bz) {
    R q = mul_rn(ax, bx);
    q = add_rn(q, mul_rn(ay, by));
    q = add_rn(q, mul_rn(az, bz));
    return q;
}
// transmit path length dv(i,m): virtual source (focused: signed by the normal;
// diverging: unsigned) or plane wave.   sign(0) == 0 as in MATLAB.
template <typename R>
__device__ __forceinline__ R tx_dist(R px, R py, R pz, R vx, R vy, R vz, R nx, R ny, R nz, bool VS, bool DV) {
    const R rx = sub_rn(px, vx), ry = sub_rn(py, vy), rz = sub_rn(pz, vz);
    if (VS) {
        const R d = norm3(rx, ry, rz);
        if (DV) return d;
        const R dp = dot3(rx, ry, rz, nx, ny, nz);
        const R s = (dp > R(0)) ? R(1) : ((dp < R(0)) ? R(-1) : ((dp == R(0)) ? R(0) : dp));
        return mul_rn(d, s);
    }
    return dot3(rx, ry, rz, nx, ny, nz);
}
template <typename R> __device__ __forceinline__ R rx_dist(R px, R py, R pz, R rx_, R ry_, R rz_) {
    return norm3(sub_rn(px, rx_), sub_rn(py, ry_), sub_rn(pz, rz_));
}
// 1-based fractional sample position: xq = 1 + (cinv*(dv+dr) - t0)*fs
template <typename R> __device__ __forceinline__ R sample_pos(R dv, R dr, R cinv, R t0, R fs) {
    R tau = mul_rn(cinv, add_rn(dv, dr));
    tau = sub_rn(tau, t0);
    return add_rn(R(1), mul_rn(tau, fs));
}

// ---- interp1(v, xq, method, 0) on the grid 1..T ---------------
// Full edge semantics of MATLAB's interp1 as used by the CPU branch of das_spec
// (kern/das_spec.m:477): used by the generic kernels and by the slow path of
// the tiled kernel.  Operation order mirrors oracle/oracle_body.inc so the
// generic path is bit-exact against the fp32 oracle.
template <typename D, typename R = typename data_traits<D>::real>
__device__ __forceinline__ cplx<R> tap_padded(const D *v, long T, long k1) {
    using TR = data_traits<D>;
    if (k1 >= 1 && k1 <= T) return TR::load(v, (size_t)(k1 - 1));
    long a, b, c;
    if (k1 < 1) { a = 1; b = 2; c = 3; } else { a = T; b = T - 1; c = T - 2; }
    const cplx<R> va = TR::load(v, a - 1), vb = TR::load(v, b - 1), vc = TR::load(v, c - 1);
    cplx<R> o;
    o.re = add_rn(sub_rn(mul_rn(R(3), va.re), mul_rn(R(3), vb.re)), vc.re);
    o.im = add_rn(sub_rn(mul_rn(R(3), va.im), mul_rn(R(3), vb.im)), vc.im);
    return o;
}

__device__ __forceinline__ float lanczos_w(float u) {
    if (u == 0.f) return 1.f;
    const double PI = 3.14159265358979323846;
    const double s1 = sinpi((double)u), s2 = sinpi((double)u * 0.5);
    return (float)(2.0 * s1 * s2 / (PI * PI * (double)u * (double)u));
}
__device__ __forceinline__ double lanczos_w(double u) {
    if (u == 0.0) return 1.0;
    const double PI = 3.14159265358979323846;
    return 2.0 * sinpi(u) * sinpi(u * 0.5) / (PI * PI * u * u);
}

template <typename D, typename R = typename data_traits<D>::real>
__device__ __forceinline__ cplx<R> interp1(const D *v, long T, R xq, int method) {
    using TR = data_traits<D>;
    cplx<R> o = {R(0), R(0)};
    if (!(xq >= R(1) && xq <= (R)T)) return o; // also NaN
    if (method == 0) { // nearest, half away from zero
        long k = (long)round_(xq);
        k = k < 1 ? 1 : (k > T ? T : k);
        return TR::load(v, (size_t)(k - 1));
    } else if (method == 1) { // linear
        long k = (long)floor_(xq);
        k = k > T - 1 ? T - 1 : k;
        k = k < 1 ? 1 : k;
        const R s = sub_rn(xq, (R)k);
        const cplx<R> v0 = TR::load(v, (size_t)(k - 1)), v1 = TR::load(v, (size_t)k);
        o.re = add_rn(v0.re, mul_rn(s, sub_rn(v1.re, v0.re)));
        o.im = add_rn(v0.im, mul_rn(s, sub_rn(v1.im, v0.im)));
        return o;
    } else if (method == 2) { // cubic convolution, Keys a = -1/2, padded ends
        long k = (long)floor_(xq);
        k = k > T - 1 ? T - 1 : k;
        k = k < 1 ? 1 : k;
        const R s = sub_rn(xq, (R)k);
        const R s2 = mul_rn(s, s), s3 = mul_rn(s2, s);
        const R w0 = sub_rn(add_rn(mul_rn(R(-1), s3), mul_rn(R(2), s2)), s);
        const R w1 = add_rn(sub_rn(mul_rn(R(3), s3), mul_rn(R(5), s2)), R(2));
        const R w2 = add_rn(add_rn(mul_rn(R(-3), s3), mul_rn(R(4), s2)), s);
        const R w3 = sub_rn(s3, s2);
        const cplx<R> a = tap_padded<D>(v, T, k - 1), b = tap_padded<D>(v, T, k);
        const cplx<R> c = tap_padded<D>(v, T, k + 1), d = tap_padded<D>(v, T, k + 2);
        R ar = add_rn(mul_rn(w0, a.re), mul_rn(w1, b.re));
        ar = add_rn(ar, mul_rn(w2, c.re));
        ar = add_rn(ar, mul_rn(w3, d.re));
        R ai = add_rn(mul_rn(w0, a.im), mul_rn(w1, b.im));
        ai = add_rn(ai, mul_rn(w2, c.im));
        ai = add_rn(ai, mul_rn(w3, d.im));
        o.re = mul_rn(R(0.5), ar);
        o.im = mul_rn(R(0.5), ai);
        return o;
    } else if (method == 3) { // lanczos3 (GPU-only in the reference: src/interpd.cu:133-150)
        const R tau = sub_rn(xq, R(1));
        const R kf = floor_(tau);
        const long ti = (long)kf;
        const R u = sub_rn(tau, kf);
        if (!(ti - 1 >= 0 && ti + 2 < T)) return o;
        R ar = 0, ai = 0;
#pragma unroll
        for (int j = -1; j <= 2; ++j) {
            const R w = lanczos_w(sub_rn(u, (R)j));
            const cplx<R> s = TR::load(v, (size_t)(ti + j));
            ar = add_rn(ar, mul_rn(w, s.re));
            ai = add_rn(ai, mul_rn(w, s.im));
        }
        o.re = ar;
        o.im = ai;
        return o;
    }
    return o;
}

// complex multiply with individually rounded products (MATLAB .* on complex)
template <typename R> __device__ __forceinline__ cplx<R> cmul(cplx<R> a, cplx<R> b) {
    cplx<R> o;
    o.re = sub_rn(mul_rn(a.re, b.re), mul_rn(a.im, b.im));
    o.im = add_rn(mul_rn(a.re, b.im), mul_rn(a.im, b.re));
    return o;
}

} // namespace qups
