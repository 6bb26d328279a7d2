// convd.cu — batched direct 1-D convolution along one dimension.
//
// Replaces src/convd.cu:95-156 (conv_temp / conv, convf, convc, convcf; launcher kern/convd.m:135-201).
// Canonical layout: x is (C, Lx, S), y is (C|1, Ly, S|1), z is (C, Lz, S) with C = elements before the working
// dimension (stride 1) and S = batches after it; singleton C / S of y broadcast.
//   z(c, l, s) = sum_i x(c, i, s) * y(c, l + l0 - i, s)        l0 = 0 ('full'), ceil((Ly-1)/2) ('same'), Ly-1 ('valid')
// One thread per output element, i ascending: deterministic.  Compiled with -fmad=false.
#include "common.cuh"
#include "other_kernels.cuh"

namespace qups {
void count_launch(uint64_t n);

template <typename T> struct cv;
template <> struct cv<float> { static __device__ float mac(float a, float x, float y) { return add_rn(a, mul_rn(x, y)); } static __device__ float zero() { return 0.f; } };
template <> struct cv<double> { static __device__ double mac(double a, double x, double y) { return add_rn(a, mul_rn(x, y)); } static __device__ double zero() { return 0.0; } };
template <> struct cv<float2> {
    static __device__ float2 mac(float2 a, float2 x, float2 y) {
        a.x = add_rn(a.x, sub_rn(mul_rn(x.x, y.x), mul_rn(x.y, y.y)));
        a.y = add_rn(a.y, add_rn(mul_rn(x.x, y.y), mul_rn(x.y, y.x)));
        return a;
    }
    static __device__ float2 zero() { return make_float2(0.f, 0.f); }
};
template <> struct cv<double2> {
    static __device__ double2 mac(double2 a, double2 x, double2 y) {
        a.x = add_rn(a.x, sub_rn(mul_rn(x.x, y.x), mul_rn(x.y, y.y)));
        a.y = add_rn(a.y, add_rn(mul_rn(x.x, y.y), mul_rn(x.y, y.x)));
        return a;
    }
    static __device__ double2 zero() { return make_double2(0.0, 0.0); }
};

// storage type -> accumulation type: half data (convh / convch, src/convd.cu:141,153) is widened per element — exactly — and
// accumulated in fp32 (the reference accumulates in half), rounded to half once per output element
template <typename T> struct io { using acc = T; static __device__ T ld(const T *p, uint64_t i) { return p[i]; } static __device__ void st(T *p, uint64_t i, T v) { p[i] = v; } };
template <> struct io<__half> {
    using acc = float;
    static __device__ float ld(const __half *p, uint64_t i) { return __half2float(p[i]); }
    static __device__ void st(__half *p, uint64_t i, float v) { p[i] = __float2half_rn(v); }
};
template <> struct io<__half2> {
    using acc = float2;
    static __device__ float2 ld(const __half2 *p, uint64_t i) { return __half22float2(p[i]); }
    static __device__ void st(__half2 *p, uint64_t i, float2 v) { p[i] = __floats2half2_rn(v.x, v.y); }
};

template <typename T>
__global__ void __launch_bounds__(256) convd_kernel(T *z, const T *x, const T *y, uint64_t C, uint64_t S, long long Lx, long long Ly,
                                                    long long Lz, long long l0, uint64_t yC, uint64_t yS) {
    const uint64_t total = C * (uint64_t)Lz * S;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = o % C, l = (o / C) % (uint64_t)Lz, s = o / (C * (uint64_t)Lz);
        const T *xp = x + c + C * (uint64_t)Lx * s;
        const T *yp = y + (yC > 1 ? c : 0) + yC * (uint64_t)Ly * (yS > 1 ? s : 0);
        const long long k = (long long)l + l0;               // full-convolution index
        long long i0 = k - (Ly - 1); if (i0 < 0) i0 = 0;
        long long i1 = k; if (i1 > Lx - 1) i1 = Lx - 1;
        using A = typename io<T>::acc;
        A acc = cv<A>::zero();
        for (long long i = i0; i <= i1; ++i) acc = cv<A>::mac(acc, io<T>::ld(xp, (uint64_t)i * C), io<T>::ld(yp, (uint64_t)(k - i) * yC));
        io<T>::st(z, o, acc);
    }
}

int launch_convd(const qups_convd_params &p, void *z, const void *x, const void *y, cudaStream_t st) {
    const long long Lx = (long long)p.Lx, Ly = (long long)p.Ly;
    long long Lz, l0;
    if (p.shape == 0) { Lz = Lx + Ly - 1; l0 = 0; }
    else if (p.shape == 1) { Lz = Lx; l0 = (Ly - 1 + 1) / 2; }
    else { Lz = Lx - Ly + 1; if (Lz < 0) Lz = 0; l0 = Ly - 1; }
    if (Lx == 0 || Ly == 0) Lz = (p.shape == 0 || p.shape == 2) ? 0 : Lx;
    const uint64_t total = p.C * (uint64_t)Lz * p.S;
    if (total == 0) return 0;
    const uint64_t g = (total + 255) / 256;
    const unsigned grid = (unsigned)(g < (1u << 20) ? g : (1u << 20));
#define QUPS_CV(T) convd_kernel<T><<<grid, 256, 0, st>>>((T *)z, (const T *)x, (const T *)y, p.C, p.S, Lx, Ly, Lz, l0, p.yC, p.yS)
    if (p.dtype == QUPS_F32) { if (p.is_complex) QUPS_CV(float2); else QUPS_CV(float); }
    else if (p.dtype == QUPS_F64) { if (p.is_complex) QUPS_CV(double2); else QUPS_CV(double); }
    else if (p.dtype == QUPS_F16) { if (p.is_complex) QUPS_CV(__half2); else QUPS_CV(__half); }
    else return -3;
#undef QUPS_CV
    count_launch(1);
    return (int)cudaGetLastError();
}
} // namespace qups
