// exp_half2_ring.cu — would a native half2 ring pay?  ("missing" #1: stage __half2 windows straight from the
// fp16 cube, LDS.32 per tap, convert in registers, accumulate fp32.)
//
// Both kernels run the all-FAST cubic stage body of das_tiled (csrc/das_tiled.cu fast_pair2: scalar individually rounded delay
// sequence, add.rd magic-number index, Keys weights in 9 packed ops, 8 packed FFMA2 accumulates, two pixel rows per thread, 16
// traces per stage, per-trace slot offsets) from windows that are already resident in shared memory — no producer, no
// barriers, no edges: the consumer body alone, 8 warps per CTA, 2 CTAs per SM.
//   f32 ring : float2 samples, LDS.64 per tap          (what ships: fp16 cubes are widened once into an fp32 scratch cube)
//   f16 ring : __half2 samples, LDS.32 per tap + one half2 -> float2 conversion per tap (2 cvt), fp32 weights / accumulators
// The half ring halves the shared-memory wavefronts (a warp of 4-byte loads is one wavefront, of 8-byte loads two) and adds 16
// conversion instructions to the ~41 of the body per 2 pairs.
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exph scripts/exp_half2_ring.cu && /tmp/exph
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kNT = 16;      // traces per stage
constexpr int kWmax = 128;   // samples per slot
constexpr int kWarps = 8;

__device__ __forceinline__ float sample_pos(float dv, float dr, float cinv, float t0, float fs) {
    // canonical sequence (csrc/common.cuh): tau = cinv (dv + dr); tau -= t0; xq = 1 + tau fs — each op individually rounded
    const float tau = __fsub_rn(__fmul_rn(cinv, __fadd_rn(dv, dr)), t0);
    return __fadd_rn(1.0f, __fmul_rn(tau, fs));
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds32h(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(a));
    return __half22float2(*reinterpret_cast<const __half2 *>(&r));
}

template <int HALF>
__device__ __forceinline__ void pair2(float2 dv, float2 dr, float cinv, float t0, float fs, uint32_t soff, float2 &acc0, float2 &acc1) {
    float2 xq;
    xq.x = sample_pos(dv.x, dr.x, cinv, t0, fs);
    xq.y = sample_pos(dv.y, dr.y, cinv, t0, fs);
    const float2 tm = make_float2(__fadd_rd(xq.x, 8388608.f), __fadd_rd(xq.y, 8388608.f));
    const float2 kf = __fadd2_rn(tm, make_float2(-8388608.f, -8388608.f));
    const float2 u = __ffma2_rn(kf, make_float2(-1.f, -1.f), xq);
    constexpr int sh = HALF ? 2 : 3, st = HALF ? 4 : 8;
    const uint32_t a0 = soff + (__float_as_uint(tm.x) << sh), a1 = soff + (__float_as_uint(tm.y) << sh);
    float2 p0, p1, p2, p3, q0, q1, q2, q3;
    if (HALF) {
        p0 = lds32h(a0); p1 = lds32h(a0 + st); p2 = lds32h(a0 + 2 * st); p3 = lds32h(a0 + 3 * st);
        q0 = lds32h(a1); q1 = lds32h(a1 + st); q2 = lds32h(a1 + 2 * st); q3 = lds32h(a1 + 3 * st);
    } else {
        p0 = lds64(a0); p1 = lds64(a0 + st); p2 = lds64(a0 + 2 * st); p3 = lds64(a0 + 3 * st);
        q0 = lds64(a1); q1 = lds64(a1 + st); q2 = lds64(a1 + 2 * st); q3 = lds64(a1 + 3 * st);
    }
    const float2 one2 = make_float2(1.f, 1.f), m2 = make_float2(-2.f, -2.f);
    const float2 v = __ffma2_rn(u, make_float2(-1.f, -1.f), one2);
    const float2 h = __fmul2_rn(__fmul2_rn(u, v), make_float2(-0.5f, -0.5f));
    const float2 w0 = __fmul2_rn(h, v), w3 = __fmul2_rn(h, u);
    const float2 w1 = __ffma2_rn(m2, w0, __fadd2_rn(v, w3)), w2 = __ffma2_rn(m2, w3, __fadd2_rn(u, w0));
    acc0 = __ffma2_rn(p0, make_float2(w0.x, w0.x), acc0); acc1 = __ffma2_rn(q0, make_float2(w0.y, w0.y), acc1);
    acc0 = __ffma2_rn(p1, make_float2(w1.x, w1.x), acc0); acc1 = __ffma2_rn(q1, make_float2(w1.y, w1.y), acc1);
    acc0 = __ffma2_rn(p2, make_float2(w2.x, w2.x), acc0); acc1 = __ffma2_rn(q2, make_float2(w2.y, w2.y), acc1);
    acc0 = __ffma2_rn(p3, make_float2(w3.x, w3.x), acc0); acc1 = __ffma2_rn(q3, make_float2(w3.y, w3.y), acc1);
}

// geo[stage] = (dv base, t0): changes per stage like the transmit does; the per-lane part mimics an 8 x 4 x 2 lane patch
template <int HALF>
__global__ void __launch_bounds__(kWarps * 32, 2) ring_kernel(const float2 *x, const float2 *geo, float2 *y, int stages, float cinv, float fs) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (HALF) {
        __half2 *w = reinterpret_cast<__half2 *>(smem);
        for (int i = threadIdx.x; i < kNT * kWmax; i += blockDim.x) w[i] = __float22half2_rn(x[i]);
    } else {
        float2 *w = reinterpret_cast<float2 *>(smem);
        for (int i = threadIdx.x; i < kNT * kWmax; i += blockDim.x) w[i] = x[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float pa = (float)((warp & 3) * 8 + (lane & 7)), pb = (float)(((warp >> 2) * 4 + (lane >> 3)) * 2);
    float2 dr[kNT];
#pragma unroll
    for (int j = 0; j < kNT; ++j) { // receive path lengths: a few samples of spread over the tile, drifting with the trace
        dr[j].x = 10.f + 0.31f * pa + 0.83f * pb + 0.5f * (float)j;
        dr[j].y = 10.f + 0.31f * pa + 0.83f * (pb + 1.f) + 0.5f * (float)j;
    }
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) - ((0x4B000000u << (HALF ? 2 : 3)));
    constexpr uint32_t slot = kWmax * (HALF ? 4 : 8);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
    for (int s = 0; s < stages; ++s) {
        const float2 g = __ldg(geo + (s & 1023));
        const float2 dv = make_float2(g.x + 0.11f * pa + 0.6f * pb, g.x + 0.11f * pa + 0.6f * (pb + 1.f));
        float2 sa0 = make_float2(0.f, 0.f), sa1 = sa0;
#pragma unroll
        for (int j = 0; j < kNT; ++j) pair2<HALF>(dv, dr[j], cinv, g.y, fs, base + j * slot, sa0, sa1);
        acc0.x += sa0.x; acc0.y += sa0.y; acc1.x += sa1.x; acc1.y += sa1.y;
    }
    y[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
}

template <int HALF> static double run(const float2 *dx, const float2 *dgeo, float2 *dy, int grid, int stages, double *sum) {
    const int smem = kNT * kWmax * (HALF ? 4 : 8);
    cudaFuncSetAttribute(ring_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) ring_kernel<HALF><<<grid, kWarps * 32, smem>>>(dx, dgeo, dy, stages, 1.0f, 1.0f);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) ring_kernel<HALF><<<grid, kWarps * 32, smem>>>(dx, dgeo, dy, stages, 1.0f, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const size_t ny = (size_t)grid * kWarps * 32;
    float2 *hy = (float2 *)malloc(ny * sizeof(float2));
    cudaMemcpy(hy, dy, ny * sizeof(float2), cudaMemcpyDeviceToHost);
    double s = 0;
    for (size_t i = 0; i < ny; ++i) s += (double)hy[i].x + (double)hy[i].y;
    free(hy);
    *sum = s;
    return ms / reps;
}

int main() {
    const int grid = 148 * 2 * 4, stages = 4096;
    float2 *hx = (float2 *)malloc(kNT * kWmax * sizeof(float2)), *hg = (float2 *)malloc(1024 * sizeof(float2));
    srand(1);
    for (int i = 0; i < kNT * kWmax; ++i) hx[i] = make_float2((float)(rand() % 2001 - 1000) / 64.f, (float)(rand() % 2001 - 1000) / 64.f); // exact in fp16
    for (int i = 0; i < 1024; ++i) hg[i] = make_float2(5.f + (float)(rand() % 4000) / 100.f, (float)(rand() % 100) / 50.f);   // xq stays inside [2, 120]
    float2 *dx, *dg, *dy;
    cudaMalloc(&dx, kNT * kWmax * sizeof(float2)); cudaMalloc(&dg, 1024 * sizeof(float2)); cudaMalloc(&dy, (size_t)grid * kWarps * 32 * sizeof(float2));
    cudaMemcpy(dx, hx, kNT * kWmax * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(dg, hg, 1024 * sizeof(float2), cudaMemcpyHostToDevice);
    double s32 = 0, s16 = 0;
    const double pairs = (double)grid * kWarps * 32 * 2 * kNT * stages;
    const double t32 = run<0>(dx, dg, dy, grid, stages, &s32);
    const double t16 = run<1>(dx, dg, dy, grid, stages, &s16);
    printf("cubic stage body, fp32 ring (LDS.64 per tap)            : %8.3f ms  %.3e pairs/s\n", t32, pairs / (t32 * 1e-3));
    printf("cubic stage body, half2 ring (LDS.32 + 2 cvt per tap)   : %8.3f ms  %.3e pairs/s  (%.3fx)\n", t16, pairs / (t16 * 1e-3), t32 / t16);
    printf("sum over all outputs: fp32 ring %.9e  half2 ring %.9e  rel diff %.2e (data exact in fp16)\n", s32, s16, fabs(s32 - s16) / fabs(s32));
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
