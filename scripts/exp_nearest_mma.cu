// exp_nearest_mma.cu — the experiment BASELINE.json's north_star asks for before tensor cores are ruled out:
// "tensor cores used only if recasting the apodized rx-sum as a dense GEMV is measured faster".
//
// Nearest-neighbour DAS for one trace is y(i) += x[k(i)]: a product of a one-hot selection matrix S (pixels x window samples)
// with the window.  Both kernels below consume the SAME index stream (a synthetic, smoothly varying k(i, trace) with the
// spread of the headline geometry: neighbouring pixels 0-2 samples apart) from a shared-memory window and accumulate in fp32:
//   gather_kernel : what das_tiled's 1-tap body does — one LDS.64 per (pixel, trace) pair, FADD2 accumulate;
//   mma_kernel    : per 16 pixels x 16-sample window one mma.sync.m16n8k16 (bf16 in, fp32 accumulate): the A fragment is the
//                   one-hot S built in registers by compare/select from the pixels' indices (obtained by shuffle), the B
//                   fragment is the window pre-split into three bf16 planes per component (hi + mid + lo = the fp32 value
//                   exactly, so the selected sample is reproduced bit for bit), read from shared memory; the fp32
//                   accumulator fragment persists across traces, so there is no per-trace epilogue.
// The MMA variant is given every advantage: the window is already split and resident, all 16 pixels of a fragment share one
// 16-sample window (base index = min over the fragment, guaranteed by construction here), no producer, no edge handling.
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exp scripts/exp_nearest_mma.cu && /tmp/exp
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kWin = 256;      // window samples resident in shared memory
constexpr int kTraces = 4096;  // traces per CTA pass (loop length)
constexpr int kWarps = 8;

__device__ __forceinline__ int index_of(int pixel, int trace) { // smooth in the pixel, drifting with the trace; fits kWin - 16
    return ((pixel * 3) >> 2) + ((trace * 7) & 127);            // neighbouring pixels 0..1 apart, 16 pixels span <= 12 samples
}

__global__ void __launch_bounds__(kWarps * 32) gather_kernel(const float2 *x, float2 *y, int iters) {
    __shared__ float2 win[kWin];
    for (int i = threadIdx.x; i < kWin; i += blockDim.x) win[i] = x[i];
    __syncthreads();
    const int pixel = threadIdx.x & 31; // 32 pixels per warp, two passes of 16 like the MMA variant's two fragments
    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
    for (int it = 0; it < iters; ++it)
#pragma unroll 8
        for (int t = 0; t < kTraces; t += 2) {
            const float2 a = win[index_of(pixel, t)], b = win[index_of(pixel, t + 1)];
            acc0.x += a.x; acc0.y += a.y; acc1.x += b.x; acc1.y += b.y;
        }
    y[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
}

__global__ void __launch_bounds__(kWarps * 32) mma_kernel(const float2 *x, float2 *y, int iters) {
    // window pre-split: planes[c][k] packs for sample k the bf16 pair (plane c of re/im): B fragment rows k, columns
    // n = 0..5 = re_hi, re_mid, re_lo, im_hi, im_mid, im_lo (6, 7 zero).  B fragment of m16n8k16: thread holds b0,b1 = rows
    // (lane%4)*2 + {0,1}, b2,b3 = rows + 8, column lane/4.
    __shared__ __nv_bfloat16 Bs[8][kWin]; // [column][sample]
    for (int i = threadIdx.x; i < kWin; i += blockDim.x) {
        const float2 v = x[i];
        float r = v.x, q = v.y;
        for (int c = 0; c < 3; ++c) {
            const __nv_bfloat16 hr = __float2bfloat16_rz(r), hq = __float2bfloat16_rz(q);
            Bs[c][i] = hr; Bs[3 + c][i] = hq;
            r -= __bfloat162float(hr); q -= __bfloat162float(hq);
        }
        Bs[6][i] = Bs[7][i] = __float2bfloat16(0.f);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, pixel = lane;
    float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}; // accumulator fragments of the two 16-pixel groups
    const int col = lane >> 2, r0 = (lane & 3) * 2;
    for (int it = 0; it < iters; ++it)
#pragma unroll 4
        for (int t = 0; t < kTraces; ++t) {
            const int k = index_of(pixel, t);
#pragma unroll
            for (int g = 0; g < 2; ++g) { // pixels 16 g .. 16 g + 15
                // window base of the fragment = index of its first pixel (indices are non-decreasing in the pixel here)
                const int base = __shfl_sync(0xffffffffu, k, 16 * g);
                const int ka = __shfl_sync(0xffffffffu, k, 16 * g + (lane >> 2)) - base;     // row lane/4
                const int kb = __shfl_sync(0xffffffffu, k, 16 * g + (lane >> 2) + 8) - base; // row lane/4 + 8
                // A fragment (row-major 16 x 16 bf16): a0a1 = (row, cols r0, r0+1), a2a3 = (row + 8, same), a4a5 = (row, cols + 8), a6a7
                auto onehot = [&](int kr, int c) -> uint32_t { return kr == c ? 0x00003F80u : (kr == c + 1 ? 0x3F800000u : 0u); };
                const uint32_t a0 = onehot(ka, r0), a1 = onehot(kb, r0), a2 = onehot(ka, r0 + 8), a3 = onehot(kb, r0 + 8);
                const __nv_bfloat16 *bp = &Bs[col][base + r0];
                const uint32_t b0 = (uint32_t)__bfloat16_as_ushort(bp[0]) | ((uint32_t)__bfloat16_as_ushort(bp[1]) << 16);
                const uint32_t b1 = (uint32_t)__bfloat16_as_ushort(bp[8]) | ((uint32_t)__bfloat16_as_ushort(bp[9]) << 16);
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[g][0]), "+f"(d[g][1]), "+f"(d[g][2]), "+f"(d[g][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            }
        }
    // accumulator fragment: d0,d1 = (row lane/4, cols (lane%4)*2 + {0,1}), d2,d3 = (row + 8, same): store raw (the sum over
    // the 3 planes is a 6-column reduction the real kernel would do once per pixel at the very end)
    y[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(d[0][0] + d[0][1] + d[1][0] + d[1][1], d[0][2] + d[0][3] + d[1][2] + d[1][3]);
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 2, iters = 8;
    float2 *x, *y;
    cudaMalloc(&x, sizeof(float2) * kWin);
    cudaMalloc(&y, sizeof(float2) * grid * kWarps * 32);
    float2 h[kWin];
    for (int i = 0; i < kWin; ++i) h[i] = make_float2((float)(i * 37 % 101) * 1.0009765f, (float)(i * 53 % 89) * 0.99951f);
    cudaMemcpy(x, h, sizeof(h), cudaMemcpyHostToDevice);
    const double pairs = (double)grid * kWarps * 32 * kTraces * iters;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const size_t ny = (size_t)grid * kWarps * 32;
    float2 *hy = (float2 *)malloc(sizeof(float2) * ny);
    auto total = [&]() { cudaMemcpy(hy, y, sizeof(float2) * ny, cudaMemcpyDeviceToHost); double s = 0; for (size_t i = 0; i < ny; ++i) s += (double)hy[i].x + hy[i].y; return s; };
    double sg = 0, sm = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); gather_kernel<<<grid, kWarps * 32>>>(x, y, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("gather (LDS.64 per pair)        : %8.3f ms  %.3e pairs/s\n", ms, pairs / (ms * 1e-3));
        sg = total();
        cudaEventRecord(e0); mma_kernel<<<grid, kWarps * 32>>>(x, y, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("one-hot mma.sync.m16n8k16 bf16x3: %8.3f ms  %.3e pairs/s\n", ms, pairs / (ms * 1e-3));
        sm = total();
    }
    // correctness of the recast: both kernels sum the same samples (sum over all pixels of re + im; fp32 summation order differs)
    printf("sum over all outputs: gather %.9e  mma %.9e  rel diff %.2e\n", sg, sm, (sg - sm) / sg);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
