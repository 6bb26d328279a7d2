// apod_gen.cu — dense image of the closed-form apodization generators (the array API of SURVEY.md §8f-1).
//
// apAcceptanceAngle / apCosineAngle / apApertureGrowth / apTranslatingAperture / apScanline / apTxParallelogram
// (src/UltrasoundSystem.m:4892-5429) return ND arrays broadcastable to I1 x I2 x I3 x N x M.  This kernel writes
// the same arrays on the device, real fp32 (or complex with zero imaginary part, the type the reference's GPU
// branch forces, kern/das_spec.m:237-243), from the SAME device functions the fused DAS kernel evaluates
// (apod_fused.cuh) — the two paths agree bit for bit.  One thread per output element, coalesced along I1.
#include "das_args.cuh"

namespace qups {

__global__ void __launch_bounds__(256) apod_generate_kernel(FusedApod fa, int which, float *out, int as_complex, const float *Pi,
                                                            const float *Pr, uint32_t I1, uint32_t I2, uint32_t I3, uint64_t I, uint64_t NM) {
    const uint64_t total = I * NM;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = e % I;
        const uint32_t k = (uint32_t)(e / I); // receive n or transmit m
        const float px = __ldg(Pi + 3 * i), py = __ldg(Pi + 3 * i + 1), pz = __ldg(Pi + 3 * i + 2);
        const uint32_t i1 = (uint32_t)(i % I1), i2 = (uint32_t)((i / I1) % I2), i3 = (uint32_t)(i / ((uint64_t)I1 * I2));
        const float plat = ap_lateral(fa, px, i1, i2, i3);
        const float w = which == 0 ? ap_rx_weight(fa, px, py, pz, plat, Pr, k) : ap_tx_weight(fa, px, py, pz, plat, k);
        if (as_complex) reinterpret_cast<float2 *>(out)[e] = make_float2(w, 0.f);
        else out[e] = w;
    }
}

int launch_apod_generate(const FusedApod &fa, int which, float *out, int as_complex, const float *Pi, const float *Pr,
                         uint64_t I1, uint64_t I2, uint64_t I3, uint64_t NM, cudaStream_t st) {
    const uint64_t I = I1 * I2 * I3, total = I * NM;
    if (total == 0) return 0;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (total + 255) / 256, cap = (uint64_t)sms * 16;
    apod_generate_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(fa, which, out, as_complex, Pi, Pr, (uint32_t)I1,
                                                                              (uint32_t)I2, (uint32_t)I3, I, NM);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace qups
