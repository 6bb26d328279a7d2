// apod_fused.cuh — closed-form apodization generators evaluated on the device.
//
// The reference builds dense ND masks on the host and hands them to DAS as 'apod' arrays
// (src/UltrasoundSystem.m:4892-5429); at the headline size a pixel x receive mask is 1-2 GB and makes DAS
// HBM-bound.  Every generator is a closed-form function of (pixel, element) geometry that the DAS kernel already
// holds in registers, so it is evaluated in-kernel from this small parameter block instead.  The same functions
// back qups_apod_generate (the dense-array API), so the fused and the array path agree bit for bit.
//
// Numerics: one individually rounded fp32 operation per step (explicit _rn intrinsics, never contracted), in the
// order of the reference's expressions; oracle/apod_np.py restates the same sequence in NumPy float32.
//   rx (pixel i, receive n):
//     ACCEPTANCE  apAcceptanceAngle :5303   r = Pi - Pn; r /= |r|; w = (n . r >= cosd(theta))
//     COSINE      apCosineAngle     :5377   c = clamp(n . r, -1, 1); w = cosd(min(90, (90/theta) acosd(c)))
//     GROWTH      apApertureGrowth  :5165   planar: d = Xn - Xi, z = Zi; else d,z in the element's frame;
//                                           w = (z > f |2d|) & (|2d| < Dmax)
//     TRANSLATING apTranslatingAperture :5074 (receive factor)  w = |xi - xn| <= tol
//   tx (pixel i, transmit m):
//     SCANLINE    apScanline :4892          w = |xi - xv| <  tol
//     TRANSLATING apTranslatingAperture (transmit factor)       w = |xi - xv| <= tol
//     PARALLELOGRAM apTxParallelogram :5269 xp_k = x - sind(phi_k+theta_m) (z / cosd(phi_k+theta_m));
//                                           w = any_k(lo < xp_k) & any_k(xp_k <= hi)
#pragma once
#include "common.cuh"

namespace qups {

enum { AP_RX_NONE = 0, AP_RX_ACCEPTANCE = 1, AP_RX_COSINE = 2, AP_RX_GROWTH = 3, AP_RX_TRANSLATING = 4 };
enum { AP_TX_NONE = 0, AP_TX_SCANLINE = 1, AP_TX_TRANSLATING = 2, AP_TX_PARALLELOGRAM = 3 };

struct FusedApod {
    int rx_kind, tx_kind;
    float rx_p[4];        // ACCEPTANCE: [cosd(theta)] | COSINE: [90/theta] | GROWTH: [f, Dmax, nonplanar] | TRANSLATING: [tol]
    float tx_p[4];        // SCANLINE / TRANSLATING: [tol] | PARALLELOGRAM: [lo, hi]
    const float *rx_aux;  // ACCEPTANCE/COSINE: 3 x N element normals | GROWTH (non-planar): 2 x N [cosd(ae); sind(ae)] | TRANSLATING: N lateral coords xn
    const float *tx_aux;  // SCANLINE/TRANSLATING: M lateral coords xv | PARALLELOGRAM: 4 x M [sind(t+p1); cosd(t+p1); sind(t+p2); cosd(t+p2)]
    const float *lat;     // optional lateral coordinate per pixel index along lat_dim (ScanPolar: scan.a); NULL -> the pixel's x
    int lat_dim;          // 1..3 (which pixel-grid dimension `lat` runs along)
};

// direction cosine between the element normal and the element -> pixel ray (shared by ACCEPTANCE and COSINE)
__device__ __forceinline__ float ap_dircos(float px, float py, float pz, float ex, float ey, float ez, float nx, float ny, float nz) {
    const float rx = sub_rn(px, ex), ry = sub_rn(py, ey), rz = sub_rn(pz, ez);
    const float d = norm3(rx, ry, rz);
    return dot3(nx, ny, nz, div_rn(rx, d), div_rn(ry, d), div_rn(rz, d));
}

// weight of receive element n for a pixel at (px,py,pz) with lateral coordinate plat
__device__ __forceinline__ float ap_rx_weight(const FusedApod &f, float px, float py, float pz, float plat, const float *Pr, uint32_t n) {
    switch (f.rx_kind) {
        case AP_RX_ACCEPTANCE: {
            const float c = ap_dircos(px, py, pz, __ldg(Pr + 3 * n), __ldg(Pr + 3 * n + 1), __ldg(Pr + 3 * n + 2),
                                      __ldg(f.rx_aux + 3 * n), __ldg(f.rx_aux + 3 * n + 1), __ldg(f.rx_aux + 3 * n + 2));
            return (c >= f.rx_p[0]) ? 1.f : 0.f; // NaN (pixel on the element) -> 0, as MATLAB's >=
        }
        case AP_RX_COSINE: {
            float c = ap_dircos(px, py, pz, __ldg(Pr + 3 * n), __ldg(Pr + 3 * n + 1), __ldg(Pr + 3 * n + 2),
                                __ldg(f.rx_aux + 3 * n), __ldg(f.rx_aux + 3 * n + 1), __ldg(f.rx_aux + 3 * n + 2));
            c = fmaxf(-1.f, fminf(1.f, c)); // max(-1, min(1, r)): MATLAB min/max ignore NaN -> +1
            if (!(c == c)) c = 1.f;
            const float t = mul_rn(f.rx_p[0], acosf(c)); // (90/theta) * angle, in radians
            return (t >= 1.57079632679489662f) ? 0.f : cosf(t);
        }
        case AP_RX_GROWTH: {
            const float ex = __ldg(Pr + 3 * n), ez = __ldg(Pr + 3 * n + 2);
            float d, z;
            if (f.rx_p[2] != 0.f) { // non-planar array: width / depth in the element's frame
                const float ca = __ldg(f.rx_aux + 2 * n), sa = __ldg(f.rx_aux + 2 * n + 1);
                const float dx = sub_rn(px, ex), dz = sub_rn(pz, ez);
                d = sub_rn(mul_rn(dx, ca), mul_rn(dz, sa));
                z = fabsf(add_rn(mul_rn(dz, ca), mul_rn(dx, sa)));
            } else {
                d = sub_rn(ex, px);
                z = pz;
            }
            const float a2d = fabsf(mul_rn(2.f, d));
            return (z > mul_rn(f.rx_p[0], a2d) && a2d < f.rx_p[1]) ? 1.f : 0.f;
        }
        case AP_RX_TRANSLATING:
            return (fabsf(sub_rn(plat, __ldg(f.rx_aux + n))) <= f.rx_p[0]) ? 1.f : 0.f;
        default:
            return 1.f;
    }
}

// weight of transmit m for a pixel
__device__ __forceinline__ float ap_tx_weight(const FusedApod &f, float px, float py, float pz, float plat, uint32_t m) {
    switch (f.tx_kind) {
        case AP_TX_SCANLINE:
            return (fabsf(sub_rn(plat, __ldg(f.tx_aux + m))) < f.tx_p[0]) ? 1.f : 0.f;
        case AP_TX_TRANSLATING:
            return (fabsf(sub_rn(plat, __ldg(f.tx_aux + m))) <= f.tx_p[0]) ? 1.f : 0.f;
        case AP_TX_PARALLELOGRAM: {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(f.tx_aux) + m);
            const float x0 = sub_rn(px, mul_rn(q.x, div_rn(pz, q.y)));
            const float x1 = sub_rn(px, mul_rn(q.z, div_rn(pz, q.w)));
            return ((f.tx_p[0] < x0 || f.tx_p[0] < x1) && (x0 <= f.tx_p[1] || x1 <= f.tx_p[1])) ? 1.f : 0.f;
        }
        default:
            return 1.f;
    }
}

// lateral coordinate of the pixel with grid indices (i1,i2,i3)
__device__ __forceinline__ float ap_lateral(const FusedApod &f, float px, uint32_t i1, uint32_t i2, uint32_t i3) {
    if (!f.lat) return px;
    return __ldg(f.lat + (f.lat_dim == 1 ? i1 : (f.lat_dim == 2 ? i2 : i3)));
}

} // namespace qups
