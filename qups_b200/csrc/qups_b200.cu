// qups_b200.cu — the extern "C" boundary of libqups_b200.so (see include/qups_b200.h).
// Argument validation, dtype dispatch, frame loop (kern/das_spec.m:371-373),
// modulation pre-pass (kern/das_spec.m:413-417) and the host-buffer variants.
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <type_traits>
#include <mutex>

#include "../../include/qups_b200.h"
#include "das_args.cuh"
#include "other_kernels.cuh"

namespace qups {
static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;
static thread_local const char *g_last_das = "none";
void count_launch(uint64_t n) { g_launches += n; }

static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {};
static bool g_pool_failed[64] = {};
cudaError_t ws_alloc(void **p, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMallocAsync(p, bytes, st);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool_failed[dev]) return cudaMallocAsync(p, bytes, st);
        if (!g_pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            if ((e = cudaMemPoolCreate(&g_pools[dev], &props)) != cudaSuccess) { // no pool on this device: default pool from now on
                (void)cudaGetLastError();
                g_pools[dev] = nullptr; g_pool_failed[dev] = true;
                return cudaMallocAsync(p, bytes, st);
            }
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool = g_pools[dev];
    }
    e = cudaMallocFromPoolAsync(p, bytes, pool, st);
    if (e != cudaSuccess) { // fall back to the default pool rather than fail the call
        (void)cudaGetLastError();
        e = cudaMallocAsync(p, bytes, st);
    }
    return e;
}
cudaError_t ws_free(void *p, cudaStream_t st) { return p ? cudaFreeAsync(p, st) : cudaSuccess; }

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(int e, const char *what) {
    return fail(QUPS_ERR_CUDA, "%s: %s", what, cudaGetErrorString((cudaError_t)e));
}

template <typename R>
static int fill_args(DasArgs<R> &a, const qups_das_params *p, const void *Pi, const void *Pr, const void *Pv4,
                     const void *Nv, const void *apod, const void *cinv, const uint64_t *acstride, int need_astride,
                     const qups_apod_fused *fz = nullptr) {
    a.I1 = p->I1; a.I2 = p->I2; a.I3 = p->I3;
    a.I = p->I1 * p->I2 * p->I3;
    a.N = p->N; a.M = p->M; a.T = p->T;
    a.S = (int)p->S;
    a.interp = p->flag & QUPS_FLAG_INTERP_MASK;
    a.keep_rx = (p->flag & QUPS_FLAG_KEEP_RX) != 0;
    a.keep_tx = (p->flag & QUPS_FLAG_KEEP_TX) != 0;
    a.tpose = (p->flag & QUPS_FLAG_TRANSPOSE) != 0;
    a.VS = p->vs != 0; a.DV = p->dv != 0;
    a.apod_real = p->apod_real != 0;
    a.accumulate = p->accumulate != 0;
    a.fs = (R)p->fs;
    a.Pi = (const R *)Pi; a.Pr = (const R *)Pr; a.Pv4 = (const R *)Pv4; a.Nv = (const R *)Nv;
    a.cinv = (const R *)cinv;
    a.apod = apod;
    for (int d = 0; d < 6; ++d) a.cstride[d] = acstride ? acstride[d] : 0;
    a.cstride[5] = 0; // entry 6 of cstride is unused padding in the reference (kern/das_spec.m:259)
    for (int s = 0; s < MAX_APOD; ++s)
        for (int d = 0; d < 6; ++d) a.astride[s][d] = (need_astride && s < a.S) ? acstride[6 + 6 * s + d] : 0;
    a.pitch_hint[0] = p->pitch_hint[0]; a.pitch_hint[1] = p->pitch_hint[1]; a.c_hint = p->c_hint;
    a.fused = 0;
    a.fa = FusedApod{};
    if (fz && need_astride) {
        a.fused = 1;
        a.fa.rx_kind = fz->rx_kind; a.fa.tx_kind = fz->tx_kind;
        for (int k = 0; k < 4; ++k) { a.fa.rx_p[k] = fz->rx_p[k]; a.fa.tx_p[k] = fz->tx_p[k]; }
        a.fa.rx_aux = (const float *)fz->rx_aux; a.fa.tx_aux = (const float *)fz->tx_aux;
        a.fa.lat = (const float *)fz->lat; a.fa.lat_dim = fz->lat_dim;
    }
    return 0;
}

static int validate_fused(const qups_apod_fused *f, const qups_das_params *p) {
    if (!f) return fail(QUPS_ERR_INVALID, "fused apodization spec is NULL");
    if (f->struct_size != sizeof(qups_apod_fused))
        return fail(QUPS_ERR_INVALID, "apod.struct_size %u != %zu (header/library mismatch)", f->struct_size, sizeof(qups_apod_fused));
    if (f->rx_kind < QUPS_AP_RX_NONE || f->rx_kind > QUPS_AP_RX_TRANSLATING) return fail(QUPS_ERR_INVALID, "unknown receive apodization kind %d", f->rx_kind);
    if (f->tx_kind < QUPS_AP_TX_NONE || f->tx_kind > QUPS_AP_TX_PARALLELOGRAM) return fail(QUPS_ERR_INVALID, "unknown transmit apodization kind %d", f->tx_kind);
    const bool rx_needs_aux = f->rx_kind == QUPS_AP_RX_ACCEPTANCE_ANGLE || f->rx_kind == QUPS_AP_RX_COSINE_ANGLE ||
                              f->rx_kind == QUPS_AP_RX_TRANSLATING || (f->rx_kind == QUPS_AP_RX_APERTURE_GROWTH && f->rx_p[2] != 0.f);
    if (rx_needs_aux && !f->rx_aux) return fail(QUPS_ERR_INVALID, "receive apodization kind %d needs rx_aux", f->rx_kind);
    if (f->tx_kind != QUPS_AP_TX_NONE && !f->tx_aux) return fail(QUPS_ERR_INVALID, "transmit apodization kind %d needs tx_aux", f->tx_kind);
    if (f->tx_kind == QUPS_AP_TX_PARALLELOGRAM && (reinterpret_cast<uintptr_t>(f->tx_aux) & 15)) return fail(QUPS_ERR_INVALID, "tx_aux must be 16-byte aligned");
    if (f->lat && (f->lat_dim < 1 || f->lat_dim > 3)) return fail(QUPS_ERR_INVALID, "lat_dim must be 1, 2 or 3");
    if (p && p->dtype == QUPS_F64) return fail(QUPS_ERR_UNSUPPORTED, "closed-form apodization is implemented for fp32 geometry (dtype F32 / F16)");
    return 0;
}

static int validate(const qups_das_params *p, bool is_delays) {
    if (!p) return fail(QUPS_ERR_INVALID, "params is NULL");
    if (p->struct_size != sizeof(qups_das_params))
        return fail(QUPS_ERR_INVALID, "params.struct_size %u != %zu (header/library mismatch)", p->struct_size,
                    sizeof(qups_das_params));
    if (p->dtype < QUPS_F32 || p->dtype > QUPS_F64) return fail(QUPS_ERR_INVALID, "unknown dtype %d", p->dtype);
    if (p->S > (uint64_t)MAX_APOD) return fail(QUPS_ERR_UNSUPPORTED, "at most %d apodization arrays", MAX_APOD);
    const int interp = p->flag & QUPS_FLAG_INTERP_MASK;
    if (!is_delays && interp > QUPS_LANCZOS3)
        return fail(QUPS_ERR_INVALID, "Unrecognized interpolation id %d: must be one of nearest(0), linear(1), cubic(2), lanczos3(3)", interp);
    if (!is_delays) {
        if (interp == QUPS_LINEAR && p->T < 2) return fail(QUPS_ERR_INVALID, "linear interpolation needs T >= 2");
        if (interp == QUPS_CUBIC && p->T < 3) return fail(QUPS_ERR_INVALID, "cubic interpolation needs T >= 3");
        if (!(p->fs > 0.0) && p->fs == p->fs && p->fs != 0.0) { /* negative fs allowed on the generic path */ }
    }
    return 0;
}

template <typename DIN, typename DA, typename DOUT, typename R>
static int run_das_typed(const qups_das_params *p, void *y, const void *Pi, const void *Pr, const void *Pv4,
                         const void *Nv, const void *apod, const void *cinv, const uint64_t *acstride, const void *x,
                         cudaStream_t st, int dtype_in, int dtype_out, const qups_apod_fused *fz) {
    DasArgs<R> a;
    fill_args<R>(a, p, Pi, Pr, Pv4, Nv, apod, cinv, acstride, 1, fz);
    const uint64_t F = p->F ? p->F : 1;
    const uint64_t On = a.keep_rx ? a.N : 1, Om = a.keep_tx ? a.M : 1;
    const uint64_t xfs = p->x_frame_stride ? p->x_frame_stride : a.T * a.N * a.M;
    const uint64_t yfs = p->y_frame_stride ? p->y_frame_stride : a.I * On * Om;
    if (a.I == 0 || On * Om == 0) return 0;
    if (a.N == 0 || a.M == 0) { // empty apertures: the sums are empty => zeros
        for (uint64_t f = 0; f < F; ++f) {
            cudaError_t e = cudaMemsetAsync((DOUT *)y + f * yfs, 0, sizeof(DOUT) * a.I * On * Om, st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        }
        return 0;
    }
    for (uint64_t f = 0; f < F; ++f) {
        a.x = (const DIN *)x + f * xfs;
        a.y = (DOUT *)y + f * yfs;
        int rc;
        bool tiled = false;
        if constexpr (sizeof(R) == 4) {
            if (p->path != QUPS_PATH_GENERIC) {
                TiledPlan plan = das_tiled_plan(a, dtype_in, dtype_out);
                if (plan.eligible) tiled = true;
                else if (p->path == QUPS_PATH_TILED)
                    return fail(QUPS_ERR_UNSUPPORTED, "tiled DAS path not applicable: %s", plan.why);
            }
            if (tiled) {
                rc = launch_das_tiled(a, st);
                g_last_das = "das_tiled";
            } else if (a.fused) {
                // closed-form apodization outside the staged kernel's envelope (kept apertures, sound-speed maps, complex
                // or > 1 arrays, ...): materialise the dense weights next to the caller's arrays and run the generic kernel
                rc = 0;
                const int extra = (a.fa.rx_kind != AP_RX_NONE) + (a.fa.tx_kind != AP_TX_NONE);
                if (a.S + extra > MAX_APOD) return fail(QUPS_ERR_UNSUPPORTED, "at most %d apodization arrays incl. the closed-form ones", MAX_APOD);
                const uint64_t dims[5] = {a.I1, a.I2, a.I3, a.N, a.M};
                uint64_t have = 0; // elements of the caller's concatenated apod buffer
                for (int q = 0; q < a.S; ++q) {
                    uint64_t ne = 1;
                    for (int d = 0; d < 5; ++d) if (a.astride[q][d]) ne *= dims[d];
                    have = have > a.astride[q][5] + ne ? have : a.astride[q][5] + ne;
                }
                const uint64_t nrx = a.fa.rx_kind != AP_RX_NONE ? a.I * a.N : 0, ntx = a.fa.tx_kind != AP_TX_NONE ? a.I * a.M : 0;
                DA *buf = nullptr;
                cudaError_t ce = ws_alloc((void **)&buf, sizeof(DA) * (have + nrx + ntx), st); // complex-sized: enough for either element type
                if (ce != cudaSuccess) return fail(QUPS_ERR_ALLOC, "ws_alloc(dense apodization, %llu elements): %s", (unsigned long long)(have + nrx + ntx), cudaGetErrorString(ce));
                if (have) ce = cudaMemcpyAsync(buf, a.apod, (a.apod_real ? sizeof(DA) / 2 : sizeof(DA)) * have, cudaMemcpyDeviceToDevice, st);
                DasArgs<R> g = a;
                g.fused = 0;
                g.apod = buf;
                if constexpr (std::is_same<DA, float2>::value) {
                    if (a.S == 0) g.apod_real = 1;                 // no caller arrays: real weights (half the bytes)
                    const int cplx = g.apod_real ? 0 : 1;          // otherwise match the caller's element type
                    float *base = reinterpret_cast<float *>(buf);
                    const uint64_t ew = cplx ? 2 : 1;              // floats per element
                    uint64_t off = have;
                    if (nrx && ce == cudaSuccess) {
                        rc = launch_apod_generate(a.fa, 0, base + off * ew, cplx, (const float *)a.Pi, (const float *)a.Pr, a.I1, a.I2, a.I3, a.N, st);
                        const uint64_t st6[6] = {a.I1 > 1 ? 1 : 0, a.I2 > 1 ? a.I1 : 0, a.I3 > 1 ? a.I1 * a.I2 : 0, a.N > 1 ? a.I : 0, 0, off};
                        for (int d = 0; d < 6; ++d) g.astride[g.S][d] = st6[d];
                        ++g.S; off += nrx;
                    }
                    if (ntx && ce == cudaSuccess && rc == 0) {
                        rc = launch_apod_generate(a.fa, 1, base + off * ew, cplx, (const float *)a.Pi, (const float *)a.Pr, a.I1, a.I2, a.I3, a.M, st);
                        const uint64_t st6[6] = {a.I1 > 1 ? 1 : 0, a.I2 > 1 ? a.I1 : 0, a.I3 > 1 ? a.I1 * a.I2 : 0, 0, a.M > 1 ? a.I : 0, off};
                        for (int d = 0; d < 6; ++d) g.astride[g.S][d] = st6[d];
                        ++g.S;
                    }
                } else {
                    ws_free(buf, st);
                    return fail(QUPS_ERR_UNSUPPORTED, "closed-form apodization outside the staged kernel needs fp32 apodization arrays");
                }
                if (ce == cudaSuccess && rc == 0) rc = launch_das_generic<DIN, DA, DOUT, R>(g, st);
                ws_free(buf, st);
                if (ce != cudaSuccess) return cuda_fail(ce, "dense apodization copy");
                g_last_das = "das_generic+apod_generate";
            } else {
                rc = launch_das_generic<DIN, DA, DOUT, R>(a, st);
                g_last_das = "das_generic";
            }
        } else {
            if (p->path == QUPS_PATH_TILED) return fail(QUPS_ERR_UNSUPPORTED, "tiled DAS path is fp32 only");
            rc = launch_das_generic<DIN, DA, DOUT, R>(a, st);
            g_last_das = "das_generic";
        }
        if (rc != 0) return cuda_fail(rc, "DAS kernel launch");
    }
    return 0;
}

// fp16 data on the staged kernel: the cube is widened to fp32 ONCE per frame — exactly, and with the re-modulation of
// kern/das_spec.m:413-417 folded into the same pass when fmod != 0 (1 read of the half2 cube + 1 write of the fp32 scratch) —
// and takes the fp32 staged kernel; geometry / accumulation are fp32 either way.  Doing the widening inside the
// kernel instead (a converter warp behind the bulk copies) would repeat it once per pixel tile, ~80x at the headline size.
// Returns 1 when the call was not eligible (caller falls back to the generic mixed-type kernel), else a status <= 0.
static int das_half_tiled(const qups_das_params *p, void *y, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                          const void *cinv, const uint64_t *acstride, const void *x, cudaStream_t st, const qups_apod_fused *fz) {
    if (p->path == QUPS_PATH_GENERIC || p->S != 0) return 1;
    DasArgs<float> a;
    fill_args<float>(a, p, Pi, Pr, Pv4, Nv, nullptr, cinv, acstride, 1, fz);
    a.x = (const void *)16; // alignment probe only: the scratch cube is 256-byte aligned
    if (!das_tiled_plan(a, 0, 0).eligible) return 1;
    const uint64_t F = p->F ? p->F : 1, nel = p->T * p->N * p->M, I = p->I1 * p->I2 * p->I3;
    const uint64_t On = a.keep_rx ? p->N : 1, Om = a.keep_tx ? p->M : 1, ny = I * On * Om;
    const uint64_t xfs = p->x_frame_stride ? p->x_frame_stride : nel, yfs = p->y_frame_stride ? p->y_frame_stride : ny;
    if (nel == 0 || ny == 0) return 1;
    float2 *xs = nullptr, *ys = nullptr;
    const size_t need = sizeof(float2) * nel;
    const bool own_x = !(p->workspace && p->workspace_bytes >= need);
    cudaError_t e = cudaSuccess;
    if (own_x) e = ws_alloc((void **)&xs, need, st); else xs = (float2 *)p->workspace;
    if (e == cudaSuccess && !p->y_f32) e = ws_alloc((void **)&ys, sizeof(float2) * ny, st);
    if (e != cudaSuccess) { if (xs && own_x) ws_free(xs, st); return fail(QUPS_ERR_ALLOC, "ws_alloc(fp32 staging of the fp16 cube): %s", cudaGetErrorString(e)); }
    qups_das_params q = *p;
    q.dtype = QUPS_F32; q.F = 1; q.path = QUPS_PATH_TILED; q.fmod = 0.0; q.workspace = nullptr; q.workspace_bytes = 0;
    int rc = 0;
    for (uint64_t f = 0; f < F && rc == 0; ++f) {
        const __half2 *xf = (const __half2 *)x + f * xfs;
        int ce = (p->fmod != 0.0)
                     ? launch_modulate<__half2, float2, float>(xs, xf, (const float *)Pv4 + 3, 4, p->T, p->N, p->M, a.tpose, (float)p->fs, p->fmod, st)
                     : launch_half2_to_float2(xs, xf, nel, st);
        if (ce) { rc = cuda_fail(ce, "half2 -> float2 staging"); break; }
        void *yo = p->y_f32 ? (void *)((float2 *)y + f * yfs) : (void *)ys;
        if (!p->y_f32 && p->accumulate) // running sum kept by the caller in half precision: widen, add in fp32, narrow
            if ((ce = launch_half2_to_float2(ys, (const __half2 *)y + f * yfs, ny, st))) { rc = cuda_fail(ce, "half2 -> float2"); break; }
        rc = run_das_typed<float2, float2, float2, float>(&q, yo, Pi, Pr, Pv4, Nv, nullptr, cinv, acstride, xs, st, 0, 0, fz);
        if (rc == 0 && !p->y_f32)
            if ((ce = launch_float2_to_half2((__half2 *)y + f * yfs, ys, ny, st))) rc = cuda_fail(ce, "float2 -> half2");
    }
    if (own_x) ws_free(xs, st);
    if (ys) ws_free(ys, st);
    return rc;
}

static int das_impl(const qups_das_params *p, void *y, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                    const void *apod, const void *cinv, const uint64_t *acstride, const void *x, cudaStream_t st,
                    const qups_apod_fused *fz = nullptr) {
    if (int rc = validate(p, false)) return rc;
    if (!y || !Pi || !Pr || !Pv4 || !Nv || !cinv || !x) {
        const uint64_t I = p->I1 * p->I2 * p->I3;
        if (I != 0 && p->N != 0 && p->M != 0) return fail(QUPS_ERR_INVALID, "NULL array argument");
    }
    if (p->S > 0 && (!apod || !acstride)) return fail(QUPS_ERR_INVALID, "S > 0 but apod/acstride is NULL");

    if (p->dtype == QUPS_F16) {
        const int rc = das_half_tiled(p, y, Pi, Pr, Pv4, Nv, cinv, acstride, x, st, fz);
        if (rc <= 0) return rc;
        if (p->path == QUPS_PATH_TILED) return fail(QUPS_ERR_UNSUPPORTED, "tiled DAS path not applicable to this fp16 call");
    }

    // (de)modulation pre-pass — the CPU-branch convention (kern/das_spec.m:413-417): the DATA are
    // re-modulated at absolute time before interpolation.  One frame of scratch.
    if (p->fmod != 0.0 && p->T * p->N * p->M != 0) {
        const uint64_t F = p->F ? p->F : 1;
        const uint64_t nel = p->T * p->N * p->M;
        const int tpose = (p->flag & QUPS_FLAG_TRANSPOSE) != 0;
        const size_t esz = (p->dtype == QUPS_F64) ? sizeof(double2) : sizeof(float2); // fp16 data -> fp32 scratch
        void *scratch = nullptr;
        bool own = false;
        if (p->workspace && p->workspace_bytes >= esz * nel) scratch = p->workspace;
        else {
            cudaError_t e = ws_alloc(&scratch, esz * nel, st);
            if (e != cudaSuccess) return fail(QUPS_ERR_ALLOC, "ws_alloc(%zu): %s", esz * nel, cudaGetErrorString(e));
            own = true;
        }
        qups_das_params q = *p;
        q.fmod = 0.0; q.F = 1; q.workspace = nullptr; q.workspace_bytes = 0;
        const uint64_t On = (p->flag & QUPS_FLAG_KEEP_RX) ? p->N : 1, Om = (p->flag & QUPS_FLAG_KEEP_TX) ? p->M : 1;
        const uint64_t xfs = p->x_frame_stride ? p->x_frame_stride : nel;
        const uint64_t yfs = p->y_frame_stride ? p->y_frame_stride : p->I1 * p->I2 * p->I3 * On * Om;
        int rc = 0;
        for (uint64_t f = 0; f < F && rc == 0; ++f) {
            int e;
            if (p->dtype == QUPS_F32) {
                e = launch_modulate<float2, float2, float>((float2 *)scratch, (const float2 *)x + f * xfs,
                                                           (const float *)Pv4 + 3, 4, p->T, p->N, p->M, tpose, (float)p->fs, p->fmod, st);
                if (e) { rc = cuda_fail(e, "modulate"); break; }
                rc = run_das_typed<float2, float2, float2, float>(&q, (float2 *)y + f * yfs, Pi, Pr, Pv4, Nv, apod, cinv, acstride, scratch, st, 0, 0, fz);
            } else if (p->dtype == QUPS_F16) {
                e = launch_modulate<__half2, float2, float>((float2 *)scratch, (const __half2 *)x + f * xfs,
                                                            (const float *)Pv4 + 3, 4, p->T, p->N, p->M, tpose, (float)p->fs, p->fmod, st);
                if (e) { rc = cuda_fail(e, "modulate"); break; }
                q.path = QUPS_PATH_GENERIC; // apod stays half: generic mixed-type kernel
                if (p->y_f32) rc = run_das_typed<float2, __half2, float2, float>(&q, (float2 *)y + f * yfs, Pi, Pr, Pv4, Nv, apod, cinv, acstride, scratch, st, 0, 0, fz);
                else rc = run_das_typed<float2, __half2, __half2, float>(&q, (__half2 *)y + f * yfs, Pi, Pr, Pv4, Nv, apod, cinv, acstride, scratch, st, 0, 1, fz);
            } else {
                e = launch_modulate<double2, double2, double>((double2 *)scratch, (const double2 *)x + f * xfs,
                                                              (const double *)Pv4 + 3, 4, p->T, p->N, p->M, tpose, p->fs, p->fmod, st);
                if (e) { rc = cuda_fail(e, "modulate"); break; }
                rc = run_das_typed<double2, double2, double2, double>(&q, (double2 *)y + f * yfs, Pi, Pr, Pv4, Nv, apod, cinv, acstride, scratch, st, 2, 2, fz);
            }
        }
        if (own) ws_free(scratch, st);
        return rc;
    }

    switch (p->dtype) {
        case QUPS_F32:
            return run_das_typed<float2, float2, float2, float>(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, st, 0, 0, fz);
        case QUPS_F16:
            if (p->y_f32) return run_das_typed<__half2, __half2, float2, float>(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, st, 1, 0, fz);
            return run_das_typed<__half2, __half2, __half2, float>(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, st, 1, 1, fz);
        default:
            return run_das_typed<double2, double2, double2, double>(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, st, 2, 2, fz);
    }
}

} // namespace qups

using namespace qups;

extern "C" {

int qups_version(void) { return QUPS_B200_VERSION; }
const char *qups_last_error(void) { return g_err; }
const char *qups_last_das_kernel(void) { return g_last_das; }
const char *qups_last_ws2_kernel(void) { return qups::last_ws2_kernel_name(); }
uint64_t qups_launch_count(int reset) {
    const uint64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int qups_das(const qups_das_params *p, void *y, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
             const void *apod, const void *cinv, const uint64_t *acstride, const void *x, qups_stream_t stream) {
    g_err[0] = 0;
    return das_impl(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, (cudaStream_t)stream);
}

int qups_das_fused(const qups_das_params *p, const qups_apod_fused *fz, void *y, const void *Pi, const void *Pr, const void *Pv4,
                   const void *Nv, const void *apod, const void *cinv, const uint64_t *acstride, const void *x, qups_stream_t stream) {
    g_err[0] = 0;
    if (int rc = validate(p, false)) return rc;
    if (int rc = validate_fused(fz, p)) return rc;
    if (fz->rx_kind == QUPS_AP_RX_NONE && fz->tx_kind == QUPS_AP_TX_NONE)
        return das_impl(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, (cudaStream_t)stream);
    if (p->dtype == QUPS_F16 && (p->S != 0 || (p->flag & (QUPS_FLAG_KEEP_RX | QUPS_FLAG_KEEP_TX))))
        return fail(QUPS_ERR_UNSUPPORTED, "closed-form apodization with fp16 data: no apodization arrays / kept apertures");
    return das_impl(p, y, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, (cudaStream_t)stream, fz);
}

int qups_apod_generate(const qups_apod_fused *fz, int32_t which, void *out, int32_t as_complex, const void *Pi, const void *Pr,
                       uint64_t I1, uint64_t I2, uint64_t I3, uint64_t NM, qups_stream_t stream) {
    g_err[0] = 0;
    if (int rc = validate_fused(fz, nullptr)) return rc;
    if (which != 0 && which != 1) return fail(QUPS_ERR_INVALID, "which must be 0 (receive weights) or 1 (transmit weights)");
    if (I1 * I2 * I3 * NM == 0) return 0;
    if (!out || !Pi || (which == 0 && !Pr)) return fail(QUPS_ERR_INVALID, "NULL array argument");
    FusedApod fa{};
    fa.rx_kind = fz->rx_kind; fa.tx_kind = fz->tx_kind;
    for (int k = 0; k < 4; ++k) { fa.rx_p[k] = fz->rx_p[k]; fa.tx_p[k] = fz->tx_p[k]; }
    fa.rx_aux = (const float *)fz->rx_aux; fa.tx_aux = (const float *)fz->tx_aux; fa.lat = (const float *)fz->lat; fa.lat_dim = fz->lat_dim;
    if (int e = launch_apod_generate(fa, which, (float *)out, as_complex, (const float *)Pi, (const float *)Pr, I1, I2, I3, NM, (cudaStream_t)stream))
        return cuda_fail(e, "apod_generate kernel");
    return 0;
}

int qups_chd_prep(const qups_prep_params *p, void *out, const void *in, const void *t0, qups_stream_t stream) {
    g_err[0] = 0;
    if (!p) return fail(QUPS_ERR_INVALID, "params is NULL");
    if (p->struct_size != sizeof(qups_prep_params)) return fail(QUPS_ERR_INVALID, "params.struct_size mismatch");
    if (p->in_dtype < QUPS_IN_REAL_F32 || p->in_dtype > QUPS_IN_REAL_F64) return fail(QUPS_ERR_INVALID, "unknown in_dtype %d", p->in_dtype);
    if (p->out_dtype != QUPS_F32 && p->out_dtype != QUPS_F16) return fail(QUPS_ERR_INVALID, "out_dtype must be F32 or F16");
    if (!(p->fs > 0.0)) return fail(QUPS_ERR_INVALID, "fs must be positive");
    if (p->K * (p->B + p->T + p->A) == 0) return 0;
    if (!out || (!in && p->T)) return fail(QUPS_ERR_INVALID, "NULL array argument");
    PrepArgs a{};
    a.in = in; a.out = out; a.t0 = (const float *)t0;
    a.T = p->T; a.K = p->K; a.B = p->B; a.A = p->A;
    a.traces_per_t0 = p->traces_per_t0 ? p->traces_per_t0 : 1;
    a.n_t0 = p->n_t0 ? p->n_t0 : 1;
    a.in_dtype = p->in_dtype; a.out_half = p->out_dtype == QUPS_F16; a.hilbert = p->hilbert != 0; a.downmix = p->fmix != 0.0;
    a.fs = (float)p->fs;
    a.cmix = (float)(-2.0 * 3.14159265358979323846 * p->fmix);
    const int e = launch_chd_prep(a, (cudaStream_t)stream);
    if (e == -1000) return fail(QUPS_ERR_UNSUPPORTED, "hilbert length %llu: power of two <= 16384 or any length <= 4096", (unsigned long long)(p->B + p->T + p->A));
    if (e) return cuda_fail(e, "chd_prep kernel");
    return 0;
}

int qups_aperture(const qups_aperture_params *p, void *out, void *out2, const void *b, const uint32_t *lags, qups_stream_t stream) {
    g_err[0] = 0;
    if (!p) return fail(QUPS_ERR_INVALID, "params is NULL");
    if (p->struct_size != sizeof(qups_aperture_params)) return fail(QUPS_ERR_INVALID, "params.struct_size mismatch");
    if (p->dtype != QUPS_F32 && p->dtype != QUPS_F64) return fail(QUPS_ERR_INVALID, "dtype must be F32 or F64");
    if (p->op < QUPS_APD_COHFAC || p->op > QUPS_APD_SLSC_ENSEMBLE) return fail(QUPS_ERR_INVALID, "unknown aperture op %d", p->op);
    if (p->C * p->S == 0) return 0;
    if (!out || (!b && p->A)) return fail(QUPS_ERR_INVALID, "NULL array argument");
    const bool need_lags = p->op == QUPS_APD_DMAS || p->op == QUPS_APD_SLSC_AVERAGE || p->op == QUPS_APD_SLSC_ENSEMBLE;
    if (need_lags && p->nlags && !lags) return fail(QUPS_ERR_INVALID, "lags is NULL");
    const int e = launch_aperture(*p, out, out2, b, lags, (cudaStream_t)stream);
    if (e == -4) return fail(QUPS_ERR_ALLOC, "host allocation failed");
    if (e) return cuda_fail(e, "aperture kernel");
    return 0;
}

int qups_delays(const qups_das_params *p, void *tau, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                const void *cinv, const uint64_t *cstride, qups_stream_t stream) {
    g_err[0] = 0;
    if (int rc = validate(p, true)) return rc;
    if (p->dtype == QUPS_F64) {
        DasArgs<double> a;
        fill_args<double>(a, p, Pi, Pr, Pv4, Nv, nullptr, cinv, cstride, 0);
        if (int e = launch_delays<double>(a, (double *)tau, (cudaStream_t)stream)) return cuda_fail(e, "delays kernel");
    } else {
        DasArgs<float> a;
        fill_args<float>(a, p, Pi, Pr, Pv4, Nv, nullptr, cinv, cstride, 0);
        if (int e = launch_delays<float>(a, (float *)tau, (cudaStream_t)stream)) return cuda_fail(e, "delays kernel");
    }
    return 0;
}

int qups_modulate(int32_t dtype, void *xout, const void *x, const void *t0, uint64_t T, uint64_t N, uint64_t M,
                  int32_t transpose, double fs, double fmod, qups_stream_t stream) {
    g_err[0] = 0;
    int e;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == QUPS_F32) e = launch_modulate<float2, float2, float>((float2 *)xout, (const float2 *)x, (const float *)t0, 1, T, N, M, transpose, (float)fs, fmod, st);
    else if (dtype == QUPS_F16) e = launch_modulate<__half2, __half2, float>((__half2 *)xout, (const __half2 *)x, (const float *)t0, 1, T, N, M, transpose, (float)fs, fmod, st);
    else if (dtype == QUPS_F64) e = launch_modulate<double2, double2, double>((double2 *)xout, (const double2 *)x, (const double *)t0, 1, T, N, M, transpose, fs, fmod, st);
    else return fail(QUPS_ERR_INVALID, "unknown dtype %d", dtype);
    if (e) return cuda_fail(e, "modulate kernel");
    return 0;
}

// ---- host-buffer entry point -------------------------------------------------------------------
namespace qups {
// per-host-thread staging state, kept between calls (see qups_host_release)
struct HostWs {
    static constexpr int NB = 9, NEV = 32;
    int dev = -1;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev[NEV] = {};
    void *buf[NB] = {};
    size_t cap[NB] = {};
    void release() {
        if (dev >= 0) cudaSetDevice(dev);
        for (int i = 0; i < NB; ++i) { if (buf[i]) cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
        for (int i = 0; i < NEV; ++i) { if (ev[i]) cudaEventDestroy(ev[i]); ev[i] = nullptr; }
        if (s_copy) cudaStreamDestroy(s_copy);
        if (s_comp) cudaStreamDestroy(s_comp);
        s_copy = s_comp = nullptr;
        dev = -1;
    }
    int ensure(int device) {
        if (dev == device && s_copy) return 0;
        release();
        cudaError_t e;
        if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
        if ((e = cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
        if ((e = cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
        for (int i = 0; i < NEV; ++i)
            if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        dev = device;
        return 0;
    }
    int get(int i, size_t bytes, void **out) {
        *out = nullptr;
        if (bytes == 0) return 0;
        if (cap[i] < bytes) {
            if (buf[i]) cudaFree(buf[i]);
            buf[i] = nullptr; cap[i] = 0;
            cudaError_t e = cudaMalloc(&buf[i], bytes);
            if (e != cudaSuccess) return fail(QUPS_ERR_ALLOC, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
            cap[i] = bytes;
        }
        *out = buf[i];
        return 0;
    }
    ~HostWs() { /* process teardown: the CUDA context may already be gone; leak by design */ }
};
static thread_local HostWs g_ws;
} // namespace qups

void qups_host_release(void) { g_ws.release(); }

int qups_das_host(const qups_das_params *p, void *y, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                  const void *apod, uint64_t apod_elems, const void *cinv, uint64_t cinv_elems,
                  const uint64_t *acstride, const void *x, int device) {
    g_err[0] = 0;
    if (int rc = validate(p, false)) return rc;
    if (int rc = g_ws.ensure(device)) return rc;
    cudaError_t e;
    // the geometry is in host memory here: derive the launcher hints from it so the staged kernel never probes the device
    qups_das_params hinted = *p;
    if (Pi && cinv && cinv_elems == 1 && !(p->pitch_hint[0] > 0 && p->pitch_hint[1] > 0 && p->c_hint > 0)) {
        auto pitch = [&](uint64_t j) -> double {
            double s2 = 0;
            for (int k = 0; k < 3; ++k) {
                const double d = (p->dtype == QUPS_F64) ? ((const double *)Pi)[3 * j + k] - ((const double *)Pi)[k]
                                                        : (double)((const float *)Pi)[3 * j + k] - (double)((const float *)Pi)[k];
                s2 += d * d;
            }
            return sqrt(s2);
        };
        const double ci = (p->dtype == QUPS_F64) ? *(const double *)cinv : (double)*(const float *)cinv;
        if (p->I1 > 1 && p->I2 > 1 && ci > 0) {
            hinted.pitch_hint[0] = pitch(1); hinted.pitch_hint[1] = pitch(p->I1); hinted.c_hint = 1.0 / ci;
        }
    }
    p = &hinted;
    const size_t rsz = (p->dtype == QUPS_F64) ? 8 : 4;                       // geometry element
    const size_t csz = (p->dtype == QUPS_F64) ? 16 : (p->dtype == QUPS_F16 ? 4 : 8); // complex data element
    const size_t ysz = (p->dtype == QUPS_F16 && p->y_f32) ? 8 : csz;
    const size_t asz = p->apod_real ? csz / 2 : csz;
    const uint64_t I = p->I1 * p->I2 * p->I3, F = p->F ? p->F : 1;
    const uint64_t On = (p->flag & QUPS_FLAG_KEEP_RX) ? p->N : 1, Om = (p->flag & QUPS_FLAG_KEEP_TX) ? p->M : 1;
    const uint64_t xfs = p->x_frame_stride ? p->x_frame_stride : p->T * p->N * p->M;
    const uint64_t yfs = p->y_frame_stride ? p->y_frame_stride : I * On * Om;
    const size_t xb = csz * ((F - 1) * xfs + p->T * p->N * p->M), yb = ysz * ((F - 1) * yfs + I * On * Om);
    void *dX, *dPi, *dPr, *dPv, *dNv, *dA, *dC, *dY;
    int rc = 0;
    if ((rc = g_ws.get(0, xb, &dX)) || (rc = g_ws.get(1, rsz * 3 * I, &dPi)) || (rc = g_ws.get(2, rsz * 3 * p->N, &dPr)) ||
        (rc = g_ws.get(3, rsz * 4 * p->M, &dPv)) || (rc = g_ws.get(4, rsz * 3 * p->M, &dNv)) ||
        (rc = g_ws.get(5, asz * apod_elems, &dA)) || (rc = g_ws.get(6, rsz * cinv_elems, &dC)) ||
        (rc = p->y_device ? 0 : g_ws.get(7, yb, &dY)))
        return rc;
    if (p->y_device) dY = y;
    cudaStream_t sc = g_ws.s_copy, sx = g_ws.s_comp;
#define QUPS_UP(dst, src, bytes) \
    if (rc == 0 && (bytes) && (e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sc)) != cudaSuccess) rc = cuda_fail(e, "H2D copy");
    QUPS_UP(dPi, Pi, rsz * 3 * I)
    QUPS_UP(dPr, Pr, rsz * 3 * p->N)
    QUPS_UP(dPv, Pv4, rsz * 4 * p->M)
    QUPS_UP(dNv, Nv, rsz * 3 * p->M)
    QUPS_UP(dA, apod, asz * apod_elems)
    QUPS_UP(dC, cinv, rsz * cinv_elems)
    // transmit-chunked pipeline: DAS is a plain sum over transmits (kern/das_spec.m:480), so chunk c+1 is copied
    // while chunk c is beamformed and accumulated.  Only for the contiguous-in-m layout and summed transmits.
    const bool chunkable = rc == 0 && F == 1 && !(p->flag & (QUPS_FLAG_TRANSPOSE | QUPS_FLAG_KEEP_TX)) &&
                           p->S == 0 && !p->accumulate &&
                           ((p->M >= 16 && p->T * p->N * p->M * csz >= (64u << 20)) || (p->host_chunks > 1 && p->M >= (uint64_t)p->host_chunks));
    if (chunkable) {
        bool scalar_c = true;
        for (int d = 0; d < 5; ++d) scalar_c = scalar_c && (!acstride || acstride[d] == 0);
        uint64_t nch = scalar_c ? ((p->M >= 128) ? 16 : 4) : 1;
        if (scalar_c && p->host_chunks > 1) nch = (uint64_t)p->host_chunks < 16 ? (uint64_t)p->host_chunks : 16;
        if (nch > 1) {
            if ((e = cudaEventRecord(g_ws.ev[HostWs::NEV - 1], sc)) != cudaSuccess) rc = cuda_fail(e, "cudaEventRecord");
            if (rc == 0 && (e = cudaStreamWaitEvent(sx, g_ws.ev[HostWs::NEV - 1], 0)) != cudaSuccess) rc = cuda_fail(e, "cudaStreamWaitEvent");
            uint64_t bounds[HostWs::NEV];
            uint64_t nb = 0, pos = 0;
            if (p->host_chunks > 1) {
                for (uint64_t c = 0; c <= nch; ++c) bounds[c] = p->M * c / nch;
                nb = nch;
            } else {
                // the first chunk is what cannot overlap anything (with the 12.6 MB of pixel positions in front of it), so it is
                // short: 8 transmits; each later chunk is 2-3x longer — the copy runs ~3x faster than the beamforming on an idle
                // host (55 GB/s vs 4 transmits/ms at C2), so chunk c + 1 lands before chunk c is done, and still does when 8
                // ranks share the host's memory bandwidth.  (Round 1 started at 32 transmits because short launches lost
                // > 8 % to the per-CTA bounds pass; with das_bounds_kernel and the finer receive split they no longer do.)
                // short transmit shards (a rank of a multi-GPU job: the upload is the contended resource) start smaller still —
                // 8 GPUs, M = 32 per rank, end to end per step: first chunk 32 (no pipeline) 15.2 ms, 16 -> 13.7, 8 -> 13.1, 4 -> 12.4
                double sz = p->M >= 128 ? 8.0 : 4.0;
                const char *eg = getenv("QUPS_B200_CHUNK_GROWTH");
                // measured, C2 through this call on one GPU (device-resident kernel 64.4 ms): chunks 32 x1.3 -> 71.3 ms, 8 x2 -> 69.3,
                // 8 x3 -> 68.5 (fewer launches; every launch recomputes dr per receive tile and has its own tail).  Short
                // transmit shards (one rank of a multi-GPU job, where the copy is the contended resource) keep x2
                const double growth = eg && atof(eg) >= 1.0 ? atof(eg) : (p->M >= 128 ? 3.0 : 2.0);
                if (const char *e0 = getenv("QUPS_B200_CHUNK0")) { const double v = atof(e0); if (v >= 1.0) sz = v; }
                bounds[0] = 0;
                while (pos < p->M && nb < HostWs::NEV - 3) {
                    uint64_t step = (uint64_t)sz;
                    if (p->M - pos < step + step / 2) step = p->M - pos;
                    pos += step;
                    bounds[++nb] = pos;
                    sz *= growth;
                }
                bounds[nb] = p->M;
            }
            // fp16 with half2 output: the running sum over the chunks lives in an fp32 image and is narrowed ONCE at the end
            // (accumulating in y itself would round the image to half precision after every chunk)
            void *dYacc = dY;
            const bool widen_y = p->dtype == QUPS_F16 && !p->y_f32;
            if (widen_y && (rc = g_ws.get(8, sizeof(float2) * I * On * Om, &dYacc))) return rc;
            for (uint64_t c = 0; c < nb && rc == 0; ++c) {
                const uint64_t m0 = bounds[c], m1 = bounds[c + 1];
                const size_t off = csz * p->T * p->N * m0, len = csz * p->T * p->N * (m1 - m0);
                if ((e = cudaMemcpyAsync((char *)dX + off, (const char *)x + off, len, cudaMemcpyHostToDevice, sc)) != cudaSuccess) { rc = cuda_fail(e, "H2D copy"); break; }
                cudaEventRecord(g_ws.ev[c], sc);
                cudaStreamWaitEvent(sx, g_ws.ev[c], 0);
                qups_das_params q = *p;
                q.M = m1 - m0;
                q.accumulate = c > 0;
                if (widen_y) q.y_f32 = 1;
                rc = das_impl(&q, dYacc, dPi, dPr, (char *)dPv + rsz * 4 * m0, (char *)dNv + rsz * 3 * m0, dA, dC, acstride,
                              (char *)dX + off, sx);
            }
            if (rc == 0 && widen_y)
                if (int ce = launch_float2_to_half2((__half2 *)dY, (const float2 *)dYacc, I * On * Om, sx)) rc = cuda_fail(ce, "float2 -> half2");
            if (rc == 0 && yb && !p->y_device && (e = cudaMemcpyAsync(y, dY, yb, cudaMemcpyDeviceToHost, sx)) != cudaSuccess) rc = cuda_fail(e, "D2H copy");
            if ((e = cudaStreamSynchronize(sx)) != cudaSuccess && rc == 0) rc = cuda_fail(e, "cudaStreamSynchronize");
            cudaStreamSynchronize(sc);
            return rc;
        }
    }
    QUPS_UP(dX, x, xb)
#undef QUPS_UP
    if (rc == 0 && (e = cudaEventRecord(g_ws.ev[0], sc)) != cudaSuccess) rc = cuda_fail(e, "cudaEventRecord");
    if (rc == 0 && (e = cudaStreamWaitEvent(sx, g_ws.ev[0], 0)) != cudaSuccess) rc = cuda_fail(e, "cudaStreamWaitEvent");
    if (rc == 0) rc = das_impl(p, dY, dPi, dPr, dPv, dNv, dA, dC, acstride, dX, sx);
    if (rc == 0 && yb && !p->y_device && (e = cudaMemcpyAsync(y, dY, yb, cudaMemcpyDeviceToHost, sx)) != cudaSuccess) rc = cuda_fail(e, "D2H copy");
    if ((e = cudaStreamSynchronize(sx)) != cudaSuccess && rc == 0) rc = cuda_fail(e, "cudaStreamSynchronize");
    cudaStreamSynchronize(sc);
    return rc;
}

int qups_wsinterpd2(const qups_ws2_params *p, void *y, const void *w, const void *x, const void *t1, const void *t2,
                    qups_stream_t stream) {
    g_err[0] = 0;
    if (!p || p->struct_size != sizeof(qups_ws2_params)) return fail(QUPS_ERR_INVALID, "bad qups_ws2_params");
    if (p->D == 0 || p->D > 8) return fail(QUPS_ERR_INVALID, "D must be in 1..8");
    if (p->interp < 0 || p->interp > 3) return fail(QUPS_ERR_INVALID, "Interp option not recognized: %d", p->interp);
    if (int e = launch_wsinterpd2(*p, y, w, x, t1, t2, (cudaStream_t)stream)) {
        if (e == -3) return fail(QUPS_ERR_UNSUPPORTED, "unsupported dtype/layout for wsinterpd2");
        return cuda_fail(e, "wsinterpd2 kernel");
    }
    return 0;
}

int qups_wsinterpd(const qups_ws2_params *p, void *y, const void *w, const void *x, const void *t, qups_stream_t stream) {
    return qups_wsinterpd2(p, y, w, x, t, nullptr, stream);
}

int qups_greens(const qups_greens_params *p, void *y, const void *Pi, const void *a, const void *Pr, const void *Pv,
                const void *kern, qups_stream_t stream) {
    g_err[0] = 0;
    if (!p || p->struct_size != sizeof(qups_greens_params)) return fail(QUPS_ERR_INVALID, "bad qups_greens_params");
    if (p->interp < 0 || p->interp > 3) return fail(QUPS_ERR_INVALID, "Interp option not recognized: %d", p->interp);
    if (!(p->fs > 0) || !(p->fsr > 0) || !(p->c0 > 0)) return fail(QUPS_ERR_INVALID, "fs, fsr and c0 must be positive");
    if (int e = launch_greens(*p, y, Pi, a, Pr, Pv, kern, (cudaStream_t)stream)) {
        if (e == -3) return fail(QUPS_ERR_UNSUPPORTED, "unsupported dtype for greens");
        if (e == -4) return fail(QUPS_ERR_ALLOC, "greens scratch allocation failed");
        return cuda_fail(e, "greens kernel");
    }
    return 0;
}

int qups_convd(const qups_convd_params *p, void *z, const void *x, const void *y, qups_stream_t stream) {
    g_err[0] = 0;
    if (!p || p->struct_size != sizeof(qups_convd_params)) return fail(QUPS_ERR_INVALID, "bad qups_convd_params");
    if (p->shape < 0 || p->shape > 2) return fail(QUPS_ERR_INVALID, "shape must be full(0), same(1) or valid(2)");
    if (!((p->yC == 1 || p->yC == p->C) && (p->yS == 1 || p->yS == p->S))) return fail(QUPS_ERR_INVALID, "A and B must have compatible dimensions");
    if (int e = launch_convd(*p, z, x, y, (cudaStream_t)stream)) {
        if (e == -3) return fail(QUPS_ERR_UNSUPPORTED, "unsupported dtype for convd");
        return cuda_fail(e, "convd kernel");
    }
    return 0;
}

int qups_das_cohfac(const qups_das_params *p, void *y, void *cf, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                    const void *cinv, const void *x, qups_stream_t stream) {
    g_err[0] = 0;
    if (int rc = validate(p, false)) return rc;
    if (p->dtype != QUPS_F32) return fail(QUPS_ERR_UNSUPPORTED, "das_cohfac: dtype must be F32");
    if (p->S != 0 || (p->flag & (QUPS_FLAG_KEEP_RX | QUPS_FLAG_KEEP_TX)) || p->fmod != 0.0 || (p->F > 1) || p->accumulate)
        return fail(QUPS_ERR_UNSUPPORTED, "das_cohfac: plain weights only (S = 0, no kept aperture, fmod = 0, F = 1)");
    const uint64_t I = p->I1 * p->I2 * p->I3;
    if (I == 0) return 0;
    if (!y || !cf || !Pi || !Pr || !Pv4 || !Nv || !cinv || !x) return fail(QUPS_ERR_INVALID, "NULL array argument");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t cst[6] = {0, 0, 0, 0, 0, 0};
    DasArgs<float> a;
    fill_args<float>(a, p, Pi, Pr, Pv4, Nv, nullptr, cinv, cst, 1, nullptr);
    a.x = x; a.y = y;
    a.cohfac = 1; a.cf = (float *)cf;
    if (a.N != 0 && a.M != 0 && p->path != QUPS_PATH_GENERIC && das_tiled_plan(a, 0, 0).eligible) {
        g_last_das = "das_tiled";
        if (int e = launch_das_tiled(a, st)) return cuda_fail(e, "das_tiled (coherence mode)");
        return 0;
    }
    // outside the staged envelope: the reference's own sequence — keep_rx cube, cohfac along the receive dimension, sum
    float2 *cube = nullptr;
    cudaError_t ce = ws_alloc((void **)&cube, sizeof(float2) * I * (p->N ? p->N : 1), st);
    if (ce != cudaSuccess) return fail(QUPS_ERR_ALLOC, "das_cohfac: ws_alloc(%llu): %s", (unsigned long long)(sizeof(float2) * I * p->N), cudaGetErrorString(ce));
    qups_das_params q = *p;
    q.flag |= QUPS_FLAG_KEEP_RX;
    int rc = das_impl(&q, cube, Pi, Pr, Pv4, Nv, nullptr, cinv, cst, x, st);
    if (rc == 0) {
        qups_aperture_params ap{};
        ap.struct_size = sizeof(ap); ap.dtype = QUPS_F32; ap.op = QUPS_APD_COHFAC; ap.C = I; ap.A = p->N; ap.S = 1; ap.gamma = 1.0;
        if (int e = launch_aperture(ap, cf, nullptr, cube, nullptr, st)) rc = cuda_fail(e, "cohfac");
    }
    if (rc == 0) rc = das_impl(p, y, Pi, Pr, Pv4, Nv, nullptr, cinv, cst, x, st);
    ws_free(cube, st);
    return rc;
}

int qups_pwznxcorr(const qups_xcorr_params *p, void *y, const void *x, const void *x0, const void *w, const int32_t *lags,
                   qups_stream_t stream) {
    g_err[0] = 0;
    if (!p || p->struct_size != sizeof(qups_xcorr_params)) return fail(QUPS_ERR_INVALID, "bad qups_xcorr_params");
    if (p->dtype != QUPS_F32 && p->dtype != QUPS_F64) return fail(QUPS_ERR_UNSUPPORTED, "pwznxcorr: dtype must be F32 or F64");
    if (p->ref < 0 || p->ref > 2) return fail(QUPS_ERR_INVALID, "pwznxcorr: ref must be neighbor(0), center(1) or x0(2)");
    if (p->L == 0 || p->W == 0 || p->T == 0 || p->N == 0 || p->F == 0) return fail(QUPS_ERR_INVALID, "pwznxcorr: empty input");
    if (p->ref == QUPS_XC_NEIGHBOR && (p->stride == 0 || p->stride >= p->N)) return fail(QUPS_ERR_INVALID, "pwznxcorr: stride must be in [1, N)");
    if (p->ref == QUPS_XC_X0 && (!x0 || !((p->x0N == 1 || p->x0N == p->N) && (p->x0F == 1 || p->x0F == p->F))))
        return fail(QUPS_ERR_INVALID, "pwznxcorr: x0 must be T x {1|N} x {1|F}");
    if (p->T >= (1ull << 30) || p->N > 65535 || p->F > 65535) return fail(QUPS_ERR_UNSUPPORTED, "pwznxcorr: T < 2^30, N and F <= 65535");
    if (xcorr_smem_bytes(p->W, p->dtype == QUPS_F64) > 200 * 1024) return fail(QUPS_ERR_UNSUPPORTED, "pwznxcorr: window too long for the shared-memory tile");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t amax = 0;
    for (uint32_t l = 0; l < p->L; ++l) { const int64_t v = lags[l] < 0 ? -(int64_t)lags[l] : lags[l]; if (v > amax) amax = v; }
    const uint32_t P = p->pad ? (uint32_t)amax : 0u; // kern/pwznxcorr.m:182
    int32_t *dl = nullptr;
    if (cudaError_t ce = ws_alloc((void **)&dl, sizeof(int32_t) * p->L, st)) return cuda_fail((int)ce, "pwznxcorr: lag table");
    cudaError_t ce = cudaMemcpyAsync(dl, lags, sizeof(int32_t) * p->L, cudaMemcpyHostToDevice, st);
    int e = ce ? (int)ce : launch_pwznxcorr(p->dtype == QUPS_F64, y, x, x0, w, dl, (uint32_t)p->T, P, (uint32_t)p->N, (uint32_t)p->F, p->L, p->W,
                                             p->ref, p->zero != 0, p->norm != 0, p->is_complex != 0, p->stride, (uint32_t)p->x0N, (uint32_t)p->x0F, st);
    // the lag table is pageable host memory: the copy above has been staged by the driver when cudaMemcpyAsync returns
    ws_free(dl, st);
    if (e) return cuda_fail(e, "pwznxcorr kernel");
    return 0;
}

int qups_refocus(const qups_refocus_params *p, void *y, const void *x, const void *Hi, const double *t0, double *t0_out,
                 qups_stream_t stream) {
    g_err[0] = 0;
    if (!p || p->struct_size != sizeof(qups_refocus_params)) return fail(QUPS_ERR_INVALID, "bad qups_refocus_params");
    if (p->dtype != QUPS_F32) return fail(QUPS_ERR_UNSUPPORTED, "refocus: dtype must be F32");
    if (!t0 || !(p->n_t0 == 1 || p->n_t0 == p->V)) return fail(QUPS_ERR_INVALID, "refocus: t0 must hold 1 or V start times");
    if (!(p->fs > 0)) return fail(QUPS_ERR_INVALID, "refocus: fs must be positive");
    double tmin = t0[0];
    for (uint32_t v = 1; v < p->n_t0; ++v) tmin = t0[v] < tmin ? t0[v] : tmin;
    std::vector<double> dt(p->V ? p->V : 1, 0.0);
    for (uint64_t v = 0; v < p->V; ++v) dt[v] = t0[p->n_t0 == 1 ? 0 : v] - tmin;
    if (t0_out) *t0_out = tmin;
    if (int e = launch_refocus(y, x, Hi, dt.data(), p->T, p->N, p->V, p->E, p->fs, (cudaStream_t)stream)) {
        if (e == -3) return fail(QUPS_ERR_UNSUPPORTED, "refocus: T must be a power of two <= 8192 (zero-pad the data)");
        return cuda_fail(e, "refocus kernels");
    }
    return 0;
}

} // extern "C"
