// das_tiled.cu — the hot DAS kernel for sm_100a (B200): fp32 complex data, scalar sound speed (the modes beyond the plain
// sum over both apertures are listed below).
//
// What it replaces: the serial M x N loop of src/bf.cu:96-139 (one thread per
// pixel, two sqrt + 64-bit stride math + 1/2/4 dependent global gathers per
// (pixel,rx,tx) pair).  Design:
//
//   * CTA = 32 x 16 pixel tile (lanes along the slow image axis, where delays
//     vary slowly => few shared-memory wavefronts per gather), 8 consumer warps
//     (2 pixel rows per thread) + 1 producer warp.
//   * loop nest: receive tile (16 traces) outer, transmit m inner.  dr(i,n) for
//     the 16 receives lives in registers for all M transmits; dv(i,m) is
//     recomputed once per (pixel, m, receive tile)  (1 sqrt per 16 pairs).
//   * the producer warp derives, per (n,m) trace, the exact range of sample
//     indices the tile can touch from per-tile min/max of dv and dr — every
//     operation of the delay sequence is monotone and individually rounded, so
//     [k(dvmin,drmin), k(dvmax,drmax)] is a rigorous bound in floating point —
//     and stages just that window of the trace into shared memory with one
//     cp.async.bulk (UBLKCP) per trace, completion on an mbarrier
//     (4-stage full/empty ring).
//   * consumers gather taps with 64-bit LDS at 32-bit shared addresses; no
//     bounds checks in the inner loop (the window is proven to cover them).
//   * traces whose window leaves [first+1, last-1] of the trace, does not fit
//     the slot, or has a NaN bound take a slow path with the full interp1 edge
//     semantics straight from global memory; traces entirely outside the data
//     are skipped (they contribute exactly 0).
//
//   * work decomposition: grid = tiles x nsplit (receive-axis split, partial images summed in split order by
//     das_reduce_kernel: deterministic).  For split launches the per-tile path-length bounds come from das_bounds_kernel
//     (once per tile instead of once per CTA), the split is the smallest divisor of the receive-tile count giving ~48
//     waves of CTAs, and the CTAs are ordered split-major so that the CTAs resident at one time read windows of the same
//     1/nsplit of the cube (DRAM reads at C2: 2.6 GB per launch, tile-major 22.6 GB).
//   * modes (template parameters): real apodization arrays (NAP), closed-form apodization (FUSED), kept apertures MUL / SYN
//     (KEEP = 1 / 2), table-driven delays for bfDAS (LUT), and the coherence mode (KEEP = 3: DAS image + coherence factor of
//     the per-receive images in one pass, 8 receives x 2 transmits per stage).
//
// Numerics: the sample position xq uses the canonical individually-rounded
// sequence (common.cuh), so tap indices are bit-identical to the oracle; the
// interpolation weights / accumulation use FMAs (tolerance-level difference).
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include "das_args.cuh"

namespace qups {

#ifndef QUPS_NT
#define QUPS_NT 16
#endif
#ifndef QUPS_CW
#define QUPS_CW 8
#endif
#ifndef QUPS_STAGES
#define QUPS_STAGES 4
#endif
#ifndef QUPS_MINBLOCKS
#define QUPS_MINBLOCKS 2
#endif
#ifndef QUPS_WMAX
#define QUPS_WMAX 128
#endif
constexpr int kNT = QUPS_NT;          // traces (receives) per stage
constexpr int kR = 2;                 // pixel rows per thread
constexpr int kCW = QUPS_CW;          // consumer warps
constexpr int kStages = QUPS_STAGES;  // smem ring depth
#ifndef QUPS_PW
#define QUPS_PW 2
#endif
constexpr int kPW = QUPS_PW;          // producer warps: they take the published stages in turn (stage index mod kPW)
// threads per CTA: the 1-tap (nearest) kernel runs kPW producer warps, linear / cubic one (a tenth warp cost them 2 %)
constexpr int threads_of(int interp) { return (kCW + (interp == 0 ? kPW : 1)) * 32; }
#ifndef QUPS_STATS
#define QUPS_STATS 0
#endif
#ifndef QUPS_MAGIC
#define QUPS_MAGIC 1
#endif
#ifndef QUPS_W9
#define QUPS_W9 1
#endif
#ifndef QUPS_ACC2
#define QUPS_ACC2 1
#endif
#ifndef QUPS_LPA
#define QUPS_LPA 8
#endif
#ifndef QUPS_WAITHINT
#define QUPS_WAITHINT 0   // mbarrier.try_wait with a suspend-time hint (measured on C2: 68.05 vs 67.69 ms without — no gain, off)
#endif
#ifndef QUPS_SOFF4
#define QUPS_SOFF4 1      // all-fast stages: slot offsets of 4 traces per LDS.128 from a compact table (12 fewer LDS per stage)
#endif
#ifndef QUPS_HDRGEO
#define QUPS_HDRGEO 1     // the producer puts the stage's transmit geometry (Pv, t0, Nv, uniform sign of dv) into the stage header
#endif
constexpr int kBarBytes = ((2 * kStages * 8 + 63) / 64) * 64;  // full[] + empty[] mbarriers
// per-stage header block: [0] int4 (kind, outer index, inner tile, uniform sign of dv: +1 / -1 / 0 = mixed) | [1] float4 geometry A
// (Pv.xyz, t0 of transmit `outer`; Pr.xyz when the inner traces are transmits) | [2] float4 geometry B (Nv.xyz) | [3], [4] the
// same two for the SECOND transmit of a coherence-mode stage (KEEP == 3: 8 receives x 2 transmits; the sign word then holds
// (sgn0 + 1) | (sgn1 + 1) << 2) | then the compact slot-offset table of all-fast stages (kNT x uint32)
constexpr int kHdrTab = 80;
constexpr int kHdrBytes = kHdrTab + 4 * kNT;
constexpr int kTilePix = kCW * 32 * kR; // pixels per tile; its SHAPE (tA x tB, lane patch lpa) is chosen per call from the pixel spacing

// QUPS_MAGIC: the cubic fast path derives the tap address from the bits of 2^23 + floor(xq); the constant
// (0x4B000000 << 3) mod 2^32 is folded into the published slot offset
template <int INTERP> struct magic_off { static constexpr uint32_t value = QUPS_MAGIC ? (0x4B000000u << 3) : 0u; };

#if QUPS_STATS
__device__ unsigned long long g_stats[12]; // [0..3] traces by flag, [4] split traces, [5] all-fast stages, [6] general stages, [7] empty, [8] general stages made of single-window FAST + SKIP traces only, [9] general stages with a dual-window trace, [10] with an EDGE trace, [11] with a SLOW trace
#endif
enum { TR_FAST = 0, TR_SKIP = 1, TR_SLOW = 2, TR_EDGE = 3 };   // per (n,m) trace
enum { ST_MIXED = 0, ST_ALL_FAST = 1, ST_END = 2 };             // per published stage of kNT traces

struct TiledArgs {
    const float *Pi, *Pr, *Pv4, *Nv, *cinv;
    const float2 *x;
    float2 *y;
    uint32_t N, M, T;
    uint32_t IA, IB, IC;    // extents: lane axis, row axis, slice axis
    uint64_t sA, sB, sC;    // pixel-index strides of those axes
    uint32_t tilesA, tilesB;
    uint32_t tA, tB, lpa;   // tile extents along the lane / row axes (tA * tB = kTilePix) and lanes of a warp along the lane axis
    uint32_t wmax;          // samples per smem slot (even)
    uint32_t stages;        // ring depth in use (2 .. kStages): fewer, longer slots on coarse pixel grids
    uint32_t numNT;
    uint32_t nsplit;        // receive-axis split: CTA (tile, split) handles a contiguous range of receive tiles
    float2 *part;           // nsplit > 1: partial images [nsplit][I], summed in order by das_reduce_kernel
    uint64_t I;
    int rev;                // launch the tiles in reverse order (largest I1 first)
    // real apodization arrays (NAP = 1 or 2): a_s[base + i1*st0 + i2*st1 + i3*st2 + n*st3 + m*st4], 32-bit indices
    const float *ap[2];
    uint32_t ast[2][5];
    uint32_t I1, I2;
    float fs;
    int VS, DV, tpose, accumulate;
    uint64_t total_elems;   // T*N*M
    FusedApod fa;           // closed-form apodization evaluated in-kernel (FUSED > 0)
    uint32_t I3;
    // LUT mode (bfDAS / bfDASLUT / sample2sep -> wsinterpd2 with separable delay tables, src/interpd.cu:344-396): the path
    // lengths come from tables in SAMPLES instead of geometry, xq = 1 + (tm(i,m) + tn(i,n))   (kern/wsinterpd2.m:290)
    const float *tn, *tm;   // tn[n * I + i] (the aperture whose traces are contiguous in x), tm[m * I + i]
    const float *wscal;     // scalar weight applied to the sum (real, or complex when wcplx)
    int wcplx;
    // receive-split launches: per-tile path-length bounds computed ONCE by das_bounds_kernel instead of once per (tile, split)
    // CTA — [tile][4 M + 2 N] ordered ints: dv < 0 cluster min / max [M], dv >= 0 cluster min / max [M], dr min / max [N]
    int *bounds;
    int split_major;
    float *partP;           // KEEP == 3: partial sums of |b_n|^2 [nsplit][I] next to the partial images in `part`
};

// ---- small PTX wrappers -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if QUPS_WAITHINT
    // potentially-blocking wait with a suspend-time hint: the warp is parked by the hardware until the phase completes (or the
    // hint expires) instead of re-issuing try_wait (ncu, round 1: 1.5e9 SYNCS + as many BRA = 6 % of all issued instructions)
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "QUPS_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra QUPS_DONE_%=;\n\t"
        "bra QUPS_WAIT_%=;\n\t"
        "QUPS_DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
#else
    while (!mbar_try_wait(bar, parity)) {}
#endif
}
// global -> shared bulk async copy (TMA engine, SASS UBLKCP), completes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
// order-preserving float <-> int map (involution)
__device__ __forceinline__ int f2o(float f) {
    const int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float o2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }

// ---- inner bodies ---------------------------------------------------------------
// xq = 1-based sample position (bit-identical to the oracle); soff = shared
// address such that the first tap of index k = floor(xq) is at soff + 8*k.
__device__ __forceinline__ void cubic_weights(float u, float &w0, float &w1, float &w2, float &w3) {
    // Keys cubic convolution a = -1/2 (interior identical to interp1 'cubic', R2020b+), 1/2 folded in
    const float u2 = u * u;
    w0 = fmaf(fmaf(-0.5f, u, 1.0f), u, -0.5f) * u;
    w1 = fmaf(fmaf(1.5f, u, -2.5f), u2, 1.0f);
    w2 = fmaf(fmaf(-1.5f, u, 2.0f), u, 0.5f) * u;
    w3 = fmaf(0.5f, u, -0.5f) * u2;
}
template <int INTERP> __device__ __forceinline__ void fast_pair(float xq, uint32_t soff, float &ar, float &ai) {
    if (INTERP == 2) {
        const float kf = floorf(xq);
        const float u = xq - kf; // exact
        const uint32_t addr = soff + ((uint32_t)__float2int_rz(kf) << 3);
        const float2 v0 = lds64(addr), v1 = lds64(addr + 8), v2 = lds64(addr + 16), v3 = lds64(addr + 24);
        float w0, w1, w2, w3;
        cubic_weights(u, w0, w1, w2, w3);
        ar = fmaf(w0, v0.x, ar); ai = fmaf(w0, v0.y, ai);
        ar = fmaf(w1, v1.x, ar); ai = fmaf(w1, v1.y, ai);
        ar = fmaf(w2, v2.x, ar); ai = fmaf(w2, v2.y, ai);
        ar = fmaf(w3, v3.x, ar); ai = fmaf(w3, v3.y, ai);
    } else if (INTERP == 1) {
        const float kf = floorf(xq);
        const float u = xq - kf;
        const uint32_t addr = soff + ((uint32_t)__float2int_rz(kf) << 3);
        const float2 v0 = lds64(addr), v1 = lds64(addr + 8);
        ar += fmaf(u, v1.x - v0.x, v0.x);
        ai += fmaf(u, v1.y - v0.y, v0.y);
    } else {
        // round half away from zero == floor(xq + 0.5) exactly for xq >= 1
        const float kf = floorf(__fadd_rn(xq, 0.5f));
        const uint32_t addr = soff + ((uint32_t)__float2int_rz(kf) << 3);
        const float2 v0 = lds64(addr);
        ar += v0.x;
        ai += v0.y;
    }
}

// Two pixel rows of one thread against one staged trace.  The delay sequence stays SCALAR with
// explicit _rn intrinsics: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (observed in SASS),
// which would break bit-identity of xq with the oracle; scalar mul.rn/add.rn are never contracted.
// The interpolation weights (tolerance-level arithmetic) use packed fp32x2 FFMA2/FMUL2 (sm_100).
struct Pack2 {
    float2 dv;
    float cinv, t0, fs;
};
template <int INTERP, int LUT = 0>
__device__ __forceinline__ void fast_pair2(const Pack2 &c, float2 dr, uint32_t soff, uint32_t soff1, float2 &acc0, float2 &acc1) {
    float2 xq;
    if (LUT) { // tables already in samples: 1 + (t_m + t_n), two individually rounded adds (identical to the general form
               // with cinv = 1, t0 = 0, fs = 1, whose extra operations are exact)
        xq.x = __fadd_rn(1.0f, __fadd_rn(c.dv.x, dr.x));
        xq.y = __fadd_rn(1.0f, __fadd_rn(c.dv.y, dr.y));
    } else {
        xq.x = sample_pos(c.dv.x, dr.x, c.cinv, c.t0, c.fs);
        xq.y = sample_pos(c.dv.y, dr.y, c.cinv, c.t0, c.fs);
    }
    if (INTERP == 2) {
#if QUPS_MAGIC
        // floor + float->int + address in one go: for 0 <= xq < 2^23, add.rm(xq, 2^23) = 2^23 + floor(xq) exactly
        // (ulp = 1 in [2^23, 2^24), round toward -inf), so its low mantissa bits ARE the index: no FRND / F2I (both on
        // the quarter-rate XU pipe).  (bits << 3) wraps mod 2^32; the constant part is folded into the base address.
        const float2 tm = make_float2(__fadd_rd(xq.x, 8388608.f), __fadd_rd(xq.y, 8388608.f));
        const float2 kf = __fadd2_rn(tm, make_float2(-8388608.f, -8388608.f)); // exact
        const float2 u = __ffma2_rn(kf, make_float2(-1.f, -1.f), xq);          // exact
        // (the producer pre-subtracts kMagicOff from the slot offset it publishes for all-fast stages)
        const uint32_t a0 = soff + (__float_as_uint(tm.x) << 3);
        const uint32_t a1 = soff1 + (__float_as_uint(tm.y) << 3);
#else
        const float2 kf = make_float2(floorf(xq.x), floorf(xq.y));
        const float2 u = __ffma2_rn(kf, make_float2(-1.f, -1.f), xq); // exact
        const uint32_t a0 = soff + ((uint32_t)__float2int_rz(kf.x) << 3);
        const uint32_t a1 = soff1 + ((uint32_t)__float2int_rz(kf.y) << 3);
#endif
        const float2 p0 = lds64(a0), p1 = lds64(a0 + 8), p2 = lds64(a0 + 16), p3 = lds64(a0 + 24);
        const float2 q0 = lds64(a1), q1 = lds64(a1 + 8), q2 = lds64(a1 + 16), q3 = lds64(a1 + 24);
#if QUPS_W9
        // Keys weights through the partition-of-unity / linear-reproduction identities (9 packed ops instead of 11):
        // w0 = -u(1-u)^2/2, w3 = -u^2(1-u)/2, w1 = (1-u) - 2 w0 + w3, w2 = u + w0 - 2 w3
        const float2 one2 = make_float2(1.f, 1.f), m2 = make_float2(-2.f, -2.f);
        const float2 v = __ffma2_rn(u, make_float2(-1.f, -1.f), one2);
        const float2 h = __fmul2_rn(__fmul2_rn(u, v), make_float2(-0.5f, -0.5f));
        const float2 w0 = __fmul2_rn(h, v);
        const float2 w3 = __fmul2_rn(h, u);
        const float2 w1 = __ffma2_rn(m2, w0, __fadd2_rn(v, w3));
        const float2 w2 = __ffma2_rn(m2, w3, __fadd2_rn(u, w0));
#else
        const float2 u2 = __fmul2_rn(u, u);
        const float2 w0 = __fmul2_rn(__ffma2_rn(__ffma2_rn(make_float2(-0.5f, -0.5f), u, make_float2(1.f, 1.f)), u,
                                                make_float2(-0.5f, -0.5f)), u);
        const float2 w1 = __ffma2_rn(__ffma2_rn(make_float2(1.5f, 1.5f), u, make_float2(-2.5f, -2.5f)), u2,
                                     make_float2(1.f, 1.f));
        const float2 w2 = __fmul2_rn(__ffma2_rn(__ffma2_rn(make_float2(-1.5f, -1.5f), u, make_float2(2.f, 2.f)), u,
                                                make_float2(0.5f, 0.5f)), u);
        const float2 w3 = __fmul2_rn(__ffma2_rn(make_float2(0.5f, 0.5f), u, make_float2(-0.5f, -0.5f)), u2);
#endif
#if QUPS_ACC2
        // complex accumulate as ONE packed FFMA2 per tap: (re,im) += w * (v.re, v.im), the weight broadcast
        acc0 = __ffma2_rn(p0, make_float2(w0.x, w0.x), acc0);
        acc1 = __ffma2_rn(q0, make_float2(w0.y, w0.y), acc1);
        acc0 = __ffma2_rn(p1, make_float2(w1.x, w1.x), acc0);
        acc1 = __ffma2_rn(q1, make_float2(w1.y, w1.y), acc1);
        acc0 = __ffma2_rn(p2, make_float2(w2.x, w2.x), acc0);
        acc1 = __ffma2_rn(q2, make_float2(w2.y, w2.y), acc1);
        acc0 = __ffma2_rn(p3, make_float2(w3.x, w3.x), acc0);
        acc1 = __ffma2_rn(q3, make_float2(w3.y, w3.y), acc1);
#else
        acc0.x = fmaf(w0.x, p0.x, acc0.x); acc0.y = fmaf(w0.x, p0.y, acc0.y);
        acc1.x = fmaf(w0.y, q0.x, acc1.x); acc1.y = fmaf(w0.y, q0.y, acc1.y);
        acc0.x = fmaf(w1.x, p1.x, acc0.x); acc0.y = fmaf(w1.x, p1.y, acc0.y);
        acc1.x = fmaf(w1.y, q1.x, acc1.x); acc1.y = fmaf(w1.y, q1.y, acc1.y);
        acc0.x = fmaf(w2.x, p2.x, acc0.x); acc0.y = fmaf(w2.x, p2.y, acc0.y);
        acc1.x = fmaf(w2.y, q2.x, acc1.x); acc1.y = fmaf(w2.y, q2.y, acc1.y);
        acc0.x = fmaf(w3.x, p3.x, acc0.x); acc0.y = fmaf(w3.x, p3.y, acc0.y);
        acc1.x = fmaf(w3.y, q3.x, acc1.x); acc1.y = fmaf(w3.y, q3.y, acc1.y);
#endif
    } else {
#if QUPS_MAGIC
        // linear / nearest with the same magic-number index (published offsets carry -kMagicOff for every interpolator)
        const float2 xr = (INTERP == 1) ? xq : make_float2(__fadd_rn(xq.x, 0.5f), __fadd_rn(xq.y, 0.5f)); // round half away == floor(xq + .5)
        const float2 tm = make_float2(__fadd_rd(xr.x, 8388608.f), __fadd_rd(xr.y, 8388608.f));
        const uint32_t a0 = soff + (__float_as_uint(tm.x) << 3);
        const uint32_t a1 = soff1 + (__float_as_uint(tm.y) << 3);
        if (INTERP == 1) {
            const float2 kf = __fadd2_rn(tm, make_float2(-8388608.f, -8388608.f)); // exact
            const float2 u = __ffma2_rn(kf, make_float2(-1.f, -1.f), xq);          // exact
            const float2 p0 = lds64(a0), p1 = lds64(a0 + 8), q0 = lds64(a1), q1 = lds64(a1 + 8);
            // v0 + u (v1 - v0), (re, im) packed: same three roundings as the scalar form
            const float2 d0 = __ffma2_rn(p0, make_float2(-1.f, -1.f), p1), d1 = __ffma2_rn(q0, make_float2(-1.f, -1.f), q1);
            acc0 = __fadd2_rn(acc0, __ffma2_rn(d0, make_float2(u.x, u.x), p0));
            acc1 = __fadd2_rn(acc1, __ffma2_rn(d1, make_float2(u.y, u.y), q0));
        } else {
            acc0 = __fadd2_rn(acc0, lds64(a0));
            acc1 = __fadd2_rn(acc1, lds64(a1));
        }
#else
        fast_pair<INTERP>(xq.x, soff, acc0.x, acc0.y);
        fast_pair<INTERP>(xq.y, soff1, acc1.x, acc1.y);
#endif
    }
}

// full-semantics slow path (edges, oversize windows, NaN): straight from global memory
__device__ __noinline__ void slow_pair(const float2 *trace, uint32_t T, float xq, int interp, float &ar, float &ai) {
    const cplx<float> v = interp1<float2>(trace, (long)T, xq, interp);
    ar += v.re;
    ai += v.im;
}

// may the unchecked gather be used for this sample position?
template <int INTERP> __device__ __forceinline__ bool interior(float xq, float Tf) {
    if (INTERP == 2) return (xq >= 2.0f) && (xq < Tf - 1.0f); // taps k-1..k+2 are real samples
    if (INTERP == 1) return (xq >= 1.0f) && (xq < Tf);
    return (xq >= 1.0f) && (xq <= Tf);
}

// One (pixel, trace) pair of an EDGE / FAST trace in a mixed stage.  The staged window holds every tap an
// in-range pixel can touch, including the first / last three samples the interp1 end padding needs
// (v(0) = 3v(1)-3v(2)+v(3), v(T+1) = 3v(T)-3v(T-1)+v(T-2)), so the trace ends are served from shared
// memory too; out-of-range pixels contribute exactly 0 (extrapval).
template <int INTERP>
__device__ __forceinline__ void edge_pair(float xq, uint32_t soff, float Tf, int T, float &ar, float &ai) {
    if (interior<INTERP>(xq, Tf)) { fast_pair<INTERP>(xq, soff, ar, ai); return; }
    if (!(xq >= 1.0f && xq <= Tf)) return;
    if (INTERP == 2) {
        int k = (int)floorf(xq);
        k = min(k, T - 1);
        const float u = xq - (float)k;
        float w[4];
        cubic_weights(u, w[0], w[1], w[2], w[3]);
        // 1-based sample q lives at soff + 8*(q+1) (first tap of floor index k is q = k-1 at soff + 8k)
        const uint32_t e0 = (k <= 1) ? 1u : (uint32_t)T; // padded end: samples e0, e0 +- 1, e0 +- 2
        const int dir = (k <= 1) ? 1 : -1;
        const float2 a = lds64(soff + 8u * (e0 + 1u)), b = lds64(soff + 8u * (uint32_t)((int)e0 + dir + 1)),
                     c = lds64(soff + 8u * (uint32_t)((int)e0 + 2 * dir + 1));
        const float2 pad = make_float2(fmaf(3.f, a.x, fmaf(-3.f, b.x, c.x)), fmaf(3.f, a.y, fmaf(-3.f, b.y, c.y)));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = k - 1 + j;
            const float2 v = (q >= 1 && q <= T) ? lds64(soff + 8u * (uint32_t)(q + 1)) : pad;
            ar = fmaf(w[j], v.x, ar);
            ai = fmaf(w[j], v.y, ai);
        }
    } else if (INTERP == 1) { // only xq == T lands here: k clamps to T-1, s = 1  (first tap of index k at soff + 8k)
        const float2 v0 = lds64(soff + 8u * (uint32_t)(T - 1)), v1 = lds64(soff + 8u * (uint32_t)T);
        ar += fmaf(1.0f, v1.x - v0.x, v0.x);
        ai += fmaf(1.0f, v1.y - v0.y, v0.y);
    }
}

// EDGE / SLOW traces of a general stage, both pixel rows of a thread (rare: kept out of line so the unrolled
// stage body stays small).  so0/so1 are the published (magic-adjusted) slot offsets of the two pixels' clusters.
template <int INTERP>
__device__ __noinline__ void rare_pair2(const float2 *trace, uint32_t T, int flag, float xq0, float xq1, uint32_t so0,
                                        uint32_t so1, float2 &t0, float2 &t1) {
    if (flag != TR_SLOW) { // EDGE (or FAST next to one): everything comes from the staged window(s)
        edge_pair<INTERP>(xq0, so0 + magic_off<INTERP>::value, (float)T, (int)T, t0.x, t0.y);
        edge_pair<INTERP>(xq1, so1 + magic_off<INTERP>::value, (float)T, (int)T, t1.x, t1.y);
    } else {               // window does not fit the slot / NaN bound: full interp1 from global memory
        slow_pair(trace, T, xq0, INTERP, t0.x, t0.y);
        slow_pair(trace, T, xq1, INTERP, t1.x, t1.y);
    }
}

// sample-index window [t_lo, t_hi] (0-based taps) that covers every pixel whose position lies in [xlo, xhi];
// returns 0 if the whole range is outside the trace (contributes exactly 0), 1 otherwise; inr = window interior
// to the trace (unchecked gather allowed), else clipped to the trace incl. the end samples interp1's padding needs
template <int INTERP>
__device__ __forceinline__ int tap_window(float xlo, float xhi, float Tf, int T, int &t_lo, int &t_hi, bool &inr, int &tap0) {
    if (xhi < 1.0f || xlo > Tf) return 0;
    // every delay operation is monotone and individually rounded, so all tap indices lie in [k(xlo), k(xhi)]
    inr = interior<INTERP>(xlo, Tf) && interior<INTERP>(xhi, Tf);
    const float xl = fmaxf(xlo, 1.0f), xh = fminf(xhi, Tf);
    int klo, khi, tap1;
    if (INTERP == 2)      { klo = (int)floorf(xl); khi = (int)floorf(xh); tap0 = 2; tap1 = 1; }
    else if (INTERP == 1) { klo = (int)floorf(xl); khi = (int)floorf(xh); tap0 = 1; tap1 = 0; }
    else { klo = (int)floorf(__fadd_rn(xl, 0.5f)); khi = (int)floorf(__fadd_rn(xh, 0.5f)); tap0 = 1; tap1 = -1; }
    t_lo = klo - tap0; t_hi = khi + tap1;
    if (!inr) {
        // EDGE: keep the three end samples the interp1 padding needs (edge_pair), then clip
        if (INTERP == 2) {
            if (klo <= 1) t_hi = max(t_hi, 2);
            if (khi >= T - 1) t_lo = min(t_lo, T - 3);
        } else if (INTERP == 1) {
            if (khi >= T) t_lo = min(t_lo, T - 2);
        }
        t_lo = max(t_lo, 0);
        t_hi = min(t_hi, T - 1);
    }
    return 1;
}

// FUSED: 0 = none; 1 = closed-form TRANSMIT weights only (one weight per pixel and stage, folded into the stage sum);
//        2 = closed-form RECEIVE weights (per-thread table in shared memory, refreshed per receive tile) + optional transmit weights
// KEEP: 0 = sum both apertures (DAS); 1 = keep the transmit dimension (MUL, y is I x M): every stage (receive tile, m) adds
//        its sum to y(:,m); 2 = keep the receive dimension (SYN, y is I x N): the ROLES of the apertures are swapped — the 16
//        traces of a stage are 16 transmits of ONE receive, the registers hold dv(i, m) of a transmit tile and the stage
//        scalar is dr(i, n) — so every stage adds to y(:,n).  (sample_pos only adds dv + dr: the swap is bit-neutral.)
//        y is pre-zeroed by the launcher; a CTA owns its pixels (nsplit = 1), so the read-modify-write needs no atomics.
//        3 = COHERENCE mode: DAS(..., 'keep_rx') -> cohfac / sum over the receive dimension WITHOUT the I x N cube
//        (kern/cohfac.m: |sum_n b_n|^2 / sum_n |b_n|^2 / N needs every per-receive sum b_n = sum_m complete).  A stage holds
//        8 receives x 2 consecutive transmits, so the 8 x 2 per-receive accumulators of a thread fit next to 8 (not 16) receive
//        path lengths; after the last transmit of a receive group S += b_n, P += |b_n|^2.  Outputs S (the DAS image) and P go
//        to the partial buffers; das_cf_reduce_kernel sums the receive splits and forms the factor.
template <int INTERP, int NAP, int FUSED, int KEEP, int LUT = 0>
__global__ void __launch_bounds__(threads_of(INTERP), QUPS_MINBLOCKS) das_tiled_kernel(const TiledArgs a) {
    constexpr int kThreads = threads_of(INTERP);
    constexpr bool kInnerTx = (KEEP == 2); // the 16 traces of a stage run over transmits instead of receives
    constexpr bool kCF = (KEEP == 3);      // coherence mode: trace j of a stage = receive (j & 7) of the group, transmit 2 outer + (j >> 3)
    constexpr int kRX = kCF ? 8 : kNT;     // receives per inner tile
    static_assert(KEEP == 0 || (NAP == 0 && FUSED == 0), "kept apertures: plain weights only");
    static_assert(LUT == 0 || (NAP == 0 && FUSED == 0 && KEEP == 0), "table-driven delays: plain sum over both apertures");
    static_assert(kR == 2, "the packed fp32x2 inner loop assumes two pixel rows per thread");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [0,64) full/empty mbarriers | stage_hdr[kStages] int4 | desc[kStages][kNT] int4 |
    //         (stage header blocks: see kHdrBytes)  dv cluster bounds: nmin[M] nmax[M] pmin[M] pmax[M] | drmin[N] drmax[N] (ordered ints) | 128B-aligned ring
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *stage_blk = smem_raw + kBarBytes;   // kStages blocks of kHdrBytes
    auto stage_hdr = [&](uint32_t st_) -> int4 * { return reinterpret_cast<int4 *>(stage_blk + st_ * kHdrBytes); };
    int4 *desc = reinterpret_cast<int4 *>(smem_raw + kBarBytes + kHdrBytes * kStages);
    int *s_dvnmin = reinterpret_cast<int *>(smem_raw + kBarBytes + kHdrBytes * kStages + sizeof(int4) * kStages * kNT);
    int *s_dvnmax = s_dvnmin + a.M;
    int *s_dvpmin = s_dvnmax + a.M;
    int *s_dvpmax = s_dvpmin + a.M;
    int *s_drmin = s_dvpmax + a.M;
    int *s_drmax = s_drmin + a.N;
    int *s_txany = s_drmax + a.N;   // FUSED: does any pixel of the tile have a non-zero transmit / receive weight?
    int *s_rxany = s_txany + (FUSED ? a.M : 0);
    const uint32_t ring_off = (uint32_t)((kBarBytes + kHdrBytes * kStages + sizeof(int4) * kStages * kNT + sizeof(int) * ((FUSED ? 5 : 4) * a.M + (FUSED ? 3 : 2) * a.N) + 127) & ~127u);
    const uint32_t ring = smem_u32(smem_raw) + ring_off;
    // FUSED == 2: receive-weight table [kNT][kCW*32] float2 (.x/.y = the thread's two pixel rows) behind the ring
    const uint32_t wtab = ring + a.stages * kNT * a.wmax * 8u;
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- tile coordinates -------------------------------------------------------
    const uint32_t bid = a.rev ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    // split-major order (a.split_major): consecutive CTAs share the receive range, so the CTAs resident at one time read the
    // windows of the same 1/nsplit of the cube (L2 working set) instead of all of it
    const uint32_t ntile = gridDim.x / a.nsplit;
    const uint32_t tile = a.split_major ? bid % ntile : bid / a.nsplit, split = a.split_major ? bid / ntile : bid % a.nsplit;
    // receive tiles [nt0, nt1) of this CTA (balanced contiguous ranges)
    const uint32_t nt0 = (uint32_t)(((uint64_t)a.numNT * split) / a.nsplit), nt1 = (uint32_t)(((uint64_t)a.numNT * (split + 1)) / a.nsplit);
    const uint32_t ta = tile % a.tilesA, tb = (tile / a.tilesA) % a.tilesB, tc = tile / (a.tilesA * a.tilesB);
    // lane patch: a warp covers lpa pixels along the lane axis x (32/lpa) pixel-row pairs; wA = tA/lpa warps sit side by
    // side along the lane axis.  A compact patch keeps the tap addresses of a half-warp inside one 128-byte row of shared
    // memory; the tile shape follows the pixel spacing so the staged windows stay short on anisotropic grids.
    const uint32_t lpa = a.lpa, lpb = 32u / lpa, wA = a.tA / lpa;
    const uint32_t ia = ta * a.tA + (warp % wA) * lpa + (lane % lpa);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, kCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    const bool have_bounds = (FUSED == 0 && (KEEP == 0 || KEEP == 3)) && a.bounds != nullptr;
    if (have_bounds) { // the six tables are contiguous in shared memory and in the pre-pass output
        const int *src = a.bounds + (uint64_t)tile * (4 * a.M + 2 * a.N);
        for (uint32_t i = tid; i < 4 * a.M + 2 * a.N; i += kThreads) s_dvnmin[i] = __ldg(src + i);
    } else {
    for (uint32_t i = tid; i < a.M; i += kThreads) { s_dvnmin[i] = INT_MAX; s_dvnmax[i] = INT_MIN; s_dvpmin[i] = INT_MAX; s_dvpmax[i] = INT_MIN; }
    for (uint32_t i = tid; i < a.N; i += kThreads) { s_drmin[i] = INT_MAX; s_drmax[i] = INT_MIN; }
    }
    if constexpr (FUSED != 0) {
        for (uint32_t i = tid; i < a.M; i += kThreads) s_txany[i] = (a.fa.tx_kind == AP_TX_NONE);
        for (uint32_t i = tid; i < a.N; i += kThreads) s_rxany[i] = (FUSED != 2);
    }
    __syncthreads();

    // LUT: the tables are sample indices; the general position formula with cinv = 1, t0 = 0, fs = 1 is then exactly 1 + (tm + tn)
    const float cinv = LUT ? 1.0f : __ldg(a.cinv);
    const float fs = LUT ? 1.0f : a.fs;
    const float Tf = (float)a.T;
    const bool VS = a.VS, DV = a.DV;

    if (warp < kCW) {
        // =========================== consumers =====================================
        float px[kR], py[kR], pz[kR];
        uint64_t pix[kR];
        bool valid[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const uint32_t ib = tb * a.tB + ((warp / wA) * lpb + (lane / lpa)) * kR + r;
            valid[r] = (ia < a.IA) && (ib < a.IB);
            // out-of-image lanes shadow a valid pixel so they never widen the windows
            const uint32_t ca = ia < a.IA ? ia : a.IA - 1, cb = ib < a.IB ? ib : a.IB - 1;
            pix[r] = (uint64_t)ca * a.sA + (uint64_t)cb * a.sB + (uint64_t)tc * a.sC;
            if constexpr (LUT) { px[r] = py[r] = pz[r] = 0.f; }
            else {
                px[r] = __ldg(a.Pi + 3 * pix[r]);
                py[r] = __ldg(a.Pi + 3 * pix[r] + 1);
                pz[r] = __ldg(a.Pi + 3 * pix[r] + 2);
            }
        }
        // per-pixel part of the apodization index (NAP arrays, real weights)
        uint32_t aoff[NAP > 0 ? NAP : 1][kR];
        if constexpr (NAP > 0) {
#pragma unroll
            for (int r = 0; r < kR; ++r) {
                const uint32_t i1 = (uint32_t)(pix[r] % a.I1), i2 = (uint32_t)((pix[r] / a.I1) % a.I2), i3 = (uint32_t)(pix[r] / ((uint64_t)a.I1 * a.I2));
#pragma unroll
                for (int q = 0; q < NAP; ++q) aoff[q][r] = i1 * a.ast[q][0] + i2 * a.ast[q][1] + i3 * a.ast[q][2];
            }
        }
        float plat[kR] = {0.f, 0.f}; // lateral coordinate of the pixel (scan.x, or scan.a for polar scans)
        if constexpr (FUSED != 0) {
#pragma unroll
            for (int r = 0; r < kR; ++r) {
                const uint32_t i1 = (uint32_t)(pix[r] % a.I1), i2 = (uint32_t)((pix[r] / a.I1) % a.I2), i3 = (uint32_t)(pix[r] / ((uint64_t)a.I1 * a.I2));
                plat[r] = ap_lateral(a.fa, px[r], i1, i2, i3);
            }
        }
        // ---- phase 0: per-tile min/max of dv(.,m) and dr(.,n) -------------------------
        // dv is tracked in two clusters, dv < 0 and dv >= 0: a focused transmit flips the sign of dv at the focal
        // plane (kern/das_spec.m:429), so a tile crossing it touches two disjoint windows of each trace
        for (uint32_t m = 0; m < (have_bounds ? 0u : a.M); ++m) {
            float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if constexpr (!LUT) {
                pv = __ldg(reinterpret_cast<const float4 *>(a.Pv4) + m);
                nx = __ldg(a.Nv + 3 * m); ny = __ldg(a.Nv + 3 * m + 1); nz = __ldg(a.Nv + 3 * m + 2);
            }
            int nlo = INT_MAX, nhi = INT_MIN, plo = INT_MAX, phi = INT_MIN;
#pragma unroll
            for (int r = 0; r < kR; ++r) {
                const float d = LUT ? __ldg(a.tm + (uint64_t)m * a.I + pix[r]) : tx_dist(px[r], py[r], pz[r], pv.x, pv.y, pv.z, nx, ny, nz, VS, DV);
                const int o = f2o(d);
                if (!LUT && d < 0.f) { nlo = min(nlo, o); nhi = max(nhi, o); }   // (table delays: one cluster)
                else                 { plo = min(plo, o); phi = max(phi, o); }
            }
            nlo = __reduce_min_sync(0xffffffffu, nlo);
            nhi = __reduce_max_sync(0xffffffffu, nhi);
            plo = __reduce_min_sync(0xffffffffu, plo);
            phi = __reduce_max_sync(0xffffffffu, phi);
            if (lane == 0) {
                if (nlo <= nhi) { atomicMin(&s_dvnmin[m], nlo); atomicMax(&s_dvnmax[m], nhi); }
                if (plo <= phi) { atomicMin(&s_dvpmin[m], plo); atomicMax(&s_dvpmax[m], phi); }
            }
            if constexpr (FUSED != 0) {
                if (a.fa.tx_kind != AP_TX_NONE) {
                    const bool nz = ap_tx_weight(a.fa, px[0], py[0], pz[0], plat[0], m) != 0.f || ap_tx_weight(a.fa, px[1], py[1], pz[1], plat[1], m) != 0.f;
                    if (__any_sync(0xffffffffu, nz) && lane == 0) s_txany[m] = 1;
                }
            }
        }
        for (uint32_t n = kInnerTx ? 0u : nt0 * kRX; n < (have_bounds ? 0u : (kInnerTx ? a.N : min(nt1 * kRX, a.N))); ++n) {
            float rx = 0.f, ry = 0.f, rz = 0.f;
            if constexpr (!LUT) { rx = __ldg(a.Pr + 3 * n); ry = __ldg(a.Pr + 3 * n + 1); rz = __ldg(a.Pr + 3 * n + 2); }
            int lo = INT_MAX, hi = INT_MIN;
#pragma unroll
            for (int r = 0; r < kR; ++r) {
                const int o = f2o(LUT ? __ldg(a.tn + (uint64_t)n * a.I + pix[r]) : rx_dist(px[r], py[r], pz[r], rx, ry, rz));
                lo = min(lo, o);
                hi = max(hi, o);
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) { atomicMin(&s_drmin[n], lo); atomicMax(&s_drmax[n], hi); }
            if constexpr (FUSED == 2) {
                const bool nz = ap_rx_weight(a.fa, px[0], py[0], pz[0], plat[0], a.Pr, n) != 0.f || ap_rx_weight(a.fa, px[1], py[1], pz[1], plat[1], a.Pr, n) != 0.f;
                if (__any_sync(0xffffffffu, nz) && lane == 0) s_rxany[n] = 1;
            }
        }
        __syncthreads();

        // ---- phase 1: main loop over the stages the producer publishes ----------------------
        // The ring carries only stages with work: (receive tile nt, transmit m) pairs whose 16 traces
        // are all outside the data for this tile are dropped by the producer and never cost a handshake.
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        // coherence mode: per-receive sums of the current receive group (both pixel rows), sum of their squared magnitudes
        float2 b0[kCF ? 8 : 1], b1[kCF ? 8 : 1];
        float accP0 = 0.f, accP1 = 0.f;
#pragma unroll
        for (int r = 0; r < (kCF ? 8 : 1); ++r) b0[r] = b1[r] = make_float2(0.f, 0.f);
        auto cf_fold = [&]() { // a receive group is complete: S += b_n, P += |b_n|^2   (kern/cohfac.m)
            if constexpr (kCF) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    acc0.x += b0[r].x; acc0.y += b0[r].y; acc1.x += b1[r].x; acc1.y += b1[r].y;
                    accP0 = fmaf(b0[r].x, b0[r].x, fmaf(b0[r].y, b0[r].y, accP0));
                    accP1 = fmaf(b1[r].x, b1[r].x, fmaf(b1[r].y, b1[r].y, accP1));
                    b0[r] = b1[r] = make_float2(0.f, 0.f);
                }
            }
        };
        Pack2 pk, pk2;   // pk2: the second transmit of a coherence-mode stage
        pk.cinv = pk2.cinv = cinv;
        pk.fs = pk2.fs = fs;
        pk2.dv = make_float2(0.f, 0.f); pk2.t0 = 0.f;
        float2 dr[kNT]; // .x = pixel row 0, .y = pixel row 1
#pragma unroll
        for (int j = 0; j < kNT; ++j) dr[j] = make_float2(0.f, 0.f);
        int cur_nt = -1;
        for (uint32_t it = 0;; ++it) {
            const uint32_t s = it % a.stages, ph = (it / a.stages) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            const int4 *hblk = stage_hdr(s);
            const int4 hdr = hblk[0]; // kind, m, nt, uniform sign of dv
            if (hdr.x == ST_END) break;
            if constexpr (kCF) { if ((int)hdr.z != cur_nt && cur_nt >= 0) cf_fold(); }
            // hdr.y = outer index of the stage (transmit m; receive n when kInnerTx), hdr.z = inner tile (16 receives; 16 transmits)
            const uint32_t outer = (uint32_t)hdr.y, nt = (uint32_t)hdr.z;
            const uint32_t m = kInnerTx ? 0u : (kCF ? 2u * outer : outer);
            float t0m = 0.f;
            if ((int)nt != cur_nt) { // new inner tile: its 16 path lengths go to registers
                cur_nt = (int)nt;
#pragma unroll
                for (int j = 0; j < kNT; ++j) {
                    if constexpr (kInnerTx) {
                        const uint32_t mj = min(nt * kNT + j, a.M - 1);
                        const float4 pvj = __ldg(reinterpret_cast<const float4 *>(a.Pv4) + mj);
                        const float nxj = __ldg(a.Nv + 3 * mj), nyj = __ldg(a.Nv + 3 * mj + 1), nzj = __ldg(a.Nv + 3 * mj + 2);
                        dr[j].x = tx_dist(px[0], py[0], pz[0], pvj.x, pvj.y, pvj.z, nxj, nyj, nzj, VS, DV);
                        dr[j].y = tx_dist(px[1], py[1], pz[1], pvj.x, pvj.y, pvj.z, nxj, nyj, nzj, VS, DV);
                    } else if constexpr (LUT) {
                        const uint32_t n = min(nt * kNT + j, a.N - 1);
                        dr[j].x = __ldg(a.tn + (uint64_t)n * a.I + pix[0]);
                        dr[j].y = __ldg(a.tn + (uint64_t)n * a.I + pix[1]);
                    } else {
                        if (kCF && j >= 8) continue;
                        const uint32_t n = min(nt * kRX + j, a.N - 1);
                        const float rx = __ldg(a.Pr + 3 * n), ry = __ldg(a.Pr + 3 * n + 1), rz = __ldg(a.Pr + 3 * n + 2);
                        dr[j].x = rx_dist(px[0], py[0], pz[0], rx, ry, rz);
                        dr[j].y = rx_dist(px[1], py[1], pz[1], rx, ry, rz);
                        if constexpr (FUSED == 2) { // this thread's receive weights for the tile: its own table column
                            const float w0 = ap_rx_weight(a.fa, px[0], py[0], pz[0], plat[0], a.Pr, n);
                            const float w1 = ap_rx_weight(a.fa, px[1], py[1], pz[1], plat[1], a.Pr, n);
                            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(wtab + (uint32_t)(j * kCW * 32 + tid) * 8u), "f"(w0), "f"(w1) : "memory");
                        }
                    }
                }
            }
            if constexpr (kInnerTx) { // stage scalar = receive path length; t0 is per trace (fetched in the loop, warp-uniform)
#if QUPS_HDRGEO
                const float4 prv = reinterpret_cast<const float4 *>(hblk)[1];
                const float rx = prv.x, ry = prv.y, rz = prv.z;
#else
                const float rx = __ldg(a.Pr + 3 * outer), ry = __ldg(a.Pr + 3 * outer + 1), rz = __ldg(a.Pr + 3 * outer + 2);
#endif
                pk.dv.x = rx_dist(px[0], py[0], pz[0], rx, ry, rz);
                pk.dv.y = rx_dist(px[1], py[1], pz[1], rx, ry, rz);
            } else if constexpr (LUT) {
                pk.dv.x = __ldg(a.tm + (uint64_t)m * a.I + pix[0]);
                pk.dv.y = __ldg(a.tm + (uint64_t)m * a.I + pix[1]);
            } else {
#if QUPS_HDRGEO
                const float4 pv = reinterpret_cast<const float4 *>(hblk)[1];
                const int sg0 = kCF ? ((hdr.w & 3) - 1) : hdr.w;
                if (VS && !DV && sg0 != 0) {
                    // every pixel of the tile lies strictly on one side of the transmit's focal plane: dv = +-|Pi - Pv| exactly
                    // (d * (+-1) is exact), no dot product with the normal
                    const float d0 = rx_dist(px[0], py[0], pz[0], pv.x, pv.y, pv.z), d1 = rx_dist(px[1], py[1], pz[1], pv.x, pv.y, pv.z);
                    pk.dv.x = sg0 > 0 ? d0 : -d0;
                    pk.dv.y = sg0 > 0 ? d1 : -d1;
                } else {
                    const float4 nv = reinterpret_cast<const float4 *>(hblk)[2];
                    pk.dv.x = tx_dist(px[0], py[0], pz[0], pv.x, pv.y, pv.z, nv.x, nv.y, nv.z, VS, DV);
                    pk.dv.y = tx_dist(px[1], py[1], pz[1], pv.x, pv.y, pv.z, nv.x, nv.y, nv.z, VS, DV);
                }
                if constexpr (kCF) { // the stage's second transmit (2 outer + 1; geometry zeroed by the producer when it does not exist)
                    const float4 pv1 = reinterpret_cast<const float4 *>(hblk)[3], nv1 = reinterpret_cast<const float4 *>(hblk)[4];
                    const int sg1 = ((hdr.w >> 2) & 3) - 1;
                    if (VS && !DV && sg1 != 0) {
                        const float d0 = rx_dist(px[0], py[0], pz[0], pv1.x, pv1.y, pv1.z), d1 = rx_dist(px[1], py[1], pz[1], pv1.x, pv1.y, pv1.z);
                        pk2.dv.x = sg1 > 0 ? d0 : -d0;
                        pk2.dv.y = sg1 > 0 ? d1 : -d1;
                    } else {
                        pk2.dv.x = tx_dist(px[0], py[0], pz[0], pv1.x, pv1.y, pv1.z, nv1.x, nv1.y, nv1.z, VS, DV);
                        pk2.dv.y = tx_dist(px[1], py[1], pz[1], pv1.x, pv1.y, pv1.z, nv1.x, nv1.y, nv1.z, VS, DV);
                    }
                    pk2.t0 = pv1.w;
                }
#else
                const float4 pv = __ldg(reinterpret_cast<const float4 *>(a.Pv4) + m);
                const float nx = __ldg(a.Nv + 3 * m), ny = __ldg(a.Nv + 3 * m + 1), nz = __ldg(a.Nv + 3 * m + 2);
                pk.dv.x = tx_dist(px[0], py[0], pz[0], pv.x, pv.y, pv.z, nx, ny, nz, VS, DV);
                pk.dv.y = tx_dist(px[1], py[1], pz[1], pv.x, pv.y, pv.z, nx, ny, nz, VS, DV);
#endif
                t0m = pv.w;
            }
            pk.t0 = t0m;
            // kept aperture: fetch the output element early so the read-modify-write at the end of the stage is latency-free
            float2 yold0 = make_float2(0.f, 0.f), yold1 = make_float2(0.f, 0.f);
            if constexpr (KEEP == 1 || KEEP == 2) {
                if (valid[0]) yold0 = a.y[pix[0] + (uint64_t)outer * a.I];
                if (valid[1]) yold1 = a.y[pix[1] + (uint64_t)outer * a.I];
            }
            auto t0_of = [&](int j) -> float { // start time of trace j of this stage
                if constexpr (kInnerTx) return __ldg(a.Pv4 + 4 * min(nt * kNT + (uint32_t)j, a.M - 1) + 3);
                else return t0m;
            };
            const int4 *dsc = desc + s * kNT; // .x = slot offset of the dv < 0 cluster, .y = flag, .z = offset of the dv >= 0 cluster
            // two-level accumulation: the 16 x 4 taps of a stage are summed into stage-local accumulators first
            // (pairwise-style error growth: sqrt(64) + sqrt(#stages) instead of sqrt(#terms))
            float2 sa0 = make_float2(0.f, 0.f), sa1 = make_float2(0.f, 0.f);
            uint32_t tb[NAP > 0 ? NAP : 1]; // trace part of the apodization index at j = 0
            if constexpr (NAP > 0) {
#pragma unroll
                for (int q = 0; q < NAP; ++q) tb[q] = nt * kNT * a.ast[q][3] + m * a.ast[q][4];
            }
            auto apw2 = [&](int j) -> float2 { // product of the array (NAP) and fused receive weights for trace j, both pixel rows
                float2 w = make_float2(1.f, 1.f);
                if constexpr (FUSED == 2) w = lds64(wtab + (uint32_t)(j * kCW * 32 + tid) * 8u);
                if constexpr (NAP > 0) {
                    w.x *= __ldg(a.ap[0] + (aoff[0][0] + tb[0] + (uint32_t)j * a.ast[0][3]));
                    w.y *= __ldg(a.ap[0] + (aoff[0][1] + tb[0] + (uint32_t)j * a.ast[0][3]));
                    if constexpr (NAP > 1) {
                        w.x *= __ldg(a.ap[1] + (aoff[1][0] + tb[1] + (uint32_t)j * a.ast[1][3]));
                        w.y *= __ldg(a.ap[1] + (aoff[1][1] + tb[1] + (uint32_t)j * a.ast[1][3]));
                    }
                }
                return w;
            };
            constexpr bool kWeighted = (NAP > 0) || (FUSED == 2);
            float wt0 = 1.f, wt1 = 1.f; // fused transmit weights of the two pixels for this stage
            bool live = true;
            if constexpr (FUSED != 0) {
                if (a.fa.tx_kind != AP_TX_NONE) {
                    wt0 = ap_tx_weight(a.fa, px[0], py[0], pz[0], plat[0], m);
                    wt1 = ap_tx_weight(a.fa, px[1], py[1], pz[1], plat[1], m);
                    live = __any_sync(0xffffffffu, wt0 != 0.f || wt1 != 0.f); // scanline-type masks: most warps idle
                }
            }
            if (live) {
            if (hdr.x == ST_ALL_FAST) {
                // every trace FAST with a single window: fully unrolled, branch-free
#if QUPS_SOFF4
                const uint4 *so4p = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(hblk) + kHdrTab);
                uint4 so4 = make_uint4(0u, 0u, 0u, 0u);
#endif
#pragma unroll
                for (int j = 0; j < kNT; ++j) {
#if QUPS_SOFF4
                    if ((j & 3) == 0) so4 = so4p[j >> 2];
                    const uint32_t so = (j & 3) == 0 ? so4.x : ((j & 3) == 1 ? so4.y : ((j & 3) == 2 ? so4.z : so4.w));
#else
                    const uint32_t so = (uint32_t)dsc[j].x;
#endif
                    if constexpr (kInnerTx) pk.t0 = t0_of(j);
                    if constexpr (kCF) {   // receive j & 7 of the group, first / second transmit of the stage
                        fast_pair2<INTERP, LUT>(j < 8 ? pk : pk2, dr[j & 7], so, so, b0[j & 7], b1[j & 7]);
                    } else if constexpr (!kWeighted) {
                        fast_pair2<INTERP, LUT>(pk, dr[j], so, so, sa0, sa1);
                    } else { // a .* interp1(...): sample into temporaries, then one weighted accumulate per pixel
                        float2 t0 = make_float2(0.f, 0.f), t1 = make_float2(0.f, 0.f);
                        fast_pair2<INTERP, LUT>(pk, dr[j], so, so, t0, t1);
                        const float2 w = apw2(j);
                        sa0.x = fmaf(w.x, t0.x, sa0.x); sa0.y = fmaf(w.x, t0.y, sa0.y);
                        sa1.x = fmaf(w.y, t1.x, sa1.x); sa1.y = fmaf(w.y, t1.y, sa1.y);
                    }
                }
            } else {
                // general stage: one CTA-uniform branch per trace (SKIP / FAST / rare); each pixel picks the window of
                // its own dv cluster.  Kept ROLLED (dr read through a local-memory copy): unrolling it a second time
                // next to the all-fast body overflows the instruction cache (measured: no_instruction stalls 0.15 -> 1.05)
                bool neg0 = pk.dv.x < 0.f, neg1 = pk.dv.y < 0.f; // dv cluster of the two pixels (per trace when kInnerTx)
                float2 drl[kNT];
#pragma unroll
                for (int j = 0; j < kNT; ++j) drl[j] = dr[j];
#pragma unroll 1
                for (int j = 0; j < kNT; ++j) {
                    const int4 d = dsc[j];
                    if (d.y == TR_SKIP) continue;
                    const float2 drj = drl[kCF ? (j & 7) : j];
                    float t0j = t0m;
                    if constexpr (kInnerTx) { neg0 = drj.x < 0.f; neg1 = drj.y < 0.f; t0j = t0_of(j); pk.t0 = t0j; }
                    const Pack2 &pkj = (kCF && j >= 8) ? pk2 : pk;
                    if constexpr (kCF) { neg0 = pkj.dv.x < 0.f; neg1 = pkj.dv.y < 0.f; t0j = pkj.t0; }
                    const uint32_t so0 = (uint32_t)(neg0 ? d.x : d.z), so1 = (uint32_t)(neg1 ? d.x : d.z);
                    float2 t0 = make_float2(0.f, 0.f), t1 = make_float2(0.f, 0.f);
                    if (d.y == TR_FAST) {
                        fast_pair2<INTERP, LUT>(pkj, drj, so0, so1, t0, t1);
                    } else {
                        const uint32_t n = kInnerTx ? outer : (kCF ? nt * 8 + (j & 7) : nt * kNT + j);
                        const uint32_t mm = kInnerTx ? nt * kNT + j : (kCF ? 2 * outer + (j >> 3) : outer);
                        const uint64_t nm = a.tpose ? ((uint64_t)mm + (uint64_t)n * a.M) : ((uint64_t)n + (uint64_t)mm * a.N);
                        const float xq0 = sample_pos(pkj.dv.x, drj.x, cinv, t0j, fs);
                        const float xq1 = sample_pos(pkj.dv.y, drj.y, cinv, t0j, fs);
                        // an EDGE trace crosses the end of the data somewhere in the TILE; most warps of the tile are
                        // still entirely interior (-> packed fast path) or entirely outside (-> contribute 0)
                        const bool in2 = interior<INTERP>(xq0, Tf) && interior<INTERP>(xq1, Tf);
                        const bool out2 = !(xq0 >= 1.0f && xq0 <= Tf) && !(xq1 >= 1.0f && xq1 <= Tf);
                        if (d.y == TR_EDGE && __all_sync(0xffffffffu, in2)) {
                            fast_pair2<INTERP, LUT>(pkj, drj, so0, so1, t0, t1);
                        } else if (d.y == TR_EDGE && __all_sync(0xffffffffu, out2)) {
                            continue;
                        } else {
                            rare_pair2<INTERP>(a.x + nm * a.T, a.T, d.y, xq0, xq1, so0, so1, t0, t1);
                        }
                    }
                    if constexpr (kCF) {   // rolled loop: the receive slot is picked by predicated adds (static register indices)
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if ((j & 7) == r) { b0[r].x += t0.x; b0[r].y += t0.y; b1[r].x += t1.x; b1[r].y += t1.y; }
                    } else if constexpr (!kWeighted) {
                        sa0.x += t0.x; sa0.y += t0.y; sa1.x += t1.x; sa1.y += t1.y;
                    } else {
                        const float2 w = apw2(j);
                        sa0.x = fmaf(w.x, t0.x, sa0.x); sa0.y = fmaf(w.x, t0.y, sa0.y);
                        sa1.x = fmaf(w.y, t1.x, sa1.x); sa1.y = fmaf(w.y, t1.y, sa1.y);
                    }
                }
            }
            } // live
            if constexpr (kCF) {
                // nothing per stage: the per-receive sums are folded when the receive group changes
            } else if constexpr (KEEP != 0) { // y(:, outer) += stage sum; this CTA is the only writer of its pixels
                if (valid[0]) a.y[pix[0] + (uint64_t)outer * a.I] = make_float2(yold0.x + sa0.x, yold0.y + sa0.y);
                if (valid[1]) a.y[pix[1] + (uint64_t)outer * a.I] = make_float2(yold1.x + sa1.x, yold1.y + sa1.y);
            } else if constexpr (FUSED != 0) { // the transmit weight is constant over the stage: one multiply per pixel
                acc0.x = fmaf(wt0, sa0.x, acc0.x); acc0.y = fmaf(wt0, sa0.y, acc0.y);
                acc1.x = fmaf(wt1, sa1.x, acc1.x); acc1.y = fmaf(wt1, sa1.y, acc1.y);
            } else {
                acc0.x += sa0.x; acc0.y += sa0.y; acc1.x += sa1.x; acc1.y += sa1.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        }
        if constexpr (LUT) { // y = w * sum (scalar weight; w = 1 leaves the sum untouched)
            const float wr = __ldg(a.wscal), wi = a.wcplx ? __ldg(a.wscal + 1) : 0.f;
            if (a.wcplx) {
                acc0 = make_float2(wr * acc0.x - wi * acc0.y, wr * acc0.y + wi * acc0.x);
                acc1 = make_float2(wr * acc1.x - wi * acc1.y, wr * acc1.y + wi * acc1.x);
            } else if (wr != 1.0f) {
                acc0.x *= wr; acc0.y *= wr; acc1.x *= wr; acc1.y *= wr;
            }
        }
        if constexpr (kCF) {   // partial image and partial power of this receive range; das_cf_reduce_kernel finishes
            cf_fold();
            if (valid[0]) { a.part[(uint64_t)split * a.I + pix[0]] = acc0; a.partP[(uint64_t)split * a.I + pix[0]] = accP0; }
            if (valid[1]) { a.part[(uint64_t)split * a.I + pix[1]] = acc1; a.partP[(uint64_t)split * a.I + pix[1]] = accP1; }
        } else if constexpr (KEEP != 0) {
            // nothing left to write: every stage updated y in place
        } else if (a.nsplit > 1) { // partial image of this receive range; das_reduce_kernel sums the splits in order
            if (valid[0]) a.part[(uint64_t)split * a.I + pix[0]] = acc0;
            if (valid[1]) a.part[(uint64_t)split * a.I + pix[1]] = acc1;
        } else {
            if (a.accumulate) {
                if (valid[0]) { const float2 o = a.y[pix[0]]; acc0.x += o.x; acc0.y += o.y; }
                if (valid[1]) { const float2 o = a.y[pix[1]]; acc1.x += o.x; acc1.y += o.y; }
            }
            if (valid[0]) a.y[pix[0]] = acc0;
            if (valid[1]) a.y[pix[1]] = acc1;
        }
    } else {
        // =========================== producer warps ================================
        // One warp needs ~400 mostly dependent instructions per stage (ncu: with one producer the consumers of the
        // 1- and 2-tap kernels spent 29 % of their time waiting for data).  kPW warps run the same candidate loop and
        // take the candidates in turn; the stage index is the CANDIDATE index, so no warp needs the other's result
        // (a candidate whose exact per-trace test finds nothing is then published as an all-SKIP stage).
        // Measured (C2): nearest 42.6 -> 37.7 ms with two producers; linear unchanged and cubic 2 % slower with a tenth warp in
        // the CTA (their consumers are shared-memory / issue bound) — so only the 1-tap kernel is launched with the second.
        const uint32_t pw = (uint32_t)warp - kCW;
        constexpr uint32_t nprod = (uint32_t)(kThreads / 32 - kCW);
        __syncthreads(); // matches the consumers' post-phase-0 barrier
        const bool cinv_ok = (cinv > 0.f) && (fs > 0.f);
        const int Ti = (int)a.T;
        uint32_t it = 0;
        // inner tiles: 16 receives (16 transmits when kInnerTx); outer index: transmit m (receive n when kInnerTx)
        // coherence mode: inner tile = 8 receives, outer index = a PAIR of transmits; lane l stages receive (l & 7) of transmit (l >> 3)
        const uint32_t n_inner = kInnerTx ? a.M : a.N, n_outer = kInnerTx ? a.N : (kCF ? (a.M + 1) / 2 : a.M);
        for (uint32_t nt = nt0; nt < nt1; ++nt) {
            const uint32_t il = kCF ? nt * 8 + (lane & 7) : nt * kNT + lane;
            const bool has = (lane < kNT) && (il < n_inner);
            // tile-level bounds of the inner path length (and of t0 when the inner traces are transmits): one conservative
            // skip test per outer index, 32 outer indices per pass
            int olo = INT_MAX, ohi = INT_MIN;
            float tlo = 0.f, thi = 0.f;
            if (has) {
                if constexpr (kInnerTx) { olo = min(s_dvnmin[il], s_dvpmin[il]); ohi = max(s_dvnmax[il], s_dvpmax[il]); tlo = thi = __ldg(a.Pv4 + 4 * il + 3); }
                else { olo = s_drmin[il]; ohi = s_drmax[il]; }
            }
            const float rlo_t = o2f(__reduce_min_sync(0xffffffffu, olo)), rhi_t = o2f(__reduce_max_sync(0xffffffffu, ohi));
            float t0lo_t = 0.f, t0hi_t = 0.f;
            if constexpr (kInnerTx) {
                t0lo_t = o2f(__reduce_min_sync(0xffffffffu, has ? f2o(tlo) : INT_MAX));
                t0hi_t = o2f(__reduce_max_sync(0xffffffffu, has ? f2o(thi) : INT_MIN));
            }
            for (uint32_t m0 = 0; m0 < n_outer; m0 += 32) {
                const uint32_t ml = m0 + lane;
                bool skip = true;
                if (ml < n_outer) {
                    float xl, xh;
                    if constexpr (kInnerTx) { // outer = receive ml
                        xl = sample_pos(rlo_t, o2f(s_drmin[ml]), cinv, t0hi_t, fs);
                        xh = sample_pos(rhi_t, o2f(s_drmax[ml]), cinv, t0lo_t, fs);
                    } else {
                        if constexpr (kCF) { // the pair of transmits 2 ml, 2 ml + 1: keep the stage if either can reach the data
                            const uint32_t ma = 2 * ml, mb = min(2 * ml + 1, a.M - 1);
                            const float ta = __ldg(a.Pv4 + 4 * ma + 3), tb = __ldg(a.Pv4 + 4 * mb + 3);
                            const float xla = sample_pos(o2f(min(s_dvnmin[ma], s_dvpmin[ma])), rlo_t, cinv, ta, fs);
                            const float xha = sample_pos(o2f(max(s_dvnmax[ma], s_dvpmax[ma])), rhi_t, cinv, ta, fs);
                            const float xlb = sample_pos(o2f(min(s_dvnmin[mb], s_dvpmin[mb])), rlo_t, cinv, tb, fs);
                            const float xhb = sample_pos(o2f(max(s_dvnmax[mb], s_dvpmax[mb])), rhi_t, cinv, tb, fs);
                            const bool ska = (xla <= xha) && (xha < 1.0f || xla > Tf), skb = (xlb <= xhb) && (xhb < 1.0f || xlb > Tf);
                            xl = (ska && skb) ? 2.0f * Tf : 1.0f;   // both out of range -> (xl > Tf) below skips the pair
                            xh = (ska && skb) ? 3.0f * Tf : Tf;
                        } else {
                        const float t0l = LUT ? 0.f : __ldg(a.Pv4 + 4 * ml + 3);
                        xl = sample_pos(o2f(min(s_dvnmin[ml], s_dvpmin[ml])), rlo_t, cinv, t0l, fs);
                        xh = sample_pos(o2f(max(s_dvnmax[ml], s_dvpmax[ml])), rhi_t, cinv, t0l, fs);
                        }
                    }
                    skip = cinv_ok && (xl <= xh) && (xh < 1.0f || xl > Tf);
                    if constexpr (FUSED != 0) skip = skip || (s_txany[ml] == 0); // no pixel of the tile uses this transmit
                }
                uint32_t todo = ~__ballot_sync(0xffffffffu, skip);
                while (todo) {
                    const uint32_t outer = m0 + (uint32_t)(__ffs((int)todo) - 1);
                    todo &= todo - 1;
                    if (nprod > 1 && (it % nprod) != pw) { ++it; continue; } // another producer warp's stage
                    // this lane's trace: (receive n, transmit m)
                    const uint32_t m = kInnerTx ? il : (kCF ? 2 * outer + (lane >> 3) : outer), n = kInnerTx ? outer : il;
                    const uint32_t s = it % a.stages, ph = (it / a.stages) & 1;
                    int flag = TR_SKIP;
                    // up to two windows per trace: [0] single window / dv < 0 cluster, [1] dv >= 0 cluster
                    uint32_t bytes[2] = {0u, 0u}, soff[2] = {0u, 0u}, dst[2] = {0u, 0u};
                    const float2 *src[2] = {nullptr, nullptr};
                    if (has && (!kCF || m < a.M)) {   // (coherence mode: an odd transmit count leaves the last pair half empty)
                        const float t0m = LUT ? 0.f : __ldg(a.Pv4 + 4 * m + 3);
                        const int nmin = s_dvnmin[m], nmax = s_dvnmax[m], pmin = s_dvpmin[m], pmax = s_dvpmax[m];
                        const float rlo = o2f(s_drmin[n]), rhi = o2f(s_drmax[n]); // receive path-length bounds of this trace
                        const float xlo = sample_pos(o2f(min(nmin, pmin)), rlo, cinv, t0m, fs);
                        const float xhi = sample_pos(o2f(max(nmax, pmax)), rhi, cinv, t0m, fs);
                        flag = TR_SLOW;
                        const uint64_t tr = a.tpose ? ((uint64_t)m + (uint64_t)n * a.M) : ((uint64_t)n + (uint64_t)m * a.N);
                        const uint32_t slot = ring + (s * kNT + lane) * a.wmax * 8u;
                        // place window [t_lo, t_hi] of this trace at slot element `at`; returns its length (0 = does not fit)
                        auto place = [&](int q, int t_lo, int t_hi, int tap0, uint32_t at) -> uint32_t {
                            const int64_t abs_lo = (int64_t)(tr * a.T) + t_lo;
                            const int64_t abs_al = abs_lo & ~(int64_t)1; // 16-byte aligned element
                            const int w0 = t_lo - (int)(abs_lo - abs_al);
                            int wlen = t_hi - w0 + 1;
                            wlen = (wlen + 1) & ~1;
                            if (!(wlen > 0 && at + (uint32_t)wlen <= a.wmax && abs_al >= 0 && (uint64_t)(abs_al + wlen) <= a.total_elems)) return 0u;
                            bytes[q] = (uint32_t)wlen * 8u;
                            dst[q] = slot + at * 8u;
                            soff[q] = dst[q] - (uint32_t)(w0 + tap0) * 8u - magic_off<INTERP>::value;
                            src[q] = a.x + abs_al;
                            return (uint32_t)wlen;
                        };
                        if (cinv_ok && xlo <= xhi) {
                            int t_lo = 0, t_hi = 0, tap0 = 0;
                            bool inr = false;
                            if (!tap_window<INTERP>(xlo, xhi, Tf, Ti, t_lo, t_hi, inr, tap0)) {
                                flag = TR_SKIP; // every pixel of the tile is outside the trace: contributes 0
                            } else if (place(0, t_lo, t_hi, tap0, 0u)) {
                                flag = inr ? TR_FAST : TR_EDGE;
                                soff[1] = soff[0];
                            } else if (nmin <= nmax && pmin <= pmax) {
                                // the tile straddles the sign flip of dv: stage the windows of the two clusters
                                const float xnl = sample_pos(o2f(nmin), rlo, cinv, t0m, fs), xnh = sample_pos(o2f(nmax), rhi, cinv, t0m, fs);
                                const float xpl = sample_pos(o2f(pmin), rlo, cinv, t0m, fs), xph = sample_pos(o2f(pmax), rhi, cinv, t0m, fs);
                                if (xnl <= xnh && xpl <= xph) {
                                    int nl = 0, nh = 0, pl = 0, ph2 = 0, tp = 0;
                                    bool inn = true, inp = true;
                                    const int hn = tap_window<INTERP>(xnl, xnh, Tf, Ti, nl, nh, inn, tp);
                                    const int hp = tap_window<INTERP>(xpl, xph, Tf, Ti, pl, ph2, inp, tp);
                                    uint32_t used = 0;
                                    bool ok = true;
                                    if (hn) { used = place(0, nl, nh, tp, 0u); ok = used != 0; }
                                    if (ok && hp) ok = place(1, pl, ph2, tp, used) != 0;
                                    if (ok) {
                                        // a cluster entirely outside the trace has no window: its pixels fail the range test
                                        // of the EDGE path and contribute 0 (the offset is never dereferenced)
                                        flag = (hn && hp && inn && inp) ? TR_FAST : TR_EDGE;
                                        if (!hn && !hp) flag = TR_SKIP;
                                    } else {
                                        bytes[0] = bytes[1] = 0u;
                                    }
                                }
                            }
                        }
                    }
                    if constexpr (FUSED == 2) { // receive n is masked for every pixel of the tile: nothing to stage
                        if (has && s_rxany[n] == 0) { flag = TR_SKIP; bytes[0] = bytes[1] = 0u; }
                    }
                    if (nprod == 1 && __all_sync(0xffffffffu, flag == TR_SKIP)) continue; // exact per-trace test: nothing to do
                    const bool all_fast = __all_sync(0xffffffffu, (flag == TR_FAST && bytes[1] == 0u) || lane >= kNT);
                    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes[0] + bytes[1]);
#if QUPS_STATS
                    if (lane == 0 && total == 0) atomicAdd(&g_stats[7], 1ull);
                    if (lane < kNT) { atomicAdd(&g_stats[flag], 1ull); if (bytes[1]) atomicAdd(&g_stats[4], 1ull); }
                    if (lane == 0) atomicAdd(&g_stats[all_fast ? 5 : 6], 1ull);
                    {
                        const bool fs_only = __all_sync(0xffffffffu, ((flag == TR_FAST && bytes[1] == 0u) || flag == TR_SKIP) || lane >= kNT);
                        const bool any_dual = __any_sync(0xffffffffu, bytes[1] != 0u && lane < kNT);
                        const bool any_edge = __any_sync(0xffffffffu, flag == TR_EDGE && lane < kNT);
                        const bool any_slow = __any_sync(0xffffffffu, flag == TR_SLOW && lane < kNT);
                        if (lane == 0 && !all_fast) { if (fs_only) atomicAdd(&g_stats[8], 1ull); if (any_dual) atomicAdd(&g_stats[9], 1ull); if (any_edge) atomicAdd(&g_stats[10], 1ull); if (any_slow) atomicAdd(&g_stats[11], 1ull); }
                    }
#endif
                    // stage header payload (warp-uniform loads, issued before the wait so their latency overlaps it)
                    int sgn = 0;
#if QUPS_HDRGEO
                    float4 gA = make_float4(0.f, 0.f, 0.f, 0.f), gB = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 gA1 = make_float4(0.f, 0.f, 0.f, 0.f), gB1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    // uniform sign of dv over the tile: the positive cluster also holds dv == 0 (sign(0) = 0, where dv = 0 and
                    // not |Pi - Pv|), so "+1" needs its minimum strictly positive
                    auto tile_sign = [&](uint32_t mm) -> int {
                        const int nmin_ = s_dvnmin[mm], nmax_ = s_dvnmax[mm], pmin_ = s_dvpmin[mm], pmax_ = s_dvpmax[mm];
                        if (nmin_ > nmax_ && pmin_ <= pmax_ && o2f(pmin_) > 0.f) return 1;
                        if (pmin_ > pmax_ && nmin_ <= nmax_) return -1;
                        return 0;
                    };
                    if constexpr (LUT) {
                        // nothing: the consumers read their delays from the tables
                    } else if constexpr (kInnerTx) {
                        gA = make_float4(__ldg(a.Pr + 3 * outer), __ldg(a.Pr + 3 * outer + 1), __ldg(a.Pr + 3 * outer + 2), 0.f);
                    } else if constexpr (kCF) {
                        const uint32_t ma = 2 * outer, mb = 2 * outer + 1;
                        gA = __ldg(reinterpret_cast<const float4 *>(a.Pv4) + ma);
                        gB = make_float4(__ldg(a.Nv + 3 * ma), __ldg(a.Nv + 3 * ma + 1), __ldg(a.Nv + 3 * ma + 2), 0.f);
                        int s1 = 0;
                        if (mb < a.M) {
                            gA1 = __ldg(reinterpret_cast<const float4 *>(a.Pv4) + mb);
                            gB1 = make_float4(__ldg(a.Nv + 3 * mb), __ldg(a.Nv + 3 * mb + 1), __ldg(a.Nv + 3 * mb + 2), 0.f);
                            s1 = tile_sign(mb);
                        }
                        sgn = (tile_sign(ma) + 1) | ((s1 + 1) << 2);
                    } else {
                        gA = __ldg(reinterpret_cast<const float4 *>(a.Pv4) + outer);
                        gB = make_float4(__ldg(a.Nv + 3 * outer), __ldg(a.Nv + 3 * outer + 1), __ldg(a.Nv + 3 * outer + 2), 0.f);
                        sgn = tile_sign(outer);
                    }
#endif
                    mbar_wait(bar_empty + 8 * s, ph ^ 1); // slot free (first lap passes immediately)
                    if (lane < kNT) desc[s * kNT + lane] = make_int4((int)soff[0], flag, (int)soff[1], 0);
                    if (lane < kNT) reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(stage_hdr(s)) + kHdrTab)[lane] = soff[0];
                    if (lane == 0) {
#if QUPS_HDRGEO
                        reinterpret_cast<float4 *>(stage_hdr(s))[1] = gA;
                        reinterpret_cast<float4 *>(stage_hdr(s))[2] = gB;
                        if constexpr (kCF) { reinterpret_cast<float4 *>(stage_hdr(s))[3] = gA1; reinterpret_cast<float4 *>(stage_hdr(s))[4] = gB1; }
#endif
                        stage_hdr(s)[0] = make_int4(all_fast ? ST_ALL_FAST : ST_MIXED, (int)outer, (int)nt, sgn);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_expect_tx(bar_full + 8 * s, total);
                    __syncwarp();
                    if (bytes[0]) bulk_g2s(dst[0], src[0], bytes[0], bar_full + 8 * s);
                    if (bytes[1]) bulk_g2s(dst[1], src[1], bytes[1], bar_full + 8 * s);
                    ++it;
                }
            }
        }
        // end-of-work marker
        {
            const uint32_t s = it % a.stages, ph = (it / a.stages) & 1;
            if ((it % nprod) == pw) mbar_wait(bar_empty + 8 * s, ph ^ 1);
            if ((it % nprod) == pw && lane == 0) {
                stage_hdr(s)[0] = make_int4(ST_END, 0, 0, 0);
                mbar_arrive(bar_full + 8 * s);
            }
        }
    }
}

// per-tile path-length bounds for receive-split launches (plain geometric delays, no kept aperture): phase 0 of the main
// kernel, run once per TILE.  A (tile, split) CTA then loads its 4 M + 2 N bounds instead of recomputing dv(.,m) for all M
// transmits (ncu, 8-GPU slab: ~30 000 instructions per warp and CTA — 1.1 % of the launch per unit of nsplit)
__global__ void __launch_bounds__(kCW * 32) das_bounds_kernel(const TiledArgs a) {
    extern __shared__ int s_b[];
    int *s_dvnmin = s_b, *s_dvnmax = s_dvnmin + a.M, *s_dvpmin = s_dvnmax + a.M, *s_dvpmax = s_dvpmin + a.M;
    int *s_drmin = s_dvpmax + a.M, *s_drmax = s_drmin + a.N;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tile = blockIdx.x;
    const uint32_t ta = tile % a.tilesA, tb = (tile / a.tilesA) % a.tilesB, tc = tile / (a.tilesA * a.tilesB);
    const uint32_t lpa = a.lpa, lpb = 32u / lpa, wA = a.tA / lpa;
    const uint32_t ia = ta * a.tA + (warp % wA) * lpa + (lane % lpa);
    for (uint32_t i = tid; i < a.M; i += kCW * 32) { s_dvnmin[i] = INT_MAX; s_dvnmax[i] = INT_MIN; s_dvpmin[i] = INT_MAX; s_dvpmax[i] = INT_MIN; }
    for (uint32_t i = tid; i < a.N; i += kCW * 32) { s_drmin[i] = INT_MAX; s_drmax[i] = INT_MIN; }
    __syncthreads();
    const bool lut = a.tn != nullptr;   // table-driven delays: the path lengths are read, one cluster per transmit
    float px[kR] = {0.f, 0.f}, py[kR] = {0.f, 0.f}, pz[kR] = {0.f, 0.f};
    uint64_t pixr[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) { // the same pixels (incl. the shadowing of out-of-image lanes) as the main kernel's consumers
        const uint32_t ib = tb * a.tB + ((warp / wA) * lpb + (lane / lpa)) * kR + r;
        const uint32_t ca = ia < a.IA ? ia : a.IA - 1, cb = ib < a.IB ? ib : a.IB - 1;
        const uint64_t pix = (uint64_t)ca * a.sA + (uint64_t)cb * a.sB + (uint64_t)tc * a.sC;
        pixr[r] = pix;
        if (!lut) { px[r] = __ldg(a.Pi + 3 * pix); py[r] = __ldg(a.Pi + 3 * pix + 1); pz[r] = __ldg(a.Pi + 3 * pix + 2); }
    }
    const bool VS = a.VS, DV = a.DV;
    for (uint32_t m = 0; m < a.M; ++m) {
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (!lut) {
            pv = __ldg(reinterpret_cast<const float4 *>(a.Pv4) + m);
            nx = __ldg(a.Nv + 3 * m); ny = __ldg(a.Nv + 3 * m + 1); nz = __ldg(a.Nv + 3 * m + 2);
        }
        int nlo = INT_MAX, nhi = INT_MIN, plo = INT_MAX, phi = INT_MIN;
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const float d = lut ? __ldg(a.tm + (uint64_t)m * a.I + pixr[r]) : tx_dist(px[r], py[r], pz[r], pv.x, pv.y, pv.z, nx, ny, nz, VS, DV);
            const int o = f2o(d);
            if (!lut && d < 0.f) { nlo = min(nlo, o); nhi = max(nhi, o); }
            else                 { plo = min(plo, o); phi = max(phi, o); }
        }
        nlo = __reduce_min_sync(0xffffffffu, nlo); nhi = __reduce_max_sync(0xffffffffu, nhi);
        plo = __reduce_min_sync(0xffffffffu, plo); phi = __reduce_max_sync(0xffffffffu, phi);
        if (lane == 0) {
            if (nlo <= nhi) { atomicMin(&s_dvnmin[m], nlo); atomicMax(&s_dvnmax[m], nhi); }
            if (plo <= phi) { atomicMin(&s_dvpmin[m], plo); atomicMax(&s_dvpmax[m], phi); }
        }
    }
    for (uint32_t n = 0; n < a.N; ++n) {
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (!lut) { rx = __ldg(a.Pr + 3 * n); ry = __ldg(a.Pr + 3 * n + 1); rz = __ldg(a.Pr + 3 * n + 2); }
        int lo = INT_MAX, hi = INT_MIN;
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const int o = f2o(lut ? __ldg(a.tn + (uint64_t)n * a.I + pixr[r]) : rx_dist(px[r], py[r], pz[r], rx, ry, rz));
            lo = min(lo, o); hi = max(hi, o);
        }
        lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { atomicMin(&s_drmin[n], lo); atomicMax(&s_drmax[n], hi); }
    }
    __syncthreads();
    int *dst = a.bounds + (uint64_t)tile * (4 * a.M + 2 * a.N);
    for (uint32_t i = tid; i < 4 * a.M + 2 * a.N; i += kCW * 32) dst[i] = s_b[i];
}

// sums the receive-split partial images in split order (deterministic) into y
__global__ void __launch_bounds__(256) das_reduce_kernel(float2 *y, const float2 *part, uint64_t I, uint32_t nsplit, int accumulate) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < I; i += (uint64_t)gridDim.x * blockDim.x) {
        float2 acc = accumulate ? y[i] : make_float2(0.f, 0.f);
        for (uint32_t s = 0; s < nsplit; ++s) {
            const float2 v = part[(uint64_t)s * I + i];
            acc.x += v.x;
            acc.y += v.y;
        }
        y[i] = acc;
    }
}

// coherence mode: sums the receive-split partial images and powers in split order, y = S, cf = |S|^2 / P / N   (kern/cohfac.m)
__global__ void __launch_bounds__(256) das_cf_reduce_kernel(float2 *y, float *cf, const float2 *part, const float *partP, uint64_t I,
                                                            uint32_t nsplit, float N) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < I; i += (uint64_t)gridDim.x * blockDim.x) {
        float2 acc = make_float2(0.f, 0.f);
        float pw = 0.f;
        for (uint32_t s = 0; s < nsplit; ++s) {
            const float2 v = part[(uint64_t)s * I + i];
            acc.x += v.x; acc.y += v.y;
            pw += partP[(uint64_t)s * I + i];
        }
        y[i] = acc;
        cf[i] = (acc.x * acc.x + acc.y * acc.y) / pw / N;   // 0 / 0 = NaN where no receive contributes, as the reference
    }
}

// ---- host side ------------------------------------------------------------------------
static size_t tiled_smem_bytes(uint32_t N, uint32_t M, uint32_t wmax, int fused = 0, uint32_t stages = kStages) {
    size_t head = kBarBytes + (size_t)kHdrBytes * kStages + sizeof(int4) * kStages * kNT + sizeof(int) * ((fused ? 5 : 4) * (size_t)M + (fused ? 3 : 2) * (size_t)N);
    head = (head + 127) & ~(size_t)127;
    return head + (size_t)stages * kNT * wmax * 8 + (fused == 2 ? (size_t)kNT * kCW * 32 * 8 : 0);
}

TiledPlan das_tiled_plan(const DasArgs<float> &a, int dtype_in, int dtype_out) {
    TiledPlan p{0, ""};
    if (dtype_in != 0 || dtype_out != 0) { p.why = "tiled path is fp32 only"; return p; }
    if (a.keep_rx && a.keep_tx) { p.why = "tiled path keeps at most one aperture"; return p; }
    if (a.cohfac && (a.keep_rx || a.keep_tx || a.S > 0 || a.fused || a.lut_tn || a.accumulate || !a.cf)) { p.why = "coherence mode: plain weights, no kept aperture"; return p; }
    if ((a.keep_rx || a.keep_tx) && (a.S > 0 || a.fused)) { p.why = "tiled path: kept apertures take no apodization"; return p; }
    if (a.S > 2 || (a.S > 0 && !a.apod_real)) { p.why = "tiled path takes at most two REAL apodization arrays"; return p; }
    if (a.fused && a.S > 1) { p.why = "tiled path: closed-form apodization combines with at most one apodization array"; return p; }
    for (int q = 0; q < a.S; ++q) { // 32-bit index arithmetic inside the kernel
        uint64_t last = a.astride[q][5] + (a.I1 - 1) * a.astride[q][0] + (a.I2 - 1) * a.astride[q][1] + (a.I3 - 1) * a.astride[q][2] +
                        (a.N - 1) * a.astride[q][3] + (a.M - 1) * a.astride[q][4];
        if (last >= (1ull << 31)) { p.why = "apodization array too large for the tiled path"; return p; }
    }
    for (int d = 0; d < 5; ++d)
        if (a.cstride[d] != 0) { p.why = "tiled path needs a scalar sound speed"; return p; }
    if (a.interp < 0 || a.interp > 2) { p.why = "tiled path: nearest|linear|cubic"; return p; }
    if (!(a.fs > 0.f)) { p.why = "fs <= 0"; return p; }
    if (a.T < 4 || a.T >= (1u << 22)) { p.why = "T out of range for the tiled path"; return p; }
    if (a.N == 0 || a.M == 0 || a.I == 0) { p.why = "empty"; return p; }
    if (a.N >= (1u << 24) || a.M >= (1u << 24) || a.I >= (1ull << 40)) { p.why = "too large"; return p; }
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0) { p.why = "x not 16-byte aligned"; return p; }
    if ((reinterpret_cast<uintptr_t>(a.Pv4) & 15) != 0) { p.why = "Pv not 16-byte aligned"; return p; }
    if (tiled_smem_bytes((uint32_t)a.N, (uint32_t)a.M, 64, 2) > 200 * 1024) { p.why = "N+M too large for smem"; return p; }
    p.eligible = 1;
    return p;
}

int launch_das_tiled(const DasArgs<float> &a, cudaStream_t st) {
    TiledArgs t{};
    t.Pi = a.Pi; t.Pr = a.Pr; t.Pv4 = a.Pv4; t.Nv = a.Nv; t.cinv = a.cinv;
    t.x = reinterpret_cast<const float2 *>(a.x);
    t.y = reinterpret_cast<float2 *>(a.y);
    t.N = (uint32_t)a.N; t.M = (uint32_t)a.M; t.T = (uint32_t)a.T;
    t.fs = a.fs; t.VS = a.VS; t.DV = a.DV; t.tpose = a.tpose; t.accumulate = a.accumulate;
    t.total_elems = a.T * a.N * a.M;
    const bool lut = a.lut_tn != nullptr;   // table-driven delays (wsinterpd2 canonical form): no geometry arrays at all
    t.tn = a.lut_tn; t.tm = a.lut_tm; t.wscal = a.lut_w; t.wcplx = a.lut_wcplx;
    const int keep = a.cohfac ? 3 : (a.keep_tx ? 1 : (a.keep_rx ? 2 : 0));
    // inner tiles: 16 receives; 16 transmits when the receive dimension is kept; 8 receives in coherence mode
    t.numNT = keep == 3 ? (t.N + 7) / 8 : ((keep == 2 ? t.M : t.N) + kNT - 1) / kNT;
    // axis assignment: lanes along I2 (the slow axis of a ZXY ScanCartesian, src/ScanCartesian.m:11) when
    // it is wide enough, rows along I1; overridable for experiments (QUPS_B200_LANE_AXIS=1|2)
    int lane_axis = (a.I2 >= 8) ? 2 : 1;
    if (const char *e = getenv("QUPS_B200_LANE_AXIS")) { int v = atoi(e); if (v == 1 || v == 2) lane_axis = v; }
    if (lane_axis == 2) { t.IA = (uint32_t)a.I2; t.sA = a.I1; t.IB = (uint32_t)a.I1; t.sB = 1; }
    else                { t.IA = (uint32_t)a.I1; t.sA = 1;    t.IB = (uint32_t)a.I2; t.sB = a.I1; }
    t.IC = (uint32_t)a.I3; t.sC = a.I1 * a.I2;
    double span_m = 0.0; // path-length spread across one tile (metres, sum of the two tile extents)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool noprobe = lut || (cudaStreamIsCapturing(st, &cap) != cudaSuccess) || cap != cudaStreamCaptureStatusNone || getenv("QUPS_B200_NOPROBE");
    // ---- tile shape: 512 pixels as tA x tB with a lpa x (32/lpa) x 2 warp patch, chosen so that the delay spread across the
    // tile (~ pixel spacing x extent) is smallest: square-ish on isotropic grids (32 x 16, patch 8 x 4 x 2 — measured best on the
    // headline grid), narrow along a coarsely sampled axis (e.g. one image column per element pitch).  The spacing is read
    // from three pixel positions; the decision is cached per (Pi, grid) so only the first call with a new grid synchronises.
    const bool hinted = a.pitch_hint[0] > 0.0 && a.pitch_hint[1] > 0.0 && a.c_hint > 0.0; // caller supplied the grid pitch: no probe, no cache
    {
        struct ShapeCache { const void *Pi; uint64_t I1, I2, I3; int lane_axis; uint32_t tA, lpa; double span; };
        static thread_local ShapeCache sc = {nullptr, 0, 0, 0, 0, 0, 0, 0.0};
        uint32_t tA = 32, lpa = QUPS_LPA;
        // tile shape from the pixel pitch along the lane (dA) and row (dB) axes
        auto choose = [&](double dA, double dB, bool okA, bool okB) {
            if (!(dA > 0) || !(dB > 0) || !(dA == dA) || !(dB == dB)) dA = dB = 1.0;
            if (!okA) dA = dB * 1e3;   // degenerate axis: make the tile as thin as possible along it
            if (!okB) dB = dA * 1e3;
            double best = 1e300;
            for (uint32_t l = 1; l <= 32; l <<= 1)
                for (uint32_t w = 1; w <= (uint32_t)kCW; w <<= 1) {
                    const uint32_t ta_ = l * w, tb_ = kTilePix / ta_;
                    const double tile = dA * ta_ + dB * tb_, warpspan = dA * l + dB * (2.0 * (32 / l));
                    // prefer the measured-best isotropic shape on ties (32 x 16, patch 8): tiny bias
                    const double cost = tile + 0.25 * warpspan + ((ta_ == 32 && l == (uint32_t)QUPS_LPA) ? -1e-9 * tile : 0.0);
                    if (cost < best) { best = cost; tA = ta_; lpa = l; span_m = tile; }
                }
            if (!okA || !okB || dA == 1.0) span_m = 0.0; // unknown spacing: keep the default ring
        };
        // the probe synchronises the stream once per new grid: not allowed while the stream is being captured into a CUDA
        // graph (and skippable with QUPS_B200_NOPROBE=1) — the default shape / ring are then used, results are unaffected
        if (hinted) {
            const double d1 = a.pitch_hint[0], d2 = a.pitch_hint[1];
            choose(lane_axis == 2 ? d2 : d1, lane_axis == 2 ? d1 : d2, t.IA > 1, t.IB > 1);
        } else if (sc.Pi == a.Pi && sc.I1 == a.I1 && sc.I2 == a.I2 && sc.I3 == a.I3 && sc.lane_axis == lane_axis) {
            tA = sc.tA; lpa = sc.lpa; span_m = sc.span;
        } else if (noprobe) {
            span_m = 0.0;
        } else {
            float P0[3], PA[3], PB[3];
            double dA = 1.0, dB = 1.0;
            const bool okA = t.IA > 1, okB = t.IB > 1;
            cudaError_t ce = cudaMemcpyAsync(P0, a.Pi, sizeof(P0), cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess && okA) ce = cudaMemcpyAsync(PA, a.Pi + 3 * t.sA, sizeof(PA), cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess && okB) ce = cudaMemcpyAsync(PB, a.Pi + 3 * t.sB, sizeof(PB), cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
            if (ce != cudaSuccess) return (int)ce;
            auto dist = [&](const float *q) { double s2 = 0; for (int k = 0; k < 3; ++k) s2 += ((double)q[k] - P0[k]) * ((double)q[k] - P0[k]); return sqrt(s2); };
            if (okA) dA = dist(PA);
            if (okB) dB = dist(PB);
            choose(dA, dB, okA, okB);
            sc = {a.Pi, a.I1, a.I2, a.I3, lane_axis, tA, lpa, span_m};
        }
        if (const char *e = getenv("QUPS_B200_TILE")) { // "tA,lpa" override for experiments
            int v1 = 0, v2 = 0;
            if (sscanf(e, "%d,%d", &v1, &v2) == 2 && v2 >= 1 && v2 <= 32 && (v2 & (v2 - 1)) == 0 && v1 >= v2 && v1 / v2 <= kCW &&
                v1 % v2 == 0 && ((v1 / v2) & (v1 / v2 - 1)) == 0) { tA = (uint32_t)v1; lpa = (uint32_t)v2; }
        }
        t.tA = tA; t.lpa = lpa; t.tB = kTilePix / tA;
    }
    t.tilesA = (t.IA + t.tA - 1) / t.tA;
    t.tilesB = (t.IB + t.tB - 1) / t.tB;
    // closed-form apodization: 1 = transmit weights only, 2 = receive weights (shared-memory table) [+ transmit weights]
    const int fused = !a.fused ? 0 : (a.fa.rx_kind != AP_RX_NONE ? 2 : (a.fa.tx_kind != AP_TX_NONE ? 1 : 0));
    t.fa = a.fa;
    t.I3 = (uint32_t)a.I3;
    // ---- ring geometry: slot length (samples) x depth.  The window one tile needs is at most ~ 2 fs/c x its spatial extent;
    // fine grids take 4 slots of 128 samples, coarse grids trade depth for length (2 x 256, or 2 x 512 at one CTA per SM)
    uint32_t wmax = QUPS_WMAX, stages = kStages;
    {
        float cinv_h = 0.f;
        static thread_local const void *c_ptr = nullptr;
        static thread_local float c_val = 0.f;
        if (span_m > 0.0 && hinted) cinv_h = (float)(1.0 / a.c_hint);
        else if (span_m > 0.0) { // (span_m stays 0 when the probe was skipped)
            if (c_ptr == a.cinv && c_val > 0.f) cinv_h = c_val;
            else if (!noprobe && cudaMemcpyAsync(&cinv_h, a.cinv, sizeof(float), cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess) { c_ptr = a.cinv; c_val = cinv_h; }
        }
        // upper bound of the window: every pixel step changes the round-trip path by at most twice its length
        const double west = (cinv_h > 0.f) ? 2.0 * (double)a.fs * cinv_h * span_m + 8.0 : 0.0;
        if (west > 256) { wmax = 512; stages = 2; }
        else if (west > 170) { wmax = 256; stages = 2; }
        else if (west > 128) { wmax = 170; stages = 3; }
    }
    if (const char *e = getenv("QUPS_B200_WMAX")) { int v = atoi(e); if (v >= 8 && v <= 1024) wmax = (uint32_t)(v & ~1); }
    if (const char *e = getenv("QUPS_B200_STAGES")) { int v = atoi(e); if (v >= 2 && v <= kStages) stages = (uint32_t)v; }
    wmax &= ~1u;
    const size_t budget = (wmax > 256 ? 200 * 1024 : 200 * 1024 / QUPS_MINBLOCKS);
    while (wmax > 16 && tiled_smem_bytes(t.N, t.M, wmax, fused, stages) > budget) wmax -= 16;
    t.wmax = wmax;
    t.stages = stages;
    const size_t smem = tiled_smem_bytes(t.N, t.M, wmax, fused, stages);
    const uint64_t tiles = (uint64_t)t.tilesA * t.tilesB * t.IC;
    if (tiles == 0 || tiles > 0x7fffffffull) return (int)cudaErrorInvalidValue;

    t.I1 = (uint32_t)a.I1; t.I2 = (uint32_t)a.I2;
    for (int q = 0; q < 2; ++q) {
        t.ap[q] = (q < a.S) ? reinterpret_cast<const float *>(a.apod) + a.astride[q][5] : nullptr;
        for (int d = 0; d < 5; ++d) t.ast[q][d] = (q < a.S) ? (uint32_t)a.astride[q][d] : 0u;
    }
    void (*kern)(const TiledArgs) = nullptr;
    const int ip = a.interp < 0 ? 0 : (a.interp > 2 ? 2 : a.interp);
#define QUPS_PICK(I_, N_, F_, K_) if (ip == I_ && a.S == N_ && fused == F_ && keep == K_) kern = das_tiled_kernel<I_, N_, F_, K_>;
#define QUPS_PICK_I(I_) QUPS_PICK(I_, 0, 0, 0) QUPS_PICK(I_, 1, 0, 0) QUPS_PICK(I_, 2, 0, 0) QUPS_PICK(I_, 0, 1, 0) QUPS_PICK(I_, 1, 1, 0) \
                        QUPS_PICK(I_, 0, 2, 0) QUPS_PICK(I_, 1, 2, 0) QUPS_PICK(I_, 0, 0, 1) QUPS_PICK(I_, 0, 0, 2) QUPS_PICK(I_, 0, 0, 3)
    QUPS_PICK_I(0) QUPS_PICK_I(1) QUPS_PICK_I(2)
#undef QUPS_PICK_I
#undef QUPS_PICK
    if (lut) {
        kern = nullptr;
        if (a.S == 0 && fused == 0 && keep == 0)
            kern = ip == 0 ? das_tiled_kernel<0, 0, 0, 0, 1> : (ip == 1 ? das_tiled_kernel<1, 0, 0, 0, 1> : das_tiled_kernel<2, 0, 0, 0, 1>);
    }
    if (!kern) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    // Work decomposition: when the pixel tiles alone cannot fill ~4 waves of the 148 SMs (small images, pixel-sharded
    // multi-GPU slabs), split the receive axis across CTAs so the grid stays balanced; partials are summed in order.
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // Candidate splits are scored by a makespan model: CTAs are scheduled dynamically onto sms * MINBLOCKS slots, so with
    // w = grid / slots "waves" the launch takes between ceil(w) (equal CTAs) and w + 1/2 (unequal CTAs, small tail) CTA times;
    // every extra split repeats the per-CTA setup (phase 0 over all M transmits ~ 9 % of one receive tile's work).
    uint32_t nsplit = 1;
    {
        const double slots = (double)(wmax > 256 ? 1 : QUPS_MINBLOCKS) * sms; // long-slot rings run one CTA per SM
        const uint32_t max_split = t.numNT >= 4 ? t.numNT / 2 : 1;   // at least two receive tiles per CTA
        double best = 1e300;
        for (uint32_t ns = 1; ns <= max_split; ++ns) {
            const double w = (double)tiles * ns / slots;
            const double span = 0.5 * ceil(w) + 0.5 * (w + 0.5);
            const double cost = span / w * (1.0 + 0.09 * ns / t.numNT);
            // a further split must win by 3 %: measured on the headline grid (2048 tiles, 6.9 waves) two splits ran 1 % slower
            // than one although the model gives them 1.2 % (second pass of phase 0, two receive working sets in L2, the reduce)
            if (cost < best * 0.97) { best = cost; nsplit = ns; }
        }
    }
    // Plain geometric delays: the per-tile bounds come from das_bounds_kernel (below), so a split costs a CTA almost nothing
    // and finer is better — the tail of a launch is about half a CTA duration.  Measured on one rank's slab of an 8-GPU job
    // (256 tiles): nsplit 4 / 8 / 16 = 9.66 / 9.77 / 9.32 ms (10.15 ms with the makespan model and in-kernel bounds); uneven
    // receive ranges (nsplit 3, 5, 6 of 16 tiles) made stragglers.  So: the smallest DIVISOR of the receive-tile count that
    // gives at least `waves_target` waves of CTAs, else one receive tile per CTA.
    // (with a handful of transmits phase 0 is cheap anyway and the extra launch is not: config C1, one plane wave, 39 -> 69 us)
    const bool can_bounds = (!lut || !getenv("QUPS_B200_LUT_NOBOUNDS")) && fused == 0 && (keep == 0 || keep == 3) && t.M >= 16 && !getenv("QUPS_B200_NOBOUNDS");
    if (can_bounds) {
        const double slots = (double)(wmax > 256 ? 1 : QUPS_MINBLOCKS) * sms;
        double waves_target = 48.0; // C2 on one GPU, same box: 12 / 24 / 48 waves = 65.12 / 64.82 / 64.40 ms (end to end 71.98 / 71.98 / 71.25)
        if (const char *ew = getenv("QUPS_B200_WAVES")) { const double v = atof(ew); if (v >= 1.0) waves_target = v; }
        nsplit = 1;
        for (uint32_t d = 1; d <= t.numNT; ++d) {
            if (t.numNT % d) continue;
            nsplit = d;
            if ((double)tiles * d >= waves_target * slots) break;
        }
    }
    if (keep == 1 || keep == 2) nsplit = 1;           // kept apertures: each CTA is the only writer of its pixels
    if (nsplit < 1) nsplit = 1;
    if (const char *e2 = getenv("QUPS_B200_NSPLIT")) { int v = atoi(e2); if (v >= 1 && (uint32_t)v <= t.numNT) nsplit = (uint32_t)v; }
    if (keep == 1 || keep == 2) nsplit = 1;
    if (tiles * nsplit > 0x7fffffffull) nsplit = 1;
    t.nsplit = nsplit;
    // measured at C2 (8-way split): DRAM reads 22.6 -> 2.60 GB per launch (compulsory: 1.07 GB), writes 1.24 -> 0.46 GB, time
    // 64.8 -> 65.1 ms (within run-to-run spread): on by default, QUPS_B200_SPLIT_MAJOR=0 restores the tile-major order
    t.split_major = 1;
    if (const char *es = getenv("QUPS_B200_SPLIT_MAJOR")) t.split_major = atoi(es) != 0;
    // launch order: deepest tiles first when a CTA is a whole tile (the expensive ones lead, the cheap shallow ones fill the
    // tail: 77.1 -> 73.4 ms in round 1); with the split-major fine-grained decomposition the natural order is better
    // (same box, C2: 64.87 -> 64.03 ms; one rank's 8-GPU slab: 9.06 -> 8.83 ms)
    t.rev = nsplit > 1 ? 0 : 1;
    if (const char *e3 = getenv("QUPS_B200_TILE_REV")) t.rev = atoi(e3) != 0;
    t.I = a.I;
    t.part = nullptr;
    t.partP = nullptr;
    if (nsplit > 1 || keep == 3) {
        e = ws_alloc((void **)&t.part, sizeof(float2) * a.I * nsplit, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (keep == 3) {
        e = ws_alloc((void **)&t.partP, sizeof(float) * a.I * nsplit, st);
        if (e != cudaSuccess) { ws_free(t.part, st); return (int)e; }
    }
    if ((keep == 1 || keep == 2) && !t.accumulate) { // every stage adds into y(:, n | m)
        e = cudaMemsetAsync(t.y, 0, sizeof(float2) * a.I * (keep == 1 ? a.M : a.N), st);
        if (e != cudaSuccess) return (int)e;
    }
    t.bounds = nullptr;
    const bool use_bounds = nsplit > 1 && can_bounds;
    if (use_bounds) {
        const size_t per_tile = sizeof(int) * (4 * (size_t)t.M + 2 * (size_t)t.N);
        e = ws_alloc((void **)&t.bounds, per_tile * tiles, st);
        if (e != cudaSuccess) { if (t.part) ws_free(t.part, st); return (int)e; }
        if (per_tile > 48 * 1024) e = cudaFuncSetAttribute(das_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_tile);
        if (e == cudaSuccess) {
            das_bounds_kernel<<<(unsigned)tiles, kCW * 32, per_tile, st>>>(t);
            count_launch();
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) { ws_free(t.bounds, st); if (t.part) ws_free(t.part, st); return (int)e; }
    }
    kern<<<(unsigned)(tiles * nsplit), threads_of(ip), smem, st>>>(t);
    count_launch();
    e = cudaGetLastError();
#if QUPS_STATS
    {
        unsigned long long h[12] = {0};
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_stats, sizeof(h));
        fprintf(stderr, "[das_tiled stats] tile %u x %u lpa %u wmax %u stages %u nsplit %u grid %llu\n", t.tA, t.tB, t.lpa, t.wmax, t.stages, t.nsplit, (unsigned long long)(tiles * nsplit));
        fprintf(stderr, "[das_tiled stats] traces FAST %llu SKIP %llu SLOW %llu EDGE %llu split %llu | stages all-fast %llu general %llu (empty %llu)\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        fprintf(stderr, "[das_tiled stats] general stages: FAST+SKIP only %llu, with dual-window %llu, with EDGE %llu, with SLOW %llu\n", h[8], h[9], h[10], h[11]);
        unsigned long long z[12] = {0};
        cudaMemcpyToSymbol(g_stats, z, sizeof(z));
    }
#endif
    if (keep == 3) {
        if (e == cudaSuccess) {
            const uint64_t g = (a.I + 255) / 256;
            das_cf_reduce_kernel<<<(unsigned)(g < 4096 ? g : 4096), 256, 0, st>>>(t.y, a.cf, t.part, t.partP, a.I, nsplit, (float)a.N);
            count_launch();
            e = cudaGetLastError();
        }
        ws_free(t.partP, st);
        ws_free(t.part, st);
    } else if (nsplit > 1) {
        if (e == cudaSuccess) {
            const uint64_t g = (a.I + 255) / 256;
            das_reduce_kernel<<<(unsigned)(g < 4096 ? g : 4096), 256, 0, st>>>(t.y, t.part, a.I, nsplit, t.accumulate);
            count_launch();
            e = cudaGetLastError();
        }
        ws_free(t.part, st);
    }
    if (t.bounds) ws_free(t.bounds, st);
    return (int)e;
}

} // namespace qups
