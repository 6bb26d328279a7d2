// fft_smem.cuh — hand-written shared-memory FFT building blocks shared by chd_prep.cu (hilbert) and refocus.cu (REFoCUS).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace qups {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ---- shared-memory FFT: register-blocked radix-8 passes, in place, never permuting --------------------------------------
// Three consecutive radix-2 stages are fused into one pass: a thread loads the 8 elements B + k' + j*q (j = 0..7) of its
// butterfly group into registers, runs the three stages there (the twiddles of the group are w[k'] of each stage times
// the constant rotations exp(-i*pi*{0, 1/4, 1/2, 3/4})) and stores them back in place.  log2(n) stages take
// ceil(log2(n) / 3) passes over shared memory instead of log2(n) (n = 2048: 4 instead of 11; the radix-2 version spent
// 93 % of its time in shared-memory wavefronts, profiles/r1_chd_prep_hilbert_ncu_summary.md).
// The FORWARD transform is decimation-in-frequency (natural order in, bit-reversed out), the INVERSE decimation-in-time on
// bit-reversed input (natural order out); everything in between (hilbert weights, Bluestein's spectrum product) is
// element-wise and indexes by the bit-reversed position.  The fused groups are the SAME radix-2 butterflies in the same
// order, so the output order is exactly that of the radix-2 transform.
// Layout: element i lives at s[padi(i)], two complex pad slots after every 16 elements: the strided accesses of the
// short-span passes (q = 4: lanes 32 elements apart; q = 1: a thread owns 8 contiguous elements, LDS.64 at a 64-byte lane
// stride) would otherwise put up to 16 lanes on one bank; with the pad every half-warp touches 32 distinct banks.
// Twiddles: one contiguous table PER STAGE, tw[half - 1 + k] = exp(-i*pi*k/half), half = 1, 2, 4 .. n/2 (n - 1 entries).
__device__ __forceinline__ uint32_t padi(uint32_t i) { return i + ((i >> 4) << 1); }
__host__ __device__ constexpr size_t padded_len(size_t n) { return n + ((n >> 4) << 1) + 2; }

__device__ inline void fft_twiddles(float2 *tw, uint32_t n) {
    for (uint32_t j = threadIdx.x; j + 1 < n; j += blockDim.x) {
        const uint32_t half = 1u << (31 - __clz(j + 1)), k = j + 1 - half;
        float sn, cs;
        sincospif(-(float)k / (float)half, &sn, &cs); // exact dyadic argument
        tw[j] = make_float2(cs, sn);
    }
    __syncthreads();
}

// v * exp(-/+ i*pi*e8/4) for the rotations inside a fused group (kk/halfj = e8/4 with kk < halfj <= 4; folded at compile time)
template <bool CONJ> __device__ __forceinline__ float2 rotc(float2 v, int e8) { // e8 = angle in quarters of pi: 0 .. 3
    const float c = 0.70710678118654752440f;
    if (e8 == 0) return v;
    if (e8 == 2) return CONJ ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
    if (e8 == 1) return CONJ ? make_float2((v.x - v.y) * c, (v.x + v.y) * c) : make_float2((v.x + v.y) * c, (v.y - v.x) * c);
    return CONJ ? make_float2((-v.x - v.y) * c, (v.x - v.y) * c) : make_float2((v.y - v.x) * c, (-v.x - v.y) * c);
}
// R fused stages on the 2^R elements of one butterfly group held in registers; q = element stride of the group, kp = k'
template <int R, bool UNIT = false> __device__ __forceinline__ void dif_regs(float2 (&e)[1 << R], const float2 *tw, uint32_t q, uint32_t kp) {
#pragma unroll
    for (int s = R; s >= 1; --s) {
        const int halfj = 1 << (s - 1);
        const float2 w = UNIT ? make_float2(1.f, 0.f) : tw[halfj * q - 1 + kp]; // UNIT: q = 1, k' = 0 -> every twiddle is exp(0)
#pragma unroll
        for (int jj = 0; jj < (1 << (R - 1)); ++jj) {
            const int kk = jj & (halfj - 1), j0 = ((jj >> (s - 1)) << s) + kk, j1 = j0 + halfj;
            const float2 a = e[j0], c = e[j1], d = make_float2(a.x - c.x, a.y - c.y);
            e[j0] = make_float2(a.x + c.x, a.y + c.y);
            e[j1] = rotc<false>(UNIT ? d : cmulf(d, w), kk * (4 / halfj));
        }
    }
}
template <int R, bool UNIT = false> __device__ __forceinline__ void dit_regs(float2 (&e)[1 << R], const float2 *tw, uint32_t q, uint32_t kp) { // unscaled inverse
#pragma unroll
    for (int s = 1; s <= R; ++s) {
        const int halfj = 1 << (s - 1);
        float2 w = UNIT ? make_float2(1.f, 0.f) : tw[halfj * q - 1 + kp];
        w.y = -w.y;
#pragma unroll
        for (int jj = 0; jj < (1 << (R - 1)); ++jj) {
            const int kk = jj & (halfj - 1), j0 = ((jj >> (s - 1)) << s) + kk, j1 = j0 + halfj;
            const float2 a = e[j0], t = rotc<true>(UNIT ? e[j1] : cmulf(e[j1], w), kk * (4 / halfj));
            e[j0] = make_float2(a.x + t.x, a.y + t.y);
            e[j1] = make_float2(a.x - t.x, a.y - t.y);
        }
    }
}
// one pass: the R stages whose top stage is `st` (block size 2^st), load / store through functors (shared or global)
template <int R, bool INV, class Ld, class St>
__device__ __forceinline__ void fft_group(uint32_t n, uint32_t st, const float2 *tw, Ld ld, St stf) {
    const uint32_t q = 1u << (st - R), tasks = n >> R;
    for (uint32_t task = threadIdx.x; task < tasks; task += blockDim.x) {
        const uint32_t kp = task & (q - 1), base = ((task >> (st - R)) << st) + kp;
        float2 e[1 << R];
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) e[j] = ld(base + j * q);
        if constexpr (R == 3) {
            // radix-8 form: a twiddle-free 8-point transform (constant rotations only) and ONE twiddle per element,
            // W^(k' m) = exp(-i*pi*k'*m/(4q)) for the element holding output m = bitrev3(j) — 7 complex multiplies per
            // group instead of the 12 of three separate radix-2 stages.  The values come from the half = 4q stage table
            // (index k'*m < 7q: the second half of the circle is the negated first)
            const float2 *t4 = tw + (4 * q - 1);
            auto W = [&](uint32_t m) -> float2 {
                const uint32_t idx = kp * m;
                float2 w = t4[idx < 4 * q ? idx : idx - 4 * q];
                if (idx >= 4 * q) { w.x = -w.x; w.y = -w.y; }
                if (INV) w.y = -w.y;
                return w;
            };
            constexpr int br3[8] = {0, 4, 2, 6, 1, 5, 3, 7};
            if (INV) {
#pragma unroll
                for (int j = 1; j < 8; ++j) e[j] = cmulf(e[j], W(br3[j]));
                dit_regs<3, true>(e, tw, q, kp);
            } else {
                dif_regs<3, true>(e, tw, q, kp);
#pragma unroll
                for (int j = 1; j < 8; ++j) e[j] = cmulf(e[j], W(br3[j]));
            }
        } else {
            if (INV) dit_regs<R>(e, tw, q, kp); else dif_regs<R>(e, tw, q, kp);
        }
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) stf(base + j * q, e[j]);
    }
}
// the bottom group (stages rb .. 1): forward stages, an element-wise spectrum operation `mid(position, value)`, inverse stages
template <int R, class Ld, class St, class Mid>
__device__ __forceinline__ void fft_bottom(uint32_t n, const float2 *tw, Ld ld, St stf, Mid mid) {
    for (uint32_t task = threadIdx.x; task < (n >> R); task += blockDim.x) {
        const uint32_t base = task << R;
        float2 e[1 << R];
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) e[j] = ld(base + j);
        dif_regs<R, true>(e, tw, 1, 0);
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) e[j] = mid(base + j, e[j]);
        dit_regs<R, true>(e, tw, 1, 0);
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) stf(base + j, e[j]);
    }
}
__device__ __forceinline__ uint32_t bottom_bits(uint32_t log2n) { const uint32_t r = log2n % 3; return r ? r : (log2n ? 3u : 0u); }

// whole transforms on a padded shared-memory array (Bluestein's inner FFTs)
__device__ inline void fft_fwd_dif(float2 *s, const float2 *tw, uint32_t n, uint32_t log2n) {
    auto ld = [&](uint32_t i) { return s[padi(i)]; };
    auto stf = [&](uint32_t i, float2 v) { s[padi(i)] = v; };
    const uint32_t rb = bottom_bits(log2n);
    for (uint32_t st = log2n; st > rb; st -= 3) { fft_group<3, false>(n, st, tw, ld, stf); __syncthreads(); }
    if (rb == 3) fft_group<3, false>(n, 3, tw, ld, stf); else if (rb == 2) fft_group<2, false>(n, 2, tw, ld, stf); else if (rb == 1) fft_group<1, false>(n, 1, tw, ld, stf);
    __syncthreads();
}
__device__ inline void fft_inv_dit(float2 *s, const float2 *tw, uint32_t n, uint32_t log2n) { // unscaled
    auto ld = [&](uint32_t i) { return s[padi(i)]; };
    auto stf = [&](uint32_t i, float2 v) { s[padi(i)] = v; };
    const uint32_t rb = bottom_bits(log2n);
    if (rb == 3) fft_group<3, true>(n, 3, tw, ld, stf); else if (rb == 2) fft_group<2, true>(n, 2, tw, ld, stf); else if (rb == 1) fft_group<1, true>(n, 1, tw, ld, stf);
    __syncthreads();
    for (uint32_t st = rb + 3; st <= log2n; st += 3) { fft_group<3, true>(n, st, tw, ld, stf); __syncthreads(); }
}

// ---- compile-time specialised passes (LG = log2 of the length) -------------------------------------------------------
// With the length known at compile time every index of a pass is `padi(base) + constant`: for a group stride q the
// padded offset of element j is j*q + 2*((j*q) >> 4) (k' < q never carries into the 16-element pad period), so the loads
// and stores take immediate offsets and the integer work per element disappears (ncu on the run-time version: ALU pipe
// 58 %, FMA 30 %, issue 74 % — index arithmetic, not butterflies, filled the issue slots).
template <int I, int N, class F> __device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}
__host__ __device__ constexpr uint32_t pad_off(uint32_t e) { return e + ((e >> 4) << 1); }

// radix-8 pass with top stage ST of a 2^LG transform.  ld(i, p) / stf(i, p, v): i = logical index, p = padded index (a
// global-memory functor ignores p, a shared-memory one ignores i)
template <int LG, int ST, bool INV, class Ld, class St>
__device__ __forceinline__ void fft_group_c(const float2 *tw, Ld ld, St stf) {
    constexpr uint32_t q = 1u << (ST - 3), tasks = (1u << LG) >> 3;
    for (uint32_t task = threadIdx.x; task < tasks; task += blockDim.x) {
        const uint32_t kp = task & (q - 1), base = ((task >> (ST - 3)) << ST) + kp, pb = padi(base);
        float2 e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = ld(base + j * q, pb + pad_off(j * q));
        // twiddles W^m = exp(-i*pi*k'*m/(4q)): m = 1, 2, 4, 3 from the stage tables (no wrap: 3k' < 4q), 5..7 as products
        float2 w[8];
        w[1] = tw[4 * q - 1 + kp]; w[2] = tw[2 * q - 1 + kp]; w[4] = tw[q - 1 + kp]; w[3] = tw[4 * q - 1 + 3 * kp];
        if (INV) { w[1].y = -w[1].y; w[2].y = -w[2].y; w[3].y = -w[3].y; w[4].y = -w[4].y; }
        w[5] = cmulf(w[4], w[1]); w[6] = cmulf(w[4], w[2]); w[7] = cmulf(w[4], w[3]);
        constexpr int br3[8] = {0, 4, 2, 6, 1, 5, 3, 7};
        if (INV) {
#pragma unroll
            for (int j = 1; j < 8; ++j) e[j] = cmulf(e[j], w[br3[j]]);
            dit_regs<3, true>(e, tw, q, kp);
        } else {
            dif_regs<3, true>(e, tw, q, kp);
#pragma unroll
            for (int j = 1; j < 8; ++j) e[j] = cmulf(e[j], w[br3[j]]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) stf(base + j * q, pb + pad_off(j * q), e[j]);
    }
}
// bottom group of R stages: forward, element-wise `mid(position, value)`, inverse — on 2^R contiguous elements
template <int LG, int R, class Ld, class St, class Mid>
__device__ __forceinline__ void fft_bottom_c(const float2 *tw, Ld ld, St stf, Mid mid) {
    for (uint32_t task = threadIdx.x; task < ((1u << LG) >> R); task += blockDim.x) {
        const uint32_t base = task << R, pb = padi(base);
        float2 e[1 << R];
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) e[j] = ld(base + j, pb + j);
        dif_regs<R, true>(e, tw, 1, 0);
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) e[j] = mid(base + j, e[j]);
        dit_regs<R, true>(e, tw, 1, 0);
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) stf(base + j, pb + j, e[j]);
    }
}
// forward transform, spectrum operation, inverse transform of one 2^LG trace: ldg / stg read and write the caller's
// (global) data, s is the padded shared work array.  Passes over shared memory: 2 * (groups - 1).
template <int LG, class LdG, class StG, class Mid>
__device__ __forceinline__ void fft_roundtrip_c(float2 *s, const float2 *tw, LdG ldg, StG stg, Mid mid) {
    constexpr int RB = (LG % 3) ? (LG % 3) : 3, NG = (LG - RB) / 3; // bottom stages, number of radix-8 groups above them
    auto lds = [&](uint32_t, uint32_t p) { return s[p]; };
    auto sts = [&](uint32_t, uint32_t p, float2 v) { s[p] = v; };
    auto ldG = [&](uint32_t i, uint32_t) { return ldg(i); };
    auto stG = [&](uint32_t i, uint32_t, float2 v) { stg(i, v); };
    if constexpr (NG == 0) {
        fft_bottom_c<LG, RB>(tw, ldG, stG, mid);
    } else {
        fft_group_c<LG, LG, false>(tw, ldG, sts);
        __syncthreads();
        static_for<1, NG>([&](auto I) { fft_group_c<LG, LG - 3 * decltype(I)::value, false>(tw, lds, sts); __syncthreads(); });
        fft_bottom_c<LG, RB>(tw, lds, sts, mid);
        __syncthreads();
        static_for<1, NG>([&](auto I) { fft_group_c<LG, RB + 3 * decltype(I)::value, true>(tw, lds, sts); __syncthreads(); });
        fft_group_c<LG, LG, true>(tw, lds, stG);
        __syncthreads(); // the next trace's first pass overwrites s
    }
}

// chirp c[n] = exp(-i*pi*n^2/L) with n^2 reduced mod 2L in integers
__device__ __forceinline__ float2 chirp(uint64_t n, uint64_t L) {
    const uint64_t r = (n * n) % (2 * L);
    float sn, cs;
    sincospif(-(float)((double)r / (double)L), &sn, &cs);
    return make_float2(cs, sn);
}

// DFT of length L (arbitrary) of s (padded layout) via Bluestein; cb = FFT of the wrapped conjugate chirp in bit-reversed
// order (plain layout, read element-wise), ch = chirp
__device__ inline void dft_bluestein(float2 *s, const float2 *tw, const float2 *cb, const float2 *ch, uint64_t L, uint32_t nfft, uint32_t log2n, bool inverse) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    // inverse DFT = conj(DFT(conj(x)))
    for (uint32_t i = tid; i < nfft; i += nt) {
        float2 v = make_float2(0.f, 0.f);
        if (i < L) {
            v = s[padi(i)];
            if (inverse) v.y = -v.y;
            v = cmulf(v, ch[i]);
        }
        s[padi(i)] = v;
    }
    __syncthreads();
    fft_fwd_dif(s, tw, nfft, log2n);
    for (uint32_t i = tid; i < nfft; i += nt) s[padi(i)] = cmulf(s[padi(i)], cb[i]); // both spectra in bit-reversed order
    __syncthreads();
    fft_inv_dit(s, tw, nfft, log2n);
    const float sc = 1.0f / (float)nfft;
    for (uint32_t i = tid; i < L; i += nt) {
        const float2 u = s[padi(i)];
        float2 v = cmulf(make_float2(u.x * sc, u.y * sc), ch[i]);
        if (inverse) v.y = -v.y;
        s[padi(i)] = v;
    }
    __syncthreads();
}

} // namespace qups
