/* qups_oracle.h — CPU oracle for the QUPS DAS / wsinterpd2 / greens hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker or the timed
 * CPU baseline — never as the shipped compute path.
 *
 * PARITY STATUS: "parity unpinned" for exact image values. The reference's
 * CPU path is MATLAB (kern/das_spec.m + built-in interp1); neither MATLAB nor
 * Octave exists in the build image, and the reference ships no golden
 * vectors for DAS. This file restates the algorithm line
 * by line (citations inline); it is pinned by (i) an independent NumPy
 * restatement (oracle/oracle_np.py), (ii) the reference's own physical
 * known-answer tests restated in tests/ (test/BFTest.m:230-317,
 * test/SimTest.m:299-324), (iii) the reference's interpTest data generator
 * (test/interpTest.m:28-47), and (iv) on the GPU box, the reference's own
 * src/bf.cu DASf kernel compiled unmodified into oracle/_ref/ (interior
 * samples only; it has different edge semantics).
 */
#ifndef QUPS_ORACLE_H
#define QUPS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_DAS = 0, ORACLE_SYN = 1, ORACLE_MUL = 2, ORACLE_BF = 3, ORACLE_DELAYS = 4 };
enum { ORACLE_NEAREST = 0, ORACLE_LINEAR = 1, ORACLE_CUBIC = 2, ORACLE_LANCZOS3 = 3 };

#define ORACLE_MAX_APOD 8

/* das_spec(fun, Pi, Pr, Pv, Nv, x, t0, fs, c, ...)   kern/das_spec.m:1 */
typedef struct {
    int32_t fun;            /* ORACLE_DAS.. */
    int32_t interp;         /* ORACLE_NEAREST.. */
    int32_t VS, DV;         /* 'plane-waves' => VS=0 ; 'diverging-waves' => DV=1   :118-126 */
    int32_t tpose;          /* 'transpose': data is T x M x N                        :139-141 */
    int32_t S;              /* number of apodization arrays (0 => none)              :133-135 */
    int32_t apod_complex;   /* apod arrays are complex (interleaved) else real */
    int32_t pad_;
    uint64_t I[3], N, M, T, F;
    double fs, fmod;
    const void *Pi;         /* 3 x I   */
    const void *Pr;         /* 3 x N   */
    const void *Pv;         /* 3 x M   */
    const void *Nv;         /* 3 x M   */
    const void *x;          /* complex T x N x M x F (or T x M x N x F if tpose) */
    const void *t0;         /* M values (host expands a scalar)                      :424-425 */
    const void *cinv;       /* broadcastable to I1 x I2 x I3 x N x M                 :170 */
    uint64_t csz[5];
    const void *apod[ORACLE_MAX_APOD];
    uint64_t asz[ORACLE_MAX_APOD][5];
} oracle_das_args;

/* canonical (I,N,M) form of wsinterpd2   kern/wsinterpd2.m:1 */
typedef struct {
    int32_t interp, sum_n, sum_m, w_complex;
    uint64_t I, N, M, T;
    double omega;           /* imaginary part of omega (2*pi*fmod/fs) */
    const void *x;          /* complex T x N x M */
    const void *t1; uint64_t t1sz[3];   /* each dim 1 or full (I,N,M) */
    const void *t2; uint64_t t2sz[3];
    const void *w;  uint64_t wsz[3];
} oracle_ws2_args;

/* greens CPU math   src/UltrasoundSystem.m:720-863 */
typedef struct {
    int32_t interp, pad_;
    uint64_t S, N, M, T, K, E;
    int64_t n0;             /* first output sample index: t = (n0 : n0+T-1)          :613-615 */
    double c0, fs, fsr, R0, wv_t0;
    const void *ps;         /* 3 x S scatterer positions */
    const void *amp;        /* S real amplitudes */
    const void *pn;         /* 3 x N x E receive (sub-)element positions */
    const void *pv;         /* 3 x M x E transmit (sub-)element positions */
    const void *kern;       /* complex K : convolved waveform samples                :584-588 */
} oracle_greens_args;

void oracle_interp1_f(const float *v, long T, float xq, int method, float *yr, float *yi);
void oracle_interp1_d(const double *v, long T, double xq, int method, double *yr, double *yi);
int oracle_das_f(const oracle_das_args *A, float *y);
int oracle_das_d(const oracle_das_args *A, double *y);
int oracle_wsinterpd2_f(const oracle_ws2_args *A, float *y);
int oracle_wsinterpd2_d(const oracle_ws2_args *A, double *y);
int oracle_greens_f(const oracle_greens_args *A, float *x);
int oracle_greens_d(const oracle_greens_args *A, double *x);
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
