/* qups_b200.h — C ABI of libqups_b200.so: the B200-native (sm_100a) drop-in for
 * the one data-parallel hot path of thorstone25/qups — delay-and-sum
 * beamforming and the Green's-function point-scatterer simulator.
 *
 * Every entry point replaces one `parallel.gpu.CUDAKernel.feval` call site of
 * the reference (MATLAB -> PTX-by-name).  Argument order and meaning follow the
 * reference kernels' argument lists so that the MATLAB-side glue
 * (mex/qups_b200_mex.cu, matlab/*.m) is a direct forward:
 *
 *   qups_das          <-  kern/das_spec.m:371-373   k.feval(yg,Pi,Pr,Pv,Nv,apod,cinv,[cstride,astride],x,[fs,fmod])
 *                         src/bf.cu:144-172         DAS / DASf / DASh
 *   qups_delays       <-  kern/das_spec.m:376-377   fun='delays'  (src/bf.cu:209-298 delays / delaysf)
 *   qups_wsinterpd2   <-  kern/wsinterpd2.m:226-235 k.feval(y,w,x,t1,t2,sizes,iflags,strides,flag,extrapval,omega)
 *                         src/interpd.cu:451-476    wsinterpd2 / wsinterpd2f / wsinterpd2h
 *   qups_wsinterpd    <-  kern/wsinterpd.m:205-213  (src/interpd.cu:422-447)
 *   qups_greens       <-  src/UltrasoundSystem.m:718 k.feval(x,ps,as,pn,pv,kn,sb,iblock,[t0k,t0x,fs,fsr,cinv,R0],[E,E],flag)
 *                         src/greens.cu:88-122      greens / greensf / greensh
 *   qups_convd        <-  kern/convd.m:194-201      k.feval(z, x, y, sizes)   (src/convd.cu:133-156 conv / convf / convc / convcf)
 *   qups_modulate     <-  kern/das_spec.m:413-417   x .* exp(2i*pi*fmod.*t)  (CPU-branch convention, fused pre-pass)
 *   qups_*_host       <-  the same calls for callers holding HOST arrays (plain MEX, no gpuArray)
 *
 * Conventions
 *   - column-major (MATLAB) layouts, complex = interleaved (re,im)
 *   - all array pointers are DEVICE pointers unless the function name ends in
 *     _host; small descriptor arrays (strides, sizes) are HOST pointers
 *   - buffers are caller-owned; the library never frees or retains them
 *   - return value: 0 on success, negative qups_status on error;
 *     qups_last_error() returns a thread-local message
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); calls are
 *     asynchronous w.r.t. the host unless stated
 *   - re-entrant.  State the library keeps between calls, all of it invisible in results: the thread-local error text,
 *     launch counter and last-kernel name; one stream-ordered memory pool per device for scratch (created on first use,
 *     memory is returned to it, not to the driver, at synchronisation points); and, for the *_host entry points only,
 *     per-host-thread device staging buffers + two streams (the ~1 GB cube buffer is kept between calls on purpose;
 *     qups_host_release() frees them).  No result depends on previous calls (see pitch_hint for the one perf-only cache).
 *   - results follow the reference's CPU semantics (kern/das_spec.m CPU branch +
 *     MATLAB interp1(…, extrapval=0)), not the edge quirks of src/interpd.cu
 */
#ifndef QUPS_B200_H
#define QUPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QUPS_B200_VERSION 200

#if defined(__GNUC__)
#define QUPS_API __attribute__((visibility("default")))
#else
#define QUPS_API
#endif

typedef void *qups_stream_t; /* cudaStream_t */

typedef enum {
    QUPS_OK = 0,
    QUPS_ERR_INVALID = -1,     /* bad argument / inconsistent sizes          */
    QUPS_ERR_CUDA = -2,        /* CUDA runtime error (see qups_last_error)   */
    QUPS_ERR_UNSUPPORTED = -3, /* valid request this build does not handle   */
    QUPS_ERR_ALLOC = -4        /* device / host allocation failed            */
} qups_status;

typedef enum { QUPS_F32 = 0, QUPS_F16 = 1, QUPS_F64 = 2 } qups_dtype;
/* interpolation ids == bits 0-2 of QUPS_BF_FLAG (kern/das_spec.m:198-203) */
typedef enum { QUPS_NEAREST = 0, QUPS_LINEAR = 1, QUPS_CUBIC = 2, QUPS_LANCZOS3 = 3 } qups_interp;

/* QUPS_BF_FLAG bits (kern/das_spec.m:199-213, src/bf.cu:127-137) */
#define QUPS_FLAG_INTERP_MASK 7
#define QUPS_FLAG_KEEP_RX 8
#define QUPS_FLAG_KEEP_TX 16
#define QUPS_FLAG_TRANSPOSE 32

typedef enum { QUPS_PATH_AUTO = 0, QUPS_PATH_GENERIC = 1, QUPS_PATH_TILED = 2 } qups_path;

/* ---- DAS -------------------------------------------------------------- */
/* Image of the reference's __constant__ symbols QUPS_{I1,I2,I3,N,M,T,S,VS,DV,
 * BF_FLAG} (src/sizes.cu:6-53, src/bf.cu:45-47) set at kern/das_spec.m:294-298. */
typedef struct {
    uint32_t struct_size; /* = sizeof(qups_das_params) */
    int32_t dtype;        /* qups_dtype of x / apod / y ('f'/'h'/'' kernel suffix, kern/das_spec.m:218-222) */
    uint64_t I1, I2, I3;  /* pixel grid; I = I1*I2*I3 */
    uint64_t N, M, T;     /* receives, transmits, time samples */
    uint64_t F;           /* frames looped by the library (kern/das_spec.m:371); 0 -> 1 */
    uint64_t S;           /* number of apodization arrays (0 -> none, i.e. apod = 1) */
    int32_t flag;         /* QUPS_BF_FLAG */
    int32_t vs, dv;       /* QUPS_VS (0 = plane waves), QUPS_DV (1 = diverging waves) */
    int32_t apod_real;    /* extension: apod arrays are real (the reference forces complex, :237-243) */
    int32_t y_f32;        /* extension: with dtype F16 write float2 output instead of half2 */
    int32_t path;         /* qups_path; AUTO picks the tiled kernel when eligible */
    int32_t accumulate;   /* extension: y += result instead of y = result (transmit-chunked pipelines) */
    int32_t host_chunks;  /* qups_das_host only: transmit chunks of the copy/compute pipeline (0 = automatic) */
    int32_t y_device;     /* qups_das_host only: y is a DEVICE pointer on `device` and the image stays there (no read-back):
                             multi-GPU transmit partitions reduce the partial images with NCCL before one rank reads the sum */
    int32_t reserved_;    /* 0 */
    double fs;            /* sampling frequency */
    double fmod;          /* modulation frequency (data re-modulated at absolute time, kern/das_spec.m:413-417) */
    uint64_t x_frame_stride; /* complex elements between frames of x; 0 -> T*N*M */
    uint64_t y_frame_stride; /* complex elements between frames of y; 0 -> I*[N]*[M] */
    void *workspace;         /* optional device scratch for the modulated cube (fmod != 0); NULL -> stream-ordered alloc */
    uint64_t workspace_bytes;
    /* Optional tuning hints for the staged kernel (never affect results): distance in metres between neighbouring pixels along
     * I1 and I2 (scan.dz / scan.dx of a ScanCartesian) and the scalar sound speed.  With all three set the launcher picks the
     * tile shape and ring geometry from them and never touches device memory from the host; with any of them 0 it reads three
     * pixel positions and 1/c back once per (pixel pointer, grid size) — a host synchronisation on the first call with a new
     * grid, cached for later calls (the one cache keyed on a caller pointer; a stale entry only costs speed). */
    double pitch_hint[2];
    double c_hint;
} qups_das_params;

/* y      : out, complex, I x [N if keep_rx] x [M if keep_tx] (x F)
 * Pi     : 3 x I real (float for F32/F16, double for F64)
 * Pr     : 3 x N ;  Pv4 : 4 x M (row 4 = t0 per transmit, kern/das_spec.m:361) ; Nv : 3 x M
 * apod   : the S arrays flattened and concatenated (kern/das_spec.m:344-345), complex (or real if apod_real); NULL if S == 0
 * cinv   : real, 1/c broadcastable to I1 x I2 x I3 x N x M
 * acstride (HOST): uint64[6 + 6*S] = [cstride(6), astride(6 x S)] exactly as built at kern/das_spec.m:256-260
 *          (per-dim element strides, 0 for singleton dims; entry 6 of each column = base offset into apod)
 * x      : complex T x N x M (x F)   (T x M x N if QUPS_FLAG_TRANSPOSE)                                              */
QUPS_API int qups_das(const qups_das_params *p, void *y, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
             const void *apod, const void *cinv, const uint64_t *acstride, const void *x, qups_stream_t stream);

/* ---- closed-form apodization -------------------------------- */
/* The reference's apodization generators return dense ND masks that DAS receives as 'apod' arrays
 * (src/UltrasoundSystem.m:4892-5429: apScanline, apMultiline, apTranslatingAperture, apApertureGrowth,
 * apTxParallelogram, apAcceptanceAngle, apCosineAngle).  Each is a closed-form function of (pixel, element)
 * geometry; this block lets DAS evaluate it in-kernel instead of reading 4-8 bytes per (pixel, receive) pair.
 * The weight of a (pixel i, receive n, transmit m) term is  rx(i,n) * tx(i,m) * prod_s apod_s(i,n,m).
 *   rx_kind                       rx_p                         rx_aux (device, fp32)
 *   ACCEPTANCE_ANGLE  (:5303)     [cosd(theta)]                3 x N element normals
 *   COSINE_ANGLE      (:5377)     [90/theta]                   3 x N element normals
 *   APERTURE_GROWTH   (:5165)     [f, Dmax, nonplanar(0|1)]    nonplanar: 2 x N [cosd(ae); sind(ae)] of the element angles
 *   TRANSLATING       (:5074)     [tol_rx]                     N lateral coordinates of the receivers (x, or angle)
 *   tx_kind                       tx_p                         tx_aux (device, fp32)
 *   SCANLINE          (:4892)     [tol]   |xi - xv| <  tol     M lateral coordinates of the transmits (focus x, or angle)
 *   TRANSLATING       (:5074)     [tol]   |xi - xv| <= tol     M lateral coordinates
 *   PARALLELOGRAM     (:5269)     [xlo, xhi] (xdc bounds)      4 x M [sind(th+phi1); cosd(th+phi1); sind(th+phi2); cosd(th+phi2)], 16-byte aligned
 * lat / lat_dim: lateral coordinate of the pixels per index along grid dimension lat_dim (ScanPolar: scan.a along adim);
 * NULL -> the pixel's x coordinate (ScanCartesian).  apMultiline (:4970) is a small I_lat x M matrix: pass it as an array. */
typedef enum {
    QUPS_AP_RX_NONE = 0, QUPS_AP_RX_ACCEPTANCE_ANGLE = 1, QUPS_AP_RX_COSINE_ANGLE = 2, QUPS_AP_RX_APERTURE_GROWTH = 3,
    QUPS_AP_RX_TRANSLATING = 4
} qups_ap_rx_kind;
typedef enum { QUPS_AP_TX_NONE = 0, QUPS_AP_TX_SCANLINE = 1, QUPS_AP_TX_TRANSLATING = 2, QUPS_AP_TX_PARALLELOGRAM = 3 } qups_ap_tx_kind;
typedef struct {
    uint32_t struct_size; /* = sizeof(qups_apod_fused) */
    int32_t rx_kind, tx_kind;
    int32_t lat_dim;
    float rx_p[4], tx_p[4];
    const void *rx_aux, *tx_aux, *lat;
} qups_apod_fused;

/* qups_das with closed-form apodization; `apod` arrays (p->S) still multiply in.  fp32 geometry (dtype F32, or F16 data on
 * the plain-DAS configuration).  Calls outside the staged kernel's envelope materialise the dense weights internally. */
QUPS_API int qups_das_fused(const qups_das_params *p, const qups_apod_fused *apf, void *y, const void *Pi, const void *Pr,
                   const void *Pv4, const void *Nv, const void *apod, const void *cinv, const uint64_t *acstride, const void *x,
                   qups_stream_t stream);

/* DAS and the coherence factor of its per-receive images in ONE pass, without the I x N cube:
 *   b  = DAS(us, chd, 'keep_rx', true);  y = sum(b, rxdim);  cf = cohfac(b, rxdim) = |sum_n b_n|^2 ./ sum_n |b_n|^2 / N
 * (src/UltrasoundSystem.m:3172 + kern/cohfac.m).  y: I complex, cf: I real.  Plain weights only: dtype F32, S = 0, scalar cinv,
 * no kept aperture, fmod = 0, F = 1.  Calls outside the staged kernel's envelope (lanczos3, very large N + M) are served by
 * DAS(keep_rx) into library scratch + the cohfac reduction, same results. */
QUPS_API int qups_das_cohfac(const qups_das_params *p, void *y, void *cf, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                             const void *cinv, const void *x, qups_stream_t stream);
/* The dense array the reference's generator would return: which = 0 -> receive weights I1 x I2 x I3 x NM(=N),
 * which = 1 -> transmit weights I1 x I2 x I3 x 1 x NM(=M); real fp32, or complex (imag 0) when as_complex. */
QUPS_API int qups_apod_generate(const qups_apod_fused *apf, int32_t which, void *out, int32_t as_complex, const void *Pi, const void *Pr,
                       uint64_t I1, uint64_t I2, uint64_t I3, uint64_t NM, qups_stream_t stream);

/* tau(i,n,m) = cinv .* (dv + dr)   (no -t0)   kern/das_spec.m:448-449; tau : real I x N x M */
QUPS_API int qups_delays(const qups_das_params *p, void *tau, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                const void *cinv, const uint64_t *cstride, qups_stream_t stream);

/* Same as qups_das but every array pointer is a HOST pointer; performs H2D, compute, D2H and synchronises.
 * Pinned host memory is used at full PCIe rate; pageable memory works but is slower.
 * When the call is eligible for the staged kernel the transmit axis is cut into chunks whose H2D copies overlap
 * the beamforming of the previous chunk (two streams), so end-to-end time ~ max(copy, compute). */
QUPS_API int qups_das_host(const qups_das_params *p, void *y, const void *Pi, const void *Pr, const void *Pv4, const void *Nv,
                  const void *apod, uint64_t apod_elems, const void *cinv, uint64_t cinv_elems,
                  const uint64_t *acstride, const void *x, int device);

/* qups_das_host keeps its device staging buffers and streams per host thread between calls (the cube is ~1 GB;
 * re-allocating it every call costs more than the beamforming). Release them explicitly with this call. */
QUPS_API void qups_host_release(void);

/* x(t,n,m) *= exp(2i*pi*fmod*(t0(m) + t/fs))  out-of-place; t0 : M reals (device). kern/das_spec.m:413-417 */
QUPS_API int qups_modulate(int32_t dtype, void *xout, const void *x, const void *t0, uint64_t T, uint64_t N, uint64_t M,
                  int32_t transpose, double fs, double fmod, qups_stream_t stream);

/* ---- ChannelData pre-processing -------------------------------- */
/* One pass over the cube doing what the reference scripts do with a chain of ChannelData methods before DAS
 * (example_.m:261-269):  zeropad (src/ChannelData.m:1153-1183) -> hilbert (:935-966) -> downmix (:757-807) ->
 * singleT / halfT cast (:452-483).  Any subset: B = A = 0, hilbert = 0, fmix = 0 disable the steps.
 *   in  : T x K traces (K = N*M*F), element type in_dtype
 *   out : (B+T+A) x K complex fp32 (QUPS_F32) or half2 (QUPS_F16)
 *   t0  : device, n_t0 fp32 start times (NULL -> 0); trace k uses t0[(k / traces_per_t0) % n_t0]
 *         (one t0 per transmit: traces_per_t0 = N, n_t0 = M).  The caller's new t0 is t0 - B/fs.
 * hilbert: analytic signal over the padded length L = B+T+A (MATLAB hilbert(x): fft, [1 2..2 1 0..0], ifft); real part
 *          of complex input is used.  L a power of two up to 16384, or any L <= 4096 (Bluestein). */
typedef enum { QUPS_IN_REAL_F32 = 0, QUPS_IN_CPLX_F32 = 1, QUPS_IN_REAL_I16 = 2, QUPS_IN_REAL_F64 = 3 } qups_prep_in;
typedef struct {
    uint32_t struct_size;
    int32_t in_dtype;   /* qups_prep_in */
    int32_t out_dtype;  /* QUPS_F32 | QUPS_F16 */
    int32_t hilbert;
    uint64_t T, K, B, A;
    uint64_t traces_per_t0, n_t0;
    double fs, fmix;    /* downmix by fmix: x .* exp(-2i*pi*fmix*t) */
} qups_prep_params;
QUPS_API int qups_chd_prep(const qups_prep_params *p, void *out, const void *in, const void *t0, qups_stream_t stream);

/* ---- aperture-domain post-processing -------------------------------- */
/* One-pass reductions along the aperture dimension of a beamformed cube (DAS 'keep_rx' output), viewed as
 * C x A x S complex (A = the reduced dimension `dim`, C / S = product of the dimensions before / after it):
 *   COHFAC         kern/cohfac.m   out real  C x S :  |sum b|^2 / sum |b|^2 / A
 *   DMAS           kern/dmas.m     out cplx  C x S :  exp(1j angle(z)) sqrt(|z|), z = sum_{lag} sum_n b(n) b(n+lag)
 *   PCF            kern/pcf.m      out real  C x S (weights w), out2 real C x S (sf, may be NULL); gamma
 *   SLSC_AVERAGE / SLSC_ENSEMBLE   kern/slsc.m (kdim singleton)   out cplx C x S
 * lags: HOST array of nlags lag values (dmas: only 1..A-1 are used; slsc: L = nlags normalises the average). */
typedef enum { QUPS_APD_COHFAC = 0, QUPS_APD_DMAS = 1, QUPS_APD_PCF = 2, QUPS_APD_SLSC_AVERAGE = 3, QUPS_APD_SLSC_ENSEMBLE = 4 } qups_aperture_op;
typedef struct {
    uint32_t struct_size;
    int32_t dtype; /* QUPS_F32 | QUPS_F64 */
    int32_t op;    /* qups_aperture_op */
    uint32_t nlags;
    uint64_t C, A, S;
    double gamma;
} qups_aperture_params;
QUPS_API int qups_aperture(const qups_aperture_params *p, void *out, void *out2, const void *b, const uint32_t *lags, qups_stream_t stream);

/* ---- wsinterpd / wsinterpd2 ------------------------------------------ */
/* y(l) = sum over dims with ystride==0 of  exp(1i*omega*t) * w(k) * interp1(x(:,v), 1+t, interp, 0),  t = t1(r)+t2(u)
 * Image of the reference argument list (src/interpd.cu:344-349): D broadcast dims of size sizes[d]; dstride is
 * uint64[5*D], column d = element strides of {w, y, t1, t2, x-trace} along dim d (kern/wsinterpd2.m:193-219).
 * The reference reduces with global atomics; this library reduces deterministically inside one thread.
 * dtype F16 = wsinterpd2h / wsinterpdh (src/interpd.cu:422-429,451-458): x, w half2 (w half when w_real), t1 / t2 HALF, y half2
 * (float2 with y_f32); positions, interpolation and sums in fp32.
 * Fast path: the canonical look-up-table delay-and-sum (bfDAS -> bfDASLUT -> sample2sep: dense I x N and I x M tables, both
 * apertures summed, scalar weight, omega = 0, nearest|linear|cubic, fp32 or fp16) runs on the staged DAS kernel in table mode
 * (qups_last_ws2_kernel() == "ws2_tiled"); everything else takes the generic strided kernel. */
typedef struct {
    uint32_t struct_size;
    int32_t dtype;
    uint64_t T;        /* samples per trace */
    uint32_t D;        /* number of broadcast dims (<= 8) */
    int32_t interp;
    int32_t w_real;    /* extension: weights are real */
    int32_t y_f32;
    double omega;      /* imaginary part of omega = 2*pi*fmod/fs  (src/ChannelData.m:1439) */
    uint64_t sizes[8];
    uint64_t dstride[40];
} qups_ws2_params;

QUPS_API int qups_wsinterpd2(const qups_ws2_params *p, void *y, const void *w, const void *x, const void *t1, const void *t2,
                    qups_stream_t stream);
/* single-table variant (kern/wsinterpd.m): dstride columns are {w, y, t, (ignored), x} */
QUPS_API int qups_wsinterpd(const qups_ws2_params *p, void *y, const void *w, const void *x, const void *t, qups_stream_t stream);

/* ---- greens ------------------------------------------------------------ */
/* Image of the scalar pack [t0k,t0x,fs,fsr,cinv,R0], [E,E], flag  (src/UltrasoundSystem.m:718, src/greens.cu:8-27).
 * Output sample s (0-based) is absolute sample index n0 + s, n0 = round(t0k*fs).
 * Semantics follow the CPU path (:797-851): R0 == 0 means "no propagation loss" (the reference GPU kernel
 * divides by R0^2 there, src/greens.cu:84), no 1/R0^2 scaling, sum order = scatterer order as given. */
typedef struct {
    uint32_t struct_size;
    int32_t dtype;
    uint64_t I;       /* scatterers   (QUPS_I) */
    uint64_t S;       /* output time samples (QUPS_S) */
    uint64_t T;       /* kernel (waveform) samples (QUPS_T) */
    uint64_t N, M;    /* receives, transmits */
    uint64_t E;       /* sub-elements per element (both apertures) */
    int64_t n0;       /* first output sample index */
    int32_t interp;
    int32_t y_f32;
    double t0x;       /* waveform start time wv.t0 */
    double fs, fsr;   /* output sampling frequency, kernel/output sampling ratio */
    double c0;        /* sound speed (the CPU path divides by c0, :797-798; the reference GPU pack carries 1/c0) */
    double R0;        /* minimum distance */
} qups_greens_params;

/* y : out complex S x N x M ; Pi : 3 x I scatterer positions ; a : I real amplitudes ;
 * Pr : 3 x N x E ; Pv : 3 x M x E ; kern : complex T
 * Scratch: the default fp32 kernel (fsr == 1) takes (N + M) * E * I * 16 bytes of path-length tables from the library's
 * stream-ordered pool for the duration of the call (82 MB at 10 k scatterers, 256 + 256 elements); above 4 GB, or if the
 * allocation fails, every trace computes its own path lengths instead (same results to 2 ulp of fp64, slower). */
QUPS_API int qups_greens(const qups_greens_params *p, void *y, const void *Pi, const void *a, const void *Pr, const void *Pv,
                const void *kern, qups_stream_t stream);

/* ---- convd ---------------------------------------------------------------- */
/* Batched direct 1-D convolution along one dimension: replaces conv/convf/convc/convcf (src/convd.cu:133-156,
 * launcher kern/convd.m:135-201).  x is C x Lx x S, y is yC x Ly x yS (yC in {1,C}, yS in {1,S}), z is C x Lz x S;
 * shape 0 'full' (Lz = Lx+Ly-1), 1 'same' (Lz = Lx, centred as MATLAB conv), 2 'valid' (Lz = max(Lx-Ly+1, 0)). */
typedef struct {
    uint32_t struct_size;
    int32_t dtype;      /* QUPS_F32 | QUPS_F64 | QUPS_F16 (half / half2 storage, fp32 accumulation: convh / convch) */
    int32_t is_complex; /* interleaved complex data */
    int32_t shape;
    uint64_t C, S, Lx, Ly, yC, yS;
} qups_convd_params;
QUPS_API int qups_convd(const qups_convd_params *p, void *z, const void *x, const void *y, qups_stream_t stream);

/* ---- pwznxcorr -------------------------------------------------------------------------- */
/* Pair-wise windowed zero-normalised cross-correlation: replaces the whole-array expressions of kern/pwznxcorr.m:142-266
 * (native branch `iflt = false`; integer lags; U = 1; multi = false).  x is T x N x F (time, channels, frames; real or
 * interleaved complex), y is T x N' x F x L complex (N' = N - stride for ref NEIGHBOR, else N):
 *   y(t,n,f,l) = kernfun(xlz .* xrz_l) ./ (sqrt(kernfun(|xlz|^2)) .* sqrt(kernfun(|xrz_l|^2)))          (norm)
 *   xlz = xl - kernfun(xl),  xrz_l = conj(circshift(xr, -lag_l)) - kernfun(...)                           (zero)
 *   kernfun(z) = convn(z, w, 'same') along time; with `pad` ceil(max|lag|) zeros are appended before the circular shift
 * ref: NEIGHBOR channel n vs n + stride | CENTER mean of the median channel(s) | X0 a given signal x0 (T x {1|N} x {1|F},
 * complex).  lags: HOST int32[L]; w: DEVICE real[W] window weights (the reference's scalar W is ones(W), unscaled). */
typedef enum { QUPS_XC_NEIGHBOR = 0, QUPS_XC_CENTER = 1, QUPS_XC_X0 = 2 } qups_xcorr_ref;
typedef struct {
    uint32_t struct_size;
    int32_t dtype;      /* QUPS_F32 | QUPS_F64 */
    int32_t is_complex; /* x is interleaved complex (y and x0 always are) */
    int32_t ref;        /* qups_xcorr_ref */
    int32_t zero, norm, pad;
    uint32_t stride;
    uint32_t L, W;
    uint64_t T, N, F;
    uint64_t x0N, x0F;  /* extents of x0 along channels / frames (1 = broadcast); ignored unless ref == X0 */
} qups_xcorr_params;
QUPS_API int qups_pwznxcorr(const qups_xcorr_params *p, void *y, const void *x, const void *x0, const void *w, const int32_t *lags,
                            qups_stream_t stream);

/* ---- refocus (REFoCUS transmit decoding) -------------------------------------------------------- */
/* Applies a per-frequency decoding matrix to channel data: replaces src/UltrasoundSystem.m:3729-3757
 *   x = fft(x, T, tdim) .* exp(-2i*pi*f.*t0);  y(:,:,e) = sum_v Hi(e,v,:) .* x(:,:,v);  y = ifft(y .* exp(+2i*pi*f.*min(t0)))
 * with f = (0:T-1)*fs/T (src/ChannelData.m:1491).  x: T x N x V complex (time, receives, pulses), Hi: E x V x T complex
 * (elements x pulses x frequency, the reference's `Hi` output), y: T x N x E complex.  t0: HOST array of n_t0 (1 or V)
 * start times; *t0_out (HOST, may be NULL) receives min(t0), the start time of the decoded data (:3762).
 * The decoder itself (tikhonov / adjoint / pinv, :3690-3727) depends only on the sequence and stays host code.
 * T must be a power of two <= 8192 (QUPS_ERR_UNSUPPORTED otherwise: zero-pad, as the reference's help recommends). */
typedef struct {
    uint32_t struct_size;
    int32_t dtype;      /* QUPS_F32 */
    uint64_t T, N, V, E;
    uint32_t n_t0, reserved_;
    double fs;
} qups_refocus_params;
QUPS_API int qups_refocus(const qups_refocus_params *p, void *y, const void *x, const void *Hi, const double *t0, double *t0_out,
                          qups_stream_t stream);

/* ---- misc --------------------------------------------------------------- */
QUPS_API const char *qups_last_error(void);
QUPS_API int qups_version(void);
/* number of kernels this library launched on this thread since the last reset (bench.py's gpu_launches) */
QUPS_API uint64_t qups_launch_count(int reset);
/* name of the DAS kernel variant the last qups_das call on this thread dispatched to */
QUPS_API const char *qups_last_das_kernel(void);
/* "ws2_tiled" (canonical look-up-table delay-and-sum on the staged kernel) or "wsinterpd2" (generic strided kernel) for the
 * last qups_wsinterpd2 / qups_wsinterpd call of the process */
QUPS_API const char *qups_last_ws2_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* QUPS_B200_H */
