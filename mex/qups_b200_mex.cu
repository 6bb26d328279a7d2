// qups_b200_mex.cu — MATLAB gateway (mexcuda) to libqups_b200.so.
//
// Build (on a machine with MATLAB + Parallel Computing Toolbox; NOT buildable in this repo's image):
//   mexcuda -output qups_b200_mex mex/qups_b200_mex.cu -Iinclude -Lqups_b200 -lqups_b200
//
// It replaces, one for one, the three `parallel.gpu.CUDAKernel.feval` call sites of the reference:
//   kern/das_spec.m:371-373     y{f} = k.feval(yg, Pi, Pr, Pv, Nv, apod, cinv, [cstride, astride], x(:,:,:,f), [fs, fmod])
//   kern/wsinterpd2.m:235       y_   = k.feval(y_, w_, x_, t1_, t2_, dsizes, iflags, strides, flagnum, imag(omega))
//   src/UltrasoundSystem.m:718  x    = k.feval(x, ps, as, pn, pv, kn, sb, iblock, [t0k,t0x,fs,fsr,cinv,R0], [E,E], flag)
//
// Usage from MATLAB (all array arguments are gpuArrays, exactly what the reference passes to feval):
//   y = qups_b200_mex('das',    C, yg, Pi, Pr, Pv4, Nv, apod, cinv, acstride, x, [fs fmod])
//   y = qups_b200_mex('ws2',    C, y0, w, x, t1, t2, dsizes, strides)
//   x = qups_b200_mex('greens', C, x0, ps, as, pn, pv, kn)
//   y = qups_b200_mex('das',    C, yg, ..., [fs fmod], rx_aux, tx_aux, lat)   closed-form apodization (C.ap_rx_kind / C.ap_tx_kind,
//                                                                               C.ap_rx_p, C.ap_tx_p, C.ap_lat_dim; empty [] for unused arrays)
//   a = qups_b200_mex('apod',   C, a0, Pi, Pr, rx_aux, tx_aux, lat)           dense image of the same generator (C.which = 0 rx | 1 tx)
//   r = qups_b200_mex('aperture', C, r0, b)                                    cohfac / dmas / pcf / slsc along one dimension
//                                                                               (C.op, C.C, C.A, C.S, C.lags (uint32), C.gamma)
//   y = qups_b200_mex('prep',   C, y0, x, t0)                                 zeropad -> hilbert -> downmix -> cast (C.B, C.A, C.hilbert, C.fmix, C.fs, C.N)
//   y = qups_b200_mex('das_cohfac', C, yg, Pi, Pr, Pv4, Nv, cinv, x, cf0)       DAS image; the coherence factor is written into the
//                                                                               gpuArray cf0 (single, size of the image) in place
//   y = qups_b200_mex('xcorr',  C, y0, x, x0, w)                               pwznxcorr: C.ref, C.zero, C.norm, C.pad, C.stride, C.lags (int32, host),
//                                                                               x is T x N x F, w the window weights (single/double gpuArray), x0 may be []
//   y = qups_b200_mex('refocus',C, y0, x, Hi)                                  REFoCUS decode: C.fs, C.t0 (double, host, 1 or V values); y0 T x N x E prototype
// where C is a scalar struct holding what the reference puts in __constant__ memory with k.setConstantMemory
// (kern/das_spec.m:294-298): C.I1,C.I2,C.I3,C.N,C.M,C.T,C.S,C.VS,C.DV,C.flag  (+ optional das hints C.pitch = [dz dx], C.c0 ;
// ws2: C.T,C.interp,C.omega ;
// greens: C.n0,C.t0x,C.fs,C.fsr,C.c0,C.R0,C.E,C.interp).  The output is a new gpuArray of the size/type of the
// first array argument (feval's convention for non-const pointer parameters).
#include "mex.h"
#include "gpu/mxGPUArray.h"
#include <string.h>
#include "qups_b200.h"

static double fld(const mxArray *s, const char *name, double dflt) {
    const mxArray *f = mxGetField(s, 0, name);
    return f ? mxGetScalar(f) : dflt;
}
static int dtype_of(const mxGPUArray *a) {
    switch (mxGPUGetClassID(a)) {
        case mxDOUBLE_CLASS: return QUPS_F64;
        case mxSINGLE_CLASS: return QUPS_F32;
        case mxUINT16_CLASS: return QUPS_F16; /* halfT aliases its storage as uint16 (kern/das_spec.m:356-357) */
        default: mexErrMsgIdAndTxt("QUPS:b200:type", "Unsupported data class."); return -1;
    }
}
static void fill_apf(qups_apod_fused *f, const mxArray *C, const mxGPUArray *rx, const mxGPUArray *tx, const mxGPUArray *lat) {
    memset(f, 0, sizeof(*f));
    f->struct_size = sizeof(*f);
    f->rx_kind = (int32_t)fld(C, "ap_rx_kind", 0); f->tx_kind = (int32_t)fld(C, "ap_tx_kind", 0);
    f->lat_dim = (int32_t)fld(C, "ap_lat_dim", 2);
    const mxArray *rp = mxGetField(C, 0, "ap_rx_p"), *tp = mxGetField(C, 0, "ap_tx_p");
    for (size_t k = 0; rp && k < mxGetNumberOfElements(rp) && k < 4; ++k) f->rx_p[k] = (float)mxGetPr(rp)[k];
    for (size_t k = 0; tp && k < mxGetNumberOfElements(tp) && k < 4; ++k) f->tx_p[k] = (float)mxGetPr(tp)[k];
    f->rx_aux = rx && mxGPUGetNumberOfElements(rx) ? mxGPUGetDataReadOnly(rx) : NULL;   /* single gpuArrays */
    f->tx_aux = tx && mxGPUGetNumberOfElements(tx) ? mxGPUGetDataReadOnly(tx) : NULL;
    f->lat = lat && mxGPUGetNumberOfElements(lat) ? mxGPUGetDataReadOnly(lat) : NULL;
}
static void check(int rc) {
    if (rc != 0) mexErrMsgIdAndTxt("QUPS:b200:error", "libqups_b200 error %d: %s", rc, qups_last_error());
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    mxInitGPU();
    if (nrhs < 3 || !mxIsChar(prhs[0]) || !mxIsStruct(prhs[1]))
        mexErrMsgIdAndTxt("QUPS:b200:usage", "qups_b200_mex(op, C, out_prototype, ...)");
    char op[16];
    mxGetString(prhs[0], op, sizeof(op));
    const mxArray *C = prhs[1];
    /* the output buffer: a copy of the prototype, as feval returns non-const pointer arguments */
    const mxGPUArray *proto = mxGPUCreateFromMxArray(prhs[2]);
    mxGPUArray *out = mxGPUCopyGPUArray(proto);
    mxGPUDestroyGPUArray(proto);
    void *y = mxGPUGetData(out);
#define IN(i) mxGPUCreateFromMxArray(prhs[i])
#define RO(a) mxGPUGetDataReadOnly(a)

    if (!strcmp(op, "das")) {
        if (nrhs != 12 && nrhs != 15) mexErrMsgIdAndTxt("QUPS:b200:usage", "'das' takes 12 arguments (15 with closed-form apodization)");
        const mxGPUArray *Pi = IN(3), *Pr = IN(4), *Pv = IN(5), *Nv = IN(6), *ap = IN(7), *ci = IN(8), *x = IN(10);
        qups_das_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        p.dtype = dtype_of(x);
        p.I1 = (uint64_t)fld(C, "I1", 1); p.I2 = (uint64_t)fld(C, "I2", 1); p.I3 = (uint64_t)fld(C, "I3", 1);
        p.N = (uint64_t)fld(C, "N", 0); p.M = (uint64_t)fld(C, "M", 0); p.T = (uint64_t)fld(C, "T", 0);
        p.S = (uint64_t)fld(C, "S", 0); p.F = 1;
        p.flag = (int32_t)fld(C, "flag", 1);
        p.vs = (int32_t)fld(C, "VS", 1); p.dv = (int32_t)fld(C, "DV", 0);
        const double *fsfc = mxGetPr(prhs[11]);           /* [fs, fmod] is a host array (kern/das_spec.m:372) */
        p.fs = fsfc[0]; p.fmod = fsfc[1];
        /* optional launcher hints (never affect results): C.pitch = [scan.dz scan.dx] (metres between neighbouring pixels along
         * I1 / I2), C.c0 = scalar sound speed.  With them the library picks its tile shape without reading pixel positions
         * back from the device (no host synchronisation, no pointer-keyed cache). */
        const mxArray *pitch = mxGetField(C, 0, "pitch");
        if (pitch && mxGetNumberOfElements(pitch) >= 2) { p.pitch_hint[0] = mxGetPr(pitch)[0]; p.pitch_hint[1] = mxGetPr(pitch)[1]; }
        p.c_hint = fld(C, "c0", 0);
        /* [cstride, astride] arrives as a host uint64 array (MATLAB copies small arrays for feval) */
        if (!mxIsUint64(prhs[9])) mexErrMsgIdAndTxt("QUPS:b200:type", "strides must be uint64");
        const uint64_t *acs = (const uint64_t *)mxGetData(prhs[9]);
        if (nrhs == 15) { /* closed-form apodization evaluated in-kernel (src/UltrasoundSystem.m:4892-5429 generators) */
            const mxGPUArray *rxa = IN(12), *txa = IN(13), *lat = IN(14);
            qups_apod_fused f;
            fill_apf(&f, C, rxa, txa, lat);
            check(qups_das_fused(&p, &f, y, RO(Pi), RO(Pr), RO(Pv), RO(Nv), p.S ? RO(ap) : NULL, RO(ci), acs, RO(x), NULL));
            mxGPUDestroyGPUArray(rxa); mxGPUDestroyGPUArray(txa); mxGPUDestroyGPUArray(lat);
        } else
        check(qups_das(&p, y, RO(Pi), RO(Pr), RO(Pv), RO(Nv), p.S ? RO(ap) : NULL, RO(ci), acs, RO(x), NULL));
        mxGPUDestroyGPUArray(Pi); mxGPUDestroyGPUArray(Pr); mxGPUDestroyGPUArray(Pv); mxGPUDestroyGPUArray(Nv);
        mxGPUDestroyGPUArray(ap); mxGPUDestroyGPUArray(ci); mxGPUDestroyGPUArray(x);
    } else if (!strcmp(op, "ws2")) {
        if (nrhs != 9) mexErrMsgIdAndTxt("QUPS:b200:usage", "'ws2' takes 9 arguments");
        const mxGPUArray *w = IN(3), *x = IN(4), *t1 = IN(5), *t2 = IN(6);
        qups_ws2_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        p.dtype = dtype_of(x);
        p.T = (uint64_t)fld(C, "T", 0);
        p.interp = (int32_t)fld(C, "interp", 1);
        p.omega = fld(C, "omega", 0);
        const size_t D = mxGetNumberOfElements(prhs[7]);
        if (D > 8) mexErrMsgIdAndTxt("QUPS:b200:usage", "at most 8 dimensions");
        p.D = (uint32_t)D;
        const uint64_t *ds = (const uint64_t *)mxGetData(prhs[7]);   /* dsizes (uint64)                 */
        const uint64_t *st = (const uint64_t *)mxGetData(prhs[8]);   /* strides: 5 x D = [w;y;t1;t2;x]  */
        for (size_t d = 0; d < D; ++d) {
            p.sizes[d] = ds[d];
            for (int r = 0; r < 5; ++r) p.dstride[r + 5 * d] = st[r + 5 * d];
        }
        check(qups_wsinterpd2(&p, y, RO(w), RO(x), RO(t1), RO(t2), NULL));
        mxGPUDestroyGPUArray(w); mxGPUDestroyGPUArray(x); mxGPUDestroyGPUArray(t1); mxGPUDestroyGPUArray(t2);
    } else if (!strcmp(op, "greens")) {
        if (nrhs != 8) mexErrMsgIdAndTxt("QUPS:b200:usage", "'greens' takes 8 arguments");
        const mxGPUArray *ps = IN(3), *as = IN(4), *pn = IN(5), *pv = IN(6), *kn = IN(7);
        qups_greens_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        p.dtype = dtype_of(kn);
        const mwSize *osz = mxGPUGetDimensions(out);                  /* x is 1 x T x N x M (:619) */
        p.S = osz[1]; p.N = osz[2]; p.M = mxGPUGetNumberOfDimensions(out) > 3 ? osz[3] : 1;
        p.I = mxGPUGetNumberOfElements(as);
        p.T = mxGPUGetNumberOfElements(kn);
        p.E = (uint64_t)fld(C, "E", 1);
        p.n0 = (int64_t)fld(C, "n0", 0);
        p.interp = (int32_t)fld(C, "interp", 2);
        p.t0x = fld(C, "t0x", 0); p.fs = fld(C, "fs", 1); p.fsr = fld(C, "fsr", 1);
        p.c0 = fld(C, "c0", 1540); p.R0 = fld(C, "R0", 0);
        check(qups_greens(&p, y, RO(ps), RO(as), RO(pn), RO(pv), RO(kn), NULL));
        mxGPUDestroyGPUArray(ps); mxGPUDestroyGPUArray(as); mxGPUDestroyGPUArray(pn); mxGPUDestroyGPUArray(pv);
        mxGPUDestroyGPUArray(kn);
    } else if (!strcmp(op, "apod")) {
        if (nrhs != 8) mexErrMsgIdAndTxt("QUPS:b200:usage", "'apod' takes 8 arguments");
        const mxGPUArray *Pi = IN(3), *Pr = IN(4), *rxa = IN(5), *txa = IN(6), *lat = IN(7);
        qups_apod_fused f;
        fill_apf(&f, C, rxa, txa, lat);
        const int which = (int)fld(C, "which", 0);
        check(qups_apod_generate(&f, which, y, mxGPUGetComplexity(out) == mxCOMPLEX, RO(Pi), RO(Pr), (uint64_t)fld(C, "I1", 1),
                                 (uint64_t)fld(C, "I2", 1), (uint64_t)fld(C, "I3", 1), (uint64_t)fld(C, which ? "M" : "N", 1), NULL));
        mxGPUDestroyGPUArray(Pi); mxGPUDestroyGPUArray(Pr); mxGPUDestroyGPUArray(rxa); mxGPUDestroyGPUArray(txa); mxGPUDestroyGPUArray(lat);
    } else if (!strcmp(op, "aperture")) {
        if (nrhs != 4) mexErrMsgIdAndTxt("QUPS:b200:usage", "'aperture' takes 4 arguments");
        const mxGPUArray *b = IN(3);
        qups_aperture_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        p.dtype = dtype_of(b);
        p.op = (int32_t)fld(C, "op", 0);
        p.C = (uint64_t)fld(C, "C", 1); p.A = (uint64_t)fld(C, "A", 1); p.S = (uint64_t)fld(C, "S", 1);
        p.gamma = fld(C, "gamma", 1);
        const mxArray *lg = mxGetField(C, 0, "lags");          /* host uint32 vector */
        p.nlags = lg ? (uint32_t)mxGetNumberOfElements(lg) : 0;
        check(qups_aperture(&p, y, NULL, RO(b), lg ? (const uint32_t *)mxGetData(lg) : NULL, NULL));
        mxGPUDestroyGPUArray(b);
    } else if (!strcmp(op, "prep")) {
        if (nrhs != 5) mexErrMsgIdAndTxt("QUPS:b200:usage", "'prep' takes 5 arguments");
        const mxGPUArray *x = IN(3), *t0 = IN(4);            /* y0: complex single (B+T+A) x N x M prototype; t0: single */
        qups_prep_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        const bool cplx = mxGPUGetComplexity(x) == mxCOMPLEX;
        switch (mxGPUGetClassID(x)) {
            case mxSINGLE_CLASS: p.in_dtype = cplx ? QUPS_IN_CPLX_F32 : QUPS_IN_REAL_F32; break;
            case mxINT16_CLASS: p.in_dtype = QUPS_IN_REAL_I16; break;
            case mxDOUBLE_CLASS: p.in_dtype = QUPS_IN_REAL_F64; break;
            default: mexErrMsgIdAndTxt("QUPS:b200:type", "Unsupported data class.");
        }
        p.out_dtype = mxGPUGetClassID(out) == mxUINT16_CLASS ? QUPS_F16 : QUPS_F32;
        const mwSize *xs = mxGPUGetDimensions(x);
        p.T = xs[0]; p.K = mxGPUGetNumberOfElements(x) / (xs[0] ? xs[0] : 1);
        p.B = (uint64_t)fld(C, "B", 0); p.A = (uint64_t)fld(C, "A", 0); p.hilbert = (int32_t)fld(C, "hilbert", 0);
        p.traces_per_t0 = (uint64_t)fld(C, "N", 1); p.n_t0 = mxGPUGetNumberOfElements(t0);
        p.fs = fld(C, "fs", 1); p.fmix = fld(C, "fmix", 0);
        check(qups_chd_prep(&p, y, RO(x), p.n_t0 ? RO(t0) : NULL, NULL));
        mxGPUDestroyGPUArray(x); mxGPUDestroyGPUArray(t0);
    } else if (!strcmp(op, "das_cohfac")) {
        if (nrhs != 10) mexErrMsgIdAndTxt("QUPS:b200:usage", "'das_cohfac' takes 10 arguments");
        const mxGPUArray *Pi = IN(3), *Pr = IN(4), *Pv = IN(5), *Nv = IN(6), *cinv = IN(7), *x = IN(8), *cf = IN(9);
        qups_das_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        p.dtype = dtype_of(x);
        p.I1 = (uint64_t)fld(C, "I1", 1); p.I2 = (uint64_t)fld(C, "I2", 1); p.I3 = (uint64_t)fld(C, "I3", 1);
        p.N = (uint64_t)fld(C, "N", 1); p.M = (uint64_t)fld(C, "M", 1); p.T = (uint64_t)fld(C, "T", 1); p.F = 1;
        p.flag = (int32_t)fld(C, "flag", 0); p.vs = (int32_t)fld(C, "VS", 1); p.dv = (int32_t)fld(C, "DV", 0);
        p.fs = fld(C, "fs", 1);
        /* cf0 is written in place: the factor comes back through the caller's gpuArray (a second output would need a second
         * prototype argument; the reference's feval convention returns one array per non-const pointer) */
        check(qups_das_cohfac(&p, y, (void *)mxGPUGetDataReadOnly(cf), RO(Pi), RO(Pr), RO(Pv), RO(Nv), RO(cinv), RO(x), NULL));
        mxGPUDestroyGPUArray(Pi); mxGPUDestroyGPUArray(Pr); mxGPUDestroyGPUArray(Pv); mxGPUDestroyGPUArray(Nv);
        mxGPUDestroyGPUArray(cinv); mxGPUDestroyGPUArray(x); mxGPUDestroyGPUArray(cf);
    } else if (!strcmp(op, "xcorr")) {
        if (nrhs != 6) mexErrMsgIdAndTxt("QUPS:b200:usage", "'xcorr' takes 6 arguments");
        const mxGPUArray *x = IN(3), *x0 = IN(4), *w = IN(5);
        qups_xcorr_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        p.dtype = dtype_of(x);
        p.is_complex = mxGPUGetComplexity(x) == mxCOMPLEX;
        p.ref = (int32_t)fld(C, "ref", 0); p.zero = (int32_t)fld(C, "zero", 1); p.norm = (int32_t)fld(C, "norm", 1);
        p.pad = (int32_t)fld(C, "pad", 1); p.stride = (uint32_t)fld(C, "stride", 1);
        const mxArray *lg = mxGetField(C, 0, "lags");          /* host int32 vector */
        if (!lg || mxGetClassID(lg) != mxINT32_CLASS) mexErrMsgIdAndTxt("QUPS:b200:usage", "C.lags must be an int32 vector");
        p.L = (uint32_t)mxGetNumberOfElements(lg); p.W = (uint32_t)mxGPUGetNumberOfElements(w);
        const mwSize *xs = mxGPUGetDimensions(x);
        const mwSize nd = mxGPUGetNumberOfDimensions(x);
        p.T = xs[0]; p.N = nd > 1 ? xs[1] : 1; p.F = mxGPUGetNumberOfElements(x) / (p.T * p.N ? p.T * p.N : 1);
        if (mxGPUGetNumberOfElements(x0)) {
            const mwSize *zs = mxGPUGetDimensions(x0);
            p.x0N = mxGPUGetNumberOfDimensions(x0) > 1 ? zs[1] : 1;
            p.x0F = mxGPUGetNumberOfElements(x0) / (zs[0] * p.x0N ? zs[0] * p.x0N : 1);
        }
        check(qups_pwznxcorr(&p, y, RO(x), mxGPUGetNumberOfElements(x0) ? RO(x0) : NULL, RO(w), (const int32_t *)mxGetData(lg), NULL));
        mxGPUDestroyGPUArray(x); mxGPUDestroyGPUArray(x0); mxGPUDestroyGPUArray(w);
    } else if (!strcmp(op, "refocus")) {
        if (nrhs != 5) mexErrMsgIdAndTxt("QUPS:b200:usage", "'refocus' takes 5 arguments");
        const mxGPUArray *x = IN(3), *Hi = IN(4);             /* x: T x N x V, Hi: E x V x T, both complex single */
        qups_refocus_params p;
        memset(&p, 0, sizeof(p));
        p.struct_size = sizeof(p);
        p.dtype = dtype_of(x);
        const mwSize *xs = mxGPUGetDimensions(x), *hs = mxGPUGetDimensions(Hi);
        p.T = xs[0]; p.N = xs[1]; p.V = mxGPUGetNumberOfDimensions(x) > 2 ? xs[2] : 1; p.E = hs[0];
        const mxArray *t0 = mxGetField(C, 0, "t0");            /* host double vector, 1 or V start times */
        if (!t0 || !mxIsDouble(t0)) mexErrMsgIdAndTxt("QUPS:b200:usage", "C.t0 must be a double vector");
        p.n_t0 = (uint32_t)mxGetNumberOfElements(t0); p.fs = fld(C, "fs", 1);
        check(qups_refocus(&p, y, RO(x), RO(Hi), mxGetPr(t0), NULL, NULL));
        mxGPUDestroyGPUArray(x); mxGPUDestroyGPUArray(Hi);
    } else {
        mexErrMsgIdAndTxt("QUPS:b200:usage", "unknown op '%s'", op);
    }
    plhs[0] = mxGPUCreateMxArrayOnGPU(out);
    mxGPUDestroyGPUArray(out);
}
