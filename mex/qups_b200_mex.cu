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
// where C is a scalar struct holding what the reference puts in __constant__ memory with k.setConstantMemory
// (kern/das_spec.m:294-298): C.I1,C.I2,C.I3,C.N,C.M,C.T,C.S,C.VS,C.DV,C.flag  (+ ws2: C.T,C.interp,C.omega ;
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
        if (nrhs != 12) mexErrMsgIdAndTxt("QUPS:b200:usage", "'das' takes 12 arguments");
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
        /* [cstride, astride] arrives as a host uint64 array (MATLAB copies small arrays for feval) */
        if (!mxIsUint64(prhs[9])) mexErrMsgIdAndTxt("QUPS:b200:type", "strides must be uint64");
        const uint64_t *acs = (const uint64_t *)mxGetData(prhs[9]);
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
    } else {
        mexErrMsgIdAndTxt("QUPS:b200:usage", "unknown op '%s'", op);
    }
    plhs[0] = mxGPUCreateMxArrayOnGPU(out);
    mxGPUDestroyGPUArray(out);
}
