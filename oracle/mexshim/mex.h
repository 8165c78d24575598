/*
 * mex.h -- small stand-in for MATLAB's MEX API (there is no MATLAB in this image), enough for the reference's
 * ucodes/omc_matrad/omc_matrad.c to compile AND for its parseInput() / initPhantom() / initSource() / mexFunction() to run on
 * arrays built in C.  TEST INFRASTRUCTURE (oracle/): used by oracle/ref_harness.c (-DOMC_REF_MATRAD), by
 * oracle/matrad_mex_harness.c and, as the header the drop-in is compiled against here, by ompmc_b200/host/omc_matrad_dropin.c
 * (a maintainer compiles that file with MATLAB's own mex.h instead).
 *
 * Only what the reference uses: real double / int32 / logical / char arrays, 1x1 structs, cell arrays, real sparse matrices;
 * mexCallMATLAB() knows "num2str" (integers as %d, other values as %.5g-like text, row vectors blank-separated -- what the
 * reference then feeds to atoi / atof / sscanf), "waitbar" and "close" (no-ops).  Column-major storage as in MATLAB.
 * mx_new_*() / mx_set_field() are this shim's own constructors for building inputs without MATLAB.
 */
#ifndef OMC_MEX_SHIM_H
#define OMC_MEX_SHIM_H
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef size_t mwSize;
typedef size_t mwIndex;
typedef unsigned char mxLogical;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { MX_DOUBLE, MX_INT32, MX_LOGICAL, MX_CHAR, MX_STRUCT, MX_CELL, MX_SPARSE } mx_shim_class;

typedef struct mxArray_tag {
    mx_shim_class cls;
    mwSize ndim, dims[4];
    void *data;                    /* double* / int* / mxLogical* / char* (NUL-terminated); sparse: pr */
    int owns;                      /* data allocated by the shim (mxDestroyArray frees it) */
    int nfields;                   /* struct */
    char **names;
    struct mxArray_tag **fields;   /* struct fields or cell elements */
    mwIndex *ir, *jc;              /* sparse */
    mwSize nzmax;
} mxArray;

#define MX_UNUSED __attribute__((unused))

static MX_UNUSED mwSize mx_numel(const mxArray *a) {
    mwSize n = 1;
    for (mwSize i = 0; i < a->ndim; i++) n *= a->dims[i];
    return n;
}
static MX_UNUSED mxArray *mx_new(mx_shim_class cls, mwSize ndim, const mwSize *dims, size_t elem) {
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->cls = cls; a->ndim = ndim < 2 ? 2 : ndim;
    for (mwSize i = 0; i < 4; i++) a->dims[i] = 1;
    for (mwSize i = 0; i < ndim && i < 4; i++) a->dims[i] = dims[i];
    if (elem) { a->data = calloc(mx_numel(a) ? mx_numel(a) : 1, elem); a->owns = 1; }
    return a;
}
/* constructors of this shim (not MATLAB API) */
static MX_UNUSED mxArray *mx_new_double(mwSize ndim, const mwSize *dims) { return mx_new(MX_DOUBLE, ndim, dims, sizeof(double)); }
static MX_UNUSED mxArray *mx_new_int32(mwSize ndim, const mwSize *dims) { return mx_new(MX_INT32, ndim, dims, sizeof(int)); }
static MX_UNUSED mxArray *mx_new_double_row(mwSize n, const double *v) {
    const mwSize d[2] = {1, n};
    mxArray *a = mx_new_double(2, d);
    if (v) memcpy(a->data, v, n * sizeof(double));
    return a;
}
static MX_UNUSED mxArray *mx_new_logical_scalar(int v) {
    const mwSize d[2] = {1, 1};
    mxArray *a = mx_new(MX_LOGICAL, 2, d, sizeof(mxLogical));
    ((mxLogical *)a->data)[0] = (mxLogical)(v != 0);
    return a;
}
static MX_UNUSED mxArray *mx_new_struct(void) {
    const mwSize d[2] = {1, 1};
    return mx_new(MX_STRUCT, 2, d, 0);
}
static MX_UNUSED void mx_set_field(mxArray *s, const char *name, mxArray *v) {
    for (int i = 0; i < s->nfields; i++)
        if (!strcmp(s->names[i], name)) { s->fields[i] = v; return; }
    s->names = (char **)realloc(s->names, (size_t)(s->nfields + 1) * sizeof(char *));
    s->fields = (mxArray **)realloc(s->fields, (size_t)(s->nfields + 1) * sizeof(mxArray *));
    s->names[s->nfields] = strdup(name);
    s->fields[s->nfields] = v;
    s->nfields += 1;
}
static MX_UNUSED mxArray *mx_new_cell(mwSize rows, mwSize cols) {
    const mwSize d[2] = {rows, cols};
    mxArray *a = mx_new(MX_CELL, 2, d, 0);
    a->fields = (mxArray **)calloc(rows * cols ? rows * cols : 1, sizeof(mxArray *));
    return a;
}

/* ---- the MATLAB API the reference uses ------------------------------------------------------------------------------------- */
static MX_UNUSED mxArray *mxCreateString(const char *s) {
    const mwSize d[2] = {1, strlen(s)};
    mxArray *a = mx_new(MX_CHAR, 2, d, 0);
    a->data = strdup(s); a->owns = 1;
    return a;
}
static MX_UNUSED mxArray *mxCreateDoubleScalar(double v) { return mx_new_double_row(1, &v); }
static MX_UNUSED mxArray *mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity c) {
    const mwSize d[2] = {m, n};
    (void)c;
    mxArray *a = mx_new(MX_SPARSE, 2, d, 0);
    if (nzmax < 1) nzmax = 1;
    a->nzmax = nzmax;
    a->data = calloc(nzmax, sizeof(double)); a->owns = 1;
    a->ir = (mwIndex *)calloc(nzmax, sizeof(mwIndex));
    a->jc = (mwIndex *)calloc(n + 1, sizeof(mwIndex));
    return a;
}
static MX_UNUSED mxArray *mxGetField(const mxArray *a, mwIndex i, const char *f) {
    if (!a || a->cls != MX_STRUCT || i != 0) return NULL;
    for (int k = 0; k < a->nfields; k++)
        if (!strcmp(a->names[k], f)) return a->fields[k];
    return NULL;
}
static MX_UNUSED double *mxGetPr(const mxArray *a) { return a ? (double *)a->data : NULL; }
static MX_UNUSED char *mxArrayToString(const mxArray *a) { return (a && a->cls == MX_CHAR) ? strdup((const char *)a->data) : NULL; }
static MX_UNUSED void *mxRealloc(void *p, size_t n) { return realloc(p, n ? n : 1); }
static MX_UNUSED mwIndex *mxGetIr(const mxArray *a) { return a->ir; }
static MX_UNUSED mwIndex *mxGetJc(const mxArray *a) { return a->jc; }
static MX_UNUSED void mxSetPr(mxArray *a, double *p) { a->data = p; }
static MX_UNUSED void mxSetIr(mxArray *a, mwIndex *p) { a->ir = p; }
static MX_UNUSED void mxSetNzmax(mxArray *a, mwSize n) { a->nzmax = n; }
static MX_UNUSED void mxDestroyArray(mxArray *a) {
    if (!a) return;
    if (a->owns) free(a->data);
    free(a->ir); free(a->jc);
    for (int k = 0; k < a->nfields; k++) free(a->names[k]);
    free(a->names); free(a->fields);
    free(a);
}
static MX_UNUSED int mxIsStruct(const mxArray *a) { return a && a->cls == MX_STRUCT; }
static MX_UNUSED int mxIsDouble(const mxArray *a) { return a && a->cls == MX_DOUBLE; }
static MX_UNUSED int mxIsInt32(const mxArray *a) { return a && a->cls == MX_INT32; }
static MX_UNUSED int mxIsInt64(const mxArray *a) { (void)a; return 0; }
static MX_UNUSED int mxIsInt8(const mxArray *a) { (void)a; return 0; }
static MX_UNUSED int mxIsInt16(const mxArray *a) { (void)a; return 0; }
static MX_UNUSED const mwSize *mxGetDimensions(const mxArray *a) { return a->dims; }
static MX_UNUSED mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->ndim; }
static MX_UNUSED int mxGetNumberOfFields(const mxArray *a) { return a->nfields; }
static MX_UNUSED mxLogical *mxGetLogicals(const mxArray *a) { return (a && a->cls == MX_LOGICAL) ? (mxLogical *)a->data : NULL; }
static MX_UNUSED mxArray *mxGetCell(const mxArray *a, mwIndex i) { return (a && a->cls == MX_CELL && i < mx_numel(a)) ? a->fields[i] : NULL; }
static MX_UNUSED double mxGetScalar(const mxArray *a) {
    if (!a || !a->data) return 0.0;
    if (a->cls == MX_INT32) return (double)((int *)a->data)[0];
    if (a->cls == MX_LOGICAL) return (double)((mxLogical *)a->data)[0];
    return ((double *)a->data)[0];
}
static MX_UNUSED mwIndex mxCalcSingleSubscript(const mxArray *a, mwSize n, const mwIndex *s) {
    mwIndex lin = 0, mul = 1;
    for (mwSize i = 0; i < n; i++) { lin += s[i] * mul; mul *= (i < 4 ? a->dims[i] : 1); }
    return lin;
}
static MX_UNUSED int mexCallMATLAB(int nl, mxArray *pl[], int nr, mxArray *pr[], const char *f) {
    if (!strcmp(f, "num2str") && nr == 1 && nl == 1 && pr[0]) {
        char buf[512];
        size_t at = 0;
        const mxArray *a = pr[0];
        const mwSize n = mx_numel(a);
        buf[0] = 0;
        for (mwSize i = 0; i < n && at < sizeof buf - 40; i++) {
            const double v = a->cls == MX_INT32 ? (double)((int *)a->data)[i] : (a->cls == MX_LOGICAL ? (double)((mxLogical *)a->data)[i] : ((double *)a->data)[i]);
            if (i) at += (size_t)snprintf(buf + at, sizeof buf - at, "  ");
            if (v == floor(v) && fabs(v) < 1e15) at += (size_t)snprintf(buf + at, sizeof buf - at, "%.0f", v);
            else at += (size_t)snprintf(buf + at, sizeof buf - at, "%.5g", v);
        }
        pl[0] = mxCreateString(buf);
        return 0;
    }
    if (!strcmp(f, "waitbar")) { if (nl >= 1) pl[0] = mxCreateDoubleScalar(1.0); return 0; }
    if (!strcmp(f, "close")) return 0;
    fprintf(stderr, "mex shim: mexCallMATLAB(\"%s\") is not available without MATLAB\n", f);
    return 1;
}
#define mexPrintf printf
static MX_UNUSED void mexErrMsgIdAndTxt(const char *id, const char *msg, ...) {
    fprintf(stderr, "%s: %s\n", id, msg);
    fflush(stdout);
    exit(EXIT_FAILURE);            /* MATLAB unwinds to the prompt; a stand-alone process can only stop */
}
#endif
