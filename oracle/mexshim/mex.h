/*
 * mex.h -- minimal stand-in for MATLAB's MEX API, just enough for the reference's
 * ucodes/omc_matrad/omc_matrad.c to COMPILE without MATLAB (there is none in this image).
 * TEST INFRASTRUCTURE (oracle/).  Only the declarations the file uses; every function is a stub that
 * aborts if reached: the harness (oracle/ref_harness_matrad.c) never calls mexFunction()/parseInput(),
 * it fills the reference's global structs itself and calls initHistory(ibeamlet)/shower() directly.
 */
#ifndef OMC_MEX_SHIM_H
#define OMC_MEX_SHIM_H
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef unsigned char mxLogical;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

#define OMC_MEX_STUB(ret, name, args) static ret name args { fprintf(stderr, "mex shim: " #name " called\n"); abort(); }
OMC_MEX_STUB(mxArray *, mxGetField, (const mxArray *a, mwIndex i, const char *f))
OMC_MEX_STUB(double *, mxGetPr, (const mxArray *a))
OMC_MEX_STUB(char *, mxArrayToString, (const mxArray *a))
OMC_MEX_STUB(void *, mxRealloc, (void *p, size_t n))
OMC_MEX_STUB(mwIndex *, mxGetIr, (const mxArray *a))
OMC_MEX_STUB(mwIndex *, mxGetJc, (const mxArray *a))
OMC_MEX_STUB(void, mxDestroyArray, (mxArray *a))
OMC_MEX_STUB(void, mxSetPr, (mxArray *a, double *p))
OMC_MEX_STUB(void, mxSetIr, (mxArray *a, mwIndex *p))
OMC_MEX_STUB(void, mxSetNzmax, (mxArray *a, mwSize n))
OMC_MEX_STUB(int, mxIsStruct, (const mxArray *a))
OMC_MEX_STUB(int, mxIsDouble, (const mxArray *a))
OMC_MEX_STUB(int, mxIsInt32, (const mxArray *a))
OMC_MEX_STUB(int, mxIsInt64, (const mxArray *a))
OMC_MEX_STUB(int, mxIsInt8, (const mxArray *a))
OMC_MEX_STUB(int, mxIsInt16, (const mxArray *a))
OMC_MEX_STUB(const mwSize *, mxGetDimensions, (const mxArray *a))
OMC_MEX_STUB(double, mxGetScalar, (const mxArray *a))
OMC_MEX_STUB(int, mxGetNumberOfFields, (const mxArray *a))
OMC_MEX_STUB(mwSize, mxGetNumberOfDimensions, (const mxArray *a))
OMC_MEX_STUB(mxLogical *, mxGetLogicals, (const mxArray *a))
OMC_MEX_STUB(mxArray *, mxGetCell, (const mxArray *a, mwIndex i))
OMC_MEX_STUB(mxArray *, mxCreateString, (const char *s))
OMC_MEX_STUB(mxArray *, mxCreateSparse, (mwSize m, mwSize n, mwSize nz, mxComplexity c))
OMC_MEX_STUB(mxArray *, mxCreateDoubleScalar, (double v))
OMC_MEX_STUB(mwIndex, mxCalcSingleSubscript, (const mxArray *a, mwSize n, const mwIndex *s))
OMC_MEX_STUB(int, mexCallMATLAB, (int nl, mxArray *pl[], int nr, mxArray *pr[], const char *f))
#define mexPrintf printf
static void mexErrMsgIdAndTxt(const char *id, const char *msg, ...) { fprintf(stderr, "%s: %s\n", id, msg); abort(); }
#endif
