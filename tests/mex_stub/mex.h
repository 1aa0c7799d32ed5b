/* Minimal stand-in for MATLAB's mex.h: the declarations matlab/manisdp_mex.cpp needs, in a container without MATLAB.
 * tests/mex_stub/mex_stub.cpp implements them functionally so that tests/test_mex_gateway.py can EXECUTE the gateway
 * (create -> set_Y -> tr_solve -> kkt -> get_Y on the GPU).  Not used by any product code. */
#ifndef MEX_STUB_H
#define MEX_STUB_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxUINT64_CLASS = 13 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
bool mxIsChar(const mxArray*);
bool mxIsSparse(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsStruct(const mxArray*);
bool mxIsUint64(const mxArray*);
bool mxIsCell(const mxArray*);
mxArray* mxCreateCellMatrix(mwSize, mwSize);
mxArray* mxGetCell(const mxArray*, mwIndex);
void mxSetCell(mxArray*, mwIndex, mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
double mxGetScalar(const mxArray*);
double* mxGetPr(const mxArray*);
void* mxGetData(const mxArray*);
mwIndex* mxGetIr(const mxArray*);
mwIndex* mxGetJc(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
void mxSetFieldByNumber(mxArray*, mwIndex, int, mxArray*);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexLock(void);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
