/* Minimal DECLARATIONS of the MATLAB MEX C API (R2018a interleaved-real names that the shims in
 * mex/ use) -- test infrastructure only: lets `gcc -fsyntax-only` type-check the shim sources
 * against include/ssnamg.h in an image that has no MATLAB.  Nothing here is ever linked. */
#ifndef SSN_MEX_STUB_H
#define SSN_MEX_STUB_H
#include <stddef.h>
#include <stdbool.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef uint16_t mxChar;
typedef bool mxLogical;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...);
int mexAtExit(void (*fn)(void));
void mexLock(void);
size_t mxGetNumberOfElements(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
bool mxIsDouble(const mxArray *a);
bool mxIsSparse(const mxArray *a);
bool mxIsStruct(const mxArray *a);
bool mxIsChar(const mxArray *a);
bool mxIsEmpty(const mxArray *a);
bool mxIsLogical(const mxArray *a);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
mxLogical *mxGetLogicals(const mxArray *a);
double mxGetScalar(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);
mwIndex *mxGetJc(const mxArray *a);
mwIndex *mxGetIr(const mxArray *a);
mwSize mxGetNzmax(const mxArray *a);
mxArray *mxGetField(const mxArray *s, mwIndex i, const char *name);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity c);
mxArray *mxCreateLogicalMatrix(mwSize m, mwSize n);
void *mxMalloc(size_t n);
void *mxCalloc(size_t n, size_t sz);
void mxFree(void *p);
int mexCallMATLAB(int nlhs, mxArray *plhs[], int nrhs, mxArray *prhs[], const char *name);
void mxDestroyArray(mxArray *a);
const mxArray *mexGetVariablePtr(const char *workspace, const char *name);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#endif
