/* Minimal declarations of the MATLAB MEX API used by fmcw_cuda_mex.cpp, for SYNTAX CHECKING ONLY in
 * environments without MATLAB (this container).  Not a substitute for MATLAB's mex.h. */
#ifndef FMCW_STUB_MEX_H
#define FMCW_STUB_MEX_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7, mxINT16_CLASS = 10, mxINT32_CLASS = 12 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
bool mxIsChar(const mxArray*); bool mxIsStruct(const mxArray*); bool mxIsDouble(const mxArray*); bool mxIsSingle(const mxArray*);
bool mxIsInt16(const mxArray*); bool mxIsEmpty(const mxArray*); bool mxIsNumeric(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxGetField(const mxArray*, mwSize, const char*);
void mxSetField(mxArray*, mwSize, const char*, mxArray*);
double mxGetScalar(const mxArray*); double* mxGetPr(const mxArray*); void* mxGetData(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
void mxSetN(mxArray*, mwSize);
void mexLock(void); void mexUnlock(void); int mexAtExit(void (*)(void));
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexFunction(int, mxArray*[], int, const mxArray*[]);
#ifdef __cplusplus
}
#endif
#endif
