// tests/stub/mex.h -- a small functional stand-in for MATLAB's mex.h / matrix.h (R2018a interleaved API), test infrastructure
// only: it lets matlab/redmax_mex.cpp be compiled by g++ and driven from tests/stub/mex_driver.cpp in an image without
// MATLAB.  Only what the gateway uses is provided.  mexErrMsgIdAndTxt throws (MATLAB longjmps out of the MEX function).
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

typedef size_t mwSize;
typedef enum { mxUNKNOWN_CLASS = 0, mxSTRUCT_CLASS, mxCHAR_CLASS, mxDOUBLE_CLASS, mxINT32_CLASS, mxUINT64_CLASS } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;
typedef double mxDouble;
typedef int32_t mxInt32;

struct mxArray {
    mxClassID cls = mxUNKNOWN_CLASS;
    std::vector<mwSize> dims;
    std::vector<unsigned char> data;
    std::string str;
    std::map<std::string, mxArray*> fields;
    size_t numel() const {
        size_t n = 1;
        for (mwSize d : dims) n *= d;
        return dims.empty() ? 0 : n;
    }
};

struct mex_error : std::runtime_error {
    std::string id;
    mex_error(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {}
};

inline size_t mx_elsize(mxClassID c) { return c == mxDOUBLE_CLASS ? 8 : c == mxINT32_CLASS ? 4 : c == mxUINT64_CLASS ? 8 : 1; }
inline mxArray* mxCreateNumericArray(mwSize nd, const mwSize* dims, mxClassID c, mxComplexity) {
    mxArray* a = new mxArray();
    a->cls = c;
    a->dims.assign(dims, dims + nd);
    a->data.assign(a->numel() * mx_elsize(c), 0);
    return a;
}
inline mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID c, mxComplexity x) {
    const mwSize d[2] = {m, n};
    return mxCreateNumericArray(2, d, c, x);
}
inline mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity x) { return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, x); }
inline mxArray* mxDuplicateArray(const mxArray* a) { return new mxArray(*a); }
inline void mxDestroyArray(mxArray* a) { delete a; }
inline size_t mxGetNumberOfElements(const mxArray* a) { return a->cls == mxCHAR_CLASS ? a->str.size() : a->numel(); }
inline size_t mxGetM(const mxArray* a) { return a->dims.empty() ? 0 : a->dims[0]; }
inline size_t mxGetN(const mxArray* a) {
    size_t n = 1;
    for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i];
    return a->dims.empty() ? 0 : n;
}
inline bool mxIsEmpty(const mxArray* a) { return mxGetNumberOfElements(a) == 0; }
inline bool mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
inline bool mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
inline bool mxIsInt32(const mxArray* a) { return a->cls == mxINT32_CLASS; }
inline bool mxIsUint64(const mxArray* a) { return a->cls == mxUINT64_CLASS; }
inline bool mxIsStruct(const mxArray* a) { return a->cls == mxSTRUCT_CLASS; }
inline bool mxIsComplex(const mxArray*) { return false; }
inline void* mxGetData(const mxArray* a) { return (void*)a->data.data(); }
// the typed accessors of the interleaved API are only valid on arrays of their own class: MATLAB terminates the MEX file
// otherwise -- here that is a hard failure of the test
inline mxDouble* mxGetDoubles(const mxArray* a) {
    if (a->cls != mxDOUBLE_CLASS) throw std::logic_error("mxGetDoubles on a non-double array (MATLAB would terminate the MEX file)");
    return (mxDouble*)a->data.data();
}
inline mxInt32* mxGetInt32s(const mxArray* a) {
    if (a->cls != mxINT32_CLASS) throw std::logic_error("mxGetInt32s on a non-int32 array (MATLAB would terminate the MEX file)");
    return (mxInt32*)a->data.data();
}
inline double mxGetScalar(const mxArray* a) {
    if (a->cls == mxDOUBLE_CLASS) return *(const double*)a->data.data();
    if (a->cls == mxINT32_CLASS) return (double)*(const int32_t*)a->data.data();
    if (a->cls == mxUINT64_CLASS) return (double)*(const uint64_t*)a->data.data();
    return 0.0;
}
inline mxArray* mxGetField(const mxArray* s, size_t, const char* name) {
    auto it = s->fields.find(name);
    return it == s->fields.end() ? nullptr : it->second;
}
inline int mxGetString(const mxArray* a, char* buf, mwSize len) {
    std::strncpy(buf, a->str.c_str(), len - 1);
    buf[len - 1] = 0;
    return 0;
}
inline void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char msg[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof(msg), fmt, ap);
    va_end(ap);
    throw mex_error(id, msg);
}
inline void mexLock() {}
inline int mexAtExit(void (*)(void)) { return 0; }

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
