// Shared helpers for the emerge_b200 CUDA library (sm_100a).
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#define EMB_HD __host__ __device__ __forceinline__

// complex128 as two doubles; layout-compatible with numpy complex128 / double2.
struct cx {
    double re, im;
};
EMB_HD cx mk(double r, double i = 0.0) { return cx{r, i}; }
EMB_HD cx operator+(cx a, cx b) { return cx{a.re + b.re, a.im + b.im}; }
EMB_HD cx operator-(cx a, cx b) { return cx{a.re - b.re, a.im - b.im}; }
EMB_HD cx operator-(cx a) { return cx{-a.re, -a.im}; }
EMB_HD cx operator*(cx a, cx b) { return cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
EMB_HD cx operator*(double s, cx a) { return cx{s * a.re, s * a.im}; }
EMB_HD cx operator*(cx a, double s) { return cx{s * a.re, s * a.im}; }
EMB_HD cx& operator+=(cx& a, cx b) { a.re += b.re; a.im += b.im; return a; }
EMB_HD cx& operator-=(cx& a, cx b) { a.re -= b.re; a.im -= b.im; return a; }
EMB_HD cx conj(cx a) { return cx{a.re, -a.im}; }
EMB_HD double norm2(cx a) { return a.re * a.re + a.im * a.im; }
// a += s*b with real s (2 DFMA)
EMB_HD void fma_r(cx& a, double s, cx b) { a.re += s * b.re; a.im += s * b.im; }
// a += b*c (4 DFMA)
EMB_HD void fma_c(cx& a, cx b, cx c) {
    a.re += b.re * c.re - b.im * c.im;
    a.im += b.re * c.im + b.im * c.re;
}
EMB_HD cx cdiv(cx a, cx b) {
    double d = 1.0 / (b.re * b.re + b.im * b.im);
    return cx{(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}

// status codes of the C-ABI (include/emerge_b200.h)
enum {
    EMB_OK = 0,
    EMB_ERR_CUDA = -1,
    EMB_ERR_ARG = -2,
    EMB_ERR_STATE = -3,
    EMB_ERR_LIMIT = -4,
    EMB_NOT_CONVERGED = 1
};

struct emb_error_sink {
    std::string msg;
};

#define EMB_CUDA(ctx, call)                                                                       \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " at " + __FILE__ + \
                         ":" + std::to_string(__LINE__);                                          \
            return EMB_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define EMB_TRY(expr)                \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ < 0) return rc__;   \
    } while (0)
