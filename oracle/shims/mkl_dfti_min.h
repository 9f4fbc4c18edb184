/* TEST INFRASTRUCTURE ONLY.  Minimal hand-written declarations of Intel MKL's DFTI C interface (the real mkl_dfti.h is not
 * in this image).  The symbols are exported by torch's libtorch_cpu.so, which statically contains MKL; the timed build of
 * the unmodified reference (oracle/_ref/quasistatic_pic_fast) links against that library so that its CPU baseline uses
 * an FFTW/MKL-class FFT, as the reference would with FFTW.  Enum values follow MKL's public header; the shim
 * self-checks the first transform of every plan against oracle/qo_fft.c and falls back to it on any mismatch. */
#ifndef QPIC_MKL_DFTI_MIN_H
#define QPIC_MKL_DFTI_MIN_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DFTI_DESCRIPTOR * DFTI_DESCRIPTOR_HANDLE;
typedef long MKL_LONG;

enum DFTI_CONFIG_PARAM {
    DFTI_FORWARD_DOMAIN = 0,
    DFTI_DIMENSION = 1,
    DFTI_LENGTHS = 2,
    DFTI_PRECISION = 3,
    DFTI_FORWARD_SCALE = 4,
    DFTI_BACKWARD_SCALE = 5,
    DFTI_NUMBER_OF_TRANSFORMS = 7,
    DFTI_COMPLEX_STORAGE = 8,
    DFTI_REAL_STORAGE = 9,
    DFTI_CONJUGATE_EVEN_STORAGE = 10,
    DFTI_PLACEMENT = 11,
    DFTI_INPUT_STRIDES = 12,
    DFTI_OUTPUT_STRIDES = 13,
    DFTI_INPUT_DISTANCE = 14,
    DFTI_OUTPUT_DISTANCE = 15,
    DFTI_THREAD_LIMIT = 27
};

enum DFTI_CONFIG_VALUE {
    DFTI_COMPLEX = 32,
    DFTI_REAL = 33,
    DFTI_SINGLE = 35,
    DFTI_DOUBLE = 36,
    DFTI_COMPLEX_COMPLEX = 39,
    DFTI_COMPLEX_REAL = 40,
    DFTI_REAL_COMPLEX = 41,
    DFTI_REAL_REAL = 42,
    DFTI_INPLACE = 43,
    DFTI_NOT_INPLACE = 44
};

MKL_LONG DftiCreateDescriptor_d_md(DFTI_DESCRIPTOR_HANDLE *, MKL_LONG dimension, const MKL_LONG * lengths);
MKL_LONG DftiSetValue(DFTI_DESCRIPTOR_HANDLE, enum DFTI_CONFIG_PARAM, ...);
MKL_LONG DftiCommitDescriptor(DFTI_DESCRIPTOR_HANDLE);
MKL_LONG DftiComputeForward(DFTI_DESCRIPTOR_HANDLE, void *, ...);
MKL_LONG DftiComputeBackward(DFTI_DESCRIPTOR_HANDLE, void *, ...);
MKL_LONG DftiFreeDescriptor(DFTI_DESCRIPTOR_HANDLE *);

#ifdef __cplusplus
}

/* a 2-D real <-> conjugate-even descriptor with FFTW's r2c/c2r layout: n0 x n1 reals, n0 x (n1/2+1) complex */
static inline DFTI_DESCRIPTOR_HANDLE qpic_mkl_make_r2c_2d(int n0, int n1, bool forward) {
    DFTI_DESCRIPTOR_HANDLE h = nullptr;
    MKL_LONG len[2] = {n0, n1};
    if (DftiCreateDescriptor_d_md(&h, 2, len) != 0 || !h) return nullptr;
    MKL_LONG rs[3] = {0, n1, 1};
    MKL_LONG cs[3] = {0, n1 / 2 + 1, 1};
    bool ok = DftiSetValue(h, DFTI_CONJUGATE_EVEN_STORAGE, DFTI_COMPLEX_COMPLEX) == 0 && DftiSetValue(h, DFTI_PLACEMENT, DFTI_NOT_INPLACE) == 0 &&
              DftiSetValue(h, DFTI_INPUT_STRIDES, forward ? rs : cs) == 0 && DftiSetValue(h, DFTI_OUTPUT_STRIDES, forward ? cs : rs) == 0 &&
              DftiCommitDescriptor(h) == 0;
    if (!ok) {
        DftiFreeDescriptor(&h);
        return nullptr;
    }
    return h;
}
#endif

#endif
