// TEST INFRASTRUCTURE ONLY -- not part of the product path.  Backend of oracle/shims/fftw3.h.
//
// Besides the transforms it keeps a wall-clock record of the reference's xi-loop WITHOUT touching the
// reference sources: every FFT the reference executes lies inside System_3d::solve_wakefield's xi-loop
// (src/System_3d.cpp:356,460,539,567), the first one a few loop nests after the loop starts and the last
// one (slice 0's psi solve) a few loop nests before it ends.  At process exit the shim writes
//     {"first_fft_to_last_fft_s": ..., "fft_s": ..., "n_fft": ...}
// to the file named by $QPIC_REF_TIMING_FILE (if set).  bench.py --impl reference uses that span.
#include "fftw3.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../qo_fft.h"

#ifdef QPIC_SHIM_MKL
#include "mkl_dfti_min.h"
#endif

struct qpic_shim_plan_s {
    bool mkl_checked = false;
    int n0, n1;
    double * real;
    fftw_complex * cplx;
    bool forward;
    qo_fft2d plan;
#ifdef QPIC_SHIM_MKL
    DFTI_DESCRIPTOR_HANDLE handle;
#endif
};

namespace {

using clock_type = std::chrono::steady_clock;
bool g_any = false;
clock_type::time_point g_first, g_last;
double g_fft_seconds = 0.0;
long g_count = 0;
bool g_registered = false;

void write_timing() {
    const char * path = std::getenv("QPIC_REF_TIMING_FILE");
    if (!path || !g_any) return;
    FILE * f = std::fopen(path, "w");
    if (!f) return;
    const double span = std::chrono::duration<double>(g_last - g_first).count();
    std::fprintf(f, "{\"first_fft_to_last_fft_s\": %.9g, \"fft_s\": %.9g, \"n_fft\": %ld}\n", span, g_fft_seconds, g_count);
    std::fclose(f);
}

}  // namespace

extern "C" {

double * fftw_alloc_real(size_t n) { return static_cast<double *>(aligned_alloc(64, ((n * sizeof(double) + 63) / 64) * 64)); }

fftw_complex * fftw_alloc_complex(size_t n) {
    return static_cast<fftw_complex *>(aligned_alloc(64, ((n * sizeof(fftw_complex) + 63) / 64) * 64));
}

void fftw_free(void * p) { free(p); }

static fftw_plan make_plan(int n0, int n1, double * real, fftw_complex * cplx, bool forward) {
    auto * p = new qpic_shim_plan_s();
    p->n0 = n0;
    p->n1 = n1;
    p->real = real;
    p->cplx = cplx;
    p->forward = forward;
    if (qo_fft2d_init(&p->plan, n0, n1) != 0) {
        delete p;
        return nullptr;
    }
#ifdef QPIC_SHIM_MKL
    p->handle = qpic_mkl_make_r2c_2d(n0, n1, forward);
#endif
    return p;
}

fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double * in, fftw_complex * out, unsigned) {
    return make_plan(n0, n1, in, out, true);
}

fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex * in, double * out, unsigned) {
    return make_plan(n0, n1, out, in, false);
}

void fftw_execute(const fftw_plan p) {
    const auto t0 = clock_type::now();
    if (!g_any) {
        g_any = true;
        g_first = t0;
        if (!g_registered) {
            g_registered = true;
            std::atexit(write_timing);
        }
    }
#ifdef QPIC_SHIM_MKL
    if (p->handle && !p->mkl_checked) {
        // first use: run both back ends on the live data and keep MKL only if it agrees with the plain transform
        // (guards the hand-written DFTI declarations); this one call is not representative of steady-state timing
        p->mkl_checked = true;
        const size_t nr = (size_t)p->n0 * p->n1, nc = (size_t)p->n0 * (p->n1 / 2 + 1);
        std::vector<double> real_in(p->real, p->real + nr), cplx_in(reinterpret_cast<double *>(p->cplx), reinterpret_cast<double *>(p->cplx) + 2 * nc);
        double num = 0, den = 0;
        if (p->forward) {
            std::vector<double> ref(2 * nc);
            qo_rfft2_forward(&p->plan, real_in.data(), ref.data());
            DftiComputeForward(p->handle, p->real, p->cplx);
            const double * got = reinterpret_cast<double *>(p->cplx);
            for (size_t i = 0; i < 2 * nc; i++) {
                num += (got[i] - ref[i]) * (got[i] - ref[i]);
                den += ref[i] * ref[i];
            }
            if (!(num <= 1e-24 * den + 1e-300)) {
                std::fprintf(stderr, "fftw3 shim: MKL forward transform disagrees with the plain FFT (%g); using the plain FFT\n", num / (den + 1e-300));
                DftiFreeDescriptor(&p->handle);
                p->handle = nullptr;
                std::memcpy(p->cplx, ref.data(), sizeof(double) * 2 * nc);
            }
        } else {
            std::vector<double> ref(nr), scratch(cplx_in);
            qo_rfft2_backward(&p->plan, scratch.data(), ref.data());
            DftiComputeBackward(p->handle, p->cplx, p->real);
            for (size_t i = 0; i < nr; i++) {
                num += (p->real[i] - ref[i]) * (p->real[i] - ref[i]);
                den += ref[i] * ref[i];
            }
            if (!(num <= 1e-24 * den + 1e-300)) {
                std::fprintf(stderr, "fftw3 shim: MKL backward transform disagrees with the plain FFT (%g); using the plain FFT\n", num / (den + 1e-300));
                DftiFreeDescriptor(&p->handle);
                p->handle = nullptr;
                std::memcpy(p->real, ref.data(), sizeof(double) * nr);
            }
        }
    } else if (p->handle) {
        if (p->forward) {
            DftiComputeForward(p->handle, p->real, p->cplx);
        } else {
            DftiComputeBackward(p->handle, p->cplx, p->real);
        }
    } else
#endif
    {
        if (p->forward) {
            qo_rfft2_forward(&p->plan, p->real, reinterpret_cast<double *>(p->cplx));
        } else {
            qo_rfft2_backward(&p->plan, reinterpret_cast<double *>(p->cplx), p->real);
        }
    }
    const auto t1 = clock_type::now();
    g_last = t1;
    g_fft_seconds += std::chrono::duration<double>(t1 - t0).count();
    g_count++;
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    qo_fft2d_free(&p->plan);
#ifdef QPIC_SHIM_MKL
    if (p->handle) DftiFreeDescriptor(&p->handle);
#endif
    delete p;
}

int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}

}  // extern "C"
