// TEST INFRASTRUCTURE ONLY: executes the __host__ __device__ per-cell functions of fluidgym_b200/csrc/extruded3_b200.cuh on the
// CPU (plain loops over planes and cells) so that tests/test_extruded_host.py can compare the very code the CUDA kernels run
// with the numpy specification and the reference trace without a GPU.  Never linked into the product library.
#include <stddef.h>
#include <stdint.h>

#include "fluidgym_b200.h"
#define X3_HOST_ONLY
#include "extruded3_b200.cuh"

static X3Tab make(const fgb_tables *t, int nz, float hz) { X3Tab x; x.t = *t; x.nz = nz; x.hz = hz; return x; }

extern "C" void xh_setup_advection(const fgb_tables *t, int nz, float hz, const float *u, const float *ures, const float *bvel, float dt,
                                   float *coff, float *A, float *rhs, int with_matrix) {
    const X3Tab x = make(t, nz, hz);
    for (int k = 0; k < nz; ++k)
        for (int g = 0; g < t->N; ++g) x3_setup_advection_cell(x, k, g, u, ures, bvel, dt, coff, A, rhs, with_matrix);
}
extern "C" void xh_pressure_matrix(const fgb_tables *t, int nz, float hz, const float *A, float *poff, float *pdiag) {
    const X3Tab x = make(t, nz, hz);
    for (int k = 0; k < nz; ++k)
        for (int g = 0; g < t->N; ++g) x3_pressure_matrix_cell(x, k, g, A, poff, pdiag);
}
extern "C" void xh_hbya(const fgb_tables *t, int nz, float hz, const float *u, const float *ures, const float *bvel, const float *coff,
                        const float *A, float dt, float *hb) {
    const X3Tab x = make(t, nz, hz);
    for (int k = 0; k < nz; ++k)
        for (int g = 0; g < t->N; ++g) x3_hbya_cell(x, k, g, u, ures, bvel, coff, A, dt, hb);
}
extern "C" void xh_divergence(const fgb_tables *t, int nz, float hz, const float *hb, const float *bvel, const float *pprev, const float *A,
                              float *div) {
    const X3Tab x = make(t, nz, hz);
    for (int k = 0; k < nz; ++k)
        for (int g = 0; g < t->N; ++g) x3_divergence_cell(x, k, g, hb, bvel, pprev, A, div);
}
extern "C" void xh_correct(const fgb_tables *t, int nz, float hz, const float *hb, const float *p, const float *A, float *uout) {
    const X3Tab x = make(t, nz, hz);
    for (int k = 0; k < nz; ++k)
        for (int g = 0; g < t->N; ++g) x3_correct_cell(x, k, g, hb, p, A, uout);
}
