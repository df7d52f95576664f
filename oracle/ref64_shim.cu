// TEST INFRASTRUCTURE ONLY (tests/ may load the library this builds into; the product never links it).
// C entry points over the UNMODIFIED reference class `Wavelets` (pdwt/src/wt.h:20-76) compiled with -DDOUBLEPRECISION
// (pdwt/src/filters.h:16-30: DTYPE = double; pdwt/Makefile:36-39 `libpdwtd.so`).  The reference's own Python wrapper
// binds float buffers only (src/pypwt.pyx:29-61), so its double build has no binding: this shim is the missing one,
// used by tests/test_gpu_f64_vs_pdwt.py to run the reference's double build beside `Wavelets64` on the GPU.
// Built by `make -C oracle ref64` from the sources where they lie under /root/reference into oracle/_ref/.
#include "wt.h"

extern "C" {
void* r64_create(double* img, int Nr, int Nc, const char* wname, int levels, int do_separable, int do_cycle_spinning,
                 int do_swt, int ndim) {
    return new Wavelets(img, Nr, Nc, wname, levels, 1, do_separable, do_cycle_spinning, do_swt, ndim);
}
void r64_destroy(void* w) { delete static_cast<Wavelets*>(w); }
int r64_levels(void* w) { return static_cast<Wavelets*>(w)->winfos.nlevels; }
int r64_state(void* w) { return (int)static_cast<Wavelets*>(w)->state; }
int r64_shift_r(void* w) { return static_cast<Wavelets*>(w)->current_shift_r; }
int r64_shift_c(void* w) { return static_cast<Wavelets*>(w)->current_shift_c; }
void r64_forward(void* w) { static_cast<Wavelets*>(w)->forward(); }
void r64_inverse(void* w) { static_cast<Wavelets*>(w)->inverse(); }
void r64_soft_threshold(void* w, double beta, int app, int normalize) { static_cast<Wavelets*>(w)->soft_threshold(beta, app, normalize); }
void r64_hard_threshold(void* w, double beta, int app, int normalize) { static_cast<Wavelets*>(w)->hard_threshold(beta, app, normalize); }
void r64_shrink(void* w, double beta, int app) { static_cast<Wavelets*>(w)->shrink(beta, app); }
double r64_norm1(void* w) { return static_cast<Wavelets*>(w)->norm1(); }
double r64_norm2sq(void* w) { return static_cast<Wavelets*>(w)->norm2sq(); }
int r64_get_image(void* w, double* dst) { return static_cast<Wavelets*>(w)->get_image(dst); }
int r64_get_coeff(void* w, double* dst, int num) { return static_cast<Wavelets*>(w)->get_coeff(dst, num); }
void r64_set_image(void* w, double* img) { static_cast<Wavelets*>(w)->set_image(img, 0); }
}
