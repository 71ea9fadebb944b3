// model_host.cu -- host-side C entry points for the GENERATED model functions (model_generated.cuh), so
// that the CPU test-suite can check the generated code against the oracle without a GPU.  Not used by
// the solver: the kernels include model_generated.cuh directly.
#include "model_generated.cuh"
extern "C" void b2model_gl4_step(const double* x, const double* u, double h, double* xn, double* sv) { model_gl4_step(x, u, h, xn, sv); }
extern "C" void b2model_rk4_step(const double* x, const double* u, double h, double* xn, double* sv) { model_rk4_step(x, u, h, xn, sv); }
extern "C" void b2model_con_h(const double* x, double px, double py, double r2, double* h, double* cx, double* cy) { model_con_h(x, px, py, r2, h, cx, cy); }
extern "C" const char* b2model_gl4_pattern(void) { return MODEL_GL4_PATTERN; }
extern "C" double b2model_r2(void) { return MODEL_R2; }
