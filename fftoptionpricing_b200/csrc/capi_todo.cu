// Entry points not implemented yet return FOP_UNSUPPORTED (replaced as the kernels land).
#include "handle.cuh"
extern "C" {
int fop_carr_madan(fop_handle_t h, fop_model_t, const double*, int, int, double, double, double,
                   double, double, int, double*) { return fop::fail(h, FOP_UNSUPPORTED, "fop_carr_madan: not built yet"); }
int fop_fsts(fop_handle_t h, fop_model_t, const double*, int, int, double, double, double, double,
             int, int, int, int, double*) { return fop::fail(h, FOP_UNSUPPORTED, "fop_fsts: not built yet"); }
int fop_objfn_cf(fop_handle_t h, fop_model_t, const double*, int, const double*, const double*, int,
                 double, double, double*) { return fop::fail(h, FOP_UNSUPPORTED, "fop_objfn_cf: not built yet"); }
int fop_fft(fop_handle_t h, double*, int, int, int) { return fop::fail(h, FOP_UNSUPPORTED, "fop_fft: not built yet"); }
}
