/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * C ABI shared by the two CPU oracles:
 *   oracle/_ref/libfftop_oracle_ref.so  -- the reference's OWN headers (OptionPricing.h, fft.h/.hpp,
 *        OptionCalibration.h) compiled unchanged from /root/reference against oracle/shims/
 *        (ref_driver.cpp).  Parity is pinned against this one.
 *   oracle/libfftop_oracle_port.so      -- an independent restatement (port/oracle_port.cpp) that
 *        needs no reference source; it is itself pinned against the _ref flavour and against the
 *        committed golden vectors (tests/golden/).
 * Argument lists mirror include/fftop_b200.h minus the handle; all pointers are host pointers.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load these libraries.  The product library never does.
 */
#ifndef FFTOP_ORACLE_ABI_H
#define FFTOP_ORACLE_ABI_H
#ifdef __cplusplus
extern "C" {
#endif

const char* ora_flavor(void); /* "reference" or "port" */
int ora_num_threads(void);    /* OpenMP threads the batched loops will use */

int ora_carr_madan(int base, int tc, const double* params, int P, int N, double ada,
                   double alpha, double S0, double r, double T, int is_put, double* out);
int ora_cos(int base, int tc, const double* params, int P, const double* K_desc, int J,
            double S0, double r, const double* T, int M, int numU, int kind, double* out);
int ora_fsts(int base, int tc, const double* params, int P, int N, double xmax, double strike,
             double r, double T, int steps, int american, int payoff, int kind, double* out);
int ora_objfn_cf(int base, int tc, const double* params, int P, const double* u,
                 const double* phiHat, int m, double r, double T, double* loss);
int ora_cos_loss(int base, int tc, const double* params, int P, const double* K_desc, int J,
                 double S0, double r, const double* T, int M, int numU, const double* mkt,
                 const double* w, double* loss, double* prices_out);
/* reference radix-2 FFT (fft.hpp:49-61), in place, `batch` sequences of N interleaved complex */
int ora_fft(double* data, int batch, int N, int inverse);
/* integrateFFT / integrateIFFT (fft.hpp:64-110) of a function given by its N samples on the input
 * grid (the callable handed to the reference looks the sample up by index) */
int ora_integrate_fft(const double* fn_samples, int batch, int N, double outputMin, double inputMin,
                      double inputMax, int inverse, double* out);
/* CF value phi(u) and log-CF at u = (u_re, u_im) for one parameter set (unit tests of the CF lib) */
int ora_cf(int base, int tc, const double* params, double r, double T, double u_re, double u_im,
           double* cf_out2, double* logcf_out2);
/* optioncal::generateFOEstimate(...)(N, u)  (OptionCalibration.h:157-184); the port flavour does
 * not restate it and returns 4 -- it is pinned by golden vectors from the reference flavour */
int ora_fo_estimate(const double* strikes, const double* options, int n, double stock, double r,
                    double t, double minStrike, double maxStrike, int N, const double* u, int m,
                    double* phiHat_out);
int ora_carr_madan_strikes(int N, double ada, double S0, double* K_out);
int ora_fang_oost_strikes(double xMin, double xMax, double S0, int numX, double* K_out);
int ora_strike_underlying(double xMin, double xMax, double K, int numX, double* S_out);

#ifdef __cplusplus
}
#endif
#endif
