/*
 * fftop_b200.h -- C ABI of libfftop_b200.so, the B200 (sm_100a) batched Fourier option-pricing
 * engine.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * The reference (danielhstahl/FFTOptionPricing) is a header-only C++14 template library with no
 * ABI of its own.  Each entry point below is the batched, descriptor-driven
 * equivalent of one family of reference templates; the reference interface it replaces is cited
 * as file:line relative to the reference tree.  The reference takes an arbitrary C++ callable as
 * characteristic function (CF); device code cannot, so the CF is a *model descriptor*
 * (fop_model_t) covering exactly the CFs the reference's tests and calibration driver use
 * (test.cpp:54,570,731,887-905,989,1050,1090,1119; generateCharts.cpp:190-380).
 *
 * Conventions
 *  - Every array argument may be a HOST or a DEVICE pointer (detected with
 *    cudaPointerGetAttributes).  Host inputs are copied with cudaMemcpyAsync on the handle's
 *    stream into a workspace owned by the handle (pass pinned memory for full PCIe bandwidth);
 *    host outputs are copied back and the call returns after the stream is idle.  When all
 *    outputs are device pointers the call only enqueues work on the handle's stream
 *    (fop_get_stream) and returns; use fop_synchronize or your own events.
 *  - All arithmetic is IEEE fp64.  No fast-math, exact sincospi twiddles.
 *  - Return value: FOP_OK or an error class; fop_last_error(h) gives the message.
 *    Like the reference (which validates nothing) NaNs in parameters propagate to the
 *    outputs; only structural errors (null pointers, non power-of-two N, sizes out of the
 *    supported range, unknown enum) are reported.
 *  - Naming: there are no separate *_async variants of the pricers -- a call whose outputs are all
 *    device pointers IS asynchronous (previous bullet).  The one explicit *_async entry point,
 *    fop_gather_losses_async, additionally runs on a side stream.
 *  - A handle is bound to one device and is not thread-safe; use one handle per host thread/GPU.
 *  - There is NO CPU fallback: without a CUDA device fop_create fails with FOP_CUDA_ERROR.
 */
#ifndef FFTOP_B200_H
#define FFTOP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FOP_ABI_VERSION 1

typedef struct fop_handle_s* fop_handle_t;

enum fop_status {
  FOP_OK = 0,
  FOP_INVALID_ARG = 1,
  FOP_CUDA_ERROR = 2,
  FOP_OOM = 3,
  FOP_UNSUPPORTED = 4,
  FOP_NCCL_ERROR = 5
};

/* ---- characteristic-function model descriptor ------------------------------------------------
 * Levy base (per unit time, risk-neutral log CF psi(u), u complex):
 *   FOP_GAUSS  params [sigma]                    gaussLogCF(u, r-sigma^2/2, sigma)      test.cpp:570
 *   FOP_MERTON params [sigma, lambda, muJ, sigJ] mertonLogRNCF(u,lambda,muJ,sigJ,r,sigma) test.cpp:1119
 *   FOP_CGMY   params [sigma, C, G, M, Y]        cgmyLogRNCF(u,C,G,M,Y,r,sigma)         test.cpp:731
 *   FOP_VG     params [sigma, theta, nu]          vgLogRNCF(u,sigma,theta,nu,r) = u (r + log(1 - theta nu
 *                                                 - sigma^2 nu/2)/nu) - log(1 - u theta nu - sigma^2 nu u^2/2)/nu
 *              (variance gamma: named by the north star next to CGMY, NOT in the reference -- textbook
 *              form, parity unpinned; it has no Brownian part, so a time change acts with rho = 0)
 * Time change (4 more params appended: [speed, adaV, rho, v0Hat]):
 *   FOP_TC_NONE           phi(u) = exp(psi_r(u) T)
 *   FOP_TC_CIR            phi(u) = exp(r T u + cirLogMGF(-psi_0(u), speed, speed-adaV rho u sigma,
 *                                                        adaV, T, v0Hat))   test.cpp:887-894,1049-1058
 *                         (GAUSS + CIR is the reference's "Heston", generateCharts.cpp:231-250)
 *   FOP_TC_CIR_DIFFUSION  only the Brownian part is time changed, jumps run in calendar time
 *                                                                           test.cpp:896-905
 * Parameter order follows generateCharts.cpp:215,260,306,346.
 */
enum fop_base_model { FOP_GAUSS = 0, FOP_MERTON = 1, FOP_CGMY = 2, FOP_VG = 3 };
enum fop_time_change { FOP_TC_NONE = 0, FOP_TC_CIR = 1, FOP_TC_CIR_DIFFUSION = 2 };

typedef struct fop_model_s {
  int base;        /* enum fop_base_model  */
  int time_change; /* enum fop_time_change */
} fop_model_t;

/* number of doubles per parameter set for a model (0 if the descriptor is invalid) */
int fop_model_num_params(fop_model_t model);

/* COS output kinds: OptionPricing.h:362 (CallPrice) :398 (PutPrice) :423 :448 (Delta)
 * :525 :550 (Gamma) :474 :499 (Theta) */
enum fop_cos_kind {
  FOP_CALL_PRICE = 0,
  FOP_PUT_PRICE = 1,
  FOP_CALL_DELTA = 2,
  FOP_PUT_DELTA = 3,
  FOP_CALL_GAMMA = 4,
  FOP_PUT_GAMMA = 5,
  FOP_CALL_THETA = 6,
  FOP_PUT_THETA = 7
};

/* FSTS output kinds: OptionPricing.h:186 (Price) :210 (Delta) :235 (Theta) :259 (Gamma) */
enum fop_fsts_kind { FOP_FSTS_PRICE = 0, FOP_FSTS_DELTA = 1, FOP_FSTS_GAMMA = 2, FOP_FSTS_THETA = 3 };
/* payoff of the asset K*exp(x): call max(S-K,0), put max(K-S,0)   test.cpp:56-61,202-204 */
enum fop_payoff { FOP_PAYOFF_CALL = 0, FOP_PAYOFF_PUT = 1 };

/* ---- handle ---------------------------------------------------------------------------------- */
int fop_abi_version(void);
int fop_create(fop_handle_t* handle, int device_ordinal);
int fop_destroy(fop_handle_t handle);
const char* fop_last_error(fop_handle_t handle); /* valid until the next call on the handle */
const char* fop_status_string(int status);
void* fop_get_stream(fop_handle_t handle);       /* the handle's cudaStream_t */
int fop_synchronize(fop_handle_t handle);
/* number of this library's kernel launches since fop_create (for bench.py's gpu_launches) */
long long fop_launch_count(fop_handle_t handle);
/* contraction kernel the last fop_cos* call on this handle used (for benchmarks / tests):
 * 0 none yet, 1 fp64 tensor cores (mma.sync DMMA), 2 int8 digit planes on tcgen05 / TMEM
 * (model CFs, <= 16 maturities, numU <= 8192; FOP_GEMM_I8=0 disables it), 3 fused single kernel */
int fop_last_cos_contraction(fop_handle_t handle);

/* Optional per-kernel device timing: while enabled every kernel launch of this library is
 * bracketed by CUDA events on the handle's stream.  fop_profile_read synchronises, writes the
 * kernel names NUL-separated into names_buf, accumulated milliseconds and launch counts per name,
 * the number of names into *n_out, and resets the accumulators. */
int fop_profile_enable(fop_handle_t handle, int enable);
int fop_profile_read(fop_handle_t handle, char* names_buf, int buf_len, double* ms_out,
                     long long* launches_out, int max_entries, int* n_out);

/* ---- grid helpers (host, trivial) ------------------------------------------------------------ */
/* getCarrMadanStrikes            OptionPricing.h:131-138: K_i = S0 exp(-pi/ada + (2 pi/ada/N) i) */
int fop_carr_madan_strikes(int N, double ada, double S0, double* K_out);
/* getFangOostStrike              OptionPricing.h:98-105:  K_i = S0 exp(-(xMin + dx i))           */
int fop_fang_oost_strikes(double xMin, double xMax, double S0, int numX, double* K_out);
/* getStrikeUnderlying            OptionPricing.h:84-90:   S_i = K exp(xMin + dx i)               */
int fop_strike_underlying(double xMin, double xMax, double K, int numX, double* S_out);

/* ---- Carr-Madan ------------------------------------------------------------------------------
 * Replaces optionprice::CarrMadanCall / CarrMadanPut (OptionPricing.h:620-656, :670-704; pass
 * alpha = 1.5 for the 6-argument overloads) and CarrMadanGeneric (:582-607), batched over P
 * parameter sets.  N must be a power of two, 16 <= N <= 2^20.
 *   params [P][np]   out [P][N]   (index i <-> strike fop_carr_madan_strikes()[i])
 */
int fop_carr_madan(fop_handle_t h, fop_model_t model, const double* params, int P, int N,
                   double ada, double alpha, double S0, double r, double T, int is_put,
                   double* out);

/* ---- Fang-Oosterlee COS ----------------------------------------------------------------------
 * Replaces optionprice::FangOost{Call,Put}{Price,Delta,Gamma,Theta} (OptionPricing.h:358-559) and
 * FangOostGeneric (:325-346) including the absent fangoost::computeExpectationVectorLevy,
 * batched over P parameter sets and M maturities.
 *   K_desc [J] strikes in DESCENDING order, first and last are the synthetic far strikes
 *              (OptionPricing.h:303-312,351)
 *   T [M]      maturities;   numU <= 16384;  out [P][M][J]
 */
int fop_cos(fop_handle_t h, fop_model_t model, const double* params, int P,
            const double* K_desc, int J, double S0, double r, const double* T, int M, int numU,
            int kind, double* out);

/* All requested output kinds from ONE pass over the characteristic function (the
 * reference evaluates the CF again for each of FangOostCallPrice / Delta / Gamma / Theta,
 * OptionPricing.h:358-559; here price, delta, gamma and theta coefficients come from the same CF
 * value and share the strike basis).  kinds [nkinds] of enum fop_cos_kind, 1 <= nkinds <= 8;
 * out [nkinds][P][M][J].
 */
int fop_cos_greeks(fop_handle_t h, fop_model_t model, const double* params, int P,
                   const double* K_desc, int J, double S0, double r, const double* T, int M,
                   int numU, const int* kinds, int nkinds, double* out);

/* ---- FSTS ------------------------------------------------------------------------------------
 * Replaces optionprice::FSTS{Price,Delta,Gamma,Theta} (OptionPricing.h:185-281) and FSTSGeneric
 * (:155-184): steps = 1, american = 0 is the reference's single FFT -> multiply -> IFFT.
 * steps > 1 applies the one-step operator with dt = T/steps repeatedly; american != 0 takes
 * max(value, payoff) after every step (README.md:19 -- not implemented by the reference; defined
 * in SURVEY.md A.4).  steps > 1 needs a Levy model (FOP_TC_NONE) and kind FOP_FSTS_PRICE.
 * N power of two, 16 <= N <= 2^20 (N <= 4096: register-resident kernel, two sets per transform;
 * larger grids: general FFT engine).  out [P][N]  (node j <-> asset fop_strike_underlying(-xmax,
 * xmax, strike, N)[j])
 */
int fop_fsts(fop_handle_t h, fop_model_t model, const double* params, int P, int N, double xmax,
             double strike, double r, double T, int steps, int american, int payoff, int kind,
             double* out);

/* ---- calibration objectives ------------------------------------------------------------------
 * fop_objfn_cf replaces the closure returned by optioncal::getObjFn_arr
 * (OptionCalibration.h:185-201) evaluated for P parameter vectors at once:
 *   loss_p = (1/m) sum_l ( isnan(c) ? 5000 : |phiHat_l - c|^2 ),  c = logCF_p(1 + i u_l)
 * where logCF is the model's log characteristic function at maturity T with rate r
 * (generateCharts.cpp:231-250 uses r = 0).   u [m], phiHat [m][2] (re, im), loss [P].
 */
int fop_objfn_cf(fop_handle_t h, fop_model_t model, const double* params, int P,
                 const double* u, const double* phiHat, int m, double r, double T,
                 double* loss);

/* Price-space objective (BASELINE.json config 5; defined in SURVEY.md 8c-ii on top of
 * FangOostCallPrice, OptionPricing.h:362-382):
 *   loss_p = (1/(M J)) sum_{m,j} w_mj (FangOostCallPrice_pmj - mkt_mj)^2
 * mkt [M][J], w [M][J] or NULL (all ones), loss [P]; prices_out [P][M][J] or NULL.
 */
int fop_cos_loss(fop_handle_t h, fop_model_t model, const double* params, int P,
                 const double* K_desc, int J, double S0, double r, const double* T, int M,
                 int numU, const double* mkt, const double* w, double* loss,
                 double* prices_out);

/* Market quotes -> empirical log-CF estimate (HOST function, no GPU work): replaces
 * optioncal::generateFOEstimate(strikes, options, stock, r, t, minStrike, maxStrike)(N, uArray)
 * (OptionCalibration.h:157-184; getOptionSpline :105-156; monotone_spline.h).  strikes / options
 * [n] ascending strikes and their call prices, u [m]; phiHat_out [m][2] (re, im) is what
 * fop_objfn_cf takes.
 */
int fop_fo_estimate(const double* strikes, const double* options, int n, double stock, double r,
                    double t, double minStrike, double maxStrike, int N, const double* u, int m,
                    double* phiHat_out);

/* ---- on-device calibration of the CF-space objective -------------------------------------------
 * Replaces the reference's cuckoo::optimize(objFn, constraints, nestSize, numSims, tol, seed) over
 * the closure of getObjFn_arr (generateCharts.cpp:102-118; library cuckoo_search is un-vendored):
 * cuckoo search (Levy flights + abandonment, greedy acceptance) with `nests` nests for at most
 * `iterations` iterations or until the best objective value drops below `tol`, inside one kernel
 * launch.  `restarts` independent searches (seeded streams `seed`, restart index) run as one CTA
 * each -- use a multiple of the SM count to fill the GPU; restart 0 alone is the reference's
 * single search.  lower/upper[np]: box constraints (swarm_utils::upper_lower).
 * Outputs (host or device): params_out[restarts][np], fnval_out[restarts], iters_out[restarts]
 * (optional), *best_restart_out = argmin fnval (host int, optional).  fnval_out[i] is bit-identical
 * to fop_objfn_cf evaluated at params_out[i].  2 <= nests <= 64, 1 <= m <= 2048.
 */
int fop_calibrate_cf(fop_handle_t h, fop_model_t model, const double* lower, const double* upper,
                     const double* u, const double* phiHat, int m, double r, double T, int nests,
                     int iterations, double tol, unsigned long long seed, int restarts,
                     double* params_out, double* fnval_out, int* iters_out, int* best_restart_out);

/* Price-space calibration (SURVEY.md 8f-1 on the objective of fop_cos_loss): the same cuckoo search, the
 * objective of a candidate is (1/(M J)) sum w_mj (FangOostCallPrice_mj(theta) - mkt_mj)^2.  The
 * `restarts` searches of `nests` nests form one population of restarts x nests parameter sets that is
 * priced per half-iteration by the batched COS kernels (use restarts x nests >= ~1e4 to fill the GPU;
 * 2 <= nests <= 4096).  The tolerance is tested every 16 iterations.  Outputs as fop_calibrate_cf;
 * fnval_out[i] is bit-identical to fop_cos_loss evaluated at params_out[i] inside the same population
 * layout.  iters_out[i] = iterations run (the same for every restart).
 */
int fop_calibrate_cos(fop_handle_t h, fop_model_t model, const double* lower, const double* upper,
                      const double* K_desc, int J, double S0, double r, const double* T, int M, int numU,
                      const double* mkt, const double* w, int nests, int iterations, double tol,
                      unsigned long long seed, int restarts, double* params_out, double* fnval_out,
                      int* iters_out, int* best_restart_out);

/* ---- multi-GPU: the one collective of the path -----------------------------------
 * One process per GPU, one handle per process.  Parameter sets are sharded contiguously over the
 * ranks and priced without any exchange; fop_gather_losses all-gathers the per-set losses (NCCL
 * over NVLink / NVSwitch, on the handle's stream): loss_all[world][count], rank-major.  count must
 * be the same on every rank (pad the last shard).  Rank 0 calls fop_comm_unique_id and the host
 * program hands the 128 bytes to the other ranks; every rank then calls fop_comm_init.  NCCL is
 * loaded at run time (libnccl.so.2); without it these calls return FOP_NCCL_ERROR.
 */
int fop_comm_unique_id(char* id_out128);
int fop_comm_init(fop_handle_t h, const char* id128, int rank, int world);
int fop_gather_losses(fop_handle_t h, const double* loss_local, int count, double* loss_all);
/* Overlapped form (device pointers only): the all-gather is enqueued on a side stream of the handle,
 * ordered after everything enqueued so far on the handle's stream, and runs beside the kernels the
 * caller enqueues next (the next calibration step).  fop_comm_wait makes the handle's stream wait
 * for the last such gather (no host synchronisation).  loss_local / loss_all must stay untouched
 * until then -- alternate two buffer pairs between steps. */
int fop_gather_losses_async(fop_handle_t h, const double* loss_local, int count, double* loss_all);
int fop_comm_wait(fop_handle_t h);
int fop_comm_destroy(fop_handle_t h);

/* ---- batched FFT (the replacement for fft.hpp:49-61; exposed for parity tests) ---------------
 * In-place unnormalised transforms of `batch` contiguous length-N complex128 sequences
 * (interleaved re,im -- the layout of the reference's CArray, fft.h:12-13).
 * inverse = 0: X_m = sum_j x_j e^{-2 pi i jm/N} (fft.hpp:56);  inverse = 1: e^{+...} (fft.hpp:49).
 * N power of two, 2 <= N <= 2^20.
 */
int fop_fft(fop_handle_t h, double* data, int batch, int N, int inverse);

/* ---- integrateFFT / integrateIFFT  (fft.hpp:64-110, genericIntegration) -----------------------
 * Simpson-weighted Fourier integral of a function sampled on a uniform grid:
 *   inputDx = (inputMax - inputMin)/(N-1),  du = ((N-1) 2 pi/(inputMax - inputMin))/N
 *   x_j   = fn_j inputDx e^{i outputMin inputDx j}/3 * {1 at j = 0, N-1; 2 (j even); 4 (j odd)}
 *   X     = fft(x) (inverse = 0, integrateFFT) or ifft(x) (inverse = 1, integrateIFFT), unnormalised
 *   out_m = X_m e^{i u_m inputMin},  u_m = outputMin + m du
 * fn_samples[batch][N] complex128 = the reference's callable `fn` evaluated by the caller at
 * inputMin + inputDx j (a device function cannot be an arbitrary host callable); out[batch][N]
 * complex128 (may alias fn_samples).  The reference's optional `fnOutput(u, v)` edit is applied by
 * the caller to out.  N power of two, 16 <= N <= 2^20.
 */
int fop_integrate_fft(fop_handle_t h, const double* fn_samples, int batch, int N, double outputMin,
                      double inputMin, double inputMax, int inverse, double* out);

/* ---- caller-sampled characteristic functions / payoffs ------------------------------------------
 * The reference accepts ANY callable as characteristic function, payoff and asset map
 * (FangOostGeneric OptionPricing.h:325-333, CarrMadanGeneric :582-590, FSTSGeneric :155-163).  A
 * callable is only ever evaluated on a grid the library defines, so the host evaluates it there
 * (the *_u_grid helpers give the points) and hands the samples over; option transform, damping,
 * weights, transforms, contraction, time stepping, early exercise and the output map run on the
 * GPU exactly as for the built-in models.  include/fftop/OptionPricing.h wraps these behind the
 * reference's own signatures taking std::function / lambdas.
 *
 * COS:  u_out[numU] = k pi/(x_max - x_min);  cf_samples[P][M][numU][2] = cf_{p,m}(i u_k) (re, im),
 *       i.e. what the reference's `cf` returns for the complex argument (0, u_k) at maturity T_m.
 * Carr-Madan: u_out2[N][2] = (alpha + 1, j ada) -- the argument CallAug passes to cf
 *       (OptionPricing.h:564-567); cf_samples[P][N][2].
 * FSTS: w_out[N] = getUDomain(vmax, du, k) (OptionPricing.h:19-23), the CF is sampled at (0, w_k) for
 *       the step length T/steps; payoff[N] (payoff_per_set = 0) or [P][N] = payoff(getAssetPrice(x_j));
 *       asset[N] = getAssetPrice(x_j), needed for kind delta / gamma only (else NULL).
 */
int fop_cos_u_grid(const double* K_desc, int J, double S0, int numU, double* u_out);
int fop_cos_cf_samples(fop_handle_t h, const double* cf_samples, int P, const double* K_desc, int J,
                       double S0, double r, const double* T, int M, int numU, int kind, double* out);
int fop_carr_madan_u_grid(int N, double ada, double alpha, double* u_out2);
int fop_carr_madan_cf_samples(fop_handle_t h, const double* cf_samples, int P, int N, double ada,
                              double alpha, double S0, double r, double T, int is_put, double* out);
int fop_fsts_u_grid(int N, double xmax, double* w_out);
int fop_fsts_samples(fop_handle_t h, const double* cf_samples, const double* payoff, int payoff_per_set,
                     const double* asset, int P, int N, double xmax, double r, double T, int steps,
                     int american, int kind, double* out);

/* measured fp64 FMA throughput of the device in TFLOP/s (roofline denominator, bench.py) */
int fop_measure_fp64_peak(fop_handle_t h, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* FFTOP_B200_H */
