// C ABI: fop_cos, fop_cos_loss  (Fang-Oosterlee COS; include/fftop_b200.h)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cos_fused_ws.cuh"
#include "cos_gemm_i8.cuh"
#include "cos_gemm_tma.cuh"
#include "cos_gemm_variants.h"
#include "cos_host_tables.h"
#include "handle.cuh"

namespace fop {  // capi_cos_fused.cu
int cos_fused_pick_nb(int J);
int cos_fused_consts(fop_handle_t h, int base, int tc, const double* params, int np, int P, ModelConsts* out);
int cos_fused_launch(fop_handle_t h, CosFusedArgs& a, int NB);
// capi_cos_coeff.cu
bool cos_coeff_launch_spec(fop_handle_t h, const CosCoeffArgs& a, int grid_blocks_per_sm);
}  // namespace fop

namespace {
using namespace fop;

// kernel tables live in their own translation units (capi_cos_gemm.cu, capi_cos_gemm_ex.cu) so that
// the fourteen tensor-core instantiations compile in parallel with this file
const GemmExVariant* pick_gemm_ex(int J) {
  if (std::getenv("FOP_GEMM_NO_EX")) return nullptr;
  int n = 0;
  const GemmExVariant* t = fop::gemm_ex_variants(&n);
  for (int i = 0; i < n; ++i)
    if (J == 8 * t[i].NB + t[i].EX) return &t[i];
  return nullptr;
}
// ---- TMA tensor maps for the contraction (cos_gemm_tma.cuh); the driver entry point is resolved
// through the runtime, so the library still links against cudart only
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = []() -> TmapEncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<TmapEncodeFn>(p);
  }();
  return fn;
}
// fp64 matrix [dim1][dim0] (dim0 contiguous, row pitch `pitch` doubles), boxes of box1 x box0
bool make_tmap_2d(CUtensorMap* m, const double* base, unsigned long long dim0, unsigned long long dim1,
                  unsigned long long pitch, unsigned box0, unsigned box1, bool swizzle128) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {dim0, dim1};
  const cuuint64_t strides[1] = {pitch * 8};
  const cuuint32_t box[2] = {box0, box1};
  const cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// digit planes [dim2][dim1][dim0] of bytes (dim0 contiguous), boxes of box2 x box1 x box0, 128-byte swizzle
bool make_tmap_3d_u8(CUtensorMap* m, const unsigned char* base, unsigned long long dim0, unsigned long long dim1,
                     unsigned long long dim2, unsigned long long pitch1, unsigned long long pitch2, unsigned box0,
                     unsigned box1, unsigned box2) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {dim0, dim1, dim2};
  const cuuint64_t strides[2] = {pitch1, pitch2};
  const cuuint32_t box[3] = {box0, box1, box2};
  const cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<unsigned char*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
const GemmTmaVariant* pick_gemm_tma(int NB, int EX) {
  if (const char* e = std::getenv("FOP_GEMM_TMA"))
    if (e[0] == '0') return nullptr;
  int n = 0;
  const GemmTmaVariant* t = fop::gemm_tma_variants(&n);
  for (int i = 0; i < n; ++i)
    if (t[i].NB == NB && t[i].EX == EX) return &t[i];
  return nullptr;
}

// strike-block width that wastes the fewest padded columns (ties -> wider block)
const GemmVariant* pick_gemm(int J, int* nblk_out) {
  const GemmVariant* best = nullptr;
  long long best_cost = 0;
  int nv = 0;
  const GemmVariant* tab = fop::gemm_variants(&nv);
  for (int iv = 0; iv < nv; ++iv) {
    const GemmVariant& v = tab[iv];
    const int JB = 8 * v.NB, nblk = (J + JB - 1) / JB;
    const long long cost = (long long)nblk * JB;
    if (!best || cost < best_cost || (cost == best_cost && v.NB > best->NB)) {
      best = &v;
      best_cost = cost;
      *nblk_out = nblk;
    }
  }
  return best;
}

// kinds[nk]: output kinds (fop_cos_kind); out = [nk][P][M][J] (or nullptr in loss-only mode)
// cf_samples != nullptr: the characteristic function is given by its samples phi_{p,m}(i u_k)
// ([P][M][numU][2], fop_cos_cf_samples) instead of (model, params)
int cos_run(fop_handle_t h, fop_model_t model, const double* params, int P, const double* K_desc,
            int J, double S0, double r, const double* T, int M, int numU, const int* kinds, int nk,
            const double* mkt, const double* w, double* out, double* loss,
            const double* cf_samples = nullptr) {
  if (!h) return FOP_INVALID_ARG;
  const int np = cf_samples ? 0 : num_params(model.base, model.time_change);
  if (!cf_samples && np == 0)
    return fail(h, FOP_INVALID_ARG, "unknown model (%d,%d)", model.base, model.time_change);
  if ((!params && !cf_samples) || !K_desc || !T) return fail(h, FOP_INVALID_ARG, "null input pointer");
  if (!out && !loss) return fail(h, FOP_INVALID_ARG, "no output requested");
  if (P < 0 || J < 2 || M < 1 || numU < 1 || numU > 16384 || !kinds || nk < 1 || nk > 8)
    return fail(h, FOP_INVALID_ARG, "bad sizes P=%d J=%d M=%d numU=%d nkinds=%d", P, J, M, numU, nk);
  int greek_mask = 0, plane_of[4] = {0, 0, 0, 0}, ngreeks = 0;
  for (int i = 0; i < nk; ++i) {
    if (kinds[i] < 0 || kinds[i] > 7) return fail(h, FOP_INVALID_ARG, "unknown output kind %d", kinds[i]);
    greek_mask |= 1 << (kinds[i] >> 1);
  }
  for (int g = 0; g < 4; ++g)
    if (greek_mask >> g & 1) plane_of[g] = ngreeks++;
  if (loss && !mkt) return fail(h, FOP_INVALID_ARG, "loss requested without market prices");
  if (P == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));

  int nblk = 0;
  const GemmVariant* gv = pick_gemm(J, &nblk);
  // Path.  Default: the split path (coefficient kernel, then a GEMM over strike blocks that
  // re-uses each coefficient nblk times).  FOP_COS_PATH=fused selects, for few strikes and one
  // output kind, the warp-specialised single kernel of cos_fused_ws.cuh (CF producers + tensor-core
  // consumers in one CTA, no coefficient matrix in HBM).  It is kept selectable and tested, not
  // the default: on config 5 it measures 6.4 ms against 3.8 + 2.5 ms for the split path (both
  // halves want the same fp64 pipe; profiles/README.md).
  const int fnb = cos_fused_pick_nb(J);
  const int spt = M <= FW_MAX_M ? FW_ROWS / M : 0;
  bool fused = false;
  if (const char* force = std::getenv("FOP_COS_PATH"))
    fused = !std::strcmp(force, "fused") && nk == 1 && fnb > 0 && spt > 0 && !cf_samples;
  const GemmExVariant* gx = fused ? nullptr : pick_gemm_ex(J);
  if (fused || gx) nblk = 1;
  const int JB = fused ? 8 * fnb : gx ? 8 * gx->NB : 8 * gv->NB;
  const int Jp = fused ? dm_bs(fnb) : gx ? JB + 2 : nblk * JB;  // row stride of the basis table
  // rows of C / of the basis are zero-padded to a multiple of the contraction's k-stage
  const int KK = ((2 * numU + Q8_KC - 1) / Q8_KC) * Q8_KC;  // (a multiple of DM_KC, TM_KC and FW_KC too)
  // int8 tensor-core contraction (cos_gemm_i8.cuh): every output kind of a model CF with up to 16
  // maturities (market tables of a strike block in shared memory) and numU <= 8192 (int32 accumulators
  // cannot overflow); sampled CFs, longer maturity lists, or FOP_GEMM_I8=0, take the fp64 tensor-core
  // kernels
  const int nblk8 = (J + Q8_NQ - 1) / Q8_NQ, J8 = nblk8 * Q8_NQ;
  bool use_i8 = !fused && !cf_samples && cos_i8_smem(M) + 512 <= h->smem_optin && KK <= 16384 &&
                tmap_encoder() != nullptr;
  if (const char* e = std::getenv("FOP_GEMM_I8"))
    if (e[0] == '0') use_i8 = false;
  if (std::getenv("FOP_COEFF_GENERIC")) use_i8 = false;  // (the generic coefficient kernel has no quantised output)

  // Strike-grid tables (u_k, V_k, strikes, maturities, cos/sin basis): rebuilt only when
  // (K_desc, S0, T, numU) change -- a calibration loop keeps them fixed while theta varies.
  std::vector<double> Kh(J), Th(M);
  // host descriptors are read directly (a cudaMemcpy would synchronise the device on every call)
  if (is_device_ptr(K_desc)) FOP_CUDA(h, cudaMemcpy(Kh.data(), K_desc, sizeof(double) * J, cudaMemcpyDeviceToHost));
  else std::memcpy(Kh.data(), K_desc, sizeof(double) * J);
  if (is_device_ptr(T)) FOP_CUDA(h, cudaMemcpy(Th.data(), T, sizeof(double) * M, cudaMemcpyDeviceToHost));
  else std::memcpy(Th.data(), T, sizeof(double) * M);
  auto& cc = h->cos_cache;
  const size_t npow_off = 2 * (size_t)numU + 2 * (size_t)J + M;   // [M] int32 n_m, in M doubles of space
  const size_t tab_doubles = npow_off + M;
  const bool hit = cc.dev && cc.S0 == S0 && cc.numU == numU && cc.JB == JB && cc.Jp == Jp &&
                   cc.KK == KK && cc.K == Kh && cc.T == Th;
  if (!hit) {
    // the key is invalid until the new tables are completely in place: an early error return below
    // must not leave the old key over new or partial tables
    cc.numU = 0;
    cc.K.clear();
    std::vector<double> tab(tab_doubles);
    double* uk = tab.data();
    double* vk = uk + numU;
    double* x = vk + numU;
    for (int j = 0; j < J; ++j) x[j] = std::log(S0 / Kh[j]);  // getXFromK, OptionPricing.h:110-117
    std::copy(Kh.begin(), Kh.end(), x + J);
    std::copy(Th.begin(), Th.end(), x + 2 * J);
    cc.dt = 0.0;
    std::vector<int> npow;
    std::vector<int> steps;
    if (maturity_grid(Th, &cc.dt, &npow) && !std::getenv("FOP_COS_NO_TGRID")) steps = maturity_steps(npow);
    if (!steps.empty()) std::memcpy(tab.data() + npow_off, steps.data(), sizeof(int) * M);
    else cc.dt = 0.0;
    const double xMin = x[0], xMax = x[J - 1];
    const double du = M_PI / (xMax - xMin), cp = 2.0 / (xMax - xMin);
    for (int k = 0; k < numU; ++k) {
      uk[k] = du * k;
      vk[k] = vk_put(xMin, uk[k], k) * cp * (k == 0 ? 0.5 : 1.0);
    }
    cc.vk_host.assign(vk, vk + numU);
    cc.uk_host.assign(uk, uk + numU);
    const size_t bytes = (tab_doubles + (size_t)KK * Jp + 2 * (size_t)numU) * 8 + 512 + 3 * (((size_t)Q8_SLICES * J8 * KK + 255) & ~(size_t)255);
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));  // earlier launches may still read the old tables
    if (cc.bytes < bytes) {
      if (cc.dev) cudaFree(cc.dev);
      cc.dev = nullptr;
      cc.bytes = 0;
      FOP_CUDA(h, cudaMalloc(&cc.dev, bytes));
      cc.bytes = bytes;
    }
    FOP_CUDA(h, cudaMemcpy(cc.dev, tab.data(), tab_doubles * 8, cudaMemcpyHostToDevice));
    const size_t boff = (tab_doubles + 1) & ~(size_t)1;  // 16-byte aligned basis
    FOP_CUDA(h, cudaMemsetAsync(cc.dev + boff, 0, (size_t)KK * Jp * 8, h->stream));  // zero padding rows
    const int nb = numU * Jp;
    {
      ProfScope ps(h, "cos_basis_kernel");
      cos_basis_kernel<<<(nb + 255) / 256, 256, 0, h->stream>>>(cc.dev, cc.dev + 2 * numU, numU, J,
                                                               Jp, cc.dev + boff);
    }
    FOP_LAUNCH_CHECK(h);
    cc.q8_valid[0] = cc.q8_valid[1] = cc.q8_valid[2] = false;  // fixed-point tables of the int8 contraction: rebuilt on demand
    cc.q8_vp_valid = cc.q8_vt_valid = false;
    cc.K = Kh;  cc.T = Th;  cc.S0 = S0;  cc.numU = numU;  cc.JB = JB;  cc.Jp = Jp;  cc.KK = KK;
  }
  const double* d_tab = cc.dev;
  const double* d_K = cc.dev + 2 * numU + J;
  const double* d_T = cc.dev + 2 * numU + 2 * J;
  const int* d_npow = cc.dt > 0.0 ? reinterpret_cast<const int*>(cc.dev + npow_off) : nullptr;
  const double* d_basis = cc.dev + ((tab_doubles + 1) & ~(size_t)1);
  // Fixed-point tables of the int8 contraction (cos_gemm_i8.cuh).  Price, delta and gamma multiply phi
  // by a factor c_k that does not depend on the parameters (1, i u_k, -(u_k^2 + i u_k);
  // OptionPricing.h:44-73), so they share ONE set of coefficient planes (phi 2^45) and differ in the
  // basis planes only: Bas_g[2k][j], Bas_g[2k+1][j] = Re, -Im of c_k e^{i u_k (x_j - x_0)} times
  // V_k cp / amax_g.  Theta's factor -(log phi - r) depends on phi: it gets its own planes
  // (transform / (4 + |r|), a bound on |phi (log phi - r)| for |phi| <= 1) over the price basis.
  double* d_v8p = cc.dev + ((tab_doubles + 1) & ~(size_t)1) + (size_t)KK * Jp;   // [numU] 2^45
  double* d_v8t = d_v8p + numU;   // [numU] 2^45 / (4 + |r|); MUST follow d_v8p: the two-set kernel reads vk[numU]
  const size_t b8_bytes = ((size_t)Q8_SLICES * J8 * KK + 255) & ~(size_t)255;
  unsigned char* d_b8 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(d_v8t + numU) + 255) & ~(uintptr_t)255);
  bool q8_price_planes = false, q8_theta_planes = false;
  const double theta_bound = 4.0 + std::fabs(r);
  if (use_i8) {
    bool need_tab[3] = {false, false, false};
    for (int i = 0; i < nk; ++i) {
      const int g = kinds[i] >> 1;
      if (g == 3) q8_theta_planes = true;
      else q8_price_planes = true;
      need_tab[g == 3 ? 0 : g] = true;
    }
    bool synced = false;
    auto sync_once = [&]() -> int {  // earlier launches may still read the tables that are about to change
      if (!synced) FOP_CUDA(h, cudaStreamSynchronize(h->stream));
      synced = true;
      return FOP_OK;
    };
    for (int g = 0; g < 3 && use_i8; ++g) {
      if (!need_tab[g] || cc.q8_valid[g]) continue;
      double amax = 0.0;
      for (int k = 0; k < numU; ++k) {
        const double u = cc.uk_host[k];
        const double f = g == 0 ? 1.0 : g == 1 ? u : u * std::sqrt(1.0 + u * u);
        amax = std::max(amax, std::fabs(cc.vk_host[k]) * f);
      }
      if (!(amax > 0.0) || !std::isfinite(amax)) {
        use_i8 = false;
        break;
      }
      FOP_TRY(sync_once());
      const long long n8 = (long long)J8 * KK;
      {
        ProfScope ps(h, "cos_basis_q8_kernel");
        cos_basis_q8_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, h->stream>>>(
            cc.dev, cc.dev + numU, cc.dev + 2 * numU, g, 1.0 / amax, numU, KK, J, J8, d_b8 + (size_t)g * b8_bytes);
      }
      FOP_LAUNCH_CHECK(h);
      cc.q8_amax[g] = amax;
      cc.q8_valid[g] = true;
    }
    if (use_i8 && q8_price_planes && !cc.q8_vp_valid) {
      std::vector<double> v8(numU, Q8_A_SCALE);
      FOP_TRY(sync_once());
      FOP_CUDA(h, cudaMemcpy(d_v8p, v8.data(), sizeof(double) * numU, cudaMemcpyHostToDevice));
      cc.q8_vp_valid = true;
    }
    if (use_i8 && q8_theta_planes && (!cc.q8_vt_valid || cc.q8_theta_bound != theta_bound)) {
      std::vector<double> v8(numU, Q8_A_SCALE / theta_bound);
      FOP_TRY(sync_once());
      FOP_CUDA(h, cudaMemcpy(d_v8t, v8.data(), sizeof(double) * numU, cudaMemcpyHostToDevice));
      cc.q8_vt_valid = true;
      cc.q8_theta_bound = theta_bound;
    }
  }
  if (use_i8) nblk = 2 * nblk8;  // two partial row losses per (set, maturity, strike block): one per column half

  const bool out_host = out && !is_device_ptr(out);
  const bool loss_host = loss && !is_device_ptr(loss);
  // Parameter sets are processed in chunks: bounds the coefficient matrix C (4 KiB per row at
  // numU = 256) to 4 GiB and the staging of a host-bound price matrix to 1 GiB.
  long long chunkP = P;
  // bytes of coefficient matrix per chunk: 4 GiB of the 180 GB (one chunk for the 125 000-set
  // shard of config 5; 1 GiB chunks measured 1.4 % slower, 512 MiB 9 %: launch tails)
  long long c_budget = 4ll << 30;
  if (const char* e = std::getenv("FOP_COS_CHUNK_MIB")) c_budget = std::max(1ll, atoll(e)) << 20;
  if (!fused)
    chunkP = std::min<long long>(chunkP, std::max<long long>(1, c_budget / ((long long)M * KK * 8 * ngreeks)));
  if (out_host) chunkP = std::min<long long>(chunkP, std::max<long long>(1, (1ll << 30) / ((long long)M * J * 8)));
  const long long chunk_rows = chunkP * M;
  const int lossblk = nblk;

  Arena ar(h);
  size_t need = cf_samples ? stage_bytes(cf_samples, (size_t)P * M * numU * 16) : stage_bytes(params, (size_t)P * np * 8);
  if (mkt) need += stage_bytes(mkt, (size_t)M * J * 8);
  if (w) need += stage_bytes(w, (size_t)M * J * 8);
  need += fused ? Arena::pad((size_t)P * sizeof(ModelConsts)) : Arena::pad((size_t)chunk_rows * KK * 8 * ngreeks);
  if (out_host) need += Arena::pad((size_t)chunk_rows * J * 8);
  if (loss) need += Arena::pad((size_t)P * M * lossblk * 8) + (loss_host ? Arena::pad((size_t)P * 8) : 0);
  if (use_i8) need += Arena::pad(2 * (size_t)chunk_rows);
  FOP_TRY(ar.reserve(need));

  const double *d_params = nullptr, *d_mkt = nullptr, *d_w = nullptr;
  const double* d_phi = nullptr;
  if (cf_samples) FOP_TRY(stage_in(h, ar, cf_samples, (size_t)P * M * numU * 2, &d_phi));
  else FOP_TRY(stage_in(h, ar, params, (size_t)P * np, &d_params));
  if (mkt) FOP_TRY(stage_in(h, ar, mkt, (size_t)M * J, &d_mkt));
  if (w) FOP_TRY(stage_in(h, ar, w, (size_t)M * J, &d_w));
  double* d_C = fused ? nullptr : ar.alloc<double>((size_t)chunk_rows * KK * ngreeks);
  ModelConsts* d_consts = fused ? ar.alloc<ModelConsts>(P) : nullptr;
  if (fused) FOP_TRY(cos_fused_consts(h, model.base, model.time_change, d_params, np, P, d_consts));
  double* d_out_tmp = out_host ? ar.alloc<double>((size_t)chunk_rows * J) : nullptr;
  double* d_rowloss = loss ? ar.alloc<double>((size_t)P * M * lossblk) : nullptr;
  double* d_loss = loss ? (loss_host ? ar.alloc<double>(P) : loss) : nullptr;
  unsigned char* d_rowflag = use_i8 ? ar.alloc<unsigned char>(2 * (size_t)chunk_rows) : nullptr;  // price | theta planes
  CUtensorMap tmB8[3];
  if (use_i8) {
    for (int g = 0; g < 3; ++g)
      if (cc.q8_valid[g] &&
          !make_tmap_3d_u8(&tmB8[g], d_b8 + (size_t)g * b8_bytes, (unsigned long long)KK, (unsigned long long)J8, Q8_SLICES,
                           (unsigned long long)KK, (unsigned long long)J8 * KK, Q8_KC, Q8_NQ, Q8_SLICES))
        return fail(h, FOP_CUDA_ERROR, "tensor map (basis digit planes) could not be encoded");
    FOP_CUDA(h, cudaFuncSetAttribute(cos_gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)cos_i8_smem(M)));
  }

  // 32-row warp tiles (4 warps per CTA) measure 1-2 % faster than 16-row tiles (8 warps) for both
  // 66 and 1024 strikes since the epilogue runs from registers; FOP_GEMM_MB=2 selects the latter
  int mb = 4;
  if (const char* e = std::getenv("FOP_GEMM_MB")) mb = atoi(e) == 2 ? 2 : 4;
  if (gx) mb = 4;
  auto gemm_kernel = gx ? gx->kern : mb == 2 ? gv->dmma2 : gv->dmma4;
  const size_t gemm_smem = cos_dmma_smem(gx ? gx->NB : gv->NB);
  FOP_CUDA(h, cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
  // default: operands by TMA into an mbarrier ring (cos_gemm_tma.cuh); FOP_GEMM_TMA=0 / FOP_GEMM_MB=2:
  // the cp.async kernels above
  const int g_nb = gx ? gx->NB : gv->NB, g_ex = gx ? gx->EX : 0;
  const GemmTmaVariant* gt = (!fused && mb == 4 && tmap_encoder()) ? pick_gemm_tma(g_nb, g_ex) : nullptr;
  CUtensorMap tmB;
  if (gt) {
    if (!make_tmap_2d(&tmB, d_basis, (unsigned long long)Jp, (unsigned long long)KK, (unsigned long long)Jp,
                      (unsigned)dm_bs(g_nb), TM_KC, false))
      gt = nullptr;
    else
      FOP_CUDA(h, cudaFuncSetAttribute(gt->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cos_tma_smem(g_nb)));
  }

  h->last_cos_contraction = fused ? 3 : use_i8 ? 2 : 1;
  for (long long p0 = 0; p0 < P; p0 += chunkP) {
    const int pc = (int)std::min<long long>(chunkP, P - p0);
    const long long rows = (long long)pc * M;
    if (fused) {
      const int kind = kinds[0];
      double* out_k = out ? out + (size_t)p0 * M * J : nullptr;
      CosFusedArgs fa;
      CosCoeffArgs& ca = fa.ca;
      ca.base = model.base;  ca.tc = model.time_change;
      ca.greek_mask = greek_mask;  ca.plane = 0;
      ca.P = pc;  ca.M = M;  ca.K = numU;  ca.Kp = KK / 2;  ca.np = np;  ca.spb = 0;
      ca.r = r;
      ca.params = d_params + (size_t)p0 * np;
      ca.T = d_T;  ca.npow = d_npow;  ca.dt = cc.dt;  ca.uk = d_tab;  ca.vk = d_tab + numU;  ca.C = nullptr;
      ca.C8 = nullptr;  ca.rowflag = nullptr;
      fa.kind = kind;  fa.J = J;  fa.nchunks = KK / FW_KC;  fa.SPT = spt;  fa.depth = 0;
      fa.S0 = S0;
      fa.consts = d_consts + p0;
      fa.basis = d_basis;  fa.strikes = d_K;
      fa.out = out ? (out_host ? d_out_tmp : out_k) : nullptr;
      fa.vec2 = (J % 2 == 0) && (reinterpret_cast<uintptr_t>(fa.out) % 16 == 0);
      fa.mkt = d_mkt;  fa.w = d_w;
      fa.rowloss = loss ? d_rowloss + (size_t)p0 * M : nullptr;
      FOP_TRY(cos_fused_launch(h, fa, fnb));
      if (out_host)
        FOP_CUDA(h, cudaMemcpyAsync(out_k, d_out_tmp, (size_t)rows * J * 8, cudaMemcpyDeviceToHost,
                                    h->stream));
      continue;
    }
    if (cf_samples) {
      CosSampleArgs sa;
      sa.greek_mask = greek_mask;  sa.plane = (size_t)rows * KK;  sa.rows = rows;
      sa.K = numU;  sa.Kp = KK / 2;  sa.r = r;
      sa.phi = d_phi + (size_t)p0 * M * numU * 2;
      sa.uk = d_tab;  sa.vk = d_tab + numU;  sa.C = d_C;
      const long long n = rows * sa.Kp;
      {
        ProfScope ps(h, "cos_coeff_samples_kernel");
        cos_coeff_samples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(sa);
      }
      FOP_LAUNCH_CHECK(h);
    } else if (use_i8) {
      // digit planes in place of the fp64 matrix (6 of its 8 bytes per entry): one set for price / delta /
      // gamma (phi itself), one for theta; then one contraction per requested kind
      unsigned char* planes[2] = {reinterpret_cast<unsigned char*>(d_C), nullptr};
      planes[1] = planes[0] + (q8_price_planes ? (size_t)chunk_rows * KK * Q8_SLICES : 0);
      unsigned char* flags[2] = {d_rowflag, d_rowflag + chunk_rows};
      FOP_CUDA(h, cudaMemsetAsync(d_rowflag, 0, 2 * (size_t)chunk_rows, h->stream));
      const bool both = q8_price_planes && q8_theta_planes;   // one evaluation feeds both sets of planes
      for (int set = 0; set < 2; ++set) {
        if (!(set == 0 ? q8_price_planes : q8_theta_planes) || (both && set == 1)) continue;
        CosCoeffArgs ca;
        ca.base = model.base;  ca.tc = model.time_change;
        ca.greek_mask = both ? 9 : set == 0 ? 1 : 8;
        ca.plane = both ? (size_t)chunk_rows * KK * Q8_SLICES : 0;   // bytes from the price to the theta planes
        ca.P = pc;  ca.M = M;  ca.K = numU;  ca.Kp = KK / 2;  ca.np = np;  ca.spb = 1;
        ca.r = r;
        ca.params = d_params + (size_t)p0 * np;
        ca.T = d_T;  ca.npow = d_npow;  ca.dt = cc.dt;  ca.uk = d_tab;  ca.vk = set == 0 ? d_v8p : d_v8t;
        ca.C = nullptr;  ca.C8 = planes[set];  ca.rowflag = flags[set];
        {
          ProfScope ps(h, "cos_coeff_kernel");
          if (!cos_coeff_launch_spec(h, ca, 32)) return fail(h, FOP_INVALID_ARG, "no quantising coefficient kernel");
        }
        FOP_LAUNCH_CHECK(h);
      }
      for (int ik = 0; ik < nk; ++ik) {
        const int kind = kinds[ik], g = kind >> 1, set = g == 3 ? 1 : 0, tb = g == 3 ? 0 : g;
        double* out_k = out ? out + (size_t)ik * P * M * J + (size_t)p0 * M * J : nullptr;
        CosGemmI8Args qa;
        qa.kind = kind;  qa.M = M;  qa.J = J;  qa.KK = KK;  qa.nblk = nblk8;  qa.rows = rows;  qa.r = r;  qa.S0 = S0;
        qa.amax = cc.q8_amax[tb] * (g == 3 ? theta_bound : 1.0);
        qa.T = d_T;  qa.strikes = d_K;
        qa.out = out ? (out_host ? d_out_tmp : out_k) : nullptr;
        qa.vec2 = (J % 2 == 0) && (reinterpret_cast<uintptr_t>(qa.out) % 16 == 0);
        qa.mkt = d_mkt;  qa.w = d_w;
        qa.rowloss = (loss && ik == 0) ? d_rowloss + (size_t)p0 * M * nblk : nullptr;
        qa.rowflag = flags[set];
        CUtensorMap tmA8;
        // [rows][6 planes][KK] bytes, addressed as (kk, row, plane)
        if (!make_tmap_3d_u8(&tmA8, planes[set], (unsigned long long)KK, (unsigned long long)rows, Q8_SLICES,
                             (unsigned long long)Q8_SLICES * KK, (unsigned long long)KK, Q8_KC, Q8_ROWS, 1))
          return fail(h, FOP_CUDA_ERROR, "tensor map (coefficient digit planes) could not be encoded");
        const long long ntiles = (rows + Q8_ROWS - 1) / Q8_ROWS;
        {
          ProfScope ps(h, "cos_gemm_kernel");
          // persistent, one CTA per SM: the SMs are divided among the strike blocks
          const dim3 qgrid((unsigned)std::min<long long>(ntiles, std::max(1, h->sm_count / nblk8)), (unsigned)nblk8);
          cos_gemm_i8_kernel<<<qgrid, Q8_THREADS, cos_i8_smem(M), h->stream>>>(qa, tmA8, tmB8[tb]);
        }
        FOP_LAUNCH_CHECK(h);
        if (out_host)
          FOP_CUDA(h, cudaMemcpyAsync(out_k, d_out_tmp, (size_t)rows * J * 8, cudaMemcpyDeviceToHost, h->stream));
      }
      continue;
    } else {
      CosCoeffArgs ca;
      ca.base = model.base;  ca.tc = model.time_change;
      ca.greek_mask = greek_mask;  ca.plane = (size_t)rows * KK;
      ca.P = pc;  ca.M = M;  ca.K = numU;  ca.Kp = KK / 2;  ca.np = np;
      ca.spb = std::max(1, std::min(COEFF_MAX_SPB, 1024 / ca.Kp));
      ca.r = r;
      ca.params = d_params + (size_t)p0 * np;
      ca.T = d_T;  ca.npow = d_npow;  ca.dt = cc.dt;  ca.uk = d_tab;  ca.vk = d_tab + numU;  ca.C = d_C;
      ca.C8 = nullptr;  ca.rowflag = nullptr;
      const int ngroups = (pc + ca.spb - 1) / ca.spb;
      const int cgrid = std::min(ngroups, h->sm_count * 32);
      {
        ProfScope ps(h, "cos_coeff_kernel");
        // default: model-specialised kernel (capi_cos_coeff.cu); FOP_COEFF_GENERIC=1: the generic one
        if (std::getenv("FOP_COEFF_GENERIC") || !cos_coeff_launch_spec(h, ca, 32)) {
          int ilp = M >= 2 ? 2 : 1;
          if (const char* e = std::getenv("FOP_COEFF_ILP")) ilp = std::max(1, std::min(atoi(e), M >= 2 ? 2 : 1));
          if (ilp >= 2) cos_coeff_kernel<2><<<cgrid, 256, 0, h->stream>>>(ca);
          else cos_coeff_kernel<1><<<cgrid, 256, 0, h->stream>>>(ca);
        }
      }
      FOP_LAUNCH_CHECK(h);
    }
    for (int ik = 0; ik < nk; ++ik) {
      const int kind = kinds[ik];
      double* out_k = out ? out + (size_t)ik * P * M * J + (size_t)p0 * M * J : nullptr;
      CosGemmArgs ga;
      ga.kind = kind;  ga.M = M;  ga.J = J;  ga.KK = KK;  ga.Jp = Jp;  ga.nblk = nblk;
      ga.rows = rows;  ga.r = r;  ga.S0 = S0;
      ga.C = d_C + (size_t)plane_of[kind >> 1] * rows * KK;
      ga.basis = d_basis;  ga.T = d_T;  ga.strikes = d_K;
      ga.out = out ? (out_host ? d_out_tmp : out_k) : nullptr;
      ga.vec2 = (J % 2 == 0) && (JB % 2 == 0) && (reinterpret_cast<uintptr_t>(ga.out) % 16 == 0);
      ga.mkt = d_mkt;  ga.w = d_w;
      ga.mvec2 = (J % 2 == 0) && (JB % 2 == 0) && (reinterpret_cast<uintptr_t>(d_mkt) % 16 == 0) &&
                 (reinterpret_cast<uintptr_t>(d_w) % 16 == 0);
      ga.rowloss = loss ? d_rowloss + (size_t)p0 * M * nblk : nullptr;
      ga.tab_smem = 0;
      const dim3 ggrid((unsigned)((rows + DM_RT - 1) / DM_RT), (unsigned)nblk);
      CUtensorMap tmA;
      const bool use_tma = gt && make_tmap_2d(&tmA, ga.C, (unsigned long long)KK, (unsigned long long)rows,
                                              (unsigned long long)KK, TM_KC, DM_RT, true);
      {
        ProfScope ps(h, "cos_gemm_kernel");
        if (use_tma) {  // persistent: two CTAs per SM (FOP_GEMM_TMA_CTAS=n: per SM, A/B runs)
          int per_sm = 2;
          if (const char* e = std::getenv("FOP_GEMM_TMA_CTAS")) per_sm = std::max(1, atoi(e));
          const dim3 pgrid((unsigned)std::min<long long>(ggrid.x, (long long)h->sm_count * per_sm), (unsigned)nblk);
          // market / weight tables in shared memory when two CTAs per SM still fit (227 KB - 1 KB
          // reserved per CTA, ~2.2 KB static)
          size_t smem = cos_tma_smem(g_nb);
          const size_t tabb = cos_tma_tab_bytes(g_nb, M);
          if (loss && d_mkt && 2 * (smem + tabb + 3400) <= h->smem_per_sm && !std::getenv("FOP_GEMM_NO_TAB")) {
            ga.tab_smem = 1;
            smem += tabb;
            FOP_CUDA(h, cudaFuncSetAttribute(gt->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          }
          // warp tiles of 16 rows (8 warps per CTA, 16 per SM) measure best: 2.19 vs 2.22 ms on config 5,
          // 33.9 vs 32.5 TFLOP/s on config 2 (FOP_GEMM_TMA_MB=4 | 1: 32- / 8-row tiles for A/B runs)
          auto tk = gt->kern2;
          int tthreads = 256;
          if (const char* e = std::getenv("FOP_GEMM_TMA_MB")) {
            if (atoi(e) == 4) { tk = gt->kern; tthreads = 128; }
            if (atoi(e) == 1 && gt->kern1) { tk = gt->kern1; tthreads = 512; }
          }
          FOP_CUDA(h, cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          tk<<<pgrid, tthreads, smem, h->stream>>>(ga, tmA, tmB);
        }
        else gemm_kernel<<<ggrid, 32 * (DM_RT / (8 * mb)), gemm_smem, h->stream>>>(ga);
      }
      FOP_LAUNCH_CHECK(h);
      if (out_host)
        FOP_CUDA(h, cudaMemcpyAsync(out_k, d_out_tmp, (size_t)rows * J * 8, cudaMemcpyDeviceToHost,
                                    h->stream));
    }
  }
  if (loss) {
    {
      ProfScope ps(h, "cos_loss_finalize_kernel");
      cos_loss_finalize2_kernel<<<(P + 255) / 256, 256, 0, h->stream>>>(d_rowloss, P, M, J, lossblk, d_loss);
    }
    FOP_LAUNCH_CHECK(h);
    if (loss_host)
      FOP_CUDA(h, cudaMemcpyAsync(loss, d_loss, (size_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (out_host || loss_host) FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  return FOP_OK;
}

}  // namespace

extern "C" {

int fop_cos(fop_handle_t h, fop_model_t model, const double* params, int P, const double* K_desc,
            int J, double S0, double r, const double* T, int M, int numU, int kind, double* out) {
  if (h && !out) return fop::fail(h, FOP_INVALID_ARG, "null output pointer");
  return cos_run(h, model, params, P, K_desc, J, S0, r, T, M, numU, &kind, 1, nullptr, nullptr, out,
                 nullptr);
}

int fop_cos_u_grid(const double* K_desc, int J, double S0, int numU, double* u_out) {
  if (!K_desc || !u_out || J < 2 || numU < 1) return FOP_INVALID_ARG;
  // u_k = k pi / (x_max - x_min), x = log(S0 / K)  (getXFromK OptionPricing.h:110-117; the grid of
  // fangoost::computeExpectationVectorLevy)
  const double xMin = std::log(S0 / K_desc[0]), xMax = std::log(S0 / K_desc[J - 1]);
  const double du = M_PI / (xMax - xMin);
  for (int k = 0; k < numU; ++k) u_out[k] = du * k;
  return FOP_OK;
}

int fop_cos_cf_samples(fop_handle_t h, const double* cf_samples, int P, const double* K_desc, int J,
                       double S0, double r, const double* T, int M, int numU, int kind, double* out) {
  if (h && (!out || !cf_samples)) return fop::fail(h, FOP_INVALID_ARG, "null pointer");
  const fop_model_t none = {0, 0};
  return cos_run(h, none, nullptr, P, K_desc, J, S0, r, T, M, numU, &kind, 1, nullptr, nullptr, out,
                 nullptr, cf_samples);
}

int fop_cos_greeks(fop_handle_t h, fop_model_t model, const double* params, int P,
                   const double* K_desc, int J, double S0, double r, const double* T, int M,
                   int numU, const int* kinds, int nkinds, double* out) {
  if (h && !out) return fop::fail(h, FOP_INVALID_ARG, "null output pointer");
  return cos_run(h, model, params, P, K_desc, J, S0, r, T, M, numU, kinds, nkinds, nullptr, nullptr,
                 out, nullptr);
}

int fop_cos_loss(fop_handle_t h, fop_model_t model, const double* params, int P,
                 const double* K_desc, int J, double S0, double r, const double* T, int M,
                 int numU, const double* mkt, const double* w, double* loss, double* prices_out) {
  if (h && (!loss || !mkt)) return fop::fail(h, FOP_INVALID_ARG, "null loss / market pointer");
  const int kind = FOP_CALL_PRICE;
  return cos_run(h, model, params, P, K_desc, J, S0, r, T, M, numU, &kind, 1, mkt, w, prices_out,
                 loss);
}

}  // extern "C"
