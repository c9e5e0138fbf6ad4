// Launcher of the fused warp-specialised COS kernel (cos_fused_ws.cuh); called from capi_cos.cu.
// Kept in its own translation unit so that its six instantiations (three strike-block widths x one or two maturities per thread) compile in parallel with the
// split path.
#include <algorithm>
#include <cstdlib>

#include "cos_fused_ws.cuh"
#include "handle.cuh"

namespace fop {

// strike-block widths instantiated for the fused path (JB = 8 NB)
int cos_fused_pick_nb(int J) { return J <= 16 ? 2 : J <= 40 ? 5 : J <= 72 ? 9 : 0; }

int cos_fused_consts(fop_handle_t h, int base, int tc, const double* params, int np, int P,
                     ModelConsts* out) {
  ProfScope ps(h, "cos_consts_kernel");
  cos_consts_kernel<<<(P + 127) / 128, 128, 0, h->stream>>>(base, tc, params, np, P, out);
  FOP_LAUNCH_CHECK(h);
  return FOP_OK;
}

template <int NB>
static int launch_nb(fop_handle_t h, CosFusedArgs& a, int ilp) {
  auto kern = ilp >= 2 ? cos_fused_ws_kernel<NB, 2> : cos_fused_ws_kernel<NB, 1>;
  a.depth = fw_depth(NB, h->smem_optin);
  if (a.depth < 3) return fail(h, FOP_CUDA_ERROR, "fused COS: shared memory too small (%zu B)", h->smem_optin);
  const size_t smem = (size_t)a.depth * fw_stage_bytes(NB) + fw_tail_bytes(NB);
  FOP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = ((long long)a.ca.P + a.SPT - 1) / a.SPT;
  const int grid = (int)std::min<long long>(ntiles, h->sm_count);
  {
    ProfScope ps(h, "cos_fused_ws_kernel");
    kern<<<grid, FW_THREADS, smem, h->stream>>>(a);
  }
  FOP_LAUNCH_CHECK(h);
  return FOP_OK;
}

int cos_fused_launch(fop_handle_t h, CosFusedArgs& a, int NB) {
  int ilp = a.ca.M >= 2 ? 2 : 1;  // measured: 2 maturities interleaved beat 1 and 4 (register cap 128)
  if (const char* e = std::getenv("FOP_COEFF_ILP")) ilp = std::max(1, std::min(atoi(e), a.ca.M >= 2 ? 2 : 1));
  switch (NB) {
    case 2: return launch_nb<2>(h, a, ilp);
    case 5: return launch_nb<5>(h, a, ilp);
    case 9: return launch_nb<9>(h, a, ilp);
  }
  return fail(h, FOP_INVALID_ARG, "fused COS: no kernel for NB=%d", NB);
}

}  // namespace fop
