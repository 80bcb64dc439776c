// cosmolike.cu — C ABI of libcosmolike_b200.so (see include/cosmolike.h).
// Host side: context, device copies of the static operands, W = L^-1 in extended precision, TMA descriptors,
// pinned staging and the three-stage launch sequence.  No torch, no Python; plain CUDA runtime + one driver
// entry point (cuTensorMapEncodeTiled) resolved at run time.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <set>
#include <tuple>
#include <atomic>
#include <cmath>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cosmolike.h"
#include "chi2_gemm.cuh"
#include "chi2_ozaki.cuh"
#include "devspec.h"
#include "friedmann.cuh"
#include "multi.cuh"

using namespace cosmolike;

static thread_local std::string g_create_error;

struct cl_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  static constexpr int kRing = 64;      // timing history: one event set per evaluation call
  cudaEvent_t evring[kRing][7] = {};   // [6]: between forming the digit planes and the contraction (stage 3 split)
  cudaEvent_t* ev = evring[0];
  int64_t n_timed = 0;
  DevSpec ds{};
  std::vector<void*> dev_allocs;
  // stage 3 operands
  double* d_W = nullptr;  // [n_sn][ldW] = L^-1
  double* d_u = nullptr;  // [n_sn] u = W 1
  double uu = 0.0;        // u.u
  int64_t ldW = 0;
  int T = 0;
  CUtensorMap tmW{};
  // workspace
  int64_t cap_rows = 0, max_rows = 65536;
  int64_t ldR = 0;
  double *d_theta = nullptr, *d_out = nullptr, *d_R = nullptr, *d_aux = nullptr, *d_part = nullptr, *d_part_u = nullptr;
  double *d_scratch = nullptr;  // helper outputs
  int64_t scratch_bytes = 0;
  double *h_theta = nullptr, *h_out = nullptr;  // pinned
  int64_t h_theta_cap = 0, h_out_cap = 0;
  int64_t launches = 0;
  int opt_gemm_ctas = 0, opt_s12_ctas = 0, opt_diag_skip = 1, opt_dbg = 0, opt_gemm_dynamic = 1, opt_group_rb = 0, opt_s12_lean = 1, opt_fuse_planes = 1;   // stage 2 writes the digit planes itself where it can
  int* d_counter = nullptr;
  // stage 3 on tcgen05 (chi2_ozaki.cuh): int8 digit planes of W (static) and of the residual rows (per pass)
  int opt_engine = CL_CHI2_ENGINE_TCGEN05, opt_slices = 7, opt_slice_tpb = 128;
  int oz_slices_built = 0;           // S the W planes were built for (0 = none)
  int oz_T = 0;                      // column tiles of the sliced kernel
  int64_t oz_ld = 0;                 // bytes per row of a digit plane
  int8_t *d_Ws = nullptr, *d_Rs = nullptr;
  double *d_wscale = nullptr, *d_rscale = nullptr;
  int64_t oz_cap_rows = 0;
  CUtensorMap tmWs{};
  // accuracy guard of the tcgen05 engine (friedmann.cuh: GuardArgs): static part of the a-priori bound and the fallback state
  double oz_omega = 0.0;             // worst case: sqrt(sum_n (nnz_n 2^eW_n)^2) over the rows of W
  double oz_omega_pr = 0.0;          // probabilistic: max_n sqrt(nnz_n) 2^eW_n
  int opt_guard = 1, opt_guard_mode = 0;   // mode 0: probabilistic bound (lambda = 8), 1: worst case
  double guard_abs = 5e-7, guard_rel = 1e-12;
  int* d_guard = nullptr;            // [0] rows flagged since creation, [1] in the current pass, [2 + rb] row-block marks
  unsigned char* d_rowflag = nullptr;
  double *d_part_fb = nullptr, *d_part_u_fb = nullptr;   // [T][cap] partials of the FP64 fallback pass
  // multi-GPU (multi.cuh): NCCL communicator of this context and the gather / grid buffers
  NcclComm comm = nullptr;
  int comm_rank = 0, comm_size = 0;
  double* d_gather = nullptr;          // [comm_size][gather_cap] gathered results (host-memory variant)
  int64_t gather_cap = 0;
  double *d_grid_part = nullptr, *h_grid_part = nullptr;   // per-block partial reductions of a grid chunk
  int64_t grid_part_cap = 0;
  // device-side proposals (cl_propose_eval)
  double *d_prop_u = nullptr, *d_prop_val = nullptr, *d_prop_keep = nullptr;   // [cap][ndim], [cap], kept rows (u | theta | value)
  unsigned char* d_prop_inside = nullptr;
  int* d_prop_cnt = nullptr;             // [0] accepted total, [1] inside total, [2 ..] per-block counts / first slots
  int64_t prop_cap = 0, prop_keep_cap = 0;
  // CUDA graphs of small evaluations: one instantiated graph per (entry point, rows, selector, buffers); a call shape is
  // captured the second time it is seen (the first call settles every allocation), and every change of an option or of a
  // workspace pointer starts a new epoch, which drops the graphs
  typedef std::tuple<int, int64_t, int, int, const void*, const void*, const void*, int64_t> GraphKey;
  struct GraphEntry { cudaGraphExec_t exec; int64_t launches; };
  std::map<GraphKey, GraphEntry> graphs;
  std::set<GraphKey> graph_seen;
  int opt_graph = 1;
  int64_t graph_max_rows = 4096, graph_replays = 0;
  std::string err, desc;
  std::mutex mu;
};

static void drop_graphs(cl_ctx* c) {
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second.exec);
  c->graphs.clear();
  c->graph_seen.clear();
}

static int fail(cl_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                            \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) return fail(ctx, CL_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 2-D f64 tensor [rows][cols] with row stride ld (elements), box = [kBM rows][kBK cols], 128-byte swizzle
static int make_tmap(cl_ctx* c, CUtensorMap* tm, const double* ptr, int64_t rows, int64_t cols, int64_t ld) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(c, CL_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, CL_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CL_OK;
}

// 3-D int8 tensor [slices][rows][ld bytes] with k extent `cols`, box = {64 B of k, box_rows, box_slices}, 64-byte swizzle
static int make_tmap_planes(cl_ctx* c, CUtensorMap* tm, const int8_t* ptr, int64_t cols, int64_t rows, int slices, int64_t ld, int box_rows, int box_slices) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(c, CL_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slices};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)kOzKB, (cuuint32_t)box_rows, (cuuint32_t)box_slices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, CL_E_CUDA, "cuTensorMapEncodeTiled (digit planes) failed with CUresult %d", (int)r);
  return CL_OK;
}

template <typename T>
static int upload(cl_ctx* c, const T* src, size_t n, const T** dst) {
  *dst = nullptr;
  if (!src || n == 0) return CL_OK;
  void* p = nullptr;
  CUDA_TRY(c, cudaMalloc(&p, n * sizeof(T)));
  c->dev_allocs.push_back(p);
  CUDA_TRY(c, cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
  *dst = (const T*)p;
  return CL_OK;
}

// ---- W = L^-1 (lower triangular, row-major) in long double, column blocks in parallel ----
static bool invert_lower(const double* L, int n, int64_t ldL, std::vector<double>& W, int64_t ldW) {
  for (int i = 0; i < n; i++)
    if (!(L[(size_t)i * ldL + i] > 0.0) || !std::isfinite(L[(size_t)i * ldL + i])) return false;
  W.assign((size_t)n * ldW, 0.0);
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  int nthreads = (int)std::min<unsigned>(hw, 32u);
  if (n < 256) nthreads = 1;
  const int cb = 32;  // column block
  std::vector<std::thread> pool;
  std::atomic<int> next{0};
  auto work = [&]() {
    std::vector<long double> X((size_t)n * cb);
    for (;;) {
      int blk = next.fetch_add(1);
      int j0 = blk * cb;
      if (j0 >= n) break;
      int jw = std::min(cb, n - j0);
      // solve L X = E[:, j0:j0+jw]; rows < j0 of X are zero
      for (int i = j0; i < n; i++) {
        long double* xi = &X[(size_t)i * cb];
        for (int q = 0; q < jw; q++) xi[q] = (i == j0 + q) ? 1.0L : 0.0L;
        const double* Li = L + (size_t)i * ldL;
        for (int k = j0; k < i; k++) {
          long double lik = Li[k];
          const long double* xk = &X[(size_t)k * cb];
          for (int q = 0; q < jw; q++) xi[q] -= lik * xk[q];
        }
        long double inv = 1.0L / (long double)Li[i];
        for (int q = 0; q < jw; q++) xi[q] *= inv;
        for (int q = 0; q < jw; q++) W[(size_t)i * ldW + j0 + q] = (j0 + q <= i) ? (double)xi[q] : 0.0;
      }
    }
  };
  for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return true;
}

// Cholesky of a symmetric positive definite matrix given by its INVERSE (CL_SN_INVCOV with a large block):
// C = inv(Cinv) by Gauss-Jordan-free route: factor Cinv = G G^T, then C = G^-T G^-1 and chol(C) is formed anew.
static bool cholesky_lower_ld(std::vector<long double>& A, int n) {
  for (int j = 0; j < n; j++) {
    long double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0L)) return false;
    d = sqrtl(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      long double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
    for (int i = 0; i < j; i++) A[(size_t)i * n + j] = 0.0L;
  }
  return true;
}

static bool lower_factor_from_invcov(const double* Cinv, int n, std::vector<double>& L) {
  // Cinv = G G^T ; C = G^-T G^-1 ; L = chol(C)
  std::vector<long double> G((size_t)n * n);
  for (size_t i = 0; i < (size_t)n * n; i++) G[i] = Cinv[i];
  if (!cholesky_lower_ld(G, n)) return false;
  std::vector<double> Gd((size_t)n * n), Ginv;
  for (size_t i = 0; i < (size_t)n * n; i++) Gd[i] = (double)G[i];
  if (!invert_lower(Gd.data(), n, n, Ginv, n)) return false;
  std::vector<long double> Cm((size_t)n * n, 0.0L);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      long double s = 0.0L;
      for (int k = i; k < n; k++) s += (long double)Ginv[(size_t)k * n + i] * Ginv[(size_t)k * n + j];
      Cm[(size_t)i * n + j] = s;
      Cm[(size_t)j * n + i] = s;
    }
  if (!cholesky_lower_ld(Cm, n)) return false;
  L.resize((size_t)n * n);
  for (size_t i = 0; i < (size_t)n * n; i++) L[i] = (double)Cm[i];
  return true;
}

// ---- kernel dispatch over the (family, dark-energy) instantiations ----
typedef void (*S12Kernel)(const DevSpec, const Stage12Args);
// lean: 0 = the full kernel, 1 = large SN block alone (late-time family), 2 = large SN block on the fast path with fused digit
// planes + small probes (friedmann.cuh)
static S12Kernel pick_s12(int fam, int de, int lean = 0) {
#define CASE(F, D) if (fam == F && de == D) return k_friedmann_residuals<F, D, 0>;
#define LEAN_CASE(D) if (lean == 1 && fam == CL_FAMILY_LATE && de == D) return k_friedmann_residuals<CL_FAMILY_LATE, D, 1>;
#define MID_CASE(F, D) if (lean == 2 && fam == F && de == D) return k_friedmann_residuals<F, D, 2>;
  LEAN_CASE(CL_DE_LCDM) LEAN_CASE(CL_DE_WCDM) LEAN_CASE(CL_DE_CPL) LEAN_CASE(CL_DE_THAWING)
  MID_CASE(CL_FAMILY_LATE, CL_DE_LCDM) MID_CASE(CL_FAMILY_LATE, CL_DE_WCDM) MID_CASE(CL_FAMILY_LATE, CL_DE_CPL) MID_CASE(CL_FAMILY_LATE, CL_DE_THAWING)
  MID_CASE(CL_FAMILY_FULL, CL_DE_LCDM) MID_CASE(CL_FAMILY_FULL, CL_DE_WCDM) MID_CASE(CL_FAMILY_FULL, CL_DE_CPL) MID_CASE(CL_FAMILY_FULL, CL_DE_THAWING)
#undef MID_CASE
  CASE(CL_FAMILY_LATE, CL_DE_LCDM) CASE(CL_FAMILY_LATE, CL_DE_WCDM) CASE(CL_FAMILY_LATE, CL_DE_CPL) CASE(CL_FAMILY_LATE, CL_DE_THAWING)
  CASE(CL_FAMILY_FULL, CL_DE_LCDM) CASE(CL_FAMILY_FULL, CL_DE_WCDM) CASE(CL_FAMILY_FULL, CL_DE_CPL) CASE(CL_FAMILY_FULL, CL_DE_THAWING)
#undef CASE
#undef LEAN_CASE
  return nullptr;
}
// the lean instantiation serves plain evaluations of a large SN block alone (see friedmann.cuh)
static bool s12_lean(const DevSpec& d, int mode) {
  return mode == MODE_EVAL && d.family == CL_FAMILY_LATE && d.n_sn > 0 && !d.sn_small && d.n_bao == 0 && d.n_cc == 0 && d.cmb_mode == CL_CMB_NONE;
}
// the fast SN path of stage 2 (four consecutive supernovae per thread from the quad-interleaved static copies) with the digit
// planes written in the same kernel: a thread must be able to hold its share of the row
static bool s12_fusable(const DevSpec& d) {
  return d.n_sn > 0 && !d.sn_small && d.sn_zs4 != nullptr && ((d.n_sn + 127) & ~127) <= 8 * kS12Threads && d.grid_uniform &&
         (d.n_vel == 0 || d.vel_pm1) && d.sn_mu_fixed == nullptr && d.n_lin == 0;
}

static int validate(const cl_spec* s) {
  if (!s) return fail(nullptr, CL_E_INVALID, "spec is NULL");
  if (s->abi_version != CL_ABI_VERSION) return fail(nullptr, CL_E_INVALID, "abi_version %u != %u", s->abi_version, CL_ABI_VERSION);
  if (s->ndim < 1 || s->ndim > CL_MAX_DIM) return fail(nullptr, CL_E_INVALID, "ndim out of range");
  auto col_ok = [&](int c, bool required) { return required ? (c >= 0 && c < s->ndim) : (c >= -1 && c < s->ndim); };
  if (s->family != CL_FAMILY_LATE && s->family != CL_FAMILY_FULL) return fail(nullptr, CL_E_INVALID, "bad family");
  if (s->de_model < CL_DE_LCDM || s->de_model > CL_DE_THAWING) return fail(nullptr, CL_E_INVALID, "bad de_model");
  if (!col_ok(s->col_H0, false)) return fail(nullptr, CL_E_INVALID, "col_H0 out of range");
  if (s->family == CL_FAMILY_LATE && !col_ok(s->col_Om, true)) return fail(nullptr, CL_E_INVALID, "LATE family needs col_Om");
  if (s->family == CL_FAMILY_FULL && (!col_ok(s->col_obh2, true) || !col_ok(s->col_och2, true)))
    return fail(nullptr, CL_E_INVALID, "FULL family needs col_obh2 and col_och2");
  if (s->de_model != CL_DE_LCDM && !col_ok(s->col_w0, true)) return fail(nullptr, CL_E_INVALID, "dark-energy model needs col_w0");
  if (s->de_model == CL_DE_CPL && !col_ok(s->col_wa, true)) return fail(nullptr, CL_E_INVALID, "CPL needs col_wa");
  const bool need_grid = s->n_sn > 0 || s->n_bao > 0;
  if (need_grid || s->z_grid) {
    if (!s->z_grid || s->n_grid < 16 || s->n_grid > kS12Threads * kPPT) return fail(nullptr, CL_E_INVALID, "z_grid must have 16..4096 points");
    for (int i = 1; i < s->n_grid; i++)
      if (!(s->z_grid[i] > s->z_grid[i - 1])) return fail(nullptr, CL_E_INVALID, "z_grid must be strictly increasing");
  }
  if (s->n_sn < 0 || s->n_bao < 0 || s->n_bao > CL_MAX_BAO || s->n_cc < 0 || s->n_cc > CL_MAX_CC) return fail(nullptr, CL_E_INVALID, "block size out of range");
  if (s->n_sn > 0) {
    if (!s->sn_zcmb || !s->sn_zhel || !s->sn_obs || !s->sn_mat) return fail(nullptr, CL_E_INVALID, "SN block has NULL arrays");
    if (!col_ok(s->col_offset, false)) return fail(nullptr, CL_E_INVALID, "col_offset out of range");
    if (s->n_vel < 0 || s->n_vel > CL_MAX_VEL || (s->n_vel > 0 && !s->sn_vel_weight)) return fail(nullptr, CL_E_INVALID, "bad velocity templates");
    for (int k = 0; k < s->n_vel; k++) if (!col_ok(s->col_vel[k], true)) return fail(nullptr, CL_E_INVALID, "col_vel out of range");
    if (s->n_lin < 0 || s->n_lin > CL_MAX_VEL || (s->n_lin > 0 && !s->sn_lin_template)) return fail(nullptr, CL_E_INVALID, "bad linear templates");
    for (int k = 0; k < s->n_lin; k++) if (!col_ok(s->col_lin[k], true)) return fail(nullptr, CL_E_INVALID, "col_lin out of range");
  }
  if (s->n_bao > 0) {
    if (!s->bao_z || !s->bao_value || !s->bao_qty || !s->bao_inv_cov) return fail(nullptr, CL_E_INVALID, "BAO block has NULL arrays");
    if (s->rd_mode == CL_RD_PARAM && !col_ok(s->col_rd, true)) return fail(nullptr, CL_E_INVALID, "col_rd out of range");
    if (s->rd_mode == CL_RD_FIT && s->family == CL_FAMILY_LATE && !col_ok(s->col_obh2, true)) return fail(nullptr, CL_E_INVALID, "r_drag fit needs col_obh2");
    for (int k = 0; k < s->n_bao; k++) if (s->bao_qty[k] < 0 || s->bao_qty[k] > 3) return fail(nullptr, CL_E_INVALID, "bad BAO quantity code");
  }
  if (s->cmb_mode != CL_CMB_NONE && s->family != CL_FAMILY_FULL) return fail(nullptr, CL_E_INVALID, "CMB block needs the FULL family");
  if (s->n_gl < 0 || s->n_gl > CL_MAX_GL) return fail(nullptr, CL_E_INVALID, "n_gl out of range");
  if (s->n_cc > 0 && (!s->cc_z || !s->cc_H || !s->cc_inv_cov || !col_ok(s->col_fcc, false))) return fail(nullptr, CL_E_INVALID, "bad CC block");
  if (s->n_gauss_chi2 < 0 || s->n_gauss_chi2 > CL_MAX_GAUSS || s->n_gauss_prior < 0 || s->n_gauss_prior > CL_MAX_GAUSS)
    return fail(nullptr, CL_E_INVALID, "too many Gaussian terms");
  for (int g = 0; g < s->n_gauss_chi2; g++) if (!col_ok(s->gauss_chi2_col[g], true)) return fail(nullptr, CL_E_INVALID, "gauss_chi2_col out of range");
  for (int g = 0; g < s->n_gauss_prior; g++) if (!col_ok(s->gauss_prior_col[g], true)) return fail(nullptr, CL_E_INVALID, "gauss_prior_col out of range");
  return CL_OK;
}

static void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.resize(n); w.resize(n);
  for (int i = 0; i < n; i++) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; it++) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[n - 1 - i] = z;
    w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

extern "C" int cl_destroy(cl_ctx* c) {
  if (!c) return CL_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (void* p : c->dev_allocs) cudaFree(p);
  for (double* p : {c->d_theta, c->d_out, c->d_R, c->d_aux, c->d_part, c->d_part_u, c->d_scratch, c->d_W, c->d_u}) if (p) cudaFree(p);
  if (c->d_counter) cudaFree(c->d_counter);
  for (void* p : {(void*)c->d_guard, (void*)c->d_rowflag, (void*)c->d_part_fb, (void*)c->d_part_u_fb}) if (p) cudaFree(p);
  for (void* p : {(void*)c->d_Ws, (void*)c->d_Rs, (void*)c->d_wscale, (void*)c->d_rscale}) if (p) cudaFree(p);
  drop_graphs(c);
  if (c->comm) { nccl_api().CommDestroy(c->comm); c->comm = nullptr; }
  if (c->d_gather) cudaFree(c->d_gather);
  for (void* p : {(void*)c->d_prop_u, (void*)c->d_prop_val, (void*)c->d_prop_keep, (void*)c->d_prop_inside, (void*)c->d_prop_cnt}) if (p) cudaFree(p);
  if (c->d_grid_part) cudaFree(c->d_grid_part);
  if (c->h_grid_part) cudaFreeHost(c->h_grid_part);
  if (c->h_theta) cudaFreeHost(c->h_theta);
  if (c->h_out) cudaFreeHost(c->h_out);
  for (auto& r : c->evring) for (auto& e : r) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return CL_OK;
}

extern "C" const char* cl_last_error(const cl_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
extern "C" const char* cl_describe(const cl_ctx* c) { return c ? c->desc.c_str() : "cosmolike_b200 (no context)"; }
extern "C" int64_t cl_launch_count(const cl_ctx* c) { return c ? c->launches : 0; }

extern "C" int cl_create(const cl_spec* spec, int device, cl_ctx** out) {
  if (!out) return fail(nullptr, CL_E_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate(spec);
  if (rc != CL_OK) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, CL_E_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(nullptr, CL_E_NO_DEVICE, "device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, CL_E_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(nullptr, CL_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

  cl_ctx* c = new cl_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  auto bail = [&](int code) { g_create_error = c->err; cl_destroy(c); return code; };
#define TRY(expr) do { int r_ = (expr); if (r_ != CL_OK) return bail(r_); } while (0)
#define CTRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { fail(c, CL_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); return bail(CL_E_CUDA); } } while (0)
  CTRY(cudaSetDevice(device));
  CTRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& r : c->evring) for (auto& e : r) CTRY(cudaEventCreate(&e));

  DevSpec& d = c->ds;
  const cl_spec& s = *spec;
  d.ndim = s.ndim; d.family = s.family; d.de_model = s.de_model;
  d.col_H0 = s.col_H0; d.col_Om = s.col_Om; d.Om_is_physical = s.Om_is_physical; d.col_obh2 = s.col_obh2; d.col_och2 = s.col_och2;
  d.col_w0 = s.de_model == CL_DE_LCDM ? -1 : s.col_w0; d.col_wa = s.de_model == CL_DE_CPL ? s.col_wa : -1;
  d.H0_fixed = s.H0_fixed; d.H0_scale = s.H0_scale; d.k = s.cmbc;
  d.nu_inv_rho0 = s.cmbc.nu_rho0 != 0.0 ? 1.0 / s.cmbc.nu_rho0 : 0.0;
  // grid
  d.G = s.z_grid ? s.n_grid : 0;
  if (s.z_grid) {
    TRY(upload(c, s.z_grid, (size_t)s.n_grid, &d.z_grid));
    // np.linspace(0, stop, n): y = arange(n) * step with step = stop/(n-1), last element := stop
    double step = s.z_grid[1];
    bool uni = s.z_grid[0] == 0.0 && step > 0.0;
    for (int i = 0; uni && i < s.n_grid - 1; i++) uni = (s.z_grid[i] == (double)i * step);
    d.grid_uniform = uni ? 1 : 0;
    d.step = step; d.inv_step = 1.0 / step; d.z_last = s.z_grid[s.n_grid - 1];
    if (uni) {
      // theta-independent node tables of the grid pass, transposed: entry [k][t] belongs to node 16 t + k (friedmann.cuh)
      const int nt = kS12Threads, nk = kPPT + 1;
      std::vector<double> ln1pz((size_t)nk * nt);
      for (int t = 0; t < nt; t++)
        for (int k = 0; k < nk; k++) ln1pz[(size_t)k * nt + t] = (double)log1pl((long double)(kPPT * t + k) * (long double)step);
      TRY(upload(c, ln1pz.data(), ln1pz.size(), &d.grid_ln1pz_T));
      if (s.family == CL_FAMILY_FULL) {
        // Omnu_z(z_i) with the reference's formula (cmb/data_planck_act_compression.py:53-66); independent of theta
        std::vector<double> om((size_t)nk * nt);
        const cl_cmb_consts& k = s.cmbc;
        for (int t = 0; t < nt; t++)
          for (int kk = 0; kk < nk; kk++) {
            double zp1 = 1.0 + (double)(kPPT * t + kk) * step, r = k.nu_m0 / zp1, mz = r * r, ws = 0.0;
            for (int q = 0; q < 5; q++) ws += sqrt(k.nu_q2[q] + mz) * k.nu_w[q];
            om[(size_t)kk * nt + t] = zp1 * zp1 * zp1 * zp1 * ws / k.nu_rho0;
          }
        TRY(upload(c, om.data(), om.size(), &d.grid_omnu_T));
      }
    }
  }
  {  // fast_log10 table: c_j = 1 + (j + 1/2)/128; entry = {fl(1/c_j), -log10(fl(1/c_j))}
    std::vector<double> tab(256);
    for (int j = 0; j < 128; j++) {
      double inv = (double)(1.0L / (1.0L + ((long double)j + 0.5L) / 128.0L));
      tab[2 * j] = inv;
      tab[2 * j + 1] = (double)(-5.0L * log10l((long double)inv));
    }
    const double* dtab = nullptr;
    TRY(upload(c, tab.data(), tab.size(), &dtab));
    d.logtab = reinterpret_cast<const double2*>(dtab);
  }
  // SN
  d.n_sn = s.n_sn;
  if (s.n_sn > 0) {
    const int n = s.n_sn;
    d.sn_small = n <= CL_SN_SMALL_MAX ? 1 : 0;
    d.sn_form = s.sn_cov_form; d.col_offset = s.col_offset; d.n_vel = s.n_vel; d.vel_mode = s.vel_mode; d.vel_scale = s.vel_scale;
    for (int k = 0; k < CL_MAX_VEL; k++) d.col_vel[k] = k < s.n_vel ? s.col_vel[k] : 0;
    // vel_pm1: one template whose weights are all +-1 (the z_turn step) and the divide form -> two reciprocals per theta
    bool pm1 = s.n_vel == 1 && s.vel_mode == CL_VEL_DIVIDE;
    for (int i = 0; pm1 && i < n; i++) pm1 = (s.sn_vel_weight[i] == 1.0 || s.sn_vel_weight[i] == -1.0);
    d.vel_pm1 = pm1 ? 1 : 0;
    std::vector<double> pack((size_t)n * 4);
    for (int i = 0; i < n; i++) {
      pack[4 * i + 0] = s.sn_zcmb[i];
      pack[4 * i + 1] = s.n_vel > 0 ? s.sn_vel_weight[i] : 0.0;
      pack[4 * i + 2] = 1.0 + s.sn_zhel[i];  // (1.0 + z_hel), sn/pantheon.py:54
      pack[4 * i + 3] = s.sn_obs[i];
    }
    TRY(upload(c, pack.data(), pack.size(), &d.sn_pack));
    {
      std::vector<double> zs((size_t)n * 2), obsp(n);
      for (int i = 0; i < n; i++) {
        zs[2 * i] = pm1 ? 1.0 + s.sn_zcmb[i] : s.sn_zcmb[i];
        zs[2 * i + 1] = pm1 ? s.sn_vel_weight[i] : 0.0;
        obsp[i] = (double)((long double)s.sn_obs[i] - 25.0L - 5.0L * log10l(1.0L + (long double)s.sn_zhel[i]));
      }
      const double* dz = nullptr;
      TRY(upload(c, zs.data(), zs.size(), &dz));
      d.sn_zs = reinterpret_cast<const double2*>(dz);
      TRY(upload(c, obsp.data(), obsp.size(), &d.sn_obsp));
      // quad-interleaved copies for the fused digit-plane path (friedmann.cuh): [q][m] = supernova 4 m + q, padded with the last one
      const int q4 = ((n + 127) & ~127) / 4;   // covers the padded plane pitch (a multiple of 128 columns)
      {
        std::vector<double> zs4((size_t)4 * q4 * 2), ob4((size_t)4 * q4);
        for (int q = 0; q < 4; q++)
          for (int m = 0; m < q4; m++) {
            const int i = std::min(4 * m + q, n - 1);
            zs4[2 * ((size_t)q * q4 + m)] = zs[2 * i]; zs4[2 * ((size_t)q * q4 + m) + 1] = zs[2 * i + 1];
            ob4[(size_t)q * q4 + m] = obsp[i];
          }
        const double* dz4 = nullptr;
        TRY(upload(c, zs4.data(), zs4.size(), &dz4));
        d.sn_zs4 = reinterpret_cast<const double2*>(dz4);
        TRY(upload(c, ob4.data(), ob4.size(), &d.sn_obsp4));
        d.sn_q4 = q4;
      }
    }
    if (s.n_vel > 0) TRY(upload(c, s.sn_vel_weight, (size_t)n * s.n_vel, &d.sn_vel_w));
    if (s.sn_mu_fixed) TRY(upload(c, s.sn_mu_fixed, (size_t)n, &d.sn_mu_fixed));
    d.n_lin = s.n_lin;
    for (int k = 0; k < CL_MAX_VEL; k++) d.col_lin[k] = k < s.n_lin ? s.col_lin[k] : 0;
    if (s.n_lin > 0) TRY(upload(c, s.sn_lin_template, (size_t)n * s.n_lin, &d.sn_lin_t));
    // factor handling
    std::vector<double> Lbuf;
    const double* L = s.sn_mat;
    if (s.sn_cov_form == CL_SN_INVCOV && !d.sn_small) {
      if (!lower_factor_from_invcov(s.sn_mat, n, Lbuf)) { fail(c, CL_E_NUMERIC, "inverse covariance is not positive definite"); return bail(CL_E_NUMERIC); }
      L = Lbuf.data();
    }
    if (d.sn_small && s.sn_cov_form == CL_SN_INVCOV) {
      TRY(upload(c, s.sn_mat, (size_t)n * n, &d.sn_mat_small));
    } else {
      c->ldW = ((int64_t)n + 1) & ~1LL;  // 16-byte row stride for TMA
      std::vector<double> W;
      if (!invert_lower(L, n, n, W, c->ldW)) { fail(c, CL_E_NUMERIC, "Cholesky factor has a non-positive diagonal"); return bail(CL_E_NUMERIC); }
      std::vector<double> u(n);
      long double uu = 0.0L;
      for (int i = 0; i < n; i++) {
        long double acc = 0.0L;
        for (int j = 0; j <= i; j++) acc += W[(size_t)i * c->ldW + j];
        u[i] = (double)acc; uu += acc * acc;
      }
      c->uu = (double)uu;
      {  // static part of the digit-plane error bound: every row of W with its power-of-two scale (as k_oz_slice_rows forms it)
        long double om2 = 0.0L, ompr = 0.0L;
        for (int i = 0; i < n; i++) {
          double mx = 0.0;
          int nnz = 0;
          for (int j = 0; j <= i; j++) { const double a = fabs(W[(size_t)i * c->ldW + j]); mx = std::max(mx, a); nnz += a != 0.0; }
          const int e = std::max(mx > 0.0 ? ilogb(mx) + 1 : 0, -900);
          const long double t = (long double)nnz * ldexpl(1.0L, e);
          om2 += t * t;
          ompr = std::max(ompr, sqrtl((long double)nnz) * ldexpl(1.0L, e));
        }
        c->oz_omega = (double)sqrtl(om2);
        c->oz_omega_pr = (double)ompr;
      }
      if (d.sn_small) {
        std::vector<double> Wc((size_t)n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Wc[(size_t)i * n + j] = W[(size_t)i * c->ldW + j];
        TRY(upload(c, Wc.data(), Wc.size(), &d.sn_mat_small));
      } else {
        CTRY(cudaMalloc(&c->d_W, W.size() * sizeof(double)));
        CTRY(cudaMemcpy(c->d_W, W.data(), W.size() * sizeof(double), cudaMemcpyHostToDevice));
        CTRY(cudaMalloc(&c->d_counter, sizeof(int)));
        CTRY(cudaMalloc(&c->d_u, n * sizeof(double)));
        CTRY(cudaMemcpy(c->d_u, u.data(), n * sizeof(double), cudaMemcpyHostToDevice));
        c->T = (n + kBN - 1) / kBN;
        c->ldR = ((int64_t)n + 15) & ~15LL;
        TRY(make_tmap(c, &c->tmW, c->d_W, n, n, c->ldW));
      }
    }
  }
  // BAO
  d.n_bao = s.n_bao;
  if (s.n_bao > 0) {
    d.dh_mode = s.bao_dh_mode; d.rd_mode = s.rd_mode; d.col_rd = s.col_rd; d.rd_fixed = s.rd_fixed;
    TRY(upload(c, s.bao_z, (size_t)s.n_bao, &d.bao_z));
    TRY(upload(c, s.bao_value, (size_t)s.n_bao, &d.bao_val));
    TRY(upload(c, s.bao_qty, (size_t)s.n_bao, &d.bao_qty));
    TRY(upload(c, s.bao_inv_cov, (size_t)s.n_bao * s.n_bao, &d.bao_W));
  } else {
    d.rd_mode = s.rd_mode; d.col_rd = s.col_rd; d.rd_fixed = s.rd_fixed;
  }
  // CMB + GL nodes
  d.cmb_mode = s.cmb_mode;
  for (int i = 0; i < 3; i++) d.cmb_prior[i] = s.cmb_prior[i];
  for (int i = 0; i < 9; i++) d.cmb_W[i] = s.cmb_weight[i];
  if (s.family == CL_FAMILY_FULL) {
    std::vector<double> gx, gw;
    const double *px = s.gl_x, *pw = s.gl_w;
    int ngl = s.n_gl;
    if (!px || !pw || ngl <= 0) { ngl = 100; gauss_legendre(ngl, gx, gw); px = gx.data(); pw = gw.data(); }
    d.n_gl = ngl;
    TRY(upload(c, px, (size_t)ngl, &d.gl_x));
    TRY(upload(c, pw, (size_t)ngl, &d.gl_w));
  }
  // CC
  d.n_cc = s.n_cc; d.col_fcc = s.n_cc > 0 ? s.col_fcc : -1; d.cc_logdet = s.cc_logdet; d.cc_norm_sign = s.cc_norm_sign;
  if (s.n_cc > 0) {
    TRY(upload(c, s.cc_z, (size_t)s.n_cc, &d.cc_z));
    TRY(upload(c, s.cc_H, (size_t)s.n_cc, &d.cc_H));
    TRY(upload(c, s.cc_inv_cov, (size_t)s.n_cc * s.n_cc, &d.cc_W));
  }
  // Gaussian terms, prior, guard
  d.n_gc = s.n_gauss_chi2; d.n_gp = s.n_gauss_prior; d.has_bounds = s.has_bounds; d.guard_cpl = s.guard_cpl;
  for (int g = 0; g < CL_MAX_GAUSS; g++) {
    d.gc_col[g] = s.gauss_chi2_col[g]; d.gc_mean[g] = s.gauss_chi2_mean[g]; d.gc_sigma[g] = s.gauss_chi2_sigma[g];
    d.gp_col[g] = s.gauss_prior_col[g]; d.gp_mean[g] = s.gauss_prior_mean[g]; d.gp_sigma[g] = s.gauss_prior_sigma[g];
  }
  for (int j = 0; j < CL_MAX_DIM; j++) { d.lo[j] = s.lo[j]; d.hi[j] = s.hi[j]; }
  d.lp_norm = s.log_prior_norm; d.guard_value = s.guard_value;

  // kernel attributes
  S12Kernel k12 = pick_s12(d.family, d.de_model);
  CTRY(cudaFuncSetAttribute(k12, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S12Smem)));
  if (s12_lean(d, MODE_EVAL)) CTRY(cudaFuncSetAttribute(pick_s12(d.family, d.de_model, 1), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S12Smem)));
  else if (s12_fusable(d)) CTRY(cudaFuncSetAttribute(pick_s12(d.family, d.de_model, 2), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S12Smem)));
  CTRY(cudaFuncSetAttribute(k_chi2_gemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
  CTRY(cudaFuncSetAttribute(k_chi2_gemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
  CTRY(cudaFuncSetAttribute(k_chi2_ozaki<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzCfg<5>::SMEM));
  CTRY(cudaFuncSetAttribute(k_chi2_ozaki<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzCfg<6>::SMEM));
  CTRY(cudaFuncSetAttribute(k_chi2_ozaki<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzCfg<7>::SMEM));

  char buf[256];
  snprintf(buf, sizeof buf, "cosmolike_b200 abi %u, sm_100a, %s (%d SMs), n_sn=%d n_bao=%d cmb=%d n_cc=%d grid=%d%s", CL_ABI_VERSION, prop.name,
           prop.multiProcessorCount, d.n_sn, d.n_bao, d.cmb_mode, d.n_cc, d.G, d.grid_uniform ? " uniform" : "");
  c->desc = buf;
  *out = c;
  return CL_OK;
#undef TRY
#undef CTRY
}

extern "C" int cl_set_option(cl_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return CL_E_INVALID;
  std::string n(name);
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graphs(c);   // captured launches carry the options they were captured with
  }
  if (n == "cuda_graphs") { c->opt_graph = value != 0; return CL_OK; }
  if (n == "cuda_graph_max_rows") { if (value < 0) return fail(c, CL_E_INVALID, "cuda_graph_max_rows must be >= 0"); c->graph_max_rows = value; return CL_OK; }
  if (n == "max_rows_per_pass") { if (value < 128) return fail(c, CL_E_INVALID, "max_rows_per_pass must be >= 128"); c->max_rows = value; return CL_OK; }
  if (n == "gemm_ctas") { c->opt_gemm_ctas = (int)value; return CL_OK; }
  if (n == "stage12_ctas") { c->opt_s12_ctas = (int)value; return CL_OK; }
  if (n == "dbg") { c->opt_dbg = (int)value; return CL_OK; }
  if (n == "stage12_lean") { c->opt_s12_lean = value != 0; return CL_OK; }
  if (n == "fuse_planes") { c->opt_fuse_planes = value != 0; return CL_OK; }
  if (n == "chi2_guard") { c->opt_guard = value != 0; return CL_OK; }
  if (n == "chi2_guard_mode") { if (value != 0 && value != 1) return fail(c, CL_E_INVALID, "chi2_guard_mode must be 0 (probabilistic bound) or 1 (worst case)"); c->opt_guard_mode = (int)value; return CL_OK; }
  if (n == "gemm_dynamic") { c->opt_gemm_dynamic = value ? 1 : 0; return CL_OK; }
  if (n == "gemm_group_rb") { c->opt_group_rb = (int)value; return CL_OK; }
  if (n == "gemm_diag_skip") { c->opt_diag_skip = value ? 1 : 0; return CL_OK; }   // DMMA engine only
  if (n == "chi2_engine") {
    if (value != CL_CHI2_ENGINE_DMMA && value != CL_CHI2_ENGINE_TCGEN05) return fail(c, CL_E_INVALID, "chi2_engine must be 0 (FP64 DMMA) or 1 (tcgen05 int8 digit planes)");
    c->opt_engine = (int)value; return CL_OK;
  }
  if (n == "chi2_slice_tpb") { if (value != 64 && value != 128 && value != 256) return fail(c, CL_E_INVALID, "chi2_slice_tpb must be 64, 128 or 256"); c->opt_slice_tpb = (int)value; return CL_OK; }
  if (n == "chi2_slices") {
    if (value < 5 || value > 7) return fail(c, CL_E_INVALID, "chi2_slices must be 5, 6 or 7");
    c->opt_slices = (int)value; return CL_OK;
  }
  return fail(c, CL_E_INVALID, "unknown option %s", name);
}

extern "C" int cl_set_option_f64(cl_ctx* c, const char* name, double value) {
  if (!c || !name) return CL_E_INVALID;
  std::string n(name);
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graphs(c);
  }
  if (!(value >= 0.0)) return fail(c, CL_E_INVALID, "%s must be >= 0", name);
  if (n == "chi2_guard_abs") { c->guard_abs = value; return CL_OK; }
  if (n == "chi2_guard_rel") { c->guard_rel = value; return CL_OK; }
  return fail(c, CL_E_INVALID, "unknown option %s", name);
}

// eps_S of the digit-plane error bound: 2^(2 - 8 S) (1 + (S - 1) 256 / 255)
static double oz_eps(int S) { return ldexp(1.0 + (S - 1) * (256.0 / 255.0), 2 - 8 * S); }
constexpr double kGuardLambda = 8.0;   // the probabilistic bound is exceeded with probability < 2 exp(-lambda^2 / 2) = 2.5e-14 per row
// coefficient of the linear term of the selected a-priori bound (friedmann.cuh: GuardArgs)
static double oz_kappa(const cl_ctx* c) {
  return c->opt_guard_mode ? oz_eps(c->opt_slices) * c->oz_omega : kGuardLambda * oz_eps(c->opt_slices) * c->oz_omega_pr;
}

extern "C" int cl_graph_info(cl_ctx* c, int64_t out[2]) {
  if (!c || !out) return CL_E_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  out[0] = (int64_t)c->graphs.size();
  out[1] = c->graph_replays;
  return CL_OK;
}

extern "C" int cl_guard_info(cl_ctx* c, double out[4]) {
  if (!c || !out) return CL_E_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  out[0] = out[1] = 0.0; out[2] = c->opt_guard_mode ? c->oz_omega : c->oz_omega_pr; out[3] = oz_kappa(c);
  if (c->d_guard) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    int h[2] = {0, 0};
    CUDA_TRY(c, cudaDeviceSynchronize());   // the counters may be in flight on a caller-supplied stream
    CUDA_TRY(c, cudaMemcpy(h, c->d_guard, sizeof h, cudaMemcpyDeviceToHost));
    out[0] = h[0]; out[1] = h[1];
  }
  return CL_OK;
}

// ---- workspace ----
static int ensure_rows(cl_ctx* c, int64_t rows) {
  if (rows <= c->cap_rows) return CL_OK;
  int64_t cap = std::max<int64_t>(rows, std::min<int64_t>(c->max_rows, std::max<int64_t>(1024, c->cap_rows * 2)));
  CUDA_TRY(c, cudaDeviceSynchronize());  // the workspace may be in use on a caller-supplied stream
  drop_graphs(c);
  for (double** p : {&c->d_theta, &c->d_out, &c->d_R, &c->d_aux, &c->d_part, &c->d_part_u, &c->d_part_fb, &c->d_part_u_fb}) { if (*p) cudaFree(*p); *p = nullptr; }
  int guard_total = 0;   // rows flagged since creation survive a regrowth of the workspace
  if (c->d_guard) { cudaMemcpy(&guard_total, c->d_guard, sizeof(int), cudaMemcpyDeviceToHost); cudaFree(c->d_guard); c->d_guard = nullptr; }
  if (c->d_rowflag) { cudaFree(c->d_rowflag); c->d_rowflag = nullptr; }
  c->cap_rows = 0;
  CUDA_TRY(c, cudaMalloc(&c->d_theta, cap * CL_MAX_DIM * sizeof(double)));
  CUDA_TRY(c, cudaMalloc(&c->d_out, cap * 4 * sizeof(double)));
  CUDA_TRY(c, cudaMalloc(&c->d_aux, cap * AUX_COUNT * sizeof(double)));
  if (c->d_W) {
    CUDA_TRY(c, cudaMalloc(&c->d_R, cap * c->ldR * sizeof(double)));
    const int64_t t_max = std::max<int64_t>(c->T, 2 * ((c->ds.n_sn + OzCfg<7>::NT - 1) / OzCfg<7>::NT));  // either engine
    CUDA_TRY(c, cudaMalloc(&c->d_part, cap * t_max * sizeof(double)));
    CUDA_TRY(c, cudaMalloc(&c->d_part_u, cap * t_max * sizeof(double)));
    CUDA_TRY(c, cudaMalloc(&c->d_part_fb, cap * c->T * sizeof(double)));
    CUDA_TRY(c, cudaMalloc(&c->d_part_u_fb, cap * c->T * sizeof(double)));
    const size_t gbytes = (2 + (size_t)(cap + 127) / 128) * sizeof(int);
    CUDA_TRY(c, cudaMalloc(&c->d_guard, gbytes));
    CUDA_TRY(c, cudaMemset(c->d_guard, 0, gbytes));
    CUDA_TRY(c, cudaMemcpy(c->d_guard, &guard_total, sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMalloc(&c->d_rowflag, cap));
  }
  c->cap_rows = cap;
  return CL_OK;
}

static int ensure_pinned(cl_ctx* c, int64_t theta_elems, int64_t out_elems) {
  if (theta_elems > c->h_theta_cap || out_elems > c->h_out_cap) { CUDA_TRY(c, cudaDeviceSynchronize()); drop_graphs(c); }
  if (theta_elems > c->h_theta_cap) {
    if (c->h_theta) cudaFreeHost(c->h_theta);
    c->h_theta = nullptr; c->h_theta_cap = 0;
    CUDA_TRY(c, cudaMallocHost(&c->h_theta, theta_elems * sizeof(double)));
    c->h_theta_cap = theta_elems;
  }
  if (out_elems > c->h_out_cap) {
    if (c->h_out) cudaFreeHost(c->h_out);
    c->h_out = nullptr; c->h_out_cap = 0;
    CUDA_TRY(c, cudaMallocHost(&c->h_out, out_elems * sizeof(double)));
    c->h_out_cap = out_elems;
  }
  return CL_OK;
}

static int ensure_scratch(cl_ctx* c, int64_t bytes) {
  if (bytes <= c->scratch_bytes) return CL_OK;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  if (c->d_scratch) cudaFree(c->d_scratch);
  c->d_scratch = nullptr; c->scratch_bytes = 0;
  CUDA_TRY(c, cudaMalloc(&c->d_scratch, bytes));
  c->scratch_bytes = bytes;
  return CL_OK;
}

static int launch_s12(cl_ctx* c, const Stage12Args& a, cudaStream_t st) {
  // with digit planes to write: the SN-only or the small-probe instantiation (run_pass has checked that the spec qualifies)
  const int lean = !c->opt_s12_lean ? 0 : s12_lean(c->ds, a.mode) ? 1 : (a.planes != nullptr ? 2 : 0);
  S12Kernel k = pick_s12(c->ds.family, c->ds.de_model, lean);
  // ~64 CTAs per resident slot (three rows per CTA at B = 65536: measured 0.924 -> 0.887 ms against 8 per slot, one row per CTA is
  // slower again), and the same number of rows for every CTA (no ragged tail at small batches)
  int64_t want = c->opt_s12_ctas > 0 ? c->opt_s12_ctas : (int64_t)c->sm_count * 3 * 64;
  int64_t rows_per_cta = (a.B + want - 1) / want;
  int grid = (int)((a.B + rows_per_cta - 1) / rows_per_cta);
  k<<<grid, kS12Threads, sizeof(S12Smem), st>>>(c->ds, a);
  c->launches++;
  CUDA_TRY(c, cudaGetLastError());
  return CL_OK;
}

// digit planes for the tcgen05 engine: W planes once per slice count, residual planes sized like the workspace
template <int S>
static int oz_slice_launch(cl_ctx* c, const double* src, int64_t ld_src, int64_t rows, int n, int8_t* dst, double* scale, cudaStream_t st) {
  const int tpb = c->opt_slice_tpb;   // threads per block of the slicing kernel (one warp per row)
  const unsigned grid = (unsigned)((rows + tpb / 32 - 1) / (tpb / 32));
  const int trips = (int)(c->oz_ld / 128);
  // register-resident rows (one HBM read) when the row is 32-byte aligned and short enough
  const bool reg_ok = (ld_src % 4 == 0) && (((uintptr_t)src & 31) == 0) && trips <= 16;
#define OZ_REG(T) k_oz_slice_rows_reg<S, T><<<grid, tpb, 0, st>>>(src, ld_src, rows, n, dst, c->oz_ld, scale)
  if (reg_ok && trips <= 8) OZ_REG(8);
  else if (reg_ok && trips <= 12) OZ_REG(12);
  else if (reg_ok && trips <= 14) OZ_REG(14);
  else if (reg_ok) OZ_REG(16);
  else k_oz_slice_rows<S><<<grid, tpb, 0, st>>>(src, ld_src, rows, n, dst, c->oz_ld, scale);
#undef OZ_REG
  c->launches++;
  CUDA_TRY(c, cudaGetLastError());
  return CL_OK;
}
static int oz_slice(cl_ctx* c, int S, const double* src, int64_t ld_src, int64_t rows, int n, int8_t* dst, double* scale, cudaStream_t st) {
  return S == 5 ? oz_slice_launch<5>(c, src, ld_src, rows, n, dst, scale, st)
       : S == 6 ? oz_slice_launch<6>(c, src, ld_src, rows, n, dst, scale, st)
                : oz_slice_launch<7>(c, src, ld_src, rows, n, dst, scale, st);
}
static int oz_tile_cols(int S) { return S == 5 ? OzCfg<5>::NT : S == 6 ? OzCfg<6>::NT : OzCfg<7>::NT; }

static int ensure_planes(cl_ctx* c, int64_t rows, cudaStream_t st) {
  const int S = c->opt_slices, n = c->ds.n_sn;
  if (c->oz_slices_built != S) {
    CUDA_TRY(c, cudaDeviceSynchronize());
    drop_graphs(c);
    for (void** p : {(void**)&c->d_Ws, (void**)&c->d_Rs, (void**)&c->d_wscale, (void**)&c->d_rscale}) { if (*p) cudaFree(*p); *p = nullptr; }
    c->oz_cap_rows = 0;
    c->oz_ld = ((int64_t)n + 127) & ~127LL;
    c->oz_T = (n + oz_tile_cols(S) - 1) / oz_tile_cols(S);
    CUDA_TRY(c, cudaMalloc(&c->d_Ws, (size_t)S * n * c->oz_ld));
    CUDA_TRY(c, cudaMalloc(&c->d_wscale, n * sizeof(double)));
    int rc = oz_slice(c, S, c->d_W, c->ldW, n, n, c->d_Ws, c->d_wscale, st);
    if (rc != CL_OK) return rc;
    rc = make_tmap_planes(c, &c->tmWs, c->d_Ws, n, n, S, c->oz_ld, oz_tile_cols(S), S);
    if (rc != CL_OK) return rc;
    c->oz_slices_built = S;
  }
  if (rows > c->oz_cap_rows) {
    CUDA_TRY(c, cudaDeviceSynchronize());
    drop_graphs(c);
    for (void** p : {(void**)&c->d_Rs, (void**)&c->d_rscale}) { if (*p) cudaFree(*p); *p = nullptr; }
    c->oz_cap_rows = 0;
    const int64_t cap = std::max(rows, c->cap_rows);
    CUDA_TRY(c, cudaMalloc(&c->d_Rs, (size_t)S * cap * c->oz_ld));
    CUDA_TRY(c, cudaMalloc(&c->d_rscale, cap * sizeof(double)));
    c->oz_cap_rows = cap;
  }
  return CL_OK;
}

template <int S>
static void oz_launch(int grid, cudaStream_t st, const CUtensorMap& tmR, const CUtensorMap& tmW, const OzArgs& g) {
  k_chi2_ozaki<S><<<grid, kOzThreads, OzCfg<S>::SMEM, st>>>(tmR, tmW, g);
}

// stage 3 with the tcgen05 engine: slice the residual rows, then the int8 contraction
static int run_stage3_planes(cl_ctx* c, int64_t rows, cudaStream_t st, bool record, bool moments, bool planes_ready) {
  int rc = ensure_planes(c, rows, st);
  if (rc != CL_OK) return rc;
  const int S = c->opt_slices, n = c->ds.n_sn;
  if (!planes_ready) {   // stage 2 has not written the planes itself
    rc = oz_slice(c, S, c->d_R, c->ldR, rows, n, c->d_Rs, c->d_rscale, st);
    if (rc != CL_OK) return rc;
  }
  if (record) CUDA_TRY(c, cudaEventRecord(c->ev[6], st));
  CUtensorMap tmRs;
  rc = make_tmap_planes(c, &tmRs, c->d_Rs, n, rows, S, c->oz_ld, kOzM, S);   // OzCfg<S>::SLO planes per box
  if (rc != CL_OK) return rc;
  OzArgs g{};
  g.B = rows; g.N = n; g.T = c->oz_T; g.n_rb = (int)((rows + kOzM - 1) / kOzM);
  g.part = c->d_part; g.rowscale = c->d_rscale; g.colscale = c->d_wscale; g.counter = c->d_counter;
  g.part_u = moments ? c->d_part_u : nullptr; g.u = c->d_u;
  g.dbg_skip = (c->opt_dbg >> 4) & 31;   // dbg bits 4, 5: skip the W / R plane loads (timing experiments, results invalid)
  g.prof = nullptr; g.trace = nullptr;
  if (c->opt_dbg & 4) {   // cycle counters of the contraction kernel, printed after the launch (profiling only)
    rc = ensure_scratch(c, 8 * 8 * 1024 + 8 * 16384);
    if (rc != CL_OK) return rc;
    g.prof = reinterpret_cast<long long*>(c->d_scratch);
    CUDA_TRY(c, cudaMemsetAsync(g.prof, 0, 8 * 8 * 1024 + 8 * 16384, st));
    if (c->opt_dbg & 8) g.trace = g.prof + 8 * 1024;   // event trace of CTA 0, written to $COSMOLIKE_TRACE (default oz_trace.txt)
  }
  // row blocks per L2 group: the S digit planes of a group's rows (+ the W planes) stay L2-resident across its column tiles
  int grp = c->opt_group_rb > 0 ? c->opt_group_rb : (int)std::max<int64_t>(8, ((28LL << 20) / ((int64_t)kOzM * c->oz_ld * S)) & ~7LL);
  g.group_rb = std::min(grp, g.n_rb);
  CUDA_TRY(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int), st));
  const int64_t items = (int64_t)g.n_rb * g.T;
  if (items >= (1LL << 30) || g.n_rb >= (1 << 24)) return fail(c, CL_E_INVALID, "tcgen05 engine: too many tiles in one pass (lower max_rows_per_pass)");
  const int grid = (int)std::min<int64_t>(items, c->opt_gemm_ctas > 0 ? c->opt_gemm_ctas : c->sm_count);
  if (S == 5) oz_launch<5>(grid, st, tmRs, c->tmWs, g);
  else if (S == 6) oz_launch<6>(grid, st, tmRs, c->tmWs, g);
  else oz_launch<7>(grid, st, tmRs, c->tmWs, g);
  c->launches++;
  CUDA_TRY(c, cudaGetLastError());
  if (g.prof) {
    std::vector<long long> h(8 * grid);
    CUDA_TRY(c, cudaMemcpyAsync(h.data(), g.prof, h.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    double s[8] = {0};
    double pre = 0;
    for (int i = 0; i < grid; i++) { pre += (double)(h[8 * i + 4] >> 32) / grid; h[8 * i + 4] &= 0xffffffffLL; }
    for (int i = 0; i < grid; i++) for (int j = 0; j < 8; j++) s[j] += (double)h[8 * i + j] / grid;
    fprintf(stderr, "[oz prof] epilogue warp: tcgen05.ld+wait %.0f cyc, column-scale prologue %.0f cyc\n", s[7], pre);
    fprintf(stderr, "[oz prof] per CTA: total %.0f cyc, wait smem-full %.0f (%.1f%%), wait level-empty %.0f (%.1f%%), wait queue %.0f, k-blocks %.0f | epilogue warp: wait level-full %.0f (%.1f%%), tiles %.0f\n",
            s[0], s[1], 100 * s[1] / s[0], s[2], 100 * s[2] / s[0], s[3], s[4], s[5], 100 * s[5] / s[0], s[6]);
    if (g.trace) {
      std::vector<long long> t(1 + 11 * kOzTraceCap);
      CUDA_TRY(c, cudaMemcpy(t.data(), g.trace, t.size() * 8, cudaMemcpyDeviceToHost));
      const char* path = getenv("COSMOLIKE_TRACE");
      if (FILE* f = fopen(path ? path : "oz_trace.txt", "w")) {
        for (size_t i = 1; i < t.size(); i++) if (t[i]) fprintf(f, "%lld %lld\n", t[i] >> 44, t[i] & ((1LL << 44) - 1));
        fclose(f);
      }
    }
  }
  return CL_OK;
}

// stage 3 on the FP64 tensor pipe (chi2_gemm.cuh).  `guard` != nullptr: the accuracy-guard fallback pass of the tcgen05 engine,
// restricted to the flagged row blocks (returns at once on the device when nothing is flagged).
static int run_stage3_dmma(cl_ctx* c, int64_t rows, cudaStream_t st, bool moments, double* part, double* part_u, const int* guard) {
  CUtensorMap tmR;
  int rc = make_tmap(c, &tmR, c->d_R, rows, c->ds.n_sn, c->ldR);
  if (rc != CL_OK) return rc;
  GemmArgs g{};
  g.B = rows; g.N = c->ds.n_sn; g.T = c->T; g.n_rb = (int)((rows + kBM - 1) / kBM);
  g.part = part; g.part_u = part_u; g.u = c->d_u; g.diag_skip = c->opt_diag_skip;
  g.counter = nullptr; g.guard = guard;
  if (c->opt_gemm_dynamic || guard) {
    // row blocks per L2 group: <= ~58 MB of residual rows, a multiple of 8 (the R rows of a group + the 23 MB of W stay
    // L2-resident across the group's column tiles; measured on B200: 32 row blocks at N=1701 -> 1.2 GB of DRAM reads per
    // launch instead of 6.5 GB, and multiples of 8 schedule ~2 % better than odd group sizes)
    int grp = c->opt_group_rb > 0 ? c->opt_group_rb
                                  : (int)std::max<int64_t>(8, ((58LL << 20) / ((int64_t)kBM * c->ldR * 8)) & ~7LL);
    g.group_rb = std::min(grp, g.n_rb);
    g.counter = c->d_counter;
    CUDA_TRY(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int), st));
  }
  int64_t items = (int64_t)g.n_rb * g.T;
  int grid = (int)std::min<int64_t>(items, c->opt_gemm_ctas > 0 ? c->opt_gemm_ctas : c->sm_count);
  if (moments) k_chi2_gemm<true><<<grid, kGemmThreads, kGemmSmemBytes, st>>>(tmR, c->tmW, g);
  else k_chi2_gemm<false><<<grid, kGemmThreads, kGemmSmemBytes, st>>>(tmR, c->tmW, g);
  c->launches++;
  CUDA_TRY(c, cudaGetLastError());
  return CL_OK;
}

// one pass over `rows` device-resident parameter vectors: stage 1+2 -> stage 3 -> finalize
static int run_pass(cl_ctx* c, const double* d_theta, int64_t rows, int64_t ld, int what, double* d_out, double* d_comps,
                    bool moments, cudaStream_t st, bool record) {
  int rc = ensure_rows(c, rows);
  if (rc != CL_OK) return rc;
  const bool large = c->d_W != nullptr;
  Stage12Args a{};
  a.theta = d_theta; a.B = rows; a.ld = ld; a.mode = MODE_EVAL; a.what = moments ? CL_OUT_CHI2 : what;
  a.R = c->d_R; a.ldR = c->ldR; a.aux = c->d_aux; a.zero_offset = moments ? 1 : 0; a.dbg = c->opt_dbg;
  // int32 level accumulators hold S products of |d_i d_j| <= 2^14 over n_sn terms: exact up to n_sn = 2^31 / (7 * 2^14) = 18724
  const bool planes = large && c->opt_engine == CL_CHI2_ENGINE_TCGEN05 && c->ds.n_sn <= 16384;
  // stage 2 writes the digit planes itself when the lean kernel runs its fast SN path and a thread can hold its share of the row
  const DevSpec& d = c->ds;
  const bool fused = planes && c->opt_fuse_planes && c->opt_s12_lean && s12_fusable(d) && !(c->opt_dbg & 2);
  if (fused) {
    rc = ensure_planes(c, rows, st);
    if (rc != CL_OK) return rc;
    a.planes = reinterpret_cast<signed char*>(c->d_Rs); a.planes_ld = c->oz_ld; a.rowscale = c->d_rscale; a.planes_S = c->opt_slices;
  }
  if (record) CUDA_TRY(c, cudaEventRecord(c->ev[1], st));
  rc = launch_s12(c, a, st);
  if (rc != CL_OK) return rc;
  if (record) CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
  if (planes) {
    rc = run_stage3_planes(c, rows, st, record, moments, fused);
    if (rc != CL_OK) return rc;
  } else if (large) {
    if (record) CUDA_TRY(c, cudaEventRecord(c->ev[6], st));
    rc = run_stage3_dmma(c, rows, st, moments, c->d_part, c->d_part_u, nullptr);
    if (rc != CL_OK) return rc;
  }
  if (record) CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
  // accuracy guard (tcgen05 engine): rows whose a-priori error bound exceeds the tolerance are flagged by the finalize kernel
  // and recomputed on the FP64 tensor pipe; the three fallback launches return at once when nothing is flagged
  const bool guard = planes && c->opt_guard;
  GuardArgs q{};
  if (guard) {
    q.rowscale = c->d_rscale; q.kappa = oz_kappa(c); q.kappa_sq = oz_eps(c->opt_slices) * c->oz_omega; q.tol_abs = c->guard_abs; q.tol_rel = c->guard_rel;
    q.guard = c->d_guard; q.rowflag = c->d_rowflag; q.only_flagged = 0;
    CUDA_TRY(c, cudaMemsetAsync(c->d_guard + 1, 0, (1 + (size_t)(rows + 127) / 128) * sizeof(int), st));
  }
  const unsigned fgrid = (unsigned)((rows + 255) / 256);
  FinalizeArgs f{};
  f.B = rows; f.what = what; f.n_part = planes ? 2 * c->oz_T : c->T; f.sn_large = large ? 1 : 0;
  f.part = c->d_part; f.aux = c->d_aux; f.theta = d_theta; f.ld = ld; f.out = d_out; f.comps = d_comps; f.guard_value = c->ds.guard_value; f.q = q;
  if (moments) k_sum_parts<<<fgrid, 256, 0, st>>>(c->d_part, c->d_part_u, f.n_part, rows, c->uu, d_out, q);   // d_out[rows][3] = (yy, yu, uu)
  else k_finalize<<<fgrid, 256, 0, st>>>(c->ds, f);
  c->launches++;
  CUDA_TRY(c, cudaGetLastError());
  if (guard) {
    if (fused) {   // the FP64 residual rows of the flagged blocks were never written: stage 1+2 again for those rows
      Stage12Args a2 = a;
      a2.planes = nullptr; a2.rowscale = nullptr; a2.guard = c->d_guard;
      S12Kernel k = pick_s12(c->ds.family, c->ds.de_model, s12_lean(c->ds, MODE_EVAL) ? 1 : 0);
      k<<<c->sm_count * 3, kS12Threads, sizeof(S12Smem), st>>>(c->ds, a2);
      c->launches++;
      CUDA_TRY(c, cudaGetLastError());
    }
    rc = run_stage3_dmma(c, rows, st, moments, c->d_part_fb, c->d_part_u_fb, c->d_guard);
    if (rc != CL_OK) return rc;
    q.only_flagged = 1;
    f.q = q; f.part = c->d_part_fb; f.n_part = c->T;
    if (moments) k_sum_parts<<<fgrid, 256, 0, st>>>(c->d_part_fb, c->d_part_u_fb, c->T, rows, c->uu, d_out, q);
    else k_finalize<<<fgrid, 256, 0, st>>>(c->ds, f);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
  }
  if (record) CUDA_TRY(c, cudaEventRecord(c->ev[4], st));
  return CL_OK;
}

// ---- CUDA graphs of small evaluations ----
// Returns 1 when the call was served by a graph replay, 0 when the caller must run it the ordinary way, < 0 on error.
// `body` enqueues the work on `st` (no allocations, no synchronisation: the first call of a shape ran the ordinary way).
template <typename Body>
static int graph_run(cl_ctx* c, const cl_ctx::GraphKey& key, cudaStream_t st, Body body) {
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    if (!c->graph_seen.count(key)) { c->graph_seen.insert(key); return 0; }   // first sight: ordinary launches settle the workspace
    if (c->graphs.size() >= 64) drop_graphs(c);
    const int64_t l0 = c->launches;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return 0; }
    const int rc = body();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    const int64_t n_launch = c->launches - l0;
    c->launches = l0;
    if (rc != CL_OK || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return rc != CL_OK ? rc : 0;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) { cudaGetLastError(); return 0; }
    it = c->graphs.emplace(key, cl_ctx::GraphEntry{exec, n_launch}).first;
  }
  CUDA_TRY(c, cudaGraphLaunch(it->second.exec, st));
  c->launches += it->second.launches;
  c->graph_replays++;
  return 1;
}
static bool graph_ok(const cl_ctx* c, int64_t rows) {
  return c->opt_graph && c->opt_dbg == 0 && rows <= c->graph_max_rows && rows <= c->max_rows && rows <= c->cap_rows;
}

static int check_eval_args(cl_ctx* c, const double* theta, int64_t B, int64_t ld, const void* out) {
  if (!c) return CL_E_INVALID;
  if (B < 0 || (B > 0 && (!theta || !out))) return fail(c, CL_E_INVALID, "NULL buffer");
  if (ld < c->ds.ndim) return fail(c, CL_E_INVALID, "ld (%lld) < ndim (%d)", (long long)ld, c->ds.ndim);
  return CL_OK;
}

extern "C" int cl_eval_device(cl_ctx* c, const double* d_theta, int64_t B, int64_t ld, int what, double* d_out, void* stream) {
  int rc = check_eval_args(c, d_theta, B, ld, d_out);
  if (rc != CL_OK || B == 0) return rc;
  if (what < CL_OUT_CHI2 || what > CL_OUT_LOGPROB) return fail(c, CL_E_INVALID, "bad output selector");
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  if (graph_ok(c, B) && st != cudaStreamLegacy && st != cudaStreamPerThread) {   // small batch on fixed buffers: one graph launch
    const cl_ctx::GraphKey key{1, B, what, 1, d_theta, d_out, (const void*)st, ld};
    rc = graph_run(c, key, st, [&]() { return run_pass(c, d_theta, B, ld, what, d_out, nullptr, false, st, false); });
    if (rc != 0) return rc < 0 ? rc : CL_OK;
  }
  c->ev = c->evring[c->n_timed % cl_ctx::kRing];
  CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
  for (int64_t r0 = 0; r0 < B; r0 += c->max_rows) {
    int64_t rows = std::min(c->max_rows, B - r0);
    rc = run_pass(c, d_theta + r0 * ld, rows, ld, what, d_out + r0, nullptr, false, st, r0 == 0);
    if (rc != CL_OK) return rc;
  }
  CUDA_TRY(c, cudaEventRecord(c->ev[5], st));
  c->n_timed++;
  return CL_OK;
}

// page-locked host memory (cudaHostAlloc / cudaHostRegister / cl_host_alloc) can be the source or target of a DMA as it is
static bool is_page_locked(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

extern "C" int cl_host_alloc(cl_ctx* c, size_t bytes, void** ptr) {
  if (!ptr) return fail(c, CL_E_INVALID, "ptr is NULL");
  *ptr = nullptr;
  if (c) CUDA_TRY(c, cudaSetDevice(c->device));
  cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(c, CL_E_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
  return CL_OK;
}

extern "C" int cl_host_free(cl_ctx* c, void* ptr) {
  if (!ptr) return CL_OK;
  cudaError_t e = cudaFreeHost(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(c, CL_E_CUDA, "cudaFreeHost failed: %s", cudaGetErrorString(e)); }
  return CL_OK;
}

// host-memory evaluation with `width` outputs per row (1: out, 4: components, 3: moments)
static int eval_host(cl_ctx* c, const double* theta, int64_t B, int64_t ld, int what, double* out, int width, bool moments) {
  int rc = check_eval_args(c, theta, B, ld, out);
  if (rc != CL_OK || B == 0) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int nd = c->ds.ndim;
  cudaStream_t st = c->stream;
  if (graph_ok(c, B) && c->h_theta_cap >= B * nd && c->h_out_cap >= B * 4) {
    // small batch: host copy into the staging area, ONE graph launch (H2D, the whole pass, D2H), host copy out
    const cl_ctx::GraphKey key{0, B, what, width + (moments ? 8 : 0), nullptr, nullptr, nullptr, 0};
    if (ld == nd) memcpy(c->h_theta, theta, (size_t)B * nd * sizeof(double));
    else for (int64_t i = 0; i < B; i++) memcpy(c->h_theta + i * nd, theta + i * ld, nd * sizeof(double));
    rc = graph_run(c, key, st, [&]() {
      cudaError_t e = cudaMemcpyAsync(c->d_theta, c->h_theta, B * nd * sizeof(double), cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return fail(c, CL_E_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
      int r;
      if (width == 1) r = run_pass(c, c->d_theta, B, nd, what, c->d_out, nullptr, false, st, false);
      else if (width == 4) r = run_pass(c, c->d_theta, B, nd, CL_OUT_CHI2, nullptr, c->d_out, false, st, false);
      else r = run_pass(c, c->d_theta, B, nd, CL_OUT_CHI2, c->d_out, nullptr, true, st, false);
      if (r != CL_OK) return r;
      e = cudaMemcpyAsync(c->h_out, c->d_out, B * width * sizeof(double), cudaMemcpyDeviceToHost, st);
      return e == cudaSuccess ? (int)CL_OK : fail(c, CL_E_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
    });
    if (rc < 0) return rc;
    if (rc == 1) {
      CUDA_TRY(c, cudaStreamSynchronize(st));
      memcpy(out, c->h_out, (size_t)B * width * sizeof(double));
      return CL_OK;
    }
  }
  c->ev = c->evring[c->n_timed % cl_ctx::kRing];
  const bool theta_pinned = is_page_locked(theta), out_pinned = is_page_locked(out);
  CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
  for (int64_t r0 = 0; r0 < B; r0 += c->max_rows) {
    int64_t rows = std::min(c->max_rows, B - r0);
    rc = ensure_rows(c, rows);
    if (rc != CL_OK) return rc;
    rc = ensure_pinned(c, theta_pinned ? 0 : rows * nd, out_pinned ? 0 : rows * 4);
    if (rc != CL_OK) return rc;
    if (r0 > 0) CUDA_TRY(c, cudaStreamSynchronize(st));  // staging buffers are reused
    if (theta_pinned) {   // DMA straight from the caller's page-locked rows
      CUDA_TRY(c, cudaMemcpy2DAsync(c->d_theta, nd * sizeof(double), theta + r0 * ld, ld * sizeof(double), nd * sizeof(double),
                                    (size_t)rows, cudaMemcpyHostToDevice, st));
    } else {
      if (ld == nd) memcpy(c->h_theta, theta + r0 * ld, (size_t)rows * nd * sizeof(double));   // contiguous rows: one copy
      else for (int64_t i = 0; i < rows; i++) memcpy(c->h_theta + i * nd, theta + (r0 + i) * ld, nd * sizeof(double));
      CUDA_TRY(c, cudaMemcpyAsync(c->d_theta, c->h_theta, rows * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    double* d_res = c->d_out;
    if (width == 1) rc = run_pass(c, c->d_theta, rows, nd, what, d_res, nullptr, false, st, r0 == 0);
    else if (width == 4) rc = run_pass(c, c->d_theta, rows, nd, CL_OUT_CHI2, nullptr, d_res, false, st, r0 == 0);
    else rc = run_pass(c, c->d_theta, rows, nd, CL_OUT_CHI2, d_res, nullptr, true, st, r0 == 0);   // d_res[rows][3] = (yy, yu, uu)
    if (rc != CL_OK) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out_pinned ? out + r0 * width : c->h_out, d_res, rows * width * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaEventRecord(c->ev[5], st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    if (!out_pinned) memcpy(out + r0 * width, c->h_out, rows * width * sizeof(double));
  }
  c->n_timed++;
  return CL_OK;
}

extern "C" int cl_eval(cl_ctx* c, const double* theta, int64_t B, int64_t ld, int what, double* out) {
  if (what < CL_OUT_CHI2 || what > CL_OUT_LOGPROB) return fail(c, CL_E_INVALID, "bad output selector");
  return eval_host(c, theta, B, ld, what, out, 1, false);
}

extern "C" int cl_eval_components(cl_ctx* c, const double* theta, int64_t B, int64_t ld, double* out) {
  return eval_host(c, theta, B, ld, CL_OUT_CHI2, out, 4, false);
}

extern "C" int cl_eval_sn_moments(cl_ctx* c, const double* theta, int64_t B, int64_t ld, double* out) {
  if (!c) return CL_E_INVALID;
  if (!c->d_W || c->ds.col_offset < 0) return fail(c, CL_E_INVALID, "sn moments need a large SN block with an offset column");
  return eval_host(c, theta, B, ld, CL_OUT_CHI2, out, 3, true);
}

// ---- multi-GPU: NCCL communicator per context, sharded evaluation with an all-gather of the results ----
#define NCCL_TRY(ctx, expr)                                                                                              \
  do {                                                                                                                   \
    int r_ = (expr);                                                                                                     \
    if (r_ != 0) return fail(ctx, CL_E_CUDA, "%s failed: %s", #expr, nccl_api().GetErrorString ? nccl_api().GetErrorString(r_) : "NCCL error"); \
  } while (0)

extern "C" int cl_comm_unique_id(void* uid) {
  if (!uid) return fail(nullptr, CL_E_INVALID, "uid is NULL");
  NcclApi& api = nccl_api();
  if (!api.handle) return fail(nullptr, CL_E_INVALID, "NCCL unavailable: %s", api.why ? api.why : "?");
  static_assert(sizeof(NcclUid) == CL_NCCL_UID_BYTES, "ncclUniqueId is 128 bytes");
  int r = api.GetUniqueId(reinterpret_cast<NcclUid*>(uid));
  if (r != 0) return fail(nullptr, CL_E_CUDA, "ncclGetUniqueId failed: %s", api.GetErrorString(r));
  return CL_OK;
}

extern "C" int cl_comm_init(cl_ctx* c, int rank, int nranks, const void* uid) {
  if (!c || !uid) return CL_E_INVALID;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, CL_E_INVALID, "bad rank %d of %d", rank, nranks);
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->comm) return fail(c, CL_E_INVALID, "the context already has a communicator");
  NcclApi& api = nccl_api();
  if (!api.handle) return fail(c, CL_E_INVALID, "NCCL unavailable: %s", api.why ? api.why : "?");
  CUDA_TRY(c, cudaSetDevice(c->device));
  NcclUid id;
  memcpy(&id, uid, sizeof id);
  NCCL_TRY(c, api.CommInitRank(&c->comm, nranks, id, rank));
  c->comm_rank = rank; c->comm_size = nranks;
  return CL_OK;
}

extern "C" int cl_comm_destroy(cl_ctx* c) {
  if (!c) return CL_E_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->comm) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    nccl_api().CommDestroy(c->comm);
    c->comm = nullptr; c->comm_size = 0; c->comm_rank = 0;
  }
  return CL_OK;
}

extern "C" int cl_comm_info(const cl_ctx* c, int* rank, int* nranks) {
  if (!c) return CL_E_INVALID;
  if (rank) *rank = c->comm_rank;
  if (nranks) *nranks = c->comm_size;
  return CL_OK;
}

// host -> device copy of rows [r0, r0 + rows) of a caller's theta (page-locked: DMA as it is; pageable: through the staging area)
static int upload_theta(cl_ctx* c, const double* theta, int64_t r0, int64_t rows, int64_t ld, bool pinned, cudaStream_t st) {
  const int nd = c->ds.ndim;
  if (pinned) {
    CUDA_TRY(c, cudaMemcpy2DAsync(c->d_theta, nd * sizeof(double), theta + r0 * ld, ld * sizeof(double), nd * sizeof(double),
                                  (size_t)rows, cudaMemcpyHostToDevice, st));
  } else {
    if (ld == nd) memcpy(c->h_theta, theta + r0 * ld, (size_t)rows * nd * sizeof(double));
    else for (int64_t i = 0; i < rows; i++) memcpy(c->h_theta + i * nd, theta + (r0 + i) * ld, nd * sizeof(double));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_theta, c->h_theta, rows * nd * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  return CL_OK;
}

extern "C" int cl_eval_allgather_device(cl_ctx* c, const double* d_theta, int64_t B, int64_t ld, int what, double* d_out_all, void* stream) {
  int rc = check_eval_args(c, d_theta, B, ld, d_out_all);
  if (rc != CL_OK) return rc;
  if (what < CL_OUT_CHI2 || what > CL_OUT_LOGPROB) return fail(c, CL_E_INVALID, "bad output selector");
  if (!c->comm) return fail(c, CL_E_INVALID, "no communicator: call cl_comm_init first");
  if (B == 0) return CL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  c->ev = c->evring[c->n_timed % cl_ctx::kRing];
  CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
  double* mine = d_out_all + (int64_t)c->comm_rank * B;
  for (int64_t r0 = 0; r0 < B; r0 += c->max_rows) {
    int64_t rows = std::min(c->max_rows, B - r0);
    rc = run_pass(c, d_theta + r0 * ld, rows, ld, what, mine + r0, nullptr, false, st, r0 == 0);
    if (rc != CL_OK) return rc;
  }
  NCCL_TRY(c, nccl_api().AllGather(mine, d_out_all, (size_t)B, kNcclFloat64, c->comm, st));   // in place: mine = recv + rank * count
  CUDA_TRY(c, cudaEventRecord(c->ev[5], st));
  c->n_timed++;
  return CL_OK;
}

extern "C" int cl_eval_allgather(cl_ctx* c, const double* theta, int64_t B, int64_t ld, int what, double* out_all, int root) {
  if (!c) return CL_E_INVALID;
  if (!c->comm) return fail(c, CL_E_INVALID, "no communicator: call cl_comm_init first");
  if (what < CL_OUT_CHI2 || what > CL_OUT_LOGPROB) return fail(c, CL_E_INVALID, "bad output selector");
  if (root >= c->comm_size) return fail(c, CL_E_INVALID, "root %d out of range", root);
  const bool receive = root < 0 || root == c->comm_rank;
  if (B < 0 || (B > 0 && (!theta || (receive && !out_all)))) return fail(c, CL_E_INVALID, "NULL buffer");
  if (ld < c->ds.ndim) return fail(c, CL_E_INVALID, "ld (%lld) < ndim (%d)", (long long)ld, c->ds.ndim);
  if (B == 0) return CL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int nd = c->ds.ndim, W = c->comm_size;
  if (B > c->gather_cap) {
    CUDA_TRY(c, cudaStreamSynchronize(st));
    if (c->d_gather) cudaFree(c->d_gather);
    c->d_gather = nullptr; c->gather_cap = 0;
    CUDA_TRY(c, cudaMalloc(&c->d_gather, (size_t)W * B * sizeof(double)));
    c->gather_cap = B;
  }
  c->ev = c->evring[c->n_timed % cl_ctx::kRing];
  const bool theta_pinned = is_page_locked(theta), out_pinned = receive && is_page_locked(out_all);
  CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
  double* mine = c->d_gather + (int64_t)c->comm_rank * B;
  int rc;
  for (int64_t r0 = 0; r0 < B; r0 += c->max_rows) {
    int64_t rows = std::min(c->max_rows, B - r0);
    rc = ensure_rows(c, rows);
    if (rc != CL_OK) return rc;
    rc = ensure_pinned(c, theta_pinned ? 0 : rows * nd, (receive && !out_pinned) ? (int64_t)W * B : 0);
    if (rc != CL_OK) return rc;
    if (r0 > 0 && !theta_pinned) CUDA_TRY(c, cudaStreamSynchronize(st));   // the staging buffer is reused
    rc = upload_theta(c, theta, r0, rows, ld, theta_pinned, st);
    if (rc != CL_OK) return rc;
    rc = run_pass(c, c->d_theta, rows, nd, what, mine + r0, nullptr, false, st, r0 == 0);
    if (rc != CL_OK) return rc;
    if (r0 + rows < B) CUDA_TRY(c, cudaStreamSynchronize(st));   // d_theta is reused by the next pass
  }
  NCCL_TRY(c, nccl_api().AllGather(mine, c->d_gather, (size_t)B, kNcclFloat64, c->comm, st));
  if (receive) CUDA_TRY(c, cudaMemcpyAsync(out_pinned ? out_all : c->h_out, c->d_gather, (size_t)W * B * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaEventRecord(c->ev[5], st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  if (receive && !out_pinned) memcpy(out_all, c->h_out, (size_t)W * B * sizeof(double));
  c->n_timed++;
  return CL_OK;
}

// ---- profile-likelihood grids generated on the device ----
extern "C" int cl_eval_grid(cl_ctx* c, const cl_grid* grid, int64_t first, int64_t count, int what, double* out, cl_grid_stats* stats) {
  if (!c || !grid || !stats) return CL_E_INVALID;
  const bool moments = what == CL_GRID_PROFILE || what == CL_GRID_MARGINAL;
  if (!moments && (what < CL_OUT_CHI2 || what > CL_OUT_LOGPROB)) return fail(c, CL_E_INVALID, "bad output selector");
  if (moments && (!c->d_W || c->ds.col_offset < 0)) return fail(c, CL_E_INVALID, "offset profiling needs a large SN block with an offset column");
  const int nd = c->ds.ndim;
  if (grid->n_axes < 1 || grid->n_axes > nd) return fail(c, CL_E_INVALID, "n_axes out of range");
  DevGrid dg{};
  dg.ndim = nd; dg.n_axes = grid->n_axes;
  long double total = 1.0L;
  bool used[CL_MAX_DIM] = {};
  for (int a = 0; a < grid->n_axes; a++) {
    const int col = grid->col[a];
    if (col < 0 || col >= nd || used[col]) return fail(c, CL_E_INVALID, "axis %d: bad or repeated theta column", a);
    if (grid->n[a] < 1) return fail(c, CL_E_INVALID, "axis %d has no points", a);
    used[col] = true;
    dg.col[a] = col; dg.n[a] = grid->n[a]; dg.lo[a] = grid->lo[a]; dg.hi[a] = grid->hi[a];
    dg.step[a] = grid->n[a] > 1 ? (grid->hi[a] - grid->lo[a]) / (double)(grid->n[a] - 1) : 0.0;   // np.linspace
    total *= (long double)grid->n[a];
  }
  for (int j = 0; j < nd; j++) dg.fixed[j] = grid->fixed[j];
  if (first < 0 || count < 0 || (long double)first + (long double)count > total) return fail(c, CL_E_INVALID, "slice [%lld, +%lld) outside the grid", (long long)first, (long long)count);
  const bool as_chi2 = moments || what == CL_OUT_CHI2;
  stats->best = as_chi2 ? INFINITY : -INFINITY; stats->index = -1; stats->log_sum = -INFINITY; stats->count = count;
  stats->larger_is_better = as_chi2 ? 0 : 1; stats->reserved = 0;
  if (count == 0) return CL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  GridStats gs;
  const bool out_pinned = out && is_page_locked(out);
  c->ev = c->evring[c->n_timed % cl_ctx::kRing];
  CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
  for (int64_t r0 = 0; r0 < count; r0 += c->max_rows) {
    const int64_t rows = std::min(c->max_rows, count - r0);
    int rc = ensure_rows(c, rows);
    if (rc != CL_OK) return rc;
    const int64_t blocks = (rows + 255) / 256;
    if (blocks > c->grid_part_cap) {
      CUDA_TRY(c, cudaStreamSynchronize(st));
      if (c->d_grid_part) cudaFree(c->d_grid_part);
      if (c->h_grid_part) cudaFreeHost(c->h_grid_part);
      c->d_grid_part = c->h_grid_part = nullptr; c->grid_part_cap = 0;
      const int64_t cap = std::max<int64_t>(blocks, (c->max_rows + 255) / 256);
      CUDA_TRY(c, cudaMalloc(&c->d_grid_part, cap * 4 * sizeof(double)));
      CUDA_TRY(c, cudaMallocHost(&c->h_grid_part, cap * 4 * sizeof(double)));
      c->grid_part_cap = cap;
    }
    if (out && !out_pinned) { rc = ensure_pinned(c, 0, rows); if (rc != CL_OK) return rc; }
    k_grid_theta<<<(unsigned)blocks, 256, 0, st>>>(dg, first + r0, rows, c->d_theta);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    // d_out holds 4 doubles per row: the engine's output in [0, 3 rows) and, when the caller wants them, the values behind it
    rc = run_pass(c, c->d_theta, rows, nd, moments ? CL_OUT_CHI2 : what, c->d_out, nullptr, moments, st, r0 == 0);
    if (rc != CL_OK) return rc;
    GridReduceArgs ra{};
    ra.src = c->d_out; ra.moments = moments ? (what == CL_GRID_MARGINAL ? 2 : 1) : 0; ra.as_chi2 = as_chi2 ? 1 : 0;
    ra.rows = rows; ra.first = first + r0; ra.vals = out ? c->d_out + 3 * rows : nullptr; ra.part = c->d_grid_part;
    k_grid_reduce<<<(unsigned)blocks, 256, 0, st>>>(ra);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(c->h_grid_part, c->d_grid_part, blocks * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out) CUDA_TRY(c, cudaMemcpyAsync(out_pinned ? out + r0 : c->h_out, ra.vals, rows * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (r0 + rows >= count) CUDA_TRY(c, cudaEventRecord(c->ev[5], st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    if (out && !out_pinned) memcpy(out + r0, c->h_out, rows * sizeof(double));
    for (int64_t b = 0; b < blocks; b++) grid_fold(gs, c->h_grid_part[4 * b], c->h_grid_part[4 * b + 1], c->h_grid_part[4 * b + 2], c->h_grid_part[4 * b + 3]);
  }
  c->n_timed++;
  stats->best = as_chi2 ? gs.best : -0.5 * gs.best;
  stats->index = gs.index >= 0.0 ? (int64_t)gs.index : -1;
  stats->log_sum = gs.sum > 0.0 ? gs.lmax + log(gs.sum) : -INFINITY;
  return CL_OK;
}

extern "C" int cl_grid_allreduce(cl_ctx* c, cl_grid_stats* stats) {
  if (!c || !stats) return CL_E_INVALID;
  if (!c->comm) return fail(c, CL_E_INVALID, "no communicator: call cl_comm_init first");
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int W = c->comm_size;
  int rc = ensure_scratch(c, (int64_t)W * 4 * sizeof(double));
  if (rc != CL_OK) return rc;
  rc = ensure_pinned(c, 0, (int64_t)W * 4);
  if (rc != CL_OK) return rc;
  double mine[4] = {stats->best, (double)stats->index, stats->log_sum, (double)stats->count};
  double* slot = c->d_scratch + 4 * c->comm_rank;
  CUDA_TRY(c, cudaMemcpyAsync(slot, mine, sizeof mine, cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));   // `mine` is a stack buffer
  NCCL_TRY(c, nccl_api().AllGather(slot, c->d_scratch, 4, kNcclFloat64, c->comm, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->h_out, c->d_scratch, (size_t)W * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  // fold in rank order: the same bits on every rank
  const bool larger = stats->larger_is_better != 0;
  double best = larger ? -INFINITY : INFINITY, index = -1.0, lmax = -INFINITY, sum = 0.0;
  int64_t count = 0;
  for (int r = 0; r < W; r++) {
    const double* p = c->h_out + 4 * r;
    const bool better = larger ? p[0] > best : p[0] < best;
    if (p[1] >= 0.0 && (better || (p[0] == best && (index < 0.0 || p[1] < index)))) { best = p[0]; index = p[1]; }
    if (p[2] > -INFINITY) {
      if (p[2] > lmax) { sum = sum * exp(lmax - p[2]) + 1.0; lmax = p[2]; }
      else sum += exp(p[2] - lmax);
    }
    count += (int64_t)p[3];
  }
  stats->best = best; stats->index = index >= 0.0 ? (int64_t)index : -1;
  stats->log_sum = sum > 0.0 ? lmax + log(sum) : -INFINITY;
  stats->count = count;
  return CL_OK;
}

// ---- proposals of a nested sampler: generate, evaluate, select on the device ----
extern "C" int cl_propose_eval(cl_ctx* c, const cl_proposal* prop, int64_t n, uint64_t seed, uint64_t offset, int what, double thresh,
                               int64_t max_keep, double* u_out, double* theta_out, double* val_out, int64_t counts[3]) {
  if (!c || !prop || !counts) return CL_E_INVALID;
  if (what < CL_OUT_CHI2 || what > CL_OUT_LOGPROB) return fail(c, CL_E_INVALID, "bad output selector");
  const int nd = c->ds.ndim;
  if (prop->ndim != nd) return fail(c, CL_E_INVALID, "proposal ndim %d != spec ndim %d", prop->ndim, nd);
  if (n < 1 || max_keep < 1 || !u_out || !theta_out || !val_out) return fail(c, CL_E_INVALID, "bad proposal arguments");
  if (n > c->max_rows) return fail(c, CL_E_INVALID, "n (%lld) exceeds max_rows_per_pass (%lld)", (long long)n, (long long)c->max_rows);
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  int rc = ensure_rows(c, n);
  if (rc != CL_OK) return rc;
  const int64_t blocks = (n + 255) / 256;
  if (n > c->prop_cap) {
    CUDA_TRY(c, cudaStreamSynchronize(st));
    for (void** p : {(void**)&c->d_prop_u, (void**)&c->d_prop_val, (void**)&c->d_prop_inside, (void**)&c->d_prop_cnt}) { if (*p) cudaFree(*p); *p = nullptr; }
    c->prop_cap = 0;
    const int64_t cap = std::max<int64_t>(n, std::min<int64_t>(c->max_rows, 65536));
    CUDA_TRY(c, cudaMalloc(&c->d_prop_u, cap * nd * sizeof(double)));
    CUDA_TRY(c, cudaMalloc(&c->d_prop_val, cap * sizeof(double)));
    CUDA_TRY(c, cudaMalloc(&c->d_prop_inside, cap));
    CUDA_TRY(c, cudaMalloc(&c->d_prop_cnt, (2 + (cap + 255) / 256) * sizeof(int)));
    c->prop_cap = cap;
  }
  if (max_keep > c->prop_keep_cap) {
    CUDA_TRY(c, cudaStreamSynchronize(st));
    if (c->d_prop_keep) cudaFree(c->d_prop_keep);
    c->d_prop_keep = nullptr; c->prop_keep_cap = 0;
    CUDA_TRY(c, cudaMalloc(&c->d_prop_keep, max_keep * (2 * nd + 1) * sizeof(double)));
    c->prop_keep_cap = max_keep;
  }
  DevProposal dp{};
  dp.ndim = nd;
  for (int j = 0; j < nd; j++) {
    dp.mu[j] = prop->mu[j]; dp.lo[j] = prop->lo[j]; dp.hi[j] = prop->hi[j]; dp.mean[j] = prop->mean[j]; dp.sigma[j] = prop->sigma[j];
    dp.gauss[j] = prop->gauss[j];
    for (int k = 0; k < nd; k++) dp.L[j * nd + k] = prop->L[j * nd + k];
  }
  c->ev = c->evring[c->n_timed % cl_ctx::kRing];
  CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
  k_propose<<<(unsigned)blocks, 256, 0, st>>>(dp, n, seed, offset, c->d_prop_u, c->d_theta, c->d_prop_inside);
  c->launches++;
  CUDA_TRY(c, cudaGetLastError());
  rc = run_pass(c, c->d_theta, n, nd, what, c->d_prop_val, nullptr, false, st, true);
  if (rc != CL_OK) return rc;
  CUDA_TRY(c, cudaMemsetAsync(c->d_prop_cnt, 0, 2 * sizeof(int), st));
  k_select_count<<<(unsigned)blocks, 256, 0, st>>>(c->d_prop_val, c->d_prop_inside, thresh, n, c->d_prop_cnt + 2, c->d_prop_cnt + 1);
  k_select_scan<<<1, 1024, 0, st>>>(c->d_prop_cnt + 2, (int)blocks, c->d_prop_cnt);
  double* keep_u = c->d_prop_keep; double* keep_t = keep_u + max_keep * nd; double* keep_v = keep_t + max_keep * nd;
  k_select_scatter<<<(unsigned)blocks, 256, 0, st>>>(c->d_prop_val, c->d_prop_inside, thresh, n, nd, c->d_prop_cnt + 2, c->d_prop_u, c->d_theta,
                                                     max_keep, keep_u, keep_t, keep_v);
  c->launches += 3;
  CUDA_TRY(c, cudaGetLastError());
  int h_cnt[2] = {0, 0};
  CUDA_TRY(c, cudaMemcpyAsync(h_cnt, c->d_prop_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  const int64_t kept = std::min<int64_t>(h_cnt[0], max_keep);
  counts[0] = h_cnt[1]; counts[1] = h_cnt[0]; counts[2] = kept;
  if (kept > 0) {   // the kept rows are few (the sampler asks for what it still needs): plain copies
    CUDA_TRY(c, cudaMemcpyAsync(u_out, keep_u, kept * nd * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(theta_out, keep_t, kept * nd * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(val_out, keep_v, kept * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(c, cudaEventRecord(c->ev[5], st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  c->n_timed++;
  return CL_OK;
}

// ---- helper exports: one stage-1/2 launch in a non-EVAL mode, results through pinned staging ----
static int helper(cl_ctx* c, const double* theta, int64_t B, int64_t ld, int mode, const double* zq, int64_t nq, double* o1, double* o2, int64_t width) {
  int rc = check_eval_args(c, theta, B, ld, o1 ? o1 : o2);
  if (rc != CL_OK || B == 0) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int nd = c->ds.ndim;
  cudaStream_t st = c->stream;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, (64LL << 20) / std::max<int64_t>(8, width * 16)));
  std::vector<double> hbuf;
  for (int64_t r0 = 0; r0 < B; r0 += chunk) {
    int64_t rows = std::min(chunk, B - r0);
    int64_t per = rows * width;
    rc = ensure_scratch(c, (rows * nd + 2 * per + nq) * (int64_t)sizeof(double));
    if (rc != CL_OK) return rc;
    double* d_th = c->d_scratch;
    double* d_o1 = d_th + rows * nd;
    double* d_o2 = d_o1 + per;
    double* d_zq = d_o2 + per;
    hbuf.resize(rows * nd);
    for (int64_t i = 0; i < rows; i++) memcpy(hbuf.data() + i * nd, theta + (r0 + i) * ld, nd * sizeof(double));
    CUDA_TRY(c, cudaMemcpyAsync(d_th, hbuf.data(), rows * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nq > 0) CUDA_TRY(c, cudaMemcpyAsync(d_zq, zq, nq * sizeof(double), cudaMemcpyHostToDevice, st));
    Stage12Args a{};
    a.theta = d_th; a.B = rows; a.ld = nd; a.mode = mode; a.what = CL_OUT_CHI2;
    a.zq = d_zq; a.nq = (int)nq;
    if (mode == MODE_DIST) { a.outDM = o1 ? d_o1 : nullptr; a.outDH = o2 ? d_o2 : nullptr; }
    else if (mode == MODE_RESID) { a.R = d_o1; }
    else a.out = d_o1;
    rc = launch_s12(c, a, st);
    if (rc != CL_OK) return rc;
    if (o1) CUDA_TRY(c, cudaMemcpyAsync(o1 + r0 * width, d_o1, per * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (o2) CUDA_TRY(c, cudaMemcpyAsync(o2 + r0 * width, d_o2, per * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
  }
  return CL_OK;
}

extern "C" int cl_distances(cl_ctx* c, const double* theta, int64_t B, int64_t ld, const double* zq, int64_t nq, double* DM, double* DH) {
  if (!c) return CL_E_INVALID;
  if (!c->ds.z_grid) return fail(c, CL_E_INVALID, "spec has no z_grid");
  if (nq <= 0 || !zq || (!DM && !DH)) return fail(c, CL_E_INVALID, "bad query arguments");
  return helper(c, theta, B, ld, MODE_DIST, zq, nq, DM, DH, nq);
}
extern "C" int cl_bao_theory(cl_ctx* c, const double* theta, int64_t B, int64_t ld, double* out) {
  if (!c) return CL_E_INVALID;
  if (c->ds.n_bao <= 0) return fail(c, CL_E_INVALID, "spec has no BAO block");
  return helper(c, theta, B, ld, MODE_BAO, nullptr, 0, out, nullptr, c->ds.n_bao);
}
extern "C" int cl_cmb(cl_ctx* c, const double* theta, int64_t B, int64_t ld, double* out) {
  if (!c) return CL_E_INVALID;
  if (c->ds.family != CL_FAMILY_FULL) return fail(c, CL_E_INVALID, "CMB quantities need the FULL family");
  return helper(c, theta, B, ld, MODE_CMB, nullptr, 0, out, nullptr, 8);
}
extern "C" int cl_sn_residuals(cl_ctx* c, const double* theta, int64_t B, int64_t ld, double* out) {
  if (!c) return CL_E_INVALID;
  if (c->ds.n_sn <= 0) return fail(c, CL_E_INVALID, "spec has no SN block");
  return helper(c, theta, B, ld, MODE_RESID, nullptr, 0, out, nullptr, c->ds.n_sn);
}

static int timing_of(cl_ctx* c, int64_t idx, double ms[4]) {
  cudaEvent_t* ev = c->evring[idx % cl_ctx::kRing];
  CUDA_TRY(c, cudaEventSynchronize(ev[5]));
  float t;
  CUDA_TRY(c, cudaEventElapsedTime(&t, ev[1], ev[2])); ms[0] = t;
  CUDA_TRY(c, cudaEventElapsedTime(&t, ev[2], ev[3])); ms[1] = t;
  CUDA_TRY(c, cudaEventElapsedTime(&t, ev[3], ev[4])); ms[2] = t;
  CUDA_TRY(c, cudaEventElapsedTime(&t, ev[0], ev[5])); ms[3] = t;
  return CL_OK;
}

extern "C" int cl_last_timing(cl_ctx* c, double ms[4]) {
  if (!c || !ms) return CL_E_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->n_timed == 0) return fail(c, CL_E_INVALID, "no evaluation has been timed yet");
  return timing_of(c, c->n_timed - 1, ms);
}

extern "C" int cl_stage3_split(cl_ctx* c, int n, double* ms) {
  if (!c || !ms || n < 0) return CL_E_INVALID;
  if (!c->d_W) return fail(c, CL_E_INVALID, "the spec has no large SN block");
  std::lock_guard<std::mutex> lk(c->mu);
  int avail = (int)std::min<int64_t>(c->n_timed, cl_ctx::kRing);
  int k = std::min(n, avail);
  for (int i = 0; i < k; i++) {
    cudaEvent_t* ev = c->evring[(c->n_timed - k + i) % cl_ctx::kRing];
    CUDA_TRY(c, cudaEventSynchronize(ev[5]));
    float t;
    CUDA_TRY(c, cudaEventElapsedTime(&t, ev[2], ev[6])); ms[2 * i] = t;
    CUDA_TRY(c, cudaEventElapsedTime(&t, ev[6], ev[3])); ms[2 * i + 1] = t;
  }
  return k;
}

extern "C" int cl_timing_history(cl_ctx* c, int n, double* ms) {
  if (!c || !ms || n < 0) return CL_E_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  int avail = (int)std::min<int64_t>(c->n_timed, cl_ctx::kRing);
  int k = std::min(n, avail);
  for (int i = 0; i < k; i++) {
    int rc = timing_of(c, c->n_timed - k + i, ms + 4 * i);
    if (rc != CL_OK) return rc;
  }
  return k;
}
