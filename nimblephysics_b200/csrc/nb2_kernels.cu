// libnb2.so — kernels + C ABI (include/nb2.h).  sm_100a only.
//
// Kernel shape (contact-free step): ONE THREAD PER WORLD, 32 worlds per warp.  The model is a __grid_constant__
// kernel parameter (constant-bank, warp-uniform loads); per-world working storage lives in dynamic shared memory,
// interleaved [word][lane] so every access is bank-conflict free; the only HBM traffic is the fp32 state/action
// rows, the outputs and the saved-for-backward stream ([word][B]: coalesced).  All branches depend on the model
// only, so warps never diverge.  See DESIGN.md for the roofline discussion (the path is FP32-latency bound).
#include <cuda_runtime.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/nb2.h"
#include "nb2_dyn.cuh"
#include "nb2_contact.cuh"
#include "nb2_host_model.h"

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

#define NB2_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
      return NB2_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

namespace {

template <class R>
__global__ void __launch_bounds__(128)
k_step_fwd(const __grid_constant__ Nb2ModelDev<R> M, int B, const float* __restrict__ state,
           const float* __restrict__ action, float* __restrict__ next, R* __restrict__ saved, int words) {
  extern __shared__ __align__(16) unsigned char nb2_smem[];
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= B) return;
  R* scr = reinterpret_cast<R*>(nb2_smem) + (size_t)(threadIdx.x >> 5) * words * 32 + (threadIdx.x & 31);
  nb2::world_forward<R, 32>(M, scr, state + (size_t)w * 2 * M.ndof, action + (size_t)w * M.na,
                            next + (size_t)w * 2 * M.ndof, saved ? saved + w : nullptr, (size_t)B, saved != nullptr);
}

template <class R>
__global__ void __launch_bounds__(128)
k_step_bwd(const __grid_constant__ Nb2ModelDev<R> M, int B, const float* __restrict__ state,
           const float* __restrict__ action, const R* __restrict__ saved, const float* __restrict__ gnext,
           float* __restrict__ gstate, float* __restrict__ gaction, int words) {
  extern __shared__ __align__(16) unsigned char nb2_smem[];
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= B) return;
  R* scr = reinterpret_cast<R*>(nb2_smem) + (size_t)(threadIdx.x >> 5) * words * 32 + (threadIdx.x & 31);
  nb2::world_backward<R, 32>(M, scr, state + (size_t)w * 2 * M.ndof, action + (size_t)w * M.na,
                             gnext + (size_t)w * 2 * M.ndof, saved + w, (size_t)B,
                             gstate + (size_t)w * 2 * M.ndof, gaction + (size_t)w * M.na);
}

// contact / boxed-LCP stage: one thread per world, fp64, per-world workspace in global memory (L1/L2 cached).
// The pivoting LCP solve is data dependent, so lanes of a warp diverge here by construction.
__global__ void __launch_bounds__(64)
k_contact_fwd(const __grid_constant__ Nb2ModelDev<double> M, const __grid_constant__ Nb2ContactDev C, int B,
              const float* __restrict__ state, float* __restrict__ next, const double* __restrict__ saved,
              double* __restrict__ workspace, size_t ws_doubles, double* __restrict__ x_lcp, int* __restrict__ m_lcp,
              int* __restrict__ labels, int* __restrict__ status, int* __restrict__ ncontacts, float* __restrict__ cinfo,
              double* __restrict__ crec, size_t rec_doubles) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= B) return;
  // the 32 worlds of a warp share one lane-interleaved workspace block
  nb2::world_contact<32>(M, C, state + (size_t)w * 2 * M.ndof, next + (size_t)w * 2 * M.ndof, saved + w, (size_t)B,
                     workspace + (size_t)(w >> 5) * ws_doubles * 32, w & 31, x_lcp + (size_t)w * NB2_MAX_ROWS, m_lcp + w,
                     labels + (size_t)w * NB2_MAX_ROWS, status + w, ncontacts + w,
                     cinfo ? cinfo + (size_t)w * NB2_MAX_CONTACTS * 10 : nullptr, crec ? crec + (size_t)w * rec_doubles : nullptr);
}

// backward of a step with the contact stage (fp64): world_backward<double, 32, CONTACT=true>
__global__ void __launch_bounds__(32)
k_step_bwd_contact(const __grid_constant__ Nb2ModelDev<double> M, const __grid_constant__ Nb2ContactDev C, int B,
                   const float* __restrict__ state, const float* __restrict__ action, const double* __restrict__ saved,
                   const double* __restrict__ crec, size_t rec_doubles, double* __restrict__ workspace, size_t ws_doubles,
                   const float* __restrict__ gnext, float* __restrict__ gstate, float* __restrict__ gaction, int words) {
  extern __shared__ __align__(16) unsigned char nb2_smem[];
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= B) return;
  double* scr = reinterpret_cast<double*>(nb2_smem) + (size_t)(threadIdx.x >> 5) * words * 32 + (threadIdx.x & 31);
  nb2::BwdContactHook H;
  H.model_contact = &C; H.ws = workspace + (size_t)(w >> 5) * ws_doubles * 32; H.lane = w & 31; H.crec = crec + (size_t)w * rec_doubles;
  nb2::world_backward<double, 32, true>(M, scr, state + (size_t)w * 2 * M.ndof, action + (size_t)w * M.na,
                                        gnext + (size_t)w * 2 * M.ndof, saved + w, (size_t)B,
                                        gstate + (size_t)w * 2 * M.ndof, gaction + (size_t)w * M.na, &H);
}

constexpr int kMaxSmem = 227 * 1024;

// warps per block: spread small batches over all SMs first (1 warp per block), pack up to 4 warps per block
// once there are more warps than SMs can hold singly; always bounded by the shared-memory budget.
int pick_warps(int B, size_t bytes_per_warp, int sm_count) {
  const int total_warps = (B + 31) / 32;
  int fit = (int)(kMaxSmem / bytes_per_warp);
  if (fit < 1) return 0;
  int want = (total_warps <= 2 * sm_count) ? 1 : 4;
  if (want > fit) want = fit;
  if (want > 4) want = 4;
  return want;
}

}  // namespace

struct nb2_model {
  Nb2ModelDev<float> mf;
  Nb2ModelDev<double> md;
  Nb2ContactDev contact;
  bool has_contacts = false;
  int fwd_words, bwd_words, saved_words;
  int sm_count;
  bool attr_set[4] = {false, false, false, false};
  // device buffers owned by the *_host entry points
  float *d_state = nullptr, *d_action = nullptr, *d_next = nullptr, *d_gnext = nullptr,
        *d_gstate = nullptr, *d_gaction = nullptr;
  void* d_saved = nullptr;  // sized for fp64 words
  int host_cap = 0;
  int host_B = 0;  // batch of the last forward_host kept for backward
  cudaStream_t host_stream = nullptr;
  std::mutex mu;
};

template <class R>
static int launch_fwd(nb2_model* m, const Nb2ModelDev<R>& M, int B, const float* state, const float* action,
                      float* next, R* saved, cudaStream_t st, int attr_idx) {
  const size_t per_warp = (size_t)m->fwd_words * 32 * sizeof(R);
  const int warps = pick_warps(B, per_warp, m->sm_count);
  if (warps == 0) { g_err = "model needs " + std::to_string(per_warp) + " B of shared memory per warp (> 227 KB)"; return NB2_ERR_UNSUPPORTED; }
  if (!m->attr_set[attr_idx]) {
    NB2_CUDA(cudaFuncSetAttribute(k_step_fwd<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    m->attr_set[attr_idx] = true;
  }
  const int threads = warps * 32;
  const int blocks = (B + threads - 1) / threads;
  k_step_fwd<R><<<blocks, threads, per_warp * warps, st>>>(M, B, state, action, next, saved, m->fwd_words);
  g_launches++;
  NB2_CUDA(cudaGetLastError());
  return NB2_OK;
}
template <class R>
static int launch_bwd(nb2_model* m, const Nb2ModelDev<R>& M, int B, const float* state, const float* action,
                      const R* saved, const float* gnext, float* gstate, float* gaction, cudaStream_t st,
                      int attr_idx) {
  const size_t per_warp = (size_t)m->bwd_words * 32 * sizeof(R);
  const int warps = pick_warps(B, per_warp, m->sm_count);
  if (warps == 0) { g_err = "model needs " + std::to_string(per_warp) + " B of shared memory per warp (> 227 KB)"; return NB2_ERR_UNSUPPORTED; }
  if (!m->attr_set[attr_idx]) {
    NB2_CUDA(cudaFuncSetAttribute(k_step_bwd<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    m->attr_set[attr_idx] = true;
  }
  const int threads = warps * 32;
  const int blocks = (B + threads - 1) / threads;
  k_step_bwd<R><<<blocks, threads, per_warp * warps, st>>>(M, B, state, action, saved, gnext, gstate, gaction, m->bwd_words);
  g_launches++;
  NB2_CUDA(cudaGetLastError());
  return NB2_OK;
}

extern "C" {

const char* nb2_last_error(void) { return g_err.c_str(); }
const char* nb2_version(void) { return "nb2 0.1 (sm_100a, thread-per-world ABA + adjoint)"; }
long long nb2_launch_count(void) { return g_launches.load(); }

int nb2_model_create(const nb2_model_desc* desc, nb2_model** out) {
  if (!desc || !out) { g_err = "null argument"; return NB2_ERR_INVALID; }
  nb2_model* m = new nb2_model();
  std::string err;
  if (!nb2_fill_model(*desc, m->mf, err) || !nb2_fill_model(*desc, m->md, err)) {
    g_err = err; delete m;
    return (desc->nb > NB2_MAX_BODIES || desc->ndof > NB2_MAX_DOFS) ? NB2_ERR_UNSUPPORTED : NB2_ERR_INVALID;
  }
  if (desc->nshapes > 0 && desc->npairs > 0) {
    if (!nb2_fill_contact(*desc, m->contact, err)) { g_err = err; delete m; return NB2_ERR_UNSUPPORTED; }
    m->has_contacts = true;
  }
  m->fwd_words = nb2::fwd_layout(m->mf.nb, m->mf.ndof, m->mf.nslots, m->mf.nfree).total;
  m->bwd_words = nb2::bwd_layout(m->mf.nb, m->mf.ndof, m->mf.nslots, m->mf.nfree).total;
  m->saved_words = nb2_saved_words(m->mf.nb, m->mf.ndof, m->mf.nfree);
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    g_err = "no CUDA device available: nimblephysics_b200 has no CPU fallback";
    delete m;
    return NB2_ERR_CUDA;
  }
  m->sm_count = prop.multiProcessorCount;
  *out = m;
  return NB2_OK;
}

void nb2_model_destroy(nb2_model* m) {
  if (!m) return;
  cudaFree(m->d_state); cudaFree(m->d_action); cudaFree(m->d_next); cudaFree(m->d_saved);
  cudaFree(m->d_gnext); cudaFree(m->d_gstate); cudaFree(m->d_gaction);
  if (m->host_stream) cudaStreamDestroy(m->host_stream);
  delete m;
}
int nb2_model_has_contacts(const nb2_model* m) { return (m && m->has_contacts) ? 1 : 0; }
size_t nb2_contact_workspace_bytes(const nb2_model* m, int B) {
  if (!m || !m->has_contacts || B <= 0) return 0;
  return nb2::contact_ws_doubles(m->mf.nb, m->mf.ndof) * sizeof(double) * (size_t)((B + 31) / 32) * 32;
}
int nb2_step_forward_contact(const nb2_model* cm, int B, const float* state, const float* action, float* next_state,
                             void* saved_fp64, void* workspace, double* x_lcp, int32_t* m_lcp, int32_t* labels,
                             int32_t* status, int32_t* ncontacts, float* cinfo, double* contact_record, void* stream) {
  nb2_model* m = const_cast<nb2_model*>(cm);
  if (!m || B < 0 || !state || !action || !next_state || !saved_fp64 || !workspace || !x_lcp || !m_lcp || !labels || !status || !ncontacts) {
    g_err = "nb2_step_forward_contact: bad argument"; return NB2_ERR_INVALID;
  }
  if (!m->has_contacts) { g_err = "nb2_step_forward_contact: the model has no collision pairs"; return NB2_ERR_INVALID; }
  if (B == 0) return NB2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_fwd<double>(m, m->md, B, state, action, next_state, (double*)saved_fp64, st, 1);
  if (rc) return rc;
  const int threads = 32;
  k_contact_fwd<<<(B + threads - 1) / threads, threads, 0, st>>>(m->md, m->contact, B, state, next_state, (const double*)saved_fp64,
                                                                 (double*)workspace, nb2::contact_ws_doubles(m->mf.nb, m->mf.ndof), x_lcp,
                                                                 m_lcp, labels, status, ncontacts, cinfo, contact_record,
                                                                 nb2::contact_rec_doubles(m->mf.ndof));
  g_launches++;
  NB2_CUDA(cudaGetLastError());
  return NB2_OK;
}
size_t nb2_contact_record_bytes(const nb2_model* m, int B) {
  if (!m || !m->has_contacts || B <= 0) return 0;
  return nb2::contact_rec_doubles(m->mf.ndof) * sizeof(double) * (size_t)B;
}
int nb2_step_backward_contact(const nb2_model* cm, int B, const float* state, const float* action, const void* saved_fp64,
                              const double* contact_record, void* workspace, const float* grad_next_state, float* grad_state,
                              float* grad_action, void* stream) {
  nb2_model* m = const_cast<nb2_model*>(cm);
  if (!m || B < 0 || !state || !action || !saved_fp64 || !contact_record || !workspace || !grad_next_state || !grad_state || !grad_action) {
    g_err = "nb2_step_backward_contact: bad argument"; return NB2_ERR_INVALID;
  }
  if (!m->has_contacts) { g_err = "nb2_step_backward_contact: the model has no collision pairs"; return NB2_ERR_INVALID; }
  if (B == 0) return NB2_OK;
  const int words = nb2::bwd_layout(m->mf.nb, m->mf.ndof, m->mf.nslots, m->mf.nfree, 42).total;
  const size_t smem = (size_t)words * 32 * sizeof(double);
  if (smem > (size_t)kMaxSmem) { g_err = "model needs " + std::to_string(smem) + " B of shared memory per warp (> 227 KB)"; return NB2_ERR_UNSUPPORTED; }
  static bool attr_done = false;
  if (!attr_done) { NB2_CUDA(cudaFuncSetAttribute(k_step_bwd_contact, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)); attr_done = true; }
  k_step_bwd_contact<<<(B + 31) / 32, 32, smem, (cudaStream_t)stream>>>(m->md, m->contact, B, state, action, (const double*)saved_fp64,
                                                                         contact_record, nb2::contact_rec_doubles(m->mf.ndof), (double*)workspace,
                                                                         nb2::contact_ws_doubles(m->mf.nb, m->mf.ndof), grad_next_state,
                                                                         grad_state, grad_action, words);
  g_launches++;
  NB2_CUDA(cudaGetLastError());
  return NB2_OK;
}
int nb2_model_ndof(const nb2_model* m) { return m ? m->mf.ndof : -1; }
int nb2_model_na(const nb2_model* m) { return m ? m->mf.na : -1; }
int nb2_saved_words_per_world(const nb2_model* m) { return m ? m->saved_words : -1; }

int nb2_step_forward(const nb2_model* cm, int B, const float* state, const float* action, float* next_state,
                     void* saved, int precision, void* stream) {
  nb2_model* m = const_cast<nb2_model*>(cm);
  if (!m || B < 0 || !state || !action || !next_state) { g_err = "nb2_step_forward: bad argument"; return NB2_ERR_INVALID; }
  if (B == 0) return NB2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == NB2_FP64) return launch_fwd<double>(m, m->md, B, state, action, next_state, (double*)saved, st, 1);
  return launch_fwd<float>(m, m->mf, B, state, action, next_state, (float*)saved, st, 0);
}

int nb2_step_backward(const nb2_model* cm, int B, const float* state, const float* action, const void* saved,
                      const float* grad_next_state, float* grad_state, float* grad_action, int precision,
                      void* stream) {
  nb2_model* m = const_cast<nb2_model*>(cm);
  if (!m || B < 0 || !state || !action || !saved || !grad_next_state || !grad_state || !grad_action) {
    g_err = "nb2_step_backward: bad argument"; return NB2_ERR_INVALID;
  }
  if (B == 0) return NB2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == NB2_FP64) return launch_bwd<double>(m, m->md, B, state, action, (const double*)saved, grad_next_state, grad_state, grad_action, st, 3);
  return launch_bwd<float>(m, m->mf, B, state, action, (const float*)saved, grad_next_state, grad_state, grad_action, st, 2);
}

static int ensure_host_buffers(nb2_model* m, int B) {
  if (!m->host_stream) NB2_CUDA(cudaStreamCreateWithFlags(&m->host_stream, cudaStreamNonBlocking));
  if (B <= m->host_cap) return NB2_OK;
  cudaFree(m->d_state); cudaFree(m->d_action); cudaFree(m->d_next); cudaFree(m->d_saved);
  cudaFree(m->d_gnext); cudaFree(m->d_gstate); cudaFree(m->d_gaction);
  m->d_state = m->d_action = m->d_next = m->d_gnext = m->d_gstate = m->d_gaction = nullptr;
  m->d_saved = nullptr;
  m->host_cap = 0;
  const size_t n2 = (size_t)2 * m->mf.ndof, na = (size_t)(m->mf.na > 0 ? m->mf.na : 1);
  NB2_CUDA(cudaMalloc(&m->d_state, n2 * B * sizeof(float)));
  NB2_CUDA(cudaMalloc(&m->d_action, na * B * sizeof(float)));
  NB2_CUDA(cudaMalloc(&m->d_next, n2 * B * sizeof(float)));
  NB2_CUDA(cudaMalloc(&m->d_saved, (size_t)m->saved_words * B * sizeof(double)));
  NB2_CUDA(cudaMalloc(&m->d_gnext, n2 * B * sizeof(float)));
  NB2_CUDA(cudaMalloc(&m->d_gstate, n2 * B * sizeof(float)));
  NB2_CUDA(cudaMalloc(&m->d_gaction, na * B * sizeof(float)));
  m->host_cap = B;
  return NB2_OK;
}

int nb2_step_forward_host(nb2_model* m, int B, const float* state, const float* action, float* next_state,
                          int keep_for_backward, int precision) {
  if (!m || B <= 0 || !state || !action || !next_state) { g_err = "nb2_step_forward_host: bad argument"; return NB2_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(m->mu);
  int rc = ensure_host_buffers(m, B);
  if (rc) return rc;
  const size_t n2 = (size_t)2 * m->mf.ndof, na = (size_t)m->mf.na;
  cudaStream_t st = m->host_stream;
  NB2_CUDA(cudaMemcpyAsync(m->d_state, state, n2 * B * sizeof(float), cudaMemcpyHostToDevice, st));
  NB2_CUDA(cudaMemcpyAsync(m->d_action, action, na * B * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = nb2_step_forward(m, B, m->d_state, m->d_action, m->d_next, keep_for_backward ? m->d_saved : nullptr, precision, st);
  if (rc) return rc;
  NB2_CUDA(cudaMemcpyAsync(next_state, m->d_next, n2 * B * sizeof(float), cudaMemcpyDeviceToHost, st));
  NB2_CUDA(cudaStreamSynchronize(st));
  m->host_B = keep_for_backward ? B : 0;
  return NB2_OK;
}

int nb2_step_backward_host(nb2_model* m, int B, const float* grad_next_state, float* grad_state, float* grad_action,
                           int precision) {
  if (!m || !grad_next_state || !grad_state || !grad_action) { g_err = "nb2_step_backward_host: bad argument"; return NB2_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(m->mu);
  if (B <= 0 || B != m->host_B) { g_err = "nb2_step_backward_host: no matching forward_host(keep_for_backward=1) precedes this call"; return NB2_ERR_INVALID; }
  const size_t n2 = (size_t)2 * m->mf.ndof, na = (size_t)m->mf.na;
  cudaStream_t st = m->host_stream;
  NB2_CUDA(cudaMemcpyAsync(m->d_gnext, grad_next_state, n2 * B * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = nb2_step_backward(m, B, m->d_state, m->d_action, m->d_saved, m->d_gnext, m->d_gstate, m->d_gaction, precision, st);
  if (rc) return rc;
  NB2_CUDA(cudaMemcpyAsync(grad_state, m->d_gstate, n2 * B * sizeof(float), cudaMemcpyDeviceToHost, st));
  NB2_CUDA(cudaMemcpyAsync(grad_action, m->d_gaction, na * B * sizeof(float), cudaMemcpyDeviceToHost, st));
  NB2_CUDA(cudaStreamSynchronize(st));
  return NB2_OK;
}

}  // extern "C"
