// C ABI of libcolvo_b200.so (include/colvo.h).  Host-side only: descriptor validation,
// workspace carving and kernel launches on the caller's stream.  No allocation, no
// synchronisation, no retained state; there is no CPU fallback.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/colvo.h"
#include "colvo_kernels.cuh"

using namespace colvo;

namespace colvo {
KernelTimer g_timer = {0, nullptr, nullptr};

cudaError_t ensure_dyn_smem(const void* kernel, size_t bytes) {
  // (kernel, device) pairs already opted in; a handful of entries, appended under a lock, read lock-free
  struct Entry { const void* fn; int dev; size_t bytes; };
  static Entry table[64];
  static std::atomic<int> count{0};
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const int n = count.load(std::memory_order_acquire);
  for (int i = 0; i < n; ++i)
    if (table[i].fn == kernel && table[i].dev == dev && table[i].bytes >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  const int m = count.load(std::memory_order_relaxed);
  if (m < 64) {
    table[m] = Entry{kernel, dev, bytes};
    count.store(m + 1, std::memory_order_release);
  }
  return cudaSuccess;
}

cudaError_t make_tensor_map_3d(CUtensorMap* tm, const void* base, const unsigned long long (&dims)[3],
                               const unsigned long long (&strides_bytes)[2], const unsigned (&box)[3]) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;      // resolved once per process (idempotent, so a race only repeats the lookup)
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t gd[3] = {dims[0], dims[1], dims[2]}, gs[2] = {strides_bytes[0], strides_bytes[1]};
  const cuuint32_t bx[3] = {box[0], box[1], box[2]}, es[3] = {1, 1, 1};
  auto call = [&]() {
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = call();
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    // The encoder is a DRIVER entry point: it needs the device's primary context current on the calling thread, which the
    // runtime only guarantees after its first call there -- autograd runs the backward on its own thread, and with
    // COLVO_F_NO_SRC_GRAD this is the first CUDA call of that thread.  Bind the context through the runtime and retry.
    cudaFree(nullptr);
    r = call();
  }
  if (r != CUDA_SUCCESS && getenv("COLVO_DEBUG"))      // diagnostics on request only: the library itself never prints
    fprintf(stderr, "colvo: cuTensorMapEncodeTiled -> %d  base %p dims %llu %llu %llu strides %llu %llu box %u %u %u\n", (int)r, base,
            dims[0], dims[1], dims[2], strides_bytes[0], strides_bytes[1], box[0], box[1], box[2]);
  return r == CUDA_SUCCESS ? cudaSuccess : static_cast<cudaError_t>(COLVO_E_TENSOR_MAP);
}
}

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline int div_up(int a, int b) { return (a + b - 1) / b; }

struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off = align_up(off + count * sizeof(T));
    return p;
  }
};

int check_desc(const ColvoDesc* d) {
  if (!d) return COLVO_E_NULL_PTR;
  if (d->B < 1 || d->N < 1 || d->N > COLVO_MAX_SOURCES || d->S < 1 || d->S > COLVO_MAX_SCALES) return COLVO_E_BAD_DESC;
  if (d->H < 2 || d->W < 2) return COLVO_E_BAD_DESC;
  if ((long long)d->H * d->W > (1ll << 26)) return COLVO_E_BAD_DESC;   // 9 * HW element offsets stay 32-bit
  if (d->B > 65535 || (long long)d->B * d->S > 65535) return COLVO_E_BAD_DESC;   // (b, k) pairs ride on grid.y
  for (int k = 0; k < d->S; ++k) {
    if (d->h[k] != (d->H >> k) || d->w[k] != (d->W >> k)) return COLVO_E_BAD_DESC;
    if (d->h[k] < 1 || d->w[k] < 1) return COLVO_E_BAD_DESC;
  }
  return 0;
}

void fill_params(KP& P, const ColvoDesc* d) {
  memset(&P, 0, sizeof(P));
  P.B = d->B; P.N = d->N; P.S = d->S; P.H = d->H; P.W = d->W; P.HW = d->H * d->W;
  for (int k = 0; k < d->S; ++k) {
    P.h[k] = d->h[k];
    P.w[k] = d->w[k];
    P.ry[k] = (float)((double)d->h[k] / (double)d->H);
    P.rx[k] = (float)((double)d->w[k] / (double)d->W);
    P.depth_bs[k] = (long long)d->h[k] * d->w[k];
  }
  P.alpha = d->alpha; P.c1 = d->c1; P.c2 = d->c2; P.eps_proj = d->eps_proj; P.eps_lcc = d->eps_lcc;
  P.eps_disp = d->eps_disp; P.z_min = d->z_min; P.smooth_weight = d->smooth_weight;
  P.geo_weight = d->geo_weight;
  P.flags = d->flags;
  P.tgt_bf = 1;
  P.src_nf = 1;
  P.src_bf = d->N;
  P.frame_el = (d->flags & COLVO_F_PACKED_BF16) ? (long long)P.HW : 3ll * P.HW;
  P.K_bs = 9; P.T_ns = 16; P.T_bs = 16 * d->N;
  P.tiles_x = div_up(d->W, kTileW);
  P.tiles_y = div_up(d->H, kTileH);
  P.btiles_y = div_up(d->H, kBwdTileH);
  P.ftiles_x = div_up(d->W, 32);
  P.ftiles_y = div_up(d->H, kFwdTileH);
  P.sm_blocks = div_up(d->W, kSmBW) * div_up(d->H, kSmBH);
  for (int k = 0; k < d->S; ++k) {
    const double lam = (double)d->smooth_weight / (double)(1 << k) / (double)d->S;
    const double nx = (double)d->B * d->h[k] * (d->w[k] - 1), ny = (double)d->B * (d->h[k] - 1) * d->w[k];
    P.sm_cx[k] = nx > 0 ? (float)(lam / nx) : 0.f;
    P.sm_cy[k] = ny > 0 ? (float)(lam / ny) : 0.f;
  }
}

size_t smooth_tiles(const ColvoDesc* d) { return (size_t)div_up(d->W, kSmBW) * div_up(d->H, kSmBH) * d->S; }

int stat_chunks(const ColvoDesc* d) { return div_up(d->H * d->W, stats_pixels_per_cta(d->S)); }

size_t carve_fwd(const ColvoDesc* d, void* ws, FwdBuffers& F) {
  Carver c(ws);
  const size_t BNS = (size_t)d->B * d->N * d->S;
  const size_t tiles = (size_t)div_up(d->W, kTileW) * div_up(d->H, kTileH);
  F.stat_chunks = stat_chunks(d);
  F.stat_part = c.take<double>(BNS * F.stat_chunks * kStatVals);
  F.smooth_part = c.take<double>((size_t)d->B * smooth_tiles(d) * kSmVals);
  F.loss_part = c.take<double>((size_t)d->B * tiles);
  F.g_part = c.take<double>((size_t)d->B * tiles * d->N * kMaxS * 2);
  F.iw = c.take<float4>(BNS * (size_t)d->H * d->W);
  return c.off;
}

size_t carve_bwd(const ColvoDesc* d, void* ws, BwdBuffers& Bw) {
  Carver c(ws);
  const size_t tiles = (size_t)div_up(d->W, kTileW) * div_up(d->H, kBwdTileH);
  const size_t HW = (size_t)d->H * d->W;
  Bw.dDhat[0] = nullptr;
  for (int k = 1; k < kMaxS; ++k) Bw.dDhat[k] = (k < d->S) ? c.take<float>((size_t)d->B * HW) : nullptr;
  Bw.pose_part = c.take<double>((size_t)d->B * tiles * d->N * 12);
  Bw.gsrc4 = (d->flags & COLVO_F_NO_SRC_GRAD) ? nullptr : c.take<float4>((size_t)d->B * d->N * HW);
  return c.off;
}

// the caller-owned `saved` buffer: doubles first, then the fp32 smoothness adjoint fields
size_t carve_saved(const ColvoDesc* d, double* saved, SavedView& sv) {
  const size_t BNS = (size_t)d->B * d->N * d->S, BS = (size_t)d->B * d->S;
  size_t nd = BNS * kSavedPerFrame + BS * kSavedPerScale;
  sv.frame = saved;
  sv.scale = saved ? saved + BNS * kSavedPerFrame : nullptr;
  float* f = saved ? reinterpret_cast<float*>(saved + nd) : nullptr;
  size_t nf = 0;
  for (int k = 0; k < kMaxS; ++k) {
    if (k < d->S) {
      sv.s_field[k] = f ? f + nf : nullptr;
      nf += (size_t)d->B * d->h[k] * d->w[k];
    } else {
      sv.s_field[k] = nullptr;
    }
  }
  nf = (nf + 3) / 4 * 4;                       // 16-byte texels (the doubles in front keep the base 16 B aligned)
  if ((nd & 1) != 0) nf += 2;                  // ... also when the double count is odd
  sv.coef = f ? f + nf : nullptr;
  nf += (size_t)d->B * d->S * 12 * d->H * d->W;
  sv.geo = f ? reinterpret_cast<float4*>(f + nf) : nullptr;
  nf += BNS * 4 * (size_t)d->H * d->W;
  return nd + (nf + 1) / 2;
}

__global__ void k_fill_scalar(float* p, float v) { *p = v; }

// uint8 frames -> fp32 in [0, 1] (COLVO_F_HOST_U8): x = u8 * (1.0f / 255.0f), one rounding; four pixels per thread
__global__ void __launch_bounds__(256) k_widen_u8(const uint8_t* __restrict__ in, float* __restrict__ out, long long n) {
  const float s = 1.0f / 255.0f;
  const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uchar4 u = __ldcs(reinterpret_cast<const uchar4*>(in) + i);
    reinterpret_cast<float4*>(out)[i] = make_float4((float)u.x * s, (float)u.y * s, (float)u.z * s, (float)u.w * s);
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i] * s;
}

}  // namespace

extern "C" {

int colvo_version(void) { return COLVO_VERSION; }

const char* colvo_error_string(int rc) {
  switch (rc) {
    case 0: return "success";
    case COLVO_E_BAD_DESC: return "colvo: bad descriptor (sizes, pyramid shapes h_k = H >> k, 1 <= N <= 2, 1 <= S <= 4)";
    case COLVO_E_WORKSPACE: return "colvo: workspace too small (see colvo_workspace_bytes)";
    case COLVO_E_NULL_PTR: return "colvo: required pointer is NULL";
    case COLVO_E_MISALIGNED: return "colvo: pointer not aligned (fp32 buffers 4 B, saved 8 B, workspace 256 B)";
    case COLVO_E_UNSUPPORTED: return "colvo: unsupported configuration (e.g. grad_srcs with packed bf16 images)";
    case COLVO_E_TENSOR_MAP: return "colvo: the driver refused a TMA tensor map (cuTensorMapEncodeTiled)";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString(static_cast<cudaError_t>(rc));
  return "colvo: unknown error code";
}

int colvo_desc_init(ColvoDesc* d, int32_t B, int32_t N, int32_t S, int32_t H, int32_t W, uint32_t flags) {
  if (!d) return COLVO_E_NULL_PTR;
  memset(d, 0, sizeof(*d));
  d->B = B; d->N = N; d->S = S; d->H = H; d->W = W;
  for (int k = 0; k < COLVO_MAX_SCALES && k < S; ++k) { d->h[k] = H >> k; d->w[k] = W >> k; }
  d->alpha = 0.85f; d->c1 = 1e-4f; d->c2 = 9e-4f; d->eps_proj = 1e-7f; d->eps_lcc = 1e-6f; d->eps_disp = 1e-7f;
  d->z_min = 1e-3f; d->smooth_weight = 1e-3f; d->geo_weight = 0.f;
  d->flags = flags;
  return check_desc(d);
}

int colvo_workspace_bytes(const ColvoDesc* d, size_t* bytes) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!bytes) return COLVO_E_NULL_PTR;
  FwdBuffers F;
  BwdBuffers Bw;
  size_t a = carve_fwd(d, nullptr, F), b = carve_bwd(d, nullptr, Bw);
  *bytes = a > b ? a : b;
  return 0;
}

int colvo_saved_doubles(const ColvoDesc* d, size_t* count) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!count) return COLVO_E_NULL_PTR;
  SavedView sv;
  *count = carve_saved(d, nullptr, sv);
  return 0;
}

int colvo_photo_forward(const ColvoDesc* d, const void* tgt, const void* srcs, const float* const* depth,
                        const float* K, const float* T, const float* src_depth, float* loss, float* ab,
                        uint8_t* valid, uint8_t* sel, double* saved, void* ws, size_t ws_bytes, void* stream) {
  return colvo_photo_forward_occ(d, tgt, srcs, depth, K, T, src_depth, loss, ab, valid, sel, nullptr, saved, ws, ws_bytes, stream);
}

int colvo_photo_forward_occ(const ColvoDesc* d, const void* tgt, const void* srcs, const float* const* depth,
                            const float* K, const float* T, const float* src_depth, float* loss, float* ab,
                            uint8_t* valid, uint8_t* sel, float* occ, double* saved, void* ws, size_t ws_bytes,
                            void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!tgt || !srcs || !depth || !K || !T || !loss || !ab || !ws) return COLVO_E_NULL_PTR;
  for (int k = 0; k < d->S; ++k)
    if (!depth[k]) return COLVO_E_NULL_PTR;
  if ((d->flags & COLVO_F_SAVE_FOR_BWD) && (!sel || !saved)) return COLVO_E_NULL_PTR;
  if (((uintptr_t)ws & 255u) || ((uintptr_t)saved & 15u)) return COLVO_E_MISALIGNED;
  if ((d->flags & COLVO_F_PACKED_BF16) && ((((uintptr_t)tgt) | ((uintptr_t)srcs)) & 7u)) return COLVO_E_MISALIGNED;
  FwdBuffers F;
  if (carve_fwd(d, ws, F) > ws_bytes) return COLVO_E_WORKSPACE;
  KP P;
  fill_params(P, d);
  P.tgt = tgt; P.srcs = srcs; P.K = K; P.T = T;
  P.src_depth = (d->geo_weight != 0.f) ? src_depth : nullptr;
  for (int k = 0; k < d->S; ++k) P.depth[k] = depth[k];
  SavedView sv;
  carve_saved(d, (d->flags & COLVO_F_SAVE_FOR_BWD) ? saved : nullptr, sv);
  if (occ && !P.src_depth) return COLVO_E_UNSUPPORTED;     // the occlusion mask is a by-product of the geometric term
  return (int)launch_forward(P, F, loss, ab, valid, sel, occ, sv, static_cast<cudaStream_t>(stream));
}

int colvo_photo_backward(const ColvoDesc* d, const void* tgt, const void* srcs, const float* const* depth,
                         const float* K, const float* T, const float* src_depth, const float* grad_loss,
                         const uint8_t* sel, const double* saved, float* const* grad_depth, float* grad_T,
                         float* grad_srcs, float* grad_src_depth, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!tgt || !srcs || !depth || !K || !T || !grad_loss || !sel || !saved || !grad_depth || !grad_T || !ws)
    return COLVO_E_NULL_PTR;
  for (int k = 0; k < d->S; ++k)
    if (!depth[k] || !grad_depth[k]) return COLVO_E_NULL_PTR;
  const bool want_src = !(d->flags & COLVO_F_NO_SRC_GRAD);
  if (want_src && !grad_srcs) return COLVO_E_NULL_PTR;
  if (want_src && (d->flags & COLVO_F_PACKED_BF16)) return COLVO_E_UNSUPPORTED;   // quantised images carry no gradient
  if (((uintptr_t)ws & 255u) || ((uintptr_t)saved & 15u)) return COLVO_E_MISALIGNED;
  if ((d->flags & COLVO_F_PACKED_BF16) && ((((uintptr_t)tgt) | ((uintptr_t)srcs)) & 7u)) return COLVO_E_MISALIGNED;
  BwdBuffers Bw;
  if (carve_bwd(d, ws, Bw) > ws_bytes) return COLVO_E_WORKSPACE;
  KP P;
  fill_params(P, d);
  P.tgt = tgt; P.srcs = srcs; P.K = K; P.T = T;
  P.src_depth = (d->geo_weight != 0.f) ? src_depth : nullptr;
  for (int k = 0; k < d->S; ++k) P.depth[k] = depth[k];
  SavedView sv;
  carve_saved(d, const_cast<double*>(saved), sv);
  return (int)launch_backward(P, Bw, grad_loss, sel, sv, grad_depth, grad_T, want_src ? grad_srcs : nullptr,
                              P.src_depth ? grad_src_depth : nullptr, static_cast<cudaStream_t>(stream));
}

int colvo_debug_time_kernel(int which, void* ev_start, void* ev_stop) {
  if (which < 0 || which > 4) return COLVO_E_UNSUPPORTED;
  if (which != 0 && (!ev_start || !ev_stop)) return COLVO_E_NULL_PTR;
  g_timer.which = which;
  g_timer.start = static_cast<cudaEvent_t>(ev_start);
  g_timer.stop = static_cast<cudaEvent_t>(ev_stop);
  return 0;
}

// ---- consistency sweep ----------------------------------------------------------------------
// pairs whose warped frames share one pass of the scratch buffer: at most COLVO_SWEEP_PASS_MB of texels
#ifndef COLVO_SWEEP_PASS_MB
#define COLVO_SWEEP_PASS_MB 256
#endif
static int consistency_pairs_per_pass(int F, int H, int W) {
  const long long per = (long long)H * W * (long long)sizeof(float4);
  long long n = ((long long)COLVO_SWEEP_PASS_MB << 20) / per;
  if (n < 1) n = 1;
  if (n > F - 1) n = F - 1;
  return (int)n;
}
static size_t carve_consistency(int F, int H, int W, void* ws, double** stat, int* chunks, double** pe_part,
                                float** ab, float4** iw) {
  Carver c(ws);
  const size_t Pn = (size_t)(F - 1);
  const size_t tiles = (size_t)div_up(W, 32) * div_up(H, kFwdTileH);
  *chunks = div_up(H * W, kThreads * kStatPPT);
  *stat = c.take<double>(Pn * (*chunks) * kStatVals);
  *pe_part = c.take<double>(Pn * tiles * 2);
  *ab = c.take<float>(Pn * 2);
  *iw = c.take<float4>((size_t)consistency_pairs_per_pass(F, H, W) * H * W);
  return c.off;
}

int colvo_consistency_workspace_bytes(int32_t F, int32_t H, int32_t W, size_t* bytes) {
  if (!bytes) return COLVO_E_NULL_PTR;
  if (F < 2 || H < 2 || W < 2 || F - 1 > 65535) return COLVO_E_BAD_DESC;
  double *a, *b;
  float* c;
  float4* iw;
  int ch;
  *bytes = carve_consistency(F, H, W, nullptr, &a, &ch, &b, &c, &iw);
  return 0;
}

int colvo_consistency(int32_t F, int32_t H, int32_t W, uint32_t flags, const float* frames, const float* depth,
                      const float* T, const float* K, int32_t k_per_pair, float* out, void* ws, size_t ws_bytes,
                      void* stream) {
  if (F < 2 || H < 2 || W < 2 || F - 1 > 65535) return COLVO_E_BAD_DESC;
  if (!frames || !depth || !T || !K || !out || !ws) return COLVO_E_NULL_PTR;
  if ((uintptr_t)ws & 255u) return COLVO_E_MISALIGNED;
  double *stat, *pe_part;
  float* ab;
  float4* iw;
  int chunks;
  if (carve_consistency(F, H, W, ws, &stat, &chunks, &pe_part, &ab, &iw) > ws_bytes) return COLVO_E_WORKSPACE;
  ColvoDesc d;
  int rc = colvo_desc_init(&d, F - 1, 1, 1, H, W, flags & COLVO_F_LCC);
  if (rc) return rc;
  KP P;
  fill_params(P, &d);
  // pair i: target = frames[i], source = frames[i+1]: one array, two views
  P.tgt = frames;
  P.srcs = frames + 3ll * P.HW;
  P.tgt_bf = 1;
  P.src_bf = 1;
  P.src_nf = 0;
  P.depth[0] = depth;
  P.K = K;
  P.K_bs = k_per_pair ? 9 : 0;
  P.T = T;
  P.T_bs = 16;
  P.T_ns = 0;
  return (int)launch_consistency(P, stat, chunks, pe_part, ab, out, iw, consistency_pairs_per_pass(F, H, W),
                                 static_cast<cudaStream_t>(stream));
}

// ---- end-to-end step on host buffers ----------------------------------------------------------
struct Arena {
  float *tgt, *srcs, *depth[kMaxS], *K, *T, *loss, *ab, *one, *grad_depth[kMaxS], *grad_T, *grad_srcs;
  uint8_t* sel;
  double* saved;
  void* ws;
  size_t ws_bytes;
  uint8_t *tgt_u8, *srcs_u8;     // byte staging of the frames (COLVO_F_HOST_U8), else null
};

static size_t carve_arena(const ColvoDesc* d, void* base, Arena& A) {
  Carver c(base);
  const size_t HW = (size_t)d->H * d->W, B = d->B, N = d->N, S = d->S;
  A.tgt = c.take<float>(B * 3 * HW);
  A.srcs = c.take<float>(B * N * 3 * HW);
  for (int k = 0; k < kMaxS; ++k) A.depth[k] = (k < d->S) ? c.take<float>(B * d->h[k] * d->w[k]) : nullptr;
  A.K = c.take<float>(B * 9);
  A.T = c.take<float>(B * N * 16);
  A.loss = c.take<float>(1);
  A.ab = c.take<float>(B * N * S * 2);
  A.one = c.take<float>(1);
  for (int k = 0; k < kMaxS; ++k) A.grad_depth[k] = (k < d->S) ? c.take<float>(B * d->h[k] * d->w[k]) : nullptr;
  A.grad_T = c.take<float>(B * N * 16);
  A.grad_srcs = c.take<float>(B * N * 3 * HW);
  A.sel = c.take<uint8_t>(B * S * HW);
  size_t ns = 0;
  colvo_saved_doubles(d, &ns);
  A.saved = c.take<double>(ns);
  colvo_workspace_bytes(d, &A.ws_bytes);
  A.ws = c.take<char>(A.ws_bytes);
  const bool u8 = (d->flags & COLVO_F_HOST_U8) != 0;       // last, so that the gradient offsets do not depend on the flag
  A.tgt_u8 = u8 ? c.take<uint8_t>(B * 3 * HW) : nullptr;
  A.srcs_u8 = u8 ? c.take<uint8_t>(B * N * 3 * HW) : nullptr;
  return c.off;
}

int colvo_step_host_arena_bytes(const ColvoDesc* d, size_t* bytes) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!bytes) return COLVO_E_NULL_PTR;
  Arena A;
  *bytes = carve_arena(d, nullptr, A);
  return 0;
}

int colvo_step_host_arena_grads(const ColvoDesc* d, size_t* grad_depth_off, size_t* grad_T_off, size_t* grad_srcs_off) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!grad_depth_off || !grad_T_off || !grad_srcs_off) return COLVO_E_NULL_PTR;
  Arena A;
  carve_arena(d, reinterpret_cast<void*>(uintptr_t(256)), A);     // carve at a dummy aligned base: offsets = pointer - base
  const char* base = reinterpret_cast<const char*>(uintptr_t(256));
  for (int k = 0; k < d->S; ++k) grad_depth_off[k] = (size_t)(reinterpret_cast<const char*>(A.grad_depth[k]) - base);
  *grad_T_off = (size_t)(reinterpret_cast<const char*>(A.grad_T) - base);
  *grad_srcs_off = (size_t)(reinterpret_cast<const char*>(A.grad_srcs) - base);
  return 0;
}

int colvo_photo_step_host(const ColvoDesc* d_in, const void* h_tgt, const void* h_srcs, const float* const* h_depth,
                          const float* h_K, const float* h_T, float* h_loss, float* const* h_grad_depth,
                          float* h_grad_T, float* h_grad_srcs, float grad_scale, void* arena, size_t arena_bytes,
                          void* stream) {
  int rc = check_desc(d_in);
  if (rc) return rc;
  if (!h_tgt || !h_srcs || !h_depth || !h_K || !h_T || !h_loss || !arena) return COLVO_E_NULL_PTR;
  const bool grads_to_host = h_grad_depth != nullptr;     // NULL: the gradients stay in the device arena
  if (grads_to_host && !h_grad_T) return COLVO_E_NULL_PTR;
  ColvoDesc d = *d_in;
  d.flags |= COLVO_F_SAVE_FOR_BWD;
  const bool want_src = !(d.flags & COLVO_F_NO_SRC_GRAD);
  if (grads_to_host && want_src && !h_grad_srcs) return COLVO_E_NULL_PTR;
  for (int k = 0; k < d.S; ++k)              // every pointer is checked before the first copy is enqueued
    if (!h_depth[k] || (grads_to_host && !h_grad_depth[k])) return COLVO_E_NULL_PTR;
  if ((uintptr_t)arena & 255u) return COLVO_E_MISALIGNED;
  Arena A;
  if (carve_arena(&d, arena, A) > arena_bytes) return COLVO_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t HW = (size_t)d.H * d.W, B = d.B, N = d.N;
  cudaError_t e;
#define CV_COPY(dst, src, count, kind)                                                      \
  do {                                                                                      \
    e = cudaMemcpyAsync((dst), (src), sizeof(float) * (count), (kind), st);                 \
    if (e != cudaSuccess) return (int)e;                                                    \
  } while (0)
  if (d.flags & COLVO_F_HOST_U8) {      // a quarter of the image bytes over PCIe, widened on the device
    e = cudaMemcpyAsync(A.tgt_u8, h_tgt, B * 3 * HW, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyAsync(A.srcs_u8, h_srcs, B * N * 3 * HW, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    k_widen_u8<<<148 * 4, 256, 0, st>>>(A.tgt_u8, A.tgt, (long long)(B * 3 * HW));
    k_widen_u8<<<148 * 8, 256, 0, st>>>(A.srcs_u8, A.srcs, (long long)(B * N * 3 * HW));
  } else {
    CV_COPY(A.tgt, h_tgt, B * 3 * HW, cudaMemcpyHostToDevice);
    CV_COPY(A.srcs, h_srcs, B * N * 3 * HW, cudaMemcpyHostToDevice);
  }
  for (int k = 0; k < d.S; ++k) CV_COPY(A.depth[k], h_depth[k], B * d.h[k] * d.w[k], cudaMemcpyHostToDevice);
  CV_COPY(A.K, h_K, B * 9, cudaMemcpyHostToDevice);
  CV_COPY(A.T, h_T, B * N * 16, cudaMemcpyHostToDevice);
  k_fill_scalar<<<1, 1, 0, st>>>(A.one, grad_scale);
  const float* depth_p[kMaxS] = {A.depth[0], A.depth[1], A.depth[2], A.depth[3]};
  float* gdepth_p[kMaxS] = {A.grad_depth[0], A.grad_depth[1], A.grad_depth[2], A.grad_depth[3]};
  d.geo_weight = 0.f;   // the host step carries no source depth maps
  d.flags &= ~(COLVO_F_PACKED_BF16 | COLVO_F_HOST_U8);   // ... and the kernels see planar fp32 images
  rc = colvo_photo_forward(&d, A.tgt, A.srcs, depth_p, A.K, A.T, nullptr, A.loss, A.ab, nullptr, A.sel, A.saved, A.ws,
                           A.ws_bytes, stream);
  if (rc) return rc;
  rc = colvo_photo_backward(&d, A.tgt, A.srcs, depth_p, A.K, A.T, nullptr, A.one, A.sel, A.saved, gdepth_p, A.grad_T,
                            want_src ? A.grad_srcs : nullptr, nullptr, A.ws, A.ws_bytes, stream);
  if (rc) return rc;
  CV_COPY(h_loss, A.loss, 1, cudaMemcpyDeviceToHost);
  if (grads_to_host) {
    for (int k = 0; k < d.S; ++k) CV_COPY(h_grad_depth[k], A.grad_depth[k], B * d.h[k] * d.w[k], cudaMemcpyDeviceToHost);
    CV_COPY(h_grad_T, A.grad_T, B * N * 16, cudaMemcpyDeviceToHost);
    if (want_src) CV_COPY(h_grad_srcs, A.grad_srcs, B * N * 3 * HW, cudaMemcpyDeviceToHost);
  }
#undef CV_COPY
  return 0;
}

}  // extern "C"
