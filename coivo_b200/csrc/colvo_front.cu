// Front-end of the path: raw network outputs -> the tensors the loss takes.
//   k_pose_fwd / k_pose_bwd   axis-angle + translation (+ optional inversion) -> T [B,N,4,4] and its adjoint
//   k_disp_fwd / k_disp_bwd   sigmoid disparity -> depth, depth gradient -> disparity gradient
// oracle/frontend.py is the arithmetic contract (Monodepth2's transformation_from_parameters /
// disp_to_depth, assumption A0).  The pose adjoint is evaluated with forward-mode dual numbers over
// the six parameters: 24 threads' worth of work per step, so clarity wins over cleverness.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/colvo.h"

namespace {

struct Dual6 {
  float v, d[6];
};
__device__ __forceinline__ Dual6 mk(float v) { Dual6 r; r.v = v; for (int i = 0; i < 6; ++i) r.d[i] = 0.f; return r; }
__device__ __forceinline__ Dual6 var(float v, int i) { Dual6 r = mk(v); r.d[i] = 1.f; return r; }
__device__ __forceinline__ Dual6 operator+(const Dual6& a, const Dual6& b) { Dual6 r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ Dual6 operator-(const Dual6& a, const Dual6& b) { Dual6 r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ Dual6 operator-(const Dual6& a) { Dual6 r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }
__device__ __forceinline__ Dual6 operator*(const Dual6& a, const Dual6& b) { Dual6 r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ Dual6 operator/(const Dual6& a, const Dual6& b) { Dual6 r; float ib = 1.0f / b.v; r.v = a.v * ib; for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
__device__ __forceinline__ Dual6 dsqrt(const Dual6& a) { Dual6 r; r.v = sqrtf(a.v); float g = a.v > 0.f ? 0.5f / r.v : 0.f; for (int i = 0; i < 6; ++i) r.d[i] = g * a.d[i]; return r; }
__device__ __forceinline__ Dual6 dsin(const Dual6& a) { Dual6 r; r.v = sinf(a.v); float g = cosf(a.v); for (int i = 0; i < 6; ++i) r.d[i] = g * a.d[i]; return r; }
__device__ __forceinline__ Dual6 dcos(const Dual6& a) { Dual6 r; r.v = cosf(a.v); float g = -sinf(a.v); for (int i = 0; i < 6; ++i) r.d[i] = g * a.d[i]; return r; }

// T[0..11] = rows 0..2 of the 4x4 (row-major 3x4), as duals over (axisangle[3], translation[3])
__device__ void pose_dual(const float* aa, const float* tr, bool invert, Dual6 (&T)[12]) {
  Dual6 vx = var(aa[0], 0), vy = var(aa[1], 1), vz = var(aa[2], 2);
  Dual6 t[3] = {var(tr[0], 3), var(tr[1], 4), var(tr[2], 5)};
  Dual6 angle = dsqrt(vx * vx + vy * vy + vz * vz);
  Dual6 den = angle + mk(1e-7f);
  Dual6 x = vx / den, y = vy / den, z = vz / den;
  Dual6 ca = dcos(angle), sa = dsin(angle), C = mk(1.0f) - ca;
  Dual6 xs = x * sa, ys = y * sa, zs = z * sa, xC = x * C, yC = y * C, zC = z * C;
  Dual6 xyC = x * yC, yzC = y * zC, zxC = z * xC;
  Dual6 R[9] = {x * xC + ca, xyC - zs, zxC + ys, xyC + zs, y * yC + ca, yzC - xs, zxC - ys, yzC + xs, z * zC + ca};
  if (!invert) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
      T[4 * i + 3] = t[i];
    }
  } else {  // R^T | -R^T t
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * j + i];
      T[4 * i + 3] = -(R[0 + i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);
    }
  }
}

__global__ void k_pose_fwd(int count, int N, unsigned invert_mask, const float* __restrict__ aa,
                           const float* __restrict__ tr, float* __restrict__ T) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Dual6 D[12];
  pose_dual(aa + 3 * i, tr + 3 * i, (invert_mask >> (i % N)) & 1u, D);
  float* o = T + 16 * (long long)i;
  for (int j = 0; j < 12; ++j) o[j] = D[j].v;
  o[12] = o[13] = o[14] = 0.f;
  o[15] = 1.f;
}

__global__ void k_pose_bwd(int count, int N, unsigned invert_mask, const float* __restrict__ aa,
                           const float* __restrict__ tr, const float* __restrict__ gT, float* __restrict__ gaa,
                           float* __restrict__ gtr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Dual6 D[12];
  pose_dual(aa + 3 * i, tr + 3 * i, (invert_mask >> (i % N)) & 1u, D);
  const float* g = gT + 16 * (long long)i;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 12; ++j)
    for (int p = 0; p < 6; ++p) acc[p] += g[j] * D[j].d[p];
  for (int p = 0; p < 3; ++p) { gaa[3 * i + p] = acc[p]; gtr[3 * i + p] = acc[3 + p]; }
}

struct PtrPack { const float* in[COLVO_MAX_SCALES]; const float* in2[COLVO_MAX_SCALES]; float* out[COLVO_MAX_SCALES]; long long n[COLVO_MAX_SCALES]; };

__global__ void k_disp_fwd(PtrPack P, float min_disp, float range) {
  const int k = blockIdx.y;
  const float* d = P.in[k];
  float* o = P.out[k];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P.n[k]; i += (long long)gridDim.x * blockDim.x)
    o[i] = 1.0f / (min_disp + range * d[i]);
}
// grad_disp = -range * depth^2 * grad_depth
__global__ void k_disp_bwd(PtrPack P, float range) {
  const int k = blockIdx.y;
  const float* depth = P.in[k];
  const float* g = P.in2[k];
  float* o = P.out[k];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P.n[k]; i += (long long)gridDim.x * blockDim.x) {
    float z = depth[i];
    o[i] = -range * z * z * g[i];
  }
}

}  // namespace

extern "C" {

int colvo_pose_from_axisangle(int32_t B, int32_t N, uint32_t invert_mask, const float* axisangle,
                              const float* translation, float* T, void* stream) {
  if (B < 1 || N < 1 || N > 32) return COLVO_E_BAD_DESC;
  if (!axisangle || !translation || !T) return COLVO_E_NULL_PTR;
  const int count = B * N;
  k_pose_fwd<<<(count + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(count, N, invert_mask, axisangle, translation, T);
  return (int)cudaGetLastError();
}

int colvo_pose_from_axisangle_backward(int32_t B, int32_t N, uint32_t invert_mask, const float* axisangle,
                                       const float* translation, const float* grad_T, float* grad_axisangle,
                                       float* grad_translation, void* stream) {
  if (B < 1 || N < 1 || N > 32) return COLVO_E_BAD_DESC;
  if (!axisangle || !translation || !grad_T || !grad_axisangle || !grad_translation) return COLVO_E_NULL_PTR;
  const int count = B * N;
  k_pose_bwd<<<(count + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(count, N, invert_mask, axisangle, translation,
                                                                              grad_T, grad_axisangle, grad_translation);
  return (int)cudaGetLastError();
}

int colvo_disp_to_depth(int32_t S, const int64_t* counts, const float* const* disp, float* const* depth, float min_depth,
                        float max_depth, void* stream) {
  if (S < 1 || S > COLVO_MAX_SCALES || !(min_depth > 0.f) || !(max_depth > min_depth)) return COLVO_E_BAD_DESC;
  if (!counts || !disp || !depth) return COLVO_E_NULL_PTR;
  PtrPack P = {};
  long long mx = 0;
  for (int k = 0; k < S; ++k) {
    if (!disp[k] || !depth[k] || counts[k] < 0) return COLVO_E_NULL_PTR;
    P.in[k] = disp[k]; P.out[k] = depth[k]; P.n[k] = counts[k];
    mx = counts[k] > mx ? counts[k] : mx;
  }
  const float min_disp = 1.0f / max_depth, range = 1.0f / min_depth - min_disp;
  int blocks = (int)((mx + 1023) / 1024);
  blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
  k_disp_fwd<<<dim3(blocks, S), 256, 0, static_cast<cudaStream_t>(stream)>>>(P, min_disp, range);
  return (int)cudaGetLastError();
}

int colvo_disp_to_depth_backward(int32_t S, const int64_t* counts, const float* const* depth,
                                 const float* const* grad_depth, float* const* grad_disp, float min_depth,
                                 float max_depth, void* stream) {
  if (S < 1 || S > COLVO_MAX_SCALES || !(min_depth > 0.f) || !(max_depth > min_depth)) return COLVO_E_BAD_DESC;
  if (!counts || !depth || !grad_depth || !grad_disp) return COLVO_E_NULL_PTR;
  PtrPack P = {};
  long long mx = 0;
  for (int k = 0; k < S; ++k) {
    if (!depth[k] || !grad_depth[k] || !grad_disp[k] || counts[k] < 0) return COLVO_E_NULL_PTR;
    P.in[k] = depth[k]; P.in2[k] = grad_depth[k]; P.out[k] = grad_disp[k]; P.n[k] = counts[k];
    mx = counts[k] > mx ? counts[k] : mx;
  }
  const float range = 1.0f / min_depth - 1.0f / max_depth;
  int blocks = (int)((mx + 1023) / 1024);
  blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
  k_disp_bwd<<<dim3(blocks, S), 256, 0, static_cast<cudaStream_t>(stream)>>>(P, range);
  return (int)cudaGetLastError();
}

}  // extern "C"
