// Forward kernels of the ColVO photometric-loss path,
// hand-written for sm_100a.  oracle/photometric.py is the arithmetic contract.
//
//   k_tgt_pyramid   box-averaged target pyramid for the smoothness term        (row 9)
//   k_disp_sum      partial sums of 1/D per (b,k)                               (row 9)
//   k_warp_stats    warp every (b,n,k) frame once, fp64 LCC sums, valid mask    (rows 0-5)
//   k_lcc_solve     (a, b) per warped frame                                     (row 5)
//   k_photo_fwd     per 32x8 tile: identity + re-projection candidates, SSIM+L1,
//                   min-reprojection / auto-mask, loss partials, dL/da, dL/db   (rows 0-8)
//   k_smooth_fwd    edge-aware smoothness partials                              (row 9)
//   k_finalize_fwd  deterministic final sums -> loss, G_a, G_b                  (row 10)
#include "colvo_kernels.cuh"

namespace colvo {

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_tgt_pyramid(KP P, float* p1, float* p2, float* p3) {
  // one thread per pooled pixel per channel, all scales k >= 1 in one launch (blockIdx.y = k-1)
  const int k = blockIdx.y + 1;
  if (k >= P.S) return;
  float* out = (k == 1) ? p1 : (k == 2 ? p2 : p3);
  const int hk = P.h[k], wk = P.w[k];
  const long long total = (long long)P.B * 3 * hk * wk;
  const int f = 1 << k;
  const float inv = 1.0f / (float)(f * f);
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kThreads) {
    int x = (int)(i % wk);
    long long r = i / wk;
    int y = (int)(r % hk);
    long long bc = r / hk;                       // b*3 + c
    int b = (int)(bc / 3), c = (int)(bc - 3 * b);
    const float* src = P.tgt + (long long)b * P.tgt_bs + (long long)c * P.HW + (long long)(y * f) * P.W + x * f;
    float s = 0.f;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) s += __ldg(src + dy * P.W + dx);
    out[i] = s * inv;
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_disp_sum(KP P, double* __restrict__ disp_part) {
  __shared__ double sm[kThreads / 32];
  const int bk = blockIdx.y, b = bk / P.S, k = bk % P.S;
  const int n = P.h[k] * P.w[k];
  const float* D = P.depth[k] + (long long)b * P.depth_bs[k];
  double acc = 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    acc += (double)(1.0f / __ldg(D + i));
  double v[1] = {acc};
  block_reduce_store<1, double>(v, sm, disp_part + (long long)bk * kSmoothChunks + blockIdx.x);
}

// mean inverse depth of (b,k) from the chunk partials, fixed order -> deterministic
__device__ __forceinline__ double disp_mean(const KP& P, const double* __restrict__ disp_part, int bk, int k) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kSmoothChunks; ++i) s += disp_part[(long long)bk * kSmoothChunks + i];
  return s / (double)(P.h[k] * P.w[k]);
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    k_warp_stats(KP P, double* __restrict__ part, uint8_t* __restrict__ valid_out) {
  __shared__ double sm[(kThreads / 32) * 5];
  const int bnk = blockIdx.y;
  const int k = bnk % P.S, n = (bnk / P.S) % P.N, b = bnk / (P.S * P.N);
  const Cam cam = load_cam(P, b);
  const Pose pose = load_pose(P, b, n);
  const float* Dk = P.depth[k] + (long long)b * P.depth_bs[k];
  const float* src = P.srcs + (long long)b * P.src_bs + (long long)n * P.src_ns;
  const float* tg = P.tgt + (long long)b * P.tgt_bs;
  int pix = blockIdx.x * (kThreads * kStatPPT) + threadIdx.x;
  int py = pix / P.W, px = pix - py * P.W;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
  for (int i = 0; i < kStatPPT; ++i) {
    if (pix < P.HW) {
      Geo g; Taps t; Texels tx; float x[3];
      warp_pixel(P, Dk, k, src, cam, pose, px, py, g, t, tx, x);
      if (valid_out) valid_out[(long long)bnk * P.HW + pix] = g.valid ? 1 : 0;
      if (g.valid) {
        float y0 = __ldg(tg + pix), y1 = __ldg(tg + P.HW + pix), y2 = __ldg(tg + 2 * P.HW + pix);
        acc[0] += 3.0;
        acc[1] += (double)(x[0] + x[1] + x[2]);
        acc[2] += (double)(y0 + y1 + y2);
        acc[3] += (double)(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        acc[4] += (double)(x[0] * y0 + x[1] * y1 + x[2] * y2);
      }
    }
    pix += kThreads;
    px += kThreads;
    while (px >= P.W) { px -= P.W; ++py; }
  }
  block_reduce_store<5, double>(acc, sm, part + ((long long)bnk * gridDim.x + blockIdx.x) * 5);
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    k_lcc_solve(KP P, const double* __restrict__ part, int chunks, float* __restrict__ ab, double* __restrict__ saved) {
  const int bnk = blockIdx.x, lane = threadIdx.x;
  double s[5] = {0, 0, 0, 0, 0};
  if (P.flags & 1u) {
    for (int c = lane; c < chunks; c += 32) {
#pragma unroll
      for (int j = 0; j < 5; ++j) s[j] += part[((long long)bnk * chunks + c) * 5 + j];
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) s[j] = warp_sum(s[j]);
  }
  if (lane == 0) {
    double a = 1.0, b = 0.0, mx = 0.0, my = 0.0, inv_nvar = 0.0, n = s[0];
    if ((P.flags & 1u) && n > 0.0) {
      mx = s[1] / n;
      my = s[2] / n;
      double var = s[3] / n - mx * mx;
      double cov = s[4] / n - mx * my;
      a = cov / (var + (double)P.eps_lcc);
      b = my - a * mx;
      inv_nvar = 1.0 / (n * (var + (double)P.eps_lcc));
    }
    // the backward differentiates the fp32-rounded pair it actually applied
    float af = (float)a, bf = (float)b;
    ab[2 * bnk + 0] = af;
    ab[2 * bnk + 1] = bf;
    if (saved) {
      double* o = saved + (long long)bnk * kSavedPerFrame;
      o[0] = n; o[1] = mx; o[2] = my; o[3] = inv_nvar; o[4] = (double)af; o[5] = (double)bf; o[6] = 0.0; o[7] = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// The fused tile kernel.  One CTA = one 32x8 output tile of one triplet; the (k, n) loops run
// inside the CTA so the target tile, its SSIM moments and the identity candidates are
// computed once and shared by all 2*S warps.
constexpr int kFH = kTileH + 2, kFW = kTileW + 2;   // tile + 1-pixel SSIM halo

template <int NS>
__device__ __forceinline__ float pe_own(const float (&xs)[3][kFH][kFW], const float (&y9)[3][9], const float (&muy)[3],
                                        const float (&sgy)[3], int ty, int tx, float a, float b, const KP& P,
                                        float* dpa, float* dpb) {
  float pe = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = 0.f, sxx = 0.f, sxy = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float v = xs[c][ty + j / 3][tx + j % 3];
      s += v;
      sxx = fmaf(v, v, sxx);
      sxy = fmaf(v, y9[c][j], sxy);
    }
    const float i9 = 1.0f / 9.0f;
    pe += pe_channel(s * i9, sxx * i9, sxy * i9, muy[c], sgy[c], xs[c][ty + 1][tx + 1], y9[c][4], a, b, P.alpha, P.c1,
                     P.c2, dpa, dpb);
  }
  return pe * (1.0f / 3.0f);
}

template <int NS>
__global__ void __launch_bounds__(kThreads, 2)
    k_photo_fwd(KP P, const float* __restrict__ ab, uint8_t* __restrict__ sel_out, double* __restrict__ loss_part,
                double* __restrict__ g_part, int need_g) {
  constexpr int NV = 1 + NS * kMaxS * 2;
  __shared__ float ys[3][kFH][kFW];
  __shared__ float xs[3][kFH][kFW];
  __shared__ double red[(kThreads / 32) * NV];

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int b = blockIdx.z, x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int px = x0 + tx, py = y0 + ty;
  const bool in_img = (px < P.W) && (py < P.H);
  const float* tg = P.tgt + (long long)b * P.tgt_bs;
  const Cam cam = load_cam(P, b);

  for (int idx = tid; idx < kFH * kFW; idx += kThreads) {
    int r = idx / kFW, c = idx - r * kFW;
    int gy = reflect_clamp(y0 - 1 + r, P.H), gx = reflect_clamp(x0 - 1 + c, P.W);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ys[ch][r][c] = __ldg(tg + (long long)ch * P.HW + gy * P.W + gx);
  }
  __syncthreads();

  float y9[3][9], muy[3], sgy[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float v = ys[c][ty + j / 3][tx + j % 3];
      y9[c][j] = v;
      s += v;
      ss = fmaf(v, v, ss);
    }
    muy[c] = s * (1.0f / 9.0f);
    sgy[c] = ss * (1.0f / 9.0f) - muy[c] * muy[c];
  }

  // identity candidates: raw sources, no calibration (oracle A10)
  float ident[NS];
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const float* src = P.srcs + (long long)b * P.src_bs + (long long)n * P.src_ns;
    for (int idx = tid; idx < kFH * kFW; idx += kThreads) {
      int r = idx / kFW, c = idx - r * kFW;
      int gy = reflect_clamp(y0 - 1 + r, P.H), gx = reflect_clamp(x0 - 1 + c, P.W);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) xs[ch][r][c] = __ldg(src + (long long)ch * P.HW + gy * P.W + gx);
    }
    __syncthreads();
    ident[n] = in_img ? pe_own<NS>(xs, y9, muy, sgy, ty, tx, 1.0f, 0.0f, P, nullptr, nullptr) : 0.f;
    __syncthreads();
  }

  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

#pragma unroll
  for (int k = 0; k < kMaxS; ++k) {
    if (k < P.S) {
      float best = ident[0];
      int sel = 0;
#pragma unroll
      for (int n = 1; n < NS; ++n)
        if (ident[n] < best) { best = ident[n]; sel = n; }
      float dpa[NS], dpb[NS];
      const float* Dk = P.depth[k] + (long long)b * P.depth_bs[k];
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float* src = P.srcs + (long long)b * P.src_bs + (long long)n * P.src_ns;
        const Pose pose = load_pose(P, b, n);
        const int bnk = (b * P.N + n) * P.S + k;
        const float a = __ldg(ab + 2 * bnk), bb = __ldg(ab + 2 * bnk + 1);
        for (int idx = tid; idx < kFH * kFW; idx += kThreads) {
          int r = idx / kFW, c = idx - r * kFW;
          int ry = y0 - 1 + r, rx = x0 - 1 + c;
          if (ry <= P.H && rx <= P.W) {           // positions further out are never read
            int gy = reflect_clamp(ry, P.H), gx = reflect_clamp(rx, P.W);
            Geo g; Taps t; Texels tx4; float x[3];
            warp_pixel(P, Dk, k, src, cam, pose, gx, gy, g, t, tx4, x);
            xs[0][r][c] = x[0];
            xs[1][r][c] = x[1];
            xs[2][r][c] = x[2];
          }
        }
        __syncthreads();
        dpa[n] = 0.f;
        dpb[n] = 0.f;
        if (in_img) {
          float pe = pe_own<NS>(xs, y9, muy, sgy, ty, tx, a, bb, P, need_g ? &dpa[n] : nullptr,
                                need_g ? &dpb[n] : nullptr);
          if (pe < best) { best = pe; sel = NS + n; }
        }
        __syncthreads();
      }
      if (in_img) {
        acc[0] += best;
        if (sel_out) sel_out[((long long)b * P.S + k) * P.HW + py * P.W + px] = (uint8_t)sel;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          if (sel == NS + n) {
            acc[1 + (n * kMaxS + k) * 2 + 0] += dpa[n] * (1.0f / 3.0f);
            acc[1 + (n * kMaxS + k) * 2 + 1] += dpb[n] * (1.0f / 3.0f);
          }
        }
      }
    }
  }

  // Per-tile partials.  dL/da and dL/db are sums of large terms of both signs: reduce them in
  // fp64 (each thread contributes at most one fp32 term per slot, so nothing is lost before).
  const int blk = (b * P.tiles_y + blockIdx.y) * P.tiles_x + blockIdx.x;
  {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i == 0 || need_g) {
        double s = warp_sum((double)acc[i]);
        if (lane == 0) red[wid * NV + i] = s;
      }
    }
    __syncthreads();
    if (tid < NV && (tid == 0 || need_g)) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) s += red[w * NV + tid];
      if (tid == 0) loss_part[blk] = s;
      else g_part[(long long)blk * (NS * kMaxS * 2) + (tid - 1)] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    k_smooth_fwd(KP P, const double* __restrict__ disp_part, const float* p1, const float* p2, const float* p3,
                 double* __restrict__ smooth_part, double* __restrict__ saved_mean) {
  __shared__ double sm[(kThreads / 32) * 2];
  const int bk = blockIdx.y, b = bk / P.S, k = bk % P.S;
  const int hk = P.h[k], wk = P.w[k], n = hk * wk;
  const double mean = disp_mean(P, disp_part, bk, k);
  if (saved_mean && blockIdx.x == 0 && threadIdx.x == 0) saved_mean[bk] = mean;
  const float inv = (float)(1.0 / (mean + (double)P.eps_disp));
  const float* D = P.depth[k] + (long long)b * P.depth_bs[k];
  const float* I;
  long long cs;
  if (k == 0) { I = P.tgt + (long long)b * P.tgt_bs; cs = P.HW; }
  else { I = ((k == 1) ? p1 : (k == 2 ? p2 : p3)) + (long long)b * 3 * n; cs = n; }
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    int y = i / wk, x = i - y * wk;
    float d = (1.0f / __ldg(D + i)) * inv;
    float i0 = __ldg(I + i), i1 = __ldg(I + cs + i), i2 = __ldg(I + 2 * cs + i);
    if (x + 1 < wk) {
      float dn = (1.0f / __ldg(D + i + 1)) * inv;
      float e = (fabsf(i0 - __ldg(I + i + 1)) + fabsf(i1 - __ldg(I + cs + i + 1)) + fabsf(i2 - __ldg(I + 2 * cs + i + 1))) *
                (1.0f / 3.0f);
      acc[0] += (double)(fabsf(d - dn) * expf(-e));
    }
    if (y + 1 < hk) {
      float dn = (1.0f / __ldg(D + i + wk)) * inv;
      float e = (fabsf(i0 - __ldg(I + i + wk)) + fabsf(i1 - __ldg(I + cs + i + wk)) +
                 fabsf(i2 - __ldg(I + 2 * cs + i + wk))) * (1.0f / 3.0f);
      acc[1] += (double)(fabsf(d - dn) * expf(-e));
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = 0; j < 2; ++j) {
    double s = warp_sum(acc[j]);
    if (lane == 0) sm[wid * 2 + j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += sm[w * 2 + threadIdx.x];
    smooth_part[((long long)bk * kSmoothChunks + blockIdx.x) * 2 + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------
// block 0: the scalar loss.  block 1 + bnk: G_a, G_b of warped frame bnk (dL/da, dL/db).
__global__ void __launch_bounds__(kThreads)
    k_finalize_fwd(KP P, const double* __restrict__ loss_part, const double* __restrict__ g_part,
                   const double* __restrict__ smooth_part, float* __restrict__ loss, double* __restrict__ saved,
                   int need_g) {
  __shared__ double sm[(kThreads / 32) * 2];
  const int tiles = P.tiles_x * P.tiles_y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (blockIdx.x == 0) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < P.B * tiles; i += kThreads) acc += loss_part[i];
    double s = warp_sum(acc);
    if (lane == 0) sm[wid] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double photo = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) photo += sm[w];
      double total = photo / ((double)P.B * (double)P.HW);     // sum_k L_photo,k
      for (int k = 0; k < P.S; ++k) {
        double sx = 0.0, sy = 0.0;
        for (int b = 0; b < P.B; ++b)
          for (int c = 0; c < kSmoothChunks; ++c) {
            sx += smooth_part[(((long long)b * P.S + k) * kSmoothChunks + c) * 2 + 0];
            sy += smooth_part[(((long long)b * P.S + k) * kSmoothChunks + c) * 2 + 1];
          }
        double nx = (double)P.B * P.h[k] * (P.w[k] - 1), ny = (double)P.B * (P.h[k] - 1) * P.w[k];
        double ls = (nx > 0 ? sx / nx : 0.0) + (ny > 0 ? sy / ny : 0.0);
        total += (double)P.smooth_weight / (double)(1 << k) * ls;
      }
      *loss = (float)(total / (double)P.S);
    }
    return;
  }
  if (!need_g) return;
  const int bnk = blockIdx.x - 1;
  const int k = bnk % P.S, n = (bnk / P.S) % P.N, b = bnk / (P.S * P.N);
  const int nv = P.N * kMaxS * 2;
  const int slot = (n * kMaxS + k) * 2;
  double acc[2] = {0.0, 0.0};
  for (int t = threadIdx.x; t < tiles; t += kThreads) {
    const double* g = g_part + ((long long)b * tiles + t) * nv + slot;
    acc[0] += g[0];
    acc[1] += g[1];
  }
  for (int j = 0; j < 2; ++j) {
    double s = warp_sum(acc[j]);
    if (lane == 0) sm[wid * 2 + j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += sm[w * 2 + threadIdx.x];
    // scale to dL/da, dL/db of the total loss: 1/S * 1/(B*HW)
    saved[(long long)bnk * kSavedPerFrame + 6 + threadIdx.x] = s / ((double)P.S * (double)P.B * (double)P.HW);
  }
}

// ------------------------------------------------------------------------------------------
// Consistency sweep (BASELINE config 5): per pair, mean pe over valid pixels.  N = 1, S = 1.
__global__ void __launch_bounds__(kThreads, 2)
    k_consistency_pe(KP P, const float* __restrict__ ab, double* __restrict__ pe_part) {
  __shared__ float ys[3][kFH][kFW];
  __shared__ float xs[3][kFH][kFW];
  __shared__ unsigned char vs[kFH][kFW];
  __shared__ double red[(kThreads / 32) * 2];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int b = blockIdx.z, x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int px = x0 + tx, py = y0 + ty;
  const bool in_img = (px < P.W) && (py < P.H);
  const float* tg = P.tgt + (long long)b * P.tgt_bs;
  const float* src = P.srcs + (long long)b * P.src_bs;
  const float* Dk = P.depth[0] + (long long)b * P.depth_bs[0];
  const Cam cam = load_cam(P, b);
  const Pose pose = load_pose(P, b, 0);
  for (int idx = tid; idx < kFH * kFW; idx += kThreads) {
    int r = idx / kFW, c = idx - r * kFW;
    int ry = y0 - 1 + r, rx = x0 - 1 + c;
    int gy = reflect_clamp(ry, P.H), gx = reflect_clamp(rx, P.W);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ys[ch][r][c] = __ldg(tg + (long long)ch * P.HW + gy * P.W + gx);
    if (ry <= P.H && rx <= P.W) {
      Geo g; Taps t; Texels tx4; float x[3];
      warp_pixel(P, Dk, 0, src, cam, pose, gx, gy, g, t, tx4, x);
      xs[0][r][c] = x[0];
      xs[1][r][c] = x[1];
      xs[2][r][c] = x[2];
      vs[r][c] = g.valid ? 1 : 0;
    }
  }
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  if (in_img && vs[ty + 1][tx + 1]) {
    float y9[3][9], muy[3], sgy[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        float v = ys[c][ty + j / 3][tx + j % 3];
        y9[c][j] = v;
        s += v;
        ss = fmaf(v, v, ss);
      }
      muy[c] = s * (1.0f / 9.0f);
      sgy[c] = ss * (1.0f / 9.0f) - muy[c] * muy[c];
    }
    float pe = pe_own<1>(xs, y9, muy, sgy, ty, tx, __ldg(ab + 2 * b), __ldg(ab + 2 * b + 1), P, nullptr, nullptr);
    acc[0] = (double)pe;
    acc[1] = 1.0;
  }
  const int blk = (b * P.tiles_y + blockIdx.y) * P.tiles_x + blockIdx.x;
  block_reduce_store<2, double>(acc, red, pe_part + (long long)blk * 2);
}

__global__ void __launch_bounds__(kThreads)
    k_consistency_final(KP P, const double* __restrict__ pe_part, const float* __restrict__ ab, float* __restrict__ out) {
  __shared__ double sm[(kThreads / 32) * 2];
  const int b = blockIdx.x, tiles = P.tiles_x * P.tiles_y;
  double acc[2] = {0.0, 0.0};
  for (int t = threadIdx.x; t < tiles; t += kThreads) {
    acc[0] += pe_part[((long long)b * tiles + t) * 2 + 0];
    acc[1] += pe_part[((long long)b * tiles + t) * 2 + 1];
  }
  double res[2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = 0; j < 2; ++j) {
    double s = warp_sum(acc[j]);
    if (lane == 0) sm[wid * 2 + j] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    res[0] = res[1] = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { res[0] += sm[w * 2]; res[1] += sm[w * 2 + 1]; }
    out[4 * b + 0] = (float)(res[1] > 0.0 ? res[0] / res[1] : 0.0);
    out[4 * b + 1] = ab[2 * b];
    out[4 * b + 2] = ab[2 * b + 1];
    out[4 * b + 3] = (float)(res[1] / (double)P.HW);
  }
}

// ------------------------------------------------------------------------------------------
static inline int div_up(int a, int b) { return (a + b - 1) / b; }

cudaError_t launch_tgt_pyramid(const KP& P, float* const* pyr, cudaStream_t st) {
  if (P.S > 1) {
    int blocks = div_up(P.B * 3 * P.h[1] * P.w[1], kThreads);
    k_tgt_pyramid<<<dim3(blocks, P.S - 1), kThreads, 0, st>>>(P, pyr[1], pyr[2], pyr[3]);
  }
  return cudaGetLastError();
}

cudaError_t launch_forward(const KP& P, const FwdBuffers& Wk, float* loss, float* ab, uint8_t* valid, uint8_t* sel,
                           double* saved, cudaStream_t st) {
  const bool lcc = (P.flags & 1u) != 0;
  const bool save = (P.flags & 4u) != 0;
  const int need_g = (save && lcc && !(P.flags & 2u)) ? 1 : 0;
  const int BNS = P.B * P.N * P.S;
  cudaError_t e = launch_tgt_pyramid(P, Wk.pyr, st);
  if (e != cudaSuccess) return e;
  k_disp_sum<<<dim3(kSmoothChunks, P.B * P.S), kThreads, 0, st>>>(P, Wk.disp_part);
  if (lcc || valid) {
    ScopedKernelTimer tm(3, st);
    k_warp_stats<<<dim3(Wk.stat_chunks, BNS), kThreads, 0, st>>>(P, Wk.stat_part, valid);
  }
  k_lcc_solve<<<BNS, 32, 0, st>>>(P, Wk.stat_part, Wk.stat_chunks, ab, saved);
  dim3 grid(P.tiles_x, P.tiles_y, P.B);
  {
    ScopedKernelTimer tm(1, st);
    if (P.N == 1)
      k_photo_fwd<1><<<grid, kThreads, 0, st>>>(P, ab, sel, Wk.loss_part, Wk.g_part, need_g);
    else
      k_photo_fwd<2><<<grid, kThreads, 0, st>>>(P, ab, sel, Wk.loss_part, Wk.g_part, need_g);
  }
  double* saved_mean = saved ? saved + (long long)BNS * kSavedPerFrame : nullptr;
  k_smooth_fwd<<<dim3(kSmoothChunks, P.B * P.S), kThreads, 0, st>>>(P, Wk.disp_part, Wk.pyr[1], Wk.pyr[2], Wk.pyr[3],
                                                                     Wk.smooth_part, saved_mean);
  k_finalize_fwd<<<1 + (need_g ? BNS : 0), kThreads, 0, st>>>(P, Wk.loss_part, Wk.g_part, Wk.smooth_part, loss, saved,
                                                               need_g);
  return cudaGetLastError();
}

cudaError_t launch_consistency(const KP& P, double* stat_part, int stat_chunks, double* pe_part, float* ab, float* out,
                               cudaStream_t st) {
  k_warp_stats<<<dim3(stat_chunks, P.B), kThreads, 0, st>>>(P, stat_part, nullptr);
  k_lcc_solve<<<P.B, 32, 0, st>>>(P, stat_part, stat_chunks, ab, nullptr);
  k_consistency_pe<<<dim3(P.tiles_x, P.tiles_y, P.B), kThreads, 0, st>>>(P, ab, pe_part);
  k_consistency_final<<<P.B, kThreads, 0, st>>>(P, pe_part, ab, out);
  return cudaGetLastError();
}

}  // namespace colvo
