// Backward kernels of the ColVO photometric-loss path (SURVEY.md section 8(a) row 11 and
// appendix A), hand-written for sm_100a.  The adjoint is analytic (no tape): the forward
// saves only sel, (a, b), the LCC statistics and dL/da, dL/db; everything else is recomputed.
//
//   k_photo_bwd      per 32x8 tile (+2 px halo): re-warp, SSIM adjoint in gather form, LCC
//                    adjoint, bilinear scatter-add (REDG) into grad_srcs, projection adjoint
//                    -> full-resolution depth adjoint + per-tile pose-gradient partials
//   k_depth_gather   adjoint of the bilinear depth up-sample in gather form (no atomics)
//   k_smooth_bwd_a/b edge-aware smoothness adjoint incl. the mean-normalisation term
//   k_pose_final     deterministic reduction of the pose-gradient partials
#include "colvo_kernels.cuh"

namespace colvo {

constexpr int kBH = kTileH + 4, kBW = kTileW + 4;   // tile + 2-pixel halo (raw warped image)
constexpr int kCH = kTileH + 2, kCW = kTileW + 2;   // tile + 1-pixel halo (window centres)
constexpr int kRing = 4 * kBW + 4 * kTileH;         // halo positions of the kBH x kBW tile

__device__ __forceinline__ void ring_pos(int j, int& r, int& c) {
  if (j < 2 * kBW) {
    r = j / kBW;
    c = j - r * kBW;
  } else if (j < 4 * kBW) {
    j -= 2 * kBW;
    int rr = j / kBW;
    r = kTileH + 2 + rr;
    c = j - rr * kBW;
  } else {
    j -= 4 * kBW;
    r = 2 + (j >> 2);
    int cc = j & 3;
    c = (cc < 2) ? cc : (kTileW + cc);
  }
}

template <int NS>
__global__ void __launch_bounds__(kThreads, 2)
    k_photo_bwd(KP P, const float* __restrict__ grad_loss, const uint8_t* __restrict__ sel,
                const double* __restrict__ saved, float* __restrict__ grad_d0, float* __restrict__ dD1,
                float* __restrict__ dD2, float* __restrict__ dD3, float* __restrict__ grad_srcs,
                double* __restrict__ pose_part) {
  __shared__ float xs[3][kBH][kBW];
  __shared__ float ys[3][kBH][kBW];
  __shared__ float ymu[3][kCH][kCW];
  __shared__ float ysg[3][kCH][kCW];
  __shared__ float coef[9][kCH][kCW];
  __shared__ unsigned char sels[kMaxS][kCH][kCW];
  __shared__ float cst[NS][kMaxS][6];               // a, b, P, Q, mean_x, mean_y per warped frame
  __shared__ double red[(kThreads / 32) * NS * 12];

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int b = blockIdx.z, x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int px = x0 + tx, py = y0 + ty;
  const bool in_img = (px < P.W) && (py < P.H);
  const int qx = reflect_clamp(px, P.W), qy = reflect_clamp(py, P.H);   // addressable stand-in when outside
  const float* tg = P.tgt + (long long)b * P.tgt_bs;
  const Cam cam = load_cam(P, b);
  const float go = __ldg(grad_loss);
  const float wscale = go / ((float)P.S * (float)P.B * (float)P.HW);

  // ---- phase 0: target tile (+2), selection masks (+1), per-frame constants ----
  for (int idx = tid; idx < kBH * kBW; idx += kThreads) {
    int r = idx / kBW, c = idx - r * kBW;
    int gy = reflect_clamp(y0 - 2 + r, P.H), gx = reflect_clamp(x0 - 2 + c, P.W);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ys[ch][r][c] = __ldg(tg + (long long)ch * P.HW + gy * P.W + gx);
  }
  for (int idx = tid; idx < kCH * kCW; idx += kThreads) {
    int r = idx / kCW, c = idx - r * kCW;
    int gy = y0 - 1 + r, gx = x0 - 1 + c;
    bool inside = gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
    for (int k = 0; k < P.S; ++k)
      sels[k][r][c] = inside ? sel[((long long)b * P.S + k) * P.HW + gy * P.W + gx] : (unsigned char)255;
  }
  if (tid < NS * kMaxS) {
    int n = tid / kMaxS, k = tid % kMaxS;
    float a = 1.f, bb = 0.f, Pc = 0.f, Qc = 0.f, mx = 0.f, my = 0.f;
    if (k < P.S) {
      const double* s = saved + ((long long)(b * P.N + n) * P.S + k) * kSavedPerFrame;
      a = (float)s[4];
      bb = (float)s[5];
      mx = (float)s[1];
      my = (float)s[2];
      if ((P.flags & 1u) && !(P.flags & 2u) && s[0] > 0.0) {
        Pc = (float)((double)go * (s[6] - s[7] * s[1]) * s[3]);
        Qc = (float)((double)go * s[7] * s[4] / s[0]);
      }
    }
    cst[n][k][0] = a; cst[n][k][1] = bb; cst[n][k][2] = Pc; cst[n][k][3] = Qc; cst[n][k][4] = mx; cst[n][k][5] = my;
  }
  __syncthreads();
  for (int idx = tid; idx < kCH * kCW; idx += kThreads) {
    int r = idx / kCW, c = idx - r * kCW;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        float v = ys[ch][r + j / 3][c + j % 3];
        s += v;
        ss = fmaf(v, v, ss);
      }
      float m = s * (1.0f / 9.0f);
      ymu[ch][r][c] = m;
      ysg[ch][r][c] = ss * (1.0f / 9.0f) - m * m;
    }
  }
  // (the first __syncthreads of the (k, n) loop orders these writes before their readers)

  // reflect-padding multiplicities of the 3x3 gather at the own pixel
  float my3[3], mx3[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    my3[d] = reflect_mult(py, py + d - 1, P.H);
    mx3[d] = reflect_mult(px, px + d - 1, P.W);
  }
  float yq[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) yq[ch] = ys[ch][ty + 2][tx + 2];

  float gp[NS][12];
#pragma unroll
  for (int n = 0; n < NS; ++n)
#pragma unroll
    for (int j = 0; j < 12; ++j) gp[n][j] = 0.f;

#pragma unroll 1
  for (int k = 0; k < P.S; ++k) {
    const float* Dk = P.depth[k] + (long long)b * P.depth_bs[k];
    float dD = 0.f;
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const float* src = P.srcs + (long long)b * P.src_bs + (long long)n * P.src_ns;
      const Pose pose = load_pose(P, b, n);
      const float a = cst[n][k][0], bb = cst[n][k][1];
      // ---- stage A: raw warped image on the tile + 2 halo; own pixel kept in registers ----
      Geo g; Taps t; Texels tx4; float xq[3];
      warp_pixel(P, Dk, k, src, cam, pose, qx, qy, g, t, tx4, xq);
      xs[0][ty + 2][tx + 2] = xq[0];
      xs[1][ty + 2][tx + 2] = xq[1];
      xs[2][ty + 2][tx + 2] = xq[2];
      if (tid < kRing) {
        int r, c;
        ring_pos(tid, r, c);
        int ry = y0 - 2 + r, rx = x0 - 2 + c;
        if (ry <= P.H && rx <= P.W) {             // windows of in-image centres reach at most index n
          Geo g2; Taps t2; Texels tx2; float x2[3];
          warp_pixel(P, Dk, k, src, cam, pose, reflect_clamp(rx, P.W), reflect_clamp(ry, P.H), g2, t2, tx2, x2);
          xs[0][r][c] = x2[0];
          xs[1][r][c] = x2[1];
          xs[2][r][c] = x2[2];
        }
      }
      __syncthreads();
      // ---- stage B: SSIM adjoint coefficient fields at the window centres (tile + 1) ----
      for (int idx = tid; idx < kCH * kCW; idx += kThreads) {
        int r = idx / kCW, c = idx - r * kCW;
        const bool on = sels[k][r][c] == (unsigned char)(NS + n);
        if (on) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            float s = 0.f, sxx = 0.f, sxy = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              float v = xs[ch][r + j / 3][c + j % 3];
              s += v;
              sxx = fmaf(v, v, sxx);
              sxy = fmaf(v, ys[ch][r + j / 3][c + j % 3], sxy);
            }
            const float i9 = 1.0f / 9.0f;
            Coef q = ssim_coef(s * i9, sxx * i9, sxy * i9, ymu[ch][r][c], ysg[ch][r][c], a, bb, P.alpha, P.c1, P.c2,
                               wscale);
            coef[3 * ch + 0][r][c] = q.ca;
            coef[3 * ch + 1][r][c] = q.cb;
            coef[3 * ch + 2][r][c] = q.cg;
          }
        } else {
#pragma unroll
          for (int f = 0; f < 9; ++f) coef[f][r][c] = 0.f;
        }
      }
      __syncthreads();
      // ---- stage C: gather, LCC adjoint, bilinear adjoint, projection adjoint ----
      if (in_img) {
        const float Pc = cst[n][k][2], Qc = cst[n][k][3], mx = cst[n][k][4], my = cst[n][k][5];
        const float wq = (sels[k][ty + 1][tx + 1] == (unsigned char)(NS + n)) ? wscale * (1.f - P.alpha) * (1.0f / 3.0f) : 0.f;
        float hq[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float A = 0.f, Bc = 0.f, G = 0.f;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              float m = my3[dy] * mx3[dx];
              A = fmaf(m, coef[3 * ch + 0][ty + dy][tx + dx], A);
              Bc = fmaf(m, coef[3 * ch + 1][ty + dy][tx + dx], Bc);
              G = fmaf(m, coef[3 * ch + 2][ty + dy][tx + dx], G);
            }
          float gq = A + xq[ch] * Bc + yq[ch] * G + wq * a * sgn(fmaf(a, xq[ch], bb) - yq[ch]);
          float lcc = g.valid ? (Pc * ((yq[ch] - my) - 2.f * a * (xq[ch] - mx)) - Qc) : 0.f;
          hq[ch] = gq + lcc;
        }
        float du = 0.f, dv = 0.f;
        const float w00 = (1.f - t.wx) * (1.f - t.wy), w01 = t.wx * (1.f - t.wy);
        const float w10 = (1.f - t.wx) * t.wy, w11 = t.wx * t.wy;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          du += hq[ch] * ((1.f - t.wy) * (tx4.i01[ch] - tx4.i00[ch]) + t.wy * (tx4.i11[ch] - tx4.i10[ch]));
          dv += hq[ch] * ((1.f - t.wx) * (tx4.i10[ch] - tx4.i00[ch]) + t.wx * (tx4.i11[ch] - tx4.i01[ch]));
          if (grad_srcs) {
            float* gs = grad_srcs + (long long)b * P.src_bs + (long long)n * P.src_ns + (long long)ch * P.HW;
            atomicAdd(gs + t.y0 * P.W + t.x0, w00 * hq[ch]);
            atomicAdd(gs + t.y0 * P.W + t.x1, w01 * hq[ch]);
            atomicAdd(gs + t.y1 * P.W + t.x0, w10 * hq[ch]);
            atomicAdd(gs + t.y1 * P.W + t.x1, w11 * hq[ch]);
          }
        }
        if (!t.gx) du = 0.f;
        if (!t.gy) dv = 0.f;
        dD += project_adjoint(g, cam, pose, du, dv, gp[n]);
      }
    }
    if (in_img) {
      float* o = (k == 0) ? grad_d0 + (long long)b * P.depth_bs[0]
                          : ((k == 1) ? dD1 : (k == 2 ? dD2 : dD3)) + (long long)b * P.HW;
      o[py * P.W + px] = dD;
    }
  }

  // ---- per-tile pose-gradient partials (fp64: sums of terms of both signs) ----
  const int blk = (b * P.tiles_y + blockIdx.y) * P.tiles_x + blockIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int n = 0; n < NS; ++n)
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      double s = warp_sum((double)gp[n][j]);
      if (lane == 0) red[wid * (NS * 12) + n * 12 + j] = s;
    }
  __syncthreads();
  if (tid < NS * 12) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w * (NS * 12) + tid];
    pose_part[(long long)blk * (NS * 12) + tid] = s;
  }
}

// ------------------------------------------------------------------------------------------
// grad_T[b,n] from the per-tile partials: one CTA per (b,n); warp w reduces entries w, w+8.
__global__ void __launch_bounds__(kThreads)
    k_pose_final(KP P, const double* __restrict__ pose_part, float* __restrict__ grad_T) {
  const int bn = blockIdx.x, b = bn / P.N, n = bn % P.N;
  const int tiles = P.tiles_x * P.tiles_y, nv = P.N * 12;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = wid; j < 16; j += kThreads / 32) {
    if (j < 12) {
      double acc = 0.0;
      for (int t = lane; t < tiles; t += 32) acc += pose_part[((long long)b * tiles + t) * nv + n * 12 + j];
      acc = warp_sum(acc);
      if (lane == 0) {
        int row, col;
        if (j < 9) { row = j / 3; col = j % 3; } else { row = j - 9; col = 3; }
        grad_T[(long long)bn * 16 + row * 4 + col] = (float)acc;
      }
    } else if (lane == 0) {
      grad_T[(long long)bn * 16 + 12 + (j - 12)] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Adjoint of upsample_depth in gather form: low-res texel (i, j) sums every full-res pixel whose
// 2x2 bilinear footprint contains it.  blockIdx.y = b * (S-1) + (k-1).
__global__ void __launch_bounds__(kThreads)
    k_depth_gather(KP P, const float* __restrict__ dD1, const float* __restrict__ dD2, const float* __restrict__ dD3,
                   float* g1, float* g2, float* g3) {
  const int b = blockIdx.y / (P.S - 1), k = blockIdx.y % (P.S - 1) + 1;
  const int hk = P.h[k], wk = P.w[k];
  const int idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= hk * wk) return;
  const int i = idx / wk, j = idx - i * wk;
  const float* dD = ((k == 1) ? dD1 : (k == 2 ? dD2 : dD3)) + (long long)b * P.HW;
  float* out = ((k == 1) ? g1 : (k == 2 ? g2 : g3)) + (long long)b * P.depth_bs[k];
  const float ry = P.ry[k], rx = P.rx[k];
  // full-res rows v with source coordinate in (i-1, i+1): conservative bounds, exact test inside
  int v_lo = imax(0, (int)floorf(((float)i - 0.5f) / ry - 0.5f) - 1);
  int v_hi = imin(P.H - 1, (int)ceilf(((float)i + 1.5f) / ry - 0.5f) + 1);
  int u_lo = imax(0, (int)floorf(((float)j - 0.5f) / rx - 0.5f) - 1);
  int u_hi = imin(P.W - 1, (int)ceilf(((float)j + 1.5f) / rx - 0.5f) + 1);
  if (i == hk - 1) v_hi = P.H - 1;   // clamped border rows / columns all land on the last texel
  if (j == wk - 1) u_hi = P.W - 1;
  float acc = 0.f;
  for (int v = v_lo; v <= v_hi; ++v) {
    Axis ay = upsample_axis(v, ry, hk);
    float wy = ((ay.i0 == i) ? (1.0f - ay.w1) : 0.f) + ((ay.i1 == i) ? ay.w1 : 0.f);
    if (wy == 0.f) continue;
    float row = 0.f;
    for (int u = u_lo; u <= u_hi; ++u) {
      Axis ax = upsample_axis(u, rx, wk);
      float wx = ((ax.i0 == j) ? (1.0f - ax.w1) : 0.f) + ((ax.i1 == j) ? ax.w1 : 0.f);
      if (wx != 0.f) row = fmaf(wx, __ldg(dD + v * P.W + u), row);
    }
    acc = fmaf(wy, row, acc);
  }
  out[idx] = acc;
}

// ------------------------------------------------------------------------------------------
// Smoothness adjoint.  Pass a: s_p = dL/dd*_p per pixel (stored) and partial sums of s_p * d_p.
__global__ void __launch_bounds__(kThreads)
    k_smooth_bwd_a(KP P, const float* __restrict__ grad_loss, const double* __restrict__ saved_mean, const float* p1,
                   const float* p2, const float* p3, float* s1, float* s2, float* s3, float* s0,
                   double* __restrict__ sd_part) {
  __shared__ double sm[kThreads / 32];
  const int bk = blockIdx.y, b = bk / P.S, k = bk % P.S;
  const int hk = P.h[k], wk = P.w[k], n = hk * wk;
  const float inv = (float)(1.0 / (saved_mean[bk] + (double)P.eps_disp));
  const float* D = P.depth[k] + (long long)b * P.depth_bs[k];
  float* sf = ((k == 0) ? s0 : (k == 1 ? s1 : (k == 2 ? s2 : s3))) + (long long)b * n;
  const float* I;
  long long cs;
  if (k == 0) { I = P.tgt + (long long)b * P.tgt_bs; cs = P.HW; }
  else { I = ((k == 1) ? p1 : (k == 2 ? p2 : p3)) + (long long)b * 3 * n; cs = n; }
  const double lam = (double)__ldg(grad_loss) * (double)P.smooth_weight / (double)(1 << k) / (double)P.S;
  const double nx = (double)P.B * hk * (wk - 1), ny = (double)P.B * (hk - 1) * wk;
  const float cx = nx > 0 ? (float)(lam / nx) : 0.f, cy = ny > 0 ? (float)(lam / ny) : 0.f;
  double acc = 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    int y = i / wk, x = i - y * wk;
    float dr = 1.0f / __ldg(D + i);
    float d = dr * inv;
    float i0 = __ldg(I + i), i1 = __ldg(I + cs + i), i2 = __ldg(I + 2 * cs + i);
    float s = 0.f;
    auto edge = [&](int j, float cw, float sign_self) {
      float dn = (1.0f / __ldg(D + j)) * inv;
      float e = (fabsf(i0 - __ldg(I + j)) + fabsf(i1 - __ldg(I + cs + j)) + fabsf(i2 - __ldg(I + 2 * cs + j))) *
                (1.0f / 3.0f);
      // term |d_lo - d_hi| * exp(-e): this pixel is `lo` (sign_self = +1) or `hi` (-1)
      float diff = sign_self > 0 ? (d - dn) : (dn - d);
      s += sign_self * sgn(diff) * expf(-e) * cw;
    };
    if (x + 1 < wk) edge(i + 1, cx, 1.f);
    if (x > 0) edge(i - 1, cx, -1.f);
    if (y + 1 < hk) edge(i + wk, cy, 1.f);
    if (y > 0) edge(i - wk, cy, -1.f);
    sf[i] = s;
    acc += (double)(s * dr);
  }
  double v[1] = {acc};
  block_reduce_store<1, double>(v, sm, sd_part + (long long)bk * kSmoothChunks + blockIdx.x);
}

// Pass b: dL/dD_q += -(s_q / (mu+eps) - sum(s d) / (hw (mu+eps)^2)) / D_q^2
__global__ void __launch_bounds__(kThreads)
    k_smooth_bwd_b(KP P, const double* __restrict__ saved_mean, const float* s0, const float* s1, const float* s2,
                   const float* s3, const double* __restrict__ sd_part, float* g0, float* g1, float* g2, float* g3) {
  const int bk = blockIdx.y, b = bk / P.S, k = bk % P.S;
  const int n = P.h[k] * P.w[k];
  double sd = 0.0;
#pragma unroll
  for (int i = 0; i < kSmoothChunks; ++i) sd += sd_part[(long long)bk * kSmoothChunks + i];
  const double me = saved_mean[bk] + (double)P.eps_disp;
  const float inv = (float)(1.0 / me);
  const float corr = (float)(sd / ((double)n * me * me));
  const float* D = P.depth[k] + (long long)b * P.depth_bs[k];
  const float* sf = ((k == 0) ? s0 : (k == 1 ? s1 : (k == 2 ? s2 : s3))) + (long long)b * n;
  float* g = ((k == 0) ? g0 : (k == 1 ? g1 : (k == 2 ? g2 : g3))) + (long long)b * P.depth_bs[k];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    float dr = 1.0f / __ldg(D + i);
    float dd = sf[i] * inv - corr;
    g[i] += -dd * dr * dr;
  }
}

// ------------------------------------------------------------------------------------------
static inline int div_up(int a, int b) { return (a + b - 1) / b; }

cudaError_t launch_backward(const KP& P, const BwdBuffers& Wk, const float* grad_loss, const uint8_t* sel,
                            const double* saved, float* const* grad_depth, float* grad_T, float* grad_srcs,
                            cudaStream_t st) {
  cudaError_t e;
  if (grad_srcs) {
    e = cudaMemsetAsync(grad_srcs, 0, sizeof(float) * (size_t)P.B * P.N * 3 * P.HW, st);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(P.tiles_x, P.tiles_y, P.B);
  {
  ScopedKernelTimer tm(2, st);
  if (P.N == 1)
    k_photo_bwd<1><<<grid, kThreads, 0, st>>>(P, grad_loss, sel, saved, grad_depth[0], Wk.dDhat[1], Wk.dDhat[2],
                                              Wk.dDhat[3], grad_srcs, Wk.pose_part);
  else
    k_photo_bwd<2><<<grid, kThreads, 0, st>>>(P, grad_loss, sel, saved, grad_depth[0], Wk.dDhat[1], Wk.dDhat[2],
                                              Wk.dDhat[3], grad_srcs, Wk.pose_part);
  }
  k_pose_final<<<P.B * P.N, kThreads, 0, st>>>(P, Wk.pose_part, grad_T);
  if (P.S > 1) {
    int blocks = div_up(P.h[1] * P.w[1], kThreads);
    k_depth_gather<<<dim3(blocks, P.B * (P.S - 1)), kThreads, 0, st>>>(
        P, Wk.dDhat[1], Wk.dDhat[2], Wk.dDhat[3], P.S > 1 ? grad_depth[1] : nullptr, P.S > 2 ? grad_depth[2] : nullptr,
        P.S > 3 ? grad_depth[3] : nullptr);
  }
  const double* saved_mean = saved + (long long)P.B * P.N * P.S * kSavedPerFrame;
  // smoothness adjoint; its "+=" into grad_depth is ordered after the kernels above by the stream
  e = launch_tgt_pyramid(P, Wk.pyr, st);
  if (e != cudaSuccess) return e;
  dim3 sg(kSmoothChunks, P.B * P.S);
  k_smooth_bwd_a<<<sg, kThreads, 0, st>>>(P, grad_loss, saved_mean, Wk.pyr[1], Wk.pyr[2], Wk.pyr[3], Wk.s_field[1],
                                          Wk.s_field[2], Wk.s_field[3], Wk.s_field[0], Wk.sd_part);
  k_smooth_bwd_b<<<sg, kThreads, 0, st>>>(P, saved_mean, Wk.s_field[0], Wk.s_field[1], Wk.s_field[2], Wk.s_field[3],
                                          Wk.sd_part, grad_depth[0], P.S > 1 ? grad_depth[1] : nullptr,
                                          P.S > 2 ? grad_depth[2] : nullptr, P.S > 3 ? grad_depth[3] : nullptr);
  return cudaGetLastError();
}

}  // namespace colvo
