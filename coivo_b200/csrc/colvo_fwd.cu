// Forward kernels of the ColVO photometric-loss path,
// hand-written for sm_100a.  oracle/photometric.py is the arithmetic contract.
//
//   k_prepass       box-averaged target pyramid + partial sums of 1/D per (b,k)      (row 9)
//   k_warp_stats    warp every (b,n,k) frame once, fp64 LCC sums, valid mask          (rows 0-5)
//   k_lcc_solve     (a, b) per warped frame                                           (row 5)
//   k_photo_fwd     per 32x8 tile: identity + re-projection candidates, SSIM+L1,
//                   min-reprojection / auto-mask, loss partials, dL/da, dL/db         (rows 0-8)
//   k_smooth_fwd    edge-aware smoothness partials (+ its adjoint field when saving)  (row 9)
//   k_finalize_fwd  deterministic final sums -> loss, G_a, G_b, sum s*d               (row 10)
#include "colvo_kernels.cuh"
#include "colvo_photo_fwd.cuh"

// occupancy knobs (CTAs per SM the register allocator must allow) -- tuned on B200, see DESIGN.md
#ifndef COLVO_MINB_STATS
#define COLVO_MINB_STATS 4
#endif
#ifndef COLVO_Y_REGS        // 1: keep the 3x3 target window of the own pixel in registers (27 regs)
#define COLVO_Y_REGS 0
#endif

namespace colvo {

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// block j of a "(b, k, chunk)" launch -> its coordinates; chunks per scale are P.sm_chunks[k]
__device__ __forceinline__ void decode_chunk(const KP& P, int j, int& b, int& k, int& c) {
  int tot = 0;
#pragma unroll
  for (int i = 0; i < kMaxS; ++i) tot += (i < P.S) ? P.sm_chunks[i] : 0;
  b = j / tot;
  int r = j - b * tot;
  k = 0;
#pragma unroll
  for (int i = 0; i < kMaxS - 1; ++i) {
    if (i < P.S - 1 && k == i && r >= P.sm_chunks[i]) { r -= P.sm_chunks[i]; k = i + 1; }
  }
  c = r;
}
static inline int total_chunks(const KP& P) {
  int t = 0;
  for (int k = 0; k < P.S; ++k) t += P.sm_chunks[k];
  return t;
}

// ------------------------------------------------------------------------------------------
// blocks [0, n_pyr): target pyramid, one thread per output texel (block ranges per scale: pyr_off[k]);
// blocks [n_pyr, ...): sum of 1/D chunks
struct PyrOff { int off[kMaxS + 1]; };
template <bool PK>
__global__ void __launch_bounds__(kThreads)
    k_prepass(KP P, int n_pyr, PyrOff po, float* p1, float* p2, float* p3, double* __restrict__ disp_part) {
  __shared__ double sm[kThreads / 32];
  if ((int)blockIdx.x < n_pyr) {
    int k = 1;
    while (k + 1 < P.S && (int)blockIdx.x >= po.off[k + 1]) ++k;
    float* out = (k == 1) ? p1 : (k == 2 ? p2 : p3);
    const int hk = P.h[k], wk = P.w[k], f = 1 << k;
    const int total = P.B * 3 * hk * wk;
    const int i = (blockIdx.x - po.off[k]) * kThreads + threadIdx.x;
    if (i >= total) return;
    const int x = i % wk, r = i / wk;
    const int y = r % hk, bc = r / hk;
    const int b = bc / 3, c = bc - 3 * b;
    const Img<PK> im = img_at<PK>(P, P.tgt, b * P.tgt_bf);
    const int o = (y * f) * P.W + x * f;
    float s = 0.f;
    for (int dy = 0; dy < f; ++dy) {
      float rs = 0.f;
      for (int dx = 0; dx < f; ++dx) {
        if constexpr (PK) {
          float v[3];
          im.load3(o + dy * P.W + dx, v);
          rs += (c == 0) ? v[0] : (c == 1 ? v[1] : v[2]);
        } else {
          rs += __ldg(im.p + (c * P.HW + o + dy * P.W + dx));
        }
      }
      s += rs;
    }
    out[i] = s * (1.0f / (float)(f * f));
    return;
  }
  int b, k, c;
  decode_chunk(P, blockIdx.x - n_pyr, b, k, c);
  const int n = P.h[k] * P.w[k], C = P.sm_chunks[k];
  const float* D = P.depth[k] + (long long)b * P.depth_bs[k];
  double acc = 0.0;
  for (int i = c * kThreads + threadIdx.x; i < n; i += C * kThreads) acc += (double)f_rcp(__ldg(D + i));
  double v[1] = {acc};
  block_reduce_store<1, double>(v, sm, disp_part + ((long long)(b * P.S + k)) * kSmoothMaxChunks + c);
}

// ------------------------------------------------------------------------------------------
// One CTA = one (b, k) and a chunk of pixels; both sources are warped by the same thread so the
// ray, the up-sampled depth and the target pixel are loaded once.  The raw warped frames are also
// written out ([B,N,S,3,H,W] scratch): the tile kernel needs them with a halo, and re-warping there
// costs more issue slots than the 12 B/pixel round trip costs bandwidth on this ALU-bound path.
template <int NS, bool GEO, bool PK>
__global__ void __launch_bounds__(kThreads, COLVO_MINB_STATS)
    k_warp_stats(KP P, double* __restrict__ part, uint8_t* __restrict__ valid_out, float4* __restrict__ iw_out,
                 float4* __restrict__ geo_out) {
  constexpr int NA = GEO ? kStatVals : 5;      // accumulators per source (the 6th only with the geometric term)
  constexpr int NV = NA * NS;
  __shared__ double sm[(kThreads / 32) * NV];
  const int bk = blockIdx.y, k = bk % P.S, b = bk / P.S;
  const Cam cam = load_cam(P, b);
  const float* Dk = P.depth[k] + (long long)b * P.depth_bs[k];
  const Img<PK> tg = img_at<PK>(P, P.tgt, b * P.tgt_bf);
  Pose pose[NS];
#pragma unroll
  for (int n = 0; n < NS; ++n) pose[n] = load_pose(P, b, n);
  int pix = blockIdx.x * (kThreads * kStatPPT) + threadIdx.x;
  int py = pix / P.W, px = pix - py * P.W;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
#pragma unroll 2
  for (int i = 0; i < kStatPPT; ++i) {
    if (pix < P.HW) {
      const float rx = ray_x(px, cam), ry = ray_y(py, cam);
      const float D = depth_at(P, Dk, k, px, py);
      float y0, y1, y2;
      if constexpr (PK) {
        float yv[3];
        tg.load3(pix, yv);
        y0 = yv[0]; y1 = yv[1]; y2 = yv[2];
      } else {
        y0 = __ldg(tg.p + pix); y1 = __ldg(tg.p + P.HW + pix); y2 = __ldg(tg.p + 2 * P.HW + pix);
      }
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        // (built per use on purpose: a hoisted 64-bit base costs this 64-register kernel more than re-deriving it)
        const Img<PK> src = img_at<PK>(P, P.srcs, b * P.src_bf + n * P.src_nf);
        Geo g; Taps t; Texels tx; float x[3];
        warp_sample<PK>(P, src, cam, pose[n], rx, ry, D, g, t, tx, x);
        const int bnk = (b * P.N + n) * P.S + k;
        if (valid_out) valid_out[(long long)bnk * P.HW + pix] = g.valid ? 1 : 0;
        // raw warped frame, re-used by k_photo_fwd instead of warping again (+halo): one 16-byte texel
        if (iw_out) iw_out[(long long)bnk * P.HW + pix] = make_float4(x[0], x[1], x[2], 0.f);
        // the projection itself, for the backward (valid rides in the mantissa LSB of the depth)
        if (geo_out)
          geo_out[(long long)bnk * P.HW + pix] =
              make_float4(g.u, g.v, g.iz, __uint_as_float((__float_as_uint(D) & ~1u) | (g.valid ? 1u : 0u)));
        if (g.valid) {
          acc[NA * n + 0] += 3.0;
          acc[NA * n + 1] += (double)(x[0] + x[1] + x[2]);
          acc[NA * n + 2] += (double)(y0 + y1 + y2);
          acc[NA * n + 3] += (double)(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
          acc[NA * n + 4] += (double)(x[0] * y0 + x[1] * y1 + x[2] * y2);
          if (GEO) {               // geometric consistency (f-2): per-pixel, so it lives in this pass
            float d4[4], dZ, dS;
            const float ds = sample_plane(P.src_depth + (long long)(b * P.N + n) * P.HW, t, P.W, d4);
            acc[NA * n + (NA - 1)] += (double)geo_diff(g.Zp, ds, dZ, dS);
          }
        }
      }
    }
    pix += kThreads;
    px += kThreads;
    while (px >= P.W) { px -= P.W; ++py; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = warp_sum(acc[i]);
    if (lane == 0) sm[wid * NV + i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += sm[w * NV + threadIdx.x];
    const int n = threadIdx.x / NA, j = threadIdx.x - NA * n;
    const int bnk = (b * P.N + n) * P.S + k;
    part[((long long)bnk * gridDim.x + blockIdx.x) * kStatVals + j] = s;
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    k_lcc_solve(KP P, const double* __restrict__ part, int chunks, float* __restrict__ ab, double* __restrict__ saved) {
  const int bnk = blockIdx.x, lane = threadIdx.x;
  double s[5] = {0, 0, 0, 0, 0};
  if (P.flags & 1u) {
    for (int c = lane; c < chunks; c += 32) {
#pragma unroll
      for (int j = 0; j < 5; ++j) s[j] += part[((long long)bnk * chunks + c) * kStatVals + j];
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
// 3x3-window helpers of the consistency sweep's tile kernel (k_consistency_pe below): one window per
// thread over a 32x8 tile (+1 halo) in shared memory.  The training loss uses the strip kernel in
// colvo_photo_fwd.cuh instead.
constexpr int kFH = kTileH + 2, kFW = kTileW + 2;   // tile + 1-pixel SSIM halo
constexpr int kFN = kFH * kFW;

// The 3x3 target window of the own pixel: either 27 registers or re-read from the target tile.
struct YWin {
#if COLVO_Y_REGS
  float v[3][9];
#endif
  const float* ys;      // [3][kFN] target tile in shared memory
  int o;                // ty*kFW + tx
  float mu[3], sg[3];   // window mean and variance of the target
  __device__ __forceinline__ float at(int c, int j) const {
#if COLVO_Y_REGS
    return v[c][j];
#else
    return ys[c * kFN + o + (j / 3) * kFW + (j % 3)];
#endif
  }
};

__device__ __forceinline__ float pe_own(const float* __restrict__ xb /* [3][kFN] */, const YWin& y, float a, float b,
                                        const KP& P, float* dpa, float* dpb, bool want_cf, Coef (&cf)[3]) {
  float pe = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = 0.f, sxx = 0.f, sxy = 0.f, xc = 0.f, yc = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const float v = xb[c * kFN + y.o + (j / 3) * kFW + (j % 3)];
      const float w = y.at(c, j);
      if (j == 4) { xc = v; yc = w; }
      s += v;
      sxx = fmaf(v, v, sxx);
      sxy = fmaf(v, w, sxy);
    }
    const float i9 = 1.0f / 9.0f;
    pe += pe_channel(s * i9, sxx * i9, sxy * i9, y.mu[c], y.sg[c], xc, yc, a, b, P.alpha, P.c1, P.c2, dpa, dpb, want_cf,
                     cf[c]);
  }
  return pe * (1.0f / 3.0f);
}
__device__ __forceinline__ float pe_own(const float* __restrict__ xb, const YWin& y, float a, float b, const KP& P) {
  Coef unused[3];
  return pe_own(xb, y, a, b, P, nullptr, nullptr, false, unused);
}
__device__ __forceinline__ void ywin_init(YWin& y, const float* ys, int own) {
  y.ys = ys;
  y.o = own;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float v = ys[c * kFN + own + (j / 3) * kFW + (j % 3)];
#if COLVO_Y_REGS
      y.v[c][j] = v;
#endif
      s += v;
      ss = fmaf(v, v, ss);
    }
    y.mu[c] = s * (1.0f / 9.0f);
    y.sg[c] = ss * (1.0f / 9.0f) - y.mu[c] * y.mu[c];
  }
}

// ------------------------------------------------------------------------------------------
// Smoothness (row 9).  Each pixel visits its four edges: the right/down ones give the loss, all
// four give s_p = dL/dd*_p (for grad_loss = 1), which the backward only has to rescale.
template <bool SAVE, bool PK>
__global__ void __launch_bounds__(kThreads)
    k_smooth_fwd(KP P, const double* __restrict__ disp_part, const float* p1, const float* p2, const float* p3,
                 double* __restrict__ smooth_part, double* __restrict__ saved_scale, float* s0, float* s1, float* s2,
                 float* s3) {
  __shared__ double sm[(kThreads / 32) * 3];
  __shared__ double mean_s;
  int b, k, c;
  decode_chunk(P, blockIdx.x, b, k, c);
  const int bk = b * P.S + k;
  const int hk = P.h[k], wk = P.w[k], n = hk * wk, C = P.sm_chunks[k];
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int i = threadIdx.x; i < C; i += 32) s += disp_part[(long long)bk * kSmoothMaxChunks + i];
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      mean_s = s / (double)n;
      if (SAVE && c == 0) saved_scale[bk * kSavedPerScale + 0] = mean_s;
    }
  }
  __syncthreads();
  const float inv = (float)(1.0 / (mean_s + (double)P.eps_disp));
  const float* D = P.depth[k] + (long long)b * P.depth_bs[k];
  // image of this scale: the target itself (either storage format) at k = 0, the fp32 pyramid above
  const Img<PK> I0 = img_at<PK>(P, P.tgt, b * P.tgt_bf);
  Img<false> Ik;
  Ik.p = (k == 0) ? nullptr : ((k == 1) ? p1 : (k == 2 ? p2 : p3)) + (long long)b * 3 * n;
  Ik.HW = n;
  float* sf = nullptr;
  if (SAVE) sf = ((k == 0) ? s0 : (k == 1 ? s1 : (k == 2 ? s2 : s3))) + (long long)b * n;
  const double lam = (double)P.smooth_weight / (double)(1 << k) / (double)P.S;
  const double nx = (double)P.B * hk * (wk - 1), ny = (double)P.B * (hk - 1) * wk;
  const float cx = nx > 0 ? (float)(lam / nx) : 0.f, cy = ny > 0 ? (float)(lam / ny) : 0.f;
  double acc[3] = {0.0, 0.0, 0.0};
  auto run = [&](const auto& img) {
    for (int i = c * kThreads + threadIdx.x; i < n; i += C * kThreads) {
      const int y = i / wk, x = i - y * wk;
      const float dr = f_rcp(__ldg(D + i));
      const float d = dr * inv;
      float iv[3];
      img.load3(i, iv);
      const float i0 = iv[0], i1 = iv[1], i2 = iv[2];
      float s = 0.f;
      auto edge = [&](int j) -> float2 {   // (d_i - d_j, exp(-mean_c |I_i - I_j|))
        float dn = f_rcp(__ldg(D + j)) * inv;
        float jv[3];
        img.load3(j, jv);
        float e = (fabsf(i0 - jv[0]) + fabsf(i1 - jv[1]) + fabsf(i2 - jv[2])) * (1.0f / 3.0f);
        return make_float2(d - dn, __expf(-e));   // e in [0, 1]: MUFU.EX2 path, rel. error ~1e-6
      };
      if (x + 1 < wk) {
        float2 t = edge(i + 1);
        acc[0] += (double)(fabsf(t.x) * t.y);
        if (SAVE) s += sgn(t.x) * t.y * cx;
      }
      if (y + 1 < hk) {
        float2 t = edge(i + wk);
        acc[1] += (double)(fabsf(t.x) * t.y);
        if (SAVE) s += sgn(t.x) * t.y * cy;
      }
      if (SAVE) {
        if (x > 0) { float2 t = edge(i - 1); s += sgn(t.x) * t.y * cx; }
        if (y > 0) { float2 t = edge(i - wk); s += sgn(t.x) * t.y * cy; }
        sf[i] = s;
        acc[2] += (double)(s * dr);
      }
    }
  };
  if (k == 0) run(I0);
  else run(Ik);
  block_reduce_store<3, double>(acc, sm, smooth_part + ((long long)bk * kSmoothMaxChunks + c) * 3);
}

// ------------------------------------------------------------------------------------------
// block 0: the scalar loss.  blocks [1, 1+BNS): G_a, G_b of warped frame bnk (dL/da, dL/db).
// blocks [1+BNS, 1+BNS+B*S): sum_p s_p d_p of (b, k) for the smoothness adjoint.
__global__ void __launch_bounds__(kThreads)
    k_finalize_fwd(KP P, const double* __restrict__ loss_part, const double* __restrict__ g_part,
                   const double* __restrict__ smooth_part, const double* __restrict__ stat_part, int stat_chunks,
                   float* __restrict__ loss, double* __restrict__ saved_frame, double* __restrict__ saved_scale,
                   int need_g) {
  __shared__ double sm[(kThreads / 32) * 2];
  __shared__ double wk_s[kMaxS][2];
  const int tiles = P.ftiles_x * P.ftiles_y;    // k_photo_fwd's tiling
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int BNS = P.B * P.N * P.S;
  if (blockIdx.x == 0) {
    if (threadIdx.x < kMaxS) {
      const int k = threadIdx.x;
      double wx = 0.0, wy = 0.0;
      if (k < P.S) {
        const double lam = (double)P.smooth_weight / (double)(1 << k);
        const double nx = (double)P.B * P.h[k] * (P.w[k] - 1), ny = (double)P.B * (P.h[k] - 1) * P.w[k];
        wx = nx > 0 ? lam / nx : 0.0;
        wy = ny > 0 ? lam / ny : 0.0;
      }
      wk_s[k][0] = wx;
      wk_s[k][1] = wy;
    }
    __syncthreads();
    double acc[2] = {0.0, 0.0};                       // photometric sum, weighted smoothness (+ geometric) sum
    for (int i = threadIdx.x; i < P.B * tiles; i += kThreads) acc[0] += loss_part[i];
    if (P.src_depth) {                                // L_geo,k = sum of diffs / (B N HW), weight geo_weight
      const double wg = (double)P.geo_weight / ((double)P.B * (double)P.N * (double)P.HW);
      for (int i = threadIdx.x; i < BNS * stat_chunks; i += kThreads) acc[1] += stat_part[(long long)i * kStatVals + 5] * wg;
    }
    for (int i = threadIdx.x; i < P.B * P.S * kSmoothMaxChunks; i += kThreads) {
      const int bk = i / kSmoothMaxChunks, c = i - bk * kSmoothMaxChunks, k = bk % P.S;
      if (c < P.sm_chunks[k])
        acc[1] += smooth_part[(long long)i * 3 + 0] * wk_s[k][0] + smooth_part[(long long)i * 3 + 1] * wk_s[k][1];
    }
    for (int j = 0; j < 2; ++j) {
      double s = warp_sum(acc[j]);
      if (lane == 0) sm[wid * 2 + j] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double o0 = 0.0, o1 = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) { o0 += sm[w * 2]; o1 += sm[w * 2 + 1]; }
      *loss = (float)((o0 / ((double)P.B * (double)P.HW) + o1) / (double)P.S);
    }
    return;
  }
  if ((int)blockIdx.x <= BNS) {
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
      saved_frame[(long long)bnk * kSavedPerFrame + 6 + threadIdx.x] = s / ((double)P.S * (double)P.B * (double)P.HW);
    }
    return;
  }
  // sum_p s_p d_p
  const int bk = blockIdx.x - 1 - BNS, k = bk % P.S;
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int c = lane; c < P.sm_chunks[k]; c += 32) s += smooth_part[((long long)bk * kSmoothMaxChunks + c) * 3 + 2];
    s = warp_sum(s);
    if (lane == 0) saved_scale[bk * kSavedPerScale + 1] = s;
  }
}

// ------------------------------------------------------------------------------------------
// Consistency sweep (BASELINE config 5): per pair, mean pe over valid pixels.  N = 1, S = 1.
__global__ void __launch_bounds__(kThreads, 2)
    k_consistency_pe(KP P, const float* __restrict__ ab, double* __restrict__ pe_part) {
  __shared__ float ys[3 * kFN];
  __shared__ float xs[3 * kFN];
  __shared__ unsigned char vs[kFN];
  __shared__ double red[(kThreads / 32) * 2];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int b = blockIdx.z, x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int px = x0 + tx, py = y0 + ty;
  const bool in_img = (px < P.W) && (py < P.H);
  const int own = ty * kFW + tx;
  const float* tg = static_cast<const float*>(P.tgt) + b * P.tgt_bf * P.frame_el;
  const Img<false> src = img_at<false>(P, P.srcs, b * P.src_bf);
  const float* Dk = P.depth[0] + (long long)b * P.depth_bs[0];
  const Cam cam = load_cam(P, b);
  const Pose pose = load_pose(P, b, 0);
  for (int idx = tid; idx < kFN; idx += kThreads) {
    int r = idx / kFW, c = idx - r * kFW;
    int ry = y0 - 1 + r, rx = x0 - 1 + c;
    int gy = reflect_clamp(ry, P.H), gx = reflect_clamp(rx, P.W);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ys[ch * kFN + idx] = __ldg(tg + (ch * P.HW + gy * P.W + gx));
    if (ry <= P.H && rx <= P.W) {
      Geo g; Taps t; Texels tx4; float x[3];
      warp_sample<false>(P, src, cam, pose, ray_x(gx, cam), ray_y(gy, cam), __ldg(Dk + gy * P.W + gx), g, t, tx4, x);
      xs[idx] = x[0];
      xs[kFN + idx] = x[1];
      xs[2 * kFN + idx] = x[2];
      vs[idx] = g.valid ? 1 : 0;
    }
  }
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  if (in_img && vs[own + kFW + 1]) {
    YWin yw;
    ywin_init(yw, ys, own);
    float pe = pe_own(xs, yw, __ldg(ab + 2 * b), __ldg(ab + 2 * b + 1), P);
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
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = 0; j < 2; ++j) {
    double s = warp_sum(acc[j]);
    if (lane == 0) sm[wid * 2 + j] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double r0 = 0.0, r1 = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { r0 += sm[w * 2]; r1 += sm[w * 2 + 1]; }
    out[4 * b + 0] = (float)(r1 > 0.0 ? r0 / r1 : 0.0);
    out[4 * b + 1] = ab[2 * b];
    out[4 * b + 2] = ab[2 * b + 1];
    out[4 * b + 3] = (float)(r1 / (double)P.HW);
  }
}

// ------------------------------------------------------------------------------------------
cudaError_t launch_forward(const KP& P, const FwdBuffers& Wk, float* loss, float* ab, uint8_t* valid, uint8_t* sel,
                           const SavedView& sv, cudaStream_t st) {
  const bool lcc = (P.flags & 1u) != 0;
  const bool save = (P.flags & 4u) != 0;
  const int need_g = (save && lcc && !(P.flags & 2u)) ? 1 : 0;
  const int BNS = P.B * P.N * P.S;
  const int chunks = P.B * total_chunks(P);
  PyrOff po;
  int n_pyr = 0;
  po.off[0] = 0;
  for (int k = 1; k <= kMaxS; ++k) {
    if (k < kMaxS) po.off[k] = n_pyr;
    else po.off[kMaxS] = n_pyr;
    if (k < P.S) n_pyr += div_up(P.B * 3 * P.h[k] * P.w[k], kThreads);
  }
  const bool pk = (P.flags & 16u) != 0;
  if (pk) k_prepass<true><<<n_pyr + chunks, kThreads, 0, st>>>(P, n_pyr, po, Wk.pyr[1], Wk.pyr[2], Wk.pyr[3], Wk.disp_part);
  else k_prepass<false><<<n_pyr + chunks, kThreads, 0, st>>>(P, n_pyr, po, Wk.pyr[1], Wk.pyr[2], Wk.pyr[3], Wk.disp_part);
  {
    ScopedKernelTimer tm(3, st);
    dim3 g(Wk.stat_chunks, P.B * P.S);
    const bool geo = P.src_depth != nullptr;
    auto run = [&](auto kern) { kern<<<g, kThreads, 0, st>>>(P, Wk.stat_part, valid, Wk.iw, save ? sv.geo : nullptr); };
    if (P.N == 1) {
      if (geo) { if (pk) run(k_warp_stats<1, true, true>); else run(k_warp_stats<1, true, false>); }
      else { if (pk) run(k_warp_stats<1, false, true>); else run(k_warp_stats<1, false, false>); }
    } else {
      if (geo) { if (pk) run(k_warp_stats<2, true, true>); else run(k_warp_stats<2, true, false>); }
      else { if (pk) run(k_warp_stats<2, false, true>); else run(k_warp_stats<2, false, false>); }
    }
  }
  k_lcc_solve<<<BNS, 32, 0, st>>>(P, Wk.stat_part, Wk.stat_chunks, ab, save ? sv.frame : nullptr);
  dim3 grid(P.ftiles_x, P.ftiles_y, P.B);
  {
    ScopedKernelTimer tm(1, st);
    float4* co = save ? reinterpret_cast<float4*>(sv.coef) : nullptr;
    // opting in to > 48 KB of dynamic shared memory is a per-function, per-device attribute: cheap and idempotent
    auto run = [&](auto kern, size_t smem) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, kFwdThreads, smem, st>>>(P, ab, sel, Wk.loss_part, Wk.g_part, need_g, co, Wk.iw);
    };
    if (P.N == 1) { if (pk) run(k_photo_fwd<1, true>, photo_fwd_smem<1>()); else run(k_photo_fwd<1, false>, photo_fwd_smem<1>()); }
    else { if (pk) run(k_photo_fwd<2, true>, photo_fwd_smem<2>()); else run(k_photo_fwd<2, false>, photo_fwd_smem<2>()); }
  }
  {
    auto run = [&](auto kern, bool sv_on) {
      kern<<<chunks, kThreads, 0, st>>>(P, Wk.disp_part, Wk.pyr[1], Wk.pyr[2], Wk.pyr[3], Wk.smooth_part, sv_on ? sv.scale : nullptr,
                                        sv_on ? sv.s_field[0] : nullptr, sv_on ? sv.s_field[1] : nullptr,
                                        sv_on ? sv.s_field[2] : nullptr, sv_on ? sv.s_field[3] : nullptr);
    };
    if (save) { if (pk) run(k_smooth_fwd<true, true>, true); else run(k_smooth_fwd<true, false>, true); }
    else { if (pk) run(k_smooth_fwd<false, true>, false); else run(k_smooth_fwd<false, false>, false); }
  }
  const int nfin = 1 + (save ? BNS + P.B * P.S : 0);
  k_finalize_fwd<<<nfin, kThreads, 0, st>>>(P, Wk.loss_part, Wk.g_part, Wk.smooth_part, Wk.stat_part, Wk.stat_chunks,
                                            loss, sv.frame, sv.scale, need_g);
  return cudaGetLastError();
}

cudaError_t launch_consistency(const KP& P, double* stat_part, int stat_chunks, double* pe_part, float* ab, float* out,
                               cudaStream_t st) {
  k_warp_stats<1, false, false><<<dim3(stat_chunks, P.B), kThreads, 0, st>>>(P, stat_part, nullptr, nullptr, nullptr);
  k_lcc_solve<<<P.B, 32, 0, st>>>(P, stat_part, stat_chunks, ab, nullptr);
  k_consistency_pe<<<dim3(P.tiles_x, P.tiles_y, P.B), kThreads, 0, st>>>(P, ab, pe_part);
  k_consistency_final<<<P.B, kThreads, 0, st>>>(P, pe_part, ab, out);
  return cudaGetLastError();
}

}  // namespace colvo
