// k_photo_fwd: the fused forward tile kernel, sm_100a.
//
// One CTA = 4 warps = one 32 x 12 tile of windows of one triplet.  Every warp walks DOWN a strip of
// 3 window rows with one window column per lane: for each data row it forms the horizontal 3-sums
// of x, x^2 and x*y of both warped frames (texels (x0,x1,x2,-) read with LDS.128 from the staged
// tile), keeps the last three rows of sums in registers and adds them vertically -- the separable
// form of the 3x3 SSIM window, ~2.4x fewer issue slots than summing 9 taps per window.  Both sources
// of a scale are evaluated together, so min-reprojection is decided on the spot and the adjoint
// pieces (dL/da, dL/db terms and the SSIM adjoint coefficients the backward gathers) are computed
// for the winning candidate only.  The frames of scale k+1 are fetched with 16-byte cp.async while
// scale k is evaluated.
//
// Arithmetic contract: oracle/photometric.py (ssim3x3, photometric_error, min_reprojection_automask).
#pragma once
#include "colvo_kernels.cuh"

#ifndef COLVO_FWD_M_SMEM     // 1: park the target window moments in shared memory between scales (66 KB per CTA, 3 CTAs / SM);
#define COLVO_FWD_M_SMEM 1   // 0: recompute them in every scale's walk (50 KB per CTA, 4 CTAs / SM)
#endif
#ifndef COLVO_MINB_FWD
#define COLVO_MINB_FWD ((COLVO_FWD_M_SMEM && COLVO_FWD_ROWS > 3) ? 3 : 4)
#endif

namespace colvo {

constexpr int kDW = 32 + 2;                 // data columns of a forward tile (1-pixel SSIM halo)
constexpr int kDH = kFwdTileH + 2;          // data rows
constexpr int kDN = kDW * kDH;              // texels per staged frame
constexpr int kStageRounds = (kDN + kFwdThreads - 1) / kFwdThreads;

template <int NS>
struct FwdSmem {
  float4 y[kDN];                            // target tile (y0, y1, y2, -)
  float4 x[2][NS][kDN];                     // double-buffered frames of one scale (or the raw sources)
#if COLVO_FWD_M_SMEM
  float4 m[2][kFwdTileH * 32];              // per window: (mu_y[3], var_y[0]), (var_y[1], var_y[2], best identity pe, its index)
#endif
  double red[kFwdWarps * (1 + NS * kMaxS * 2)];
};

// horizontal 3-sums of one data row at one window column
template <int NS>
struct RowH {
  float sx[NS][3], sxx[NS][3], sxy[NS][3], xc[NS][3];
};
struct RowY {
  float sy[3], syy[3], yc[3];
};

#define CV_CH(v, c) ((c) == 0 ? (v).x : ((c) == 1 ? (v).y : (v).z))

template <int NS, bool WITH_Y>
__device__ __forceinline__ void row_sums(RowH<NS>& R, RowY& Y, const float4* __restrict__ yr, const float4* __restrict__ xr0,
                                         const float4* __restrict__ xr1) {
  const float4 ya = yr[0], yb = yr[1], yc = yr[2];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Y.yc[c] = CV_CH(yb, c);
    if (WITH_Y) {
      Y.sy[c] = CV_CH(ya, c) + CV_CH(yb, c) + CV_CH(yc, c);
      Y.syy[c] = fmaf(CV_CH(yc, c), CV_CH(yc, c), fmaf(CV_CH(yb, c), CV_CH(yb, c), CV_CH(ya, c) * CV_CH(ya, c)));
    }
  }
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const float4* xr = (n == 0) ? xr0 : xr1;
    const float4 a = xr[0], b = xr[1], d = xr[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float va = CV_CH(a, c), vb = CV_CH(b, c), vd = CV_CH(d, c);
      R.sx[n][c] = va + vb + vd;
      R.sxx[n][c] = fmaf(vd, vd, fmaf(vb, vb, va * va));
      R.sxy[n][c] = fmaf(vd, CV_CH(yc, c), fmaf(vb, CV_CH(yb, c), va * CV_CH(ya, c)));
      R.xc[n][c] = vb;
    }
  }
}

// the target side of one window
struct WinY {
  float muy[3], sgy[3], yc[3];
  float sy9[3], muy2[3], k1[3], k2[3];     // 9 mu_y, 2 mu_y, mu_y^2 + C1, var_y + C2: shared by every candidate of the window
};
__device__ __forceinline__ void winy_derive(WinY& y, float c1, float c2) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    y.sy9[c] = 9.0f * y.muy[c];
    y.muy2[c] = y.muy[c] + y.muy[c];
    y.k1[c] = fmaf(y.muy[c], y.muy[c], c1);
    y.k2[c] = y.sgy[c] + c2;
  }
}
// per-candidate constants of the calibration (a, b)
struct CalK {
  float a, b, a2_9, ta_9, ta;              // a^2 / 9 and 2a / 9 act on the 9x-scaled window moments
};
__device__ __forceinline__ CalK make_calk(float a, float b) {
  CalK k;
  k.a = a; k.b = b; k.a2_9 = a * a * (1.0f / 9.0f); k.ta = 2.f * a; k.ta_9 = k.ta * (1.0f / 9.0f);
  return k;
}

// photometric error of one candidate from its 3x3 window SUMS (value only), times 3
//   pe = alpha * mean_c clamp((1 - SSIM_c)/2) + (1 - alpha) * mean_c |a x_c + b - y_c|
// The window moments stay scaled by 9 (s9 = 9 var_x, sxy9 = 9 cov_xy): the factor rides on a^2/9 and 2a/9.
__device__ __forceinline__ float pe_value3(const float (&Sx)[3], const float (&Sxx)[3], const float (&Sxy)[3],
                                           const float (&xc)[3], const WinY& y, const CalK& k, float alpha, float c1,
                                           float c2) {
  float pe = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float mu = Sx[c] * (1.0f / 9.0f);
    const float s9 = fmaf(-mu, Sx[c], Sxx[c]);
    const float sxy9 = fmaf(-mu, y.sy9[c], Sxy[c]);
    const float mut = fmaf(k.a, mu, k.b);
    const float A1 = fmaf(mut, y.muy2[c], c1);
    const float A2 = fmaf(k.ta_9, sxy9, c2);
    const float B1 = fmaf(mut, mut, y.k1[c]);
    const float B2 = fmaf(k.a2_9, s9, y.k2[c]);
    const float S = A1 * A2 * f_rcp(B1 * B2);
    const float t = __saturatef(fmaf(-0.5f, S, 0.5f));
    const float diff = fmaf(k.a, xc[c], k.b) - y.yc[c];
    pe = fmaf(alpha, t, pe);
    pe = fmaf(1.f - alpha, fabsf(diff), pe);
  }
  return pe;
}

// adjoint pieces of the winning candidate: unit-weight SSIM adjoint coefficients (ca, cb, cg) per channel
// (colvo_math.cuh::coef_from_parts) and the terms of d pe / d a, d pe / d b
__device__ __forceinline__ void pe_adjoint(const float (&Sx)[3], const float (&Sxx)[3], const float (&Sxy)[3],
                                           const float (&xc)[3], const WinY& y, const CalK& k, float alpha, float c1,
                                           float c2, float (&ca)[3], float (&cb)[3], float (&cg)[3], float& dpa,
                                           float& dpb) {
  const float a = k.a, a2 = a * a;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float mu = Sx[c] * (1.0f / 9.0f);
    const float s9 = fmaf(-mu, Sx[c], Sxx[c]);
    const float sxy9 = fmaf(-mu, y.sy9[c], Sxy[c]);
    const float s = s9 * (1.0f / 9.0f), sxy = sxy9 * (1.0f / 9.0f);
    const float mut = fmaf(a, mu, k.b);
    const float A1 = fmaf(mut, y.muy2[c], c1);
    const float A2 = fmaf(k.ta_9, sxy9, c2);
    const float B1 = fmaf(mut, mut, y.k1[c]);
    const float B2 = fmaf(k.a2_9, s9, y.k2[c]);
    const float iB = f_rcp(B1 * B2);
    const float iB1 = iB * B2, iB2 = iB * B1;
    const float r2 = A2 * iB2;
    const float S = A1 * iB1 * r2;
    const float t = fmaf(-0.5f, S, 0.5f);
    const bool in01 = (t >= 0.f) && (t <= 1.f);
    const float dmu = 2.f * iB1 * fmaf(y.muy[c], r2, -S * mut);
    const float dsx = -S * iB2;
    const float dsxy = 2.f * A1 * iB;
    const float diff = fmaf(a, xc[c], k.b) - y.yc[c];
    const float sg = sgn_scaled(1.f - alpha, diff);
    const float act = in01 ? -0.5f * alpha : 0.f;
    dpa += fmaf(act, fmaf(dmu, mu, fmaf(dsx * k.ta, s, dsxy * sxy)), sg * xc[c]);
    dpb += fmaf(act, dmu, sg);
    const float actc = act * (1.0f / 27.0f);
    const float d1 = a * dmu, d2 = a2 * dsx, d3 = a * dsxy;
    ca[c] = actc * (d1 - 2.f * mu * d2 - y.muy[c] * d3);
    cb[c] = actc * 2.f * d2;
    cg[c] = actc * d3;
  }
}

template <int NS, bool PK>
__global__ void __launch_bounds__(kFwdThreads, COLVO_MINB_FWD)
    k_photo_fwd(KP P, const float* __restrict__ ab, uint8_t* __restrict__ sel_out, double* __restrict__ loss_part,
                double* __restrict__ g_part, int need_g, float4* __restrict__ coef_out, const float4* __restrict__ iw) {
  constexpr int NV = 1 + NS * kMaxS * 2;
  extern __shared__ __align__(16) unsigned char fwd_smem_raw[];
  FwdSmem<NS>& sm = *reinterpret_cast<FwdSmem<NS>*>(fwd_smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * kFwdTileH;
  const int px = x0 + lane;
  const bool col_in = px < P.W;

  // the staged positions of this thread (reflect-padded coordinates of the tile + halo)
  int goff[kStageRounds];
#pragma unroll
  for (int j = 0; j < kStageRounds; ++j) {
    const int idx = tid + j * kFwdThreads;
    const int r = idx / kDW, c = idx - r * kDW;
    goff[j] = (idx < kDN) ? reflect_clamp(y0 - 1 + r, P.H) * P.W + reflect_clamp(x0 - 1 + c, P.W) : -1;
  }
  auto stage_scale = [&](int k, int buf) {      // warped frames of scale k -> sm.x[buf]
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      // frame base hidden from the optimiser + 32-bit offsets: one IMAD.WIDE per copy (see Img<false>::load_taps);
      // the shared-window address is formed once, the rounds are immediate offsets
      const float4* src = iw + (long long)((b * P.N + n) * P.S + k) * P.HW;
      asm volatile("" : "+l"(src));
      const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm.x[buf][n][tid]);
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j)
        if (goff[j] >= 0) cp_async16_s(sa + j * kFwdThreads * (unsigned)sizeof(float4), src + (unsigned)goff[j]);
    }
    cp_async_commit();
  };

  // target tile and the raw sources (identity candidates), either storage format.  Planar fp32 goes through
  // 4-byte cp.async straight into the texel components, so all 9 loads per position are in flight at once;
  // packed bf16 is widened on the way in (all loads first, then the stores).
  if constexpr (!PK) {
    const unsigned hw = P.HW;
    auto stage_planar = [&](const float* fb, float4* dst0) {
      asm volatile("" : "+l"(fb));
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst0 + tid);
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j)
        if (goff[j] >= 0) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
            cp_async4_s(sa + (j * kFwdThreads * 4 + ch) * (unsigned)sizeof(float), fb + ((unsigned)goff[j] + ch * hw));
        }
    };
    stage_planar(static_cast<const float*>(P.tgt) + (long long)b * P.tgt_bf * P.frame_el, sm.y);
#pragma unroll
    for (int n = 0; n < NS; ++n)
      stage_planar(static_cast<const float*>(P.srcs) + (long long)(b * P.src_bf + n * P.src_nf) * P.frame_el, sm.x[0][n]);
  } else {
    uint2 raw[1 + NS][kStageRounds];
#pragma unroll
    for (int f = 0; f < 1 + NS; ++f) {
      const uint2* p = (f == 0) ? static_cast<const uint2*>(P.tgt) + (long long)b * P.tgt_bf * P.frame_el
                                : static_cast<const uint2*>(P.srcs) + (long long)(b * P.src_bf + (f - 1) * P.src_nf) * P.frame_el;
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j) raw[f][j] = (goff[j] >= 0) ? __ldg(p + goff[j]) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int f = 0; f < 1 + NS; ++f) {
      float4* dst = (f == 0) ? sm.y : sm.x[0][f - 1];
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j)
        if (goff[j] >= 0)
          dst[tid + j * kFwdThreads] = make_float4(__uint_as_float(raw[f][j].x << 16), __uint_as_float(raw[f][j].x & 0xffff0000u),
                                                   __uint_as_float(raw[f][j].y << 16), 0.f);
    }
  }
  cp_async_commit();
  for (int i = tid; i < kFwdWarps * NV; i += kFwdThreads) sm.red[i] = 0.0;   // slots of unused scales stay 0
  pdl_trigger();
  cp_async_wait_all();          // the raw frames have landed
  __syncthreads();

  const float alpha = P.alpha, c1 = P.c1, c2 = P.c2;
  const int trow0 = wid * kFwdRows;           // first window row of this warp = its first data row in the tile

  // ---- identity candidates (raw sources, a = 1, b = 0; oracle A10) and the target window moments ----
#if !COLVO_FWD_M_SMEM
  float id_best[kFwdRows];
  int id_sel[kFwdRows];
#endif
  {
    RowH<NS> R[3];
    RowY Y[3];
#pragma unroll
    for (int j = 0; j < kFwdRows + 2; ++j) {
      const int o = (trow0 + j) * kDW + lane;
      row_sums<NS, true>(R[j % 3], Y[j % 3], sm.y + o, sm.x[0][0] + o, sm.x[0][NS - 1] + o);
      if (j >= 2) {
        WinY wy;
        float Sx[NS][3], Sxx[NS][3], Sxy[NS][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float sy = Y[0].sy[c] + Y[1].sy[c] + Y[2].sy[c];
          const float syy = Y[0].syy[c] + Y[1].syy[c] + Y[2].syy[c];
          wy.muy[c] = sy * (1.0f / 9.0f);
          wy.sgy[c] = fmaf(-wy.muy[c], wy.muy[c], syy * (1.0f / 9.0f));
          wy.yc[c] = Y[(j - 1) % 3].yc[c];
#pragma unroll
          for (int n = 0; n < NS; ++n) {
            Sx[n][c] = R[0].sx[n][c] + R[1].sx[n][c] + R[2].sx[n][c];
            Sxx[n][c] = R[0].sxx[n][c] + R[1].sxx[n][c] + R[2].sxx[n][c];
            Sxy[n][c] = R[0].sxy[n][c] + R[1].sxy[n][c] + R[2].sxy[n][c];
          }
        }
        winy_derive(wy, c1, c2);
        const CalK kid = make_calk(1.0f, 0.0f);
        // (candidates are compared as 3 * pe: the mean over channels is a common factor)
        float best = pe_value3(Sx[0], Sxx[0], Sxy[0], R[(j - 1) % 3].xc[0], wy, kid, alpha, c1, c2);
        int sel = 0;
#pragma unroll
        for (int n = 1; n < NS; ++n) {
          const float pe = pe_value3(Sx[n], Sxx[n], Sxy[n], R[(j - 1) % 3].xc[n], wy, kid, alpha, c1, c2);
          if (pe < best) { best = pe; sel = n; }
        }
#if COLVO_FWD_M_SMEM
        const int w = (trow0 + j - 2) * 32 + lane;
        sm.m[0][w] = make_float4(wy.muy[0], wy.muy[1], wy.muy[2], wy.sgy[0]);
        sm.m[1][w] = make_float4(wy.sgy[1], wy.sgy[2], best, __int_as_float(sel));
#else
        id_best[j - 2] = best;
        id_sel[j - 2] = sel;
#endif
      }
    }
  }

  // Everything above depends on the inputs only: launched programmatically, this CTA may have run it in the tail of
  // k_warp_stats / k_smooth.  The warped frames and (a, b) are needed from here on.
  pdl_wait();
  stage_scale(0, 1);
  float loss_acc = 0.f;
#pragma unroll 1
  for (int k = 0; k < P.S; ++k) {
    const int buf = (k + 1) & 1;
    cp_async_wait_all();
    __syncthreads();              // scale k landed everywhere; every warp is done with the other buffer
    if (k + 1 < P.S) stage_scale(k + 1, buf ^ 1);
    CalK cal[NS];
    float ga[NS], gb[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const int bnk = (b * P.N + n) * P.S + k;
      cal[n] = make_calk(__ldg(ab + 2 * bnk), __ldg(ab + 2 * bnk + 1));
      ga[n] = gb[n] = 0.f;
    }
    const long long bk = (long long)b * P.S + k;
    RowH<NS> R[3];
    RowY Y[3];
#pragma unroll
    for (int j = 0; j < kFwdRows + 2; ++j) {
      const int o = (trow0 + j) * kDW + lane;
      row_sums<NS, !COLVO_FWD_M_SMEM>(R[j % 3], Y[j % 3], sm.y + o, sm.x[buf][0] + o, sm.x[buf][NS - 1] + o);
      if (j >= 2) {
        const int wr = trow0 + j - 2, py = y0 + wr;
        if (py < P.H && col_in) {
          WinY wy;
#if COLVO_FWD_M_SMEM
          const float4 m0 = sm.m[0][wr * 32 + lane], m1 = sm.m[1][wr * 32 + lane];
          wy.muy[0] = m0.x; wy.muy[1] = m0.y; wy.muy[2] = m0.z;
          wy.sgy[0] = m0.w; wy.sgy[1] = m1.x; wy.sgy[2] = m1.y;
          float best = m1.z;
          int sel = __float_as_int(m1.w);
#else
          float best = id_best[j - 2];
          int sel = id_sel[j - 2];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float sy = Y[0].sy[c] + Y[1].sy[c] + Y[2].sy[c];
            const float syy = Y[0].syy[c] + Y[1].syy[c] + Y[2].syy[c];
            wy.muy[c] = sy * (1.0f / 9.0f);
            wy.sgy[c] = fmaf(-wy.muy[c], wy.muy[c], syy * (1.0f / 9.0f));
          }
#endif
          float Sx[NS][3], Sxx[NS][3], Sxy[NS][3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            wy.yc[c] = Y[(j - 1) % 3].yc[c];
#pragma unroll
            for (int n = 0; n < NS; ++n) {
              Sx[n][c] = R[0].sx[n][c] + R[1].sx[n][c] + R[2].sx[n][c];
              Sxx[n][c] = R[0].sxx[n][c] + R[1].sxx[n][c] + R[2].sxx[n][c];
              Sxy[n][c] = R[0].sxy[n][c] + R[1].sxy[n][c] + R[2].sxy[n][c];
            }
          }
          const RowH<NS>& C = R[(j - 1) % 3];
          winy_derive(wy, c1, c2);
#pragma unroll
          for (int n = 0; n < NS; ++n) {
            const float pe = pe_value3(Sx[n], Sxx[n], Sxy[n], C.xc[n], wy, cal[n], alpha, c1, c2);
            if (pe < best) { best = pe; sel = NS + n; }
          }
          loss_acc += best;
          const int pix = py * P.W + px;
          if (sel_out) sel_out[bk * P.HW + pix] = (uint8_t)sel;
          if (coef_out != nullptr || need_g) {
            float ca[3] = {0.f, 0.f, 0.f}, cb[3] = {0.f, 0.f, 0.f}, cg[3] = {0.f, 0.f, 0.f};
            if (sel >= NS) {
              // the winner's sums, picked without a divergent copy of the arithmetic
              const bool w1 = (NS > 1) && (sel == NS + 1);
              float wSx[3], wSxx[3], wSxy[3], wxc[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                wSx[c] = w1 ? Sx[NS - 1][c] : Sx[0][c];
                wSxx[c] = w1 ? Sxx[NS - 1][c] : Sxx[0][c];
                wSxy[c] = w1 ? Sxy[NS - 1][c] : Sxy[0][c];
                wxc[c] = w1 ? C.xc[NS - 1][c] : C.xc[0][c];
              }
              const CalK wk = make_calk(w1 ? cal[NS - 1].a : cal[0].a, w1 ? cal[NS - 1].b : cal[0].b);
              float dpa = 0.f, dpb = 0.f;
              pe_adjoint(wSx, wSxx, wSxy, wxc, wy, wk, alpha, c1, c2, ca, cb, cg, dpa, dpb);
              if (w1) { ga[NS - 1] += dpa; gb[NS - 1] += dpb; }
              else { ga[0] += dpa; gb[0] += dpb; }
            }
            if (coef_out) {      // zeros where an identity candidate won: the backward stages the tile unconditionally
              const float sidx = (sel == NS + 1) ? 1.f : 0.f;
              float4* co = coef_out + bk * 3 * P.HW + pix;
#pragma unroll
              for (int c = 0; c < 3; ++c) co[(long long)c * P.HW] = make_float4(ca[c], cb[c], cg[c], sidx);
            }
          }
        }
      }
    }
    // dL/da, dL/db terms of this scale: reduced in fp64 right away (large terms of both signs)
    if (need_g) {
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const double sa = warp_sum((double)(ga[n] * (1.0f / 3.0f)));
        const double sb = warp_sum((double)(gb[n] * (1.0f / 3.0f)));
        if (lane == 0) {
          sm.red[wid * NV + 1 + (n * kMaxS + k) * 2 + 0] = sa;
          sm.red[wid * NV + 1 + (n * kMaxS + k) * 2 + 1] = sb;
        }
      }
    }
  }

  // per-tile partials: slot 0 = loss, slots 1.. = dL/da, dL/db per warped frame
  {
    const double s = warp_sum((double)(loss_acc * (1.0f / 3.0f)));      // candidates were compared as 3 * pe
    if (lane == 0) sm.red[wid * NV] = s;
  }
  __syncthreads();
  const int blk = (b * P.ftiles_y + blockIdx.y) * P.ftiles_x + blockIdx.x;
  if (tid < NV && (tid == 0 || need_g)) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kFwdWarps; ++w) s += sm.red[w * NV + tid];
    if (tid == 0) loss_part[blk] = s;
    else g_part[(long long)blk * (NS * kMaxS * 2) + (tid - 1)] = s;
  }
}

template <int NS>
static size_t photo_fwd_smem() { return sizeof(FwdSmem<NS>); }

}  // namespace colvo
