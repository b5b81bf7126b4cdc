// k_photo_fwd: the fused forward tile kernel, sm_100a.
//
// One CTA = 4 warps = one 32 x 12 tile of windows of one triplet.  Every warp walks DOWN a strip of
// 3 window rows with one window column per lane: for each data row it forms the horizontal 3-sums
// of x, x^2 and x*y of the warped frames, keeps the last three rows of sums in registers and adds them
// vertically -- the separable form of the 3x3 SSIM window.
//
// Blackwell specifics: with N = 2 sources the two warped frames of a scale are stored lane-interleaved
// ((x_c^0, x_c^1) pairs: 16 B + 8 B per pixel for the three channels of both sources) and the whole walk --
// row sums, vertical sums, SSIM, L1, and the adjoint of both candidates -- runs on packed fp32x2 registers
// (FFMA2 / FADD2 / FMUL2, colvo_f2.cuh): one issue slot per two sources.  The adjoint re-uses every
// intermediate of the value (colvo_pe.cuh::pe_fused); the winner is picked afterwards by lane.
// The frames of scale k+1 are fetched with 16 + 8 byte cp.async while scale k is evaluated.
//
// Arithmetic contract: oracle/photometric.py (ssim3x3, photometric_error, min_reprojection_automask).
#pragma once
#include "colvo_kernels.cuh"
#include "colvo_pe.cuh"

#ifndef COLVO_MINB_FWD
#define COLVO_MINB_FWD 4
#endif

namespace colvo {

constexpr int kDW = 32 + 2;                 // data columns of a forward tile (1-pixel SSIM halo)
constexpr int kDH = kFwdTileH + 2;          // data rows
constexpr int kDN = kDW * kDH;              // texels per staged frame
constexpr int kStageRounds = (kDN + kFwdThreads - 1) / kFwdThreads;

// Staged frames of one scale.  N = 1: xa = (x0, x1, x2, -).  N = 2: xa = (x0^0, x0^1, x1^0, x1^1), xb = (x2^0, x2^1):
// channel c of both sources is one aligned register pair after LDS.128 / LDS.64 (stride 16 / 8 B: conflict-free).
template <int NS>
struct FwdSmem {
  float4 y[kDN];                            // target tile (y0, y1, y2, -)
  float4 xa[2][kDN];                        // double-buffered over the scales (buffer 0 first holds the raw sources)
  float2 xb[2][NS == 2 ? kDN : 1];
  float4 m[2][kFwdTileH * 32];              // per window: (mu_y[3], var_y[0]), (var_y[1], var_y[2], best identity pe, its index)
  float gslot[NS * kMaxS * 2][kFwdThreads]; // per thread: dL/da, dL/db terms of its windows per (source, scale), summed in fp64 at the end
};

// horizontal 3-sums of one data row at one window column, all sources
template <int NS>
struct RowH {
  Vn<NS> sx[3], sxx[3], sxy[3], xc[3];
};
struct RowY {
  float sy[3], syy[3], yc[3];
};

#define CV_CH(v, c) ((c) == 0 ? (v).x : ((c) == 1 ? (v).y : (v).z))

template <int NS>
__device__ __forceinline__ void load_x3(const float4* __restrict__ xa, const float2* __restrict__ xb, int o, Vn<NS> (&v)[3]);
template <>
__device__ __forceinline__ void load_x3<1>(const float4* __restrict__ xa, const float2* __restrict__, int o, Vn<1> (&v)[3]) {
  const float4 a = xa[o];
  v[0] = Vn<1>(a.x); v[1] = Vn<1>(a.y); v[2] = Vn<1>(a.z);
}
template <>
__device__ __forceinline__ void load_x3<2>(const float4* __restrict__ xa, const float2* __restrict__ xb, int o, Vn<2> (&v)[3]) {
  const float4 a = xa[o];
  const float2 b = xb[o];
  v[0] = Vn<2>(a.x, a.y); v[1] = Vn<2>(a.z, a.w); v[2] = Vn<2>(b.x, b.y);
}

template <int NS, bool WITH_Y>
__device__ __forceinline__ void row_sums(RowH<NS>& R, RowY& Y, const float4* __restrict__ yr, const float4* __restrict__ xa,
                                         const float2* __restrict__ xb, int o) {
  const float4 ya = yr[o], yb = yr[o + 1], yc = yr[o + 2];
  Vn<NS> a[3], b[3], d[3];
  load_x3<NS>(xa, xb, o, a);
  load_x3<NS>(xa, xb, o + 1, b);
  load_x3<NS>(xa, xb, o + 2, d);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Y.yc[c] = CV_CH(yb, c);
    if (WITH_Y) {
      Y.sy[c] = CV_CH(ya, c) + CV_CH(yb, c) + CV_CH(yc, c);
      Y.syy[c] = fmaf(CV_CH(yc, c), CV_CH(yc, c), fmaf(CV_CH(yb, c), CV_CH(yb, c), CV_CH(ya, c) * CV_CH(ya, c)));
    }
    R.sx[c] = a[c] + b[c] + d[c];
    R.sxx[c] = fma2(d[c], d[c], fma2(b[c], b[c], a[c] * a[c]));
    R.sxy[c] = fma2(d[c], bc<NS>(CV_CH(yc, c)), fma2(b[c], bc<NS>(CV_CH(yb, c)), a[c] * bc<NS>(CV_CH(ya, c))));
    R.xc[c] = b[c];
  }
}

// vertical 3-sums of the last three rows
template <int NS>
__device__ __forceinline__ void window_sums(const RowH<NS> (&R)[3], Vn<NS> (&Sx)[3], Vn<NS> (&Sxx)[3], Vn<NS> (&Sxy)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Sx[c] = R[0].sx[c] + R[1].sx[c] + R[2].sx[c];
    Sxx[c] = R[0].sxx[c] + R[1].sxx[c] + R[2].sxx[c];
    Sxy[c] = R[0].sxy[c] + R[1].sxy[c] + R[2].sxy[c];
  }
}

// ADJ: also produce what the backward needs (coefficient texels, dL/da, dL/db partials, sel is always written)
template <int NS, bool PK, bool ADJ>
__global__ void __launch_bounds__(kFwdThreads, COLVO_MINB_FWD)
    k_photo_fwd(KP P, const float* __restrict__ ab, uint8_t* __restrict__ sel_out, double* __restrict__ loss_part,
                double* __restrict__ g_part, int need_g, float4* __restrict__ coef_out, const float4* __restrict__ iw) {
  constexpr int NV = 1 + NS * kMaxS * 2;
  typedef Vn<NS> V;
  extern __shared__ __align__(16) unsigned char fwd_smem_raw[];
  FwdSmem<NS>& sm = *reinterpret_cast<FwdSmem<NS>*>(fwd_smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // Triplets are walked in DESCENDING order: k_warp_stats wrote the warped frames in ascending order, so the last
  // triplets are the ones still resident in L2 when this grid starts.
  const int b = (int)gridDim.z - 1 - (int)blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * kFwdTileH;
  const int px = x0 + lane;
  const bool col_in = px < P.W;

  // the staged positions of this thread (reflect-padded coordinates of the tile + halo)
  int goff[kStageRounds];
#pragma unroll
  for (int j = 0; j < kStageRounds; ++j) {
    const int idx = tid + j * kFwdThreads;
    const int r = idx / kDW, c = idx - r * kDW;
    goff[j] = (idx < kDN) ? reflect_clamp(y0 - 1 + r, P.H) * P.W + reflect_clamp(x0 - 1 + c, P.W) : -1;
    CV_CHECK(goff[j] < P.HW);
  }
  auto stage_scale = [&](int k, int buf) {      // warped frames of scale k -> sm.xa / sm.xb [buf]
    // frame base hidden from the optimiser + 32-bit offsets: one IMAD.WIDE per copy (see Img<false>::load_taps);
    // the shared-window address is formed once, the rounds are immediate offsets
    if constexpr (NS == 1) {
      const float4* src = iw + (long long)(b * P.S + k) * P.HW;
      asm volatile("" : "+l"(src));
      const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm.xa[buf][tid]);
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j)
        if (goff[j] >= 0) cp_async16_s(sa + j * kFwdThreads * (unsigned)sizeof(float4), src + (unsigned)goff[j]);
    } else {
      const float4* srca = iw + (long long)(b * P.S + k) * P.HW;
      const float2* srcb = reinterpret_cast<const float2*>(iw + (long long)P.B * P.S * P.HW) + (long long)(b * P.S + k) * P.HW;
      asm volatile("" : "+l"(srca), "+l"(srcb));
      const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm.xa[buf][tid]);
      const unsigned sb = (unsigned)__cvta_generic_to_shared(&sm.xb[buf][tid]);
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j)
        if (goff[j] >= 0) {
          cp_async16_s(sa + j * kFwdThreads * (unsigned)sizeof(float4), srca + (unsigned)goff[j]);
          cp_async8_s(sb + j * kFwdThreads * (unsigned)sizeof(float2), srcb + (unsigned)goff[j]);
        }
    }
    cp_async_commit();
  };

  // target tile and the raw sources (identity candidates), either storage format.  Planar fp32 goes through
  // 4-byte cp.async straight into the texel components, so all 9 loads per position are in flight at once;
  // packed bf16 is widened on the way in (all loads first, then the stores).
  if constexpr (!PK) {
    const unsigned hw = P.HW;
    {
      const float* fb = static_cast<const float*>(P.tgt) + (long long)b * P.tgt_bf * P.frame_el;
      asm volatile("" : "+l"(fb));
      const unsigned sa = (unsigned)__cvta_generic_to_shared(sm.y + tid);
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j)
        if (goff[j] >= 0) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
            cp_async4_s(sa + (j * kFwdThreads * 4 + ch) * (unsigned)sizeof(float), fb + ((unsigned)goff[j] + ch * hw));
        }
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const float* fb = static_cast<const float*>(P.srcs) + (long long)(b * P.src_bf + n * P.src_nf) * P.frame_el;
      asm volatile("" : "+l"(fb));
      const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm.xa[0][tid]);
      const unsigned sb = (unsigned)__cvta_generic_to_shared(&sm.xb[0][NS == 2 ? tid : 0]);
#pragma unroll
      for (int j = 0; j < kStageRounds; ++j)
        if (goff[j] >= 0) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            // N = 1: component ch of the texel.  N = 2: lane n of pair ch (pairs 0, 1 in xa, pair 2 in xb)
            const unsigned dst = (NS == 1) ? sa + (j * kFwdThreads * 4 + ch) * (unsigned)sizeof(float)
                                 : (ch < 2 ? sa + (j * kFwdThreads * 4 + 2 * ch + n) * (unsigned)sizeof(float)
                                           : sb + (j * kFwdThreads * 2 + n) * (unsigned)sizeof(float));
            cp_async4_s(dst, fb + ((unsigned)goff[j] + ch * hw));
          }
        }
    }
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
    for (int j = 0; j < kStageRounds; ++j)
      if (goff[j] >= 0) {
        const int idx = tid + j * kFwdThreads;
        float v[1 + NS][3];
#pragma unroll
        for (int f = 0; f < 1 + NS; ++f) {
          v[f][0] = __uint_as_float(raw[f][j].x << 16);
          v[f][1] = __uint_as_float(raw[f][j].x & 0xffff0000u);
          v[f][2] = __uint_as_float(raw[f][j].y << 16);
        }
        sm.y[idx] = make_float4(v[0][0], v[0][1], v[0][2], 0.f);
        if constexpr (NS == 1) {
          sm.xa[0][idx] = make_float4(v[1][0], v[1][1], v[1][2], 0.f);
        } else {
          sm.xa[0][idx] = make_float4(v[1][0], v[NS][0], v[1][1], v[NS][1]);
          sm.xb[0][idx] = make_float2(v[1][2], v[NS][2]);
        }
      }
  }
  cp_async_commit();
  if (ADJ) {                    // slots of unused scales stay 0
#pragma unroll
    for (int i = 0; i < NS * kMaxS * 2; ++i) sm.gslot[i][tid] = 0.f;
  }
  pdl_trigger();
  cp_async_wait_all();          // the raw frames have landed
  __syncthreads();

  const float alpha = P.alpha, c1 = P.c1, c2 = P.c2;
  const int trow0 = wid * kFwdRows;           // first window row of this warp = its first data row in the tile

  // ---- identity candidates (raw sources, a = 1, b = 0; oracle A10) and the target window moments ----
  {
    float one[NS], zero[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) { one[n] = 1.0f; zero[n] = 0.0f; }
    const CalV<NS> kid = make_calv<NS>(one, zero, alpha);
    RowH<NS> R[3];
    RowY Y[3];
#pragma unroll
    for (int j = 0; j < kFwdRows + 2; ++j) {
      const int o = (trow0 + j) * kDW + lane;
      row_sums<NS, true>(R[j % 3], Y[j % 3], sm.y, sm.xa[0], sm.xb[0], o);
      if (j >= 2) {
        WinY wy;
        V Sx[3], Sxx[3], Sxy[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float sy = Y[0].sy[c] + Y[1].sy[c] + Y[2].sy[c];
          const float syy = Y[0].syy[c] + Y[1].syy[c] + Y[2].syy[c];
          wy.muy[c] = sy * (1.0f / 9.0f);
          wy.sgy[c] = fmaf(-wy.muy[c], wy.muy[c], syy * (1.0f / 9.0f));
          wy.yc[c] = Y[(j - 1) % 3].yc[c];
        }
        window_sums<NS>(R, Sx, Sxx, Sxy);
        winy_derive(wy, c1, c2);
        // (candidates are compared as 3 * pe: the mean over channels is a common factor)
        const V pe = pe_value3v<NS>(Sx, Sxx, Sxy, R[(j - 1) % 3].xc, wy, kid, alpha, c1, c2);
        float best = pe.lane(0);
        int sel = 0;
#pragma unroll
        for (int n = 1; n < NS; ++n)
          if (pe.lane(n) < best) { best = pe.lane(n); sel = n; }
        const int w = (trow0 + j - 2) * 32 + lane;
        sm.m[0][w] = make_float4(wy.muy[0], wy.muy[1], wy.muy[2], wy.sgy[0]);
        sm.m[1][w] = make_float4(wy.sgy[1], wy.sgy[2], best, __int_as_float(sel));
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
    {
      const int buf = (k + 1) & 1;
      cp_async_wait_all();
      __syncthreads();              // scale k landed everywhere; every warp is done with the other buffer
      if (k + 1 < P.S) stage_scale(k + 1, buf ^ 1);
      float av[NS], bv[NS];
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const int bnk = (b * P.N + n) * P.S + k;
        av[n] = __ldg(ab + 2 * bnk);
        bv[n] = __ldg(ab + 2 * bnk + 1);
      }
      const CalV<NS> cal = make_calv<NS>(av, bv, alpha);
      V ga = bc<NS>(0.f), gb = bc<NS>(0.f);
      const long long bk = (long long)b * P.S + k;
      RowH<NS> R[3];
      RowY Y[3];
#pragma unroll
      for (int j = 0; j < kFwdRows + 2; ++j) {
        const int o = (trow0 + j) * kDW + lane;
        CV_CHECK(o + 2 < kDN);
        row_sums<NS, false>(R[j % 3], Y[j % 3], sm.y, sm.xa[buf], sm.xb[buf], o);
        if (j >= 2) {
          const int wr = trow0 + j - 2, py = y0 + wr;
          if (py < P.H && col_in) {
            WinY wy;
            const float4 m0 = sm.m[0][wr * 32 + lane], m1 = sm.m[1][wr * 32 + lane];
            wy.muy[0] = m0.x; wy.muy[1] = m0.y; wy.muy[2] = m0.z;
            wy.sgy[0] = m0.w; wy.sgy[1] = m1.x; wy.sgy[2] = m1.y;
            float best = m1.z;
            int sel = __float_as_int(m1.w);
#pragma unroll
            for (int c = 0; c < 3; ++c) wy.yc[c] = Y[(j - 1) % 3].yc[c];
            V Sx[3], Sxx[3], Sxy[3];
            window_sums<NS>(R, Sx, Sxx, Sxy);
            const RowH<NS>& C = R[(j - 1) % 3];
            winy_derive(wy, c1, c2);
            V pe, ca[3], cb[3], cg[3], dpa, dpb;
            if constexpr (ADJ) pe = pe_fused<NS>(Sx, Sxx, Sxy, C.xc, wy, cal, alpha, c1, c2, ca, cb, cg, dpa, dpb);
            else pe = pe_value3v<NS>(Sx, Sxx, Sxy, C.xc, wy, cal, alpha, c1, c2);
#pragma unroll
            for (int n = 0; n < NS; ++n)
              if (pe.lane(n) < best) { best = pe.lane(n); sel = NS + n; }
            loss_acc += best;
            const int pix = py * P.W + px;
            if (sel_out) sel_out[bk * P.HW + pix] = (uint8_t)sel;
            if constexpr (ADJ) {
              // which source won here (none where an identity candidate did): the lane weights of the winner
              V wm;
#pragma unroll
              for (int n = 0; n < NS; ++n) wm.set(n, sel == NS + n ? 1.f : 0.f);
              ga = fma2(wm, dpa, ga);
              gb = fma2(wm, dpb, gb);
              if (coef_out) {
                // the winner's coefficients (lane pick); the winner flags ride in .w of channels 0 / 1, so the loser's
                // values need no zeroing: the backward weights every window by its flags
                const bool w1 = (NS > 1) && (sel == NS + 1);
                float4* co = coef_out + bk * 3 * P.HW + pix;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  const float fa = w1 ? ca[c].lane(NS - 1) : ca[c].lane(0);
                  const float fb = w1 ? cb[c].lane(NS - 1) : cb[c].lane(0);
                  const float fg = w1 ? cg[c].lane(NS - 1) : cg[c].lane(0);
                  const float flag = (c == 0) ? wm.lane(0) : ((c == 1 && NS > 1) ? wm.lane(NS - 1) : 0.f);
                  st_stream(co + (long long)c * P.HW, make_float4(fa, fb, fg, flag));
                }
              }
            }
          }
        }
      }
      if constexpr (ADJ) {
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          sm.gslot[(n * kMaxS + k) * 2 + 0][tid] = ga.lane(n) * (1.0f / 3.0f);
          sm.gslot[(n * kMaxS + k) * 2 + 1][tid] = gb.lane(n) * (1.0f / 3.0f);
        }
      }
    }
  }

  // per-tile partials: slot 0 = loss, slots 1.. = dL/da, dL/db per warped frame.  Each thread parks its fp32 sums
  // (at most 3 windows each) in shared memory; (slot, warp) pairs are then summed in fp64 in a fixed order.
  __syncthreads();                               // all warps are done with the staged frames: reuse them
  float* lslot = reinterpret_cast<float*>(sm.xa);        // [kFwdThreads] fp32
  double* part = reinterpret_cast<double*>(sm.y);        // [kFwdWarps * (NV - 1)]
  static_assert(sizeof(double) * kFwdWarps * NV <= sizeof(sm.y), "partials storage");
  lslot[tid] = loss_acc * (1.0f / 3.0f);                 // candidates were compared as 3 * pe
  __syncthreads();
  const int blk = (b * P.ftiles_y + blockIdx.y) * P.ftiles_x + blockIdx.x;
  block_sum_slots<1, kFwdThreads>(lslot, part, [&](int, double v) { loss_part[blk] = v; });
  if (ADJ && need_g) {
    __syncthreads();
    block_sum_slots<NV - 1, kFwdThreads>(&sm.gslot[0][0], part, [&](int slot, double v) {
      g_part[(long long)blk * (NS * kMaxS * 2) + slot] = v;
    });
  }
}

template <int NS>
static size_t photo_fwd_smem() { return sizeof(FwdSmem<NS>); }

}  // namespace colvo
