// Forward kernels of the ColVO photometric-loss path,
// hand-written for sm_100a.  oracle/photometric.py is the arithmetic contract.
//
//   k_smooth        edge-aware smoothness of all scales in one pass over the target: pyramid in
//                   shared memory, loss partials, adjoint field when saving              (row 9)
//   k_warp_stats    warp every (b,n,k) frame once, fp64 LCC sums, valid mask          (rows 0-5)
//   k_lcc_solve     (a, b) per warped frame                                           (row 5)
//   k_photo_fwd     (colvo_photo_fwd.cuh) per 32x12 tile, strip walk: identity + re-projection candidates,
//                   SSIM+L1, min-reprojection / auto-mask, loss partials, dL/da, dL/db  (rows 6-8)
//   k_finalize_fwd  deterministic final sums -> loss, G_a, G_b, sum s*d               (row 10)
#include <type_traits>

#include "colvo_kernels.cuh"
#include "colvo_photo_fwd.cuh"

// occupancy knobs (CTAs per SM the register allocator must allow) -- tuned on B200, see DESIGN.md
#ifndef COLVO_MINB_STATS
#define COLVO_MINB_STATS 4
#endif
#ifndef COLVO_STATS_UNROLL  // pixels of the statistics loop in flight per thread
#define COLVO_STATS_UNROLL 4
#endif

namespace colvo {

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// Edge-aware smoothness (row 9) of ALL scales in one pass over the target.  One CTA = one 64x32
// block of full-resolution pixels of one triplet: it loads that block (+ halo) once, builds the
// box-averaged pyramid level by level in shared memory (oracle.target_pyramid: the 2^k x 2^k mean
// is the mean of four means of the level below), and evaluates at every scale its own texels: each
// visits its four edges -- the right/down ones give the loss, all four give s_p = dL/dd*_p for
// grad_loss = 1, which the backward only has to rescale.  Level k is kept with a halo of
// 2^(S-1-k) texels so that level k+1 can be built including its own halo.
// The mean-disparity normalisation d* = d / (mean d + eps) is a positive per-(b,k) factor, so the
// sums are formed on d = 1/D and the factor is applied by k_finalize_fwd (|d*_i - d*_j| =
// |d_i - d_j| / (mean + eps); signs are unchanged).  Depends on the inputs only.
struct SmoothOut {
  double* part;            // [B][sm_blocks][S][kSmVals]
  float* sf[kMaxS];        // adjoint fields [B,h_k,w_k], or null when nothing is saved
};
__host__ __device__ constexpr int sm_img_floats() {   // image tiles of all levels at S = kMaxS (halo 8, 4, 2, 1), 3 planes
  int n = 0;
  for (int k = 0; k < kMaxS; ++k) n += 3 * ((kSmBW >> k) + 2 * (1 << (kMaxS - 1 - k))) * ((kSmBH >> k) + 2 * (1 << (kMaxS - 1 - k)));
  return n;
}
__host__ __device__ constexpr int sm_dep_floats() {   // inverse-depth tiles of all levels, halo 1
  int n = 0;
  for (int k = 0; k < kMaxS; ++k) n += ((kSmBW >> k) + 2) * ((kSmBH >> k) + 2);
  return n;
}

template <bool PK>
__global__ void __launch_bounds__(kThreads)
    k_smooth(KP P, SmoothOut O, int b0) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  float* simg = reinterpret_cast<float*>(sm_raw);
  float* sdep = simg + sm_img_floats();
  double* red = reinterpret_cast<double*>(sdep + ((sm_dep_floats() + 1) & ~1));
  pdl_trigger();
  const int tid = threadIdx.x, b = b0 + blockIdx.z;
  const int X0 = blockIdx.x * kSmBW, Y0 = blockIdx.y * kSmBH;
  const Img<PK> im = img_at<PK>(P, P.tgt, b * P.tgt_bf);

  int ioff[kMaxS], doff[kMaxS];          // smem offsets of the level tiles
  {
    int io = 0, dof = 0;
#pragma unroll
    for (int k = 0; k < kMaxS; ++k) {
      ioff[k] = io;
      doff[k] = dof;
      const int hal = (k < P.S) ? (1 << (P.S - 1 - k)) : 0;
      io += 3 * ((kSmBW >> k) + 2 * hal) * ((kSmBH >> k) + 2 * hal);
      dof += ((kSmBW >> k) + 2) * ((kSmBH >> k) + 2);
    }
  }
  // (all tile loops below are 2-D -- warp = 32 consecutive columns, 8 rows per pass -- so no integer divisions)
  const int tx = tid & 31, ty = tid >> 5;
  // ---- level 0: the full-resolution block + halo, and the inverse-depth tiles of every scale ----
  {
    const int hal = 1 << (P.S - 1), dw = kSmBW + 2 * hal, dh = kSmBH + 2 * hal, dn = dw * dh;
    for (int r = ty; r < dh; r += kThreads / 32) {
      const int gy = Y0 - hal + r;
      const bool row_in = gy >= 0 && gy < P.H;
      for (int c = tx; c < dw; c += 32) {
        const int gx = X0 - hal + c;
        float v[3] = {0.f, 0.f, 0.f};
        if (row_in && gx >= 0 && gx < P.W) im.load3(gy * P.W + gx, v);
        const int idx = r * dw + c;
        simg[idx] = v[0];
        simg[dn + idx] = v[1];
        simg[2 * dn + idx] = v[2];
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxS; ++k) {
      if (k < P.S) {
        const int hk = P.h[k], wk = P.w[k], dwk = (kSmBW >> k) + 2, dhk = (kSmBH >> k) + 2;
        const float* D = P.depth[k] + (long long)b * P.depth_bs[k];
        for (int r = ty; r < dhk; r += kThreads / 32) {
          const int gy = (Y0 >> k) - 1 + r;
          const bool row_in = gy >= 0 && gy < hk;
          for (int c = tx; c < dwk; c += 32) {
            const int gx = (X0 >> k) - 1 + c;
            sdep[doff[k] + r * dwk + c] = (row_in && gx >= 0 && gx < wk) ? f_rcp(__ldg(D + gy * wk + gx)) : 0.f;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- levels 1 .. S-1: mean of the four texels below ----
#pragma unroll
  for (int k = 1; k < kMaxS; ++k) {
    if (k < P.S) {
      const int hal = 1 << (P.S - 1 - k), dw = (kSmBW >> k) + 2 * hal, dh = (kSmBH >> k) + 2 * hal, dn = dw * dh;
      const int pdw = (kSmBW >> (k - 1)) + 4 * hal, pdn = pdw * ((kSmBH >> (k - 1)) + 4 * hal);
      const float* src = simg + ioff[k - 1];
      float* dst = simg + ioff[k];
      for (int r = ty; r < dh; r += kThreads / 32) {
        for (int c = tx; c < dw; c += 32) {
          const int o = (2 * r) * pdw + 2 * c, idx = r * dw + c;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float* q = src + ch * pdn + o;
            dst[ch * dn + idx] = ((q[0] + q[1]) + (q[pdw] + q[pdw + 1])) * 0.25f;
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- smoothness of the own texels of every scale ----
  const int blk = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
#pragma unroll 1
  for (int k = 0; k < P.S; ++k) {
    const int hk = P.h[k], wk = P.w[k];
    const int hal = 1 << (P.S - 1 - k), tw = kSmBW >> k, th = kSmBH >> k;
    const int dw = tw + 2 * hal, dn = dw * (th + 2 * hal), ddw = tw + 2;
    const float* I = simg + ioff[k];
    const float* Dr = sdep + doff[k];
    float* sf = O.sf[k] ? O.sf[k] + (long long)b * hk * wk : nullptr;
    const float cx = P.sm_cx[k], cy = P.sm_cy[k];
    float acc[kSmVals] = {0.f, 0.f, 0.f, 0.f};      // at most 8 texels per thread and scale: fp32, then fp64 across threads
    constexpr int kSmBWLog = (kSmBW == 128) ? 7 : (kSmBW == 64 ? 6 : (kSmBW == 32 ? 5 : (kSmBW == 16 ? 4 : 3)));
    static_assert((1 << kSmBWLog) == kSmBW, "kSmBW must be a power of two in [8, 128]");
    const int tw_log = kSmBWLog - k;                // tw = kSmBW >> k is a power of two
    for (int i = tid; i < tw * th; i += kThreads) {
      const int ly = i >> tw_log, lx = i & (tw - 1);
      const int gy = (Y0 >> k) + ly, gx = (X0 >> k) + lx;
      if (gy < hk && gx < wk) {
        const int o = (ly + hal) * dw + lx + hal, od = (ly + 1) * ddw + lx + 1;
        const float i0 = I[o], i1 = I[dn + o], i2 = I[2 * dn + o], dr = Dr[od];
        auto edge = [&](int oj, int odj) -> float2 {   // (d_i - d_j, exp(-mean_c |I_i - I_j|))
          const float e = (fabsf(i0 - I[oj]) + fabsf(i1 - I[dn + oj]) + fabsf(i2 - I[2 * dn + oj])) * (1.0f / 3.0f);
          return make_float2(dr - Dr[odj], __expf(-e));   // e in [0, 1]: MUFU.EX2 path, rel. error ~1e-6
        };
        float s = 0.f;
        if (gx + 1 < wk) {
          const float2 e = edge(o + 1, od + 1);
          acc[0] += fabsf(e.x) * e.y;
          s += sgn(e.x) * e.y * cx;
        }
        if (gy + 1 < hk) {
          const float2 e = edge(o + dw, od + ddw);
          acc[1] += fabsf(e.x) * e.y;
          s += sgn(e.x) * e.y * cy;
        }
        if (sf) {
          if (gx > 0) { const float2 e = edge(o - 1, od - 1); s += sgn(e.x) * e.y * cx; }
          if (gy > 0) { const float2 e = edge(o - dw, od - ddw); s += sgn(e.x) * e.y * cy; }
          sf[gy * wk + gx] = s;
          acc[2] = fmaf(s, dr, acc[2]);
        }
        acc[3] += dr;
      }
    }
    double accd[kSmVals];
#pragma unroll
    for (int j = 0; j < kSmVals; ++j) accd[j] = (double)acc[j];
    block_reduce_store<kSmVals, double>(accd, red, O.part + ((long long)blk * P.S + k) * kSmVals);
    __syncthreads();          // red is reused by the next scale
  }
  // Launched programmatically behind k_warp_stats / k_lcc_solve, whose results it does not need: it fills the tail of
  // that launch while k_photo_fwd (launched behind this one) runs its own input-only prologue.  k_photo_fwd waits
  // on THIS grid, so this grid must not complete before its predecessors have.
  pdl_wait();
}

// ------------------------------------------------------------------------------------------
// One CTA = one (b, k) and a chunk of pixels; all sources are warped by the same thread so the ray, the
// up-sampled depth and the target pixel are loaded once, and with N = 2 the two sources ride in the two lanes
// of packed fp32x2 registers (reprojection chain, bilinear blend, LCC products: FMUL2 / FADD2 / FFMA2 --
// the pinned chain stays single-rounded per lane, so `valid` is bit-exact as before).
// The raw warped frames are written out for the tile kernel, which needs them with a halo (re-warping there
// costs more issue slots than the round trip costs bandwidth on this ALU-bound path):
//   N = 1:  iw[(b,k)][pix] = (x0, x1, x2, valid)                                   16 B / pixel
//   N = 2:  iwA[(b,k)][pix] = (x0^0, x0^1, x1^0, x1^1), iwB[(b,k)][pix] = (x2^0, x2^1)   24 B / pixel for both sources
// (iwB starts B*S*HW float4 behind iw), i.e. channel c of both sources is one aligned register pair downstream.
//
// KIN (training loss, S > 1): one CTA = one chunk of pixels of one triplet for ALL scales, scale after scale -- the four
// scales gather from the same source lines, so scales 1.. find them in L1 (with one CTA per (b, k) they sat in the L1s of
// different SMs); 128-thread CTAs of kStatPPTK pixels per thread and scale keep the CTA count (and the pixels in flight per
// thread) of the one-scale form.  !KIN (consistency sweep, S = 1): blockIdx.y = (b, k), 256 threads, kStatPPT pixels each.
// MODE 0: one scale per CTA, kStatPPT pixels per thread (sweep); 1: all scales per CTA (KIN); 2: one scale per CTA,
// kStatPPTTrain pixels per thread (training loss)
template <int MODE> struct StatCfg {
  static constexpr bool KIN = MODE == 1;
  static constexpr int NT = KIN ? kStatThreadsK : kThreads, PPT = KIN ? kStatPPTK : (MODE == 2 ? kStatPPTTrain : kStatPPT);
};
template <int NS, bool GEO, bool PK, int MODE>
__global__ void __launch_bounds__(StatCfg<MODE>::NT, COLVO_MINB_STATS * kThreads / StatCfg<MODE>::NT)
    k_warp_stats(KP P, double* __restrict__ part, uint8_t* __restrict__ valid_out, float4* __restrict__ iw_out,
                 float4* __restrict__ geo_out, float* __restrict__ occ_out) {
  constexpr int NA = GEO ? kStatVals : 5;      // accumulators per source (the 6th only with the geometric term)
  constexpr int NV = NA * NS;
  constexpr bool KIN = StatCfg<MODE>::KIN;
  constexpr int NT = StatCfg<MODE>::NT, PPT = StatCfg<MODE>::PPT;
  typedef Vn<NS> V;
  __shared__ double sm[(NT / 32) * NV];
  pdl_trigger();                               // k_lcc_solve / k_photo_fwd may start their prologues in this kernel's tail
  const int b = KIN ? (int)blockIdx.y : (int)blockIdx.y / P.S;
  const int k_lo = KIN ? 0 : (int)blockIdx.y % P.S, k_hi = KIN ? P.S : k_lo + 1;
#pragma unroll 1
  for (int k = k_lo; k < k_hi; ++k) {
  const Cam cam = load_cam(P, b);               // (re-loaded per scale: cheaper than 28 registers held across the loop)
  const PoseV<NS> pose = load_pose_v<NS>(P, b);
  const float* Dk = P.depth[k] + (long long)b * P.depth_bs[k];
  Img<PK> tg = img_at<PK>(P, P.tgt, b * P.tgt_bf);
  // Loop-invariant bases, materialised once and hidden from the optimiser (which otherwise re-derives the 64-bit
  // frame offsets per use under this kernel's 64-register cap); everything inside the loop is base + 32-bit offset:
  // source n's frame sits src_noff elements behind source 0's, output frame (b, n, k) sits n * S * HW behind (b, 0, k).
  Img<PK> src0 = img_at<PK>(P, P.srcs, b * P.src_bf);
  const int src_noff = (int)(P.src_nf * P.frame_el);
  const unsigned out_nstride = (unsigned)P.S * (unsigned)P.HW, hw = P.HW;
  const long long out_bk = (long long)(b * P.N * P.S + k) * P.HW;
  uint8_t* valid_b = valid_out ? valid_out + out_bk : nullptr;
  float* occ_b = (GEO && occ_out) ? occ_out + out_bk : nullptr;     // soft occlusion mask 1 - diff (0 where invalid), [B,N,S,H,W]
  // saved projection: N = 1 one texel array; N = 2 the planes A = (u^0, u^1, v^0, v^1) and B = (iz^0, iz^1, D^, valid bits)
  float4* geo_b = geo_out ? geo_out + (long long)(b * P.S + k) * P.HW : nullptr;
  float4* geo_b2 = (geo_out && NS == 2) ? geo_out + (long long)((P.B + b) * P.S + k) * P.HW : nullptr;
  // warped frames: N = 1 one texel array; N = 2 the pair planes A (float4) and B (float2) indexed by (b, k)
  float4* iw_a = iw_out ? iw_out + (NS == 1 ? out_bk : (long long)(b * P.S + k) * P.HW) : nullptr;
  float2* iw_b2 = (iw_out && NS == 2) ? reinterpret_cast<float2*>(iw_out + (long long)P.B * P.S * P.HW) + (long long)(b * P.S + k) * P.HW
                                      : nullptr;
  asm volatile("" : "+l"(tg.p), "+l"(src0.p), "+l"(valid_b), "+l"(iw_a), "+l"(iw_b2), "+l"(geo_b), "+l"(geo_b2));
  int pix = blockIdx.x * (NT * PPT) + threadIdx.x;
  int py = pix / P.W, px = pix - py * P.W;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  constexpr int kUnroll = COLVO_STATS_UNROLL < PPT ? COLVO_STATS_UNROLL : PPT;
#pragma unroll kUnroll
  for (int i = 0; i < PPT; ++i) {
    if (pix < P.HW) {
      CV_CHECK(py < P.H && px < P.W && py * P.W + px == pix);
      const float rx = ray_x(px, cam), ry = ray_y(py, cam);
      const float D = depth_at(P, Dk, k, px, py);
      float y0, y1, y2;
      if constexpr (PK) {
        float yv[3];
        tg.load3(pix, yv);
        y0 = yv[0]; y1 = yv[1]; y2 = yv[2];
      } else {
        const unsigned upix = pix;
        y0 = __ldg(tg.p + upix); y1 = __ldg(tg.p + (upix + hw)); y2 = __ldg(tg.p + (upix + 2 * hw));
      }
      const GeoV<NS> g = reproject_v<NS>(rx, ry, D, cam, pose, P.W, P.H, P.eps_proj, P.z_min);
      Taps t[NS];
      Texels tx[NS];
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        t[n] = make_taps(g.u.lane(n), g.v.lane(n), P.W, P.H);
        CV_CHECK_TAPS(t[n], P.W, P.H);
        const int foff = n * src_noff;
        const int r0 = foff + t[n].y0 * P.W, r1 = foff + t[n].y1 * P.W;
        src0.load_taps(r0 + t[n].x0, r0 + t[n].x1, r1 + t[n].x0, r1 + t[n].x1, tx[n]);
      }
      // bilinear blend of all sources, lane by lane the arithmetic of colvo_math.cuh::bilerp
      V wx, wy, x[3];
#pragma unroll
      for (int n = 0; n < NS; ++n) { wx.set(n, t[n].wx); wy.set(n, t[n].wy); }
      const V omx = bc<NS>(1.0f) - wx, omy = bc<NS>(1.0f) - wy;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        V i00, i01, i10, i11;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          i00.set(n, tx[n].i00[c]); i01.set(n, tx[n].i01[c]); i10.set(n, tx[n].i10[c]); i11.set(n, tx[n].i11[c]);
        }
        const V top = fma2(wx, i01, omx * i00), bot = fma2(wx, i11, omx * i10);
        x[c] = fma2(wy, bot, omy * top);
      }
      // raw warped frames, re-used by k_photo_fwd instead of warping again (+halo)
      if (iw_a) {
        if constexpr (NS == 1) {
          st_stream(iw_a + pix, make_float4(x[0].lane(0), x[1].lane(0), x[2].lane(0), g.valid[0] ? 1.f : 0.f));
        } else {
          st_stream(iw_a + pix, make_float4(x[0].lane(0), x[0].lane(1), x[1].lane(0), x[1].lane(1)));
          st_stream(iw_b2 + pix, make_float2(x[2].lane(0), x[2].lane(1)));
        }
      }
      // the projection itself, for the backward
      if (geo_b) {
        if constexpr (NS == 1) {              // valid rides in the mantissa LSB of the depth
          st_stream(geo_b + pix, make_float4(g.u.lane(0), g.v.lane(0), g.iz.lane(0),
                                             __uint_as_float((__float_as_uint(D) & ~1u) | (g.valid[0] ? 1u : 0u))));
        } else {
          st_stream(geo_b + pix, make_float4(g.u.lane(0), g.u.lane(1), g.v.lane(0), g.v.lane(1)));
          st_stream(geo_b2 + pix, make_float4(g.iz.lane(0), g.iz.lane(1), D,
                                              __uint_as_float((g.valid[0] ? 1u : 0u) | (g.valid[NS - 1] ? 2u : 0u))));
        }
      }
      // LCC products of all sources
      const V s1 = x[0] + x[1] + x[2];
      const V s2 = fma2(x[2], x[2], fma2(x[1], x[1], x[0] * x[0]));
      const V s3 = fma2(x[2], bc<NS>(y2), fma2(x[1], bc<NS>(y1), x[0] * bc<NS>(y0)));
      const float sy = y0 + y1 + y2;
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const unsigned opix = (unsigned)pix + n * out_nstride;
        if (valid_b) valid_b[opix] = g.valid[n] ? 1 : 0;
        float occ = 0.f;
        if (g.valid[n]) {
          acc[NA * n + 0] += 3.0;
          acc[NA * n + 1] += (double)s1.lane(n);
          acc[NA * n + 2] += (double)sy;
          acc[NA * n + 3] += (double)s2.lane(n);
          acc[NA * n + 4] += (double)s3.lane(n);
          if (GEO) {               // geometric consistency (f-2): per-pixel, so it lives in this pass
            float d4[4], dZ, dS;
            const float ds = sample_plane(P.src_depth + (long long)(b * P.N + n) * P.HW, t[n], P.W, d4);
            const float diff = geo_diff(g.Zp.lane(n), ds, dZ, dS);
            acc[NA * n + (NA - 1)] += (double)diff;
            occ = 1.0f - diff;
          }
        }
        if (GEO && occ_b) occ_b[opix] = occ;
      }
    }
    pix += NT;
    px += NT;
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
    for (int w = 0; w < NT / 32; ++w) s += sm[w * NV + threadIdx.x];
    const int n = threadIdx.x / NA, j = threadIdx.x - NA * n;
    const int bnk = (b * P.N + n) * P.S + k;
    part[((long long)bnk * gridDim.x + blockIdx.x) * kStatVals + j] = s;
  }
  if (KIN) __syncthreads();      // sm is reused by the next scale
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    k_lcc_solve(KP P, const double* __restrict__ part, int chunks, float* __restrict__ ab, double* __restrict__ saved) {
  const int bnk = blockIdx.x, lane = threadIdx.x;
  pdl_trigger();
  pdl_wait();                          // the statistics partials of k_warp_stats
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
// block 0: the scalar loss.  blocks [1, 1+BNS): G_a, G_b of warped frame bnk (dL/da, dL/db).
// blocks [1+BNS, 1+BNS+B*S): mean inverse depth and sum_p s_p d_p of (b, k) for the smoothness adjoint.
__global__ void __launch_bounds__(kThreads)
    k_finalize_fwd(KP P, const double* __restrict__ loss_part, const double* __restrict__ g_part,
                   const double* __restrict__ smooth_part, const double* __restrict__ stat_part, int stat_chunks,
                   float* __restrict__ loss, double* __restrict__ saved_frame, double* __restrict__ saved_scale,
                   int need_g) {
  __shared__ double sm[(kThreads / 32) * 2];
  __shared__ double wk_s[kMaxS][2];
  pdl_wait();                                   // launched programmatically behind k_photo_fwd
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
    {
      // four independent partial sums per thread: the loads of four strides are in flight together (this block is
      // latency-bound: 15 k partials at 1080x1350); fixed order, so the result is reproducible
      const int n = P.B * tiles;
      double a4[4] = {0.0, 0.0, 0.0, 0.0};
      int i = threadIdx.x;
      for (; i + 3 * kThreads < n; i += 4 * kThreads) {
#pragma unroll
        for (int u = 0; u < 4; ++u) a4[u] += loss_part[i + u * kThreads];
      }
      for (; i < n; i += kThreads) a4[0] += loss_part[i];
      acc[0] = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    }
    if (P.src_depth) {                                // L_geo,k = sum of diffs / (B N HW), weight geo_weight
      const double wg = (double)P.geo_weight / ((double)P.B * (double)P.N * (double)P.HW);
      for (int i = threadIdx.x; i < BNS * stat_chunks; i += kThreads) acc[1] += stat_part[(long long)i * kStatVals + 5] * wg;
    }
    // smoothness: L lanes per (b, k) sum the partials of that image's k_smooth CTAs (kThreads / L pairs per pass of the
    // block, fixed order), then 1 / (mean d + eps) is applied (see k_smooth).  L = the largest power of two <= 32 that
    // still covers all B*S pairs in one pass when possible: 4 at config 2 (48 pairs, 80 CTAs each), 16 at config 4
    // (16 pairs, 1 496 CTAs each) -- the loop is latency-bound, so more lanes per pair = fewer dependent loads each.
    {
      int L = 4;
      while (L < 32 && (kThreads / (2 * L)) >= P.B * P.S) L *= 2;
      const int grp = threadIdx.x / L, sub = threadIdx.x % L, per_pass = kThreads / L;
      for (int bk0 = 0; bk0 < P.B * P.S; bk0 += per_pass) {
        const int bk = bk0 + grp;
        const bool on = bk < P.B * P.S;
        const int b = on ? bk / P.S : 0, k = on ? bk - b * P.S : 0;
        const double* sp = smooth_part + ((long long)b * P.sm_blocks * P.S + k) * kSmVals;
        double sx = 0.0, sy = 0.0, sd = 0.0;
        if (on) {
          for (int t = sub; t < P.sm_blocks; t += L) {
            const double* q = sp + (long long)t * P.S * kSmVals;
            sx += q[0];
            sy += q[1];
            sd += q[3];
          }
        }
        for (int o = 1; o < L; o <<= 1) {
          sx += __shfl_xor_sync(0xffffffffu, sx, o);
          sy += __shfl_xor_sync(0xffffffffu, sy, o);
          sd += __shfl_xor_sync(0xffffffffu, sd, o);
        }
        if (on && sub == 0) {
          const double mean = sd / ((double)P.h[k] * (double)P.w[k]);
          acc[1] += (sx * wk_s[k][0] + sy * wk_s[k][1]) / (mean + (double)P.eps_disp);
        }
      }
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
  // mean inverse depth and sum_p s_p d_p of (b, k), for the smoothness adjoint
  const int bk = blockIdx.x - 1 - BNS, b = bk / P.S, k = bk - b * P.S;
  if (threadIdx.x < 32) {
    const double* sp = smooth_part + ((long long)b * P.sm_blocks * P.S + k) * kSmVals;
    double s = 0.0, sd = 0.0;
    for (int t = lane; t < P.sm_blocks; t += 32) {
      const double* q = sp + (long long)t * P.S * kSmVals;
      s += q[2];
      sd += q[3];
    }
    s = warp_sum(s);
    sd = warp_sum(sd);
    if (lane == 0) {
      saved_scale[bk * kSavedPerScale + 0] = sd / ((double)P.h[k] * (double)P.w[k]);
      saved_scale[bk * kSavedPerScale + 1] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Consistency sweep (BASELINE config 5): per pair, mean pe over valid pixels.  N = 1, S = 1.
// Same strip walk as k_photo_fwd (colvo_photo_fwd.cuh) over the warped frame k_warp_stats left in scratch
// (texel .w = validity): one warp = 32 window columns x kFwdRows rows, separable 3x3 sums, value only.
struct ConsSmem {
  float4 y[kDN];
  float4 x[kDN];
  float2 xb[1];
  double red[kFwdWarps * 2];
};
__global__ void __launch_bounds__(kFwdThreads)
    k_consistency_pe(KP P, const float* __restrict__ ab, const float4* __restrict__ iw, double* __restrict__ pe_part) {
  __shared__ ConsSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * kFwdTileH;
  const int px = x0 + lane;
  {
    const float* tg = static_cast<const float*>(P.tgt) + (long long)b * P.tgt_bf * P.frame_el;
    const float4* xw = iw + (long long)b * P.HW;
    asm volatile("" : "+l"(tg), "+l"(xw));
    const unsigned say = (unsigned)__cvta_generic_to_shared(&sm.y[tid]), sax = (unsigned)__cvta_generic_to_shared(&sm.x[tid]);
    const unsigned hw = P.HW;
#pragma unroll
    for (int j = 0; j < kStageRounds; ++j) {
      const int idx = tid + j * kFwdThreads;
      if (idx < kDN) {
        const int r = idx / kDW, c = idx - r * kDW;
        const unsigned go = reflect_clamp(y0 - 1 + r, P.H) * P.W + reflect_clamp(x0 - 1 + c, P.W);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) cp_async4_s(say + (j * kFwdThreads * 4 + ch) * (unsigned)sizeof(float), tg + (go + ch * hw));
        cp_async16_s(sax + j * kFwdThreads * (unsigned)sizeof(float4), xw + go);
      }
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
  }
  const float cal_a[1] = {__ldg(ab + 2 * b)}, cal_b[1] = {__ldg(ab + 2 * b + 1)};
  const CalV<1> cal = make_calv<1>(cal_a, cal_b, P.alpha);
  const int trow0 = wid * kFwdRows;
  double acc[2] = {0.0, 0.0};
  RowH<1> R[3];
  RowY Y[3];
#pragma unroll
  for (int j = 0; j < kFwdRows + 2; ++j) {
    const int o = (trow0 + j) * kDW + lane;
    row_sums<1, true>(R[j % 3], Y[j % 3], sm.y, sm.x, sm.xb, o);
    if (j >= 2) {
      const int py = y0 + trow0 + j - 2;
      // validity of the window's own pixel: the centre texel of the middle row
      const bool valid = sm.x[(trow0 + j - 1) * kDW + lane + 1].w != 0.f;
      if (py < P.H && px < P.W && valid) {
        WinY wy;
        Vn<1> Sx[3], Sxx[3], Sxy[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float sy = Y[0].sy[c] + Y[1].sy[c] + Y[2].sy[c];
          const float syy = Y[0].syy[c] + Y[1].syy[c] + Y[2].syy[c];
          wy.muy[c] = sy * (1.0f / 9.0f);
          wy.sgy[c] = fmaf(-wy.muy[c], wy.muy[c], syy * (1.0f / 9.0f));
          wy.yc[c] = Y[(j - 1) % 3].yc[c];
        }
        window_sums<1>(R, Sx, Sxx, Sxy);
        winy_derive(wy, P.c1, P.c2);
        acc[0] += (double)(pe_value3v<1>(Sx, Sxx, Sxy, R[(j - 1) % 3].xc, wy, cal, P.alpha, P.c1, P.c2).v * (1.0f / 3.0f));
        acc[1] += 1.0;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double s = warp_sum(acc[i]);
    if (lane == 0) sm.red[wid * 2 + i] = s;
  }
  __syncthreads();
  if (tid < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kFwdWarps; ++w) s += sm.red[w * 2 + tid];
    const int blk = (b * P.ftiles_y + blockIdx.y) * P.ftiles_x + blockIdx.x;
    pe_part[(long long)blk * 2 + tid] = s;
  }
}

__global__ void __launch_bounds__(kThreads)
    k_consistency_final(KP P, const double* __restrict__ pe_part, const float* __restrict__ ab, float* __restrict__ out) {
  __shared__ double sm[(kThreads / 32) * 2];
  const int b = blockIdx.x, tiles = P.ftiles_x * P.ftiles_y;
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
                           float* occ, const SavedView& sv, cudaStream_t st) {
  const bool lcc = (P.flags & 1u) != 0;
  const bool save = (P.flags & 4u) != 0;
  const int need_g = (save && lcc && !(P.flags & 2u)) ? 1 : 0;
  const int BNS = P.B * P.N * P.S;
  const bool pk = (P.flags & 16u) != 0;
  {
    ScopedKernelTimer tm(3, st);
    const bool geo = P.src_depth != nullptr;
    const bool kin = stats_k_inner(P.S);
    dim3 g(Wk.stat_chunks, kin ? P.B : P.B * P.S);
    const int nt = kin ? kStatThreadsK : kThreads;
    auto run = [&](auto kern) { kern<<<g, nt, 0, st>>>(P, Wk.stat_part, valid, Wk.iw, save ? sv.geo : nullptr, occ); };
    auto pick = [&](auto ns, auto geoc, auto pkc) {
      constexpr int NSc = decltype(ns)::value;
      constexpr bool Gc = decltype(geoc)::value, PKc = decltype(pkc)::value;
      if (kin) run(k_warp_stats<NSc, Gc, PKc, 1>);
      else if (P.S > 1) run(k_warp_stats<NSc, Gc, PKc, 2>);
      else run(k_warp_stats<NSc, Gc, PKc, 0>);
    };
    using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
    using T = std::true_type; using F = std::false_type;
    if (P.N == 1) {
      if (geo) { if (pk) pick(I1{}, T{}, T{}); else pick(I1{}, T{}, F{}); }
      else { if (pk) pick(I1{}, F{}, T{}); else pick(I1{}, F{}, F{}); }
    } else {
      if (geo) { if (pk) pick(I2{}, T{}, T{}); else pick(I2{}, T{}, F{}); }
      else { if (pk) pick(I2{}, F{}, T{}); else pick(I2{}, F{}, F{}); }
    }
  }
  cudaError_t e = launch_pdl(k_lcc_solve, dim3(BNS), dim3(32), 0, st, P, Wk.stat_part, Wk.stat_chunks, ab,
                             save ? sv.frame : nullptr);
  if (e != cudaSuccess) return e;
  // The smoothness pass depends on the inputs only; it is launched in two halves of the batch, one behind the
  // statistics pass and one behind the tile kernel, so that each fills the tail of a big launch.
  SmoothOut smo;
  smo.part = Wk.smooth_part;
  for (int k = 0; k < kMaxS; ++k) smo.sf[k] = save ? sv.s_field[k] : nullptr;
  const size_t sm_smem = sizeof(float) * (sm_img_floats() + ((sm_dep_floats() + 1) & ~1)) + sizeof(double) * (kThreads / 32) * kSmVals;
  auto run_smooth = [&](int b0, int nb) {
    if (nb <= 0) return;
    dim3 g(div_up(P.W, kSmBW), div_up(P.H, kSmBH), nb);
    auto run = [&](auto kern) {
      e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), sm_smem);
      if (e != cudaSuccess) return;
      e = launch_pdl(kern, g, dim3(kThreads), sm_smem, st, P, smo, b0);
    };
    if (pk) run(k_smooth<true>); else run(k_smooth<false>);
  };
  const int b_first = (P.B + 1) / 2;
  run_smooth(0, b_first);
  if (e != cudaSuccess) return e;
  dim3 grid(P.ftiles_x, P.ftiles_y, P.B);
  {
    ScopedKernelTimer tm(1, st);
    float4* co = save ? reinterpret_cast<float4*>(sv.coef) : nullptr;
    auto run = [&](auto kern, size_t smem) {
      e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem);
      if (e != cudaSuccess) return;
      e = launch_pdl(kern, grid, dim3(kFwdThreads), smem, st, P, ab, sel, Wk.loss_part, Wk.g_part, need_g, co, Wk.iw);
    };
    auto pick = [&](auto ns, auto pkc) {          // ADJ: the adjoint pieces are needed only when the forward saves for a backward
      constexpr int NSc = decltype(ns)::value;
      constexpr bool PKc = decltype(pkc)::value;
      if (save) run(k_photo_fwd<NSc, PKc, true>, photo_fwd_smem<NSc>());
      else run(k_photo_fwd<NSc, PKc, false>, photo_fwd_smem<NSc>());
    };
    using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
    if (P.N == 1) { if (pk) pick(I1{}, std::true_type{}); else pick(I1{}, std::false_type{}); }
    else { if (pk) pick(I2{}, std::true_type{}); else pick(I2{}, std::false_type{}); }
  }
  if (e != cudaSuccess) return e;
  run_smooth(b_first, P.B - b_first);
  if (e != cudaSuccess) return e;
  const int nfin = 1 + (save ? BNS + P.B * P.S : 0);
  e = launch_pdl(k_finalize_fwd, dim3(nfin), dim3(kThreads), 0, st, P, (const double*)Wk.loss_part, (const double*)Wk.g_part,
                 (const double*)Wk.smooth_part, (const double*)Wk.stat_part, Wk.stat_chunks, loss, sv.frame, sv.scale, need_g);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

cudaError_t launch_consistency(const KP& P0, double* stat_part, int stat_chunks, double* pe_part, float* ab, float* out,
                               float4* iw, int pairs_per_pass, cudaStream_t st) {
  // The warped frames go through a scratch buffer of `pairs_per_pass` frames: the sweep runs in passes of that many
  // pairs (statistics + warp, (a, b), photometric error), each pass re-using the buffer.
  const int tiles = P0.ftiles_x * P0.ftiles_y;
  for (int p0 = 0; p0 < P0.B; p0 += pairs_per_pass) {
    KP P = P0;
    P.B = (P0.B - p0 < pairs_per_pass) ? P0.B - p0 : pairs_per_pass;
    P.tgt = static_cast<const float*>(P0.tgt) + (long long)p0 * P0.tgt_bf * P0.frame_el;
    P.srcs = static_cast<const float*>(P0.srcs) + (long long)p0 * P0.src_bf * P0.frame_el;
    P.depth[0] = P0.depth[0] + (long long)p0 * P0.depth_bs[0];
    P.K = P0.K + (long long)p0 * P0.K_bs;
    P.T = P0.T + (long long)p0 * P0.T_bs;
    double* sp = stat_part + (long long)p0 * stat_chunks * kStatVals;
    float* abp = ab + 2 * p0;
    {
      ScopedKernelTimer tm(3, st);
      k_warp_stats<1, false, false, 0><<<dim3(stat_chunks, P.B), kThreads, 0, st>>>(P, sp, nullptr, iw, nullptr, nullptr);
    }
    k_lcc_solve<<<P.B, 32, 0, st>>>(P, sp, stat_chunks, abp, nullptr);
    {
      ScopedKernelTimer tm(4, st);
      k_consistency_pe<<<dim3(P.ftiles_x, P.ftiles_y, P.B), kFwdThreads, 0, st>>>(P, abp, iw, pe_part + (long long)p0 * tiles * 2);
    }
  }
  k_consistency_final<<<P0.B, kThreads, 0, st>>>(P0, pe_part, ab, out);
  return cudaGetLastError();
}

}  // namespace colvo
