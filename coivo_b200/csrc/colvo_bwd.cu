// Backward kernels of the ColVO photometric-loss path (SURVEY.md section 8(a) row 11 and
// appendix A), hand-written for sm_100a.  The adjoint is analytic (no tape): the forward
// saves sel, (a, b), the LCC statistics, dL/da, dL/db, the smoothness adjoint field, the SSIM
// adjoint coefficients of the winning candidate and the projection (u', v', iz, D^) of every pixel.
//
//   k_zero           zero-fill of the scatter accumulators (a kernel, so the next one can be launched behind it
//                    programmatically)
//   k_photo_bwd      per 32 x kBwdTileH tile: SSIM adjoint gathered from the saved coefficient fields (tiles staged
//                    by cp.async + mbarriers), four taps per source at the saved coordinates, LCC adjoint,
//                    bilinear scatter-add as one vector RED per tap into a texel-interleaved accumulator,
//                    projection adjoint -> full-resolution depth adjoint + per-tile pose-gradient partials;
//                    scale 0 also gets its smoothness gradient
//   k_depth_gather   adjoint of the bilinear depth up-sample in gather form (no atomics) plus the smoothness
//                    gradient of scales k >= 1; its extra grid rows do the deterministic pose reduction
//                    (pose_final_block) and unpack the source-gradient texels into the planar grad_srcs
//   k_pose_final     the same epilogue blocks on their own when S = 1
#include <string.h>

#include <type_traits>

#include "colvo_kernels.cuh"

// Ablation builds for the timing experiments logged under profiles/ (scripts/build_variants.py): they skip the
// coefficient gather / the scatter and therefore give WRONG gradients; the package never loads such a build.
#ifndef COLVO_EXP_NOGATHER
#define COLVO_EXP_NOGATHER 0
#endif
#ifndef COLVO_EXP_NORED
#define COLVO_EXP_NORED 0
#endif
// Packed fp32x2 in the backward, measured on B200 (profiles/r2_bwd_packed_variants.log): the kernel lives on gather
// latency at 24 warps / 80 registers per thread, and the packed forms need more live registers than that budget holds
// (FFMA2 issues at half the FFMA rate, so it saves issue slots, not FMA-pipe time -- scripts/microbench/ffma2_bench.cu):
//   chain packed + gather packed 241 us,  chain scalar + gather packed 214 us,  both scalar 199 us  <- default
#ifndef COLVO_BWD_CHAIN_PACKED   // 1: the per-source chain of k_photo_bwd runs both sources in packed fp32x2 lanes; 0: one source
#define COLVO_BWD_CHAIN_PACKED 0 //    after the other
#endif
#ifndef COLVO_BWD_GATHER_PACKED  // 1: one FFMA2 per coefficient feeds both sources' accumulators; 0: scalar FFMA per source
#define COLVO_BWD_GATHER_PACKED 0
#endif
// Warp-aggregated scatter (north_star: "avoids contended global atomics"): coincident taps of neighbouring pixels are
// summed inside the warp, through shared-memory exchange slots, before they go to global memory -- two 16-byte REDs per
// pixel, source and scale instead of four (L2 atomic sectors 68 -> 36 per warp).  Built, parity-green
// (profiles/r2_scatter_merge.log) and measured SLOWER on B200: 225 us against 199 us for the plain vector-RED scatter
// (211 vs 203 us at 96 registers): the exchange costs more issue slots and shared-memory bandwidth -- which this kernel
// is short of -- than the L2 sectors it saves.  Both forms ship: the descriptor flag COLVO_F_SCATTER_MERGE selects the
// aggregated one at run time (template parameter MERGE), the default is the plain vector-RED scatter.
#ifndef COLVO_BWD_TMA       // 1: the coefficient tile of a scale (34 x 6 windows x 3 channels of 16-byte texels) is ONE TMA
#define COLVO_BWD_TMA 1     //    box load issued by one thread (UTMALDG; zero fill outside the image by the hardware);
#endif                      //    0: 16-byte cp.async (LDGSTS) by every thread
#ifndef COLVO_MINB_BWD      // CTAs per SM the register allocator must allow -- tuned on B200, see DESIGN.md
#define COLVO_MINB_BWD (20 / COLVO_BWD_TILE_H)     // 20 warps per SM at 96 registers (24 warps at 80 registers spill)
#endif

#ifdef COLVO_DEBUG_DUMP     // diagnostic builds only (tests/tools): per-pixel internals of k_photo_bwd at k = 0, source 0
__device__ float* g_colvo_dbg = nullptr;
extern "C" int colvo_debug_set_buffer(void* p) {
  float* q = static_cast<float*>(p);
  return (int)cudaMemcpyToSymbol(g_colvo_dbg, &q, sizeof(q));
}
#endif

namespace colvo {

constexpr int kCH = kBwdTileH + 2, kCW = kTileW + 2;   // tile + 1-pixel halo (window centres)
constexpr int kCN = kCH * kCW;
constexpr int kCoefBuf = (kCN * 3 + 7) & ~7;           // float4 per coefficient buffer, padded to 128 bytes (TMA destination)

// smoothness gradient of one depth texel from the saved adjoint field (grad_loss folded in by the caller)
__device__ __forceinline__ float smooth_grad(float s, float D, float inv, float corr) {
  const float dr = 1.0f / D;
  return -(s * inv - corr) * dr * dr;
}

// One CTA = one 32 x kBwdTileH (4) tile of one triplet.  The SSIM adjoint coefficients of every window were
// written by the forward (for the winning candidate; the winner flags ride in .w of channels 0 / 1), and
// k_warp_stats saved the projection (u', v', iz, D^, valid) of every pixel, so the tile needs neither a halo
// re-warp nor a re-projection: per scale the CTA stages the coefficient tile (+1 halo) in shared memory, every
// thread gathers its 3x3 neighbourhood, samples the four taps of each source at the saved coordinates, and
// pushes the result through the LCC, bilinear and projection adjoints.
//
// Blackwell specifics: with N = 2 the two sources ride in the two lanes of packed fp32x2 registers (colvo_f2.cuh)
// from the gather (one FFMA2 per coefficient feeds both sources' accumulators: the lane weights are the window's
// winner flags) through the bilinear / LCC / projection adjoints and the pose sums; only the tap addressing and
// the scatter products stay per source.  The saved projection is stored lane-interleaved for that:
//   N = 1:  geo[(b,k)][pix]  = (u', v', iz, D^ | valid in the mantissa LSB)
//   N = 2:  geoA[(b,k)][pix] = (u'^0, u'^1, v'^0, v'^1),  geoB[(b,k)][pix] = (iz^0, iz^1, D^, valid bits)
template <int NS>
struct BwdConstV {         // per scale, lane n = warped frame (n, k); built once per CTA
  Vn<NS> a, b;             // LCC gain / bias
  Vn<NS> l0, ly, lx;       // LCC adjoint of a valid sample: l0 + ly * y + lx * x
  Vn<NS> wl1;              // weight of the L1 term's sign at the own pixel: wscale * (1 - alpha) / 3 * a
};

template <int NS, bool GEO, bool PK, bool MERGE>
__global__ void __launch_bounds__(kBwdThreads, COLVO_MINB_BWD)
    k_photo_bwd(KP P, const float* __restrict__ grad_loss, const uint8_t* __restrict__ sel,
                const double* __restrict__ saved_frame, const double* __restrict__ saved_scale,
                const float* __restrict__ s_field0, const float* __restrict__ coef_in,
                const float4* __restrict__ geo_in, float* __restrict__ grad_d0,
                float* __restrict__ dD1, float* __restrict__ dD2, float* __restrict__ dD3,
                float4* __restrict__ gsrc4, float* __restrict__ grad_src_depth, double* __restrict__ pose_part,
                const __grid_constant__ CUtensorMap coef_map) {
  typedef Vn<NS> V;
  // dynamic shared memory, carved by hand
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 (*coef)[kCoefBuf] = reinterpret_cast<float4 (*)[kCoefBuf]>(smem_raw);   // [2]: (ca, cb, cg, flag) per window
                                                                               // centre and channel, double-buffered over k
  double* red = reinterpret_cast<double*>(smem_raw + sizeof(float4) * 2 * kCoefBuf);
  BwdConstV<NS>* cst = reinterpret_cast<BwdConstV<NS>*>(red + (kBwdThreads / 32) * NS * 12);    // [kMaxS]
  V* pose_s = reinterpret_cast<V*>(cst + kMaxS);                                            // [12]: R row-major, t; lane n
  float* cst_sm = reinterpret_cast<float*>(pose_s + 12);   // scale 0: 1/(mean+eps), sum(s d)/(n (mean+eps)^2)
  // full[i]: the copies into coefficient buffer i have landed (cp.async arrivals); empty[i]: every thread is done reading it
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem_raw + sizeof(float4) * 2 * kCoefBuf +
                                                                   sizeof(double) * (kBwdThreads / 32) * NS * 12 + 1024);
  static_assert(sizeof(BwdConstV<NS>) * kMaxS + sizeof(V) * 12 + 2 * sizeof(float) <= 1024, "constant area");
  // scatter exchange slots, [warp][lane][2]: lane l parks its two x1-column taps (value, texel offset) for lane l + 1
  float4 (*xch)[2] = reinterpret_cast<float4 (*)[2]>(reinterpret_cast<unsigned char*>(mbar) + 64) + (threadIdx.x >> 5) * 32;

  pdl_trigger();         // the epilogue launch may become resident while this kernel drains
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int b = blockIdx.z, x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kBwdTileH;
  const int px = x0 + tx, py = y0 + ty;
  const bool in_img = (px < P.W) && (py < P.H);
  const int qx = imin(px, P.W - 1), qy = imin(py, P.H - 1);             // addressable stand-in when outside
  const int qo = qy * P.W + qx;
  const Img<PK> tg = img_at<PK>(P, P.tgt, b * P.tgt_bf);
  const Cam cam = load_cam(P, b);
  const float go = __ldg(grad_loss);
  const float wscale = go / ((float)P.S * (float)P.B * (float)P.HW);

  // ---- phase 0: per-frame constants ----
  if (tid < NS * kMaxS) {
    const int n = tid / kMaxS, k = tid % kMaxS;
    float ca = 1.f, cb = 0.f, l0 = 0.f, ly = 0.f, lx = 0.f;
    if (k < P.S) {
      const double* s = saved_frame + ((long long)(b * P.N + n) * P.S + k) * kSavedPerFrame;
      ca = (float)s[4];
      cb = (float)s[5];
      if ((P.flags & 1u) && !(P.flags & 2u) && s[0] > 0.0) {
        // lcc_q = Pc * ((y - my) - 2 a (x - mx)) - Qc   (SURVEY.md appendix A), expanded in y and x
        const double Pc = (double)go * (s[6] - s[7] * s[1]) * s[3];
        const double Qc = (double)go * s[7] * s[4] / s[0];
        const double a = s[4];
        ly = (float)Pc;
        lx = (float)(-2.0 * a * Pc);
        l0 = (float)(-Pc * s[2] + 2.0 * a * Pc * s[1] - Qc);
      }
    }
    BwdConstV<NS>& c = cst[k];
    c.a.set(n, ca); c.b.set(n, cb); c.l0.set(n, l0); c.ly.set(n, ly); c.lx.set(n, lx);
    c.wl1.set(n, wscale * (1.f - P.alpha) * (1.0f / 3.0f) * ca);
  }
  if (tid >= 32 && tid < 32 + NS * 12) {
    const int n = (tid - 32) / 12, j = (tid - 32) % 12;
    const float* t = P.T + (long long)b * P.T_bs + (long long)n * P.T_ns;
    pose_s[j].set(n, (j < 9) ? __ldg(t + 4 * (j / 3) + (j % 3)) : __ldg(t + 4 * (j - 9) + 3));
  }
  if (tid == kBwdThreads - 1) {
    const double* sc = saved_scale + (long long)(b * P.S) * kSavedPerScale;
    const double me = sc[0] + (double)P.eps_disp;
    cst_sm[0] = (float)(1.0 / me);
    cst_sm[1] = (float)(sc[1] / ((double)P.HW * me * me));
  }

  // reflect-padding multiplicities of the 3x3 gather at the own pixel (per axis); the loss weight rides on the rows
  float my3[3], mx3[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    my3[d] = wscale * reflect_mult(py, py + d - 1, P.H);
    mx3[d] = reflect_mult(px, px + d - 1, P.W);
  }
  float yq[3];
  tg.load3(qo, yq);
  const float own_rx = ray_x(qx, cam), own_ry = ray_y(qy, cam);
  const int oc = ty * kCW + tx;

  // pose gradient: per source only sum dXp_i * D and sum dXp_i are carried (see project_adjoint)
  V pw[3], pt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pw[i] = pt[i] = bc<NS>(0.f);

  // Coefficient tile of scale k (+1 halo; zeros outside the image), fetched with cp.async one scale ahead so its
  // global latency hides behind the previous scale's arithmetic.
  auto stage_coef = [&](int k) {
#if COLVO_BWD_TMA
    // one box load: [3 channels][kCH rows][kCW windows] of 16-byte texels, box origin (x0 - 1, y0 - 1); the part of the
    // box outside the image arrives as zeros.  The bytes land on full[k & 1].
    if (tid == 0) {
      mbar_arrive_expect_tx(&mbar[k & 1], (unsigned)(sizeof(float4) * kCN * 3));
      tma_load_3d(coef[k & 1], &coef_map, 4 * (x0 - 1), y0 - 1, (b * P.S + k) * 3, &mbar[k & 1]);
    }
#else
    float4* cbf = coef[k & 1];
    const float4* cin = reinterpret_cast<const float4*>(coef_in) + (long long)(b * P.S + k) * P.HW * 3;
    for (int idx = tid; idx < kCN; idx += kBwdThreads) {
      const int r = idx / kCW, c = idx - r * kCW;
      const int gy = y0 - 1 + r, gx = x0 - 1 + c;
      const bool on = gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
      const float4* p = on ? cin + (gy * P.W + gx) : cin;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) cp_async16(cbf + ch * kCN + idx, p + (long long)ch * P.HW, on);
    }
    mbar_arrive_cp_async(&mbar[k & 1]);      // this thread's share of full[k & 1]
#endif
  };
  // warp-aggregated scatter: the active lanes of a warp are a prefix (one tile row: px grows with the lane); lane 0 has no
  // left neighbour: its exchange slot keeps a sentinel offset that matches no texel
  const unsigned amask = MERGE ? __ballot_sync(0xffffffffu, in_img) : 0u;
  if (MERGE && tx == 0) xch[0][0] = xch[0][1] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&mbar[i], (COLVO_BWD_TMA && i < 2) ? 1 : kBwdThreads);   // full[0..1], empty[0..1]
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();       // barriers initialised; per-frame constants visible
  stage_coef(0);

  // Everything above reads only what the forward saved.  Launched programmatically behind k_zero, the scatter targets
  // are guaranteed to be zero from here on.
  pdl_wait();
#pragma unroll 1
  for (int k = 0; k < P.S; ++k) {
    const float4* cb = coef[k & 1];
    // projection of the own pixel into every source, and which candidate won here
    V gu, gv, giz;
    float D_own;
    bool gvalid[NS];
    if constexpr (NS == 1) {
      const float4 g4 = ld_stream(geo_in + (long long)(b * P.S + k) * P.HW + qo);
      gu = V(g4.x); gv = V(g4.y); giz = V(g4.z);
      D_own = g4.w;
      gvalid[0] = (__float_as_uint(g4.w) & 1u) != 0u;
    } else {
      const float4 ga4 = ld_stream(geo_in + (long long)(b * P.S + k) * P.HW + qo);
      const float4 gb4 = ld_stream(geo_in + (long long)((P.B + b) * P.S + k) * P.HW + qo);
      gu = V(ga4.x, ga4.y); gv = V(ga4.z, ga4.w); giz = V(gb4.x, gb4.y);
      D_own = gb4.z;
      const unsigned vb = __float_as_uint(gb4.w);
      gvalid[0] = (vb & 1u) != 0u;
      gvalid[NS - 1] = (vb & 2u) != 0u;
    }
    const int own_sel = __ldg(sel + ((long long)b * P.S + k) * P.HW + qo);
    // Two transaction barriers per buffer instead of a block barrier per scale: the copies of scale k+1 are issued
    // once every thread has finished gathering from that buffer (scale k-1: a whole per-source phase ago), and the
    // gather of scale k starts once the copies into its buffer have landed -- warps drift by up to one scale.
    if (k + 1 < P.S) {
      if (k >= 1 && (!COLVO_BWD_TMA || tid == 0))
        mbar_wait(&mbar[2 + ((k + 1) & 1)], ((k - 1) >> 1) & 1);               // empty[(k+1)&1]: gather(k-1) done everywhere
      stage_coef(k + 1);
    }
    mbar_wait(&mbar[k & 1], (k >> 1) & 1);                                      // full[k&1]: scale k landed
    // gather once per scale: every window centre has at most one winning source (its flags weight the lanes), so
    // one pass over the 3x3 neighbourhood serves all sources with exact zeros for the others (the three sums cancel
    // heavily against each other downstream)
    V A[3], Bc[3], G[3];
    CV_CHECK(oc >= 0 && oc + 2 * kCW + 2 < kCN && 3 * kCN <= kCoefBuf);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) A[ch] = Bc[ch] = G[ch] = bc<NS>(0.f);
    if (in_img && !COLVO_EXP_NOGATHER) {
#if COLVO_BWD_GATHER_PACKED
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int o = oc + (j / 3) * kCW + (j % 3);
        const float m = my3[j / 3] * mx3[j % 3];
        const float4 q3[3] = {cb[o], cb[kCN + o], cb[2 * kCN + o]};
        V mn;
        mn.set(0, m * q3[0].w);
        if (NS > 1) mn.set(NS - 1, m * q3[1].w);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          A[ch] = fma2(mn, bc<NS>(q3[ch].x), A[ch]);
          Bc[ch] = fma2(mn, bc<NS>(q3[ch].y), Bc[ch]);
          G[ch] = fma2(mn, bc<NS>(q3[ch].z), G[ch]);
        }
      }
#else
      float As[NS][3], Bs[NS][3], Gs[NS][3];
#pragma unroll
      for (int n = 0; n < NS; ++n)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) As[n][ch] = Bs[n][ch] = Gs[n][ch] = 0.f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int o = oc + (j / 3) * kCW + (j % 3);
        const float m = my3[j / 3] * mx3[j % 3];
        const float4 q3[3] = {cb[o], cb[kCN + o], cb[2 * kCN + o]};
        float mn[NS];
        mn[0] = m * q3[0].w;
        if (NS > 1) mn[NS - 1] = m * q3[1].w;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
          for (int n = 0; n < NS; ++n) {
            As[n][ch] = fmaf(mn[n], q3[ch].x, As[n][ch]);
            Bs[n][ch] = fmaf(mn[n], q3[ch].y, Bs[n][ch]);
            Gs[n][ch] = fmaf(mn[n], q3[ch].z, Gs[n][ch]);
          }
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int n = 0; n < NS; ++n) { A[ch].set(n, As[n][ch]); Bc[ch].set(n, Bs[n][ch]); G[ch].set(n, Gs[n][ch]); }
#endif
    }
    mbar_arrive(&mbar[2 + (k & 1)]);          // done reading coefficient buffer k & 1
    float dD = 0.f;
    if (in_img) {
      const BwdConstV<NS> cc = cst[k];
      // The per-source chain runs on M lanes at a time: M = NS (both sources packed) or M = 1 (one source after the
      // other: fewer live registers, more issue slots) -- COLVO_BWD_CHAIN_PACKED.
      constexpr int M = (NS > 1 && COLVO_BWD_CHAIN_PACKED) ? NS : 1;
      typedef Vn<M> VM;
      typedef LaneSub<M, NS> LS;
#pragma unroll
      for (int n0 = 0; n0 < NS; n0 += M) {
        const VM gu_ = LS::get(gu, n0), gv_ = LS::get(gv, n0), giz_ = LS::get(giz, n0);
        const VM c_a = LS::get(cc.a, n0), c_b = LS::get(cc.b, n0), c_l0 = LS::get(cc.l0, n0), c_ly = LS::get(cc.ly, n0),
                 c_lx = LS::get(cc.lx, n0);
        // taps and texels at the saved coordinates
        Taps t[M];
        Texels tx4[M];
        int r0[M], r1[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
          const Img<PK> src = img_at<PK>(P, P.srcs, b * P.src_bf + (n0 + i) * P.src_nf);
          t[i] = make_taps(gu_.lane(i), gv_.lane(i), P.W, P.H);
          CV_CHECK_TAPS(t[i], P.W, P.H);
          r0[i] = t[i].y0 * P.W;
          r1[i] = t[i].y1 * P.W;
          src.load_taps(r0[i] + t[i].x0, r0[i] + t[i].x1, r1[i] + t[i].x0, r1[i] + t[i].x1, tx4[i]);
        }
        VM wx, wy, wl1, vm, gxm, gym;
#pragma unroll
        for (int i = 0; i < M; ++i) {
          wx.set(i, t[i].wx);
          wy.set(i, t[i].wy);
          wl1.set(i, (own_sel == NS + n0 + i) ? cc.wl1.lane(n0 + i) : 0.f);
          vm.set(i, gvalid[n0 + i] ? 1.f : 0.f);
          gxm.set(i, t[i].gx ? 1.f : 0.f);       // the coordinate gradient passes strictly inside the border only
          gym.set(i, t[i].gy ? 1.f : 0.f);
        }
        VM du = bc<M>(0.f), dv = bc<M>(0.f), hq[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          VM i00, i01, i10, i11;
#pragma unroll
          for (int i = 0; i < M; ++i) {
            i00.set(i, tx4[i].i00[ch]); i01.set(i, tx4[i].i01[ch]); i10.set(i, tx4[i].i10[ch]); i11.set(i, tx4[i].i11[ch]);
          }
          // bilinear sample and its coordinate derivatives from one cross term
          const VM d01 = i01 - i00, e10 = i10 - i00;
          const VM cross = (i11 - i10) - d01;
          const VM dux = fma2(wy, cross, d01), dvy = fma2(wx, cross, e10);
          const VM xq = fma2(wy, dvy, fma2(wx, d01, i00));
          const VM diff = fma2(c_a, xq, c_b) - bc<M>(yq[ch]);
          VM h = fma2(xq, LS::get(Bc[ch], n0), fma2(bc<M>(yq[ch]), LS::get(G[ch], n0), LS::get(A[ch], n0)));
          VM sg;
#pragma unroll
          for (int i = 0; i < M; ++i) sg.set(i, sgn_mul(wl1.lane(i), diff.lane(i)));
          const VM lcc = fma2(c_lx, xq, fma2(c_ly, bc<M>(yq[ch]), c_l0));
#ifdef COLVO_DEBUG_DUMP
          if (g_colvo_dbg && k == 0 && n0 == 0 && ch == 0) {
            float* o = g_colvo_dbg + ((long long)b * P.HW + py * P.W + px) * 12;
            o[0] = c_a.lane(0); o[1] = c_b.lane(0); o[2] = wl1.lane(0); o[3] = (float)own_sel; o[4] = diff.lane(0);
            o[5] = sg.lane(0); o[6] = h.lane(0); o[7] = xq.lane(0); o[8] = yq[0]; o[9] = lcc.lane(0); o[10] = vm.lane(0);
            o[11] = cc.wl1.lane(0);
          }
#endif
          h = fma2(vm, lcc, h + sg);
          hq[ch] = h;
          du = fma2(h, dux, du);
          dv = fma2(h, dvy, dv);
        }
        const VM w11 = wx * wy, w01 = wx - w11, w10 = wy - w11, w00 = (bc<M>(1.f) - wx) - w10;
        if (gsrc4 && !COLVO_EXP_NORED) {
          // scatter into the texel-interleaved gradient buffer: one 16-byte vector RED per tap carries the three
          // channels (a third of the L2 atomic requests of a planar scatter); null for packed sources, whose
          // quantised images carry no gradient
#pragma unroll
          for (int i = 0; i < M; ++i) {
            float4* gs = gsrc4 + (long long)(b * P.N + n0 + i) * P.HW;
            asm volatile("" : "+l"(gs));      // materialised base: one IMAD.WIDE per address (see Img<false>::load_taps)
            const float a00 = w00.lane(i), a01 = w01.lane(i), a10 = w10.lane(i), a11 = w11.lane(i);
            const float h0 = hq[0].lane(i), h1 = hq[1].lane(i), h2 = hq[2].lane(i);
            const int o00 = r0[i] + t[i].x0, o01 = r0[i] + t[i].x1, o10 = r1[i] + t[i].x0, o11 = r1[i] + t[i].x1;
            CV_CHECK(o00 >= 0 && o11 < P.HW && o01 < P.HW && o10 < P.HW);
            if constexpr (MERGE) {
            // Warp-aggregated scatter (north_star: no contended global atomics).  Neighbouring pixels of a row sample
            // neighbouring texels, so the right-hand taps (x1, y0), (x1, y1) of lane l usually ARE the left-hand taps
            // (x0, y0), (x0, y1) of lane l + 1.  Every lane parks its x1 taps -- three channel values and the texel
            // offset -- in the slot of its right neighbour, which adds them to its own x0 taps when the offsets agree and
            // forwards them unchanged when they do not (nothing is lost for any flow field): per pixel, source and
            // scale two 16-byte REDs leave the SM instead of four.  Shared-memory stores / loads, no shared atomics.
            const bool has_right = (tx < 31) && ((amask >> (tx + 1)) & 1u);
            // tap row y0, then tap row y1 (one after the other: fewer live registers; slot [0] / [1] alternate, so one
            // warp barrier per row also orders the re-use of the other slot)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const float ax0 = r ? a10 : a00, ax1 = r ? a11 : a01;
              const int ox0 = r ? o10 : o00, ox1 = r ? o11 : o01;
              if (has_right) xch[tx + 1][r] = make_float4(ax1 * h0, ax1 * h1, ax1 * h2, __int_as_float(ox1));
              __syncwarp(amask);
              const float4 d = xch[tx][r];
              const int f = __float_as_int(d.w);
              const bool m = f == ox0;
              red_add3(gs + (unsigned)ox0, fmaf(ax0, h0, m ? d.x : 0.f), fmaf(ax0, h1, m ? d.y : 0.f), fmaf(ax0, h2, m ? d.z : 0.f));
              if (!m && f >= 0) red_add3(gs + (unsigned)f, d.x, d.y, d.z);          // the neighbour's tap lies elsewhere
              if (!has_right) red_add3(gs + (unsigned)ox1, ax1 * h0, ax1 * h1, ax1 * h2);   // right-most pixel of the row segment
            }
            } else {
            red_add3(gs + (unsigned)o00, a00 * h0, a00 * h1, a00 * h2);
            red_add3(gs + (unsigned)o01, a01 * h0, a01 * h1, a01 * h2);
            red_add3(gs + (unsigned)o10, a10 * h0, a10 * h1, a10 * h2);
            red_add3(gs + (unsigned)o11, a11 * h0, a11 * h1, a11 * h2);
            }
          }
        }
        // geometric consistency (f-2): gradient to Z' directly, to the sampled source depth (scatter) and,
        // through its spatial derivative, to (u', v')
        VM dZp_direct = bc<M>(0.f);
        if (GEO) {
#pragma unroll
          for (int i = 0; i < M; ++i) {
            if (gvalid[n0 + i]) {
              float d4[4], dZ, dS;
              const float ds = sample_plane(P.src_depth + (long long)(b * P.N + n0 + i) * P.HW, t[i], P.W, d4);
              geo_diff(p_sub(p_rcp(giz_.lane(i)), P.eps_proj), ds, dZ, dS);     // Z' back from iz = 1 / (Z' + eps)
              const float wg = go * P.geo_weight / ((float)P.S * (float)P.B * (float)P.N * (float)P.HW);
              dZp_direct.set(i, wg * dZ);
              const float gS = wg * dS;
              du.set(i, du.lane(i) + gS * ((1.f - t[i].wy) * (d4[1] - d4[0]) + t[i].wy * (d4[3] - d4[2])));
              dv.set(i, dv.lane(i) + gS * ((1.f - t[i].wx) * (d4[2] - d4[0]) + t[i].wx * (d4[3] - d4[1])));
              if (grad_src_depth) {
                float* gd = grad_src_depth + (long long)(b * P.N + n0 + i) * P.HW;
                atomicAdd(gd + (r0[i] + t[i].x0), w00.lane(i) * gS);
                atomicAdd(gd + (r0[i] + t[i].x1), w01.lane(i) * gS);
                atomicAdd(gd + (r1[i] + t[i].x0), w10.lane(i) * gS);
                atomicAdd(gd + (r1[i] + t[i].x1), w11.lane(i) * gS);
              }
            }
          }
        }
        du = du * gxm;
        dv = dv * gym;
        // projection / transform / back-projection adjoint (colvo_math.cuh::project_adjoint)
        const VM dx = du * giz_, dy = dv * giz_;
        VM dXp[3];
        dXp[0] = bc<M>(cam.fx) * dx;
        dXp[1] = bc<M>(cam.fy) * dy;
        dXp[2] = fma2(bc<M>(cam.cx), dx, fma2(bc<M>(cam.cy), dy, fma2(-giz_, fma2(du, gu_, dv * gv_), dZp_direct)));
        VM R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = LS::get(pose_s[i], n0);
        const VM dX = fma2(R[0], dXp[0], fma2(R[3], dXp[1], R[6] * dXp[2]));
        const VM dY = fma2(R[1], dXp[0], fma2(R[4], dXp[1], R[7] * dXp[2]));
        const VM dZ = fma2(R[2], dXp[0], fma2(R[5], dXp[1], R[8] * dXp[2]));
        const VM dDv = fma2(bc<M>(own_rx), dX, fma2(bc<M>(own_ry), dY, dZ));
#pragma unroll
        for (int i = 0; i < M; ++i) dD += dDv.lane(i);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          LS::put(pw[i], n0, fma2(dXp[i], bc<M>(D_own), LS::get(pw[i], n0)));
          LS::put(pt[i], n0, LS::get(pt[i], n0) + dXp[i]);
        }
      }
      const int p = py * P.W + px;
      if (k == 0) {
        const float s = __ldg(s_field0 + (long long)b * P.HW + p);
        grad_d0[(long long)b * P.depth_bs[0] + p] = dD + go * smooth_grad(s, D_own, cst_sm[0], cst_sm[1]);
      } else {
        float* o = ((k == 1) ? dD1 : (k == 2 ? dD2 : dD3)) + (long long)b * P.HW;
        o[p] = dD;
      }
    }
  }

  // ---- per-tile pose-gradient partials (fp64: sums of terms of both signs) ----
  const int blk = (b * P.btiles_y + blockIdx.y) * P.tiles_x + blockIdx.x;
  __syncthreads();                                   // the coefficient buffers are free: reuse them as slot storage
  float* slots = reinterpret_cast<float*>(smem_raw); // [NS*12][kBwdThreads] <= 24 KB of the 32 KB coefficient area
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    float gp[12];
    const float w[3] = {pw[0].lane(n), pw[1].lane(n), pw[2].lane(n)}, tt[3] = {pt[0].lane(n), pt[1].lane(n), pt[2].lane(n)};
    pose_grad_expand(w, tt, own_rx, own_ry, gp);
#pragma unroll
    for (int j = 0; j < 12; ++j) slots[(n * 12 + j) * kBwdThreads + tid] = in_img ? gp[j] : 0.f;
  }
  __syncthreads();
  block_sum_slots<NS * 12, kBwdThreads>(slots, red, [&](int slot, double v) { pose_part[(long long)blk * (NS * 12) + slot] = v; });
}

// ------------------------------------------------------------------------------------------
// grad_T[b,n] from the per-tile partials: one CTA per (b,n); warp w reduces entries w, w+8.
// The blocks beyond B*N unpack the texel-interleaved source gradient into the planar grad_srcs [B,N,3,H,W]
// (kUnpackPix pixels of one (b,n) frame each; independent of the pose reduction, so it shares the launch).
constexpr int kUnpackPix = 4 * kThreads;
struct PoseFinalArgs {
  const double* pose_part;
  float* grad_T;
  const float4* gsrc4;
  float* grad_srcs;
  int unpack_chunks;
  int n_blocks;          // B*N pose blocks + B*N*unpack_chunks unpack blocks
};
__device__ __forceinline__ void pose_final_block(const KP& P, const PoseFinalArgs& A, int blk) {
  const double* __restrict__ pose_part = A.pose_part;
  float* __restrict__ grad_T = A.grad_T;
  const float4* __restrict__ gsrc4 = A.gsrc4;
  float* __restrict__ grad_srcs = A.grad_srcs;
  const int unpack_chunks = A.unpack_chunks;
  if (blk >= P.B * P.N) {
    const int j = blk - P.B * P.N, bn = j / unpack_chunks, c = j - bn * unpack_chunks;
    const float4* src = gsrc4 + (long long)bn * P.HW;
    float* dst = grad_srcs + (long long)bn * 3 * P.HW;
#pragma unroll
    for (int i = 0; i < kUnpackPix / kThreads; ++i) {
      const int p = c * kUnpackPix + i * kThreads + threadIdx.x;
      if (p < P.HW) {
        const float4 g = __ldcs(src + p);
        dst[p] = g.x;
        dst[P.HW + p] = g.y;
        dst[2 * (long long)P.HW + p] = g.z;
      }
    }
    return;
  }
  const int bn = blk, b = bn / P.N, n = bn % P.N;
  const int tiles = P.tiles_x * P.btiles_y, nv = P.N * 12;
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

__global__ void __launch_bounds__(kThreads) k_pose_final(KP P, PoseFinalArgs A) {
  pdl_wait();            // launched programmatically behind k_photo_bwd
  pose_final_block(P, A, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// Adjoint of upsample_depth in gather form: low-res texel (i, j) sums every full-res pixel whose
// 2x2 bilinear footprint contains it.  4^(k-1) lanes share one texel (its footprint grows as 4^k)
// and combine by shuffles in a fixed order, so the result is deterministic.  The smoothness
// gradient of the scale is added on the way out.  blockIdx.y = b * (S-1) + (k-1).
__global__ void __launch_bounds__(kThreads)
    k_depth_gather(KP P, const float* __restrict__ grad_loss, const double* __restrict__ saved_scale,
                   const float* __restrict__ dD1, const float* __restrict__ dD2, const float* __restrict__ dD3,
                   const float* __restrict__ sf1, const float* __restrict__ sf2, const float* __restrict__ sf3,
                   float* g1, float* g2, float* g3, PoseFinalArgs A) {
  pdl_wait();            // launched programmatically behind k_photo_bwd
  if ((int)blockIdx.y >= P.B * (P.S - 1)) {     // the grid rows beyond the (b, k) pairs: pose reduction and source-gradient unpack
    const int j = ((int)blockIdx.y - P.B * (P.S - 1)) * gridDim.x + blockIdx.x;
    if (j < A.n_blocks) pose_final_block(P, A, j);
    return;
  }
  const int b = blockIdx.y / (P.S - 1), k = blockIdx.y % (P.S - 1) + 1;
  const int hk = P.h[k], wk = P.w[k], n = hk * wk;
  const int G = 1 << (k - 1), L = G * G;            // lanes per texel
  const int gid = blockIdx.x * kThreads + threadIdx.x;
  const int idx = gid >> (2 * (k - 1)), sub = gid & (L - 1);
  const bool active = idx < n;
  const int i = active ? idx / wk : 0, j = active ? idx - (idx / wk) * wk : 0;
  const float* dD = ((k == 1) ? dD1 : (k == 2 ? dD2 : dD3)) + (long long)b * P.HW;
  const float ry = P.ry[k], rx = P.rx[k];
  float acc = 0.f;
  const int f = 1 << k;
  if ((hk << k) == P.H && (wk << k) == P.W) {
    // Exact 2^k ratio (the usual case): pixel v contributes to texel i with the triangle weight
    // 1 - |(v + 0.5) / f - 0.5 - i| -- every operation of the pinned up-sample is exact in fp32 here, so this is bit-for-bit
    // the forward's weight -- except inside the clamped half-texel borders, which put their whole weight on the border
    // texel.  The 2f x 2f footprint splits into 4 x 4 sub-blocks, one per lane: 16 loads and FMAs, no axis evaluation.
    if (active) {
      const int half = f >> 1, sr = sub >> (k - 1), sc = sub & (G - 1);
      const float inv_f = 1.0f / (float)f;
      const int vb = f * i - half + 4 * sr, ub = f * j - half + 4 * sc;
      float wx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int u = ub + c;
        float w = 1.0f - fabsf(((float)(4 * sc + c) + 0.5f) * inv_f - 1.0f);
        if ((j == 0 && u < half) || (j == wk - 1 && u >= f * (wk - 1) + half)) w = 1.0f;
        wx[c] = (u >= 0 && u < P.W) ? w : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int v = vb + r;
        float w = 1.0f - fabsf(((float)(4 * sr + r) + 0.5f) * inv_f - 1.0f);
        if ((i == 0 && v < half) || (i == hk - 1 && v >= f * (hk - 1) + half)) w = 1.0f;
        if (v >= 0 && v < P.H) {
          const float* rowp = dD + v * P.W;
          float row = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            CV_CHECK(v * P.W + imin(imax(ub + c, 0), P.W - 1) < P.HW);
            row = fmaf(wx[c], __ldg(rowp + imin(imax(ub + c, 0), P.W - 1)), row);
          }
          acc = fmaf(w, row, acc);
        }
      }
    }
  } else if (active) {
    // full-res rows v with source coordinate in (i-1, i+1): conservative bounds, exact test inside
    int v_lo = imax(0, (int)floorf(((float)i - 0.5f) / ry - 0.5f) - 1);
    int v_hi = imin(P.H - 1, (int)ceilf(((float)i + 1.5f) / ry - 0.5f) + 1);
    int u_lo = imax(0, (int)floorf(((float)j - 0.5f) / rx - 0.5f) - 1);
    int u_hi = imin(P.W - 1, (int)ceilf(((float)j + 1.5f) / rx - 0.5f) + 1);
    if (i == hk - 1) v_hi = P.H - 1;   // clamped border rows / columns all land on the last texel
    if (j == wk - 1) u_hi = P.W - 1;
    const int sr = sub / G, sc = sub - sr * G;
    // the column weights do not depend on the row: compute them once per block of 8 columns of this lane
    for (int ub = u_lo + sc; ub <= u_hi; ub += 8 * G) {
      float wxs[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int u = ub + c * G;
        float wx = 0.f;
        if (u <= u_hi) {
          Axis ax = upsample_axis(u, rx, wk);
          wx = ((ax.i0 == j) ? (1.0f - ax.w1) : 0.f) + ((ax.i1 == j) ? ax.w1 : 0.f);
        }
        wxs[c] = wx;
      }
      for (int v = v_lo + sr; v <= v_hi; v += G) {
        Axis ay = upsample_axis(v, ry, hk);
        const float wy = ((ay.i0 == i) ? (1.0f - ay.w1) : 0.f) + ((ay.i1 == i) ? ay.w1 : 0.f);
        if (wy == 0.f) continue;
        const float* rowp = dD + v * P.W + ub;
        float row = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (wxs[c] != 0.f) {
            CV_CHECK(v >= 0 && v < P.H && ub + c * G >= 0 && ub + c * G < P.W);
            row = fmaf(wxs[c], __ldg(rowp + c * G), row);
          }
        acc = fmaf(wy, row, acc);
      }
    }
  }
  for (int o = L >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (active && sub == 0) {
    const double* sc2 = saved_scale + (long long)(b * P.S + k) * kSavedPerScale;
    const double me = sc2[0] + (double)P.eps_disp;
    const float inv = (float)(1.0 / me), corr = (float)(sc2[1] / ((double)n * me * me));
    const float* sf = ((k == 1) ? sf1 : (k == 2 ? sf2 : sf3)) + (long long)b * n;
    const float D = __ldg(P.depth[k] + (long long)b * P.depth_bs[k] + idx);
    float* out = ((k == 1) ? g1 : (k == 2 ? g2 : g3)) + (long long)b * P.depth_bs[k];
    out[idx] = acc + __ldg(grad_loss) * smooth_grad(__ldg(sf + idx), D, inv, corr);
  }
}

// ------------------------------------------------------------------------------------------
static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Zero-fill of the scatter targets as a kernel (not a memset node), so that k_photo_bwd can be launched
// programmatically behind it and run its prologue (constants, barriers, first coefficient tile) meanwhile.
__global__ void __launch_bounds__(kThreads) k_zero(float4* __restrict__ a, long long na, float4* __restrict__ b, long long nb) {
  pdl_trigger();
  const long long stride = (long long)gridDim.x * kThreads;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < na; i += stride) a[i] = z;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nb; i += stride) b[i] = z;
}

template <int NS>
static size_t photo_bwd_smem(bool merge) {
  return sizeof(float4) * 2 * kCoefBuf + sizeof(double) * (kBwdThreads / 32) * NS * 12 +
         1024 /* per-frame constants, poses */ + 64 /* mbarriers */ +
         (merge ? sizeof(float4) * 2 * kBwdThreads : 0) /* scatter exchange slots */;
}

cudaError_t launch_backward(const KP& P, const BwdBuffers& Wk, const float* grad_loss, const uint8_t* sel,
                            const SavedView& sv, float* const* grad_depth, float* grad_T, float* grad_srcs,
                            float* grad_src_depth, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  const bool zero_sd = grad_src_depth != nullptr && (P.HW % 4 == 0) && (((uintptr_t)grad_src_depth & 15u) == 0);
  if (grad_src_depth && !zero_sd) {      // odd sizes / alignment: plain memset
    e = cudaMemsetAsync(grad_src_depth, 0, sizeof(float) * (size_t)P.B * P.N * P.HW, st);
    if (e != cudaSuccess) return e;
  }
  const bool zeroed = grad_srcs || zero_sd;
  if (zeroed) {
    const long long na = grad_srcs ? (long long)P.B * P.N * P.HW : 0;
    const long long nb = zero_sd ? (long long)P.B * P.N * P.HW / 4 : 0;
    k_zero<<<148 * 4, kThreads, 0, st>>>(Wk.gsrc4, na, reinterpret_cast<float4*>(grad_src_depth), nb);
  }
  dim3 grid(P.tiles_x, P.btiles_y, P.B);
  // the saved coefficient field [B,S,3,H,W] of 16-byte texels as a rank-3 fp32 tensor (4 W, H, 3 S B) for the TMA box loads
  CUtensorMap coef_map;
  memset(&coef_map, 0, sizeof(coef_map));
  if (COLVO_BWD_TMA) {
    const unsigned long long dims[3] = {4ull * P.W, (unsigned long long)P.H, 3ull * P.S * P.B};
    const unsigned long long strides[2] = {16ull * P.W, 16ull * P.W * P.H};
    const unsigned box[3] = {4u * kCW, (unsigned)kCH, 3u};
    e = make_tensor_map_3d(&coef_map, sv.coef, dims, strides, box);
    if (e != cudaSuccess) return e;
  }
  {
    ScopedKernelTimer tm(2, st);
    auto launch = [&](auto kern, size_t smem) {
      e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem);
      if (e != cudaSuccess) return;
      if (zeroed)
        e = launch_pdl(kern, grid, dim3(kBwdThreads), smem, st, P, grad_loss, sel, (const double*)sv.frame, (const double*)sv.scale,
                       (const float*)sv.s_field[0], (const float*)sv.coef, (const float4*)sv.geo, grad_depth[0], Wk.dDhat[1],
                       Wk.dDhat[2], Wk.dDhat[3], grad_srcs ? Wk.gsrc4 : nullptr, grad_src_depth, Wk.pose_part, coef_map);
      else
        kern<<<grid, kBwdThreads, smem, st>>>(P, grad_loss, sel, sv.frame, sv.scale, sv.s_field[0], sv.coef, sv.geo, grad_depth[0],
                                           Wk.dDhat[1], Wk.dDhat[2], Wk.dDhat[3], grad_srcs ? Wk.gsrc4 : nullptr, grad_src_depth,
                                           Wk.pose_part, coef_map);
      if (e == cudaSuccess) e = cudaGetLastError();
    };
    const bool geo = P.src_depth != nullptr, pk = (P.flags & 16u) != 0;
    // MERGE exists for planar sources only (packed bf16 images carry no source gradient, hence no scatter)
    const bool merge = (P.flags & 64u) != 0 && grad_srcs != nullptr && !pk;
    auto pick = [&](auto ns, auto geoc) {
      constexpr int NSc = decltype(ns)::value;
      constexpr bool Gc = decltype(geoc)::value;
      if (pk) launch(k_photo_bwd<NSc, Gc, true, false>, photo_bwd_smem<NSc>(false));
      else if (merge) launch(k_photo_bwd<NSc, Gc, false, true>, photo_bwd_smem<NSc>(true));
      else launch(k_photo_bwd<NSc, Gc, false, false>, photo_bwd_smem<NSc>(false));
    };
    using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
    if (P.N == 1) { if (geo) pick(I1{}, std::true_type{}); else pick(I1{}, std::false_type{}); }
    else { if (geo) pick(I2{}, std::true_type{}); else pick(I2{}, std::false_type{}); }
  }
  if (e != cudaSuccess) return e;
  // epilogue: pose reduction, source-gradient unpack and the up-sample adjoint are independent of each other, so they
  // share one launch (programmatic, so its CTAs are resident by the time k_photo_bwd drains)
  PoseFinalArgs A;
  A.pose_part = Wk.pose_part;
  A.grad_T = grad_T;
  A.gsrc4 = Wk.gsrc4;
  A.grad_srcs = grad_srcs;
  A.unpack_chunks = grad_srcs ? div_up(P.HW, kUnpackPix) : 0;
  A.n_blocks = P.B * P.N * (1 + A.unpack_chunks);
  if (P.S > 1) {
    int lanes = 0;
    for (int k = 1; k < P.S; ++k) lanes = imax(lanes, P.h[k] * P.w[k] << (2 * (k - 1)));
    const int gx = div_up(lanes, kThreads);
    e = launch_pdl(k_depth_gather, dim3(gx, P.B * (P.S - 1) + div_up(A.n_blocks, gx)), dim3(kThreads), 0, st, P, grad_loss,
                   (const double*)sv.scale, (const float*)Wk.dDhat[1], (const float*)Wk.dDhat[2], (const float*)Wk.dDhat[3],
                   (const float*)sv.s_field[1], (const float*)sv.s_field[2], (const float*)sv.s_field[3], grad_depth[1],
                   P.S > 2 ? grad_depth[2] : nullptr, P.S > 3 ? grad_depth[3] : nullptr, A);
  } else {
    e = launch_pdl(k_pose_final, dim3(A.n_blocks), dim3(kThreads), 0, st, P, A);
  }
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace colvo
