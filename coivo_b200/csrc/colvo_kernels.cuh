// Kernel parameter block, launch geometry and device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#include "colvo_math.cuh"
#include "colvo_f2.cuh"

// Debug build with index checks (-DCOLVO_DEBUG_BOUNDS=1; scripts/build_variants.py bounds=COLVO_DEBUG_BOUNDS=1): every
// computed gather / scatter / shared-memory index is asserted before use and a violation traps the kernel with a message.
// compute-sanitizer is closed on the GPU pool these kernels are developed on; the GPU parity tests are run once per
// kernel change against this build instead (COLVO_LIB=build/variants/lib_bounds.so pytest -m gpu).
#if defined(COLVO_DEBUG_BOUNDS) && COLVO_DEBUG_BOUNDS
#include <stdio.h>
#define CV_CHECK(cond)                                                                                              \
  do {                                                                                                              \
    if (!(cond)) {                                                                                                  \
      printf("colvo bounds check failed: %s  (%s:%d, block %d,%d,%d thread %d)\n", #cond, __FILE__, __LINE__,        \
             (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x);                                  \
      __trap();                                                                                                     \
    }                                                                                                               \
  } while (0)
#else
#define CV_CHECK(cond) ((void)0)
#endif
#define CV_CHECK_TAPS(t, W, H) CV_CHECK((t).x0 >= 0 && (t).x0 <= (t).x1 && (t).x1 < (W) && (t).y0 >= 0 && (t).y0 <= (t).y1 && (t).y1 < (H))

namespace colvo {

constexpr int kMaxS = 4;
constexpr int kMaxN = 2;
constexpr int kThreads = 256;
constexpr int kTileW = 32;       // one warp = one tile row: coalesced 128 B rows
constexpr int kTileH = 8;
// backward tile kernel (k_photo_bwd): 32 x kBwdTileH pixels, one thread each.  32 x 4 tiles at 5 CTAs (20 warps, 96
// registers) per SM with the coefficient tile fetched by TMA measured best on B200 (profiles/r2_bwd_tma_tiles.log)
#ifndef COLVO_BWD_TILE_H
#define COLVO_BWD_TILE_H 4
#endif
constexpr int kBwdTileH = COLVO_BWD_TILE_H;
constexpr int kBwdThreads = 32 * kBwdTileH;
// forward tile kernel (k_photo_fwd): every warp walks down a strip of windows, one window column per lane
constexpr int kFwdWarps = 4;     // warps per CTA
#ifndef COLVO_FWD_ROWS       // 3 rows: 50 KB of shared memory and 128 registers -> 4 CTAs (16 warps) per SM; 4 rows: 66 KB / 167
#define COLVO_FWD_ROWS 3     // registers -> 3 CTAs; measured 128 vs 133 us (profiles/r1_variants_v21.log)
#endif
constexpr int kFwdRows = COLVO_FWD_ROWS;   // window rows per warp
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kFwdTileH = kFwdWarps * kFwdRows;   // 12
#ifndef COLVO_STAT_PPT
#define COLVO_STAT_PPT 8
#endif
constexpr int kStatPPT = COLVO_STAT_PPT;   // pixels per thread in the LCC statistics pass (one scale per CTA: the sweep)
// COLVO_STATS_KINNER = 1 (training loss, S > 1): one CTA walks all scales of its pixel chunk (see k_warp_stats), 128 threads
// x 4 pixels per scale.  Measured: the kernel alone 95.8 vs 97.1 us, the whole step 0.4407 vs 0.4342 ms (the smaller CTAs
// overlap worse with the launches packed around them) -- off (profiles/r2_stats_kinner.log).
#ifndef COLVO_STATS_KINNER
#define COLVO_STATS_KINNER 0
#endif
constexpr int kStatThreadsK = 128, kStatPPTK = 4;
// one-scale-per-CTA form: 4 pixels per thread for the training loss (S > 1: -0.6 % per step against 8), kStatPPT = 8 for the
// single-scale consistency sweep (4 is 3 % slower there) -- profiles/r2_step_level_ab.log
#ifndef COLVO_STAT_PPT_TRAIN
#define COLVO_STAT_PPT_TRAIN 4
#endif
constexpr int kStatPPTTrain = COLVO_STAT_PPT_TRAIN;
inline bool stats_k_inner(int S) { return COLVO_STATS_KINNER && S > 1; }
inline int stats_pixels_per_cta(int S) {
  return stats_k_inner(S) ? kStatThreadsK * kStatPPTK : kThreads * (S > 1 ? kStatPPTTrain : kStatPPT);
}
constexpr int kStatVals = 6;     // per (frame, chunk): n, Sx, Sy, Sxx, Sxy, sum of geometric-consistency diffs
#ifndef COLVO_SM_BW
#define COLVO_SM_BW 64
#endif
#ifndef COLVO_SM_BH
#define COLVO_SM_BH 16
#endif
constexpr int kSmBW = COLVO_SM_BW, kSmBH = COLVO_SM_BH;   // full-resolution pixels per smoothness block (k_smooth); powers of two >= 8
constexpr int kSmVals = 4;            // per block and scale: sum over x edges, sum over y edges, sum s*d, sum d

// saved[] layout: doubles  [B*N*S][kSavedPerFrame]  n, mean_x, mean_y, 1/(n (var+eps)), a, b, G_a, G_b
//                 doubles  [B*S][kSavedPerScale]    mean inverse depth, sum_p s_p d_p
//                 floats   s-field of every scale   dL_smooth/dd*_p for grad_loss = 1, [B,h_k,w_k]
//                 float4   [B,S,3,H,W]              SSIM adjoint coefficients (ca, cb, cg, n) per channel of the
//                                                   winning re-projection candidate n (all zero where identity won);
//                                                   16-byte texels so the backward stages them with 16 B cp.async
//                 float4   [B,N,S,H,W]              projection of every pixel (u', v', 1/(Z'+eps), D^ with the valid
//                                                   bit in its mantissa LSB): the backward does not re-project
constexpr int kSavedPerFrame = 8;
constexpr int kSavedPerScale = 2;

struct KP {
  int B, N, S, H, W, HW;
  int h[kMaxS], w[kMaxS];
  float ry[kMaxS], rx[kMaxS];        // h_k / H, w_k / W rounded once to fp32 (oracle._upsample_axis)
  int sm_blocks;                     // k_smooth CTAs per image
  float sm_cx[kMaxS], sm_cy[kMaxS];  // smoothness edge weights lambda_k / (S * #x-edges), lambda_k / (S * #y-edges)
  float alpha, c1, c2, eps_proj, eps_lcc, eps_disp, z_min, smooth_weight;
  unsigned flags;
  const void* tgt;                   // [B,3,H,W] fp32 planar, or [B,H,W,4] bf16 packed (COLVO_F_PACKED_BF16)
  const void* srcs;                  // [B,N,3,H,W] fp32 planar, or [B,N,H,W,4] bf16 packed
  const float* depth[kMaxS];         // [B,1,h_k,w_k]
  const float* K;                    // [B,3,3]
  const float* T;                    // [B,N,4,4]
  const float* src_depth;            // [B,N,1,H,W] or null: geometric-consistency term (f-2)
  float geo_weight;
  int tgt_bf, src_bf, src_nf;        // strides in FRAMES (one frame = one [3,H,W] image): target b -> b*tgt_bf,
                                     // source (b,n) -> b*src_bf + n*src_nf (the consistency sweep aliases one array)
  long long frame_el;                // elements per frame in the storage format: 3*HW floats, or HW 8-byte texels
  long long depth_bs[kMaxS];
  int K_bs, T_bs, T_ns;
  int tiles_x, tiles_y;               // 32 x 8 tiles (consistency sweep; sizing of the forward partials)
  int btiles_y;                       // rows of 32 x kBwdTileH tiles of the backward kernel
  int ftiles_x, ftiles_y;             // 32 x kFwdTileH tiles of the forward kernel
};

__device__ __forceinline__ Cam load_cam(const KP& P, int b) {
  const float* k = P.K + (long long)b * P.K_bs;
  Cam c;
  c.fx = __ldg(k + 0);
  c.fy = __ldg(k + 4);
  c.cx = __ldg(k + 2);
  c.cy = __ldg(k + 5);
  return c;
}
__device__ __forceinline__ Pose load_pose(const KP& P, int b, int n) {
  const float* t = P.T + (long long)b * P.T_bs + (long long)n * P.T_ns;
  Pose p;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    p.r[3 * i + 0] = __ldg(t + 4 * i + 0);
    p.r[3 * i + 1] = __ldg(t + 4 * i + 1);
    p.r[3 * i + 2] = __ldg(t + 4 * i + 2);
    p.t[i] = __ldg(t + 4 * i + 3);
  }
  return p;
}

// The N source poses of a triplet, lane n = source n (one register pair per entry for N = 2)
template <int NS>
struct PoseV { Vn<NS> r[9], t[3]; };
template <int NS>
__device__ __forceinline__ PoseV<NS> load_pose_v(const KP& P, int b) {
  PoseV<NS> p;
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const Pose q = load_pose(P, b, n);
#pragma unroll
    for (int i = 0; i < 9; ++i) p.r[i].set(n, q.r[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) p.t[i].set(n, q.t[i]);
  }
  return p;
}
// rows 1-3 for all sources at once: the same pinned chain as colvo_math.cuh::reproject_ray, lane by lane
template <int NS>
struct GeoV { Vn<NS> u, v, iz, Zp; bool valid[NS]; };
template <int NS>
__device__ __forceinline__ GeoV<NS> reproject_v(float rx, float ry, float D, const Cam& c, const PoseV<NS>& p, int W, int H,
                                                float eps, float z_min) {
  typedef Vn<NS> V;
  const V X = V(p_mul(rx, D)), Y = V(p_mul(ry, D)), Z = V(D);
  const V Xp = padd(padd(padd(pmul(p.r[0], X), pmul(p.r[1], Y)), pmul(p.r[2], Z)), p.t[0]);
  const V Yp = padd(padd(padd(pmul(p.r[3], X), pmul(p.r[4], Y)), pmul(p.r[5], Z)), p.t[1]);
  const V Zp = padd(padd(padd(pmul(p.r[6], X), pmul(p.r[7], Y)), pmul(p.r[8], Z)), p.t[2]);
  const V x = padd(pmul(V(c.fx), Xp), pmul(V(c.cx), Zp));
  const V y = padd(pmul(V(c.fy), Yp), pmul(V(c.cy), Zp));
  GeoV<NS> g;
  g.Zp = Zp;
  g.iz = prcp(padd(Zp, V(eps)));
  g.u = pmul(x, g.iz);
  g.v = pmul(y, g.iz);
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const float u = g.u.lane(n), v = g.v.lane(n);
    g.valid[n] = (u >= 0.f) && (u <= (float)(W - 1)) && (v >= 0.f) && (v <= (float)(H - 1)) && (Zp.lane(n) > z_min);
  }
  return g;
}

// Row 0: up-sampled depth at a full-resolution pixel (identity at k == 0).
__device__ __forceinline__ float depth_at(const KP& P, const float* __restrict__ Dk, int k, int px, int py) {
  if (k == 0) return __ldg(Dk + py * P.W + px);
  const int wk = P.w[k];
  Axis ay = upsample_axis(py, P.ry[k], P.h[k]);
  Axis ax = upsample_axis(px, P.rx[k], wk);
  CV_CHECK(ay.i0 >= 0 && ay.i1 < P.h[k] && ax.i0 >= 0 && ax.i1 < wk);
  float d00 = __ldg(Dk + ay.i0 * wk + ax.i0);
  float d01 = __ldg(Dk + ay.i0 * wk + ax.i1);
  float d10 = __ldg(Dk + ay.i1 * wk + ax.i0);
  float d11 = __ldg(Dk + ay.i1 * wk + ax.i1);
  return upsample_blend(d00, d01, d10, d11, ax.w1, ay.w1);
}

// One [3,H,W] frame in either storage format.  Planar fp32: three planes HW apart.  Packed: one 8-byte
// RGBA-bf16 texel per pixel: a single load yields the three channels, widened
// to fp32 by a shift / mask -- all arithmetic stays fp32.
struct Texels { float i00[3], i01[3], i10[3], i11[3]; };
template <bool PK> struct Img;
template <> struct Img<false> {
  const float* p;
  int HW;
  // the four bilinear taps of all three planes: four tap pointers, then + HW per plane (one IMAD.WIDE per load)
  __device__ __forceinline__ void load_taps(int o00, int o01, int o10, int o11, Texels& tx) const {
    // The frame base is made opaque to the optimiser: otherwise it folds the 64-bit frame offset into every
    // address (4 integer instructions per load); with a materialised base each tap is one IMAD.WIDE and each
    // further plane one more.
    const float* fb = p;
    asm volatile("" : "+l"(fb));
#ifdef COLVO_EXP_NEARTAPS   // timing ablation (WRONG results): every tap lands in the same 4 KB -> always an L1 hit
    const unsigned u00 = o00 & 1023, u01 = o01 & 1023, u10 = o10 & 1023, u11 = o11 & 1023, hw = 1024;
#else
    const unsigned u00 = o00, u01 = o01, u10 = o10, u11 = o11, hw = HW;
#endif
#pragma unroll
    for (int c = 0; c < 3; ++c) {       // 32-bit offset add, then one widening multiply-add onto the base
      tx.i00[c] = __ldg(fb + (u00 + c * hw));
      tx.i01[c] = __ldg(fb + (u01 + c * hw));
      tx.i10[c] = __ldg(fb + (u10 + c * hw));
      tx.i11[c] = __ldg(fb + (u11 + c * hw));
    }
  }
  __device__ __forceinline__ void load3(int off, float (&v)[3]) const {
    const float* q = p + off;     // one IMAD.WIDE, then + HW per plane
    v[0] = __ldg(q);
    q += HW;
    v[1] = __ldg(q);
    q += HW;
    v[2] = __ldg(q);
  }
};
template <> struct Img<true> {
  const uint2* p;
  int HW;
  __device__ __forceinline__ void load_taps(int o00, int o01, int o10, int o11, Texels& tx) const {
    load3(o00, tx.i00);
    load3(o01, tx.i01);
    load3(o10, tx.i10);
    load3(o11, tx.i11);
  }
  __device__ __forceinline__ void load3(int off, float (&v)[3]) const {
    const uint2 q = __ldg(p + off);
    v[0] = __uint_as_float(q.x << 16);
    v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16);
  }
};
// frame index (small) times the per-frame element count P.frame_el
template <bool PK>
__device__ __forceinline__ Img<PK> img_at(const KP& P, const void* base, int frame);
template <>
__device__ __forceinline__ Img<false> img_at<false>(const KP& P, const void* base, int frame) {
  Img<false> im;
  im.p = static_cast<const float*>(base) + frame * P.frame_el;
  im.HW = P.HW;
  return im;
}
template <>
__device__ __forceinline__ Img<true> img_at<true>(const KP& P, const void* base, int frame) {
  Img<true> im;
  im.p = static_cast<const uint2*>(base) + frame * P.frame_el;
  im.HW = P.HW;
  return im;
}

// Rows 1-4 for one pixel whose ray (rx, ry) and up-sampled depth D are already known (both are
// shared by the N sources; the ray also by the S scales): geometry, taps, 12 texels, 3 channels.
// `src` is the [3][H][W] plane set of one source frame; offsets stay 32-bit (3*HW < 2^31).
template <bool PK>
__device__ __forceinline__ void warp_sample(const KP& P, const Img<PK>& src, const Cam& cam, const Pose& pose, float rx,
                                            float ry, float D, Geo& g, Taps& t, Texels& tx, float (&x)[3], int foff = 0) {
  g = reproject_ray(rx, ry, D, cam, pose, P.W, P.H, P.eps_proj, P.z_min);
  t = make_taps(g.u, g.v, P.W, P.H);
  const int r0 = foff + t.y0 * P.W, r1 = foff + t.y1 * P.W;     // foff: element offset of the frame from src.p
  src.load_taps(r0 + t.x0, r0 + t.x1, r1 + t.x0, r1 + t.x1, tx);
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = bilerp(tx.i00[c], tx.i01[c], tx.i10[c], tx.i11[c], t.wx, t.wy);
}

// 16-byte asynchronous global->shared copy (LDGSTS), L2-only (.cg): streaming data that should not displace the
// texels in L1; zero-fills the destination when !pred
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
// L2 eviction-priority hints for data that is streamed exactly once: it should not displace the lines the kernel
// keeps hitting (the scatter accumulator of the backward is read-modify-written ~32 times per texel by L2 atomics)
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16_hint(void* smem, const void* gmem, bool pred, unsigned long long pol) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(sa), "l"(gmem), "r"(sz), "l"(pol));
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
// 16-byte load of data that is read exactly once by this grid: no L1 allocation, so that it does not displace the source
// texels the kernel keeps gathering (COLVO_STREAM_LOADS = 0: a plain read-only load)
#ifndef COLVO_STREAM_LOADS
#define COLVO_STREAM_LOADS 1
#endif
__device__ __forceinline__ float4 ld_stream(const float4* p) {
#if COLVO_STREAM_LOADS
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ float ld_stream(const float* p) {
#if COLVO_STREAM_LOADS
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
// the same with the 32-bit shared-window address already at hand (hoisted out of an unrolled staging loop)
__device__ __forceinline__ void cp_async16_s(unsigned saddr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gmem));
}
__device__ __forceinline__ void cp_async8_s(unsigned saddr, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gmem));
}
__device__ __forceinline__ void cp_async4_s(unsigned saddr, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// bilinear sample of one extra plane (the source depth map) with the taps of a warped pixel
__device__ __forceinline__ float sample_plane(const float* __restrict__ plane, const Taps& t, int W, float (&d)[4]) {
  const int r0 = t.y0 * W, r1 = t.y1 * W;
  d[0] = __ldg(plane + (r0 + t.x0));
  d[1] = __ldg(plane + (r0 + t.x1));
  d[2] = __ldg(plane + (r1 + t.x0));
  d[3] = __ldg(plane + (r1 + t.x1));
  return bilerp(d[0], d[1], d[2], d[3], t.wx, t.wy);
}

// ---- mbarrier (shared-memory transaction barrier) helpers: producers' cp.async completions arrive on it by
// themselves, consumers poll a phase parity -- a block barrier with slack instead of a rendezvous ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
// all cp.async issued so far by this thread arrive on `bar` when they complete (counts as this thread's arrival)
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n .reg .b64 t;\n mbarrier.arrive.shared::cta.b64 t, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n .reg .pred p;\n"
      "W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra D_%=;\n bra W_%=;\n"
      "D_%=:\n}\n" ::"r"(a), "r"(parity) : "memory");
}

// ---- TMA (cp.async.bulk.tensor): one elected thread moves a whole tile global -> shared; the bytes land on an mbarrier
// (complete_tx), out-of-bounds elements of the box are zero-filled by the hardware ----
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// host side: rank-3 fp32 tensor map, no swizzle, zero fill.  dims / box in elements (innermost first), strides in bytes
// for dims 1 and 2.  Returns cudaSuccess or an error (no fallback: the caller reports it).
cudaError_t make_tensor_map_3d(CUtensorMap* tm, const void* base, const unsigned long long (&dims)[3],
                               const unsigned long long (&strides_bytes)[2], const unsigned (&box)[3]);

// fire-and-forget float add to GLOBAL memory (RED.E.ADD.F32): explicit address space, so that a base pointer hidden
// from the optimiser (see Img<false>::load_taps) does not turn into a generic atomic
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(p), "f"(v) : "memory");
}
// the same for a 16-byte texel (REDG.E.ADD.F32x4): one request for the three channels of a tap
__device__ __forceinline__ void red_add3(float4* p, float a, float b, float c) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(0.f) : "memory");
}

// store of an intermediate that a LATER kernel reads.  COLVO_STREAM_STORES = 1: st.global.cs (evict-first) -- the data is
// not re-read by the writing grid and should not displace the source texels that grid keeps gathering.
#ifndef COLVO_STREAM_STORES
#define COLVO_STREAM_STORES 1
#endif
template <typename T>
__device__ __forceinline__ void st_stream(T* p, const T& v) {
#if COLVO_STREAM_STORES
  __stcs(p, v);
#else
  *p = v;
#endif
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute
// may start while its predecessor in the stream is still draining; pdl_wait() blocks until the predecessor grid has
// completed and its writes are visible, pdl_trigger() lets the successor's CTAs be scheduled early.  Both are
// no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::); }

// Opting in to > 48 KB of dynamic shared memory is a per-function, per-device attribute: set once, then a table lookup
// (no driver call on the launch path).  Implemented in colvo_api.cu.
cudaError_t ensure_dyn_smem(const void* kernel, size_t bytes);

// launch `kern` so that it may overlap the tail of the previous kernel in `st` (see pdl_wait)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction of NV per-thread values (kThreads threads); thread i < NV stores out[i].
template <int NV, typename TV>
__device__ __forceinline__ void block_reduce_store(TV (&v)[NV], TV* sm /* [kThreads/32][NV] */, TV* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    TV s = warp_sum(v[i]);
    if (lane == 0) sm[wid * NV + i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    TV s = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += sm[w * NV + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// Deterministic block sum of NSLOT per-thread fp32 values that already sit in shared memory as
// slots[s * kThreads + tid]: (slot, warp) pairs are summed in fp64 by one thread each (lane-rotated
// reads: conflict-free), then 8 warp partials per slot are combined in a fixed order.  ~5x fewer
// issue slots than NSLOT fp64 shuffle trees.  `part` is (kThreads/32) * NSLOT doubles of scratch.
// Call with all threads after the slots are written and a __syncthreads(); result in out[s], s < NSLOT.
template <int NSLOT, int NT, typename F>
__device__ __forceinline__ void block_sum_slots(const float* slots, double* part, F&& emit) {
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x;
  for (int item = tid; item < NSLOT * NW; item += NT) {
    const int s = item / NW, w = item - s * NW;
    const float* p = slots + s * NT + w * 32;
    double acc = 0.0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) acc += (double)p[(l + tid) & 31];
    part[item] = acc;
  }
  __syncthreads();
  if (tid < NSLOT) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += part[tid * NW + w];
    emit(tid, acc);
  }
}

// ---- launchers implemented in colvo_fwd.cu / colvo_bwd.cu (called by colvo_api.cu) ----
struct FwdBuffers {
  double* stat_part;     // [B*N*S][stat_chunks][kStatVals]
  int stat_chunks;
  double* smooth_part;   // [B][sm_blocks][S][kSmVals] per k_smooth CTA and scale: sum_x, sum_y, sum s*d, sum d (un-normalised)
  double* loss_part;     // [B*tiles]
  double* g_part;        // [B*tiles][N*kMaxS*2]
  float4* iw;            // [B,N,S,H,W] raw warped frames as (x0, x1, x2, -) texels (k_warp_stats -> k_photo_fwd)
};
struct BwdBuffers {
  float* dDhat[kMaxS];   // k >= 1: [B,H,W] full-resolution depth adjoint before the up-sample adjoint
  double* pose_part;     // [B*tiles][N*12]
  float4* gsrc4;         // [B,N,H,W] texel-interleaved gradient of the sources: the scatter target (one vector RED per
                         // tap), unpacked into the planar grad_srcs by k_depth_gather's launch
};
struct SavedView {       // the caller-owned `saved` buffer, carved
  double* frame;         // [B*N*S][kSavedPerFrame]
  double* scale;         // [B*S][kSavedPerScale]
  float* s_field[kMaxS]; // [B,h_k,w_k]
  float* coef;           // [B,S,3,H,W] float4: unit-weight SSIM adjoint coefficients of the winning re-projection
  float4* geo;           // [B,N,S,H,W] float4: (u', v', iz, D^|valid) of every warped pixel
};

// one-shot event bracket around one kernel launch (colvo_debug_time_kernel)
struct KernelTimer { int which; cudaEvent_t start, stop; };
extern KernelTimer g_timer;   // process-wide: autograd runs the backward on its own thread
struct ScopedKernelTimer {
  bool on;
  cudaStream_t st;
  ScopedKernelTimer(int id, cudaStream_t s) : on(g_timer.which == id), st(s) {
    if (on) cudaEventRecord(g_timer.start, st);
  }
  ~ScopedKernelTimer() {
    if (on) { cudaEventRecord(g_timer.stop, st); g_timer.which = 0; }
  }
};

cudaError_t launch_forward(const KP& P, const FwdBuffers& W, float* loss, float* ab, uint8_t* valid, uint8_t* sel,
                           float* occ, const SavedView& saved, cudaStream_t st);
cudaError_t launch_backward(const KP& P, const BwdBuffers& W, const float* grad_loss, const uint8_t* sel,
                            const SavedView& saved, float* const* grad_depth, float* grad_T, float* grad_srcs,
                            float* grad_src_depth, cudaStream_t st);
cudaError_t launch_consistency(const KP& P, double* stat_part, int stat_chunks, double* pe_part, float* ab,
                               float* out, float4* iw, int pairs_per_pass, cudaStream_t st);

}  // namespace colvo
