// Per-pixel arithmetic of the ColVO photometric-loss path, shared by every kernel.
//
// Everything here is a small inline function usable from device code and (for the test
// harness under tests/cpu_harness/, which checks these formulas against the oracle
// before any GPU time is spent) from plain host C++.  Nothing here is a CPU fallback:
// the library in this directory only ever launches the CUDA kernels.
//
// Follows oracle/photometric.py function by function (upstream has no source:
// /root/reference/README.md:7 is the only description of this path).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CV_HD __host__ __device__ __forceinline__
#else
#define CV_HD inline
#endif

namespace colvo {

// ---- single-rounded fp32 ops (the "pinned" chain of oracle.reproject / upsample_depth) ----
#if defined(__CUDA_ARCH__)
CV_HD float p_add(float a, float b) { return __fadd_rn(a, b); }
CV_HD float p_sub(float a, float b) { return __fsub_rn(a, b); }
CV_HD float p_mul(float a, float b) { return __fmul_rn(a, b); }
CV_HD float p_div(float a, float b) { return __fdiv_rn(a, b); }
CV_HD float p_rcp(float a) { return __frcp_rn(a); }
CV_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
// approximate reciprocal, one MUFU.RCP (rel. error 2^-23; __fdividef(1, a) wraps it in four range-fix-up
// instructions that the SSIM denominators, >= 9e-8, never need): SSIM and 1/D only, never the pinned chain
CV_HD float f_rcp(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
#else
// host build: compiled with -ffp-contract=off, so each operator rounds once
CV_HD float p_add(float a, float b) { volatile float r = a + b; return r; }
CV_HD float p_sub(float a, float b) { volatile float r = a - b; return r; }
CV_HD float p_mul(float a, float b) { volatile float r = a * b; return r; }
CV_HD float p_div(float a, float b) { volatile float r = a / b; return r; }
CV_HD float p_rcp(float a) { volatile float r = 1.0f / a; return r; }
CV_HD float f_fma(float a, float b, float c) { return a * b + c; }
CV_HD float f_rcp(float a) { return 1.0f / a; }
#endif

CV_HD int imin(int a, int b) { return a < b ? a : b; }
CV_HD int imax(int a, int b) { return a > b ? a : b; }

// reflect-pad-1 index, then clamped so that far-out halo positions stay addressable
CV_HD int reflect_clamp(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return imin(imax(i, 0), n - 1);
}

struct Cam { float fx, fy, cx, cy; };
struct Pose { float r[9]; float t[3]; };  // row-major R (rows 0..2 of T[:, :3]) and t = T[:3, 3]

// ---- row 0: one axis of the bilinear depth up-sample (align_corners = False) ----
//   s = max((i + 0.5) * ratio - 0.5, 0); i0 = floor(s); w1 = s - i0; i1 = min(i0 + 1, n_in - 1)
struct Axis { int i0, i1; float w1; };
CV_HD Axis upsample_axis(int i, float ratio, int n_in) {
  float s = p_sub(p_mul(p_add((float)i, 0.5f), ratio), 0.5f);
  s = s < 0.f ? 0.f : s;
  float f = floorf(s);
  Axis a;
  a.w1 = p_sub(s, f);
  a.i0 = imin((int)f, n_in - 1);
  a.i1 = imin(a.i0 + 1, n_in - 1);
  return a;
}
CV_HD float upsample_blend(float d00, float d01, float d10, float d11, float wx, float wy) {
  float omx = p_sub(1.0f, wx), omy = p_sub(1.0f, wy);
  float top = p_add(p_mul(omx, d00), p_mul(wx, d01));
  float bot = p_add(p_mul(omx, d10), p_mul(wx, d11));
  return p_add(p_mul(omy, top), p_mul(wy, bot));
}

// ---- rows 1-3: back-project, transform, project, validity (pinned order) ----
struct Geo {
  float u, v;        // source pixel coordinates
  float X, Y, Z;     // back-projected point in the target camera (Z = depth)
  float rx, ry;      // ray
  float iz;          // 1 / (Z' + eps)
  float Zp;          // depth of the transformed point in the source camera
  bool valid;
};
// the ray of a pixel depends on the pixel and K only: kernels hoist it out of their (k, n) loops
CV_HD float ray_x(int px, const Cam& c) { return p_div(p_sub((float)px, c.cx), c.fx); }
CV_HD float ray_y(int py, const Cam& c) { return p_div(p_sub((float)py, c.cy), c.fy); }
CV_HD Geo reproject_ray(float rx, float ry, float D, const Cam& c, const Pose& p, int W, int H, float eps, float z_min) {
  Geo g;
  g.rx = rx;
  g.ry = ry;
  g.X = p_mul(g.rx, D);
  g.Y = p_mul(g.ry, D);
  g.Z = D;
  float Xp = p_add(p_add(p_add(p_mul(p.r[0], g.X), p_mul(p.r[1], g.Y)), p_mul(p.r[2], g.Z)), p.t[0]);
  float Yp = p_add(p_add(p_add(p_mul(p.r[3], g.X), p_mul(p.r[4], g.Y)), p_mul(p.r[5], g.Z)), p.t[1]);
  float Zp = p_add(p_add(p_add(p_mul(p.r[6], g.X), p_mul(p.r[7], g.Y)), p_mul(p.r[8], g.Z)), p.t[2]);
  float x = p_add(p_mul(c.fx, Xp), p_mul(c.cx, Zp));
  float y = p_add(p_mul(c.fy, Yp), p_mul(c.cy, Zp));
  g.Zp = Zp;
  g.iz = p_rcp(p_add(Zp, eps));
  g.u = p_mul(x, g.iz);
  g.v = p_mul(y, g.iz);
  g.valid = (g.u >= 0.f) && (g.u <= (float)(W - 1)) && (g.v >= 0.f) && (g.v <= (float)(H - 1)) && (Zp > z_min);
  return g;
}
CV_HD Geo reproject(int px, int py, float D, const Cam& c, const Pose& p, int W, int H, float eps, float z_min) {
  return reproject_ray(ray_x(px, c), ray_y(py, c), D, c, p, W, H, eps, z_min);
}

// ---- row 4: bilinear taps with border padding ----
struct Taps {
  int x0, x1, y0, y1;
  float wx, wy;
  bool gx, gy;       // coordinate gradient passes (strictly inside the border)
};
CV_HD Taps make_taps(float u, float v, int W, int H) {
  Taps t;
  if (!(u == u)) u = 0.f;
  if (!(v == v)) v = 0.f;
  float wm = (float)(W - 1), hm = (float)(H - 1);
  t.gx = (u > 0.f) && (u < wm);
  t.gy = (v > 0.f) && (v < hm);
  float uc = fminf(fmaxf(u, 0.f), wm);
  float vc = fminf(fmaxf(v, 0.f), hm);
  float xf = floorf(uc), yf = floorf(vc);
  t.wx = uc - xf;
  t.wy = vc - yf;
  t.x0 = (int)xf;          // uc is already clamped to [0, W-1] (NaN -> 0, +-inf -> the border): no integer clamp needed
  t.y0 = (int)yf;
  t.x1 = imin(t.x0 + 1, W - 1);
  t.y1 = imin(t.y0 + 1, H - 1);
  return t;
}
CV_HD float bilerp(float i00, float i01, float i10, float i11, float wx, float wy) {
  float omx = 1.0f - wx;
  float top = f_fma(wx, i01, omx * i00);
  float bot = f_fma(wx, i11, omx * i10);
  return f_fma(wy, bot, (1.0f - wy) * top);
}

// ---- rows 6-7: SSIM from 3x3 window moments of the RAW warped image, calibrated on the fly ----
// raw moments: mu = E[x], exx = E[x^2], exy = E[x y]; target: muy, sy = E[y^2] - muy^2.
// calibrated image a*x + b has  mu~ = a mu + b,  s~ = a^2 s,  s~xy = a sxy.
struct SsimParts {
  float S;           // SSIM value
  float t;           // (1 - S) / 2 before the clamp
  float dmu, dsx, dsxy;  // dS/d(mu~), dS/d(s~), dS/d(s~xy)
};
CV_HD SsimParts ssim_parts(float mut, float st, float stxy, float muy, float sy, float c1, float c2) {
  float A1 = 2.f * mut * muy + c1;
  float A2 = 2.f * stxy + c2;
  float B1 = mut * mut + muy * muy + c1;
  float B2 = st + sy + c2;
  float iB1 = f_rcp(B1), iB2 = f_rcp(B2);
  SsimParts o;
  float r2 = A2 * iB2;
  o.S = A1 * iB1 * r2;
  o.t = 0.5f * (1.0f - o.S);
  o.dmu = 2.f * iB1 * (muy * r2 - o.S * mut);
  o.dsx = -o.S * iB2;
  o.dsxy = 2.f * A1 * iB1 * iB2;
  return o;
}
CV_HD float clamp01(float t) { return fminf(fmaxf(t, 0.f), 1.f); }
CV_HD float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
// w * sgn(x) with sgn(0) = 0 (torch's sub-gradient of |.|): copysign + one select instead of two compares and a convert
CV_HD float sgn_scaled(float w, float x) { return (x != 0.f) ? copysignf(w, x) : 0.f; }     // w >= 0 only
// w * sgn(x) for a weight of either sign (the LCC gain, hence the L1 weight of the backward, can be negative):
// the sign bit of x is xor-ed onto w
CV_HD float sgn_mul(float w, float x) {
#if defined(__CUDA_ARCH__)
  return (x != 0.f) ? __uint_as_float(__float_as_uint(w) ^ (__float_as_uint(x) & 0x80000000u)) : 0.f;
#else
  return (x > 0.f) ? w : ((x < 0.f) ? -w : 0.f);
#endif
}

// coefficient fields of the SSIM adjoint in gather form (re-derived for
// raw moments):  d pe_p / d x_q  =  ca_p + x_q * cb_p + y_q * cg_p   for every occurrence of q in
// the reflect-padded 3x3 window of p (already includes alpha/3 * (-1/2) * 1/9 * active).
struct Coef { float ca, cb, cg; };
CV_HD Coef coef_from_parts(const SsimParts& q, float mu, float muy, float a, float alpha, float wp) {
  Coef o;
  float act = (q.t >= 0.f && q.t <= 1.f) ? wp * (-0.5f * alpha) * (1.0f / 27.0f) : 0.f;
  float dmu = a * q.dmu, ds = a * a * q.dsx, dsxy = a * q.dsxy;
  o.ca = act * (dmu - 2.f * mu * ds - muy * dsxy);
  o.cb = act * 2.f * ds;
  o.cg = act * dsxy;
  return o;
}

// photometric error of one pixel from per-channel raw window moments and the centre sample
//   pe = alpha * mean_c clamp((1 - SSIM_c)/2) + (1 - alpha) * mean_c |a x_c + b - y_c|
// Optionally also d pe / d a and d pe / d b (the LCC adjoint sums G_a, G_b of the forward) and the
// unit-weight adjoint coefficients of this window (saved by the forward for the backward).
CV_HD float pe_channel(float mu, float exx, float exy, float muy, float sy, float xc, float yc, float a, float b,
                       float alpha, float c1, float c2, float* dpa, float* dpb, bool want_cf, Coef& cf) {
  float s = exx - mu * mu;
  float sxy = exy - mu * muy;
  float mut = f_fma(a, mu, b);
  float st = a * a * s;
  float stxy = a * sxy;
  SsimParts q = ssim_parts(mut, st, stxy, muy, sy, c1, c2);
  float diff = f_fma(a, xc, b) - yc;
  float pe = alpha * clamp01(q.t) + (1.f - alpha) * fabsf(diff);
  if (dpa) {
    float act = (q.t >= 0.f && q.t <= 1.f) ? -0.5f * alpha : 0.f;
    float sg = (1.f - alpha) * sgn(diff);
    *dpa += act * (q.dmu * mu + q.dsx * 2.f * a * s + q.dsxy * sxy) + sg * xc;
    *dpb += act * q.dmu + sg;
  }
  if (want_cf) cf = coef_from_parts(q, mu, muy, a, alpha, 1.0f);
  return pe;
}
CV_HD float pe_channel(float mu, float exx, float exy, float muy, float sy, float xc, float yc, float a, float b,
                       float alpha, float c1, float c2, float* dpa, float* dpb) {
  Coef unused;
  return pe_channel(mu, exx, exy, muy, sy, xc, yc, a, b, alpha, c1, c2, dpa, dpb, false, unused);
}

CV_HD Coef ssim_coef(float mu, float exx, float exy, float muy, float sy, float a, float b, float alpha, float c1, float c2,
                     float wp) {
  float s = exx - mu * mu;
  float sxy = exy - mu * muy;
  float mut = f_fma(a, mu, b);
  SsimParts q = ssim_parts(mut, a * a * s, a * sxy, muy, sy, c1, c2);
  return coef_from_parts(q, mu, muy, a, alpha, wp);
}

// multiplicity of window centre p (= q + d, d in {-1,0,1}) in the gather at pixel q along one axis:
// 0 if p is outside the image, 2 if the reflect padding of p's window lands on q a second time.
CV_HD float reflect_mult(int q, int p, int n) {
  if (p < 0 || p >= n) return 0.f;
  float m = 1.f;
  if (q == 1 && p == 0) m += 1.f;
  if (q == n - 2 && p == n - 1) m += 1.f;
  return m;
}

// ---- geometric consistency  diff = clamp(|Z' - Ds| / (Z' + Ds), 0, 1)  (oracle A16) ----
// returns diff and its partial derivatives (torch semantics: abs' = sign with sign(0) = 0, clamp passes
// the gradient on the closed interval)
CV_HD float geo_diff(float Zp, float ds, float& dZ, float& dS) {
  const float num = Zp - ds, den = Zp + ds;
  const float iden = 1.0f / den;
  const float r = fabsf(num) * iden;
  const bool pass = (r >= 0.f) && (r <= 1.f);
  const float sg = sgn(num);
  dZ = pass ? (sg - r) * iden : 0.f;      // d/dZ' [ |n| / d ] = sg/d - |n|/d^2
  dS = pass ? (-sg - r) * iden : 0.f;
  return clamp01(r);
}

// ---- row 11: adjoint of projection / transform / back-projection for one pixel ----
// in: du, dv = dL/du', dL/dv'.  out: dD = dL/dDhat and dXp[3] = dL/dX' (gradient w.r.t. the
// transformed point).  The pose gradient follows from dXp alone:
//   dL/dt_i = sum dXp_i,   dL/dR_ij = sum dXp_i * X_j   with X = D * (rx, ry, 1)
// so a kernel only has to accumulate dXp_i and dXp_i * D per pixel (the ray is constant per pixel).
// dZp_direct: gradient that reaches Z' directly (the geometric-consistency term), 0 otherwise.
CV_HD float project_adjoint(const Geo& g, const Cam& c, const Pose& p, float du, float dv, float dZp_direct,
                            float (&dXp)[3]) {
  float dx = du * g.iz, dy = dv * g.iz;
  dXp[0] = c.fx * dx;
  dXp[1] = c.fy * dy;
  dXp[2] = c.cx * dx + c.cy * dy - g.iz * (du * g.u + dv * g.v) + dZp_direct;
  float dX = p.r[0] * dXp[0] + p.r[3] * dXp[1] + p.r[6] * dXp[2];
  float dY = p.r[1] * dXp[0] + p.r[4] * dXp[1] + p.r[7] * dXp[2];
  float dZ = p.r[2] * dXp[0] + p.r[5] * dXp[1] + p.r[8] * dXp[2];
  return g.rx * dX + g.ry * dY + dZ;
}
// expand the six per-pixel sums (w_i = sum dXp_i * D, t_i = sum dXp_i) into the 12 pose-gradient entries
// gp[3*i + j] = dL/dR_ij, gp[9 + i] = dL/dt_i
CV_HD void pose_grad_expand(const float (&w)[3], const float (&t)[3], float rx, float ry, float* gp) {
  for (int i = 0; i < 3; ++i) {
    gp[3 * i + 0] = w[i] * rx;
    gp[3 * i + 1] = w[i] * ry;
    gp[3 * i + 2] = w[i];
    gp[9 + i] = t[i];
  }
}

}  // namespace colvo
