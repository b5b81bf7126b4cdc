// TEST SCAFFOLDING ONLY -- never linked into libcolvo_b200.so and never on the product path.
//
// Whole-image, single-threaded driver around the per-pixel functions in
// coivo_b200/csrc/colvo_math.cuh (the same header the CUDA kernels include), so that the
// analytic adjoint formulas (SSIM gather coefficients, reflect multiplicities, LCC adjoint,
// projection adjoint, up-sample adjoint) can be checked against the oracle's autograd on the
// CPU container before any GPU time is spent.  tests/test_math_harness.py builds it with g++
// (-ffp-contract=off) and calls it through ctypes.  Smoothness is not covered here.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../coivo_b200/csrc/colvo_math.cuh"
#include "../../coivo_b200/csrc/colvo_pe.cuh"

using namespace colvo;

namespace {
struct Dims { int B, N, S, H, W; int h[4], w[4]; float ry[4], rx[4]; };

float depth_at(const Dims& d, const float* Dk, int k, int px, int py) {
  if (k == 0) return Dk[py * d.W + px];
  int wk = d.w[k];
  Axis ay = upsample_axis(py, d.ry[k], d.h[k]);
  Axis ax = upsample_axis(px, d.rx[k], wk);
  return upsample_blend(Dk[ay.i0 * wk + ax.i0], Dk[ay.i0 * wk + ax.i1], Dk[ay.i1 * wk + ax.i0], Dk[ay.i1 * wk + ax.i1],
                        ax.w1, ay.w1);
}
Cam cam_of(const float* K) { Cam c; c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5]; return c; }
Pose pose_of(const float* T) {
  Pose p;
  for (int i = 0; i < 3; ++i) { p.r[3 * i] = T[4 * i]; p.r[3 * i + 1] = T[4 * i + 1]; p.r[3 * i + 2] = T[4 * i + 2]; p.t[i] = T[4 * i + 3]; }
  return p;
}
struct Frame {   // one warped frame (b, n, k)
  std::vector<float> x;       // [3][HW] raw warped image
  std::vector<Geo> g;
  std::vector<Taps> t;
  std::vector<float> tex;     // [HW][12]
};
void warp_frame(const Dims& d, const float* Dk, int k, const float* src, const Cam& cam, const Pose& pose, Frame& f) {
  int HW = d.H * d.W;
  f.x.assign(3 * HW, 0.f); f.g.resize(HW); f.t.resize(HW); f.tex.assign(12 * (size_t)HW, 0.f);
  for (int py = 0; py < d.H; ++py)
    for (int px = 0; px < d.W; ++px) {
      int p = py * d.W + px;
      float D = depth_at(d, Dk, k, px, py);
      Geo g = reproject(px, py, D, cam, pose, d.W, d.H, 1e-7f, 1e-3f);
      Taps t = make_taps(g.u, g.v, d.W, d.H);
      f.g[p] = g; f.t[p] = t;
      for (int c = 0; c < 3; ++c) {
        const float* s = src + (size_t)c * HW;
        float i00 = s[t.y0 * d.W + t.x0], i01 = s[t.y0 * d.W + t.x1], i10 = s[t.y1 * d.W + t.x0], i11 = s[t.y1 * d.W + t.x1];
        f.tex[12 * (size_t)p + 4 * c + 0] = i00; f.tex[12 * (size_t)p + 4 * c + 1] = i01;
        f.tex[12 * (size_t)p + 4 * c + 2] = i10; f.tex[12 * (size_t)p + 4 * c + 3] = i11;
        f.x[(size_t)c * HW + p] = bilerp(i00, i01, i10, i11, t.wx, t.wy);
      }
    }
}
// window moments of channel image x (and cross moment with y) at pixel (px, py), reflect padding
void moments(const float* x, const float* y, int H, int W, int px, int py, float& mu, float& exx, float& exy) {
  float s = 0, sxx = 0, sxy = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      int q = reflect_clamp(py + dy, H) * W + reflect_clamp(px + dx, W);
      s += x[q]; sxx += x[q] * x[q]; sxy += x[q] * y[q];
    }
  mu = s / 9.f; exx = sxx / 9.f; exy = sxy / 9.f;
}
}  // namespace

extern "C" {

// Forward + backward of the photometric part (smooth_weight = 0) with grad_loss = 1.
// flags: bit0 lcc, bit1 lcc_detach.  Outputs: loss, valid [B,N,S,HW], sel [B,S,HW], ab [B,N,S,2],
// grad_depth[k], grad_T [B,N,16], grad_srcs [B,N,3,HW].
int harness_run(int B, int N, int S, int H, int W, unsigned flags, const float* tgt, const float* srcs,
                const float* const* depth, const float* K, const float* T, float* loss_out, uint8_t* valid,
                uint8_t* sel_out, float* ab_out, float* const* grad_depth, float* grad_T, float* grad_srcs) {
  Dims d; d.B = B; d.N = N; d.S = S; d.H = H; d.W = W;
  for (int k = 0; k < S; ++k) { d.h[k] = H >> k; d.w[k] = W >> k; d.ry[k] = (float)((double)d.h[k] / H); d.rx[k] = (float)((double)d.w[k] / W); }
  const int HW = H * W;
  const float alpha = 0.85f, c1 = 1e-4f, c2 = 9e-4f;
  const bool lcc = flags & 1u, detach = flags & 2u;
  double loss = 0.0;
  memset(grad_srcs, 0, sizeof(float) * (size_t)B * N * 3 * HW);
  memset(grad_T, 0, sizeof(float) * (size_t)B * N * 16);
  const float wscale = 1.0f / ((float)S * (float)B * (float)HW);
  for (int b = 0; b < B; ++b) {
    const float* tg = tgt + (size_t)b * 3 * HW;
    Cam cam = cam_of(K + 9 * b);
    // target window moments
    std::vector<float> muy(3 * HW), sgy(3 * HW);
    for (int c = 0; c < 3; ++c)
      for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
          float mu, eyy, t2;
          moments(tg + (size_t)c * HW, tg + (size_t)c * HW, H, W, px, py, mu, eyy, t2);
          muy[(size_t)c * HW + py * W + px] = mu; sgy[(size_t)c * HW + py * W + px] = eyy - mu * mu;
        }
    // photometric error image of one candidate, evaluated as k_photo_fwd does: window SUMS -> colvo_pe.cuh.
    // cf (optional): the unit-weight SSIM adjoint coefficients of every window, [3][HW]
    auto pe_image = [&](const float* x, float a, float bb, std::vector<float>& pe, std::vector<float>* dpa, std::vector<float>* dpb,
                        std::vector<Coef>* cf) {
      pe.assign(HW, 0.f);
      if (dpa) { dpa->assign(HW, 0.f); dpb->assign(HW, 0.f); }
      if (cf) cf->assign(3 * (size_t)HW, Coef{0, 0, 0});
      const float a1[1] = {a}, b1[1] = {bb};
      const CalV<1> cal = make_calv<1>(a1, b1, alpha);
      for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
          int p = py * W + px;
          WinY wy;
          Vn<1> Sx[3], Sxx[3], Sxy[3], xc[3], ca[3], cb[3], cg[3], da, db;
          for (int c = 0; c < 3; ++c) {
            float mu, exx, exy;
            moments(x + (size_t)c * HW, tg + (size_t)c * HW, H, W, px, py, mu, exx, exy);
            Sx[c] = Vn<1>(mu * 9.f); Sxx[c] = Vn<1>(exx * 9.f); Sxy[c] = Vn<1>(exy * 9.f);
            xc[c] = Vn<1>(x[(size_t)c * HW + p]);
            wy.muy[c] = muy[(size_t)c * HW + p]; wy.sgy[c] = sgy[(size_t)c * HW + p]; wy.yc[c] = tg[(size_t)c * HW + p];
          }
          winy_derive(wy, c1, c2);
          if (dpa || cf) {
            pe[p] = pe_fused<1>(Sx, Sxx, Sxy, xc, wy, cal, alpha, c1, c2, ca, cb, cg, da, db).v / 3.f;
            if (dpa) { (*dpa)[p] = da.v / 3.f; (*dpb)[p] = db.v / 3.f; }
            if (cf) for (int c = 0; c < 3; ++c) (*cf)[(size_t)c * HW + p] = Coef{ca[c].v, cb[c].v, cg[c].v};
          } else {
            pe[p] = pe_value3v<1>(Sx, Sxx, Sxy, xc, wy, cal, alpha, c1, c2).v / 3.f;
          }
        }
    };
    std::vector<std::vector<float>> ident(N);
    for (int n = 0; n < N; ++n) pe_image(srcs + ((size_t)b * N + n) * 3 * HW, 1.f, 0.f, ident[n], nullptr, nullptr, nullptr);
    for (int k = 0; k < S; ++k) {
      const float* Dk = depth[k] + (size_t)b * d.h[k] * d.w[k];
      std::vector<Frame> fr(N);
      std::vector<std::vector<float>> pe(N), dpa(N), dpb(N);
      std::vector<std::vector<Coef>> cfw(N);
      std::vector<float> a(N, 1.f), bb(N, 0.f);
      std::vector<double> st_n(N, 0), st_mx(N, 0), st_my(N, 0), st_inv(N, 0);
      for (int n = 0; n < N; ++n) {
        Pose pose = pose_of(T + ((size_t)b * N + n) * 16);
        warp_frame(d, Dk, k, srcs + ((size_t)b * N + n) * 3 * HW, cam, pose, fr[n]);
        double s[5] = {0, 0, 0, 0, 0};
        for (int p = 0; p < HW; ++p) {
          valid[(((size_t)b * N + n) * S + k) * HW + p] = fr[n].g[p].valid;
          if (!fr[n].g[p].valid) continue;
          for (int c = 0; c < 3; ++c) {
            double x = fr[n].x[(size_t)c * HW + p], y = tg[(size_t)c * HW + p];
            s[0] += 1; s[1] += x; s[2] += y; s[3] += x * x; s[4] += x * y;
          }
        }
        if (lcc && s[0] > 0) {
          double mx = s[1] / s[0], my = s[2] / s[0], var = s[3] / s[0] - mx * mx, cov = s[4] / s[0] - mx * my;
          double aa = cov / (var + 1e-6);
          a[n] = (float)aa; bb[n] = (float)(my - aa * mx);
          st_n[n] = s[0]; st_mx[n] = mx; st_my[n] = my; st_inv[n] = 1.0 / (s[0] * (var + 1e-6));
        }
        ab_out[(((size_t)b * N + n) * S + k) * 2 + 0] = a[n];
        ab_out[(((size_t)b * N + n) * S + k) * 2 + 1] = bb[n];
        pe_image(fr[n].x.data(), a[n], bb[n], pe[n], &dpa[n], &dpb[n], &cfw[n]);
      }
      std::vector<uint8_t> sel(HW);
      std::vector<double> Ga(N, 0), Gb(N, 0);
      for (int p = 0; p < HW; ++p) {
        float best = ident[0][p]; int s = 0;
        for (int n = 1; n < N; ++n) if (ident[n][p] < best) { best = ident[n][p]; s = n; }
        for (int n = 0; n < N; ++n) if (pe[n][p] < best) { best = pe[n][p]; s = N + n; }
        sel[p] = (uint8_t)s; sel_out[((size_t)b * S + k) * HW + p] = (uint8_t)s;
        loss += best;
        if (s >= N) { Ga[s - N] += dpa[s - N][p]; Gb[s - N] += dpb[s - N][p]; }
      }
      // ---- backward of this (b, k) ----
      std::vector<float> dD(HW, 0.f);
      for (int n = 0; n < N; ++n) {
        Pose pose = pose_of(T + ((size_t)b * N + n) * 16);
        double ga = Ga[n] * wscale, gb = Gb[n] * wscale;
        float Pc = 0, Qc = 0;
        if (lcc && !detach && st_n[n] > 0) { Pc = (float)((ga - gb * st_mx[n]) * st_inv[n]); Qc = (float)(gb * a[n] / st_n[n]); }
        // the winner's coefficient fields (unit weight in cfw; the loss weight is applied here), zeros elsewhere
        std::vector<Coef> cf(3 * (size_t)HW);
        for (int c = 0; c < 3; ++c)
          for (int p = 0; p < HW; ++p) {
            Coef q = {0, 0, 0};
            if (sel[p] == N + n) { const Coef& u = cfw[n][(size_t)c * HW + p]; q = Coef{u.ca * wscale, u.cb * wscale, u.cg * wscale}; }
            cf[(size_t)c * HW + p] = q;
          }
        float gp[12] = {0};
        float* gs = grad_srcs + ((size_t)b * N + n) * 3 * HW;
        for (int py = 0; py < H; ++py)
          for (int px = 0; px < W; ++px) {
            int p = py * W + px;
            const Geo& g = fr[n].g[p]; const Taps& t = fr[n].t[p];
            float wq = (sel[p] == N + n) ? wscale * (1.f - alpha) / 3.f : 0.f;
            float du = 0, dv = 0;
            for (int c = 0; c < 3; ++c) {
              float A = 0, Bc = 0, G = 0;
              for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                  float m = reflect_mult(py, py + dy, H) * reflect_mult(px, px + dx, W);
                  if (m == 0.f) continue;
                  const Coef& q = cf[(size_t)c * HW + (py + dy) * W + (px + dx)];
                  A += m * q.ca; Bc += m * q.cb; G += m * q.cg;
                }
              float xq = fr[n].x[(size_t)c * HW + p], yq = tg[(size_t)c * HW + p];
              float gq = A + xq * Bc + yq * G + wq * a[n] * sgn(a[n] * xq + bb[n] - yq);
              float l = g.valid ? (Pc * ((yq - (float)st_my[n]) - 2.f * a[n] * (xq - (float)st_mx[n])) - Qc) : 0.f;
              float hq = gq + l;
              const float* tx = &fr[n].tex[12 * (size_t)p + 4 * c];
              du += hq * ((1.f - t.wy) * (tx[1] - tx[0]) + t.wy * (tx[3] - tx[2]));
              dv += hq * ((1.f - t.wx) * (tx[2] - tx[0]) + t.wx * (tx[3] - tx[1]));
              float* gc = gs + (size_t)c * HW;
              gc[t.y0 * W + t.x0] += (1.f - t.wx) * (1.f - t.wy) * hq;
              gc[t.y0 * W + t.x1] += t.wx * (1.f - t.wy) * hq;
              gc[t.y1 * W + t.x0] += (1.f - t.wx) * t.wy * hq;
              gc[t.y1 * W + t.x1] += t.wx * t.wy * hq;
            }
            if (!t.gx) du = 0;
            if (!t.gy) dv = 0;
            float dXp[3];
            dD[p] += project_adjoint(g, cam, pose, du, dv, 0.f, dXp);
            float wv[3] = {dXp[0] * g.Z, dXp[1] * g.Z, dXp[2] * g.Z}, tv[3] = {dXp[0], dXp[1], dXp[2]}, e[12];
            pose_grad_expand(wv, tv, g.rx, g.ry, e);
            for (int q = 0; q < 12; ++q) gp[q] += e[q];
          }
        float* gt = grad_T + ((size_t)b * N + n) * 16;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) gt[4 * i + j] += gp[3 * i + j]; gt[4 * i + 3] += gp[9 + i]; }
      }
      // up-sample adjoint (gather form, as k_depth_gather)
      float* gd = grad_depth[k] + (size_t)b * d.h[k] * d.w[k];
      if (k == 0) { memcpy(gd, dD.data(), sizeof(float) * HW); }
      else {
        int hk = d.h[k], wk = d.w[k];
        for (int i = 0; i < hk; ++i)
          for (int j = 0; j < wk; ++j) {
            float acc = 0;
            for (int v = 0; v < H; ++v) {
              Axis ay = upsample_axis(v, d.ry[k], hk);
              float wy = ((ay.i0 == i) ? (1.f - ay.w1) : 0.f) + ((ay.i1 == i) ? ay.w1 : 0.f);
              if (wy == 0.f) continue;
              for (int u = 0; u < W; ++u) {
                Axis ax = upsample_axis(u, d.rx[k], wk);
                float wx = ((ax.i0 == j) ? (1.f - ax.w1) : 0.f) + ((ax.i1 == j) ? ax.w1 : 0.f);
                if (wx != 0.f) acc += wy * wx * dD[v * W + u];
              }
            }
            gd[i * wk + j] = acc;
          }
      }
    }
  }
  *loss_out = (float)(loss / ((double)S * B * HW));
  return 0;
}


// The packed two-source window evaluation of the forward tile kernel (colvo_pe.cuh: pe_value3v / pe_fused, lanes =
// sources) against the scalar per-channel reference (colvo_math.cuh: pe_channel / ssim_coef) on `nw` windows.
//   x [nw][2][3][9] raw warped taps, y [nw][3][9] target taps (tap 4 = centre), ab [nw][2][2]
// out_ref / out_new [nw][2][12]: pe, dpa, dpb, (ca, cb, cg) x 3 channels per source; out_one [nw][12]: pe_fused<1> on source 0.
int harness_pe_fused(int nw, const float* x, const float* y, const float* ab, float alpha, float c1, float c2,
                     float* out_ref, float* out_new, float* out_one, float* out_val) {
  for (int i = 0; i < nw; ++i) {
    const float* xi = x + (size_t)i * 54;
    const float* yi = y + (size_t)i * 27;
    WinY wy;
    float Sxs[2][3], Sxxs[2][3], Sxys[2][3], xcs[2][3];
    for (int c = 0; c < 3; ++c) {
      float sy = 0, syy = 0;
      for (int t = 0; t < 9; ++t) { sy += yi[c * 9 + t]; syy += yi[c * 9 + t] * yi[c * 9 + t]; }
      wy.muy[c] = sy / 9.f; wy.sgy[c] = syy / 9.f - wy.muy[c] * wy.muy[c]; wy.yc[c] = yi[c * 9 + 4];
      for (int n = 0; n < 2; ++n) {
        float s = 0, sxx = 0, sxy = 0;
        for (int t = 0; t < 9; ++t) { float v = xi[(n * 3 + c) * 9 + t]; s += v; sxx += v * v; sxy += v * yi[c * 9 + t]; }
        Sxs[n][c] = s; Sxxs[n][c] = sxx; Sxys[n][c] = sxy; xcs[n][c] = xi[(n * 3 + c) * 9 + 4];
      }
    }
    winy_derive(wy, c1, c2);
    const float a[2] = {ab[i * 4 + 0], ab[i * 4 + 2]}, b[2] = {ab[i * 4 + 1], ab[i * 4 + 3]};
    // scalar reference
    for (int n = 0; n < 2; ++n) {
      float pe = 0, dpa = 0, dpb = 0;
      float* o = out_ref + ((size_t)i * 2 + n) * 12;
      for (int c = 0; c < 3; ++c) {
        Coef cf;
        pe += pe_channel(Sxs[n][c] / 9.f, Sxxs[n][c] / 9.f, Sxys[n][c] / 9.f, wy.muy[c], wy.sgy[c], xcs[n][c], wy.yc[c], a[n], b[n],
                         alpha, c1, c2, &dpa, &dpb, true, cf);
        o[3 + 3 * c] = cf.ca; o[4 + 3 * c] = cf.cb; o[5 + 3 * c] = cf.cg;
      }
      o[0] = pe; o[1] = dpa; o[2] = dpb;
    }
    // packed: lanes = sources
    {
      Vn<2> Sx[3], Sxx[3], Sxy[3], xc[3], ca[3], cb[3], cg[3], dpa, dpb;
      for (int c = 0; c < 3; ++c) {
        Sx[c] = Vn<2>(Sxs[0][c], Sxs[1][c]); Sxx[c] = Vn<2>(Sxxs[0][c], Sxxs[1][c]);
        Sxy[c] = Vn<2>(Sxys[0][c], Sxys[1][c]); xc[c] = Vn<2>(xcs[0][c], xcs[1][c]);
      }
      const CalV<2> k = make_calv<2>(a, b, alpha);
      const Vn<2> pe = pe_fused<2>(Sx, Sxx, Sxy, xc, wy, k, alpha, c1, c2, ca, cb, cg, dpa, dpb);
      const Vn<2> pv = pe_value3v<2>(Sx, Sxx, Sxy, xc, wy, k, alpha, c1, c2);
      for (int n = 0; n < 2; ++n) {
        float* o = out_new + ((size_t)i * 2 + n) * 12;
        o[0] = pe.lane(n); o[1] = dpa.lane(n); o[2] = dpb.lane(n);
        for (int c = 0; c < 3; ++c) { o[3 + 3 * c] = ca[c].lane(n); o[4 + 3 * c] = cb[c].lane(n); o[5 + 3 * c] = cg[c].lane(n); }
        out_val[i * 2 + n] = pv.lane(n);
      }
    }
    {
      Vn<1> Sx[3], Sxx[3], Sxy[3], xc[3], ca[3], cb[3], cg[3], dpa, dpb;
      for (int c = 0; c < 3; ++c) { Sx[c] = Vn<1>(Sxs[0][c]); Sxx[c] = Vn<1>(Sxxs[0][c]); Sxy[c] = Vn<1>(Sxys[0][c]); xc[c] = Vn<1>(xcs[0][c]); }
      const float a1[1] = {a[0]}, b1[1] = {b[0]};
      const CalV<1> k = make_calv<1>(a1, b1, alpha);
      const Vn<1> pe = pe_fused<1>(Sx, Sxx, Sxy, xc, wy, k, alpha, c1, c2, ca, cb, cg, dpa, dpb);
      float* o = out_one + (size_t)i * 12;
      o[0] = pe.v; o[1] = dpa.v; o[2] = dpb.v;
      for (int c = 0; c < 3; ++c) { o[3 + 3 * c] = ca[c].v; o[4 + 3 * c] = cb[c].v; o[5 + 3 * c] = cg[c].v; }
    }
  }
  return 0;
}

}  // extern "C"
