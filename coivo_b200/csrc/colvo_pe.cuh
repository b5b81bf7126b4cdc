// Photometric error of a 3x3 window and its adjoint, written once for N = 1 (scalar) and N = 2 (the two source
// frames packed in one fp32x2 register pair, colvo_f2.cuh) -- SURVEY.md section 8(a) rows 6-7 and appendix A.
//
// A window is described by its SUMS over the 9 taps (Sx, Sxx, Sxy per channel and source, formed separably by
// the strip walk of k_photo_fwd) and by the target side (WinY).  pe_value3v() gives 3 * pe of every source's
// candidate; pe_fused() additionally gives, for every source, the unit-weight SSIM adjoint coefficients
// (ca, cb, cg) of the gather form  d pe_p / d x_q = ca_p + x_q cb_p + y_q cg_p  and the terms of d pe / d a,
// d pe / d b (LCC adjoint), re-using every intermediate of the value.  Host-compilable (tests/cpu_harness).
#pragma once
#include "colvo_f2.cuh"

namespace colvo {

// the target side of one window, shared by every candidate and source
struct WinY {
  float muy[3], sgy[3], yc[3];
  float muy2[3], k1[3], k2[3];             // 2 mu_y, mu_y^2 + C1, var_y + C2
};
CV_HD void winy_derive(WinY& y, float c1, float c2) {
  for (int c = 0; c < 3; ++c) {
    y.muy2[c] = y.muy[c] + y.muy[c];
    y.k1[c] = f_fma(y.muy[c], y.muy[c], c1);
    y.k2[c] = y.sgy[c] + c2;
  }
}

// per-source constants of the calibration (a, b); the window moments stay scaled by 9 (s9 = 9 var_x,
// sxy9 = 9 cov_xy), the factor rides on a^2/9 and 2a/9
template <int NS>
struct CalV {
  Vn<NS> a, b, a2_9, ta_9, fc;             // fc = -alpha a / 27: weight of the coefficient fields
};
template <int NS>
CV_HD CalV<NS> make_calv(const float (&a)[NS], const float (&b)[NS], float alpha) {
  CalV<NS> k;
  for (int n = 0; n < NS; ++n) {
    k.a.set(n, a[n]);
    k.b.set(n, b[n]);
    k.a2_9.set(n, a[n] * a[n] * (1.0f / 9.0f));
    k.ta_9.set(n, 2.f * a[n] * (1.0f / 9.0f));
    k.fc.set(n, -alpha * a[n] * (1.0f / 27.0f));
  }
  return k;
}

// 3 * pe of every source's candidate (value only)
//   pe = alpha * mean_c clamp((1 - SSIM_c)/2, 0, 1) + (1 - alpha) * mean_c |a x_c + b - y_c|
template <int NS>
CV_HD Vn<NS> pe_value3v(const Vn<NS> (&Sx)[3], const Vn<NS> (&Sxx)[3], const Vn<NS> (&Sxy)[3], const Vn<NS> (&xc)[3],
                        const WinY& y, const CalV<NS>& k, float alpha, float c1, float c2) {
  Vn<NS> pe = bc<NS>(0.f);
  for (int c = 0; c < 3; ++c) {
    const Vn<NS> mu = Sx[c] * bc<NS>(1.0f / 9.0f);
    const Vn<NS> s9 = fma2(-mu, Sx[c], Sxx[c]);
    const Vn<NS> sxy9 = fma2(-Sx[c], bc<NS>(y.muy[c]), Sxy[c]);
    const Vn<NS> mut = fma2(k.a, mu, k.b);
    const Vn<NS> A1 = fma2(mut, bc<NS>(y.muy2[c]), bc<NS>(c1));
    const Vn<NS> A2 = fma2(k.ta_9, sxy9, bc<NS>(c2));
    const Vn<NS> B1 = fma2(mut, mut, bc<NS>(y.k1[c]));
    const Vn<NS> B2 = fma2(k.a2_9, s9, bc<NS>(y.k2[c]));
    const Vn<NS> S = (A1 * rcp2(B1 * B2)) * A2;
    const Vn<NS> t = sat2(fma2(bc<NS>(-0.5f), S, bc<NS>(0.5f)));
    const Vn<NS> diff = fma2(k.a, xc[c], k.b) - bc<NS>(y.yc[c]);
    pe = fma2(bc<NS>(alpha), t, pe);
    pe = fma2(bc<NS>(1.f - alpha), abs2(diff), pe);
  }
  return pe;
}

// value and adjoint pieces of every source's candidate in one pass.  With act2 = -alpha inside the clamp (else 0),
// iB = 1 / (B1 B2), P = A1 iB:
//   dS/dmu~ = 2 dmu',  dmu' = iB (mu_y A2 - S mu~ B2);   dS/ds~ = -S iB B1 = dsx;   dS/ds~xy = 2 P
//   3 dpe/da += act2 (dmu' mu + (a dsx / 9) s9 + (P / 9) sxy9) + sg x_c;   3 dpe/db += act2 dmu' + sg
//   (ca, cb, cg) = F (dmu' - mu G - mu_y P,  G,  P),   G = a dsx,   F = -alpha a / 27 inside the clamp (else 0)
// (the same quantities as colvo_math.cuh::coef_from_parts / pe_channel, with the common factors folded).
template <int NS>
CV_HD Vn<NS> pe_fused(const Vn<NS> (&Sx)[3], const Vn<NS> (&Sxx)[3], const Vn<NS> (&Sxy)[3], const Vn<NS> (&xc)[3],
                      const WinY& y, const CalV<NS>& k, float alpha, float c1, float c2, Vn<NS> (&ca)[3],
                      Vn<NS> (&cb)[3], Vn<NS> (&cg)[3], Vn<NS>& dpa, Vn<NS>& dpb) {
  Vn<NS> pe = bc<NS>(0.f);
  dpa = bc<NS>(0.f);
  dpb = bc<NS>(0.f);
  for (int c = 0; c < 3; ++c) {
    const Vn<NS> mu = Sx[c] * bc<NS>(1.0f / 9.0f);
    const Vn<NS> s9 = fma2(-mu, Sx[c], Sxx[c]);
    const Vn<NS> sxy9 = fma2(-Sx[c], bc<NS>(y.muy[c]), Sxy[c]);
    const Vn<NS> mut = fma2(k.a, mu, k.b);
    const Vn<NS> A1 = fma2(mut, bc<NS>(y.muy2[c]), bc<NS>(c1));
    const Vn<NS> A2 = fma2(k.ta_9, sxy9, bc<NS>(c2));
    const Vn<NS> B1 = fma2(mut, mut, bc<NS>(y.k1[c]));
    const Vn<NS> B2 = fma2(k.a2_9, s9, bc<NS>(y.k2[c]));
    const Vn<NS> iB = rcp2(B1 * B2);
    const Vn<NS> P = A1 * iB;
    const Vn<NS> S = P * A2;
    const Vn<NS> t = fma2(bc<NS>(-0.5f), S, bc<NS>(0.5f));
    const Vn<NS> tc = sat2(t);
    const Vn<NS> diff = fma2(k.a, xc[c], k.b) - bc<NS>(y.yc[c]);
    pe = fma2(bc<NS>(alpha), tc, pe);
    pe = fma2(bc<NS>(1.f - alpha), abs2(diff), pe);
    // ---- adjoint ----
    const Vn<NS> nG = k.a * (S * (iB * B1));                          // -G = -a dsx
    const Vn<NS> dmup = iB * fma2(bc<NS>(y.muy[c]), A2, -((S * mut) * B2));
    const Vn<NS> T3 = fma2(P * bc<NS>(1.0f / 9.0f), sxy9, fma2(nG * bc<NS>(-1.0f / 9.0f), s9, dmup * mu));
    Vn<NS> act2, F, sg;
    for (int n = 0; n < NS; ++n) {
      const bool in01 = tc.lane(n) == t.lane(n);                      // the clamp passed t through
      act2.set(n, in01 ? -alpha : 0.f);
      F.set(n, in01 ? k.fc.lane(n) : 0.f);
      sg.set(n, sgn_scaled(1.f - alpha, diff.lane(n)));               // sub-gradient sign(0) = 0
    }
    dpa = fma2(sg, xc[c], fma2(act2, T3, dpa));
    dpb = fma2(act2, dmup, dpb) + sg;
    ca[c] = F * fma2(-bc<NS>(y.muy[c]), P, fma2(mu, nG, dmup));
    cb[c] = -(F * nG);
    cg[c] = F * P;
  }
  return pe;
}

}  // namespace colvo
