// Packed fp32x2 arithmetic (sm_100a FFMA2 / FADD2 / FMUL2) for the photometric-loss kernels.
//
// The path warps N = 2 neighbouring frames per target pixel and both go through identical arithmetic, so the
// two sources ride in the two lanes of one 64-bit register pair: `Vn<2>` wraps a float2 whose operators map to
// the sm_100 packed intrinsics (__ffma2_rn / __fadd2_rn / __fmul2_rn, crt/sm_100_rt.h) -- one issue slot for
// two FMAs.  Per-component numerics are those of fmaf / __fadd_rn / __fmul_rn, so tolerances are untouched.
// SASS accepts a scalar register broadcast to both lanes (R.F32), an immediate, |x| and -x on packed operands,
// so `bc(s)` (the same scalar in both lanes), abs2() and negation are free.  `Vn<1>` is the same interface on a
// plain float for N = 1; kernels are written once against Vn<NS>.
//
// The packed intrinsics are never contracted by the compiler: write fma2() where a fused multiply-add is wanted.
// Host build (tests/cpu_harness): the same interface, component by component.
#pragma once
#include "colvo_math.cuh"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

namespace colvo {

template <int NS> struct Vn;

template <> struct Vn<1> {
  float v;
  CV_HD Vn() {}
  CV_HD explicit Vn(float a) : v(a) {}
  CV_HD float lane(int) const { return v; }
  CV_HD void set(int, float a) { v = a; }
};

template <> struct Vn<2> {
#if defined(__CUDACC__)
  float2 v;
  CV_HD Vn() {}
  CV_HD explicit Vn(float a) { v.x = a; v.y = a; }
  CV_HD Vn(float a, float b) { v.x = a; v.y = b; }
#else
  struct { float x, y; } v;
  Vn() {}
  explicit Vn(float a) { v.x = a; v.y = a; }
  Vn(float a, float b) { v.x = a; v.y = b; }
#endif
  CV_HD float lane(int i) const { return i == 0 ? v.x : v.y; }
  CV_HD void set(int i, float a) { if (i == 0) v.x = a; else v.y = a; }
};

typedef Vn<2> f2;

// ---- N = 1: plain fp32 (the compiler contracts a * b + c on its own) ----
CV_HD Vn<1> bc1(float a) { return Vn<1>(a); }
CV_HD Vn<1> operator+(Vn<1> a, Vn<1> b) { return Vn<1>(a.v + b.v); }
CV_HD Vn<1> operator-(Vn<1> a, Vn<1> b) { return Vn<1>(a.v - b.v); }
CV_HD Vn<1> operator*(Vn<1> a, Vn<1> b) { return Vn<1>(a.v * b.v); }
CV_HD Vn<1> operator-(Vn<1> a) { return Vn<1>(-a.v); }
CV_HD Vn<1> fma2(Vn<1> a, Vn<1> b, Vn<1> c) { return Vn<1>(f_fma(a.v, b.v, c.v)); }
CV_HD Vn<1> abs2(Vn<1> a) { return Vn<1>(fabsf(a.v)); }
CV_HD Vn<1> rcp2(Vn<1> a) { return Vn<1>(f_rcp(a.v)); }
CV_HD Vn<1> sat2(Vn<1> a) { return Vn<1>(fminf(fmaxf(a.v, 0.f), 1.f)); }

// ---- N = 2: packed ----
#if defined(__CUDA_ARCH__)
CV_HD f2 operator+(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
CV_HD f2 operator-(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y)); return r; }
CV_HD f2 operator*(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
CV_HD f2 fma2(f2 a, f2 b, f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
CV_HD f2 sat2(f2 a) { return f2(__saturatef(a.v.x), __saturatef(a.v.y)); }     // no packed saturate: two FADD.SAT
#else
CV_HD f2 operator+(f2 a, f2 b) { return f2(p_add(a.v.x, b.v.x), p_add(a.v.y, b.v.y)); }
CV_HD f2 operator-(f2 a, f2 b) { return f2(p_sub(a.v.x, b.v.x), p_sub(a.v.y, b.v.y)); }
CV_HD f2 operator*(f2 a, f2 b) { return f2(p_mul(a.v.x, b.v.x), p_mul(a.v.y, b.v.y)); }
CV_HD f2 fma2(f2 a, f2 b, f2 c) { return f2(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y)); }
CV_HD f2 sat2(f2 a) { return f2(fminf(fmaxf(a.v.x, 0.f), 1.f), fminf(fmaxf(a.v.y, 0.f), 1.f)); }
#endif
CV_HD f2 operator-(f2 a) { return f2(-a.v.x, -a.v.y); }                          // folds into the consumer's operand modifier
CV_HD f2 abs2(f2 a) { return f2(fabsf(a.v.x), fabsf(a.v.y)); }                   // likewise (|R|.F32x2)
CV_HD f2 rcp2(f2 a) { return f2(f_rcp(a.v.x), f_rcp(a.v.y)); }                   // MUFU has no packed form

// ---- single-rounded ("pinned") multiply / add for the geometry chain: never contracted into an FMA ----
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry an explicit rounding mode (and
// with -fmad=false, and through volatile asm): a packed add must never consume a packed product in a pinned chain.
// So the pinned multiply is packed (FMUL2: per component == __fmul_rn) and the pinned add is one scalar __fadd_rn
// per lane, which ptxas leaves alone (checked in SASS: FMUL2, FADD, FADD).  Everywhere else in the packed code a
// written a * b + c may therefore be executed as fma(a, b, c) -- harmless outside the pinned chain.
CV_HD Vn<1> pmul(Vn<1> a, Vn<1> b) { return Vn<1>(p_mul(a.v, b.v)); }
CV_HD Vn<1> padd(Vn<1> a, Vn<1> b) { return Vn<1>(p_add(a.v, b.v)); }
CV_HD f2 pmul(f2 a, f2 b) { return a * b; }
CV_HD f2 padd(f2 a, f2 b) { return f2(p_add(a.v.x, b.v.x), p_add(a.v.y, b.v.y)); }
CV_HD Vn<1> prcp(Vn<1> a) { return Vn<1>(p_rcp(a.v)); }
CV_HD f2 prcp(f2 a) { return f2(p_rcp(a.v.x), p_rcp(a.v.y)); }

// the same scalar in every lane (a register broadcast in SASS: free)
template <int NS> CV_HD Vn<NS> bc(float a) { return Vn<NS>(a); }

// M consecutive lanes of an NS-lane value, starting at lane n0 (M == NS: the value itself; M == 1: one lane)
template <int M, int NS> struct LaneSub;
template <int NS> struct LaneSub<NS, NS> {
  static CV_HD Vn<NS> get(const Vn<NS>& v, int) { return v; }
  static CV_HD void put(Vn<NS>& d, int, const Vn<NS>& s) { d = s; }
};
template <> struct LaneSub<1, 2> {
  static CV_HD Vn<1> get(const Vn<2>& v, int n0) { return Vn<1>(v.lane(n0)); }
  static CV_HD void put(Vn<2>& d, int n0, const Vn<1>& s) { d.set(n0, s.v); }
};

// lane-wise select: the predicates are per lane (two FSEL)
CV_HD Vn<1> select2(const bool (&p)[1], Vn<1> a, Vn<1> b) { return p[0] ? a : b; }
CV_HD f2 select2(const bool (&p)[2], f2 a, f2 b) { return f2(p[0] ? a.v.x : b.v.x, p[1] ? a.v.y : b.v.y); }

}  // namespace colvo
