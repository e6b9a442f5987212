// In-register DFT building blocks for the sm_100a kernels.
//
// Every transform of length N = R1 * R2 in this library is a two-pass Stockham
// decomposition: radix-R1 butterflies held entirely in registers, one twiddle
// multiply, ONE shared-memory exchange, radix-R2 butterflies in registers.  The
// radix-R butterflies below are built at compile time from small prime kernels
// (2, 4, odd primes by the symmetric-pair formula); all internal twiddles are
// compile-time constants produced by an exact-octant constexpr sin/cos, so they
// become FFMA immediates.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

#define BFFT_HD __host__ __device__

namespace bfft
{

// ---------------------------------------------------------------- constexpr trig
constexpr double kPi = 3.14159265358979323846264338327950288;

BFFT_HD constexpr double sin_taylor(double x) // |x| <= pi/4
{
  double x2 = x * x, term = x, sum = x;
  for (int k = 1; k < 14; k++)
  {
    term *= -x2 / ((2.0 * k) * (2.0 * k + 1.0));
    sum += term;
  }
  return sum;
}
BFFT_HD constexpr double cos_taylor(double x) // |x| <= pi/4
{
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int k = 1; k < 14; k++)
  {
    term *= -x2 / ((2.0 * k - 1.0) * (2.0 * k));
    sum += term;
  }
  return sum;
}
struct cs_t
{
  double c, s;
};
// exp(2*pi*i * p / q) with exact symmetry reduction (exact 0 / +-1 where due)
BFFT_HD constexpr cs_t unit_root(long long p, long long q)
{
  p %= q;
  if (p < 0)
    p += q;
  // work in units of 1/(8q) turns: a = 8p in [0, 8q)
  bool conj = false, negc = false, swap = false;
  long long num = p, den = q; // angle = 2*pi*num/den, num/den in [0,1)
  if (2 * num > den)
  { // > half turn: conj symmetry
    num = den - num;
    conj = true;
  }
  if (4 * num > den)
  { // > quarter: cos -> -cos of (1/2 - f)
    num = den - 2 * num; // over 2*den
    den = 2 * den;
    negc = true;
  }
  if (8 * num > den)
  { // > eighth: swap sin/cos of (1/4 - f)
    num = den - 4 * num; // over 4*den
    den = 4 * den;
    swap = true;
  }
  double ang = 2.0 * kPi * (double) num / (double) den;
  double c = cos_taylor(ang), s = sin_taylor(ang);
  if (num == 0)
  {
    c = 1.0;
    s = 0.0;
  }
  if (swap)
  {
    double t = c;
    c = s;
    s = t;
  }
  if (negc)
    c = -c;
  if (conj)
    s = -s;
  return cs_t{c, s};
}

// ---------------------------------------------------------------- helpers
template <int I, int E, class F> BFFT_HD __forceinline__ void static_for(F &&f)
{
  if constexpr (I < E)
  {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

// Complex arithmetic on float2 = (re, im) with the sm_100a packed FP32 instructions
// (FADD2 / FMUL2 / FFMA2: two FP32 lanes per thread per issue slot).  ptxas folds the
// swap / sign patterns below into operand modifiers (.LO_HI, .NP, scalar broadcast), so a
// complex add is ONE instruction, a multiply by +-i is free, and a complex multiply is two.
#define BFFT_D __device__ __forceinline__
BFFT_D float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
BFFT_D float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a + i*b and a - i*b
BFFT_D float2 cadd_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }
BFFT_D float2 csub_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }
BFFT_D float2 cmul(float2 a, float2 b)
{
  return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
// a * conj(b)
BFFT_D float2 cmulc(float2 a, float2 b)
{
  return __ffma2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
// a * s for a real scalar s, and r*s + a
BFFT_D float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
BFFT_D float2 cfma_r(float s, float2 r, float2 a) { return __ffma2_rn(make_float2(s, s), r, a); }

// multiply by the compile-time constant exp(SIGN * 2*pi*i * P / Q)
template <int SIGN, int P, int Q> BFFT_D float2 cmul_root(float2 a)
{
  constexpr cs_t w = unit_root((long long) SIGN * P, Q);
  constexpr float c = (float) w.c, s = (float) w.s;
  if constexpr (w.s == 0.0 && w.c == 1.0)
    return a;
  else if constexpr (w.s == 0.0 && w.c == -1.0)
    return make_float2(-a.x, -a.y);
  else if constexpr (w.c == 0.0 && w.s == 1.0)
    return make_float2(-a.y, a.x);
  else if constexpr (w.c == 0.0 && w.s == -1.0)
    return make_float2(a.y, -a.x);
  else
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(s, s), __fmul2_rn(a, make_float2(c, c)));
}

BFFT_HD constexpr int pow2_part(int r)
{
  int p = 1;
  while (r % 2 == 0)
  {
    r /= 2;
    p *= 2;
  }
  return p;
}
BFFT_HD constexpr int gcd_of(int a, int b) { return b == 0 ? a : gcd_of(b, a % b); }
// modular inverse of a modulo m (gcd(a, m) == 1)
BFFT_HD constexpr int inv_mod(int a, int m)
{
  for (int x = 1; x < m; x++)
    if ((a * x) % m == 1)
      return x;
  return 1;
}
// first factor A of the split R = A * B used by Dft<R>: the power-of-two part when R also
// has an odd part (coprime split -> Good-Thomas, no twiddles), else 4 / 2 / smallest prime.
BFFT_HD constexpr int pick_factor(int r)
{
  const int p = pow2_part(r), q = r / p;
  if (p > 1 && q > 1)
    return p;
  if (q == 1)
    return r > 4 ? 4 : r;
  for (int f = 3; f * f <= r; f += 2)
    if (r % f == 0)
      return f;
  return r; // prime
}

// ---------------------------------------------------------------- butterflies
// Dft<R, SIGN>::run(v): in-place DFT of v[0..R), natural order in and out,
// X[k] = sum_n v[n] exp(SIGN * 2*pi*i * n*k / R).
template <int R, int SIGN, int A = pick_factor(R)> struct Dft;

template <int SIGN> struct Dft<1, SIGN, 1>
{
  BFFT_D static void run(float2 *) {}
};

template <int SIGN> struct Dft<2, SIGN, 2>
{
  BFFT_D static void run(float2 *v)
  {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <int SIGN> struct Dft<4, SIGN, 4>
{
  BFFT_D static void run(float2 *v)
  {
    float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    v[0] = cadd(t0, t2);
    v[2] = csub(t0, t2);
    if constexpr (SIGN > 0)
    { // t1 +- i*d
      v[1] = cadd_i(t1, d);
      v[3] = csub_i(t1, d);
    }
    else
    {
      v[1] = csub_i(t1, d);
      v[3] = cadd_i(t1, d);
    }
  }
};

// odd prime R: symmetric-pair formula
template <int R, int SIGN> struct Dft<R, SIGN, R>
{
  static_assert(R % 2 == 1 && R >= 3, "prime kernel expects an odd prime");
  BFFT_D static void run(float2 *v)
  {
    constexpr int H = (R - 1) / 2;
    float2 s[H + 1], d[H + 1];
    float2 x0 = v[0];
    float2 sum = x0;
    static_for<1, H + 1>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      s[k] = cadd(v[k], v[R - k]);
      d[k] = csub(v[k], v[R - k]);
      sum = cadd(sum, s[k]);
    });
    v[0] = sum;
    static_for<1, H + 1>([&](auto j_) {
      constexpr int j = decltype(j_)::value;
      float2 m = x0;
      float2 e;
      static_for<1, H + 1>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        constexpr cs_t w = unit_root((long long) j * k, R);
        constexpr float c = (float) w.c, sn = (float) (SIGN * w.s);
        m = cfma_r(c, s[k], m);
        if constexpr (k == 1)
          e = cscale(d[k], sn);
        else
          e = cfma_r(sn, d[k], e);
      });
      // X_j = m + i*e ; X_{R-j} = m - i*e
      v[j] = cadd_i(m, e);
      v[R - j] = csub_i(m, e);
    });
  }
};

// composite R = A * B (A = pick_factor(R)).
//   gcd(A, B) == 1: Good-Thomas prime-factor map, no twiddles:
//       n = (B*n1 + A*n2) mod R,  k = (B*(B^-1 mod A)*k1 + A*(A^-1 mod B)*k2) mod R
//   else Cooley-Tukey: n = n1*B + n2, k = k1 + A*k2,
//       X[k1 + A*k2] = sum_{n2} W_R^{n2*k1} W_B^{n2*k2} sum_{n1} x[n1*B+n2] W_A^{n1*k1}
template <int R, int SIGN, int A> struct Dft
{
  static_assert(R % A == 0 && A > 1 && A < R, "bad factorisation");
  BFFT_D static void run(float2 *v)
  {
    constexpr int B = R / A;
    constexpr bool PFA = gcd_of(A, B) == 1;
    float2 t[R]; // t[k1*B + n2]
    static_for<0, B>([&](auto n2_) {
      constexpr int n2 = decltype(n2_)::value;
      float2 a[A];
      static_for<0, A>([&](auto n1_) {
        constexpr int n1 = decltype(n1_)::value;
        constexpr int n = PFA ? (B * n1 + A * n2) % R : n1 * B + n2;
        a[n1] = v[n];
      });
      Dft<A, SIGN>::run(a);
      static_for<0, A>([&](auto k1_) {
        constexpr int k1 = decltype(k1_)::value;
        if constexpr (PFA)
          t[k1 * B + n2] = a[k1];
        else
          t[k1 * B + n2] = cmul_root<SIGN, n2 * k1, R>(a[k1]);
      });
    });
    static_for<0, A>([&](auto k1_) {
      constexpr int k1 = decltype(k1_)::value;
      float2 b[B];
      static_for<0, B>([&](auto n2_) {
        constexpr int n2 = decltype(n2_)::value;
        b[n2] = t[k1 * B + n2];
      });
      Dft<B, SIGN>::run(b);
      static_for<0, B>([&](auto k2_) {
        constexpr int k2 = decltype(k2_)::value;
        constexpr int k = PFA ? (B * inv_mod(B % A, A) * k1 + A * inv_mod(A % B, B) * k2) % R : k1 + A * k2;
        v[k] = b[k2];
      });
    });
  }
};

// ---------------------------------------------------------------- geometry
// Two-pass split N = R1 * R2 and tiling constants for the supported image edges.
//   KC  packed spectrum columns per column-pass chunk (NCOL = N/2 must divide)
//   PC  row pairs per row-pass chunk
template <int N> struct Geo;
//   GENERIC  1: split chosen by rule (tools/gen_geometries.py), only the unpruned variant of the fused kernel
//            is instantiated, its row slots hold exactly the window rows
#define BFFT_GEO_(N_, R1_, R2_, KC_, PC_, GEN_)                                                    \
  template <> struct Geo<N_>                                                                       \
  {                                                                                                \
    static constexpr int R1 = R1_, R2 = R2_, KC = KC_, PC = PC_;                                   \
    static constexpr bool GENERIC = GEN_;                                                          \
    static_assert(R1_ * R2_ == N_ && (R1_ % 2) == 0 && ((N_ / 2) % KC_) == 0, "geometry");         \
    static_assert(R2_ <= 32 && PC_ * R1_ <= 256 && PC_ * R2_ <= 256, "geometry");                  \
  };
#define BFFT_GEO(N_, R1_, R2_, KC_, PC_) BFFT_GEO_(N_, R1_, R2_, KC_, PC_, false)
#define BFFT_GEO_AUTO(N_, R1_, R2_, KC_, PC_) BFFT_GEO_(N_, R1_, R2_, KC_, PC_, true)
// (R1 even: the packed layout pairs sub-sequences n1 = 2k, 2k+1; R2 <= 32: one lane per sub-sequence in the
// first radix pass of a warp; PC * max(R1, R2) <= 256)
BFFT_GEO(32, 4, 8, 16, 16)
BFFT_GEO(36, 6, 6, 18, 16)
BFFT_GEO(48, 6, 8, 12, 16)
BFFT_GEO(64, 8, 8, 16, 16)
BFFT_GEO(96, 12, 8, 16, 16)
BFFT_GEO(128, 8, 16, 16, 16)
BFFT_GEO(160, 10, 16, 16, 16)
BFFT_GEO(192, 12, 16, 16, 16)
BFFT_GEO(224, 14, 16, 16, 16)
BFFT_GEO(256, 16, 16, 16, 16)
BFFT_GEO(288, 18, 16, 16, 12)
BFFT_GEO(320, 20, 16, 16, 12)
BFFT_GEO(360, 24, 15, 20, 10)
BFFT_GEO(384, 24, 16, 16, 10)
BFFT_GEO(400, 16, 25, 20, 10)
// further common box sizes (radices 2/3/5/7), not individually tuned
BFFT_GEO(100, 10, 10, 10, 16)
BFFT_GEO(120, 8, 15, 12, 16)
BFFT_GEO(144, 12, 12, 12, 16)
BFFT_GEO(200, 20, 10, 10, 12)
BFFT_GEO(216, 18, 12, 12, 12)
BFFT_GEO(240, 16, 15, 20, 16)
BFFT_GEO(300, 20, 15, 15, 12)
BFFT_GEO(336, 24, 14, 12, 10)
BFFT_GEO(420, 20, 21, 14, 12)
BFFT_GEO(432, 16, 27, 12, 8)
BFFT_GEO(448, 28, 16, 16, 8)
BFFT_GEO(480, 20, 24, 16, 10)
BFFT_GEO(500, 20, 25, 10, 10)
BFFT_GEO(512, 16, 32, 16, 8)
// every other even edge up to 512 with prime factors 2 / 3 / 5 / 7 (490 = 2 * 5 * 7^2 has no split with an even
// R1 <= 32 and R2 <= 32), generated by tools/gen_geometries.py
BFFT_GEO_AUTO(16, 4, 4, 8, 16)
BFFT_GEO_AUTO(18, 6, 3, 9, 16)
BFFT_GEO_AUTO(20, 4, 5, 10, 16)
BFFT_GEO_AUTO(24, 12, 2, 12, 16)
BFFT_GEO_AUTO(28, 14, 2, 14, 16)
BFFT_GEO_AUTO(30, 6, 5, 15, 16)
BFFT_GEO_AUTO(40, 8, 5, 20, 16)
BFFT_GEO_AUTO(42, 14, 3, 21, 16)
BFFT_GEO_AUTO(50, 10, 5, 5, 16)
BFFT_GEO_AUTO(54, 18, 3, 9, 14)
BFFT_GEO_AUTO(56, 8, 7, 14, 16)
BFFT_GEO_AUTO(60, 20, 3, 15, 12)
BFFT_GEO_AUTO(70, 14, 5, 7, 16)
BFFT_GEO_AUTO(72, 8, 9, 18, 16)
BFFT_GEO_AUTO(80, 20, 4, 20, 12)
BFFT_GEO_AUTO(84, 6, 14, 21, 16)
BFFT_GEO_AUTO(90, 10, 9, 15, 16)
BFFT_GEO_AUTO(98, 14, 7, 7, 16)
BFFT_GEO_AUTO(108, 12, 9, 18, 16)
BFFT_GEO_AUTO(112, 14, 8, 14, 16)
BFFT_GEO_AUTO(126, 14, 9, 21, 16)
BFFT_GEO_AUTO(140, 10, 14, 14, 16)
BFFT_GEO_AUTO(150, 6, 25, 15, 10)
BFFT_GEO_AUTO(162, 18, 9, 9, 14)
BFFT_GEO_AUTO(168, 12, 14, 21, 16)
BFFT_GEO_AUTO(180, 18, 10, 18, 14)
BFFT_GEO_AUTO(196, 14, 14, 14, 16)
BFFT_GEO_AUTO(210, 10, 21, 21, 12)
BFFT_GEO_AUTO(250, 10, 25, 5, 10)
BFFT_GEO_AUTO(252, 18, 14, 21, 14)
BFFT_GEO_AUTO(270, 10, 27, 15, 9)
BFFT_GEO_AUTO(280, 20, 14, 20, 12)
BFFT_GEO_AUTO(294, 14, 21, 21, 12)
BFFT_GEO_AUTO(324, 12, 27, 18, 9)
BFFT_GEO_AUTO(350, 14, 25, 7, 10)
BFFT_GEO_AUTO(378, 14, 27, 21, 9)
BFFT_GEO_AUTO(392, 28, 14, 14, 9)
BFFT_GEO_AUTO(450, 18, 25, 15, 10)
BFFT_GEO_AUTO(486, 18, 27, 9, 9)
BFFT_GEO_AUTO(504, 18, 28, 21, 9)
#undef BFFT_GEO
#undef BFFT_GEO_AUTO
#undef BFFT_GEO_

} // namespace bfft
