// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use anything in oracle/.
//
// Self-contained mixed-radix FFT used (a) by the FFTW-API shim that lets the
// UNMODIFIED reference sources link here (FFTW 3 is not installed in this image;
// the reference needs exactly 9 fftwf_* symbols, see oracle/fftw_shim/fftw3.h)
// and (b) by the restated CPU oracle (oracle/bioem_oracle.cpp).
//
// Semantics follow the FFTW 3 manual for the calls the reference makes
// (param.cpp:924-935, bioem.cpp:1458,1848, map.cpp:585, param.cpp:1521):
// row-major n0 x n1 arrays, r2c output n0 x (n1/2+1), forward sign -1,
// backward sign +1, all transforms unnormalised.
//
// Algorithm: Stockham autosort, decimation in frequency, batched with the batch
// index contiguous (so the inner loops vectorise), specialised butterflies for
// radix 2/3/4/5/7 and an O(r^2) fallback for any other prime.  Real transforms of
// even length use the half-length complex trick.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace offt
{

template <typename T> struct Stage
{
  int r;               // radix
  int m;               // n_cur / r
  std::vector<T> twr;  // [m][r-1] twiddle exp(sign*2*pi*i*p*j/n_cur), j=1..r-1
  std::vector<T> twi;
};

template <typename T> struct Plan1D
{
  int n = 0;
  int sign = -1;
  std::vector<Stage<T>> stages;

  void init(int n_, int sign_)
  {
    n = n_;
    sign = sign_;
    stages.clear();
    int rem = n;
    std::vector<int> radices;
    while (rem % 4 == 0)
    {
      radices.push_back(4);
      rem /= 4;
    }
    const int small[] = {2, 3, 5, 7};
    for (int r : small)
      while (rem % r == 0)
      {
        radices.push_back(r);
        rem /= r;
      }
    for (int r = 11; rem > 1; r += 2)
      while (rem % r == 0)
      {
        radices.push_back(r);
        rem /= r;
      }
    int ncur = n;
    for (int r : radices)
    {
      Stage<T> st;
      st.r = r;
      st.m = ncur / r;
      st.twr.resize((size_t) st.m * (r - 1));
      st.twi.resize((size_t) st.m * (r - 1));
      for (int p = 0; p < st.m; p++)
        for (int j = 1; j < r; j++)
        {
          // reduce the angle exactly in integers before going to floating point
          long long num = ((long long) p * j) % ncur;
          double ang = sign * 2.0 * M_PI * (double) num / (double) ncur;
          st.twr[(size_t) p * (r - 1) + (j - 1)] = (T) cos(ang);
          st.twi[(size_t) p * (r - 1) + (j - 1)] = (T) sin(ang);
        }
      stages.push_back(st);
      ncur /= r;
    }
  }
};

// One Stockham pass.  x, y: SoA arrays of n*batch elements.  s = stride of the
// "already transformed" dimension times batch.
template <typename T>
static void stockham_pass(const Stage<T> &st, int sign, size_t s, const T *xr,
                          const T *xi, T *yr, T *yi)
{
  const int r = st.r, m = st.m;
  const T sg = (T) sign;
  for (int p = 0; p < m; p++)
  {
    const T *wr = &st.twr[(size_t) p * (r - 1)];
    const T *wi = &st.twi[(size_t) p * (r - 1)];
    const size_t ib = s * (size_t) p;
    const size_t is = s * (size_t) m;
    const size_t ob = s * (size_t) r * (size_t) p;
    if (r == 2)
    {
      const T w1r = wr[0], w1i = wi[0];
      for (size_t q = 0; q < s; q++)
      {
        const T ar = xr[ib + q], ai = xi[ib + q];
        const T br = xr[ib + is + q], bi = xi[ib + is + q];
        yr[ob + q] = ar + br;
        yi[ob + q] = ai + bi;
        const T dr = ar - br, di = ai - bi;
        yr[ob + s + q] = dr * w1r - di * w1i;
        yi[ob + s + q] = dr * w1i + di * w1r;
      }
    }
    else if (r == 4)
    {
      const T w1r = wr[0], w1i = wi[0], w2r = wr[1], w2i = wi[1], w3r = wr[2],
              w3i = wi[2];
      for (size_t q = 0; q < s; q++)
      {
        const T ar = xr[ib + q], ai = xi[ib + q];
        const T br = xr[ib + is + q], bi = xi[ib + is + q];
        const T cr = xr[ib + 2 * is + q], ci = xi[ib + 2 * is + q];
        const T dr = xr[ib + 3 * is + q], di = xi[ib + 3 * is + q];
        const T t0r = ar + cr, t0i = ai + ci;
        const T t1r = ar - cr, t1i = ai - ci;
        const T t2r = br + dr, t2i = bi + di;
        // (b - d) * (sign * i)
        const T t3r = -sg * (bi - di), t3i = sg * (br - dr);
        yr[ob + q] = t0r + t2r;
        yi[ob + q] = t0i + t2i;
        const T u1r = t1r + t3r, u1i = t1i + t3i;
        const T u2r = t0r - t2r, u2i = t0i - t2i;
        const T u3r = t1r - t3r, u3i = t1i - t3i;
        yr[ob + s + q] = u1r * w1r - u1i * w1i;
        yi[ob + s + q] = u1r * w1i + u1i * w1r;
        yr[ob + 2 * s + q] = u2r * w2r - u2i * w2i;
        yi[ob + 2 * s + q] = u2r * w2i + u2i * w2r;
        yr[ob + 3 * s + q] = u3r * w3r - u3i * w3i;
        yi[ob + 3 * s + q] = u3r * w3i + u3i * w3r;
      }
    }
    else if (r == 3)
    {
      const T c = (T) -0.5, sn = sg * (T) 0.86602540378443864676;
      for (size_t q = 0; q < s; q++)
      {
        const T ar = xr[ib + q], ai = xi[ib + q];
        const T br = xr[ib + is + q], bi = xi[ib + is + q];
        const T cr = xr[ib + 2 * is + q], ci = xi[ib + 2 * is + q];
        const T sr = br + cr, si = bi + ci;
        const T dr = br - cr, di = bi - ci;
        yr[ob + q] = ar + sr;
        yi[ob + q] = ai + si;
        const T mr = ar + c * sr, mi = ai + c * si;
        // i*sn*(b-c)
        const T er = -sn * di, ei = sn * dr;
        const T u1r = mr + er, u1i = mi + ei;
        const T u2r = mr - er, u2i = mi - ei;
        yr[ob + s + q] = u1r * wr[0] - u1i * wi[0];
        yi[ob + s + q] = u1r * wi[0] + u1i * wr[0];
        yr[ob + 2 * s + q] = u2r * wr[1] - u2i * wi[1];
        yi[ob + 2 * s + q] = u2r * wi[1] + u2i * wr[1];
      }
    }
    else
    {
      // generic small-prime butterfly (5, 7, 11, ...): direct O(r^2) DFT with
      // the symmetric pairing trick.
      const int h = (r - 1) / 2;
      T cs[32], sn[32];
      for (int k = 1; k <= h; k++)
      {
        cs[k] = (T) cos(2.0 * M_PI * k / r);
        sn[k] = (T)(sign * sin(2.0 * M_PI * k / r));
      }
      for (size_t q = 0; q < s; q++)
      {
        T ar[64], ai[64];
        for (int k = 0; k < r; k++)
        {
          ar[k] = xr[ib + (size_t) k * is + q];
          ai[k] = xi[ib + (size_t) k * is + q];
        }
        T pr[32], pi[32], mr[32], mi[32];
        T sumr = ar[0], sumi = ai[0];
        for (int k = 1; k <= h; k++)
        {
          pr[k] = ar[k] + ar[r - k];
          pi[k] = ai[k] + ai[r - k];
          mr[k] = ar[k] - ar[r - k];
          mi[k] = ai[k] - ai[r - k];
          sumr += pr[k];
          sumi += pi[k];
        }
        yr[ob + q] = sumr;
        yi[ob + q] = sumi;
        for (int j = 1; j <= h; j++)
        {
          T cr_ = ar[0], ci_ = ai[0], sr_ = 0, si_ = 0;
          for (int k = 1; k <= h; k++)
          {
            const int idx = (j * k) % r;
            const int kk = idx <= h ? idx : r - idx;
            const T sgn = idx <= h ? (T) 1 : (T) -1;
            cr_ += cs[kk] * pr[k];
            ci_ += cs[kk] * pi[k];
            sr_ += sgn * sn[kk] * mr[k];
            si_ += sgn * sn[kk] * mi[k];
          }
          // b_j = C + i*S ; b_{r-j} = C - i*S   (S carries the sign)
          const T u1r = cr_ - si_, u1i = ci_ + sr_;
          const T u2r = cr_ + si_, u2i = ci_ - sr_;
          yr[ob + (size_t) j * s + q] = u1r * wr[j - 1] - u1i * wi[j - 1];
          yi[ob + (size_t) j * s + q] = u1r * wi[j - 1] + u1i * wr[j - 1];
          const int j2 = r - j;
          yr[ob + (size_t) j2 * s + q] = u2r * wr[j2 - 1] - u2i * wi[j2 - 1];
          yi[ob + (size_t) j2 * s + q] = u2r * wi[j2 - 1] + u2i * wr[j2 - 1];
        }
      }
    }
  }
}

// Batched complex FFT of length plan.n over `batch` interleaved (batch-contiguous)
// sequences: element (k, b) lives at index k*batch + b.  Result ends up in
// (xr, xi); (wr, wi) is scratch of the same size.
template <typename T>
static void fft_batch(const Plan1D<T> &pl, size_t batch, T *xr, T *xi, T *wr,
                      T *wi)
{
  size_t s = batch;
  T *ar = xr, *ai = xi, *br = wr, *bi = wi;
  for (const Stage<T> &st : pl.stages)
  {
    stockham_pass(st, pl.sign, s, ar, ai, br, bi);
    s *= st.r;
    T *t;
    t = ar, ar = br, br = t;
    t = ai, ai = bi, bi = t;
  }
  if (ar != xr)
  {
    memcpy(xr, ar, sizeof(T) * pl.n * batch);
    memcpy(xi, ai, sizeof(T) * pl.n * batch);
  }
}

template <typename T> struct Plan2D
{
  int n0 = 0, n1 = 0, nc = 0;
  Plan1D<T> col_f, col_b;   // length n0
  Plan1D<T> half_f, half_b; // length n1/2 (even n1)
  Plan1D<T> row_f, row_b;   // length n1 (odd n1 or c2c)
  std::vector<T> hr, hi;    // exp(+2*pi*i*k/n1), k=0..n1/2

  void init(int n0_, int n1_)
  {
    n0 = n0_;
    n1 = n1_;
    nc = n1 / 2 + 1;
    col_f.init(n0, -1);
    col_b.init(n0, +1);
    row_f.init(n1, -1);
    row_b.init(n1, +1);
    if ((n1 & 1) == 0)
    {
      half_f.init(n1 / 2, -1);
      half_b.init(n1 / 2, +1);
    }
    hr.resize(nc);
    hi.resize(nc);
    for (int k = 0; k < nc; k++)
    {
      hr[k] = (T) cos(2.0 * M_PI * k / n1);
      hi[k] = (T) sin(2.0 * M_PI * k / n1);
    }
  }
};

template <typename T> struct Scratch
{
  std::vector<T> a, b, c, d;
  void need(size_t n)
  {
    if (a.size() < n)
    {
      a.resize(n);
      b.resize(n);
      c.resize(n);
      d.resize(n);
    }
  }
};

template <typename T> static Scratch<T> &tls_scratch()
{
  static thread_local Scratch<T> s;
  return s;
}

// in: n0 x nc interleaved complex (Tio), out: n0 x n1 real.  Unnormalised, sign +1.
template <typename T, typename Tio>
static void c2r_2d(const Plan2D<T> &pl, const Tio *in, Tio *out)
{
  const int n0 = pl.n0, n1 = pl.n1, nc = pl.nc;
  Scratch<T> &S = tls_scratch<T>();
  S.need((size_t) n0 * (size_t)(n1 + 2));
  T *xr = S.a.data(), *xi = S.b.data(), *wr = S.c.data(), *wi = S.d.data();
  for (size_t i = 0; i < (size_t) n0 * nc; i++)
  {
    xr[i] = (T) in[2 * i];
    xi[i] = (T) in[2 * i + 1];
  }
  // columns: length n0, batch nc (contiguous)
  fft_batch(pl.col_b, (size_t) nc, xr, xi, wr, wi);
  // transpose to [nc][n0]
  for (int i = 0; i < n0; i++)
    for (int k = 0; k < nc; k++)
    {
      wr[(size_t) k * n0 + i] = xr[(size_t) i * nc + k];
      wi[(size_t) k * n0 + i] = xi[(size_t) i * nc + k];
    }
  if ((n1 & 1) == 0)
  {
    const int h = n1 / 2;
    // Z[k] = (X[k] + conj X[h-k]) + i (X[k] - conj X[h-k]) e^{+2 pi i k/n1}
    // FFTW's c2r never reads the imaginary parts of the DC and Nyquist bins of the
    // halved (last) dimension: a half-complex sequence has none.  The reference feeds
    // spectra that are NOT exactly Hermitian along the first dimension (CTF table quirk
    // Q1), so this matters: drop them after the column transforms, as FFTW's
    // c2c-then-hc2r decomposition of a 2-D c2r does.
    for (int i = 0; i < n0; i++)
    {
      wi[(size_t) 0 * n0 + i] = 0;
      wi[(size_t) h * n0 + i] = 0;
    }
    for (int k = 0; k < h; k++)
    {
      const T cr = pl.hr[k], ci = pl.hi[k];
      const T *ar = &wr[(size_t) k * n0], *ai = &wi[(size_t) k * n0];
      const T *br = &wr[(size_t)(h - k) * n0], *bi = &wi[(size_t)(h - k) * n0];
      T *zr = &xr[(size_t) k * n0], *zi = &xi[(size_t) k * n0];
      for (int i = 0; i < n0; i++)
      {
        const T er = ar[i] + br[i], ei = ai[i] - bi[i];
        const T dr = ar[i] - br[i], di = ai[i] + bi[i];
        const T orr = dr * cr - di * ci, oi = dr * ci + di * cr;
        zr[i] = er - oi;
        zi[i] = ei + orr;
      }
    }
    fft_batch(pl.half_b, (size_t) n0, xr, xi, wr, wi);
    for (int j = 0; j < h; j++)
      for (int i = 0; i < n0; i++)
      {
        out[(size_t) i * n1 + 2 * j] = (Tio) xr[(size_t) j * n0 + i];
        out[(size_t) i * n1 + 2 * j + 1] = (Tio) xi[(size_t) j * n0 + i];
      }
  }
  else
  {
    // odd n1: Hermitian-extend and run a full complex transform (DC imaginary part
    // ignored, as above)
    for (int i = 0; i < n0; i++)
      wi[i] = 0;
    for (int k = 0; k < nc; k++)
      for (int i = 0; i < n0; i++)
      {
        xr[(size_t) k * n0 + i] = wr[(size_t) k * n0 + i];
        xi[(size_t) k * n0 + i] = wi[(size_t) k * n0 + i];
      }
    for (int k = nc; k < n1; k++)
      for (int i = 0; i < n0; i++)
      {
        xr[(size_t) k * n0 + i] = wr[(size_t)(n1 - k) * n0 + i];
        xi[(size_t) k * n0 + i] = -wi[(size_t)(n1 - k) * n0 + i];
      }
    fft_batch(pl.row_b, (size_t) n0, xr, xi, wr, wi);
    for (int j = 0; j < n1; j++)
      for (int i = 0; i < n0; i++)
        out[(size_t) i * n1 + j] = (Tio) xr[(size_t) j * n0 + i];
  }
}

// in: n0 x n1 real, out: n0 x nc interleaved complex.  Unnormalised, sign -1.
template <typename T, typename Tio>
static void r2c_2d(const Plan2D<T> &pl, const Tio *in, Tio *out)
{
  const int n0 = pl.n0, n1 = pl.n1, nc = pl.nc;
  Scratch<T> &S = tls_scratch<T>();
  S.need((size_t) n0 * (size_t)(n1 + 2));
  T *xr = S.a.data(), *xi = S.b.data(), *wr = S.c.data(), *wi = S.d.data();
  if ((n1 & 1) == 0)
  {
    const int h = n1 / 2;
    for (int i = 0; i < n0; i++)
      for (int j = 0; j < h; j++)
      {
        xr[(size_t) j * n0 + i] = (T) in[(size_t) i * n1 + 2 * j];
        xi[(size_t) j * n0 + i] = (T) in[(size_t) i * n1 + 2 * j + 1];
      }
    fft_batch(pl.half_f, (size_t) n0, xr, xi, wr, wi);
    // X[k] = (Z[k] + conj Z[h-k])/2 - (i/2) e^{-2 pi i k/n1} (Z[k] - conj Z[h-k])
    for (int k = 0; k <= h; k++)
    {
      const T cr = pl.hr[k], ci = -pl.hi[k];
      const int ka = k % h, kb = (h - k) % h;
      const T *ar = &xr[(size_t) ka * n0], *ai = &xi[(size_t) ka * n0];
      const T *br = &xr[(size_t) kb * n0], *bi = &xi[(size_t) kb * n0];
      T *yr = &wr[(size_t) k * n0], *yi = &wi[(size_t) k * n0];
      for (int i = 0; i < n0; i++)
      {
        const T er = ar[i] + br[i], ei = ai[i] - bi[i];
        const T dr = ar[i] - br[i], di = ai[i] + bi[i];
        const T tr = dr * cr - di * ci, ti = dr * ci + di * cr;
        // -i * t = (ti, -tr)
        yr[i] = (T) 0.5 * (er + ti);
        yi[i] = (T) 0.5 * (ei - tr);
      }
    }
  }
  else
  {
    for (int i = 0; i < n0; i++)
      for (int j = 0; j < n1; j++)
      {
        xr[(size_t) j * n0 + i] = (T) in[(size_t) i * n1 + j];
        xi[(size_t) j * n0 + i] = 0;
      }
    fft_batch(pl.row_f, (size_t) n0, xr, xi, wr, wi);
    for (size_t i = 0; i < (size_t) nc * n0; i++)
    {
      wr[i] = xr[i];
      wi[i] = xi[i];
    }
  }
  // transpose [nc][n0] -> [n0][nc]
  for (int k = 0; k < nc; k++)
    for (int i = 0; i < n0; i++)
    {
      xr[(size_t) i * nc + k] = wr[(size_t) k * n0 + i];
      xi[(size_t) i * nc + k] = wi[(size_t) k * n0 + i];
    }
  fft_batch(pl.col_f, (size_t) nc, xr, xi, wr, wi);
  for (size_t i = 0; i < (size_t) n0 * nc; i++)
  {
    out[2 * i] = (Tio) xr[i];
    out[2 * i + 1] = (Tio) xi[i];
  }
}

// in/out: n0 x n1 interleaved complex.
template <typename T, typename Tio>
static void c2c_2d(const Plan2D<T> &pl, int sign, const Tio *in, Tio *out)
{
  const int n0 = pl.n0, n1 = pl.n1;
  Scratch<T> &S = tls_scratch<T>();
  S.need((size_t) n0 * (size_t)(n1 + 2));
  T *xr = S.a.data(), *xi = S.b.data(), *wr = S.c.data(), *wi = S.d.data();
  for (size_t i = 0; i < (size_t) n0 * n1; i++)
  {
    xr[i] = (T) in[2 * i];
    xi[i] = (T) in[2 * i + 1];
  }
  fft_batch(sign < 0 ? pl.col_f : pl.col_b, (size_t) n1, xr, xi, wr, wi);
  for (int i = 0; i < n0; i++)
    for (int j = 0; j < n1; j++)
    {
      wr[(size_t) j * n0 + i] = xr[(size_t) i * n1 + j];
      wi[(size_t) j * n0 + i] = xi[(size_t) i * n1 + j];
    }
  fft_batch(sign < 0 ? pl.row_f : pl.row_b, (size_t) n0, wr, wi, xr, xi);
  for (int i = 0; i < n0; i++)
    for (int j = 0; j < n1; j++)
    {
      out[2 * ((size_t) i * n1 + j)] = (Tio) wr[(size_t) j * n0 + i];
      out[2 * ((size_t) i * n1 + j) + 1] = (Tio) wi[(size_t) j * n0 + i];
    }
}

} // namespace offt
