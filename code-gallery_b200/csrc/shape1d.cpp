// 1-D tables of FE_DGQ(p) x QGauss(nq) computed in long double (the deal.II
// classes FE_DGQ / QGauss / QGaussLobatto are what the reference instantiates at NS.cc:2218-2219,
// 2320-2325).  Besides the shape matrices the kernels need the Kronecker-factored 1-D operators
// M^-1 K, M^-1 f_s, M^-1 g_s used by the Cartesian fast path (DESIGN.md "Kronecker form").
#include <cmath>
#include <map>
#include <mutex>

#include "dgx_internal.h"

namespace dgx {
namespace {
typedef long double ld;

void legendre(int n, ld t, ld &p, ld &d1, ld &d2) {
  if (n == 0) { p = 1; d1 = 0; d2 = 0; return; }
  ld p0 = 1, p1 = t;
  for (int k = 1; k < n; ++k) {
    ld pn = ((2 * k + 1) * t * p1 - k * p0) / ld(k + 1);
    p0 = p1; p1 = pn;
  }
  p = p1;
  d1 = n * (t * p1 - p0) / (t * t - 1);
  d2 = (2 * t * d1 - ld(n) * (n + 1) * p1) / (1 - t * t);
}

void gauss_legendre(int nq, std::vector<ld> &x, std::vector<ld> &w) {
  const ld pi = acosl(-1.0L);
  x.resize(nq); w.resize(nq);
  for (int k = 1; k <= nq; ++k) {
    ld t = -cosl((4 * k - 1) * pi / (4 * nq + 2));
    for (int it = 0; it < 100; ++it) {
      ld p, d1, d2; legendre(nq, t, p, d1, d2);
      ld dt = p / d1; t -= dt;
      if (fabsl(dt) < 1e-19L) break;
    }
    ld p, d1, d2; legendre(nq, t, p, d1, d2);
    x[k - 1] = (t + 1) / 2;
    w[k - 1] = 1 / ((1 - t * t) * d1 * d1);
  }
  for (int k = 0; k < nq / 2; ++k) {   // exact symmetry
    ld xs = (x[k] + (1 - x[nq - 1 - k])) / 2, ws = (w[k] + w[nq - 1 - k]) / 2;
    x[k] = xs; x[nq - 1 - k] = 1 - xs; w[k] = w[nq - 1 - k] = ws;
  }
  if (nq % 2) x[nq / 2] = 0.5L;
}

void gauss_lobatto(int n, std::vector<ld> &x) {
  const ld pi = acosl(-1.0L);
  x.assign(n, 0);
  if (n == 1) { x[0] = 0.5L; return; }
  x[0] = 0; x[n - 1] = 1;
  const int p = n - 1;
  for (int k = 1; k <= n - 2; ++k) {
    ld t = -cosl(k * pi / p);
    for (int it = 0; it < 100; ++it) {
      ld pv, d1, d2; legendre(p, t, pv, d1, d2);
      ld dt = d1 / d2; t -= dt;
      if (fabsl(dt) < 1e-19L) break;
    }
    x[k] = (t + 1) / 2;
  }
  for (int k = 0; k < n / 2; ++k) {
    ld xs = (x[k] + (1 - x[n - 1 - k])) / 2;
    x[k] = xs; x[n - 1 - k] = 1 - xs;
  }
  if (n % 2) x[n / 2] = 0.5L;
  x[0] = 0; x[n - 1] = 1;
}

void lagrange(const std::vector<ld> &nodes, ld x, std::vector<ld> &L, std::vector<ld> &D) {
  const int n = (int)nodes.size();
  L.assign(n, 0); D.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    ld denom = 1, val = 1, der = 0;
    for (int l = 0; l < n; ++l) if (l != i) { denom *= nodes[i] - nodes[l]; val *= x - nodes[l]; }
    for (int m = 0; m < n; ++m) {
      if (m == i) continue;
      ld prod = 1;
      for (int l = 0; l < n; ++l) if (l != i && l != m) prod *= x - nodes[l];
      der += prod;
    }
    L[i] = val / denom; D[i] = der / denom;
  }
}

// dense inverse by Gauss-Jordan with partial pivoting (n <= 9, SPD input)
void invert(int n, std::vector<ld> A, std::vector<ld> &inv) {
  inv.assign(n * n, 0);
  for (int i = 0; i < n; ++i) inv[i * n + i] = 1;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (fabsl(A[r * n + c]) > fabsl(A[piv * n + c])) piv = r;
    if (piv != c) for (int k = 0; k < n; ++k) { std::swap(A[c * n + k], A[piv * n + k]); std::swap(inv[c * n + k], inv[piv * n + k]); }
    ld d = 1 / A[c * n + c];
    for (int k = 0; k < n; ++k) { A[c * n + k] *= d; inv[c * n + k] *= d; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      ld f = A[r * n + c];
      if (f == 0) continue;
      for (int k = 0; k < n; ++k) { A[r * n + k] -= f * A[c * n + k]; inv[r * n + k] -= f * inv[c * n + k]; }
    }
  }
}

std::vector<double> to_d(const std::vector<ld> &v) { return std::vector<double>(v.begin(), v.end()); }

Shape1D build(int degree, int nq) {
  Shape1D s;
  const int n = degree + 1;
  s.n = n; s.nq = nq;
  std::vector<ld> nodes, xq, wq;
  gauss_lobatto(n, nodes);
  gauss_legendre(nq, xq, wq);
  std::vector<ld> S(nq * n), D(nq * n), L, Dv;
  for (int q = 0; q < nq; ++q) {
    lagrange(nodes, xq[q], L, Dv);
    for (int i = 0; i < n; ++i) { S[q * n + i] = L[i]; D[q * n + i] = Dv[i]; }
  }
  std::vector<ld> f[2], g[2];
  for (int sd = 0; sd < 2; ++sd) lagrange(nodes, ld(sd), f[sd], g[sd]);
  std::vector<ld> M(n * n, 0), K(n * n, 0), Minv, Kt(n * n, 0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      for (int q = 0; q < nq; ++q) {
        M[i * n + j] += wq[q] * S[q * n + i] * S[q * n + j];
        K[i * n + j] += wq[q] * D[q * n + i] * D[q * n + j];
      }
  invert(n, M, Minv);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      for (int k = 0; k < n; ++k) Kt[i * n + j] += Minv[i * n + k] * K[k * n + j];
  std::vector<ld> a[2], b[2];
  for (int sd = 0; sd < 2; ++sd) {
    a[sd].assign(n, 0); b[sd].assign(n, 0);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k) { a[sd][i] += Minv[i * n + k] * f[sd][k]; b[sd][i] += Minv[i * n + k] * g[sd][k]; }
  }
  s.nodes = to_d(nodes); s.xq = to_d(xq); s.wq = to_d(wq);
  s.S = to_d(S); s.D = to_d(D); s.M = to_d(M); s.K = to_d(K); s.Minv = to_d(Minv); s.Kt = to_d(Kt);
  for (int sd = 0; sd < 2; ++sd) { s.f[sd] = to_d(f[sd]); s.g[sd] = to_d(g[sd]); s.a[sd] = to_d(a[sd]); s.b[sd] = to_d(b[sd]); }
  if (nq == n) {
    // change of basis to the Lagrange polynomials on the Gauss points: u_hat = S u
    std::vector<ld> Sinv, Dh(n * n, 0);
    invert(n, S, Sinv);
    for (int q = 0; q < n; ++q)
      for (int r = 0; r < n; ++r)
        for (int i = 0; i < n; ++i) Dh[q * n + r] += D[q * n + i] * Sinv[i * n + r];
    s.Dhat = to_d(Dh);
    for (int sd = 0; sd < 2; ++sd) {
      std::vector<ld> fh(n, 0), gh(n, 0);
      for (int r = 0; r < n; ++r)
        for (int i = 0; i < n; ++i) { fh[r] += f[sd][i] * Sinv[i * n + r]; gh[r] += g[sd][i] * Sinv[i * n + r]; }
      s.fhat[sd] = to_d(fh); s.ghat[sd] = to_d(gh);
    }
  }
  for (int c = 0; c < 2; ++c) {
    std::vector<ld> P(n * n);
    for (int j = 0; j < n; ++j) {
      lagrange(nodes, (nodes[j] + c) / 2, L, Dv);
      for (int i = 0; i < n; ++i) P[j * n + i] = L[i];
    }
    s.P[c] = to_d(P);
  }
  return s;
}
}  // namespace

const Shape1D &get_shape(int degree, int nq) {
  static std::map<std::pair<int, int>, Shape1D> cache;
  static std::mutex mtx;
  std::lock_guard<std::mutex> lock(mtx);
  auto key = std::make_pair(degree, nq);
  auto it = cache.find(key);
  if (it == cache.end()) it = cache.emplace(key, build(degree, nq)).first;
  return it->second;
}

double gamma_trbdf2() { return 2.0 - std::sqrt(2.0); }

}  // namespace dgx
