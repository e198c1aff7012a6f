// Right-hand-side linear forms of NavierStokesProjectionOperator:
//   vmult_rhs_velocity      NS.cc:788-798   (assemble_rhs_{cell,face,boundary}_term_velocity, NS.cc:462-781)
//   vmult_rhs_pressure      NS.cc:915-925   (assemble_rhs_{cell,face,boundary}_term_pressure, NS.cc:806-908)
//   vmult_grad_p_projection NS.cc:1363-1366 (assemble_rhs_cell_term_projection_grad_p, NS.cc:1336-1356)
// Same construction as velocity_op.cu: every field of a cell is moved to the Lagrange basis on the Gauss points of
// the form's quadrature (QGauss(p+2) for the velocity forms, QGauss(p+1) for the pressure form, NS.cc:2324-2325;
// the pressure FE is evaluated on the velocity quadrature and vice versa, NS.cc:474, 814), cell and face terms are
// accumulated per direction by the thread owning the line, each cell computes its own side of every face in the
// outward-normal form, and the result is moved back with the transposed basis change.  zero_dst_vector = true.
#include "pressure_op.cuh"

namespace dgx {
namespace {

constexpr int kMaxN = MAX_N;

struct CellGeom {          // cell index -> integer coordinates (inflow data needs the physical y of a face point)
  int dim, L, c0[3];
  long long first_cell;
  double lo[3], h[3];
};

__device__ __forceinline__ unsigned compact1by2(unsigned x) {
  x &= 0x09249249u;
  x = (x ^ (x >> 2)) & 0x030c30c3u;
  x = (x ^ (x >> 4)) & 0x0300f00fu;
  x = (x ^ (x >> 8)) & 0x030000ffu;
  x = (x ^ (x >> 16)) & 0x000003ffu;
  return x;
}
__device__ __forceinline__ unsigned compact1by1(unsigned x) {
  x &= 0x55555555u;
  x = (x ^ (x >> 1)) & 0x33333333u;
  x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
  x = (x ^ (x >> 4)) & 0x00ff00ffu;
  x = (x ^ (x >> 8)) & 0x0000ffffu;
  return x;
}
__device__ __forceinline__ int cell_coord(const CellGeom &G, long long local_cell, int d) {
  const long long g = G.first_cell + local_cell;
  const long long per = 1LL << (G.dim * G.L);
  long long coarse = g / per;
  const unsigned loc = (unsigned)(g % per);
  int cc[3];
  for (int e = 0; e < G.dim; ++e) { cc[e] = (int)(coarse % G.c0[e]); coarse /= G.c0[e]; }
  const unsigned l = G.dim == 3 ? compact1by2(loc >> d) : compact1by1(loc >> d);
  return (cc[d] << G.L) + (int)l;
}

// EquationData::Velocity::value (equation_data.h:47-57)
template <typename T> __device__ __forceinline__ T inflow_u0(T y) {
  const T Um = T(1.5), H = T(4.1);
  return T(4.0) * Um * y * (H - y) / (H * H);
}

template <int dim, int n>
__device__ __forceinline__ void bline(int d, int l, int &base, int &stride) {   // line `l` along d in an n^dim box
  if (dim == 3) {
    if (d == 0) { base = l * n; stride = 1; }
    else if (d == 1) { base = (l % n) + n * n * (l / n); stride = n; }
    else { base = l; stride = n * n; }
  } else {
    if (d == 0) { base = l * n; stride = 1; }
    else { base = l; stride = n; }
  }
}

// in-place 1-D basis change along d of a field stored in an NM^dim box: `nin` nodal values -> `nout` values per line
// (M[nout][nin]).  ext[e] = current extent of direction e.  One thread per line of the box.
template <int dim, int NM, typename T, bool TR = false>
__device__ __forceinline__ void sweep_box(T *box, int d, const T *M, int nin, int nout, const int *ext, int l) {
  const int a = l % NM, b = l / NM;
  int oa, ob;   // the two other directions in increasing order
  if (d == 0) { oa = 1; ob = 2; } else if (d == 1) { oa = 0; ob = 2; } else { oa = 0; ob = 1; }
  if (a >= ext[oa] || (dim == 3 && b >= ext[ob])) return;
  int base, stride;
  bline<dim, NM>(d, l, base, stride);
  T u[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) u[i] = i < nin ? box[base + stride * i] : T(0);
#pragma unroll
  for (int q = 0; q < NM; ++q) {
    if (q >= nout) break;
    T acc = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) if (i < nin) acc += (TR ? M[i * nout + q] : M[q * nin + i]) * u[i];
    box[base + stride * q] = acc;
  }
}

template <typename T, int n>   // n = number of Gauss points of the form = nodes of the OUTPUT space
struct FormTables {
  T S[n * n];            // square: output-space nodal -> Gauss points
  T Dh[n * n];           // collocation derivative on the Gauss points
  T fh[2][n], gh[2][n];  // end-point value / derivative of the Gauss-point basis
  T w[n], xq[n];
  T Sx[n * (n + 1)];     // other space -> Gauss points: [n][nx], nx = n - 1 (pressure in velocity forms) or n + 1
  T gN[2][n];            // l_i'(s) of the output space's nodal basis (neighbour traces)
  T h[MAX_DIM], ih[MAX_DIM];
};

template <typename T, int n>
void fill_tables(FormTables<T, n> &F, const dgx_mf *mf, int other_degree) {
  const Shape1D &s = get_shape(n - 1, n);
  for (int i = 0; i < n * n; ++i) { F.S[i] = (T)s.S[i]; F.Dh[i] = (T)s.Dhat[i]; }
  for (int sd = 0; sd < 2; ++sd)
    for (int i = 0; i < n; ++i) { F.fh[sd][i] = (T)s.fhat[sd][i]; F.gh[sd][i] = (T)s.ghat[sd][i]; F.gN[sd][i] = (T)s.g[sd][i]; }
  for (int i = 0; i < n; ++i) { F.w[i] = (T)s.wq[i]; F.xq[i] = (T)s.xq[i]; }
  const Shape1D &x = get_shape(other_degree, n);
  const int nx = other_degree + 1;
  for (int i = 0; i < n * nx; ++i) F.Sx[i] = (T)x.S[i];
  for (int d = 0; d < MAX_DIM; ++d) { const double h = d < mf->dim ? mf->h[d] : 1.0; F.h[d] = (T)h; F.ih[d] = (T)(1.0 / h); }
}

struct VelRhsConst { double alpha, a1, a2, a1Re, a2Re, ab, abRe, C, theta; int stage; };

// ------------------------------------------------------------------------------------------ velocity rhs
template <int dim, int n>
__host__ __device__ constexpr int rhs_cpb() { return (64 / ipow(n, dim - 1)) > 0 ? (64 / ipow(n, dim - 1)) : 1; }

// Work decomposition as in velocity_op.cu: a 256-thread block is cut into independent groups of WPG warps that own CPG
// cells and all their shared memory, run every phase as a group-wide loop over work items and synchronise only with
// themselves.  Shared-memory boxes pad every x-row to n + 1 entries.
constexpr int kRhsThreads = 256;
constexpr int kRhsSmemBudget = 75 * 1024;

template <int dim, int n, typename T>
struct VelRhsCfg {
  static constexpr int NL = ipow(n, dim - 1), ND = NL * n;
  static constexpr int NFLD = 3 * dim + 1;               // u_n, second velocity, p, u_e (stage-2 boundaries)
  static constexpr int NF = 4 * dim + 1;                 // neighbour face fields: V0, G0, V1, G1 (dim each), Vp
  static constexpr int RS = n + 1, NB = RS * NL, NR = NL / n, NLF = RS * NR;
  static constexpr int PER_CELL = (NFLD + dim) * NB + 2 * NF * NLF;
  static constexpr int fit = kRhsSmemBudget / (PER_CELL * (int)sizeof(T));
  static constexpr int CPBraw = fit < 1 ? 1 : (fit > 16 ? 16 : fit);
  static constexpr int WPG = CPBraw >= 8 ? 1 : (CPBraw >= 4 ? 2 : (CPBraw >= 2 ? 4 : 8));
  static constexpr int NG = (kRhsThreads / 32) / WPG;
  static constexpr int CPG = CPBraw >= 8 ? CPBraw / 8 : 1;
  static constexpr int CPB = NG * CPG;
  static constexpr size_t smem = sizeof(T) * (size_t)CPB * PER_CELL;
};

template <int dim, int n>
__device__ __forceinline__ void rline(int d, int l, int &base, int &stride) {   // line `l` along d in a padded box
  constexpr int RS = n + 1;
  if (dim == 3) {
    if (d == 0) { base = l * RS; stride = 1; }
    else if (d == 1) { base = (l % n) + RS * n * (l / n); stride = RS; }
    else { base = (l % n) + RS * (l / n); stride = RS * n; }
  } else {
    if (d == 0) { base = l * RS; stride = 1; }
    else { base = l; stride = RS; }
  }
}

template <int WPG>
__device__ __forceinline__ void rhs_group_sync(int g) {
  if (WPG == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(WPG * 32) : "memory");
}

template <int dim, int n, typename T>
__global__ void __launch_bounds__(kRhsThreads, 3)
rhs_velocity_kernel(const __grid_constant__ FormTables<T, n> F, const __grid_constant__ VelRhsConst K, const __grid_constant__ CellGeom CG,
                    const int *__restrict__ nbr, int n_owned, const T *__restrict__ s0, const T *__restrict__ s1, const T *__restrict__ sp,
                    const T *__restrict__ s3, T *__restrict__ dst) {
  using Cfg = VelRhsCfg<dim, n, T>;
  constexpr int NL = Cfg::NL, ND = Cfg::ND, NFLD = Cfg::NFLD, NF = Cfg::NF, RS = Cfg::RS, NB = Cfg::NB, NR = Cfg::NR, NLF = Cfg::NLF;
  constexpr int PER_CELL = Cfg::PER_CELL, WPG = Cfg::WPG, CPG = Cfg::CPG, GT = WPG * 32;
  constexpr int np = n - 1, NDP = ipow(np, dim), NLP = ipow(np, dim - 1);
  extern __shared__ __align__(16) unsigned char rsm[];
  const int g = threadIdx.x / GT, tid = threadIdx.x - g * GT;
  const int cell0 = (blockIdx.x * Cfg::NG + g) * CPG;
  const int ncell = min(CPG, n_owned - cell0);
  if (ncell <= 0) return;
  T *sm = reinterpret_cast<T *>(rsm) + (size_t)g * CPG * PER_CELL;
  auto Ac = [&](int ci) { return sm + (size_t)ci * PER_CELL; };                  // [NFLD][NB]
  auto Rc = [&](int ci) { return sm + (size_t)ci * PER_CELL + NFLD * NB; };      // [dim][NB]
  auto Na = [&](int ci) { return sm + (size_t)ci * PER_CELL + (NFLD + dim) * NB; };   // [2 sides][NF][NLF]
  const int stage = K.stage;
  const int nfld = stage == 1 ? 2 * dim + 1 : NFLD;   // field slots: [0,dim) u_n ; [dim,2dim) second velocity ; 2dim: p ; then u_e (stage 2)
  // neighbour face-field slots
  auto iV0 = [](int c) { return c; };
  auto iG0 = [](int c) { return dim + c; };
  auto iV1 = [](int c) { return 2 * dim + c; };
  auto iG1 = [](int c) { return 3 * dim + c; };
  constexpr int iVp = 4 * dim;
  // 1. load one x-row per item and move it to the Gauss points of QGauss(n) on the way in (pressure: np nodes -> n points)
  for (int it = tid; it < ncell * nfld * NL; it += GT) {
    const int ci = CPG == 1 ? 0 : it / (nfld * NL), r = it - ci * (nfld * NL), f = r / NL, row = r - f * NL;
    const int cell = cell0 + ci;
    T *A = Ac(ci) + f * NB + row * RS;
    if (f == 2 * dim) {
      const int j = row % n, k = row / n;
      if (j >= np || (dim == 3 && k >= np)) continue;
      const T *gp = sp + (size_t)cell * NDP + (dim == 3 ? (j + np * k) * np : j * np);
      T u[np > 0 ? np : 1];
#pragma unroll
      for (int i = 0; i < np; ++i) u[i] = gp[i];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < np; ++i) acc += F.Sx[q * np + i] * u[i];
        A[q] = acc;
      }
    } else {
      const T *src = f < dim ? s0 : (f < 2 * dim ? s1 : s3);
      const int comp = f < dim ? f : (f < 2 * dim ? f - dim : f - 2 * dim - 1);
      T u[n];
      load_row<T, n>(src + ((size_t)cell * dim + comp) * ND + row * n, u);
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) acc += F.S[q * n + i] * u[i];
        A[q] = acc;
      }
    }
  }
  rhs_group_sync<WPG>(g);
  // 2. remaining directions of the basis change (the pressure box grows from np to n entries along d)
#pragma unroll
  for (int d = 1; d < dim; ++d) {
    for (int it = tid; it < ncell * nfld * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (nfld * NL), r = it - ci * (nfld * NL), f = r / NL, l = r - f * NL;
      int base, stride;
      rline<dim, n>(d, l, base, stride);
      T *A = Ac(ci) + f * NB;
      if (f == 2 * dim) {
        // lines along d: the directions below d are complete (n entries), the ones above still hold np
        const int a2 = l % n, b2 = l / n;
        if (dim == 3 && d == 1 && b2 >= np) continue;   // y pass: z still has np layers
        (void)a2;
        T u[np > 0 ? np : 1];
#pragma unroll
        for (int i = 0; i < np; ++i) u[i] = A[base + stride * i];
#pragma unroll
        for (int q = 0; q < n; ++q) {
          T acc = 0;
#pragma unroll
          for (int i = 0; i < np; ++i) acc += F.Sx[q * np + i] * u[i];
          A[base + stride * q] = acc;
        }
      } else {
        T u[n];
#pragma unroll
        for (int i = 0; i < n; ++i) u[i] = A[base + stride * i];
#pragma unroll
        for (int q = 0; q < n; ++q) {
          T acc = 0;
#pragma unroll
          for (int i = 0; i < n; ++i) acc += F.S[q * n + i] * u[i];
          A[base + stride * q] = acc;
        }
      }
    }
    rhs_group_sync<WPG>(g);
  }
  T vol = 1;
#pragma unroll
  for (int d = 0; d < dim; ++d) vol *= F.h[d];
  // 3. per direction: cell term (mass part with d == 0) and the two faces normal to d
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    // 3a. nodal traces of both neighbours: item = (cell, component, nodal tangential point); component 0 also fetches the
    //     pressure layer on the np^(dim-1) tangential grid
    for (int it = tid; it < ncell * dim * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, l = r - c * NL;
      const int cell = cell0 + ci;
      int base, stride;
      bline<dim, n>(d, l, base, stride);
      const int t0 = l % n, t1 = l / n, lf = t1 * RS + t0;
      T *NTa = Na(ci);
      int nb[2];
      T x0[2][n], x1[2][n], xp[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) nb[s] = nbr[(size_t)(2 * d + s) * n_owned + cell];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const size_t off = ((size_t)max(nb[s], 0) * dim + c) * ND + base;   // boundary sides read cell 0 and drop it
        if (d == 0) { load_row<T, n>(s0 + off, x0[s]); load_row<T, n>(s1 + off, x1[s]); }
        else {
#pragma unroll
          for (int i = 0; i < n; ++i) { x0[s][i] = s0[off + stride * i]; x1[s][i] = s1[off + stride * i]; }
        }
        xp[s] = 0;
        if (c == 0 && t0 < np && (dim == 2 || t1 < np)) {
          int pb, ps;
          bline<dim, np>(d, dim == 3 ? t0 + np * t1 : t0, pb, ps);
          xp[s] = sp[(size_t)max(nb[s], 0) * NDP + pb + ps * ((1 - s) * (np - 1))];
        }
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (nb[s] < 0) continue;
        T d0 = 0, d1 = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) { d0 += F.gN[1 - s][i] * x0[s][i]; d1 += F.gN[1 - s][i] * x1[s][i]; }
        const int e = (1 - s) * (n - 1);
        NTa[(s * NF + iV0(c)) * NLF + lf] = x0[s][e];
        NTa[(s * NF + iG0(c)) * NLF + lf] = F.ih[d] * d0;
        NTa[(s * NF + iV1(c)) * NLF + lf] = x1[s][e];
        NTa[(s * NF + iG1(c)) * NLF + lf] = F.ih[d] * d1;
        if (c == 0 && t0 < np && (dim == 2 || t1 < np)) NTa[(s * NF + iVp) * NLF + lf] = xp[s];
      }
    }
    rhs_group_sync<WPG>(g);
    // tangential interpolation to the face Gauss points, in place: rows, then (3-D) columns
    for (int it = tid; it < ncell * 2 * NF * NR; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (2 * NF * NR), r = it - ci * (2 * NF * NR), sf = r / NR, t1 = r - sf * NR;
      const int f = sf % NF;
      T *io = Na(ci) + sf * NLF + RS * t1;
      if (f == iVp) {
        if (dim == 3 && t1 >= np) continue;
        T x[np > 0 ? np : 1];
#pragma unroll
        for (int a2 = 0; a2 < np; ++a2) x[a2] = io[a2];
#pragma unroll
        for (int t0 = 0; t0 < n; ++t0) {
          T acc = 0;
#pragma unroll
          for (int a2 = 0; a2 < np; ++a2) acc += F.Sx[t0 * np + a2] * x[a2];
          io[t0] = acc;
        }
      } else {
        T x[n];
#pragma unroll
        for (int a2 = 0; a2 < n; ++a2) x[a2] = io[a2];
#pragma unroll
        for (int t0 = 0; t0 < n; ++t0) {
          T acc = 0;
#pragma unroll
          for (int a2 = 0; a2 < n; ++a2) acc += F.S[t0 * n + a2] * x[a2];
          io[t0] = acc;
        }
      }
    }
    rhs_group_sync<WPG>(g);
    if (dim == 3) {
      for (int it = tid; it < ncell * 2 * NF * n; it += GT) {
        const int ci = CPG == 1 ? 0 : it / (2 * NF * n), r = it - ci * (2 * NF * n), sf = r / n, t0 = r - sf * n;
        const int f = sf % NF;
        T *io = Na(ci) + sf * NLF + t0;
        if (f == iVp) {
          T x[np > 0 ? np : 1];
#pragma unroll
          for (int b2 = 0; b2 < np; ++b2) x[b2] = io[RS * b2];
#pragma unroll
          for (int t1 = 0; t1 < n; ++t1) {
            T acc = 0;
#pragma unroll
            for (int b2 = 0; b2 < np; ++b2) acc += F.Sx[t1 * np + b2] * x[b2];
            io[RS * t1] = acc;
          }
        } else {
          T x[n];
#pragma unroll
          for (int b2 = 0; b2 < n; ++b2) x[b2] = io[RS * b2];
#pragma unroll
          for (int t1 = 0; t1 < n; ++t1) {
            T acc = 0;
#pragma unroll
            for (int b2 = 0; b2 < n; ++b2) acc += F.S[t1 * n + b2] * x[b2];
            io[RS * t1] = acc;
          }
        }
      }
      rhs_group_sync<WPG>(g);
    }
    // 3b. item = (cell, component, line along d)
    for (int it = tid; it < ncell * dim * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, l = r - c * NL;
      const int cell = cell0 + ci;
      const T *A = Ac(ci);
      T *R = Rc(ci);
      const T *NTf = Na(ci);
      int base, stride;
      rline<dim, n>(d, l, base, stride);
      const int t0 = l % n, t1 = l / n, lf = t1 * RS + t0;
      const T wt = dim == 3 ? F.w[t0] * F.w[t1] : F.w[t0];
      T area = 1;
#pragma unroll
      for (int e2 = 0; e2 < dim; ++e2) if (e2 != d) area *= F.h[e2];
      const T wf = area * wt;
      // this line of the d-components (advecting velocities), of the pressure and of the own component
      T a0d[n], a1d[n], apd[n], u0[n], u1[n];
      T W0[2] = {0, 0}, W1[2] = {0, 0}, W3[2] = {0, 0}, Pv[2] = {0, 0};
#pragma unroll
      for (int q = 0; q < n; ++q) {
        a0d[q] = A[d * NB + base + stride * q]; a1d[q] = A[(dim + d) * NB + base + stride * q]; apd[q] = A[2 * dim * NB + base + stride * q];
        const T x3 = stage == 1 ? T(0) : A[(2 * dim + 1 + d) * NB + base + stride * q];
        u0[q] = A[c * NB + base + stride * q]; u1[q] = A[(dim + c) * NB + base + stride * q];
#pragma unroll
        for (int s = 0; s < 2; ++s) { W0[s] += F.fh[s][q] * a0d[q]; W1[s] += F.fh[s][q] * a1d[q]; W3[s] += F.fh[s][q] * x3; Pv[s] += F.fh[s][q] * apd[q]; }
      }
      // cell term, direction d
      T fl[n], rl[n];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T d0 = 0, d1 = 0;
#pragma unroll
        for (int q2 = 0; q2 < n; ++q2) { d0 += F.Dh[q * n + q2] * u0[q2]; d1 += F.Dh[q * n + q2] * u1[q2]; }
        d0 *= F.ih[d]; d1 *= F.ih[d];
        T f;
        if (stage == 1) f = -(T)K.a1Re * d0 + (T)K.a1 * u0[q] * a1d[q];                                           // NS.cc:497-507
        else f = (T)K.a2 * u1[q] * a1d[q] + (T)K.a1 * u0[q] * a0d[q] - (T)K.a2Re * d1 - (T)K.a1Re * d0;             // NS.cc:538-546
        if (c == d) f += apd[q];
        fl[q] = vol * wt * F.w[q] * f;
      }
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) acc += F.Dh[q * n + i] * fl[q];
        rl[i] = F.ih[d] * acc;
        if (d == 0) rl[i] += (T)K.alpha * vol * wt * F.w[i] * (stage == 1 ? u0[i] : u1[i]);   // mass part: alpha JxW (u_n | u_n_gamma)
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int nbs = nbr[(size_t)(2 * d + s) * n_owned + cell];
        const T ns = s ? T(1) : T(-1);
        T V0 = 0, G0 = 0, V1 = 0, G1 = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) { V0 += F.fh[s][q] * u0[q]; G0 += F.gh[s][q] * u0[q]; V1 += F.fh[s][q] * u1[q]; G1 += F.gh[s][q] * u1[q]; }
        G0 *= F.ih[d]; G1 *= F.ih[d];
        T val, nd = 0;
        if (nbs >= 0) {
          const T V0n = NTf[(s * NF + iV0(c)) * NLF + lf], G0n = NTf[(s * NF + iG0(c)) * NLF + lf];
          const T V1n = NTf[(s * NF + iV1(c)) * NLF + lf], G1n = NTf[(s * NF + iG1(c)) * NLF + lf];
          const T W0n = NTf[(s * NF + iV0(d)) * NLF + lf], W1n = NTf[(s * NF + iV1(d)) * NLF + lf], Pn = NTf[(s * NF + iVp) * NLF + lf];
          if (stage == 1)   // NS.cc:593-603 in outward-normal form
            val = (T)K.a1Re * T(0.5) * ns * (G0 + G0n) - (T)K.a1 * T(0.5) * ns * (V0 * W1[s] + V0n * W1n);
          else              // NS.cc:640-652
            val = (T)K.a1Re * T(0.5) * ns * (G0 + G0n) + (T)K.a2Re * T(0.5) * ns * (G1 + G1n)
                  - (T)K.a1 * T(0.5) * ns * (V0 * W0[s] + V0n * W0n) - (T)K.a2 * T(0.5) * ns * (V1 * W1[s] + V1n * W1n);
          if (c == d) val -= T(0.5) * (Pv[s] + Pn) * ns;
        } else {            // boundary faces, NS.cc:671-724 / 725-780
          const int bid = -1 - nbs;
          const T coef_jump = bid == 1 ? T(0) : (T)K.C * F.ih[d], aux = bid == 1 ? T(0) : T(1);
          const T ext = ns * (stage == 1 ? W1[s] : W3[s]);       // u_extr . n
          const T lam = bid == 1 ? T(0) : fabs(ext);
          T uD = 0;
          if (bid == 0 && c == 0) {   // inflow profile depends on y only
            T y;
            if (dim == 1 || d == 1) y = (T)(CG.lo[1] + (cell_coord(CG, cell, 1) + s) * CG.h[1]);
            else {
              const int ty = (dim == 2) ? t0 : (d == 0 ? t0 : t1);    // tangential index that runs along y
              y = (T)(CG.lo[1] + (cell_coord(CG, cell, 1) + (double)F.xq[ty]) * CG.h[1]);
            }
            uD = inflow_u0<T>(y);
          }
          if (stage == 1) val = ((T)K.a1Re * G0 - (T)K.a1 * V0 * W1[s]) * ns;
          else val = ((T)K.a1Re * G0 + (T)K.a2Re * G1 - (T)K.a1 * V0 * W0[s] - (T)K.a2 * V1 * W1[s]) * ns;
          if (c == d) val -= Pv[s] * ns;
          val += (T)K.abRe * T(2) * coef_jump * uD - aux * (T)K.ab * uD * ext + (T)K.ab * lam * uD;
          nd = -aux * (T)K.theta * (T)K.abRe * uD;
        }
#pragma unroll
        for (int q = 0; q < n; ++q) rl[q] += wf * (F.fh[s][q] * val + ns * F.ih[d] * F.gh[s][q] * nd);
      }
      if (d == 0) {
#pragma unroll
        for (int q = 0; q < n; ++q) R[c * NB + base + stride * q] = rl[q];
      } else {
#pragma unroll
        for (int q = 0; q < n; ++q) R[c * NB + base + stride * q] += rl[q];
      }
    }
    rhs_group_sync<WPG>(g);
  }
  // 4. back to the nodal basis (transposed basis change): y, z in shared memory, x on the way out
#pragma unroll
  for (int d = dim - 1; d >= 1; --d) {
    for (int it = tid; it < ncell * dim * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, l = r - c * NL;
      int base, stride;
      rline<dim, n>(d, l, base, stride);
      T *R = Rc(ci) + c * NB;
      T u[n];
#pragma unroll
      for (int q = 0; q < n; ++q) u[q] = R[base + stride * q];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) acc += F.S[q * n + i] * u[q];
        R[base + stride * i] = acc;
      }
    }
    rhs_group_sync<WPG>(g);
  }
  for (int it = tid; it < ncell * dim * NL; it += GT) {
    const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, row = r - c * NL;
    const T *R = Rc(ci) + c * NB + row * RS;
    T u[n], v[n];
#pragma unroll
    for (int q = 0; q < n; ++q) u[q] = R[q];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      T acc = 0;
#pragma unroll
      for (int q = 0; q < n; ++q) acc += F.S[q * n + i] * u[q];
      v[i] = acc;
    }
    store_row<T, n>(dst + ((size_t)(cell0 + ci) * dim + c) * ND + row * n, v, 0);
  }
}


// ------------------------------------------------------------------------------------------ pressure rhs
// n = Gauss points of QGauss(p+1) = nodes of the pressure space; the velocity FE (n + 1 nodes) is evaluated on it.
struct PresRhsConst { double inv_coeff, inv_coeff2; };

template <int dim, int n, typename T>
__global__ void __launch_bounds__(rhs_cpb<dim, n + 1>() * ipow(n + 1, dim - 1))
rhs_pressure_kernel(const __grid_constant__ FormTables<T, n> F, const __grid_constant__ PresRhsConst K, const int *__restrict__ nbr, int n_owned,
                    const T *__restrict__ us, const T *__restrict__ po, T *__restrict__ dst) {
  constexpr int NM = n + 1, NL = ipow(NM, dim - 1), NB = NL * NM, CPB = rhs_cpb<dim, NM>();
  constexpr int NDV = NB, NDP = ipow(n, dim);
  extern __shared__ __align__(16) unsigned char rsm[];
  T *sm = reinterpret_cast<T *>(rsm);
  const int tid = threadIdx.x, ci = tid / NL, l = tid % NL;
  const int cell = blockIdx.x * CPB + ci;
  const bool active = cell < n_owned;
  T *A = sm + (size_t)ci * ((dim + 2) * NB + 8 * NL);   // [dim] u_s, [1] p_old
  T *R = A + (dim + 1) * NB;
  T *NTa = R + NB, *NTb = NTa + 4 * NL;                  // [2 sides][own, nbr][NL]
  if (active) {
    for (int c = 0; c < dim; ++c) {
      const T *g = us + ((size_t)cell * dim + c) * NDV;
      for (int i = l; i < NDV; i += NL) A[c * NB + i] = g[i];
    }
    const T *g = po + (size_t)cell * NDP;
    for (int i = l; i < NDP; i += NL) {
      int r = i, o = 0, st = 1;
      for (int d = 0; d < dim; ++d) { o += (r % n) * st; r /= n; st *= NM; }
      A[dim * NB + o] = g[i];
    }
  }
  __syncthreads();
  {
    int extv[3] = {NM, NM, NM};
    const int extp[3] = {n, n, n};
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      if (active) {
        for (int c = 0; c < dim; ++c) sweep_box<dim, NM, T>(A + c * NB, d, F.Sx, NM, n, extv, l);
        sweep_box<dim, NM, T>(A + dim * NB, d, F.S, n, n, extp, l);
      }
      extv[d] = n;
      __syncthreads();
    }
  }
  T vol = 1;
#pragma unroll
  for (int d = 0; d < dim; ++d) vol *= F.h[d];
  const int t0 = l % NM, t1 = l / NM;
  const bool gline = t0 < n && (dim == 2 || t1 < n);      // this thread owns a line of the Gauss-point box
  if (active)
    for (int i = l; i < NB; i += NL) {
      int r = i;
      T jxw = vol;
      bool ok = true;
#pragma unroll
      for (int d = 0; d < dim; ++d) { const int q = r % NM; r /= NM; if (q >= n) ok = false; else jxw *= F.w[q]; }
      R[i] = ok ? (T)K.inv_coeff * jxw * A[dim * NB + i] : T(0);                                  // NS.cc:833
    }
  __syncthreads();
  const T wt = gline ? (dim == 3 ? F.w[t0] * F.w[t1] : F.w[t0]) : T(0);
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    int base, stride;
    bline<dim, NM>(d, l, base, stride);
    int nb[2] = {-1, -1};
    if (active) {
      // nodal face layers of (u*)_d: own side s, neighbour side 1 - s (exact: the basis is nodal at the end points)
      const T *own = us + ((size_t)cell * dim + d) * NDV + base;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        nb[s] = nbr[(size_t)(2 * d + s) * n_owned + cell];
        NTa[(s * 2 + 0) * NL + l] = own[stride * (s * (NM - 1))];
        NTa[(s * 2 + 1) * NL + l] = nb[s] >= 0 ? us[((size_t)nb[s] * dim + d) * NDV + base + stride * ((1 - s) * (NM - 1))] : T(0);
      }
    }
    __syncthreads();
    if (active && t0 < n) {
      for (int f = 0; f < 4; ++f) {
        T acc = 0;
#pragma unroll
        for (int a2 = 0; a2 < NM; ++a2) acc += F.Sx[t0 * NM + a2] * NTa[f * NL + a2 + NM * t1];
        NTb[f * NL + l] = acc;
      }
    }
    __syncthreads();
    if (dim == 3) {
      if (active && gline) {
        for (int f = 0; f < 4; ++f) {
          T acc = 0;
#pragma unroll
          for (int b2 = 0; b2 < NM; ++b2) acc += F.Sx[t1 * NM + b2] * NTb[f * NL + t0 + NM * b2];
          NTa[f * NL + l] = acc;
        }
      }
      __syncthreads();
    }
    const T *NTf = dim == 3 ? NTa : NTb;
    if (active && gline) {
      T area = 1;
#pragma unroll
      for (int e2 = 0; e2 < dim; ++e2) if (e2 != d) area *= F.h[e2];
      const T wf = area * wt;
      T r[n], fl[n];
#pragma unroll
      for (int q = 0; q < n; ++q) fl[q] = vol * wt * F.w[q] * (T)K.inv_coeff2 * A[d * NB + base + stride * q];   // NS.cc:834
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) acc += F.Dh[q * n + i] * fl[q];
        r[i] = F.ih[d] * acc;
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const T ns = s ? T(1) : T(-1);
        const T V = NTf[(s * 2 + 0) * NL + l], Vn = NTf[(s * 2 + 1) * NL + l];
        const T val = nb[s] >= 0 ? -(T)K.inv_coeff2 * ns * T(0.5) * (V + Vn) : -(T)K.inv_coeff2 * ns * V;   // NS.cc:866-871, 896
#pragma unroll
        for (int q = 0; q < n; ++q) r[q] += wf * F.fh[s][q] * val;
      }
#pragma unroll
      for (int q = 0; q < n; ++q) R[base + stride * q] += r[q];
    }
    __syncthreads();
  }
  {
    const int extp[3] = {n, n, n};
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      if (active) sweep_box<dim, NM, T, true>(R, d, F.S, n, n, extp, l);
      __syncthreads();
    }
  }
  if (active) {
    T *o = dst + (size_t)cell * NDP;
    for (int i = l; i < NDP; i += NL) {
      int r = i, b = 0, st = 1;
      for (int d = 0; d < dim; ++d) { b += (r % n) * st; r /= n; st *= NM; }
      o[i] = R[b];
    }
  }
}

// ------------------------------------------------------------------------------------------ grad p projection rhs
template <int dim, int n, typename T>
__global__ void __launch_bounds__(rhs_cpb<dim, n>() * ipow(n, dim - 1))
rhs_grad_p_kernel(const __grid_constant__ FormTables<T, n> F, int n_owned, const T *__restrict__ sp, T *__restrict__ dst) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = rhs_cpb<dim, n>(), np = n - 1, NDP = ipow(np, dim);
  extern __shared__ __align__(16) unsigned char rsm[];
  T *sm = reinterpret_cast<T *>(rsm);
  const int tid = threadIdx.x, ci = tid / NL, l = tid % NL;
  const int cell = blockIdx.x * CPB + ci;
  const bool active = cell < n_owned;
  T *A = sm + (size_t)ci * ((dim + 1) * ND);
  T *R = A + ND;
  if (active) {
    const T *g = sp + (size_t)cell * NDP;
    for (int i = l; i < NDP; i += NL) {
      int r = i, o = 0, st = 1;
      for (int d = 0; d < dim; ++d) { o += (r % np) * st; r /= np; st *= n; }
      A[o] = g[i];
    }
  }
  __syncthreads();
  {
    int extp[3] = {np, np, np};
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      if (active) sweep_box<dim, n, T>(A, d, F.Sx, np, n, extp, l);
      extp[d] = n;
      __syncthreads();
    }
  }
  T vol = 1;
#pragma unroll
  for (int d = 0; d < dim; ++d) vol *= F.h[d];
  const int t0 = l % n, t1 = l / n;
  const T wt = dim == 3 ? F.w[t0] * F.w[t1] : F.w[t0];
  if (active) {
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      int base, stride;
      bline<dim, n>(d, l, base, stride);
      T u[n];
#pragma unroll
      for (int q = 0; q < n; ++q) u[q] = A[base + stride * q];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T dl = 0;
#pragma unroll
        for (int q2 = 0; q2 < n; ++q2) dl += F.Dh[q * n + q2] * u[q2];
        R[d * ND + base + stride * q] = vol * wt * F.w[q] * F.ih[d] * dl;     // NS.cc:1352: submit_value(grad p)
      }
    }
  }
  __syncthreads();
  {
    const int extv[3] = {n, n, n};
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      if (active)
        for (int c = 0; c < dim; ++c) sweep_box<dim, n, T, true>(R + c * ND, d, F.S, n, n, extv, l);
      __syncthreads();
    }
  }
  if (active)
    for (int c = 0; c < dim; ++c) {
      T *o = dst + ((size_t)cell * dim + c) * ND;
      for (int i = l; i < ND; i += NL) o[i] = R[c * ND + i];
    }
}

// ------------------------------------------------------------------------------------------ host side
CellGeom make_geom(const dgx_mf *mf) {
  CellGeom G;
  const Mesh &m = mf->mesh->m;
  G.dim = m.dim; G.L = mf->level; G.first_cell = mf->first_cell;
  for (int d = 0; d < 3; ++d) { G.c0[d] = d < m.dim ? m.cells0[d] : 1; G.lo[d] = d < m.dim ? m.lo[d] : 0.0; G.h[d] = d < m.dim ? mf->h[d] : 1.0; }
  return G;
}

template <typename K> int set_smem(K kernel, size_t smem) {
  DGX_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return DGX_OK;
}

template <int dim, int n, typename T>
int launch_rhs_velocity(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs) {
  dgx_mf *mf = op->mf;
  FormTables<T, n> F;
  fill_tables<T, n>(F, mf, n - 2);
  const double gamma = gamma_trbdf2(), a21 = 0.5, a22 = 0.5, a31 = (1.0 - gamma) / (2.0 * (2.0 - gamma)), a32 = a31, a33 = 1.0 / (2.0 - gamma);
  VelRhsConst K;
  K.stage = op->stage;
  if (op->stage == 1) { K.alpha = 1.0 / (gamma * op->dt); K.a1 = a21; K.a2 = 0.0; K.ab = a22; }
  else { K.alpha = 1.0 / ((1.0 - gamma) * op->dt); K.a1 = a31; K.a2 = a32; K.ab = a33; }
  K.a1Re = K.a1 / op->Re; K.a2Re = K.a2 / op->Re; K.abRe = K.ab / op->Re;
  K.C = (double)n * n; K.theta = 1.0;
  using Cfg = VelRhsCfg<dim, n, T>;
  constexpr int CPB = Cfg::CPB;
  const size_t smem = Cfg::smem;
  int rc = set_smem(rhs_velocity_kernel<dim, n, T>, smem);
  if (rc) return rc;
  const int n_owned = (int)mf->n_owned;
  if (n_owned == 0) return DGX_OK;
  const T *s3 = op->stage == 2 ? (const T *)srcs[3]->d : (const T *)srcs[1]->d;
  rhs_velocity_kernel<dim, n, T><<<(n_owned + CPB - 1) / CPB, kRhsThreads, smem, mf->ctx->stream>>>(
      F, K, make_geom(mf), mf->d_nbr, n_owned, (const T *)srcs[0]->d, (const T *)srcs[1]->d, (const T *)srcs[2]->d, s3, (T *)dst->d);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, int n, typename T>
int launch_rhs_pressure(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs) {
  dgx_mf *mf = op->mf;
  FormTables<T, n> F;
  fill_tables<T, n>(F, mf, n);          // velocity degree = n (n + 1 nodes) on the n-point quadrature
  const double gamma = gamma_trbdf2(), dt = op->dt;
  PresRhsConst K;
  K.inv_coeff = pressure_kappa(op);
  K.inv_coeff2 = op->stage == 1 ? 1.0 / (gamma * dt) : 1.0 / ((1.0 - gamma) * dt);
  constexpr int NM = n + 1, NL = ipow(NM, dim - 1), NB = NL * NM, CPB = rhs_cpb<dim, NM>();
  const size_t smem = sizeof(T) * (size_t)CPB * ((dim + 2) * NB + 8 * NL);
  int rc = set_smem(rhs_pressure_kernel<dim, n, T>, smem);
  if (rc) return rc;
  const int n_owned = (int)mf->n_owned;
  if (n_owned == 0) return DGX_OK;
  rhs_pressure_kernel<dim, n, T><<<(n_owned + CPB - 1) / CPB, CPB * NL, smem, mf->ctx->stream>>>(F, K, mf->d_nbr, n_owned, (const T *)srcs[0]->d,
                                                                                               (const T *)srcs[1]->d, (T *)dst->d);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, int n, typename T>
int launch_rhs_grad_p(dgx_mf *mf, dgx_vec *dst, const dgx_vec *src) {
  FormTables<T, n> F;
  fill_tables<T, n>(F, mf, n - 2);
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = rhs_cpb<dim, n>();
  const size_t smem = sizeof(T) * (size_t)CPB * ((dim + 1) * ND);
  int rc = set_smem(rhs_grad_p_kernel<dim, n, T>, smem);
  if (rc) return rc;
  const int n_owned = (int)mf->n_owned;
  if (n_owned == 0) return DGX_OK;
  rhs_grad_p_kernel<dim, n, T><<<(n_owned + CPB - 1) / CPB, CPB * NL, smem, mf->ctx->stream>>>(F, n_owned, (const T *)src->d, (T *)dst->d);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

#define DGX_N_SWITCH(NV, CALL)                                                     \
  switch (NV) {                                                                    \
    case 2: return CALL(2);                                                        \
    case 3: return CALL(3);                                                        \
    case 4: return CALL(4);                                                        \
    case 5: return CALL(5);                                                        \
    case 6: return CALL(6);                                                        \
    default: set_error("rhs forms: degree out of range"); return DGX_ERR_UNSUPPORTED; \
  }

template <int dim> int rhs_velocity_dim(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs) {
#define CALL(N) launch_rhs_velocity<dim, N, double>(op, dst, srcs)
  if (op->degree + 1 < 3) { set_error("velocity rhs: velocity degree must be >= 2 (pressure degree >= 1)"); return DGX_ERR_UNSUPPORTED; }
  DGX_N_SWITCH(op->degree + 1, CALL)
#undef CALL
}
template <int dim> int rhs_pressure_dim(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs) {
#define CALL(N) launch_rhs_pressure<dim, N, double>(op, dst, srcs)
  DGX_N_SWITCH(op->degree + 1, CALL)
#undef CALL
}
template <int dim> int rhs_grad_p_dim(dgx_mf *mf, int degree_v, dgx_vec *dst, const dgx_vec *src) {
#define CALL(N) launch_rhs_grad_p<dim, N, double>(mf, dst, src)
  if (degree_v + 1 < 3) { set_error("grad p rhs: velocity degree must be >= 2"); return DGX_ERR_UNSUPPORTED; }
  DGX_N_SWITCH(degree_v + 1, CALL)
#undef CALL
}

}  // namespace
}  // namespace dgx

using namespace dgx;

extern "C" {

// vmult_rhs_velocity (NS.cc:788-798).  op: the velocity operator (degree, Re, dt, TR_BDF2_stage).  srcs as in NS.cc:2439 / 2444:
// stage 1 {u_n, u_extr, pres_n}, stage 2 {u_n, u_n_gamma, pres_int, u_extr}.
int dgx_op_vmult_rhs_velocity(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs, int n_srcs) {
  DGX_REQUIRE(op && dst && srcs, "null argument");
  DGX_REQUIRE(op->kind == DGX_OP_VELOCITY, "vmult_rhs_velocity needs the velocity operator (NS_stage 1)");
  if (op->mf->general) { set_error("right-hand-side forms do not support non-Cartesian levels yet"); return DGX_ERR_UNSUPPORTED; }
  DGX_REQUIRE(n_srcs == (op->stage == 1 ? 3 : 4), "stage 1 takes {u_n, u_extr, p_n}, stage 2 {u_n, u_n_gamma, p, u_extr}");
  DGX_REQUIRE(op->mf->dtype == DGX_F64, "rhs forms are fp64");
  for (int i = 0; i < n_srcs; ++i) {
    DGX_REQUIRE(srcs[i] && srcs[i]->mf == op->mf, "source vector of a different MatrixFree");
    const bool is_p = (i == 2);
    DGX_REQUIRE(srcs[i]->degree == (is_p ? op->degree - 1 : op->degree) && srcs[i]->ncomp == (is_p ? 1 : op->ncomp), "source vector space");
    int rc = ghost_exchange(srcs[i]);   // NS.cc:789-790
    if (rc) return rc;
  }
  DGX_REQUIRE(dst->mf == op->mf && dst->degree == op->degree && dst->ncomp == op->ncomp, "destination vector space");
  return op->mf->dim == 2 ? rhs_velocity_dim<2>(op, dst, srcs) : rhs_velocity_dim<3>(op, dst, srcs);
}

// vmult_rhs_pressure (NS.cc:915-925).  op: the pressure operator.  srcs = {u_star, pres_old} (NS.cc:2481 / 2483)
int dgx_op_vmult_rhs_pressure(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs, int n_srcs) {
  DGX_REQUIRE(op && dst && srcs && n_srcs == 2, "vmult_rhs_pressure takes {u_star, pres_old}");
  DGX_REQUIRE(op->kind == DGX_OP_PRESSURE, "vmult_rhs_pressure needs the pressure operator (NS_stage 2)");
  if (op->mf->general) { set_error("right-hand-side forms do not support non-Cartesian levels yet"); return DGX_ERR_UNSUPPORTED; }
  DGX_REQUIRE(op->mf->dtype == DGX_F64, "rhs forms are fp64");
  DGX_REQUIRE(srcs[0] && srcs[1] && srcs[0]->mf == op->mf && srcs[1]->mf == op->mf, "source vector of a different MatrixFree");
  DGX_REQUIRE(srcs[0]->degree == op->degree + 1 && srcs[0]->ncomp == op->mf->dim, "u_star must be the velocity space (degree_p + 1)");
  DGX_REQUIRE(srcs[1]->degree == op->degree && srcs[1]->ncomp == 1 && dst->degree == op->degree && dst->ncomp == 1 && dst->mf == op->mf, "pressure vector space");
  for (int i = 0; i < 2; ++i) { int rc = ghost_exchange(srcs[i]); if (rc) return rc; }
  return op->mf->dim == 2 ? rhs_pressure_dim<2>(op, dst, srcs) : rhs_pressure_dim<3>(op, dst, srcs);
}

// vmult_grad_p_projection (NS.cc:1363-1366): dst (velocity space) = int grad(p) . phi
int dgx_op_vmult_grad_p_projection(dgx_op *op, dgx_vec *dst, const dgx_vec *src) {
  DGX_REQUIRE(op && dst && src, "null argument");
  DGX_REQUIRE(op->kind == DGX_OP_MASS || op->kind == DGX_OP_VELOCITY, "vmult_grad_p_projection needs an operator on the velocity space");
  if (op->mf->general) { set_error("right-hand-side forms do not support non-Cartesian levels yet"); return DGX_ERR_UNSUPPORTED; }
  DGX_REQUIRE(op->mf->dtype == DGX_F64 && src->mf == op->mf && dst->mf == op->mf, "vectors of a different MatrixFree");
  DGX_REQUIRE(src->degree == op->degree - 1 && src->ncomp == 1 && dst->degree == op->degree && dst->ncomp == op->ncomp, "vector spaces");
  return op->mf->dim == 2 ? rhs_grad_p_dim<2>(op->mf, op->degree, dst, src) : rhs_grad_p_dim<3>(op->mf, op->degree, dst, src);
}

}  // extern "C"
