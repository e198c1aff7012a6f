// Velocity operator, NS_stage == 1 branch of apply_add (NS.cc:1399-1406): assemble_cell_term_velocity
// (NS.cc:933-988), assemble_face_term_velocity (:995-1085), assemble_boundary_term_velocity (:1092-1223), and its
// diagonal as the reference probes it (compute_diagonal NS.cc:2018-2060 with assemble_diagonal_*_velocity
// :1431-1850).  The advecting field u_extr makes the form variable-coefficient, so unlike the pressure operator
// it is evaluated at the quadrature points:
//   1. u and u_extr of a cell are moved to the Lagrange basis on the Gauss points (u_hat = (S x S x S) u): in that
//      basis values at quadrature points are the coefficients themselves, gradients are a 1-D collocation
//      derivative, and face values need a contraction along the face normal only;
//   2. cell term and the 2*dim face terms are accumulated per direction by the thread that owns the line;
//      the neighbour's traces (value, normal derivative, advecting value) are read from the neighbour's dofs and
//      interpolated to the face's Gauss points; each cell computes its own side of every face (no atomics, no
//      compress(add));
//   3. the result is moved back with the transposed basis change.
// Cell-centric outward-normal form of the face flux (equivalent to NS.cc:1028-1036 on both sides of the face):
//   FV = a/Re (-1/2 (dn u + dn u') + sigma (u - u')) + a/2 (u (ue.n) + u' (ue'.n)) + a/2 lambda (u - u'),  tests phi
//   FG = -theta a/Re 1/2 (u - u'),                                                                  tests dn phi
#include "pressure_op.cuh"

namespace dgx {
namespace {

template <typename T, int n>
struct VelParams {
  T S[n * n];        // S[q*n+i] = l_i(xi_q)
  T Dh[n * n];       // collocation derivative on the Gauss points
  T fh[2][n], gh[2][n];   // end-point value / derivative of the Gauss-point basis
  T gN[2][n];        // l_i'(s) of the nodal basis (neighbour traces are taken from nodal dofs)
  T w[n];
  T kA[MAX_DIM][n], kB[MAX_DIM][n];   // w_q aRe / h_d^2 and w_q a / h_d: quadrature-point factors of the cell term
  T aw[n];                            // alpha w_q (mass part)
  T h[MAX_DIM], ih[MAX_DIM];
  T alpha, a, aRe, C, theta;
};

// Work decomposition.  A block of 256 threads is cut into independent GROUPS of WPG warps; a group owns CPG cells and all
// the shared memory they need, runs every phase as a group-wide loop over work items (cell, field or component, line or
// point) and synchronises only with itself (__syncwarp, or a named barrier when WPG > 1).  Groups of one block drift
// apart, so one group's global loads overlap another's contractions; no phase ever waits for the whole block (the first
// version's 16 __syncthreads per block cost 12 % of all issue slots and left 25 % occupancy to hide the rest).
// Sizes: about 75 KB of shared memory per block (3 blocks per SM); small cells give WPG = 1 and several cells per warp,
// large cells several warps per cell.
constexpr int kVelThreads = 256;
constexpr int kVelSmemBudget = 75 * 1024;

template <int dim, int n, typename T>
struct VelCfg {
  static constexpr int NL = ipow(n, dim - 1), ND = NL * n;
  static constexpr int NF = 2 * dim + 1;                 // face fields: V (dim), G (dim), E_d
  static constexpr int RS = n + 1, NB = RS * NL;         // x-rows padded to n + 1: no bank conflicts for y-/z-lines
  static constexpr int NR = NL / n, NLF = RS * NR;       // face fields: rows padded the same way
  static constexpr int PER_CELL = 3 * dim * NB + 2 * NF * NLF;
  static constexpr int fit = kVelSmemBudget / (PER_CELL * (int)sizeof(T));
  static constexpr int CPBraw = fit < 1 ? 1 : (fit > 16 ? 16 : fit);
  static constexpr int WPG = CPBraw >= 8 ? 1 : (CPBraw >= 4 ? 2 : (CPBraw >= 2 ? 4 : 8));
  static constexpr int NG = (kVelThreads / 32) / WPG;    // groups per block
  static constexpr int CPG = CPBraw >= 8 ? CPBraw / 8 : 1;
  static constexpr int CPB = NG * CPG;
  static constexpr size_t smem = sizeof(T) * (size_t)CPB * PER_CELL;
};

template <int dim, int n>
__device__ __forceinline__ void vline(int d, int l, int &base, int &stride) {
  if (dim == 3) {
    if (d == 0) { base = l * n; stride = 1; }
    else if (d == 1) { base = (l % n) + n * n * (l / n); stride = n; }
    else { base = l; stride = n * n; }
  } else {
    if (d == 0) { base = l * n; stride = 1; }
    else { base = l; stride = n; }
  }
}

// same line in a box whose x-rows are padded to n + 1 entries
template <int dim, int n>
__device__ __forceinline__ void pline(int d, int l, int &base, int &stride) {
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
__device__ __forceinline__ void group_sync(int g) {
  if (WPG == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(WPG * 32) : "memory");
}

template <int dim, int n, typename T>
__global__ void __launch_bounds__(kVelThreads, 3)
velocity_kernel(const __grid_constant__ VelParams<T, n> P, const int *__restrict__ nbr, int n_owned, const T *__restrict__ src,
                const T *__restrict__ ue, T *__restrict__ dst, int add, const T *__restrict__ post_scale) {
  using Cfg = VelCfg<dim, n, T>;
  constexpr int NL = Cfg::NL, ND = Cfg::ND, NF = Cfg::NF, RS = Cfg::RS, NB = Cfg::NB, PER_CELL = Cfg::PER_CELL;
  constexpr int NR = Cfg::NR, NLF = Cfg::NLF;
  constexpr int WPG = Cfg::WPG, CPG = Cfg::CPG, GT = WPG * 32;
  extern __shared__ __align__(16) unsigned char vsm[];
  const int g = threadIdx.x / GT, tid = threadIdx.x - g * GT;
  const int cell0 = (blockIdx.x * Cfg::NG + g) * CPG;
  const int ncell = min(CPG, n_owned - cell0);
  if (ncell <= 0) return;                                  // whole group leaves; nobody else waits for it
  T *sm = reinterpret_cast<T *>(vsm) + (size_t)g * CPG * PER_CELL;
  auto Ac = [&](int ci) { return sm + (size_t)ci * PER_CELL; };                 // [2*dim][NB]: u_hat, ue_hat
  auto Rc = [&](int ci) { return sm + (size_t)ci * PER_CELL + 2 * dim * NB; };  // [dim][NB]: result
  auto Na = [&](int ci) { return sm + (size_t)ci * PER_CELL + 3 * dim * NB; };  // [2 sides][NF][NLF]: neighbour traces
  // 1. load u, u_extr one x-row per item and move the row to the Gauss-point basis on the way in.  All loads of a
  //    chunk of rounds are issued before the first contraction (the loop over items is otherwise latency-bound).
  {
    constexpr int ITEMS = CPG * 2 * dim * NL, ROUNDS = (ITEMS + GT - 1) / GT;
    constexpr int RC = (12 * 8) / (n * (int)sizeof(T)) < 1 ? 1 : ((12 * 8) / (n * (int)sizeof(T)) > ROUNDS ? ROUNDS : (12 * 8) / (n * (int)sizeof(T)));
    for (int r0 = 0; r0 < ROUNDS; r0 += RC) {
      T u[RC][n];
#pragma unroll
      for (int rr = 0; rr < RC; ++rr) {
        const int it = tid + (r0 + rr) * GT;
        if (it < ncell * 2 * dim * NL) {
          const int ci = CPG == 1 ? 0 : it / (2 * dim * NL), r = it - ci * (2 * dim * NL), f = r / NL, row = r - f * NL;
          const T *gp = (f < dim ? src : ue) + ((size_t)(cell0 + ci) * dim + (f % dim)) * ND + row * n;
          load_row<T, n>(gp, u[rr]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < RC; ++rr) {
        const int it = tid + (r0 + rr) * GT;
        if (it < ncell * 2 * dim * NL) {
          const int ci = CPG == 1 ? 0 : it / (2 * dim * NL), r = it - ci * (2 * dim * NL), f = r / NL, row = r - f * NL;
          T *A = Ac(ci) + f * NB + row * RS;
#pragma unroll
          for (int q = 0; q < n; ++q) {
            T acc = 0;
#pragma unroll
            for (int i = 0; i < n; ++i) acc += P.S[q * n + i] * u[rr][i];
            A[q] = acc;
          }
        }
      }
    }
  }
  group_sync<WPG>(g);
  // 2. remaining directions of the basis change
#pragma unroll
  for (int d = 1; d < dim; ++d) {
    for (int it = tid; it < ncell * 2 * dim * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (2 * dim * NL), r = it - ci * (2 * dim * NL), f = r / NL, l = r - f * NL;
      int base, stride;
      pline<dim, n>(d, l, base, stride);
      T *A = Ac(ci) + f * NB;
      T u[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = A[base + stride * i];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) acc += P.S[q * n + i] * u[i];
        A[base + stride * q] = acc;
      }
    }
    group_sync<WPG>(g);
  }
  T vol = 1;
#pragma unroll
  for (int d = 0; d < dim; ++d) vol *= P.h[d];
  // 3. per direction: cell term (mass part with d == 0) and the two faces normal to d
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    // 3a. nodal traces of both neighbours: item = (cell, component, nodal tangential point); loads of both sides first
    for (int it = tid; it < ncell * dim * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, l = r - c * NL;
      int base, stride;
      vline<dim, n>(d, l, base, stride);
      T *NTa = Na(ci);
      const int lf = (l / n) * RS + (l % n);
      int nb[2];
      T x[2][n], e[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) nb[s] = nbr[(size_t)(2 * d + s) * n_owned + cell0 + ci];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const size_t off = ((size_t)max(nb[s], 0) * dim + c) * ND + base;   // boundary sides read the first cell and drop it
        if (d == 0) load_row<T, n>(src + off, x[s]);
        else {
#pragma unroll
          for (int i = 0; i < n; ++i) x[s][i] = src[off + stride * i];
        }
        if (c == d) e[s] = ue[off + stride * (1 - s) * (n - 1)];
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (nb[s] < 0) continue;
        T gsum = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) gsum += P.gN[1 - s][i] * x[s][i];
        NTa[(s * NF + c) * NLF + lf] = x[s][(1 - s) * (n - 1)];
        NTa[(s * NF + dim + c) * NLF + lf] = P.ih[d] * gsum;
        if (c == d) NTa[(s * NF + 2 * dim) * NLF + lf] = e[s];
      }
    }
    group_sync<WPG>(g);
    // tangential interpolation to the face Gauss points, in place (dim - 1 passes): item = (cell, side, face field, row or
    // column)
    for (int it = tid; it < ncell * 2 * NF * NR; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (2 * NF * NR), r = it - ci * (2 * NF * NR), sf = r / NR, t1 = r - sf * NR;
      T *io = Na(ci) + sf * NLF + RS * t1;
      T x[n];
#pragma unroll
      for (int a2 = 0; a2 < n; ++a2) x[a2] = io[a2];
#pragma unroll
      for (int t0 = 0; t0 < n; ++t0) {
        T acc = 0;
#pragma unroll
        for (int a2 = 0; a2 < n; ++a2) acc += P.S[t0 * n + a2] * x[a2];
        io[t0] = acc;
      }
    }
    group_sync<WPG>(g);
    if (dim == 3) {
      for (int it = tid; it < ncell * 2 * NF * n; it += GT) {
        const int ci = CPG == 1 ? 0 : it / (2 * NF * n), r = it - ci * (2 * NF * n), sf = r / n, t0 = r - sf * n;
        T *io = Na(ci) + sf * NLF + t0;
        T x[n];
#pragma unroll
        for (int b2 = 0; b2 < n; ++b2) x[b2] = io[RS * b2];
#pragma unroll
        for (int t1 = 0; t1 < n; ++t1) {
          T acc = 0;
#pragma unroll
          for (int b2 = 0; b2 < n; ++b2) acc += P.S[t1 * n + b2] * x[b2];
          io[RS * t1] = acc;
        }
      }
      group_sync<WPG>(g);
    }
    // 3b. item = (cell, component, line along d); everything below carries the common factor cw = vol * w_t0 * w_t1
    for (int it = tid; it < ncell * dim * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, l = r - c * NL;
      const int cell = cell0 + ci;
      const T *A = Ac(ci);
      T *R = Rc(ci);
      const T *NTf = Na(ci);
      int base, stride;
      pline<dim, n>(d, l, base, stride);
      const int t0 = l % n, t1 = l / n, lf = t1 * RS + t0;
      const T cw = vol * (dim == 3 ? P.w[t0] * P.w[t1] : P.w[t0]);
      const T sigma = P.C * P.ih[d];
      T u[n], ed[n], rl[n];
      T Ed[2] = {0, 0};
#pragma unroll
      for (int q = 0; q < n; ++q) {
        u[q] = A[c * NB + base + stride * q];
        ed[q] = A[(dim + d) * NB + base + stride * q];
        Ed[0] += P.fh[0][q] * ed[q]; Ed[1] += P.fh[1][q] * ed[q];
      }
      // cell term, direction d: flux = w_q (a/Re du/dx_d - a u ue_d) / h_d, tested with d(phi)/d(xi_d)
      T fl[n];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T dl = 0;
#pragma unroll
        for (int q2 = 0; q2 < n; ++q2) dl += P.Dh[q * n + q2] * u[q2];
        fl[q] = P.kA[d][q] * dl - P.kB[d][q] * (u[q] * ed[q]);
      }
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = d == 0 ? P.aw[i] * u[i] : T(0);   // mass part: alpha w_q u
#pragma unroll
        for (int q = 0; q < n; ++q) acc += P.Dh[q * n + i] * fl[q];
        rl[i] = acc;
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int nbc = nbr[(size_t)(2 * d + s) * n_owned + cell];
        const T ns = s ? T(1) : T(-1);
        T V = 0, G = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) { V += P.fh[s][q] * u[q]; G += P.gh[s][q] * u[q]; }
        G *= P.ih[d];
        T FV, FG;
        if (nbc >= 0) {
          const T Vn = NTf[(s * NF + c) * NLF + lf], Gn = NTf[(s * NF + dim + c) * NLF + lf], En = NTf[(s * NF + 2 * dim) * NLF + lf];
          const T lam = fmax(fabs(Ed[s]), fabs(En)), jump = V - Vn;
          FV = P.aRe * (T(-0.5) * ns * (G + Gn) + sigma * jump) + P.a * T(0.5) * ns * (V * Ed[s] + Vn * En) + T(0.5) * P.a * lam * jump;
          FG = -P.theta * P.aRe * T(0.5) * jump;
        } else if (nbc != -2) {   // boundary_id != 1: weak Dirichlet, coef_trasp = 0 (NS.cc:1113-1128)
          const T lam = fabs(Ed[s]);
          FV = P.aRe * (-ns * G + T(2) * sigma * V) + P.a * lam * V;
          FG = -P.theta * P.aRe * V;
        } else {                    // boundary_id == 1: mirror state u_m = (u_0, -u_1, u_2), NS.cc:1139-1155
          const T lam = fabs(Ed[s]);
          const T dV = (c == 1) ? T(2) * V : T(0);          // u - u_m
          const T aV = (c == 1) ? T(0) : V;                 // (u + u_m) / 2
          const T dn = (c == 0 && d < 2) ? T(0) : ns * G;   // (1/2 (grad u + grad u_m) n)_c
          FV = P.aRe * (-dn + sigma * dV) + P.a * aV * (ns * Ed[s]) + T(0.5) * P.a * lam * dV;
          FG = -P.theta * P.aRe * dV;
        }
        const T FVs = P.ih[d] * FV, FGs = ns * P.ih[d] * P.ih[d] * FG;   // face area = vol / h_d
#pragma unroll
        for (int q = 0; q < n; ++q) rl[q] += P.fh[s][q] * FVs + P.gh[s][q] * FGs;
      }
      if (d == 0) {
#pragma unroll
        for (int q = 0; q < n; ++q) R[c * NB + base + stride * q] = cw * rl[q];
      } else {
#pragma unroll
        for (int q = 0; q < n; ++q) R[c * NB + base + stride * q] += cw * rl[q];
      }
    }
    group_sync<WPG>(g);
  }
  // 4. back to the nodal basis (transposed basis change): y, z in shared memory, x on the way out
#pragma unroll
  for (int d = dim - 1; d >= 1; --d) {
    for (int it = tid; it < ncell * dim * NL; it += GT) {
      const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, l = r - c * NL;
      int base, stride;
      pline<dim, n>(d, l, base, stride);
      T *R = Rc(ci) + c * NB;
      T u[n];
#pragma unroll
      for (int q = 0; q < n; ++q) u[q] = R[base + stride * q];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) acc += P.S[q * n + i] * u[q];
        R[base + stride * i] = acc;
      }
    }
    group_sync<WPG>(g);
  }
  for (int it = tid; it < ncell * dim * NL; it += GT) {
    const int ci = CPG == 1 ? 0 : it / (dim * NL), r = it - ci * (dim * NL), c = r / NL, row = r - c * NL;
    const T *R = Rc(ci) + c * NB + row * RS;
    T u[n];
#pragma unroll
    for (int q = 0; q < n; ++q) u[q] = R[q];
    T v[n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      T acc = 0;
#pragma unroll
      for (int q = 0; q < n; ++q) acc += P.S[q * n + i] * u[q];
      v[i] = acc;
    }
    if (post_scale) {   // Jacobi preconditioner of the GMRES loop riding on the store: dst = D^-1 (A src)
      T sc[n];
      load_row<T, n>(post_scale + ((size_t)(cell0 + ci) * dim + c) * ND + row * n, sc);
#pragma unroll
      for (int i = 0; i < n; ++i) v[i] *= sc[i];
    }
    store_row<T, n>(dst + ((size_t)(cell0 + ci) * dim + c) * ND + row * n, v, add);
  }
}

// ---- tensorised diagonal -------------------------------------------------------------------------------------
// The reference probes the diagonal with n^dim operator applications per cell (NS.cc:1442-1471, 1527-1530: dof i of
// every component is set to one on BOTH sides of every face and entry i is kept).  Because the probe phi_i is a
// rank-1 tensor, every term of [A phi_i]_i is a contraction of a field (1, u_extr . e_d, its face traces, lambda)
// with squared 1-D shape values: sum_q w_q ue_d(q) phi_i(q) d_d phi_i(q) etc.  These are sum-factorised moments, so
// the whole diagonal costs about ONE operator application instead of n^dim.  Same numbers as the probing to round-off.
template <typename T, int n>
struct VelDiagParams {
  T S[n * n];          // nodal -> Gauss points
  T P2[n * n];         // P2[i][q] = S[q][i]^2
  T PD[n * n];         // PD[i][q] = S[q][i] * l_i'(xi_q)
  T fh[2][n];          // end-point values of the Gauss-point basis (own traces of u_extr)
  T f[2][n], g[2][n];  // nodal basis: l_i(s), l_i'(s)
  T w[n], Md[n], Kd[n];   // weights, diag of the 1-D mass / stiffness matrix
  T h[MAX_DIM], ih[MAX_DIM];
  T alpha, a, aRe, C, theta;
};

template <int dim, int n, typename T>
__host__ __device__ constexpr int veld_cpb() {
  constexpr int per_cell = (int)sizeof(T) * ((dim + 2) * ipow(n, dim) + (2 * 3 + 2 * dim * 3) * ipow(n, dim - 1));
  return (64 * 1024 / per_cell) < 1 ? 1 : ((64 * 1024 / per_cell) > 16 ? 16 : (64 * 1024 / per_cell));
}

template <int dim, int n, typename T>
__global__ void __launch_bounds__(kVelThreads)
velocity_diag_kernel(const __grid_constant__ VelDiagParams<T, n> P, const int *__restrict__ nbr, int n_owned, const T *__restrict__ ue,
                     T *__restrict__ out) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = veld_cpb<dim, n, T>(), NT = kVelThreads;
  constexpr int PER_CELL = (dim + 2) * ND + (2 * 3 + 2 * dim * 3) * NL;
  extern __shared__ __align__(16) unsigned char dsm[];
  T *sm = reinterpret_cast<T *>(dsm);
  const int tid = threadIdx.x, cell0 = blockIdx.x * CPB, ncell = min(CPB, n_owned - cell0);
  auto Ec = [&](int ci) { return sm + (size_t)ci * PER_CELL; };                          // [dim][ND] u_extr at the Gauss points
  auto Tm = [&](int ci) { return sm + (size_t)ci * PER_CELL + dim * ND; };               // [ND] scratch
  auto Ts = [&](int ci) { return sm + (size_t)ci * PER_CELL + (dim + 1) * ND; };         // [ND] sum_d T_d
  auto Fa = [&](int ci) { return sm + (size_t)ci * PER_CELL + (dim + 2) * ND; };         // [3][NL] face fields (ping)
  auto Fb = [&](int ci) { return Fa(ci) + 3 * NL; };                                     // [3][NL] (pong)
  auto Mf = [&](int ci) { return Fa(ci) + 6 * NL; };                                     // [2*dim faces][3][NL] face moments
  for (int it = tid; it < ncell * dim * ND; it += NT) {
    const int ci = it / (dim * ND), r = it - ci * (dim * ND);
    Ec(ci)[r] = ue[(size_t)(cell0 + ci) * dim * ND + r];
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    for (int it = tid; it < ncell * dim * NL; it += NT) {
      const int ci = it / (dim * NL), r = it - ci * (dim * NL), f = r / NL, l = r - f * NL;
      int base, stride;
      vline<dim, n>(d, l, base, stride);
      T *A = Ec(ci) + f * ND;
      T u[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = A[base + stride * i];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) acc += P.S[q * n + i] * u[i];
        A[base + stride * q] = acc;
      }
    }
    __syncthreads();
  }
  T vol = 1;
#pragma unroll
  for (int d = 0; d < dim; ++d) vol *= P.h[d];
  for (int it = tid; it < ncell * ND; it += NT) Ts(it / ND)[it % ND] = 0;
  // cell moments T_d[i] = vol/h_d sum_q w_q ue_d(q) phi_i(q) d(phi_i)/d(xi_d)(q): three sweeps per direction
#pragma unroll
  for (int d = 0; d < dim; ++d) {
#pragma unroll
    for (int e = 0; e < dim; ++e) {
      for (int it = tid; it < ncell * NL; it += NT) {
        const int ci = it / NL, l = it - ci * NL;
        int base, stride;
        vline<dim, n>(e, l, base, stride);
        const T *in = e == 0 ? Ec(ci) + d * ND : Tm(ci);
        T u[n];
#pragma unroll
        for (int q = 0; q < n; ++q) u[q] = in[base + stride * q] * P.w[q];
#pragma unroll
        for (int i = 0; i < n; ++i) {
          T acc = 0;
#pragma unroll
          for (int q = 0; q < n; ++q) acc += (e == d ? P.PD[i * n + q] : P.P2[i * n + q]) * u[q];
          Tm(ci)[base + stride * i] = acc;
        }
      }
      __syncthreads();
    }
    for (int it = tid; it < ncell * ND; it += NT) Ts(it / ND)[it % ND] += vol * P.ih[d] * Tm(it / ND)[it % ND];
    __syncthreads();
  }
  // face moments: for every face, the three fields w*E_own, w*E_nbr, w*lambda contracted with S^2 (x) S^2
#pragma unroll
  for (int d = 0; d < dim; ++d) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      // nodal trace of the neighbour's advecting component d (field 1, ping buffer), interpolated below
      for (int it = tid; it < ncell * NL; it += NT) {
        const int ci = it / NL, l = it - ci * NL;
        const int nbc = nbr[(size_t)(2 * d + s) * n_owned + cell0 + ci];
        int base, stride;
        vline<dim, n>(d, l, base, stride);
        Fa(ci)[1 * NL + l] = nbc >= 0 ? ue[((size_t)nbc * dim + d) * ND + base + stride * (1 - s) * (n - 1)] : T(0);
      }
      __syncthreads();
      for (int it = tid; it < ncell * NL; it += NT) {
        const int ci = it / NL, l = it - ci * NL, t0 = l % n, t1 = l / n;
        T acc = 0;
#pragma unroll
        for (int a2 = 0; a2 < n; ++a2) acc += P.S[t0 * n + a2] * Fa(ci)[1 * NL + a2 + n * t1];
        Fb(ci)[1 * NL + l] = acc;
      }
      __syncthreads();
      if (dim == 3) {
        for (int it = tid; it < ncell * NL; it += NT) {
          const int ci = it / NL, l = it - ci * NL, t0 = l % n, t1 = l / n;
          T acc = 0;
#pragma unroll
          for (int b2 = 0; b2 < n; ++b2) acc += P.S[t1 * n + b2] * Fb(ci)[1 * NL + t0 + n * b2];
          Fa(ci)[1 * NL + l] = acc;
        }
        __syncthreads();
      }
      // own trace, lambda, quadrature weights -> Fb[0..2]
      for (int it = tid; it < ncell * NL; it += NT) {
        const int ci = it / NL, l = it - ci * NL, t0 = l % n, t1 = l / n;
        const int nbc = nbr[(size_t)(2 * d + s) * n_owned + cell0 + ci];
        int base, stride;
        vline<dim, n>(d, l, base, stride);
        T Eo = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) Eo += P.fh[s][q] * Ec(ci)[d * ND + base + stride * q];
        const T En = dim == 3 ? Fa(ci)[1 * NL + l] : Fb(ci)[1 * NL + l];
        const T wt = dim == 3 ? P.w[t0] * P.w[t1] : P.w[t0];
        const T lam = nbc >= 0 ? fmax(fabs(Eo), fabs(En)) : fabs(Eo);
        T *o = dim == 3 ? Fb(ci) : Fa(ci);
        o[0 * NL + l] = wt * Eo; o[1 * NL + l] = wt * En; o[2 * NL + l] = wt * lam;
      }
      __syncthreads();
      // moments: contract both tangential directions with S^2
      {
        for (int it = tid; it < ncell * 3 * NL; it += NT) {
          const int ci = it / (3 * NL), r = it - ci * (3 * NL), f = r / NL, l = r - f * NL, t0 = l % n, t1 = l / n;
          const T *in = (dim == 3 ? Fb(ci) : Fa(ci)) + f * NL;
          T acc = 0;
#pragma unroll
          for (int a2 = 0; a2 < n; ++a2) acc += P.P2[t0 * n + a2] * in[a2 + n * t1];
          if (dim == 3) Fa(ci)[f * NL + l] = acc;
          else Mf(ci)[((2 * d + s) * 3 + f) * NL + l] = acc;
        }
        __syncthreads();
        if (dim == 3) {
          for (int it = tid; it < ncell * 3 * NL; it += NT) {
            const int ci = it / (3 * NL), r = it - ci * (3 * NL), f = r / NL, l = r - f * NL, t0 = l % n, t1 = l / n;
            T acc = 0;
#pragma unroll
            for (int b2 = 0; b2 < n; ++b2) acc += P.P2[t1 * n + b2] * Fa(ci)[f * NL + t0 + n * b2];
            Mf(ci)[((2 * d + s) * 3 + f) * NL + l] = acc;
          }
          __syncthreads();
        }
      }
    }
  }
  // assemble the entries: item = (cell, scalar dof i); all components share everything except boundary_id == 1 faces
  for (int it = tid; it < ncell * ND; it += NT) {
    const int ci = it / ND, i = it - ci * ND, cell = cell0 + ci;
    int idx[3] = {0, 0, 0};
    { int r = i; for (int d = 0; d < dim; ++d) { idx[d] = r % n; r /= n; } }
    T mprod = vol, kk = 0;
#pragma unroll
    for (int d = 0; d < dim; ++d) mprod *= P.Md[idx[d]];
#pragma unroll
    for (int d = 0; d < dim; ++d) kk += P.ih[d] * P.ih[d] * P.Kd[idx[d]] / P.Md[idx[d]];
    const T base = P.alpha * mprod + P.aRe * mprod * kk - P.a * Ts(ci)[i];
    T val[dim];
#pragma unroll
    for (int c = 0; c < dim; ++c) val[c] = base;
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      // tangential position of dof i on faces normal to d, in the line numbering of vline
      int l;
      if (dim == 3) l = d == 0 ? idx[1] + n * idx[2] : (d == 1 ? idx[0] + n * idx[2] : idx[0] + n * idx[1]);
      else l = d == 0 ? idx[1] : idx[0];
      T area = 1, m0 = 1;
#pragma unroll
      for (int e = 0; e < dim; ++e) if (e != d) { area *= P.h[e]; m0 *= P.Md[idx[e]]; }
      m0 *= area;
      const T sigma = P.C * P.ih[d];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int nbc = nbr[(size_t)(2 * d + s) * n_owned + cell];
        const T ns = s ? T(1) : T(-1);
        const T *M = Mf(ci) + (2 * d + s) * 3 * NL;
        const T mE = area * M[0 * NL + l], mEn = area * M[1 * NL + l], mL = area * M[2 * NL + l];
        const T fs = P.f[s][idx[d]], fp = P.f[1 - s][idx[d]], gs = P.ih[d] * P.g[s][idx[d]], gp = P.ih[d] * P.g[1 - s][idx[d]];
        if (nbc >= 0) {
          const T FV = P.aRe * (T(-0.5) * ns * (gs + gp) + sigma * (fs - fp)) * m0 + P.a * T(0.5) * ns * (fs * mE + fp * mEn) + T(0.5) * P.a * (fs - fp) * mL;
          const T FG = -P.theta * P.aRe * T(0.5) * (fs - fp) * m0;
          const T add = fs * FV + ns * gs * FG;
#pragma unroll
          for (int c = 0; c < dim; ++c) val[c] += add;
        } else if (nbc != -2) {
          const T FV = P.aRe * (-ns * gs + T(2) * sigma * fs) * m0 + P.a * fs * mL;
          const T FG = -P.theta * P.aRe * fs * m0;
          const T add = fs * FV + ns * gs * FG;
#pragma unroll
          for (int c = 0; c < dim; ++c) val[c] += add;
        } else {
#pragma unroll
          for (int c = 0; c < dim; ++c) {
            const T dV = c == 1 ? T(2) * fs : T(0), aV = c == 1 ? T(0) : fs;
            const T dn = (c == 0 && d < 2) ? T(0) : ns * gs;
            const T FV = P.aRe * (-dn + sigma * dV) * m0 + P.a * aV * ns * mE + T(0.5) * P.a * dV * mL;
            const T FG = -P.theta * P.aRe * dV * m0;
            val[c] += fs * FV + ns * gs * FG;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < dim; ++c) out[((size_t)cell * dim + c) * ND + i] = T(1) / val[c];   // NS.cc:2055-2059
  }
}

// probe vector of compute_diagonal: dof `i` of every component of every cell (ghosts included) is one
template <typename T>
__global__ void probe_fill_kernel(T *v, long long total, int ND, int i) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) v[g] = (g % ND == i) ? T(1) : T(0);
}
template <typename T>
__global__ void probe_pick_kernel(T *diag, const T *res, long long n_blocks, int ND, int i) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < n_blocks; b += (long long)gridDim.x * blockDim.x) diag[b * ND + i] = res[b * ND + i];
}
template <typename T>
__global__ void reciprocal_kernel(T *v, long long n) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) v[g] = T(1) / v[g];
}

template <int dim, int n, typename T>
int launch_velocity(dgx_op *op, T *dst, const T *src, bool add) {
  dgx_mf *mf = op->mf;
  const Shape1D &s = get_shape(n - 1, n);
  VelParams<T, n> P;
  for (int i = 0; i < n * n; ++i) { P.S[i] = (T)s.S[i]; P.Dh[i] = (T)s.Dhat[i]; }
  for (int sd = 0; sd < 2; ++sd)
    for (int i = 0; i < n; ++i) { P.fh[sd][i] = (T)s.fhat[sd][i]; P.gh[sd][i] = (T)s.ghat[sd][i]; P.gN[sd][i] = (T)s.g[sd][i]; }
  for (int i = 0; i < n; ++i) P.w[i] = (T)s.wq[i];
  for (int d = 0; d < MAX_DIM; ++d) { const double h = d < dim ? mf->h[d] : 1.0; P.h[d] = (T)h; P.ih[d] = (T)(1.0 / h); }
  // stage constants exactly as NS.cc:957-958 / 982-983, 286-287, 396-397
  const double gamma = gamma_trbdf2();
  const double a22 = 0.5, a33 = 1.0 / (2.0 - gamma);
  const double a = op->stage == 1 ? a22 : a33;
  P.alpha = (T)(op->stage == 1 ? 1.0 / (gamma * op->dt) : 1.0 / ((1.0 - gamma) * op->dt));
  P.a = (T)a; P.aRe = (T)(a / op->Re);
  P.C = (T)((double)n * n);   // C_u = (p_v + 1)^2, NS.cc:293
  P.theta = (T)1.0;           // theta_v, NS.cc:290
  for (int d = 0; d < MAX_DIM; ++d)
    for (int q = 0; q < n; ++q) {
      const double ih = d < dim ? 1.0 / mf->h[d] : 1.0;
      P.kA[d][q] = (T)(s.wq[q] * (a / op->Re) * ih * ih);
      P.kB[d][q] = (T)(s.wq[q] * a * ih);
    }
  for (int q = 0; q < n; ++q) P.aw[q] = (T)(s.wq[q] * (op->stage == 1 ? 1.0 / (gamma * op->dt) : 1.0 / ((1.0 - gamma) * op->dt)));
  constexpr int CPB = VelCfg<dim, n, T>::CPB, NT = kVelThreads;
  const size_t smem = VelCfg<dim, n, T>::smem;
  static bool configured_dev[64] = {};   // per device: the attribute belongs to the device's copy of the function
  bool &configured = configured_dev[mf->ctx->device >= 0 && mf->ctx->device < 64 ? mf->ctx->device : 0];
  if (!configured) {
    DGX_CUDA_CHECK(cudaFuncSetAttribute(velocity_kernel<dim, n, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int n_owned = (int)mf->n_owned;
  if (n_owned == 0) return DGX_OK;
  const int grid = (n_owned + CPB - 1) / CPB;
  velocity_kernel<dim, n, T><<<grid, NT, smem, mf->ctx->stream>>>(P, mf->d_nbr, n_owned, src, (const T *)op->u_extr->d, dst, add ? 1 : 0,
                                                                  op->post_scale && !add ? (const T *)op->post_scale->d : nullptr);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int velocity_by_degree(dgx_op *op, T *dst, const T *src, bool add) {
  switch (op->degree) {
    case 1: return launch_velocity<dim, 2, T>(op, dst, src, add);
    case 2: return launch_velocity<dim, 3, T>(op, dst, src, add);
    case 3: return launch_velocity<dim, 4, T>(op, dst, src, add);
    case 4: return launch_velocity<dim, 5, T>(op, dst, src, add);
    case 5: return launch_velocity<dim, 6, T>(op, dst, src, add);
    case 6: return launch_velocity<dim, 7, T>(op, dst, src, add);
    default: set_error("velocity operator: degree must be in 1..6"); return DGX_ERR_UNSUPPORTED;
  }
}

template <typename T>
int velocity_dispatch(dgx_op *op, T *dst, const T *src, bool add) {
  return op->mf->dim == 2 ? velocity_by_degree<2, T>(op, dst, src, add) : velocity_by_degree<3, T>(op, dst, src, add);
}

template <int dim, int n, typename T>
int launch_velocity_diag(dgx_op *op, T *out) {
  dgx_mf *mf = op->mf;
  const Shape1D &s = get_shape(n - 1, n);
  VelDiagParams<T, n> P;
  for (int q = 0; q < n; ++q)
    for (int i = 0; i < n; ++i) {
      P.S[q * n + i] = (T)s.S[q * n + i];
      P.P2[i * n + q] = (T)(s.S[q * n + i] * s.S[q * n + i]);
      P.PD[i * n + q] = (T)(s.S[q * n + i] * s.D[q * n + i]);
    }
  for (int sd = 0; sd < 2; ++sd)
    for (int i = 0; i < n; ++i) { P.fh[sd][i] = (T)s.fhat[sd][i]; P.f[sd][i] = (T)s.f[sd][i]; P.g[sd][i] = (T)s.g[sd][i]; }
  for (int i = 0; i < n; ++i) { P.w[i] = (T)s.wq[i]; P.Md[i] = (T)s.M[i * n + i]; P.Kd[i] = (T)s.K[i * n + i]; }
  for (int d = 0; d < MAX_DIM; ++d) { const double h = d < dim ? mf->h[d] : 1.0; P.h[d] = (T)h; P.ih[d] = (T)(1.0 / h); }
  const double gamma = gamma_trbdf2();
  const double a = op->stage == 1 ? 0.5 : 1.0 / (2.0 - gamma);
  P.alpha = (T)(op->stage == 1 ? 1.0 / (gamma * op->dt) : 1.0 / ((1.0 - gamma) * op->dt));
  P.a = (T)a; P.aRe = (T)(a / op->Re); P.C = (T)((double)n * n); P.theta = (T)1.0;
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = veld_cpb<dim, n, T>();
  const size_t smem = sizeof(T) * (size_t)CPB * ((dim + 2) * ND + (2 * 3 + 2 * dim * 3) * NL);
  static bool configured_dev[64] = {};   // per device: the attribute belongs to the device's copy of the function
  bool &configured = configured_dev[mf->ctx->device >= 0 && mf->ctx->device < 64 ? mf->ctx->device : 0];
  if (!configured) {
    DGX_CUDA_CHECK(cudaFuncSetAttribute(velocity_diag_kernel<dim, n, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int n_owned = (int)mf->n_owned;
  if (n_owned == 0) return DGX_OK;
  velocity_diag_kernel<dim, n, T><<<(n_owned + CPB - 1) / CPB, kVelThreads, smem, mf->ctx->stream>>>(P, mf->d_nbr, n_owned, (const T *)op->u_extr->d, out);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int velocity_diag_by_degree(dgx_op *op, T *out) {
  switch (op->degree) {
    case 1: return launch_velocity_diag<dim, 2, T>(op, out);
    case 2: return launch_velocity_diag<dim, 3, T>(op, out);
    case 3: return launch_velocity_diag<dim, 4, T>(op, out);
    case 4: return launch_velocity_diag<dim, 5, T>(op, out);
    case 5: return launch_velocity_diag<dim, 6, T>(op, out);
    case 6: return launch_velocity_diag<dim, 7, T>(op, out);
    default: set_error("velocity diagonal: degree must be in 1..6"); return DGX_ERR_UNSUPPORTED;
  }
}

template <typename T>
int diagonal_by_probing(dgx_op *op, dgx_vec *out) {
  dgx_mf *mf = op->mf;
  long long ND = 1;
  for (int d = 0; d < mf->dim; ++d) ND *= op->degree + 1;
  dgx_vec *probe = nullptr, *res = nullptr;   // pooled with the operator: no cudaMalloc / cudaFree per call
  int rc = op_pool_vec(op, 0, &probe);
  if (!rc) rc = op_pool_vec(op, 1, &res);
  if (rc) return rc;
  const long long total = probe->n_owned + probe->n_ghost, nblocks = probe->n_owned / ND;
  cudaStream_t st = mf->ctx->stream;
  const int g1 = (int)std::min<long long>((total + 255) / 256 + 1, 148 * 8), g2 = (int)std::min<long long>((nblocks + 255) / 256 + 1, 148 * 8);
  for (int i = 0; i < ND && !rc; ++i) {
    probe_fill_kernel<T><<<g1, 256, 0, st>>>((T *)probe->d, total, (int)ND, i);
    count_launch();
    rc = velocity_dispatch<T>(op, (T *)res->d, (const T *)probe->d, false);
    if (rc) break;
    if (nblocks) { probe_pick_kernel<T><<<g2, 256, 0, st>>>((T *)out->d, (const T *)res->d, nblocks, (int)ND, i); count_launch(); }
  }
  if (!rc && out->n_owned) { reciprocal_kernel<T><<<g1, 256, 0, st>>>((T *)out->d, out->n_owned); count_launch(); }   // NS.cc:2055-2059
  if (!rc && cudaGetLastError() != cudaSuccess) { set_error("velocity diagonal: kernel launch failed"); rc = DGX_ERR_CUDA; }
  return rc;
}

}  // namespace

int velocity_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add) {
  if (!op->u_extr) { set_error("velocity operator: set_u_extr() has not been called"); return DGX_ERR_INVALID; }
  if (op->mf->dtype == DGX_F64) return velocity_dispatch<double>(op, (double *)dst->d, (const double *)src->d, add);
  return velocity_dispatch<float>(op, (float *)dst->d, (const float *)src->d, add);
}

// compute_diagonal, NS_stage == 1: unit-vector probing exactly as NS.cc:1442-1471, 1527-1530 do it -- the probe sets
// dof i of ALL components, on BOTH sides of every face, and keeps the entries of dof i.  Probing all cells at once
// with the operator kernel reproduces that (n^dim operator applications, as in the reference).
int velocity_diagonal(dgx_op *op, dgx_vec *out) {
  if (!op->u_extr) { set_error("velocity operator: set_u_extr() has not been called"); return DGX_ERR_INVALID; }
  if (op->variant == 1)   // literal probing with n^dim operator applications (kept as the cross-check of the tensorised form)
    return op->mf->dtype == DGX_F64 ? diagonal_by_probing<double>(op, out) : diagonal_by_probing<float>(op, out);
  if (op->mf->dim == 2) return op->mf->dtype == DGX_F64 ? velocity_diag_by_degree<2, double>(op, (double *)out->d) : velocity_diag_by_degree<2, float>(op, (float *)out->d);
  return op->mf->dtype == DGX_F64 ? velocity_diag_by_degree<3, double>(op, (double *)out->d) : velocity_diag_by_degree<3, float>(op, (float *)out->d);
}

}  // namespace dgx
