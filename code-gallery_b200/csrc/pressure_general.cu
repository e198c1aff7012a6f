// Pressure operator on NON-Cartesian cells: the same weak forms (assemble_cell_term_pressure NS.cc:1230-1252,
// assemble_face_term_pressure :1259-1292, assemble_boundary_term_pressure :1299-1326) with the metric of a MappingQ1 cell
// (NS.cc:2329: MatrixFree::reinit(MappingQ1, ...)) at every quadrature point -- the regime where the operator streams
// J^-1 / JxW data next to the vectors and is HBM-bound by them, not by the sum factorisation.
//
// The mesh keeps the refine_global topology (cell order, neighbour tables, ghost plan are those of mesh.cpp) and gets its vertex
// positions per level from dgx_mesh_set_vertices; a cell is the multilinear image of the unit cell.
//
// Data in HBM, per owned cell, in the operator's precision:
//   cell:  Gs[e e'] = (J^-1 J^-T)_{e e'} JxW (symmetric, dim (dim+1)/2 entries) and JxW            per cell quadrature point
//   face:  A_own[e] = (J_own^-1 n)_e JxW_f,  A_nbr[e] = (J_nbr^-1 n)_e JxW_f,  W = sigma JxW_f      per face quadrature point
// with n the OWN outward normal, sigma = C_p/2 (|grad xi_d|_own + |grad xi_d|_nbr) at the face's first quadrature point
// (NS.cc:1274-1275: get_normal_vector(0) * inverse_jacobian(0), [dim-1] = the direction normal to the face), C_p |grad xi_d| on
// boundary faces (NS.cc:1311).  With these the cell-centric outward-normal form of both face lambdas is, at a face point,
//   tested with phi:            JxW_f FV = -1/2 (grad_xi u . A_own + grad_xi u' . A_nbr) + W (u - u')
//   tested with grad_xi phi:    A_own * (-theta/2) (u - u')
// (boundary id 1: -grad_xi u . A_own + W u  and  A_own * (-theta) u; other boundary ids contribute nothing), which both sides of
// a face evaluate for themselves: no atomics, no compress(add).  Algorithmic traffic per cell of n^dim dofs:
//   2 n^dim (src, dst) + (dim (dim+1)/2 + 1) n^dim + 2 dim (2 dim + 1) n^(dim-1)  values  (3-D Q4: 16 + 56 + 67 = 139 B/DoF fp64).
//
// Kernel: cell-centric, one thread per line, Gauss-point collocation basis (values at quadrature points are the coefficients,
// gradients a 1-D collocation derivative), everything between the first load and the last store in shared memory.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pressure_op.cuh"
#include "reduce.cuh"

namespace dgx {
namespace {

template <typename T, int n>
struct GenParams {
  T S[n * n];              // nodal -> Gauss points
  T Dh[n * n];             // collocation derivative on the Gauss points: Dh[q][q2]
  T fh[2][n], gh[2][n];    // end-point value / derivative of the Gauss-point basis
  T gN[2][n];              // l_i'(s) of the nodal basis (neighbour traces are taken from nodal dofs)
  T kappa, theta, bfac;    // 1/coeff (NS.cc:1237), theta_p (NS.cc:291), factor on the boundary penalty (2 while probing: NS.cc:1997)
};

template <int dim, int n, typename T>
__host__ __device__ constexpr int gen_cpb() {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n;
  constexpr int per_cell = (int)sizeof(T) * ((dim + 2) * ND + 2 * 8 * NL);
  constexpr int by_threads = (128 / NL) > 0 ? (128 / NL) : 1;
  constexpr int by_smem = (44 * 1024) / per_cell > 0 ? (44 * 1024) / per_cell : 1;
  return by_threads < by_smem ? by_threads : by_smem;
}

template <int dim, int n>
__device__ __forceinline__ void gline(int d, int l, int &base, int &stride) {
  if (dim == 3) {
    if (d == 0) { base = l * n; stride = 1; }
    else if (d == 1) { base = (l % n) + n * n * (l / n); stride = n; }
    else { base = l; stride = n * n; }
  } else {
    if (d == 0) { base = l * n; stride = 1; }
    else { base = l; stride = n; }
  }
}

// one batch of CPB cells: dst[cells] (+)= A src.  Shared by the operator kernel (one batch per block) and the single-CTA coarse-grid CG
// kernel (loops over the batches of a tiny level).  Block-wide phases; every thread of the block must call it.
template <int dim, int n, typename T>
__device__ __forceinline__ void general_cell_batch(const GenParams<T, n> &P, const int *__restrict__ nbr, int n_owned, const T *__restrict__ cellG,
                                                   const T *__restrict__ faceA, const T *src, T *dst, int add, const Epilogue<T> &E, int cell0, T *su, T *sr,
                                                   T *st_all, T *sf) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = gen_cpb<dim, n, T>(), NT = CPB * NL;
  constexpr int NS = dim * (dim + 1) / 2, NFD = 2 * dim + 1, NFA = 8;
  T *st[dim];
#pragma unroll
  for (int e = 0; e < dim; ++e) st[e] = st_all + e * CPB * ND;
  const int tid = threadIdx.x, nblk = min(CPB, n_owned - cell0);
  const int c = tid / NL, l = tid - c * NL;
  const bool active = c < nblk;
  const int cell = cell0 + c;
  auto SF = [&](int ci, int s, int f) { return sf + ((ci * 2 + s) * NFA + f) * NL; };
  for (int idx = tid; idx < nblk * ND; idx += NT) su[idx] = src[(size_t)cell0 * ND + idx];
  __syncthreads();
  // nodal -> Gauss-point basis, direction by direction
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    if (active) {
      int base, stride;
      gline<dim, n>(d, l, base, stride);
      T u[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = su[c * ND + base + stride * i];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) acc += P.S[q * n + i] * u[i];
        su[c * ND + base + stride * q] = acc;
      }
    }
    __syncthreads();
  }
  // reference gradient at the quadrature points
#pragma unroll
  for (int e = 0; e < dim; ++e) {
    if (active) {
      int base, stride;
      gline<dim, n>(e, l, base, stride);
      T u[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = su[c * ND + base + stride * i];
#pragma unroll
      for (int q = 0; q < n; ++q) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) acc += P.Dh[q * n + i] * u[i];
        st[e][c * ND + base + stride * q] = acc;
      }
    }
  }
  __syncthreads();
  // quadrature-point work of NS.cc:1245-1248 with the metric folded in: flux = Gs grad_xi u, value = kappa JxW u
  for (int idx = tid; idx < nblk * ND; idx += NT) {
    const int ci = idx / ND, q = idx - ci * ND;
    const T *G = cellG + (size_t)(cell0 + ci) * (NS + 1) * ND + q;
    T g[dim], t[dim];
#pragma unroll
    for (int e = 0; e < dim; ++e) g[e] = st[e][idx];
    if constexpr (dim == 3) {
      const T g00 = G[0], g11 = G[ND], g22 = G[2 * ND], g01 = G[3 * ND], g02 = G[4 * ND], g12 = G[5 * ND];
      t[0] = g00 * g[0] + g01 * g[1] + g02 * g[2];
      t[1] = g01 * g[0] + g11 * g[1] + g12 * g[2];
      t[2] = g02 * g[0] + g12 * g[1] + g22 * g[2];
    } else {
      const T g00 = G[0], g11 = G[ND], g01 = G[2 * ND];
      t[0] = g00 * g[0] + g01 * g[1];
      t[1] = g01 * g[0] + g11 * g[1];
    }
#pragma unroll
    for (int e = 0; e < dim; ++e) st[e][idx] = t[e];
    sr[idx] = P.kappa * G[NS * ND] * su[idx];
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < dim; ++e) {
    if (active) {
      int base, stride;
      gline<dim, n>(e, l, base, stride);
      T t[n];
#pragma unroll
      for (int q = 0; q < n; ++q) t[q] = st[e][c * ND + base + stride * q];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) acc += P.Dh[q * n + i] * t[q];
        sr[c * ND + base + stride * i] += acc;
      }
    }
    __syncthreads();
  }
  // faces, one direction at a time
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    int base = 0, stride = 1, nb[2] = {-1, -1};
    if (active) {
      gline<dim, n>(d, l, base, stride);
      T u[n];
#pragma unroll
      for (int q = 0; q < n; ++q) u[q] = su[c * ND + base + stride * q];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        nb[s] = nbr[(size_t)(2 * d + s) * n_owned + cell];
        T V = 0, Gn = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) { V += P.fh[s][q] * u[q]; Gn += P.gh[s][q] * u[q]; }
        SF(c, s, 0)[l] = V; SF(c, s, 1)[l] = Gn;
        T Vn = 0, Gnn = 0;
        if (nb[s] >= 0) {   // nodal traces of the neighbour: value on its face (d, 1 - s) and reference derivative along d
          const T *un = src + (size_t)nb[s] * ND + base;
          T x[n];
#pragma unroll
          for (int i = 0; i < n; ++i) x[i] = un[stride * i];
          Vn = x[(1 - s) * (n - 1)];
#pragma unroll
          for (int i = 0; i < n; ++i) Gnn += P.gN[1 - s][i] * x[i];
        }
        SF(c, s, 2)[l] = Vn; SF(c, s, 3)[l] = Gnn;
      }
    }
    __syncthreads();
    // neighbour traces: nodal -> Gauss points along the tangential directions (in place, row by row)
    for (int it = tid; it < nblk * 4 * (NL / n); it += NT) {
      const int ci = it / (4 * (NL / n)), r = it - ci * (4 * (NL / n)), sfld = r / (NL / n), t1 = r - sfld * (NL / n);
      T *row = SF(ci, sfld >> 1, 2 + (sfld & 1)) + n * t1;
      T x[n];
#pragma unroll
      for (int a = 0; a < n; ++a) x[a] = row[a];
#pragma unroll
      for (int t0 = 0; t0 < n; ++t0) {
        T acc = 0;
#pragma unroll
        for (int a = 0; a < n; ++a) acc += P.S[t0 * n + a] * x[a];
        row[t0] = acc;
      }
    }
    __syncthreads();
    if (dim == 3) {
      for (int it = tid; it < nblk * 4 * n; it += NT) {
        const int ci = it / (4 * n), r = it - ci * (4 * n), sfld = r / n, t0 = r - sfld * n;
        T *col = SF(ci, sfld >> 1, 2 + (sfld & 1)) + t0;
        T x[n];
#pragma unroll
        for (int b = 0; b < n; ++b) x[b] = col[n * b];
#pragma unroll
        for (int t1 = 0; t1 < n; ++t1) {
          T acc = 0;
#pragma unroll
          for (int b = 0; b < n; ++b) acc += P.S[t1 * n + b] * x[b];
          col[n * t1] = acc;
        }
      }
      __syncthreads();
    }
    // fluxes at the face points
    const int t0 = l % n, t1 = l / n;
    const int e0 = d == 0 ? 1 : 0, e1 = d == 2 ? 1 : 2;   // tangential directions, ascending (e1 only in 3-D)
    if (active) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        T FV = 0, Fn = 0, F0 = 0, F1 = 0;
        if (nb[s] >= 0 || nb[s] == -2) {
          const T *fp = faceA + ((size_t)cell * 2 * dim + (2 * d + s)) * NFD * NL + l;
          const T *Vo = SF(c, s, 0), *Vnb = SF(c, s, 2);
          T go[dim], gn[dim];
          go[d] = SF(c, s, 1)[l]; gn[d] = SF(c, s, 3)[l];
          T a0 = 0, b0 = 0;
#pragma unroll
          for (int a = 0; a < n; ++a) { a0 += P.Dh[t0 * n + a] * Vo[a + n * t1]; b0 += P.Dh[t0 * n + a] * Vnb[a + n * t1]; }
          go[e0] = a0; gn[e0] = b0;
          if (dim == 3) {
            T a1 = 0, b1 = 0;
#pragma unroll
            for (int b = 0; b < n; ++b) { a1 += P.Dh[t1 * n + b] * Vo[t0 + n * b]; b1 += P.Dh[t1 * n + b] * Vnb[t0 + n * b]; }
            go[e1 % dim] = a1; gn[e1 % dim] = b1;
          }
          T Ao[dim], dot_o = 0, dot_n = 0;
#pragma unroll
          for (int e = 0; e < dim; ++e) { Ao[e] = fp[e * NL]; dot_o += go[e] * Ao[e]; dot_n += gn[e] * fp[(dim + e) * NL]; }
          const T W = fp[2 * dim * NL], V = Vo[l];
          T cj;
          if (nb[s] >= 0) {   // NS.cc:1284-1287
            const T jump = V - Vnb[l];
            FV = T(-0.5) * (dot_o + dot_n) + W * jump;
            cj = T(-0.5) * P.theta * jump;
          } else {            // boundary_id == 1, NS.cc:1320-1321
            FV = -dot_o + P.bfac * W * V;
            cj = -P.theta * V;
          }
          Fn = Ao[d] * cj; F0 = Ao[e0] * cj;
          if (dim == 3) F1 = Ao[e1 % dim] * cj;
        }
        SF(c, s, 7)[l] = FV; SF(c, s, 4)[l] = Fn; SF(c, s, 5)[l] = F0; SF(c, s, 6)[l] = F1;
      }
    }
    __syncthreads();
    if (active) {   // both sides of the line in one read-modify-write of the result
      T Fval[2], Fnor[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        T v = SF(c, s, 7)[l];
        const T *F0 = SF(c, s, 5), *F1 = SF(c, s, 6);
#pragma unroll
        for (int a = 0; a < n; ++a) v += P.Dh[a * n + t0] * F0[a + n * t1];
        if (dim == 3) {
#pragma unroll
          for (int b = 0; b < n; ++b) v += P.Dh[b * n + t1] * F1[t0 + n * b];
        }
        Fval[s] = v; Fnor[s] = SF(c, s, 4)[l];
      }
#pragma unroll
      for (int q = 0; q < n; ++q)
        sr[c * ND + base + stride * q] += P.fh[0][q] * Fval[0] + P.gh[0][q] * Fnor[0] + P.fh[1][q] * Fval[1] + P.gh[1][q] * Fnor[1];
    }
    __syncthreads();
  }
  // back to the nodal basis
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    if (active) {
      int base, stride;
      gline<dim, n>(d, l, base, stride);
      T u[n];
#pragma unroll
      for (int q = 0; q < n; ++q) u[q] = sr[c * ND + base + stride * q];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int q = 0; q < n; ++q) acc += P.S[q * n + i] * u[q];
        sr[c * ND + base + stride * i] = acc;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nblk * ND; idx += NT) {
    const size_t o = (size_t)cell0 * ND + idx;
    if (E.on) {
      T v = E.c * (E.rhs[o] - sr[idx]);
      if (E.d) v *= E.d[o];
      v += E.a * src[o];
      if (E.xold) v += E.b * E.xold[o];
      dst[o] = v;
    } else {
      dst[o] = add ? dst[o] + sr[idx] : sr[idx];
    }
  }
}

template <int dim, int n, typename T>
__global__ void __launch_bounds__(gen_cpb<dim, n, T>() * ipow(n, dim - 1))
pressure_general_kernel(const __grid_constant__ GenParams<T, n> P, const int *__restrict__ nbr, int n_owned, const T *__restrict__ cellG,
                        const T *__restrict__ faceA, const T *__restrict__ src, T *__restrict__ dst, int add, const __grid_constant__ Epilogue<T> E) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = gen_cpb<dim, n, T>();
  __shared__ T su[CPB * ND];            // u in the Gauss-point basis
  __shared__ T sr[CPB * ND];            // result
  __shared__ T st[dim * CPB * ND];      // reference gradient, then the flux Gs grad_xi u
  __shared__ T sf[CPB * 2 * 8 * NL];    // face fields of the current direction: [cell][side][field][tangential point]
  general_cell_batch<dim, n, T>(P, nbr, n_owned, cellG, faceA, src, dst, add, E, blockIdx.x * CPB, su, sr, st, sf);
}

// MGCoarseGridIterativeSolver<SolverCG, PreconditionIdentity> (NS.cc:2519-2529) for a metric level of a few cells, entirely inside ONE
// CTA: same algebra as pressure_coarse_cg_kernel (pressure_op.cu) / cg_core (solvers.cu) with x0 = 0; the device-side log makes the
// V-cycle capturable as a CUDA graph on MappingQ1 hierarchies too.
template <int dim, int n, typename T>
__global__ void __launch_bounds__(gen_cpb<dim, n, T>() * ipow(n, dim - 1))
pressure_general_coarse_cg_kernel(const __grid_constant__ GenParams<T, n> P, const int *nbr, int n_cells, const T *cellG, const T *faceA, const T *b, T *x,
                                  T *r, T *p, T *v, const double *tol_ptr, int max_it, int *info, double *res_out, int *log_count, int log_cap) {
  const double tol = *tol_ptr;
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = gen_cpb<dim, n, T>(), NT = CPB * NL;
  __shared__ T su[CPB * ND];
  __shared__ T sr[CPB * ND];
  __shared__ T st[dim * CPB * ND];
  __shared__ T sf[CPB * 2 * 8 * NL];
  __shared__ double red[32];
  __shared__ double bc;
  const int tid = threadIdx.x, N = n_cells * ND;
  const Epilogue<T> E;
  auto block_sum = [&](double val) -> double {
    val = warp_sum_existing(val);   // NT is not a multiple of 32 for most degrees
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    if (tid == 0) {
      double s = 0;
      for (int w = 0; w < (NT + 31) / 32; ++w) s += red[w];
      bc = s;
    }
    __syncthreads();
    return bc;
  };
  double acc = 0;
  for (int i = tid; i < N; i += NT) { const T bi = b[i]; r[i] = bi; x[i] = 0; acc += (double)bi * (double)bi; }
  double rr = block_sum(acc), res = sqrt(rr), prev_rr = 0;
  int it = 0;
  bool conv = res <= tol;
  while (!conv && it < max_it && isfinite(res)) {
    ++it;
    if (it > 1) {
      const T beta = (T)(rr / prev_rr);
      for (int i = tid; i < N; i += NT) p[i] = beta * p[i] + r[i];
    } else {
      for (int i = tid; i < N; i += NT) p[i] = r[i];
    }
    __syncthreads();
    for (int c0 = 0; c0 < n_cells; c0 += CPB) {
      general_cell_batch<dim, n, T>(P, nbr, n_cells, cellG, faceA, p, v, 0, E, c0, su, sr, st, sf);
      __syncthreads();
    }
    acc = 0;
    for (int i = tid; i < N; i += NT) acc += (double)p[i] * (double)v[i];
    const double pv = block_sum(acc);
    const T alpha = (T)(rr / pv);
    acc = 0;
    for (int i = tid; i < N; i += NT) {
      x[i] += alpha * p[i];
      const T ri = r[i] - alpha * v[i];
      r[i] = ri;
      acc += (double)ri * (double)ri;
    }
    prev_rr = rr;
    rr = block_sum(acc);
    res = sqrt(rr);
    conv = res <= tol;
  }
  if (tid == 0) {   // append to the device-side log: (iterations, converged), residual
    const int slot = atomicAdd(log_count, 1) % log_cap;
    info[2 * slot] = it; info[2 * slot + 1] = conv ? 1 : 0; res_out[slot] = res;
  }
}

// ---- metric tables (host, double) ------------------------------------------------------------------------------------------
struct CellGeom {
  int dim;
  double X[8][3];   // vertices in deal.II order (x fastest)
};

void cell_vertices(const Mesh &m, int level, const int *cc, CellGeom &g) {
  g.dim = m.dim;
  const std::vector<double> &V = m.vertices[level];
  const int nx = m.cells_dir(level, 0) + 1, ny = m.cells_dir(level, 1) + 1;
  for (int v = 0; v < (1 << m.dim); ++v) {
    const int ix = cc[0] + (v & 1), iy = cc[1] + ((v >> 1) & 1), iz = m.dim == 3 ? cc[2] + ((v >> 2) & 1) : 0;
    const size_t idx = (size_t)ix + (size_t)nx * ((size_t)iy + (size_t)ny * iz);
    for (int i = 0; i < m.dim; ++i) g.X[v][i] = V[idx * m.dim + i];
  }
}

// J[i][e] = d x_i / d xi_e of the multilinear map at xi; returns det J and inv[e][i] = d xi_e / d x_i
double jacobian(const CellGeom &g, const double *xi, double inv[3][3]) {
  const int dim = g.dim;
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int v = 0; v < (1 << dim); ++v)
    for (int e = 0; e < dim; ++e) {
      double dN = 1;
      for (int d = 0; d < dim; ++d) {
        const int b = (v >> d) & 1;
        dN *= d == e ? (b ? 1.0 : -1.0) : (b ? xi[d] : 1.0 - xi[d]);
      }
      for (int i = 0; i < dim; ++i) J[i][e] += g.X[v][i] * dN;
    }
  if (dim == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    inv[0][0] = J[1][1] / det; inv[0][1] = -J[0][1] / det;
    inv[1][0] = -J[1][0] / det; inv[1][1] = J[0][0] / det;
    return det;
  }
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  inv[0][0] = c00 / det; inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  inv[1][0] = c01 / det; inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  inv[2][0] = c02 / det; inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det; inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  return det;
}

// host tables of the owned cells for n quadrature points per direction; returns false for an inverted cell
bool build_metric(const dgx_mf *mf, int n, std::vector<double> &cellG, std::vector<double> &faceA) {
  const Mesh &m = mf->mesh->m;
  const int dim = mf->dim, level = mf->level;
  const Shape1D &sh = get_shape(n - 1, n);
  const int NL = ipow(n, dim - 1), ND = NL * n, NS = dim * (dim + 1) / 2, NFD = 2 * dim + 1;
  const double Cp = (double)n * n;   // C_p = (p + 1)^2, NS.cc:292
  cellG.assign((size_t)mf->n_owned * (NS + 1) * ND, 0.0);
  faceA.assign((size_t)mf->n_owned * 2 * dim * NFD * NL, 0.0);
  bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
  for (long long c = 0; c < mf->n_owned; ++c) {
    int cc[MAX_DIM] = {0, 0, 0};
    m.coords_of(level, mf->first_cell + c, cc);
    CellGeom g;
    cell_vertices(m, level, cc, g);
    double *G = cellG.data() + (size_t)c * (NS + 1) * ND;
    for (int q = 0; q < ND; ++q) {
      int r = q;
      double xi[3] = {0, 0, 0}, w = 1;
      for (int d = 0; d < dim; ++d) { xi[d] = sh.xq[r % n]; w *= sh.wq[r % n]; r /= n; }
      double inv[3][3];
      const double det = jacobian(g, xi, inv);
      if (!(det > 0)) ok = false;
      const double JxW = det * w;
      auto gs = [&](int e, int f) { double s = 0; for (int i = 0; i < dim; ++i) s += inv[e][i] * inv[f][i]; return s * JxW; };
      if (dim == 3) { G[q] = gs(0, 0); G[ND + q] = gs(1, 1); G[2 * ND + q] = gs(2, 2); G[3 * ND + q] = gs(0, 1); G[4 * ND + q] = gs(0, 2); G[5 * ND + q] = gs(1, 2); }
      else { G[q] = gs(0, 0); G[ND + q] = gs(1, 1); G[2 * ND + q] = gs(0, 1); }
      G[NS * ND + q] = JxW;
    }
    for (int f = 0; f < 2 * dim; ++f) {
      const int d = f / 2, s = f % 2;
      const long long nbg = m.neighbor(level, mf->first_cell + c, f);
      if (nbg < 0 && nbg != -2) continue;   // boundary ids other than 1 contribute nothing (NS.cc:1308)
      CellGeom gn;
      if (nbg >= 0) {
        int cn[MAX_DIM] = {0, 0, 0};
        m.coords_of(level, nbg, cn);
        cell_vertices(m, level, cn, gn);
      }
      double *A = faceA.data() + ((size_t)c * 2 * dim + f) * NFD * NL;
      double sigma = 0;
      for (int t = 0; t < NL; ++t) {
        const int e0 = d == 0 ? 1 : 0, e1 = d == 2 ? 1 : 2;
        double xi[3] = {0, 0, 0}, w = sh.wq[t % n];
        xi[d] = s;
        xi[e0] = sh.xq[t % n];
        if (dim == 3) { xi[e1] = sh.xq[t / n]; w *= sh.wq[t / n]; }
        double inv[3][3], invn[3][3];
        const double det = jacobian(g, xi, inv);
        if (!(det > 0)) ok = false;
        double norm = 0;
        for (int i = 0; i < dim; ++i) norm += inv[d][i] * inv[d][i];
        norm = std::sqrt(norm);
        double normal[3] = {0, 0, 0};
        for (int i = 0; i < dim; ++i) normal[i] = (s ? 1.0 : -1.0) * inv[d][i] / norm;
        const double JxWf = det * norm * w;
        double norm_n = 0;
        if (nbg >= 0) {
          double xin[3] = {xi[0], xi[1], xi[2]};
          xin[d] = 1 - s;
          const double detn = jacobian(gn, xin, invn);
          if (!(detn > 0)) ok = false;
          for (int i = 0; i < dim; ++i) norm_n += invn[d][i] * invn[d][i];
          norm_n = std::sqrt(norm_n);
        }
        if (t == 0) sigma = nbg >= 0 ? Cp * 0.5 * (norm + norm_n) : Cp * norm;   // NS.cc:1274-1275 / 1311: first face point
        for (int e = 0; e < dim; ++e) {
          double ao = 0, an = 0;
          for (int i = 0; i < dim; ++i) { ao += inv[e][i] * normal[i]; if (nbg >= 0) an += invn[e][i] * normal[i]; }
          A[(size_t)e * NL + t] = ao * JxWf;
          A[(size_t)(dim + e) * NL + t] = an * JxWf;
        }
        A[(size_t)2 * dim * NL + t] = sigma * JxWf;
      }
    }
  }
  return ok;
}

template <typename T>
int upload(const std::vector<double> &h, void **d) {
  std::vector<T> tmp(h.begin(), h.end());
  DGX_CUDA_CHECK(cudaMalloc(d, std::max<size_t>(tmp.size(), 1) * sizeof(T)));
  DGX_CUDA_CHECK(cudaMemcpy(*d, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DGX_OK;
}

int get_metric(dgx_mf *mf, int n, dgx_mf::GeneralMetric **out) {
  auto it = mf->gmetric.find(n);
  if (it == mf->gmetric.end()) {
    std::vector<double> cellG, faceA;
    if (!build_metric(mf, n, cellG, faceA)) { set_error("general mesh: a cell has a non-positive Jacobian determinant"); return DGX_ERR_INVALID; }
    dgx_mf::GeneralMetric gm;
    DGX_CUDA_CHECK(cudaSetDevice(mf->ctx->device));
    int rc = mf->dtype == DGX_F64 ? upload<double>(cellG, &gm.cellG) : upload<float>(cellG, &gm.cellG);
    if (!rc) rc = mf->dtype == DGX_F64 ? upload<double>(faceA, &gm.faceA) : upload<float>(faceA, &gm.faceA);
    if (rc) {   // e.g. out of memory: the tables of a large level are 100+ bytes per dof
      if (gm.cellG) cudaFree(gm.cellG);
      if (gm.faceA) cudaFree(gm.faceA);
      cudaGetLastError();
      return rc;
    }
    gm.bytes = (cellG.size() + faceA.size()) * (mf->dtype == DGX_F64 ? 8 : 4);
    it = mf->gmetric.emplace(n, gm).first;
  }
  *out = &it->second;
  return DGX_OK;
}

template <int dim, int n, typename T>
int launch_general(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FusedUpdate *fu, double bfac) {
  dgx_mf *mf = op->mf;
  const int n_owned = (int)mf->n_owned;
  if (n_owned == 0) return DGX_OK;
  dgx_mf::GeneralMetric *gm = nullptr;
  if (int rc = get_metric(mf, n, &gm)) return rc;
  GenParams<T, n> P;
  fill_gen_params<T, n>(P, op, bfac);
  Epilogue<T> E;
  if (fu) {
    E.on = 1;
    E.xold = fu->xold ? (const T *)fu->xold->d : nullptr;
    E.d = fu->d ? (const T *)fu->d->d : nullptr;
    E.rhs = (const T *)fu->rhs->d;
    E.a = (T)fu->a; E.b = (T)fu->b; E.c = (T)fu->c;
  }
  constexpr int CPB = gen_cpb<dim, n, T>(), NT = CPB * ipow(n, dim - 1);
  pressure_general_kernel<dim, n, T><<<(n_owned + CPB - 1) / CPB, NT, 0, mf->ctx->stream>>>(P, mf->d_nbr, n_owned, (const T *)gm->cellG, (const T *)gm->faceA,
                                                                                          (const T *)src->d, (T *)dst->d, add ? 1 : 0, E);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int general_by_degree(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FusedUpdate *fu, double bfac) {
  switch (op->degree) {
    case 1: return launch_general<dim, 2, T>(op, dst, src, add, fu, bfac);
    case 2: return launch_general<dim, 3, T>(op, dst, src, add, fu, bfac);
    case 3: return launch_general<dim, 4, T>(op, dst, src, add, fu, bfac);
    case 4: return launch_general<dim, 5, T>(op, dst, src, add, fu, bfac);
    case 5: return launch_general<dim, 6, T>(op, dst, src, add, fu, bfac);
    case 6: return launch_general<dim, 7, T>(op, dst, src, add, fu, bfac);
    default: set_error("pressure operator on a general mesh: degree must be in 1..6"); return DGX_ERR_UNSUPPORTED;
  }
}

template <typename T, int n>
void fill_gen_params(GenParams<T, n> &P, const dgx_op *op, double bfac) {
  const Shape1D &s = get_shape(n - 1, n);
  for (int i = 0; i < n * n; ++i) { P.S[i] = (T)s.S[i]; P.Dh[i] = (T)s.Dhat[i]; }
  for (int sd = 0; sd < 2; ++sd)
    for (int i = 0; i < n; ++i) { P.fh[sd][i] = (T)s.fhat[sd][i]; P.gh[sd][i] = (T)s.ghat[sd][i]; P.gN[sd][i] = (T)s.g[sd][i]; }
  P.kappa = (T)pressure_kappa(op);
  P.theta = (T)1.0;   // theta_p, NS.cc:291
  P.bfac = (T)bfac;
}

template <int dim, int n, typename T>
int launch_general_coarse(dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info, double *d_res,
                          int *d_count, int log_cap) {
  dgx_mf *mf = op->mf;
  dgx_mf::GeneralMetric *gm = nullptr;
  if (int rc = get_metric(mf, n, &gm)) return rc;
  GenParams<T, n> P;
  fill_gen_params<T, n>(P, op, 1.0);
  constexpr int NT = gen_cpb<dim, n, T>() * ipow(n, dim - 1);
  pressure_general_coarse_cg_kernel<dim, n, T><<<1, NT, 0, mf->ctx->stream>>>(P, mf->d_nbr, (int)mf->n_owned, (const T *)gm->cellG, (const T *)gm->faceA, (const T *)b->d,
                                                                              (T *)x->d, (T *)r->d, (T *)p->d, (T *)v->d, d_tol, max_it, d_info, d_res, d_count, log_cap);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int general_coarse_by_degree(dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info, double *d_res,
                             int *d_count, int log_cap) {
  switch (op->degree) {
    case 1: return launch_general_coarse<dim, 2, T>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
    case 2: return launch_general_coarse<dim, 3, T>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
    case 3: return launch_general_coarse<dim, 4, T>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
    case 4: return launch_general_coarse<dim, 5, T>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
    case 5: return launch_general_coarse<dim, 6, T>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
    case 6: return launch_general_coarse<dim, 7, T>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
    default: return DGX_ERR_UNSUPPORTED;
  }
}

template <typename T>
__global__ void gen_probe_fill_kernel(T *v, long long total, int ND, int i) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) v[g] = (g % ND == i) ? T(1) : T(0);
}
template <typename T>
__global__ void gen_probe_pick_kernel(T *diag, const T *res, long long n_cells, int ND, int i) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < n_cells; b += (long long)gridDim.x * blockDim.x) diag[b * ND + i] = T(1) / res[b * ND + i];
}

}  // namespace

int pressure_general_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FusedUpdate *fu, double bfac) {
  dgx_mf *mf = op->mf;
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? general_by_degree<2, double>(op, dst, src, add, fu, bfac) : general_by_degree<2, float>(op, dst, src, add, fu, bfac);
  return mf->dtype == DGX_F64 ? general_by_degree<3, double>(op, dst, src, add, fu, bfac) : general_by_degree<3, float>(op, dst, src, add, fu, bfac);
}

// compute_diagonal (NS.cc:2018-2060) as the reference does it: one probe per local dof i, set on BOTH sides of every face
// (NS.cc:1921-1922), the boundary probe with 2 sigma (NS.cc:1997); entry i of the result is kept and inverted (NS.cc:2055-2059).
// All cells are probed at once, so this is n^dim operator applications.
template <typename T>
static int general_diag(dgx_op *op, dgx_vec *out) {
  dgx_mf *mf = op->mf;
  int ND = 1;
  for (int d = 0; d < mf->dim; ++d) ND *= op->degree + 1;
  dgx_vec *probe = nullptr, *res = nullptr;
  int rc = op_pool_vec(op, 0, &probe);
  if (!rc) rc = op_pool_vec(op, 1, &res);
  if (rc) return rc;
  const long long total = probe->n_owned + probe->n_ghost, ncell = mf->n_owned;
  cudaStream_t st = mf->ctx->stream;
  const int g1 = (int)std::min<long long>((total + 255) / 256 + 1, 148 * 8), g2 = (int)std::min<long long>((ncell + 255) / 256 + 1, 148 * 8);
  for (int i = 0; i < ND && !rc; ++i) {
    gen_probe_fill_kernel<T><<<g1, 256, 0, st>>>((T *)probe->d, total, ND, i);
    count_launch();
    rc = pressure_general_apply(op, res, probe, false, nullptr, 2.0);
    if (rc) break;
    if (ncell) { gen_probe_pick_kernel<T><<<g2, 256, 0, st>>>((T *)out->d, (const T *)res->d, ncell, ND, i); count_launch(); }
  }
  if (!rc && cudaGetLastError() != cudaSuccess) { set_error("general pressure diagonal: kernel launch failed"); rc = DGX_ERR_CUDA; }
  return rc;
}

// The probed diagonal depends on (TR_BDF2 stage, dt) only (through 1/coeff): it is kept per pair, so a time loop with a fixed dt
// pays the n^dim probing applications once per stage and level (DGX_NO_DIAG_CACHE: probe on every call like the reference).
int pressure_general_diagonal(dgx_op *op, dgx_vec *out) {
  static const bool no_cache = getenv("DGX_NO_DIAG_CACHE") != nullptr;
  if (!no_cache)
    for (auto &dc : op->gen_diag_cache)
      if (dc.stage == op->stage && dc.dt == op->dt) return dgx_vec_copy(out, dc.v);
  int rc = op->mf->dtype == DGX_F64 ? general_diag<double>(op, out) : general_diag<float>(op, out);
  if (rc || no_cache) return rc;
  if (op->gen_diag_cache.size() >= 4) {   // keep the table short: a changing dt (last step of a run) evicts the oldest pair
    dgx_vec *old = op->gen_diag_cache.front().v;
    old->borrowed = false;
    dgx_vec_destroy(old);
    op->gen_diag_cache.erase(op->gen_diag_cache.begin());
  }
  dgx_vec *keep = nullptr;
  rc = dgx_vec_create(op->mf, op->degree, op->ncomp, &keep);
  if (rc) return rc;
  keep->borrowed = true;
  rc = dgx_vec_copy(keep, out);
  if (rc) { keep->borrowed = false; dgx_vec_destroy(keep); return rc; }
  op->gen_diag_cache.push_back({op->stage, op->dt, keep});
  return DGX_OK;
}

// ---- DG mass operator on MappingQ1 cells (NS_stage 3, assemble_cell_term_projection_grad_p NS.cc:1373-1390) ------------------------
// Per component M_cell = S^T diag(JxW) S with S = nodal -> Gauss points (n_q = n) and JxW the last row of the cell table; its exact
// inverse S^-1 diag(1 / JxW) S^-T is the same kernel with other 1-D matrices (opt-in replacement of project_grad's CG, as on
// Cartesian cells).  Algorithmic bytes: src + dst + 1 value per point.
namespace {
template <typename T, int n>
struct MassGenParams {
  T A1[n * n];   // first sweeps:  out[q] = sum_i A1[q][i] in[i]
  T A2[n * n];   // second sweeps: out[i] = sum_q A2[i][q] in[q]
};

template <int dim, int n, typename T>
__global__ void __launch_bounds__(gen_cpb<dim, n, T>() * ipow(n, dim - 1))
mass_general_kernel(const __grid_constant__ MassGenParams<T, n> P, long long n_items, int ncomp, const T *__restrict__ cellG, int invert, const T *__restrict__ src,
                    T *__restrict__ dst, int add) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = gen_cpb<dim, n, T>(), NT = CPB * NL, NS = dim * (dim + 1) / 2;
  __shared__ T sa[CPB * ND];
  const int tid = threadIdx.x;
  const long long item0 = (long long)blockIdx.x * CPB;     // item = (cell, component): n^dim contiguous values
  const int nblk = (int)min((long long)CPB, n_items - item0);
  for (int idx = tid; idx < nblk * ND; idx += NT) sa[idx] = src[(size_t)item0 * ND + idx];
  __syncthreads();
  const int c = tid / NL, l = tid - c * NL;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      if (c < nblk) {
        int base, stride;
        gline<dim, n>(d, l, base, stride);
        T u[n];
#pragma unroll
        for (int i = 0; i < n; ++i) u[i] = sa[c * ND + base + stride * i];
#pragma unroll
        for (int q = 0; q < n; ++q) {
          T acc = 0;
#pragma unroll
          for (int i = 0; i < n; ++i) acc += (pass == 0 ? P.A1[q * n + i] : P.A2[q * n + i]) * u[i];
          sa[c * ND + base + stride * q] = acc;
        }
      }
      __syncthreads();
    }
    if (pass == 0) {
      for (int idx = tid; idx < nblk * ND; idx += NT) {
        const long long item = item0 + idx / ND;
        const T jxw = cellG[(size_t)(item / ncomp) * (NS + 1) * ND + (size_t)NS * ND + (idx % ND)];
        sa[idx] = invert ? sa[idx] / jxw : sa[idx] * jxw;
      }
      __syncthreads();
    }
  }
  for (int idx = tid; idx < nblk * ND; idx += NT) {
    const size_t o = (size_t)item0 * ND + idx;
    dst[o] = add ? dst[o] + sa[idx] : sa[idx];
  }
}

template <int dim, int n, typename T>
int launch_mass_general(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse) {
  dgx_mf *mf = op->mf;
  const long long items = mf->n_owned * op->ncomp;
  if (items == 0) return DGX_OK;
  dgx_mf::GeneralMetric *gm = nullptr;
  if (int rc = get_metric(mf, n, &gm)) return rc;
  const Shape1D &s = get_shape(n - 1, n);
  MassGenParams<T, n> P;
  if (!inverse) {
    for (int q = 0; q < n; ++q)
      for (int i = 0; i < n; ++i) { P.A1[q * n + i] = (T)s.S[q * n + i]; P.A2[i * n + q] = (T)s.S[q * n + i]; }
  } else {   // S^-1 by Gauss-Jordan with partial pivoting (n <= 7)
    long double A[MAX_N * MAX_N], B[MAX_N * MAX_N];
    for (int i = 0; i < n * n; ++i) { A[i] = s.S[i]; B[i] = 0; }
    for (int i = 0; i < n; ++i) B[i * n + i] = 1;
    for (int col = 0; col < n; ++col) {
      int piv = col;
      for (int r = col + 1; r < n; ++r) if (fabsl(A[r * n + col]) > fabsl(A[piv * n + col])) piv = r;
      for (int k = 0; k < n; ++k) { std::swap(A[col * n + k], A[piv * n + k]); std::swap(B[col * n + k], B[piv * n + k]); }
      const long double d = 1 / A[col * n + col];
      for (int k = 0; k < n; ++k) { A[col * n + k] *= d; B[col * n + k] *= d; }
      for (int r = 0; r < n; ++r) {
        if (r == col) continue;
        const long double f = A[r * n + col];
        for (int k = 0; k < n; ++k) { A[r * n + k] -= f * A[col * n + k]; B[r * n + k] -= f * B[col * n + k]; }
      }
    }
    // B = S^-1 with B[i][q]; first sweeps apply S^-T: out[q] = sum_i B[i][q] in[i]; second sweeps S^-1: out[i] = sum_q B[i][q] in[q]
    for (int q = 0; q < n; ++q)
      for (int i = 0; i < n; ++i) { P.A1[q * n + i] = (T)B[i * n + q]; P.A2[i * n + q] = (T)B[i * n + q]; }
  }
  constexpr int CPB = gen_cpb<dim, n, T>(), NT = CPB * ipow(n, dim - 1);
  mass_general_kernel<dim, n, T><<<(int)((items + CPB - 1) / CPB), NT, 0, mf->ctx->stream>>>(P, items, op->ncomp, (const T *)gm->cellG, inverse ? 1 : 0,
                                                                                            (const T *)src->d, (T *)dst->d, add ? 1 : 0);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int mass_general_by_degree(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse) {
  switch (op->degree) {
    case 1: return launch_mass_general<dim, 2, T>(op, dst, src, add, inverse);
    case 2: return launch_mass_general<dim, 3, T>(op, dst, src, add, inverse);
    case 3: return launch_mass_general<dim, 4, T>(op, dst, src, add, inverse);
    case 4: return launch_mass_general<dim, 5, T>(op, dst, src, add, inverse);
    case 5: return launch_mass_general<dim, 6, T>(op, dst, src, add, inverse);
    case 6: return launch_mass_general<dim, 7, T>(op, dst, src, add, inverse);
    default: set_error("mass operator on a general mesh: degree must be in 1..6"); return DGX_ERR_UNSUPPORTED;
  }
}
}  // namespace

int mass_general_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse) {
  dgx_mf *mf = op->mf;
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? mass_general_by_degree<2, double>(op, dst, src, add, inverse) : mass_general_by_degree<2, float>(op, dst, src, add, inverse);
  return mf->dtype == DGX_F64 ? mass_general_by_degree<3, double>(op, dst, src, add, inverse) : mass_general_by_degree<3, float>(op, dst, src, add, inverse);
}

// single-CTA coarse solve on a metric level (same conditions as pressure_coarse_cg checks for Cartesian levels)
int pressure_general_coarse_cg(dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info, double *d_res,
                               int *d_count, int log_cap) {
  dgx_mf *mf = op->mf;
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? general_coarse_by_degree<2, double>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap)
                                                 : general_coarse_by_degree<2, float>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
  return mf->dtype == DGX_F64 ? general_coarse_by_degree<3, double>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap)
                              : general_coarse_by_degree<3, float>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
}

// host tables of a metric level (introspection / tests; pure host logic, works without a device): sizes first, then the data
int general_metric_host_tables(const dgx_mf *mf, int n, std::vector<double> &cellG, std::vector<double> &faceA) {
  if (!mf->general) { set_error("not a non-Cartesian level"); return DGX_ERR_INVALID; }
  if (!build_metric(mf, n, cellG, faceA)) { set_error("general mesh: a cell has a non-positive Jacobian determinant"); return DGX_ERR_INVALID; }
  return DGX_OK;
}

void general_metric_destroy(dgx_mf *mf) {
  for (auto &kv : mf->gmetric) {
    if (kv.second.cellG) cudaFree(kv.second.cellG);
    if (kv.second.faceA) cudaFree(kv.second.faceA);
  }
  mf->gmetric.clear();
}

long long general_metric_bytes(dgx_mf *mf, int n) {
  auto it = mf->gmetric.find(n);
  return it == mf->gmetric.end() ? 0 : (long long)it->second.bytes;
}

}  // namespace dgx
