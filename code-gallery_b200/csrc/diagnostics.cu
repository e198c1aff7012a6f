// Output and diagnostics of the time loop, kept on the device:
//   * compute_lift_and_drag (NS.cc:2642-2708): integral of (1/Re grad u - p I)(-n) over the boundary faces with one boundary id
//     (the reference: 4, the cylinder), QGauss(degree_p + 2) face quadrature, FEFaceValues-style point evaluation;
//   * output_results (NS.cc:2608-2633): DataOut::build_patches(MappingQ1, 1) = one patch per cell with the 2^dim cell vertices as
//     points, data v, p, v_old and PostprocessorVorticity (NS.cc:115-160: curl of v from the velocity gradients) at these points,
//     written as an UnstructuredGrid .vtu (one piece file per rank plus a .pvtu index when there are several ranks).
// Neither is on the operator hot path: the kernels are plain point evaluations (one thread per face quadrature point / per
// patch vertex, run-time degree); what matters is that u and p never leave the device except as the patch data to be written.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "dgx_internal.h"

namespace dgx {
namespace {

constexpr int kMaxNq = MAX_N + 1;

struct FaceEvalTables {
  double Sv[kMaxNq * MAX_N], Dv[kMaxNq * MAX_N];   // velocity basis and its derivative at the face quadrature points: [q][i]
  double xq[kMaxNq];                                // face quadrature points (non-Cartesian cells: the metric is evaluated there)
  double gv[2][MAX_N];                              // l_i'(s) of the velocity basis
  double Sp[kMaxNq * MAX_N];                        // pressure basis at the face quadrature points
  double w[kMaxNq];
  double h[MAX_DIM];
  int nv, np, nq, dim;
  double inv_Re;
};

// inverse Jacobian inv[e][i] = d xi_e / d x_i and det J of the multilinear (MappingQ1) cell with vertices X[2^dim][dim] at xi
__device__ __forceinline__ double q1_jacobian(int dim, const double *X, const double *xi, double inv[3][3]) {
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int v = 0; v < (1 << dim); ++v)
    for (int e = 0; e < dim; ++e) {
      double dN = 1;
      for (int d = 0; d < dim; ++d) {
        const int b = (v >> d) & 1;
        dN *= d == e ? (b ? 1.0 : -1.0) : (b ? xi[d] : 1.0 - xi[d]);
      }
      for (int i = 0; i < dim; ++i) J[i][e] += X[v * dim + i] * dN;
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

// item = (boundary face of the list, face quadrature point).  FE_DGQ is nodal at the cell boundary, so face values need the
// face layer only; the normal derivative needs all layers of the line through the point.  cellX == nullptr: Cartesian cells;
// otherwise cellX[cell][2^dim][dim] are the vertices of the MappingQ1 cells and normal, JxW and J^-1 are evaluated at the point.
template <typename T>
__global__ void __launch_bounds__(128) lift_drag_kernel(const __grid_constant__ FaceEvalTables P, const int *__restrict__ faces, int n_faces,
                                                        const T *__restrict__ u, const T *__restrict__ p, const double *__restrict__ cellX,
                                                        double *partial, double *result) {
  const int dim = P.dim, nv = P.nv, np = P.np, nq = P.nq;
  const int nqf = dim == 3 ? nq * nq : nq;
  int ndv = 1, ndp = 1;
  for (int d = 0; d < dim; ++d) { ndv *= nv; ndp *= np; }
  double drag = 0.0, lift = 0.0;
  const long long items = (long long)n_faces * nqf;
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int fi = (int)(it / nqf), q = (int)(it - (long long)fi * nqf);
    const int cell = faces[2 * fi], f = faces[2 * fi + 1], d = f >> 1, s = f & 1;
    const int e1 = d == 0 ? 1 : 0, e2 = d == 2 ? 1 : 2;      // tangential directions, ascending (e2 unused in 2-D)
    const int q1 = q % nq, q2 = q / nq;
    int strv[3] = {1, nv, nv * nv}, strp[3] = {1, np, np * np};
    // pressure value
    double pv = 0.0;
    {
      const T *pc = p + (size_t)cell * ndp + (size_t)(s ? np - 1 : 0) * strp[d];
      for (int b = 0; b < (dim == 3 ? np : 1); ++b)
        for (int a = 0; a < np; ++a) {
          const double wgt = P.Sp[q1 * np + a] * (dim == 3 ? P.Sp[q2 * np + b] : 1.0);
          pv += wgt * (double)pc[a * strp[e1] + (dim == 3 ? b * strp[e2] : 0)];
        }
    }
    // reference gradients of the velocity components at the face point: along the normal direction d from all layers of the line,
    // along the tangential directions from the face layer
    const double ns = s ? 1.0 : -1.0;
    double gref[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // [component][reference direction]
    for (int c = 0; c < dim; ++c) {
      const T *uc = u + ((size_t)cell * dim + c) * ndv;
      const T *layer = uc + (size_t)(s ? nv - 1 : 0) * strv[d];
      for (int b = 0; b < (dim == 3 ? nv : 1); ++b)
        for (int a = 0; a < nv; ++a) {
          const double sa = P.Sv[q1 * nv + a], sb = dim == 3 ? P.Sv[q2 * nv + b] : 1.0;
          const T *line = uc + a * strv[e1] + (dim == 3 ? b * strv[e2] : 0);
          double acc = 0.0;
          for (int i = 0; i < nv; ++i) acc += P.gv[s][i] * (double)line[(size_t)i * strv[d]];
          gref[c][d] += sa * sb * acc;
          const double val = (double)layer[a * strv[e1] + (dim == 3 ? b * strv[e2] : 0)];
          gref[c][e1] += P.Dv[q1 * nv + a] * sb * val;
          if (dim == 3) gref[c][e2] += sa * P.Dv[q2 * nv + b] * val;
        }
    }
    // forces = (1/Re grad u - p I) (-n_out) JxW   (NS.cc:2679-2688)
    double force[3] = {0, 0, 0};
    if (cellX == nullptr) {   // Cartesian: n_out = ns e_d, grad = diag(1/h) grad_xi
      double JxW = P.w[q1] * P.h[e1];
      if (dim == 3) JxW *= P.w[q2] * P.h[e2];
      for (int c = 0; c < dim; ++c) force[c] = -ns * (P.inv_Re * gref[c][d] / P.h[d] - (c == d ? pv : 0.0)) * JxW;
    } else {
      double xi[3] = {0, 0, 0}, inv[3][3], w = P.w[q1];
      xi[d] = s; xi[e1] = P.xq[q1];
      if (dim == 3) { xi[e2] = P.xq[q2]; w *= P.w[q2]; }
      const double det = q1_jacobian(dim, cellX + (size_t)cell * (1 << dim) * dim, xi, inv);
      double norm = 0;
      for (int i = 0; i < dim; ++i) norm += inv[d][i] * inv[d][i];
      norm = sqrt(norm);
      const double JxW = det * norm * w;
      for (int c = 0; c < dim; ++c) {
        double f = 0;
        for (int k = 0; k < dim; ++k) {
          double g = 0;   // d u_c / d x_k = sum_e d u_c / d xi_e * d xi_e / d x_k
          for (int e = 0; e < dim; ++e) g += gref[c][e] * inv[e][k];
          const double nk = ns * inv[d][k] / norm;
          f += (P.inv_Re * g - (c == k ? pv : 0.0)) * (-nk);
        }
        force[c] = f * JxW;
      }
    }
    drag += force[0];
    lift += force[1];
  }
  __shared__ double sm[2][4];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) { drag += __shfl_down_sync(0xffffffffu, drag, o); lift += __shfl_down_sync(0xffffffffu, lift, o); }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = drag; sm[1][threadIdx.x >> 5] = lift; }
  __syncthreads();
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = sm[0][0] + sm[0][1] + sm[0][2] + sm[0][3];
    partial[2 * blockIdx.x + 1] = sm[1][0] + sm[1][1] + sm[1][2] + sm[1][3];
    __threadfence();
    const unsigned int t = atomicInc(scratch_ticket(partial), gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {   // fixed-order combination of the block partials
    __threadfence();
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += ((volatile double *)partial)[2 * i]; s1 += ((volatile double *)partial)[2 * i + 1]; }
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_down_sync(0xffffffffu, s0, o); s1 += __shfl_down_sync(0xffffffffu, s1, o); }
    if (threadIdx.x == 0) { result[0] = s0; result[1] = s1; }
  }
}

struct PatchTables {
  double gv[2][MAX_N];   // l_i'(0), l_i'(1) of the velocity basis
  double ih[MAX_DIM];
  int nv, np, dim;
};

// item = (cell, vertex); out[(cell * 2^dim + vertex) * nf + field], fields = v (dim), p, v_old (dim), vorticity (1 or 3)
template <typename T>
__global__ void __launch_bounds__(128) patch_kernel(const __grid_constant__ PatchTables P, long long n_cells, const T *__restrict__ u, const T *__restrict__ p,
                                                    const T *__restrict__ uold, const double *__restrict__ cellX, double *__restrict__ out) {
  const int dim = P.dim, nv = P.nv, np = P.np, NV = 1 << dim, nvort = dim == 2 ? 1 : 3, nf = 2 * dim + 1 + nvort;
  int ndv = 1, ndp = 1;
  for (int d = 0; d < dim; ++d) { ndv *= nv; ndp *= np; }
  const int strv[3] = {1, nv, nv * nv}, strp[3] = {1, np, np * np};
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < n_cells * NV; it += (long long)gridDim.x * blockDim.x) {
    const long long cell = it / NV;
    const int v = (int)(it - cell * NV);
    int offv = 0, offp = 0;
    for (int d = 0; d < dim; ++d) { const int s = (v >> d) & 1; offv += s * (nv - 1) * strv[d]; offp += s * (np - 1) * strp[d]; }
    double *o = out + (size_t)it * nf;
    double grad[3][3];
    for (int c = 0; c < dim; ++c) {
      const T *uc = u + ((size_t)cell * dim + c) * ndv;
      o[c] = (double)uc[offv];
      o[dim + 1 + c] = (double)uold[((size_t)cell * dim + c) * ndv + offv];
      for (int d = 0; d < dim; ++d) {
        const int s = (v >> d) & 1;
        const T *line = uc + offv - s * (nv - 1) * strv[d];
        double acc = 0.0;
        for (int i = 0; i < nv; ++i) acc += P.gv[s][i] * (double)line[(size_t)i * strv[d]];
        grad[c][d] = cellX ? acc : acc * P.ih[d];
      }
    }
    if (cellX) {   // MappingQ1 cell: grad_x = J^-T grad_xi at the vertex
      double xi[3] = {0, 0, 0}, inv[3][3];
      for (int d = 0; d < dim; ++d) xi[d] = (v >> d) & 1;
      q1_jacobian(dim, cellX + (size_t)cell * NV * dim, xi, inv);
      for (int c = 0; c < dim; ++c) {
        double g[3] = {0, 0, 0};
        for (int k = 0; k < dim; ++k)
          for (int e = 0; e < dim; ++e) g[k] += grad[c][e] * inv[e][k];
        for (int k = 0; k < dim; ++k) grad[c][k] = g[k];
      }
    }
    o[dim] = (double)p[(size_t)cell * ndp + offp];
    if (dim == 2) o[2 * dim + 1] = grad[1][0] - grad[0][1];                    // NS.cc:146-148
    else {                                                                     // NS.cc:150-155
      o[2 * dim + 1] = grad[2][1] - grad[1][2];
      o[2 * dim + 2] = grad[0][2] - grad[2][0];
      o[2 * dim + 3] = grad[1][0] - grad[0][1];
    }
  }
}

struct IndicatorTables {
  double S[kMaxNq * MAX_N], D[kMaxNq * MAX_N];   // velocity basis and derivative at the cell quadrature points: [q][i]
  double xq[kMaxNq], w[kMaxNq];
  double h[MAX_DIM];
  int n, nq, dim;
};

// refine_mesh's indicator (NS.cc:2727-2752): one warp per cell, lanes over the quadrature points of QGauss<dim>(nq);
// out[cell] = diameter^2 * sum_q (d u_1 / d x_0 - d u_0 / d x_1)^2 JxW   (the reference takes this component in 3-D too)
template <typename T>
__global__ void __launch_bounds__(128) vorticity_indicator_kernel(const __grid_constant__ IndicatorTables P, long long n_cells, const T *__restrict__ u,
                                                                  const double *__restrict__ cellX, float *__restrict__ out) {
  const int dim = P.dim, n = P.n, nq = P.nq, lane = threadIdx.x & 31;
  const long long cell = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cell >= n_cells) return;
  int nd = 1, nqt = 1;
  for (int d = 0; d < dim; ++d) { nd *= n; nqt *= nq; }
  double acc = 0.0;
  for (int q = lane; q < nqt; q += 32) {
    int qi[3] = {q % nq, (q / nq) % nq, dim == 3 ? q / (nq * nq) : 0};
    double gref[2][3] = {{0, 0, 0}, {0, 0, 0}};   // reference gradients of the components 0 and 1
    for (int c = 0; c < 2; ++c) {
      const T *uc = u + ((size_t)cell * dim + c) * nd;
      for (int k = 0; k < (dim == 3 ? n : 1); ++k)
        for (int j = 0; j < n; ++j)
          for (int i = 0; i < n; ++i) {
            const double v = (double)uc[i + n * (j + n * k)];
            const double sx = P.S[qi[0] * n + i], sy = P.S[qi[1] * n + j], sz = dim == 3 ? P.S[qi[2] * n + k] : 1.0;
            gref[c][0] += P.D[qi[0] * n + i] * sy * sz * v;
            gref[c][1] += sx * P.D[qi[1] * n + j] * sz * v;
            if (dim == 3) gref[c][2] += sx * sy * P.D[qi[2] * n + k] * v;
          }
    }
    double w = P.w[qi[0]] * P.w[qi[1]] * (dim == 3 ? P.w[qi[2]] : 1.0), d10, d01, JxW;
    if (cellX == nullptr) {
      d10 = gref[1][0] / P.h[0]; d01 = gref[0][1] / P.h[1];
      JxW = w * P.h[0] * P.h[1] * (dim == 3 ? P.h[2] : 1.0);
    } else {
      double xi[3] = {P.xq[qi[0]], P.xq[qi[1]], dim == 3 ? P.xq[qi[2]] : 0.0}, inv[3][3];
      const double det = q1_jacobian(dim, cellX + (size_t)cell * (1 << dim) * dim, xi, inv);
      d10 = 0; d01 = 0;
      for (int e = 0; e < dim; ++e) { d10 += gref[1][e] * inv[e][0]; d01 += gref[0][e] * inv[e][1]; }
      JxW = det * w;
    }
    const double vort = d10 - d01;
    acc += vort * vort * JxW;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    double diam2 = 0;
    if (cellX == nullptr) {
      for (int d = 0; d < dim; ++d) diam2 += P.h[d] * P.h[d];
    } else {   // TriaAccessor::diameter(): the longest diagonal of the quadrilateral / hexahedron
      const double *X = cellX + (size_t)cell * (1 << dim) * dim;
      const int NV = 1 << dim;
      for (int v = 0; v < NV / 2; ++v) {
        double s2 = 0;
        for (int d = 0; d < dim; ++d) { const double t = X[v * dim + d] - X[(NV - 1 - v) * dim + d]; s2 += t * t; }
        diam2 = fmax(diam2, s2);
      }
    }
    out[cell] = (float)(diam2 * acc);
  }
}

int check_pair(const dgx_vec *u, const dgx_vec *p) {
  DGX_REQUIRE(u && p, "null argument");
  DGX_REQUIRE(u->mf == p->mf, "velocity and pressure must live on the same MatrixFree");
  DGX_REQUIRE(u->ncomp == u->mf->dim && p->ncomp == 1, "velocity (dim components) and pressure (scalar) expected");
  DGX_REQUIRE(u->dtype == p->dtype, "dtype mismatch");
  DGX_REQUIRE(u->degree >= 1 && p->degree >= 1, "degree >= 1");
  return DGX_OK;
}

}  // namespace

// boundary faces (local cell, face number) with the given boundary id, device copy cached per MatrixFree and id
int boundary_face_list(dgx_mf *mf, int boundary_id, const int **d_faces, int *n_faces) {
  auto it = mf->bfaces.find(boundary_id);
  if (it == mf->bfaces.end()) {
    std::vector<int> list;
    const int nf = 2 * mf->dim;
    for (long long c = 0; c < mf->n_owned; ++c)
      for (int f = 0; f < nf; ++f)
        if (mf->nbr_host[(size_t)f * mf->n_owned + c] == -1 - boundary_id) { list.push_back((int)c); list.push_back(f); }
    dgx_mf::BFaces bf;
    bf.n = (int)(list.size() / 2);
    if (bf.n) {
      DGX_CUDA_CHECK(cudaMalloc(&bf.d, list.size() * sizeof(int)));
      DGX_CUDA_CHECK(cudaMemcpy(bf.d, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    it = mf->bfaces.emplace(boundary_id, bf).first;
  }
  *d_faces = it->second.d;
  *n_faces = it->second.n;
  return DGX_OK;
}

// vertices of the owned cells of a non-Cartesian level on the device, [cell][2^dim][dim] (deal.II vertex order); nullptr on Cartesian levels
int cell_vertex_table(dgx_mf *mf, const double **d_X) {
  *d_X = nullptr;
  if (!mf->general) return DGX_OK;
  if (!mf->d_cellX && mf->n_owned) {
    const Mesh &m = mf->mesh->m;
    const int dim = mf->dim, NV = 1 << dim;
    const std::vector<double> &V = m.vertices[mf->level];
    const int nx = m.cells_dir(mf->level, 0) + 1, ny = m.cells_dir(mf->level, 1) + 1;
    std::vector<double> X((size_t)mf->n_owned * NV * dim);
    for (long long c = 0; c < mf->n_owned; ++c) {
      int cc[MAX_DIM] = {0, 0, 0};
      m.coords_of(mf->level, mf->first_cell + c, cc);
      for (int v = 0; v < NV; ++v) {
        const int ix = cc[0] + (v & 1), iy = cc[1] + ((v >> 1) & 1), iz = dim == 3 ? cc[2] + ((v >> 2) & 1) : 0;
        const size_t idx = (size_t)ix + (size_t)nx * ((size_t)iy + (size_t)ny * iz);
        for (int i = 0; i < dim; ++i) X[((size_t)c * NV + v) * dim + i] = V[idx * dim + i];
      }
    }
    DGX_CUDA_CHECK(cudaMalloc(&mf->d_cellX, X.size() * sizeof(double)));
    DGX_CUDA_CHECK(cudaMemcpy(mf->d_cellX, X.data(), X.size() * sizeof(double), cudaMemcpyHostToDevice));
    mf->h_cellX = std::move(X);
  }
  *d_X = mf->d_cellX;
  return DGX_OK;
}

}  // namespace dgx

using namespace dgx;

extern "C" {

int dgx_compute_lift_and_drag(const dgx_vec *u, const dgx_vec *p, double Re, int boundary_id, double *lift, double *drag) {
  if (int rc = check_pair(u, p)) return rc;
  DGX_REQUIRE(lift && drag, "null argument");
  dgx_mf *mf = u->mf;
  dgx_ctx *ctx = mf->ctx;
  DGX_REQUIRE(ctx->d_scratch, "no device");
  const int nq = p->degree + 2;   // QGauss<dim - 1>(EquationData::degree_p + 2), NS.cc:2643
  DGX_REQUIRE(nq <= kMaxNq && u->degree + 1 <= MAX_N, "degree too large");
  DGX_CUDA_CHECK(cudaSetDevice(ctx->device));
  const int *d_faces = nullptr;
  int n_faces = 0;
  if (int rc = boundary_face_list(mf, boundary_id, &d_faces, &n_faces)) return rc;
  FaceEvalTables P;
  std::memset(&P, 0, sizeof(P));
  const Shape1D &sv = get_shape(u->degree, nq), &sp = get_shape(p->degree, nq);
  P.nv = u->degree + 1; P.np = p->degree + 1; P.nq = nq; P.dim = mf->dim; P.inv_Re = 1.0 / Re;
  for (int q = 0; q < nq; ++q) {
    for (int i = 0; i < P.nv; ++i) { P.Sv[q * P.nv + i] = sv.S[q * P.nv + i]; P.Dv[q * P.nv + i] = sv.D[q * P.nv + i]; }
    for (int i = 0; i < P.np; ++i) P.Sp[q * P.np + i] = sp.S[q * P.np + i];
    P.w[q] = sv.wq[q]; P.xq[q] = sv.xq[q];
  }
  const double *d_X = nullptr;
  if (int rc = cell_vertex_table(mf, &d_X)) return rc;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < P.nv; ++i) P.gv[s][i] = sv.g[s][i];
  for (int d = 0; d < MAX_DIM; ++d) P.h[d] = d < mf->dim ? mf->h[d] : 1.0;
  double *partial = ctx->d_scratch + kScratchPartial, *result = ctx->d_scratch;
  const long long items = (long long)n_faces * (mf->dim == 3 ? nq * nq : nq);
  const int grid = (int)std::max<long long>(1, std::min<long long>((items + 127) / 128, 148 * 4));
  if (u->dtype == DGX_F64)
    lift_drag_kernel<double><<<grid, 128, 0, ctx->stream>>>(P, d_faces, n_faces, (const double *)u->d, (const double *)p->d, d_X, partial, result);
  else
    lift_drag_kernel<float><<<grid, 128, 0, ctx->stream>>>(P, d_faces, n_faces, (const float *)u->d, (const float *)p->d, d_X, partial, result);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  if (!mf->replicated) {   // Utilities::MPI::sum, NS.cc:2698-2699
    if (int rc = allreduce_sum(ctx, result, 2)) return rc;
  }
  DGX_CUDA_CHECK(cudaMemcpyAsync(ctx->h_scratch, result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  *drag = ctx->h_scratch[0];
  *lift = ctx->h_scratch[1];
  return DGX_OK;
}

int dgx_estimate_vorticity_indicator(const dgx_vec *u, int n_q_points_1d, float *out_host, long long capacity) {
  DGX_REQUIRE(u && out_host, "null argument");
  dgx_mf *mf = u->mf;
  dgx_ctx *ctx = mf->ctx;
  DGX_REQUIRE(ctx->d_scratch, "no device");
  DGX_REQUIRE(u->ncomp == mf->dim && u->degree >= 1 && u->degree + 1 <= MAX_N, "velocity vector (dim components) expected");
  DGX_REQUIRE(n_q_points_1d >= 1 && n_q_points_1d <= kMaxNq, "number of quadrature points");
  DGX_REQUIRE(capacity >= mf->n_owned, "output buffer too small");
  if (mf->n_owned == 0) return DGX_OK;
  DGX_CUDA_CHECK(cudaSetDevice(ctx->device));
  IndicatorTables P;
  std::memset(&P, 0, sizeof(P));
  const Shape1D &sv = get_shape(u->degree, n_q_points_1d);
  P.n = u->degree + 1; P.nq = n_q_points_1d; P.dim = mf->dim;
  for (int q = 0; q < P.nq; ++q) {
    for (int i = 0; i < P.n; ++i) { P.S[q * P.n + i] = sv.S[q * P.n + i]; P.D[q * P.n + i] = sv.D[q * P.n + i]; }
    P.xq[q] = sv.xq[q]; P.w[q] = sv.wq[q];
  }
  for (int d = 0; d < MAX_DIM; ++d) P.h[d] = d < mf->dim ? mf->h[d] : 1.0;
  const double *d_X = nullptr;
  if (int rc = cell_vertex_table(mf, &d_X)) return rc;
  float *d_out = nullptr;
  DGX_CUDA_CHECK(cudaMalloc(&d_out, (size_t)mf->n_owned * sizeof(float)));
  const int grid = (int)((mf->n_owned + 3) / 4);
  if (u->dtype == DGX_F64) vorticity_indicator_kernel<double><<<grid, 128, 0, ctx->stream>>>(P, mf->n_owned, (const double *)u->d, d_X, d_out);
  else vorticity_indicator_kernel<float><<<grid, 128, 0, ctx->stream>>>(P, mf->n_owned, (const float *)u->d, d_X, d_out);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, (size_t)mf->n_owned * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_out);
  if (e != cudaSuccess) { set_error(std::string("vorticity indicator: ") + cudaGetErrorString(e)); return DGX_ERR_CUDA; }
  return DGX_OK;
}

// GridRefinement::refine_and_coarsen_fixed_number (NS.cc:2771) on one rank: the top_fraction * N cells with the largest indicators are
// flagged for refinement (value >= the threshold, as GridRefinement::refine), the bottom_fraction * N with the smallest for coarsening
// (value <= the threshold, as GridRefinement::coarsen); a cell never gets both.  Pure host logic.
int dgx_refine_and_coarsen_fixed_number(const float *indicators, long long n, double top_fraction, double bottom_fraction, signed char *flags) {
  DGX_REQUIRE(indicators && flags && n >= 0, "null argument");
  DGX_REQUIRE(top_fraction >= 0 && bottom_fraction >= 0 && top_fraction + bottom_fraction <= 1.0, "fractions");
  for (long long i = 0; i < n; ++i) flags[i] = 0;
  if (n == 0) return DGX_OK;
  const long long n_refine = (long long)(top_fraction * (double)n), n_coarsen = (long long)(bottom_fraction * (double)n);
  std::vector<float> tmp(indicators, indicators + n);
  if (n_refine > 0) {
    std::nth_element(tmp.begin(), tmp.begin() + (n_refine - 1), tmp.end(), std::greater<float>());
    const float thr = tmp[n_refine - 1];
    for (long long i = 0; i < n; ++i) if (indicators[i] >= thr) flags[i] = 1;
  }
  if (n_coarsen > 0) {
    std::nth_element(tmp.begin(), tmp.begin() + (n_coarsen - 1), tmp.end());
    const float thr = tmp[n_coarsen - 1];
    for (long long i = 0; i < n; ++i) if (indicators[i] <= thr && flags[i] == 0) flags[i] = -1;
  }
  return DGX_OK;
}

int dgx_output_n_fields(int dim) { return 2 * dim + 1 + (dim == 2 ? 1 : 3); }

int dgx_output_patch_data(const dgx_vec *u_n, const dgx_vec *pres_n, const dgx_vec *u_n_minus_1, double *out_host, long long capacity) {
  if (int rc = check_pair(u_n, pres_n)) return rc;
  DGX_REQUIRE(u_n_minus_1 && u_n_minus_1->mf == u_n->mf && u_n_minus_1->degree == u_n->degree && u_n_minus_1->ncomp == u_n->ncomp &&
                  u_n_minus_1->dtype == u_n->dtype,
              "u_n_minus_1 must have the shape of u_n");
  dgx_mf *mf = u_n->mf;
  dgx_ctx *ctx = mf->ctx;
  DGX_REQUIRE(ctx->d_scratch, "no device");
  const int dim = mf->dim, NV = 1 << dim, nf = dgx_output_n_fields(dim);
  const long long total = mf->n_owned * NV * nf;
  DGX_REQUIRE(out_host && capacity >= total, "output buffer too small");
  if (total == 0) return DGX_OK;
  DGX_CUDA_CHECK(cudaSetDevice(ctx->device));
  PatchTables P;
  std::memset(&P, 0, sizeof(P));
  const Shape1D &sv = get_shape(u_n->degree, u_n->degree + 1);
  P.nv = u_n->degree + 1; P.np = pres_n->degree + 1; P.dim = dim;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < P.nv; ++i) P.gv[s][i] = sv.g[s][i];
  for (int d = 0; d < MAX_DIM; ++d) P.ih[d] = d < dim ? 1.0 / mf->h[d] : 1.0;
  const double *d_X = nullptr;
  if (int rc = cell_vertex_table(mf, &d_X)) return rc;
  double *d_out = nullptr;
  DGX_CUDA_CHECK(cudaMalloc(&d_out, (size_t)total * sizeof(double)));
  const int grid = (int)std::max<long long>(1, std::min<long long>((mf->n_owned * NV + 127) / 128, 148 * 8));
  if (u_n->dtype == DGX_F64)
    patch_kernel<double><<<grid, 128, 0, ctx->stream>>>(P, mf->n_owned, (const double *)u_n->d, (const double *)pres_n->d, (const double *)u_n_minus_1->d, d_X, d_out);
  else
    patch_kernel<float><<<grid, 128, 0, ctx->stream>>>(P, mf->n_owned, (const float *)u_n->d, (const float *)pres_n->d, (const float *)u_n_minus_1->d, d_X, d_out);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_out);
  if (e != cudaSuccess) { set_error(std::string("output patches: ") + cudaGetErrorString(e)); return DGX_ERR_CUDA; }
  return DGX_OK;
}

// host side of output_results: pure format work (pieces in the layout DataOutBase::write_vtu uses: every patch owns its points)
int dgx_write_vtu_piece(const char *path, int dim, long long n_cells, const double *points, const double *patch_data) {
  DGX_REQUIRE(path && (dim == 2 || dim == 3) && n_cells >= 0 && (n_cells == 0 || (points && patch_data)), "arguments");
  FILE *f = std::fopen(path, "w");
  if (!f) { set_error(std::string("cannot open ") + path); return DGX_ERR_INVALID; }
  const int NV = 1 << dim, nf = dgx_output_n_fields(dim), nvort = dim == 2 ? 1 : 3;
  const long long np = n_cells * NV;
  std::fprintf(f, "<?xml version=\"1.0\" ?>\n<!--\n# vtk DataFile Version 3.0\n# written by libdgx in the layout of deal.II's DataOut::write_vtu\n-->\n");
  std::fprintf(f, "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n");
  std::fprintf(f, "<Piece NumberOfPoints=\"%lld\" NumberOfCells=\"%lld\">\n", np, n_cells);
  std::fprintf(f, "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n");
  for (long long i = 0; i < np; ++i)
    std::fprintf(f, "%.17g %.17g %.17g\n", points[i * dim], points[i * dim + 1], dim == 3 ? points[i * dim + 2] : 0.0);
  std::fprintf(f, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n");
  static const int vtk2[4] = {0, 1, 3, 2}, vtk3[8] = {0, 1, 3, 2, 4, 5, 7, 6};   // deal.II vertex order -> VTK_QUAD / VTK_HEXAHEDRON
  for (long long c = 0; c < n_cells; ++c) {
    for (int v = 0; v < NV; ++v) std::fprintf(f, "%lld ", c * NV + (dim == 2 ? vtk2[v] : vtk3[v]));
    std::fputc('\n', f);
  }
  std::fprintf(f, "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n");
  for (long long c = 0; c < n_cells; ++c) std::fprintf(f, "%lld\n", (c + 1) * NV);
  std::fprintf(f, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
  for (long long c = 0; c < n_cells; ++c) std::fprintf(f, "%d\n", dim == 2 ? 9 : 12);
  std::fprintf(f, "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n");
  auto vec3 = [&](const char *name, int first, int ncomp) {
    std::fprintf(f, "<DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"3\" format=\"ascii\">\n", name);
    for (long long i = 0; i < np; ++i) {
      const double *o = patch_data + i * nf + first;
      std::fprintf(f, "%.17g %.17g %.17g\n", o[0], ncomp > 1 ? o[1] : 0.0, ncomp > 2 ? o[2] : 0.0);
    }
    std::fprintf(f, "</DataArray>\n");
  };
  auto scalar = [&](const char *name, int first) {
    std::fprintf(f, "<DataArray type=\"Float64\" Name=\"%s\" format=\"ascii\">\n", name);
    for (long long i = 0; i < np; ++i) std::fprintf(f, "%.17g\n", patch_data[i * nf + first]);
    std::fprintf(f, "</DataArray>\n");
  };
  vec3("v", 0, dim);                  // NS.cc:2612-2616
  scalar("p", dim);                   // NS.cc:2617-2618
  vec3("v_old", dim + 1, dim);        // NS.cc:2620-2622
  if (nvort == 1) scalar("vorticity", 2 * dim + 1);   // NS.cc:2625-2626, component_is_scalar in 2-D
  else vec3("vorticity", 2 * dim + 1, 3);
  std::fprintf(f, "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
  const bool ok = std::fclose(f) == 0;
  if (!ok) { set_error(std::string("write failed: ") + path); return DGX_ERR_INVALID; }
  return DGX_OK;
}

int dgx_output_results(const dgx_vec *u_n, const dgx_vec *pres_n, const dgx_vec *u_n_minus_1, const char *path) {
  if (int rc = check_pair(u_n, pres_n)) return rc;
  DGX_REQUIRE(path, "null argument");
  dgx_mf *mf = u_n->mf;
  dgx_ctx *ctx = mf->ctx;
  const Mesh &m = mf->mesh->m;
  const int dim = mf->dim, NV = 1 << dim, nf = dgx_output_n_fields(dim);
  std::vector<double> data((size_t)mf->n_owned * NV * nf), pts((size_t)mf->n_owned * NV * dim);
  if (int rc = dgx_output_patch_data(u_n, pres_n, u_n_minus_1, data.data(), (long long)data.size())) return rc;
  if (mf->general) pts = mf->h_cellX;   // filled by dgx_output_patch_data above
  else
    for (long long c = 0; c < mf->n_owned; ++c) {
      int cc[MAX_DIM] = {0, 0, 0};
      m.coords_of(mf->level, mf->first_cell + c, cc);
      for (int v = 0; v < NV; ++v)
        for (int d = 0; d < dim; ++d) pts[((size_t)c * NV + v) * dim + d] = m.lo[d] + (cc[d] + ((v >> d) & 1)) * mf->h[d];
    }
  const bool parallel = ctx->nranks > 1 && !mf->replicated;
  if (mf->replicated && ctx->rank != 0) return DGX_OK;   // every rank holds the same cells: rank 0 writes
  std::string base(path), piece = base;
  if (parallel) {   // one piece per rank + an index written by rank 0 (the reference writes one file through MPI I/O, NS.cc:2632)
    const size_t dot = base.rfind(".vtu");
    const std::string stem = dot == std::string::npos ? base : base.substr(0, dot);
    char suf[32];
    std::snprintf(suf, sizeof(suf), ".%04d.vtu", ctx->rank);
    piece = stem + suf;
    if (ctx->rank == 0) {
      FILE *f = std::fopen((stem + ".pvtu").c_str(), "w");
      if (!f) { set_error("cannot open " + stem + ".pvtu"); return DGX_ERR_INVALID; }
      std::fprintf(f, "<?xml version=\"1.0\" ?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n");
      std::fprintf(f, "<PPointData Scalars=\"scalars\">\n<PDataArray type=\"Float64\" Name=\"v\" NumberOfComponents=\"3\" format=\"ascii\"/>\n");
      std::fprintf(f, "<PDataArray type=\"Float64\" Name=\"p\" format=\"ascii\"/>\n<PDataArray type=\"Float64\" Name=\"v_old\" NumberOfComponents=\"3\" format=\"ascii\"/>\n");
      if (dim == 2) std::fprintf(f, "<PDataArray type=\"Float64\" Name=\"vorticity\" format=\"ascii\"/>\n");
      else std::fprintf(f, "<PDataArray type=\"Float64\" Name=\"vorticity\" NumberOfComponents=\"3\" format=\"ascii\"/>\n");
      std::fprintf(f, "</PPointData>\n<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n");
      const size_t slash = stem.find_last_of('/');
      const std::string leaf = slash == std::string::npos ? stem : stem.substr(slash + 1);
      for (int r = 0; r < ctx->nranks; ++r) std::fprintf(f, "<Piece Source=\"%s.%04d.vtu\"/>\n", leaf.c_str(), r);
      std::fprintf(f, "</PUnstructuredGrid>\n</VTKFile>\n");
      std::fclose(f);
    }
  }
  return dgx_write_vtu_piece(piece.c_str(), dim, mf->n_owned, pts.data(), data.data());
}

}  // extern "C"
