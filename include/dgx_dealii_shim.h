// dgx_dealii_shim.h -- header-only C++17 mirror of the deal.II surface that NavierStokes_TRBDF2_DG uses on its hot path,
// implemented on the C ABI of dgx.h (libdgx.so).  The reference has no FFI: its operator is consumed through C++ duck
// typing (MatrixFreeOperators::Base, SolverCG, SolverGMRES, PreconditionMG, LinearAlgebra::distributed::Vector).  With
// this header a driver written like NavierStokesProjection<dim>::diffusion_step / projection_step / project_grad / run
// (NS.cc:2424-2577, 2928-3010) keeps its statements; examples/trbdf2_tgv.cpp is that driver for the periodic Taylor-Green
// vortex.  "NS.cc" = NavierStokes_TRBDF2_DG/navier_stokes_TRBDF2_DG.cc of the code gallery.
//
// Same names, argument meaning and error behaviour as the reference: failures throw (AssertThrow -> std::runtime_error,
// SolverControl::NoConvergence -> dgxshim::SolverControl::NoConvergence, caught by main like NS.cc:3055-3077).
#ifndef DGX_DEALII_SHIM_H
#define DGX_DEALII_SHIM_H

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "dgx.h"

namespace dgxshim {

struct ExcDGX : std::runtime_error { using std::runtime_error::runtime_error; };

// SolverControl (deal.II lac/solver_control.h): absolute tolerance + iteration cap, keeps the last step / value
class SolverControl {
public:
  struct NoConvergence : std::runtime_error {
    unsigned int last_step;
    double last_residual;
    NoConvergence(unsigned int s, double r, const std::string &m) : std::runtime_error(m), last_step(s), last_residual(r) {}
  };
  SolverControl(unsigned int max_iterations, double tolerance) : max_it_(max_iterations), tol_(tolerance) {}   // NS.cc:2450, 2486, 2574
  unsigned int max_steps() const { return max_it_; }
  double tolerance() const { return tol_; }
  unsigned int last_step() const { return last_step_; }
  double last_value() const { return last_value_; }
  void record(int its, double res) { last_step_ = (unsigned)its; last_value_ = res; }

private:
  unsigned int max_it_;
  double tol_;
  unsigned int last_step_ = 0;
  double last_value_ = 0;
};

inline void check(int rc) {
  if (rc != DGX_OK) throw ExcDGX(std::string(dgx_last_error()));
}

// MPI_InitFinalize + MPI_COMM_WORLD (NS.cc:3038, 2216): one context per GPU / rank
class Context {
public:
  explicit Context(int device = 0, int rank = 0, int nranks = 1, const void *nccl_id = nullptr) { check(dgx_ctx_create(device, rank, nranks, nccl_id, &h)); }
  ~Context() { if (h) dgx_ctx_destroy(h); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  void synchronize() { check(dgx_ctx_synchronize(h)); }
  dgx_ctx *h = nullptr;
};

// parallel::distributed::Triangulation restricted to what the synthetic configurations need:
// GridGenerator::subdivided_hyper_rectangle + periodicity + refine_global (NS.cc:2261-2270)
template <int dim>
class Triangulation {
public:
  Triangulation() = default;
  ~Triangulation() { if (h) dgx_mesh_destroy(h); }
  Triangulation(const Triangulation &) = delete;
  void subdivided_hyper_rectangle(const std::array<int, dim> &repetitions, const std::array<double, dim> &p1, const std::array<double, dim> &p2,
                                  const std::array<bool, dim> &periodic, int n_global_refinements, const int *boundary_ids = nullptr) {
    int c0[3] = {1, 1, 1}, mask = 0;
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { c0[d] = repetitions[d]; lo[d] = p1[d]; hi[d] = p2[d]; if (periodic[d]) mask |= 1 << d; }
    check(dgx_mesh_create_cartesian(dim, c0, n_global_refinements, lo, hi, mask, boundary_ids, &h));
    n_refine = n_global_refinements;
    for (int d = 0; d < dim; ++d) cell_size[d] = (p2[d] - p1[d]) / (repetitions[d] * (1 << n_global_refinements));
    for (int d = 0; d < dim; ++d) { origin[d] = p1[d]; extent[d] = p2[d] - p1[d]; }
  }
  unsigned int n_global_levels() const { return (unsigned)dgx_mesh_n_levels(h); }                    // NS.cc:2351
  long long n_global_active_cells() const { return dgx_mesh_n_cells(h, n_refine); }
  dgx_mesh *h = nullptr;
  int n_refine = 0;
  std::array<double, dim> cell_size{}, origin{}, extent{};
};

// GridTools::transform(f, triangulation): every vertex of every level of the hierarchy is moved to f(vertex); MatrixFree objects
// created afterwards treat the cells as MappingQ1 images of the unit cell (NS.cc:2329 reinit(MappingQ1, ...)): the pressure and mass
// operators, lift / drag and output then use the metric at every quadrature point.  f: std::array<double, dim> -> std::array<double, dim>.
namespace GridTools {
template <int dim, typename F>
void transform(F &&f, Triangulation<dim> &tria) {
  for (int level = 0; level <= tria.n_refine; ++level) {
    int nv[3] = {1, 1, 1};
    double h[3] = {1, 1, 1};
    // cells per direction on this level = coarse repetitions << level: recover them from the finest cell size
    for (int d = 0; d < dim; ++d) {
      const double extent = tria.extent[d];
      const long long fine = std::llround(extent / tria.cell_size[d]);
      const long long here = fine >> (tria.n_refine - level);
      nv[d] = (int)here + 1;
      h[d] = extent / (double)here;
    }
    std::vector<double> coords((std::size_t)nv[0] * nv[1] * nv[2] * dim);
    std::size_t k = 0;
    for (int iz = 0; iz < nv[2]; ++iz)
      for (int iy = 0; iy < nv[1]; ++iy)
        for (int ix = 0; ix < nv[0]; ++ix) {
          const int idx[3] = {ix, iy, iz};
          std::array<double, dim> X;
          for (int d = 0; d < dim; ++d) X[d] = tria.origin[d] + idx[d] * h[d];
          const std::array<double, dim> x = f(X);
          for (int d = 0; d < dim; ++d) coords[k++] = x[d];
        }
    check(dgx_mesh_set_vertices(tria.h, level, coords.data(), (long long)coords.size()));
  }
}
}  // namespace GridTools

// MatrixFree<dim, Number>::reinit (NS.cc:2329 fine / double)
template <int dim, typename Number = double>
class MatrixFree {
public:
  MatrixFree(Context &ctx, Triangulation<dim> &tria, int level = -1) : tria_(&tria) {
    check(dgx_mf_create(ctx.h, tria.h, level, sizeof(Number) == 8 ? DGX_F64 : DGX_F32, &h));
  }
  ~MatrixFree() { if (h) dgx_mf_destroy(h); }
  MatrixFree(const MatrixFree &) = delete;
  long long n_owned_cells() const { return dgx_mf_n_owned_cells(h); }
  long long first_owned_cell() const { return dgx_mf_first_owned_cell(h); }
  const Triangulation<dim> &triangulation() const { return *tria_; }
  dgx_mf *h = nullptr;

private:
  Triangulation<dim> *tria_;
};

// LinearAlgebra::distributed::Vector<Number> (NS.cc:2102-2115)
template <typename Number>
class Vector {
public:
  Vector() = default;
  ~Vector() { if (h) dgx_vec_destroy(h); }
  Vector(const Vector &) = delete;
  // MatrixFree::initialize_dof_vector(v, dof_handler_index) (NS.cc:2330-2341): FESystem(FE_DGQ(degree), ncomp)
  template <int dim> void reinit(const MatrixFree<dim, Number> &mf, int degree, int n_components) {
    if (h) dgx_vec_destroy(h);
    check(dgx_vec_create(mf.h, degree, n_components, &h));
  }
  Vector &operator=(Number s) { check(dgx_vec_set(h, (double)s)); return *this; }
  Vector &operator=(const Vector &v) { if (this != &v) check(dgx_vec_copy(h, v.h)); return *this; }   // NS.cc:2441, 2541, 2963
  void equ(Number a, const Vector &v) { check(dgx_vec_equ(h, (double)a, v.h)); }                       // NS.cc:2408-2415, 2953
  void add(Number a, const Vector &v) { check(dgx_vec_add(h, (double)a, v.h)); }
  void sadd(Number s, Number a, const Vector &v) { check(dgx_vec_sadd(h, (double)s, (double)a, v.h)); }
  Vector &operator+=(const Vector &v) { add(Number(1), v); return *this; }                             // NS.cc:2954, 2962
  Vector &operator-=(const Vector &v) { add(Number(-1), v); return *this; }                            // NS.cc:2410, 2416
  Vector &operator*=(Number a) { check(dgx_vec_scale(h, (double)a)); return *this; }
  Number operator*(const Vector &v) const { double r; check(dgx_vec_dot(h, v.h, &r)); return (Number)r; }
  Number l2_norm() const { double r; check(dgx_vec_l2_norm(h, &r)); return (Number)r; }               // NS.cc:2450, 2486, 2574
  Number linfty_norm() const { double r; check(dgx_vec_linfty_norm(h, &r)); return (Number)r; }       // NS.cc:2585, 2597
  Number mean_value() const { double r; check(dgx_vec_mean_value(h, &r)); return (Number)r; }
  Number add_and_dot(Number a, const Vector &v, const Vector &w) { double r; check(dgx_vec_add_and_dot(h, (double)a, v.h, w.h, &r)); return (Number)r; }
  void update_ghost_values() const { check(dgx_vec_update_ghost_values(h)); }                          // NS.cc:452, 790, 917
  void zero_out_ghost_values() const { check(dgx_vec_zero_out_ghost_values(h)); }
  void compress() { check(dgx_vec_compress_add(h)); }
  std::size_t size() const { return (std::size_t)dgx_vec_global_size(h); }
  std::size_t locally_owned_size() const { return (std::size_t)dgx_vec_size(h); }
  void copy_from_host(const std::vector<Number> &v) { check(dgx_vec_copy_from_host(h, v.data(), (long long)v.size())); }
  void copy_to_host(std::vector<Number> &v) const { v.resize(locally_owned_size()); check(dgx_vec_copy_to_host(h, v.data(), (long long)v.size())); }
  dgx_vec *h = nullptr;
};

// VectorTools::interpolate(dof_handler, function, vector) for FESystem(FE_DGQ(degree), ncomp) on the Cartesian hierarchy
// (NS.cc:2393-2396): f(point, component)
template <int dim, typename Number, typename F>
void interpolate(const MatrixFree<dim, Number> &mf, int degree, int n_components, F &&f, Vector<Number> &v) {
  const int n = degree + 1;
  std::vector<double> nodes(n);
  check(dgx_fe_unit_support_points(degree, nodes.data()));
  const long long nc = mf.n_owned_cells(), first = mf.first_owned_cell();
  const auto &tria = mf.triangulation();
  std::vector<int> coords((std::size_t)nc * dim);
  check(dgx_mesh_cell_coords(tria.h, tria.n_refine, first, nc, coords.data()));
  int nd = 1;
  for (int d = 0; d < dim; ++d) nd *= n;
  std::vector<Number> host((std::size_t)nc * n_components * nd);
  for (long long c = 0; c < nc; ++c)
    for (int comp = 0; comp < n_components; ++comp)
      for (int i = 0; i < nd; ++i) {
        std::array<double, dim> p;
        int r = i;
        for (int d = 0; d < dim; ++d) { p[d] = tria.origin[d] + (coords[c * dim + d] + nodes[r % n]) * tria.cell_size[d]; r /= n; }
        host[((std::size_t)c * n_components + comp) * nd + i] = (Number)f(p, comp);
      }
  v.copy_from_host(host);
}

// NavierStokesProjectionOperator<dim, fe_degree_p, fe_degree_v, n_q_points_1d_p, n_q_points_1d_v, Vec> behind the
// MatrixFreeOperators::Base surface (NS.cc:242-2060).  The polynomial degrees are run-time values here.
template <int dim, typename Number = double>
class NavierStokesProjectionOperator {
public:
  using Vec = Vector<Number>;
  NavierStokesProjectionOperator(double Reynolds, double dt, int fe_degree_p) : Re(Reynolds), dt_(dt), degree_p(fe_degree_p) {}   // NS.cc:386-398
  ~NavierStokesProjectionOperator() { clear(); }
  NavierStokesProjectionOperator(const NavierStokesProjectionOperator &) = delete;
  void clear() {
    for (dgx_op *&o : ops_) { if (o) dgx_op_destroy(o); o = nullptr; }
    h = nullptr;
  }
  // Base::initialize(matrix_free, {k}, {k}) + set_NS_stage(stage) (NS.cc:2429-2433, 2475-2478, 2562-2571): the reference
  // re-targets ONE operator object before every solve; here each NS_stage keeps its device operator (work vectors, u_extr,
  // inverse diagonal) and initialize() selects it
  void initialize(const MatrixFree<dim, Number> &mf, unsigned int NS_stage) {
    if (NS_stage < 1 || NS_stage > 3) throw ExcDGX("NS_stage must be 1, 2 or 3");                      // AssertIndexRange, NS.cc:436-437
    mf_ = &mf; stage_ = NS_stage;
    if (!ops_[NS_stage]) check(dgx_op_create(mf.h, (int)NS_stage, NS_stage == DGX_OP_PRESSURE ? degree_p : degree_p + 1, Re, dt_, &ops_[NS_stage]));
    h = ops_[NS_stage];
    check(dgx_op_set_dt(h, dt_));
    check(dgx_op_set_TR_BDF2_stage(h, (int)trbdf2_));
  }
  void set_dt(double dt) { dt_ = dt; if (h) check(dgx_op_set_dt(h, dt)); }                             // NS.cc:415
  void set_TR_BDF2_stage(unsigned int s) { trbdf2_ = s; if (h) check(dgx_op_set_TR_BDF2_stage(h, (int)s)); }   // NS.cc:424
  void set_u_extr(const Vec &src) { check(dgx_op_set_u_extr(h, src.h)); }                              // NS.cc:448
  void initialize_dof_vector(Vec &v) const { v.reinit(*mf_, stage_ == DGX_OP_PRESSURE ? degree_p : degree_p + 1, stage_ == DGX_OP_PRESSURE ? 1 : dim); }
  void vmult(Vec &dst, const Vec &src) const { check(dgx_op_vmult(h, dst.h, src.h)); }
  void vmult_add(Vec &dst, const Vec &src) const { check(dgx_op_vmult_add(h, dst.h, src.h)); }
  void Tvmult(Vec &dst, const Vec &src) const { check(dgx_op_Tvmult(h, dst.h, src.h)); }
  void Tvmult_add(Vec &dst, const Vec &src) const { check(dgx_op_Tvmult_add(h, dst.h, src.h)); }
  std::size_t m() const { return (std::size_t)dgx_op_m(h); }                                           // NS.cc:2512
  std::size_t n() const { return m(); }
  Number el(std::size_t row, std::size_t col) const {                                                  // Base::el: diagonal entries only
    if (row != col) throw ExcDGX("el(i, j): only the diagonal is available (MatrixFreeOperators::Base::el)");
    double v; check(dgx_op_el(h, (long long)row, &v)); return (Number)v;
  }
  const MatrixFree<dim, Number> *get_matrix_free() const { return mf_; }
  void compute_diagonal() { check(dgx_op_compute_diagonal(h)); }                                       // NS.cc:2018
  void precondition_Jacobi(Vec &dst, const Vec &src, double omega) const { check(dgx_op_precondition_Jacobi(h, dst.h, src.h, omega)); }
  void vmult_rhs_velocity(Vec &dst, const std::vector<Vec *> &src) const {                             // NS.cc:788
    std::vector<dgx_vec *> hs;
    for (Vec *v : src) hs.push_back(v->h);
    check(dgx_op_vmult_rhs_velocity(h, dst.h, hs.data(), (int)hs.size()));
  }
  void vmult_rhs_pressure(Vec &dst, const std::vector<Vec *> &src) const {                             // NS.cc:915
    std::vector<dgx_vec *> hs;
    for (Vec *v : src) hs.push_back(v->h);
    check(dgx_op_vmult_rhs_pressure(h, dst.h, hs.data(), (int)hs.size()));
  }
  void vmult_grad_p_projection(Vec &dst, const Vec &src) const { check(dgx_op_vmult_grad_p_projection(h, dst.h, src.h)); }   // NS.cc:1363
  dgx_op *h = nullptr;   // the operator of the current NS_stage

private:
  double Re, dt_;
  int degree_p;
  unsigned int stage_ = 1, trbdf2_ = 1;
  dgx_op *ops_[4] = {nullptr, nullptr, nullptr, nullptr};
  const MatrixFree<dim, Number> *mf_ = nullptr;
};

// preconditioners as the solvers see them
struct PreconditionIdentity {                                                                          // NS.cc:2576
  int kind() const { return 0; }
  void *handle() const { return nullptr; }
};
template <typename Op>
struct PreconditionJacobi {                                                                            // NS.cc:2454-2461
  void initialize(const Op &op) { op_ = &op; }
  int kind() const { return 1; }
  void *handle() const { return op_ ? (void *)op_->h : nullptr; }
  const Op *op_ = nullptr;
};

// ---- multigrid wiring of projection_step, NS.cc:2490-2537, class by class ----------------------------------------------
// The reference builds the preconditioner from separate deal.II objects every solve.  Here every object is a thin, typed
// view on ONE device engine (dgx_mg: level MatrixFree / operators / vectors, transfer, smoothers, coarse solver) that
// lives with mg_matrices; each class forwards exactly the piece of configuration its deal.II namesake holds, so
// projection_step keeps its statements (examples/trbdf2_tgv.cpp).

// DoFHandler<dim> of the pressure space as far as MGTransferMatrixFree::build and PreconditionMG look at it
template <int dim>
struct DoFHandler {
  Triangulation<dim> *tria = nullptr;
  int fe_degree = 1;
};

template <int dim> class MGLevelOperators;

// mg_matrices[level]: the float level operator (NS.cc:2376-2379) = a borrowed dgx_op of the engine
template <int dim>
class LevelOperator {
public:
  LevelOperator(dgx_mg *mg, int level) : mg_(mg), level_(level) { check(dgx_mg_level_op(mg, level, &h)); }
  void compute_diagonal() { check(dgx_op_compute_diagonal(h)); }                                       // NS.cc:2514
  dgx_vec *get_matrix_diagonal_inverse() const { dgx_vec *d = nullptr; check(dgx_op_get_matrix_diagonal_inverse(h, &d)); return d; }   // NS.cc:2515
  std::size_t m() const { return (std::size_t)dgx_op_m(h); }                                           // NS.cc:2512
  std::size_t n() const { return m(); }
  float el(std::size_t row, std::size_t col) const {                                                   // Base::el: diagonal entries only
    if (row != col) throw ExcDGX("el(i, j): only the diagonal is available (MatrixFreeOperators::Base::el)");
    double v; check(dgx_op_el(h, (long long)row, &v)); return (float)v;
  }
  int level() const { return level_; }
  dgx_op *h = nullptr;

private:
  dgx_mg *mg_;
  int level_;
};

// MGLevelObject<NavierStokesProjectionOperator<..., Vector<float>>> mg_matrices (NS.cc:2351-2379): owns the engine
template <int dim>
class MGLevelOperators {
public:
  MGLevelOperators(Context &ctx, Triangulation<dim> &tria, int fe_degree_p, double Reynolds, double dt) : tria_(&tria) {
    check(dgx_mg_create(ctx.h, tria.h, fe_degree_p, DGX_F32, Reynolds, dt, &h));
  }
  ~MGLevelOperators() { if (h) dgx_mg_destroy(h); }
  MGLevelOperators(const MGLevelOperators &) = delete;
  LevelOperator<dim> operator[](unsigned int level) const { return LevelOperator<dim>(h, (int)level); }
  unsigned int min_level() const { return 0; }
  unsigned int max_level() const { return (unsigned)dgx_mg_n_levels(h) - 1; }
  void set_dt(double dt) { check(dgx_mg_set_dt(h, dt)); }                                              // NS.cc:2945-2948 on every level
  void set_TR_BDF2_stage(unsigned int s) { check(dgx_mg_set_TR_BDF2_stage(h, (int)s)); }               // NS.cc:2941-2943
  dgx_mg *h = nullptr;

private:
  Triangulation<dim> *tria_;
};

// MGLevelObject<T>: resize(min, max), operator[]
template <typename T>
class MGLevelObject {
public:
  void resize(unsigned int min_level, unsigned int max_level) { min_ = min_level; v_.assign(max_level - min_level + 1, T()); }
  T &operator[](unsigned int level) { return v_.at(level - min_); }
  const T &operator[](unsigned int level) const { return v_.at(level - min_); }
  unsigned int min_level() const { return min_; }
  unsigned int max_level() const { return min_ + (unsigned)v_.size() - 1; }

private:
  unsigned int min_ = 0;
  std::vector<T> v_;
};

namespace numbers { constexpr unsigned int invalid_unsigned_int = static_cast<unsigned int>(-1); }

// MGTransferMatrixFree<dim, Number> (NS.cc:2490-2491).  For FE_DGQ on the globally refined hierarchy the transfer is the fixed
// embedding of a parent's basis into its children (dgx_transfer_prolongate_and_add / restrict_and_add): build() has no tables
// to fill, it only checks the space.
template <int dim, typename Number>
class MGTransferMatrixFree {
public:
  void build(const DoFHandler<dim> &dof_handler) {
    if (!dof_handler.tria || dof_handler.fe_degree < 1) throw ExcDGX("MGTransferMatrixFree::build: DoFHandler without triangulation / element");
    built_ = true;
  }
  bool built() const { return built_; }

private:
  bool built_ = false;
};

// PreconditionChebyshev<Operator, Vector> (NS.cc:2493-2499): only its AdditionalData appears in the driver
template <typename Op, typename Vec>
struct PreconditionChebyshev {
  struct AdditionalData {
    double smoothing_range = 0.0;
    unsigned int degree = 1;
    unsigned int eig_cg_n_iterations = 8;
    dgx_vec *preconditioner = nullptr;      // the inverse diagonal of the level operator (NS.cc:2515)
  };
};

namespace mg {
// mg::SmootherRelaxation<SmootherType, Vec> (NS.cc:2500, 2517)
template <typename SmootherType, typename Vec>
class SmootherRelaxation {
public:
  template <int dim>
  void initialize(const MGLevelOperators<dim> &mg_matrices, const MGLevelObject<typename SmootherType::AdditionalData> &data) {
    for (unsigned int level = data.min_level(); level <= data.max_level(); ++level) {
      const auto &d = data[level];
      if (level > 0 && !d.preconditioner) throw ExcDGX("SmootherRelaxation::initialize: level without inverse diagonal (compute_diagonal first)");
      check(dgx_mg_set_smoother(mg_matrices.h, (int)level, d.degree == numbers::invalid_unsigned_int ? 1 : (int)d.degree, d.smoothing_range,
                                (int)std::min<unsigned int>(d.eig_cg_n_iterations, 1u << 30)));
    }
    engine = mg_matrices.h;
  }
  dgx_mg *engine = nullptr;
};

// mg::Matrix<Vec> (NS.cc:2531)
template <typename Vec>
class Matrix {
public:
  template <int dim> explicit Matrix(const MGLevelOperators<dim> &mg_matrices) : engine(mg_matrices.h) {}
  dgx_mg *engine;
};
}  // namespace mg

// MGCoarseGridIterativeSolver<Vec, SolverCG<Vec>, Operator, PreconditionIdentity> (NS.cc:2519-2529): CG on level 0 with
// the SolverControl of the solver handed in (the reference shares the outer solver_control, NS.cc:2520)
template <typename Vec> class SolverCG;
template <typename Vec, typename Solver, typename Op, typename Pre>
class MGCoarseGridIterativeSolver {
public:
  MGCoarseGridIterativeSolver(Solver &solver, const Op &coarse_matrix, const Pre &) : control(&solver.control()) {
    if (coarse_matrix.level() != 0) throw ExcDGX("MGCoarseGridIterativeSolver: the coarse matrix must be level 0");
  }
  const SolverControl *control;
};

// Multigrid<Vec>(matrix, coarse, transfer, pre_smooth, post_smooth) (NS.cc:2533): V-cycle over all levels
template <typename Vec>
class Multigrid {
public:
  template <class Coarse, class Transfer, class Smoother>
  Multigrid(const mg::Matrix<Vec> &matrix, const Coarse &coarse, const Transfer &transfer, const Smoother &pre, const Smoother &post)
      : engine(matrix.engine), coarse_control(coarse.control) {
    if (!transfer.built()) throw ExcDGX("Multigrid: transfer.build() has not been called");
    if (pre.engine != engine || post.engine != engine) throw ExcDGX("Multigrid: smoother initialised on other level matrices");
  }
  dgx_mg *engine;
  const SolverControl *coarse_control;
};

// PreconditionMG<dim, Vec, Transfer>(dof_handler, mg, transfer) (NS.cc:2535-2537): what SolverCG applies.  Construction
// finishes the per-solve set-up of the engine (eigenvalue estimates of the smoothers, coarse solver control).
template <int dim, typename Vec, typename Transfer>
class PreconditionMG {
public:
  PreconditionMG(const DoFHandler<dim> &, Multigrid<Vec> &mg, const Transfer &) : h(mg.engine) {
    check(dgx_mg_setup(h, mg.coarse_control->tolerance(), (int)mg.coarse_control->max_steps()));
  }
  void vmult(Vector<double> &dst, Vector<double> &src) const { check(dgx_mg_vmult(h, dst.h, src.h)); }
  int kind() const { return 2; }
  void *handle() const { return (void *)h; }
  dgx_mg *h;
};

template <typename Vec>
class SolverCG {                                                                                       // NS.cc:2487, 2542, 2546, 2575
public:
  explicit SolverCG(SolverControl &c) : control_(c) {}
  SolverControl &control() const { return control_; }
  template <class Op, class Pre>
  void solve(const Op &A, Vec &x, Vec &b, const Pre &P) {
    int its = 0;
    double res = 0;
    const int rc = dgx_solve_cg(A.h, x.h, b.h, P.kind(), P.handle(), (int)control_.max_steps(), control_.tolerance(), &its, &res);
    control_.record(its, res);
    if (rc == DGX_ERR_NO_CONVERGENCE) throw SolverControl::NoConvergence((unsigned)its, res, dgx_last_error());
    check(rc);
  }

private:
  SolverControl &control_;
};

template <typename Vec>
class SolverGMRES {                                                                                    // NS.cc:2451, 2463
public:
  struct AdditionalData { unsigned int max_basis_size = 30; };
  explicit SolverGMRES(SolverControl &c, const AdditionalData &d = AdditionalData()) : control_(c), data_(d) {}
  template <class Op, class Pre>
  void solve(const Op &A, Vec &x, Vec &b, const Pre &P) {
    int its = 0;
    double res = 0;
    const int rc = dgx_solve_gmres(A.h, x.h, b.h, P.kind(), P.handle(), (int)control_.max_steps(), control_.tolerance(), (int)data_.max_basis_size, &its, &res);
    control_.record(its, res);
    if (rc == DGX_ERR_NO_CONVERGENCE) throw SolverControl::NoConvergence((unsigned)its, res, dgx_last_error());
    check(rc);
  }

private:
  SolverControl &control_;
  AdditionalData data_;
};

}  // namespace dgxshim
#endif  // DGX_DEALII_SHIM_H
