// TR-BDF2 projection solver for the periodic 3-D Taylor-Green vortex, written against include/dgx_dealii_shim.h the way
// NavierStokesProjection<dim> is written against deal.II: the members below follow interpolate_velocity (NS.cc:2403-2418),
// diffusion_step (:2424-2464), projection_step (:2470-2548), project_grad (:2554-2577) and the body of run()
// (:2928-3010) statement by statement; only set-up (mesh, MatrixFree, multigrid object) differs, as INTEGRATION.md
// describes.  Prints one JSON line with the Krylov iteration counts of every solve and the device time per step.
//
//   usage: trbdf2_tgv [n_refines=4] [n_steps=2] [degree_p=2] [Reynolds=1600]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "dgx_dealii_shim.h"

using namespace dgxshim;

template <int dim>
class NavierStokesProjection {
public:
  using Vec = Vector<double>;
  NavierStokesProjection(int n_refines, int degree_p_, double Reynolds, double cfl)
      : degree_p(degree_p_), Re(Reynolds), gamma(2.0 - std::sqrt(2.0)),                               // NS.cc:2210
        dt(cfl * (2.0 * M_PI / (1 << n_refines)) / (degree_p_ + 2)), max_its(10000), eps(1e-8),        // parameter-file.prm:27-28
        navier_stokes_matrix(Reynolds, cfl * (2.0 * M_PI / (1 << n_refines)) / (degree_p_ + 2), degree_p_) {
    const double L = 2.0 * M_PI;
    triangulation.subdivided_hyper_rectangle({1, 1, 1}, {0.0, 0.0, 0.0}, {L, L, L}, {true, true, true}, n_refines);
    matrix_free_storage = std::make_unique<MatrixFree<dim>>(context, triangulation);                   // NS.cc:2329
    mg_matrices = std::make_unique<MGLevelOperators<dim>>(context, triangulation, degree_p, Re, dt);      // NS.cc:2351-2379
    dof_handler_pressure.tria = &triangulation;
    dof_handler_pressure.fe_degree = degree_p;                                                              // NS.cc:2282-2283
    for (Vec *v : {&u_star, &rhs_u, &u_n, &u_extr, &u_n_minus_1, &u_n_gamma, &u_tmp, &grad_pres_int})
      v->reinit(*matrix_free_storage, degree_p + 1, dim);                                               // NS.cc:2330-2337
    for (Vec *v : {&pres_int, &pres_n, &rhs_p}) v->reinit(*matrix_free_storage, degree_p, 1);            // NS.cc:2339-2341
    initialize();
  }

  // NS.cc:2389-2397
  void initialize() {
    auto vel = [](const std::array<double, dim> &p, int c) {
      return c == 0 ? std::sin(p[0]) * std::cos(p[1]) * std::cos(p[2]) : (c == 1 ? -std::cos(p[0]) * std::sin(p[1]) * std::cos(p[2]) : 0.0);
    };
    auto pres = [](const std::array<double, dim> &p, int) { return (std::cos(2 * p[0]) + std::cos(2 * p[1])) * (std::cos(2 * p[2]) + 2.0) / 16.0; };
    interpolate(*matrix_free_storage, degree_p, 1, pres, pres_n);
    interpolate(*matrix_free_storage, degree_p + 1, dim, vel, u_n_minus_1);
    interpolate(*matrix_free_storage, degree_p + 1, dim, vel, u_n);
  }

  void interpolate_velocity() {
    if (TR_BDF2_stage == 1) {
      u_extr.equ(1.0 + gamma / (2.0 * (1.0 - gamma)), u_n);
      u_tmp.equ(gamma / (2.0 * (1.0 - gamma)), u_n_minus_1);
      u_extr -= u_tmp;
    } else {
      u_extr.equ(1.0 + (1.0 - gamma) / gamma, u_n_gamma);
      u_tmp.equ((1.0 - gamma) / gamma, u_n);
      u_extr -= u_tmp;
    }
  }

  void diffusion_step() {
    navier_stokes_matrix.initialize(*matrix_free_storage, 1);
    if (TR_BDF2_stage == 1) {
      navier_stokes_matrix.vmult_rhs_velocity(rhs_u, {&u_n, &u_extr, &pres_n});
      navier_stokes_matrix.set_u_extr(u_extr);
      u_star = u_extr;
    } else {
      navier_stokes_matrix.vmult_rhs_velocity(rhs_u, {&u_n, &u_n_gamma, &pres_int, &u_extr});
      navier_stokes_matrix.set_u_extr(u_extr);
      u_star = u_extr;
    }
    SolverControl solver_control(max_its, eps * rhs_u.l2_norm());
    SolverGMRES<Vec> gmres(solver_control);
    PreconditionJacobi<NavierStokesProjectionOperator<dim>> preconditioner;
    navier_stokes_matrix.compute_diagonal();
    preconditioner.initialize(navier_stokes_matrix);
    gmres.solve(navier_stokes_matrix, u_star, rhs_u, preconditioner);
    iterations.emplace_back("gmres_velocity_stage" + std::to_string(TR_BDF2_stage), solver_control.last_step());
  }

  void projection_step() {
    navier_stokes_matrix.initialize(*matrix_free_storage, 2);
    if (TR_BDF2_stage == 1) navier_stokes_matrix.vmult_rhs_pressure(rhs_p, {&u_star, &pres_n});
    else navier_stokes_matrix.vmult_rhs_pressure(rhs_p, {&u_star, &pres_int});
    /*--- Build the linear solver (NS.cc:2485-2487) ---*/
    SolverControl solver_control(max_its, eps * rhs_p.l2_norm());
    SolverCG<Vec> cg(solver_control);

    /*--- Build the preconditioner (NS.cc:2489-2537, statement by statement) ---*/
    MGTransferMatrixFree<dim, float> mg_transfer;
    mg_transfer.build(dof_handler_pressure);

    using SmootherType = PreconditionChebyshev<LevelOperator<dim>, Vector<float>>;
    mg::SmootherRelaxation<SmootherType, Vector<float>> mg_smoother;
    MGLevelObject<typename SmootherType::AdditionalData> smoother_data;
    smoother_data.resize(0, triangulation.n_global_levels() - 1);
    for (unsigned int level = 0; level < triangulation.n_global_levels(); ++level) {
      auto level_matrix = (*mg_matrices)[level];
      if (level > 0) {
        smoother_data[level].smoothing_range = 15.0;
        smoother_data[level].degree = 3;
        smoother_data[level].eig_cg_n_iterations = 10;
      } else {
        smoother_data[0].smoothing_range = 2e-2;
        smoother_data[0].degree = numbers::invalid_unsigned_int;
        smoother_data[0].eig_cg_n_iterations = (unsigned int)level_matrix.m();
      }
      level_matrix.compute_diagonal();
      smoother_data[level].preconditioner = level_matrix.get_matrix_diagonal_inverse();
    }
    mg_smoother.initialize(*mg_matrices, smoother_data);

    PreconditionIdentity identity;
    SolverCG<Vector<float>> cg_mg(solver_control);
    MGCoarseGridIterativeSolver<Vector<float>, SolverCG<Vector<float>>, LevelOperator<dim>, PreconditionIdentity> mg_coarse(cg_mg, (*mg_matrices)[0], identity);

    mg::Matrix<Vector<float>> mg_matrix(*mg_matrices);

    Multigrid<Vector<float>> mg(mg_matrix, mg_coarse, mg_transfer, mg_smoother, mg_smoother);

    PreconditionMG<dim, Vector<float>, MGTransferMatrixFree<dim, float>> preconditioner(dof_handler_pressure, mg, mg_transfer);

    /*--- Solve the linear system (NS.cc:2539-2547) ---*/
    if (TR_BDF2_stage == 1) {
      pres_int = pres_n;
      cg.solve(navier_stokes_matrix, pres_int, rhs_p, preconditioner);
    } else {
      pres_n = pres_int;
      cg.solve(navier_stokes_matrix, pres_n, rhs_p, preconditioner);
    }
    iterations.emplace_back("cg_pressure_stage" + std::to_string(TR_BDF2_stage), solver_control.last_step());
  }

  void project_grad(const unsigned int flag) {
    if (flag != 1 && flag != 2) throw ExcDGX("project_grad: flag must be 1 or 2");                     // NS.cc:2557
    navier_stokes_matrix.initialize(*matrix_free_storage, 3);
    if (flag == 1) navier_stokes_matrix.vmult_grad_p_projection(rhs_u, pres_n);
    else navier_stokes_matrix.vmult_grad_p_projection(rhs_u, pres_int);
    SolverControl solver_control(max_its, 1e-12 * rhs_u.l2_norm());
    SolverCG<Vec> cg(solver_control);
    cg.solve(navier_stokes_matrix, u_tmp, rhs_u, PreconditionIdentity());
    iterations.emplace_back("cg_mass_flag" + std::to_string(flag), solver_control.last_step());
  }

  void set_stage(unsigned int s) {                                                                     // NS.cc:2940-2943, 2967-2969
    TR_BDF2_stage = s;
    navier_stokes_matrix.set_TR_BDF2_stage(s);
    mg_matrices->set_TR_BDF2_stage(s);
  }

  // one pass of the while loop of run(), NS.cc:2934-2986
  double advance() {
    set_stage(1);
    interpolate_velocity();
    diffusion_step();
    project_grad(1);
    u_tmp.equ(gamma * dt, u_tmp);
    u_star += u_tmp;
    projection_step();
    u_n_gamma.equ(1.0, u_star);
    project_grad(2);
    grad_pres_int.equ(1.0, u_tmp);
    u_tmp.equ(-gamma * dt, u_tmp);
    u_n_gamma += u_tmp;
    u_n_minus_1 = u_n;
    set_stage(2);
    interpolate_velocity();
    diffusion_step();
    u_tmp.equ((1.0 - gamma) * dt, grad_pres_int);
    u_star += u_tmp;
    projection_step();
    u_n.equ(1.0, u_star);
    project_grad(1);
    u_tmp.equ((gamma - 1.0) * dt, u_tmp);
    u_n += u_tmp;
    return u_n.linfty_norm();                                                                          // get_maximal_velocity, NS.cc:2583-2587
  }

  // compute_lift_and_drag, NS.cc:2642-2708: face integrals on the device, summed over the ranks (no face of the periodic box
  // carries id 4: the call is here because run() makes it after every step, NS.cc:2991)
  std::pair<double, double> compute_lift_and_drag() {
    double lift = 0, drag = 0;
    check(dgx_compute_lift_and_drag(u_n.h, pres_n.h, Re, 4, &lift, &drag));
    return {lift, drag};
  }

  // output_results, NS.cc:2608-2633: v, p, v_old and the vorticity at the patch vertices, evaluated on the device
  void output_results(const std::string &saving_dir, unsigned int step) {
    char name[64];
    std::snprintf(name, sizeof(name), "/solution-%05u.vtu", step);
    check(dgx_output_results(u_n.h, pres_n.h, u_n_minus_1.h, (saving_dir + name).c_str()));
  }

  Context context;
  int degree_p;
  double Re, gamma, dt;
  unsigned int max_its;
  double eps;
  unsigned int TR_BDF2_stage = 1;
  Triangulation<dim> triangulation;
  std::unique_ptr<MatrixFree<dim>> matrix_free_storage;
  std::unique_ptr<MGLevelOperators<dim>> mg_matrices;
  DoFHandler<dim> dof_handler_pressure;
  NavierStokesProjectionOperator<dim> navier_stokes_matrix;
  Vec pres_int, pres_n, rhs_p, u_star, rhs_u, u_n, u_extr, u_n_minus_1, u_n_gamma, u_tmp, grad_pres_int;
  std::vector<std::pair<std::string, unsigned int>> iterations;
};

int main(int argc, char **argv) {
  try {
    const int n_refines = argc > 1 ? std::atoi(argv[1]) : 4, n_steps = argc > 2 ? std::atoi(argv[2]) : 2, degree_p = argc > 3 ? std::atoi(argv[3]) : 2;
    const double Re = argc > 4 ? std::atof(argv[4]) : 1600.0;
    const std::string saving_dir = argc > 5 ? argv[5] : "";   // output_results after the last step when given
    NavierStokesProjection<3> test(n_refines, degree_p, Re, 0.2);
    std::printf("{\"cells\": \"%d^3\", \"degree_p\": %d, \"dt\": %.17g, \"launches_setup\": %lld, \"steps\": [", 1 << n_refines, degree_p, test.dt,
                dgx_kernel_launch_count());
    for (int s = 0; s < n_steps; ++s) {
      test.iterations.clear();
      test.context.synchronize();
      const long long l0 = dgx_kernel_launch_count();
      check(dgx_ctx_timer_start(test.context.h));
      const double vmax = test.advance();
      double ms = 0;
      check(dgx_ctx_timer_stop(test.context.h, &ms));
      const long long launches = dgx_kernel_launch_count() - l0;
      const auto forces = test.compute_lift_and_drag();                                                 // NS.cc:2991
      if (!saving_dir.empty() && s == n_steps - 1) test.output_results(saving_dir, (unsigned int)(s + 2));
      std::printf("%s{\"step\": %d, \"ms\": %.3f, \"max_velocity\": %.17g, \"lift\": %.17g, \"drag\": %.17g, \"launches\": %lld, \"iterations\": [", s ? ", " : "",
                  s, ms, vmax, forces.first, forces.second, launches);
      for (std::size_t i = 0; i < test.iterations.size(); ++i)
        std::printf("%s[\"%s\", %u]", i ? ", " : "", test.iterations[i].first.c_str(), test.iterations[i].second);
      std::printf("]}");
    }
    std::printf("]}\n");
  } catch (std::exception &exc) {                                                                       // NS.cc:3055-3077
    std::cerr << "Exception on processing: " << exc.what() << std::endl << "Aborting!" << std::endl;
    return 1;
  }
  return 0;
}
