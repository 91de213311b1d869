// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// Dense convex QP   min 1/2 z'Hz + c'z   s.t.  A z <= ub      (the form HoQp hands to qpOASES:
// qm_wbc/src/HoQp.cpp:135-150 — QProblem(nV,nC), init(H,g,A,nullptr,nullptr,nullptr,ubA), no bounds,
// no lower constraint limits).  qpOASES@268b2f2 is not available offline (qpoases_catkin/CMakeLists.txt:27-29),
// so the oracle uses a textbook primal active-set method (Nocedal & Wright Alg. 16.3) with a null-space
// step and a rank-revealing reduced-Hessian solve.  Where the QP optimum is unique both return it;
// where H is only semidefinite (HoQp adds 1e-12 I, HoQp.cpp:66) the choice inside the flat directions
// differs from qpOASES but does not change the final HoQp solution.
#pragma once
#include "linalg.h"

namespace orc {

struct QpResult { Vec z; int iterations = 0; int status = 0; /*0 ok, 1 iteration cap*/ std::vector<int> active; };

// z0 must be feasible.
QpResult solve_qp_active_set(const Mat& H, const Vec& c, const Mat& A, const Vec& ub, const Vec& z0, int max_iter = 1000);

}  // namespace orc
