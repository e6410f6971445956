// TEST TOOL ONLY: compiles the device-resident bisection walk (smc-nuts_b200/csrc/bisect.cuh) with g++ and drives it with
// a host callback as the objective, so the `not gpu` suite can check it against scipy.optimize.bisect bit for bit.
#include "../../smc-nuts_b200/csrc/bisect.cuh"

using namespace smcb;

typedef void (*objective_fn)(const double* x, int m, double* f);

// out4 = (result, status, iterations, nan_at); returns the number of passes used
extern "C" int hostsim_bisect(double xa, double xb, objective_fn fobj, double* out4, int* evaluations) {
    BisectState s;
    bisect_init(s, xa, xb, 0.0, 2e-12, 8.881784197001252e-16, 100);
    int passes = 0, evals = 0;
    for (; passes < kBisectPasses && s.status == kBisectRunning; ++passes) {
        double f[kBisectMaxCand];
        fobj(s.cand, s.n_cand, f);
        evals += s.n_cand;
        bisect_advance(s, f);
    }
    out4[0] = s.result; out4[1] = (double)s.status; out4[2] = (double)s.iterations; out4[3] = s.nan_at;
    if (evaluations) *evaluations = evals;
    return passes;
}
