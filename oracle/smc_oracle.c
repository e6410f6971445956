/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked, imported or executed by the product path.
 *
 * Plain-C restatement of the reference's per-particle hot path, kept structurally identical to
 * the Python it follows (recursive build_tree, two half-kicks, slice variable, uniform merges):
 *
 *   orc_nuts_batch        /root/reference/smcnuts/proposal/nuts.py:34-175  (rvs, generate_nuts_samples,
 *                         build_tree, stop_criterion, NUTSLeapfrog)
 *                         + /root/reference/smcnuts/proposal/nuts_acc_rej.py:42-49 and utils.py:22-34
 *   orc_logp_split        /root/reference/stan_models/arma/arma.stan:16-30,
 *                         /root/reference/stan_models/PRMwCD/PRMwCD.stan:17-39 (via BridgeStan in the
 *                         reference, bridgestan.py:46,78 -- PARITY UNPINNED for the model arithmetic,
 *                         see oracle/models.py) and the synthetic Gaussian of SURVEY.md section 8d.
 *   orc_uniform/normals   the Philox4x32-10 stream layout documented in oracle/philox.py.
 *
 * Randomness: the reference pulls from one shared sequential numpy stream; here every particle owns
 * the Philox stream (seed, iteration, STREAM_NUTS, particle) and consumes it in exactly the
 * reference's order (exponential; per doubling: direction uniform, merge uniforms in post-order,
 * top-level uniform only when the subtree did not stop).  oracle/philox.py::ReplayRNG feeds the same
 * numbers to the unmodified reference; tests/golden pins this file against it.
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "devmath.h"

#define MAXD 128
#define LOG_2PI 1.8378770664093454835606594728112
#define LOG_PI 1.1447298858494001741434273513531

/* ------------------------------------------------------------------ elementary functions
 * Default: glibc (the mode pinned against the reference goldens).  orc_set_devmath(1): the table-driven exp / log of
 * the CUDA kernels restated in oracle/devmath.h, used to predict the parity device build bit for bit. */
static int g_devmath = 0;
void orc_set_devmath(int on) { g_devmath = on; }
int orc_get_devmath(void) { return g_devmath; }
static double m_exp(double x) { return g_devmath ? dm_exp(x) : exp(x); }
static double m_log1p(double q) { return g_devmath ? dm_log(1.0 + q) : log1p(q); }
/* log(1 - u) of the slice variable (nuts.py:69: logu = H0 - Exp(1), Exp(1) = -log1p(-u)) */
static double m_log_1mu(double u) { return g_devmath ? dm_log(1.0 - u) : log1p(-u); }
void orc_devmath_exp(const double* x, long n, double* out) { for (long i = 0; i < n; ++i) out[i] = dm_exp(x[i]); }
void orc_devmath_log(const double* x, long n, double* out) { for (long i = 0; i < n; ++i) out[i] = dm_log(x[i]); }

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
}

static double stream_uniform(uint64_t seed, uint32_t iter, uint32_t stream, uint64_t particle, uint64_t draw) {
    uint32_t c[4] = {(uint32_t)(draw >> 1), (iter << 8) | stream, (uint32_t)particle, (uint32_t)(particle >> 32)};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    int h = (int)(draw & 1);
    uint64_t u64 = ((uint64_t)c[2 * h + 1] << 32) | c[2 * h];
    return (double)(u64 >> 11) * 0x1.0p-53;
}

void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox(c, key[0], key[1]);
    memcpy(out, c, sizeof c);
}

void orc_uniform(uint64_t seed, uint32_t iter, uint32_t stream, uint64_t particle0, long n, uint64_t draw, double* u) {
    for (long i = 0; i < n; ++i) u[i] = stream_uniform(seed, iter, stream, particle0 + (uint64_t)i, draw);
}

void orc_normals(uint64_t seed, uint32_t iter, uint32_t stream, uint64_t particle0, long n, int dim, double* z) {
    for (long i = 0; i < n; ++i)
        for (int j = 0; j < dim; j += 2) {
            double u1 = stream_uniform(seed, iter, stream, particle0 + (uint64_t)i, (uint64_t)j);
            double u2 = stream_uniform(seed, iter, stream, particle0 + (uint64_t)i, (uint64_t)j + 1);
            double rad = sqrt(-2.0 * log1p(-u1)), ang = 2.0 * M_PI * u2;
            z[i * dim + j] = rad * cos(ang);
            if (j + 1 < dim) z[i * dim + j + 1] = rad * sin(ang);
        }
}

/* ------------------------------------------------------------------ models */
typedef struct {
    int kind, dim;
    long n;
    double* data;
} Model;

void* orc_model_create(int kind, const double* data, long n) {
    Model* m = (Model*)calloc(1, sizeof(Model));
    m->kind = kind;
    m->n = n;
    m->data = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(m->data, data, sizeof(double) * (size_t)n);
    if (kind == 0) m->dim = 4;
    else if (kind == 1) m->dim = 13;
    else { m->dim = (int)llround(sqrt((double)n)); }
    return m;
}
void orc_model_destroy(void* h) { Model* m = (Model*)h; free(m->data); free(m); }
int orc_model_dim(void* h) { return ((Model*)h)->dim; }

/* arma.stan:16-30.  x = (mu, beta, theta, s), sigma = exp(s).  data = y[T]. */
static void arma_split(const Model* m, const double* x, double* A, double* B, double* gA, double* gB) {
    const double* y = m->data;
    const int T = (int)m->n;
    double mu = x[0], beta = x[1], theta = x[2], s = x[3];
    double sigma = m_exp(s), sig2 = sigma * sigma, q = sig2 / 6.25;
    *A = (-0.5 * LOG_2PI - log(10.0) - mu * mu / 200.0) + (-0.5 * LOG_2PI - log(2.0) - beta * beta / 8.0) +
         (-0.5 * LOG_2PI - log(2.0) - theta * theta / 8.0) + (-LOG_PI - log(2.5) - m_log1p(q)) + s;
    gA[0] = -mu / 100.0; gA[1] = -beta / 4.0; gA[2] = -theta / 4.0; gA[3] = 1.0 - 2.0 * q / (1.0 + q);
    double e = y[0] - (mu + beta * mu);
    double dm = -(1.0 + beta), db = -mu, dt = 0.0;
    double S = e * e, Sm = e * dm, Sb = e * db, St = e * dt;
    /* err[t] = y[t] - (mu + beta*y[t-1] + theta*err[t-1]) (arma.stan:26-27), associated as
     * (y[t] - (mu + beta*y[t-1])) - theta*err[t-1] -- the same order as the device function in
     * smc-nuts_b200/csrc/models.cuh, so the g++ build of the device lane code is bit-identical to this oracle.
     * oracle/models.py keeps the literal Stan order; the two agree to ~1e-15 (tests/test_oracle_golden.py). */
    const double ntheta = -theta;
    for (int t = 1; t < T; ++t) {
        double c = y[t] - (mu + beta * y[t - 1]);
        double en = ntheta * e + c;
        double dmn = ntheta * dm - 1.0, dbn = ntheta * db - y[t - 1], dtn = ntheta * dt - e;
        e = en; dm = dmn; db = dbn; dt = dtn;
        S += e * e; Sm += e * dm; Sb += e * db; St += e * dt;
    }
    double inv = 1.0 / sig2;
    *B = -0.5 * T * LOG_2PI - T * s - 0.5 * S * inv;
    gB[0] = -Sm * inv; gB[1] = -Sb * inv; gB[2] = -St * inv; gB[3] = -T + S * inv;
    if (!isfinite(sigma) || sigma <= 0.0) *A = -INFINITY;
}

/* PRMwCD.stan:17-39.  x = (Beta_1..12, g).  data = [q, y(100), lgamma(y+1)(100), X(100x11)].
 * sum_i [y_i eta_i - exp(eta_i) - lgamma(y_i+1)] is evaluated with the y-weighted sums hoisted
 * (sum_i y_i eta_i = Beta_1 sum y + sum_j Beta_{j+1} (X'y)_j) and eta_i as two interleaved partial sums -- the same
 * association as the device function in smc-nuts_b200/csrc/models.cuh, so the g++ build of the device lane code is
 * bit-identical to this oracle.  oracle/models.py keeps the literal per-observation order of the Stan program; the
 * two agree to ~1e-14 relative (tests/test_oracle_golden.py). */
static void prm_split(const Model* m, const double* x, double* A, double* B, double* gA, double* gB) {
    const int NO = 100, C = 11, M = 12;
    const double q = m->data[0];
    const double *y = m->data + 1, *lg = y + NO, *X = lg + NO;
    double hdr[13] = {0};
    for (int i = 0; i < NO; ++i) {
        hdr[0] += y[i];
        for (int j = 0; j < C; ++j) hdr[1 + j] += y[i] * X[i * C + j];
        hdr[12] += lg[i];
    }
    double g = x[M];
    double slam = 0.0, min_eta = 1e308, gl[12] = {0};
    for (int i = 0; i < NO; ++i) {
        const double* row = X + i * C;
        double e0 = x[0], e1 = 0.0;
        for (int j = 0; j < C; j += 2) {
            e0 += x[j + 1] * row[j];
            if (j + 1 < C) e1 += x[j + 2] * row[j + 1];
        }
        double eta = e0 + e1;
        min_eta = eta < min_eta ? eta : min_eta;
        double lam = m_exp(eta);
        slam += lam;
        gl[0] += lam;
        for (int j = 0; j < C; ++j) gl[j + 1] += lam * row[j];
    }
    double ydot = x[0] * hdr[0];
    for (int j = 0; j < C; ++j) ydot += x[j + 1] * hdr[1 + j];
    double b = ydot - slam - hdr[12];
    if (m_exp(min_eta) == 0.0) { /* Stan: lambda == 0 with y != 0 -> -inf */
        for (int i = 0; i < NO; ++i) {
            double eta = x[0];
            for (int j = 0; j < C; ++j) eta += x[j + 1] * X[i * C + j];
            if (m_exp(eta) == 0.0 && y[i] > 0.0) b = -INFINITY;
        }
    }
    *B = b;
    for (int j = 0; j < M; ++j) gB[j] = hdr[j] - gl[j];
    gB[M] = 0.0;
    double ig = m_exp(-g), sum = 0.0;
    gA[0] = 0.0;
    for (int i = 1; i < M; ++i) {
        double aq = (q == 0.5) ? sqrt(fabs(x[i]) * ig) : pow(fabs(x[i]) * ig, q);
        sum += aq;
        gA[i] = -q * aq / x[i];
    }
    *A = (2.0 * log(1.3) - lgamma(2.0) - 3.0 * g - 1.3 * ig) + g + (-(M - 1) * g - sum);
    gA[M] = -3.0 + 1.3 * ig + 1.0 - (M - 1) + q * sum;
    double Gam = m_exp(g);
    if (!isfinite(Gam) || Gam <= 0.0) *A = -INFINITY;
}

/* synthetic Gaussian: data = P (D x D, row-major, symmetric).  A = 0, B = -x'Px/2. */
static void gauss_split(const Model* m, const double* x, double* A, double* B, double* gA, double* gB) {
    const int D = m->dim;
    const double* P = m->data;
    double qf = 0.0;
    for (int i = 0; i < D; ++i) {
        double acc = 0.0;
        for (int k = 0; k < D; ++k) acc += P[i * D + k] * x[k];
        gB[i] = -acc; gA[i] = 0.0;
        qf += x[i] * acc;
    }
    *A = 0.0; *B = -0.5 * qf;
}

static void model_split(const Model* m, const double* x, double* A, double* B, double* gA, double* gB) {
    if (m->kind == 0) arma_split(m, x, A, B, gA, gB);
    else if (m->kind == 1) prm_split(m, x, A, B, gA, gB);
    else gauss_split(m, x, A, B, gA, gB);
}

/* logp(x, phi) = A + phi*B with the failure mapping of bridgestan.py:47-49,79-80 */
static double model_eval(const Model* m, const double* x, double phi, double* grad) {
    double A, B, gA[MAXD], gB[MAXD];
    model_split(m, x, &A, &B, gA, gB);
    double lp = A + phi * B;
    int bad = !isfinite(lp);
    if (bad) lp = -INFINITY;
    if (grad)
        for (int d = 0; d < m->dim; ++d) grad[d] = bad ? -INFINITY : gA[d] + phi * gB[d];
    return lp;
}

void orc_logp_split(void* h, const double* x, long N, double* A, double* B, double* gA, double* gB) {
    const Model* m = (const Model*)h;
    const int D = m->dim;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        double ga[MAXD], gb[MAXD];
        model_split(m, x + i * D, &A[i], &B[i], ga, gb);
        if (gA) memcpy(gA + i * D, ga, sizeof(double) * D);
        if (gB) memcpy(gB + i * D, gb, sizeof(double) * D);
    }
}

void orc_logp_grad(void* h, const double* x, long N, double phi, double* lp, double* grad) {
    const Model* m = (const Model*)h;
    const int D = m->dim;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) lp[i] = model_eval(m, x + i * D, phi, grad ? grad + i * D : NULL);
}

/* ------------------------------------------------------------------ NUTS (nuts.py) */
typedef struct {
    const Model* m;
    int D, max_depth;
    double eps, phi, logu;
    uint64_t seed, particle, draw;
    uint32_t iter;
    long n_leapfrog;
    const double* scale;   /* diagonal metric (NULL: identity, the reference): NUTS runs on z = x / scale */
    double H0, accept_sum; /* NUTS acceptance statistic (step-size adaptation; not in the reference): sum over the
                              leaves of min(1, exp(joint - joint_0)) */
} Ctx;

/* the model seen from the sampler: pi_z(z) = pi_x(scale * z), grad_z = scale * grad_x (identity when scale == NULL) */
static double ctx_eval(const Ctx* c, const double* z, double* grad) {
    if (!c->scale) return model_eval(c->m, z, c->phi, grad);
    double x[MAXD];
    for (int d = 0; d < c->D; ++d) x[d] = z[d] * c->scale[d];
    double lp = model_eval(c->m, x, c->phi, grad);
    for (int d = 0; d < c->D; ++d) grad[d] = grad[d] * c->scale[d];
    return lp;
}

static double next_uniform(Ctx* c) { return stream_uniform(c->seed, c->iter, 0u, c->particle, c->draw++); }

static double dot(const double* a, const double* b, int D) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) s += a[d] * b[d];
    return s;
}

/* nuts.py:152-160 */
static int stop_criterion(const double* xm, const double* xp, const double* rm, const double* rp, int D) {
    double a = 0.0, b = 0.0;
    for (int d = 0; d < D; ++d) { double dx = xp[d] - xm[d]; a += dx * rm[d]; b += dx * rp[d]; }
    return (a < 0) || (b < 0);
}

typedef struct {
    double xm[MAXD], rm[MAXD], gm[MAXD], xp[MAXD], rp[MAXD], gp[MAXD], xc[MAXD], rc[MAXD];
    double lpc;
    long n;
    int s;
} Tree;

/* nuts.py:114-150 */
static void build_tree(Ctx* c, const double* x, const double* r, const double* g, int dir, int depth, Tree* t) {
    const int D = c->D;
    const size_t sz = sizeof(double) * (size_t)D;
    if (depth == 0) {
        /* NUTSLeapfrog, nuts.py:162-175 */
        double half = dir * c->eps / 2, full = dir * c->eps;
        double xn[MAXD], rn[MAXD], gn[MAXD];
        for (int d = 0; d < D; ++d) rn[d] = r[d] + half * g[d];
        for (int d = 0; d < D; ++d) xn[d] = x[d] + full * rn[d];
        double lp = ctx_eval(c, xn, gn);
        for (int d = 0; d < D; ++d) rn[d] = rn[d] + half * gn[d];
        c->n_leapfrog++;
        double joint = lp - 0.5 * dot(rn, rn, D);
        {
            double dj = joint - c->H0;
            c->accept_sum += (dj >= 0.0) ? 1.0 : ((dj == dj) ? m_exp(dj) : 0.0);
        }
        t->n = (c->logu < joint);
        t->s = ((c->logu - 100.) >= joint);
        memcpy(t->xm, xn, sz); memcpy(t->xp, xn, sz); memcpy(t->xc, xn, sz);
        memcpy(t->rm, rn, sz); memcpy(t->rp, rn, sz); memcpy(t->rc, rn, sz);
        memcpy(t->gm, gn, sz); memcpy(t->gp, gn, sz);
        t->lpc = lp;
        return;
    }
    build_tree(c, x, r, g, dir, depth - 1, t);
    if (t->s == 0) {
        Tree* u = (Tree*)malloc(sizeof(Tree));
        if (dir == -1) {
            build_tree(c, t->xm, t->rm, t->gm, dir, depth - 1, u);
            memcpy(t->xm, u->xm, sz); memcpy(t->rm, u->rm, sz); memcpy(t->gm, u->gm, sz);
        } else {
            build_tree(c, t->xp, t->rp, t->gp, dir, depth - 1, u);
            memcpy(t->xp, u->xp, sz); memcpy(t->rp, u->rp, sz); memcpy(t->gp, u->gp, sz);
        }
        double tot = (double)(t->n + u->n);
        if (next_uniform(c) < ((double)u->n / (tot > 1. ? tot : 1.))) {
            memcpy(t->xc, u->xc, sz); memcpy(t->rc, u->rc, sz);
            t->lpc = u->lpc;
        }
        t->n = t->n + u->n;
        t->s = (t->s || u->s || stop_criterion(t->xm, t->xp, t->rm, t->rp, D));
        free(u);
    }
}

/* generate_nuts_samples, nuts.py:58-112.  Returns selected (x, r) and its logp; lp0 = logp(x0). */
static void nuts_one(Ctx* c, const double* x0, const double* r0, double* xo, double* ro, double* lp0_out,
                     double* lpsel_out, int* depth_out) {
    const int D = c->D;
    const size_t sz = sizeof(double) * (size_t)D;
    double g0[MAXD];
    double logp = ctx_eval(c, x0, g0);
    double H0 = logp - 0.5 * dot(r0, r0, D);
    c->H0 = H0; c->accept_sum = 0.0;
    if (g_devmath) {
        c->logu = H0 + m_log_1mu(next_uniform(c));
    } else {
        double expo = -log1p(-next_uniform(c));
        c->logu = H0 - expo;
    }

    double xm[MAXD], rm[MAXD], gm[MAXD], xp[MAXD], rp[MAXD], gp[MAXD];
    memcpy(xm, x0, sz); memcpy(xp, x0, sz); memcpy(rm, r0, sz); memcpy(rp, r0, sz);
    memcpy(gm, g0, sz); memcpy(gp, g0, sz);
    memcpy(xo, x0, sz); memcpy(ro, r0, sz);
    double lpsel = logp;
    int depth = 0, stop = 0;
    long n = 1;
    Tree* t = (Tree*)malloc(sizeof(Tree));
    while (stop == 0) {
        int dir = (next_uniform(c) < 0.5) ? 1 : -1;
        if (dir == -1) {
            build_tree(c, xm, rm, gm, dir, depth, t);
            memcpy(xm, t->xm, sz); memcpy(rm, t->rm, sz); memcpy(gm, t->gm, sz);
        } else {
            build_tree(c, xp, rp, gp, dir, depth, t);
            memcpy(xp, t->xp, sz); memcpy(rp, t->rp, sz); memcpy(gp, t->gp, sz);
        }
        if (t->s == 0) {
            double ratio = (double)t->n / (double)n;
            if (next_uniform(c) < (ratio < 1. ? ratio : 1.)) {
                memcpy(xo, t->xc, sz); memcpy(ro, t->rc, sz);
                lpsel = t->lpc;
            }
        }
        n += t->n;
        stop = t->s || stop_criterion(xm, xp, rm, rp, D);
        depth += 1;
        if (depth > c->max_depth) break;
    }
    free(t);
    *lp0_out = logp; *lpsel_out = lpsel; *depth_out = depth;
}

void orc_nuts_batch_stat(void* h, const double* x, const double* r, long N, double eps, double phi, int max_depth,
                         uint64_t seed, uint32_t iter, uint64_t particle0, int accrej, double* x_new, double* r_new,
                         double* lp_old, double* lp_new, int* n_leapfrog, int* accepted, int* depth_out,
                         double* accept_stat, int nthreads);
void orc_nuts_batch_metric(void* h, const double* x, const double* r, long N, double eps, double phi, int max_depth,
                           uint64_t seed, uint32_t iter, uint64_t particle0, int accrej, double* x_new, double* r_new,
                           double* lp_old, double* lp_new, int* n_leapfrog, int* accepted, int* depth_out,
                           double* accept_stat, const double* scale, int nthreads);

/* NUTSProposal.rvs (nuts.py:34-56) and, when accrej != 0, NUTSProposalWithAccRej.rvs
 * (nuts_acc_rej.py:27-52) + hmc_accept_reject (utils.py:22-34).
 * lp_old = logp(x_cond, phi); lp_new = logp(returned x_prime, phi). */
void orc_nuts_batch(void* h, const double* x, const double* r, long N, double eps, double phi, int max_depth,
                    uint64_t seed, uint32_t iter, uint64_t particle0, int accrej, double* x_new, double* r_new,
                    double* lp_old, double* lp_new, int* n_leapfrog, int* accepted, int* depth_out, int nthreads) {
    orc_nuts_batch_stat(h, x, r, N, eps, phi, max_depth, seed, iter, particle0, accrej, x_new, r_new, lp_old, lp_new,
                        n_leapfrog, accepted, depth_out, 0, nthreads);
}

/* same, also returning accept_stat[i] = accept_sum / n_leapfrog (nullable) */
void orc_nuts_batch_stat(void* h, const double* x, const double* r, long N, double eps, double phi, int max_depth,
                         uint64_t seed, uint32_t iter, uint64_t particle0, int accrej, double* x_new, double* r_new,
                         double* lp_old, double* lp_new, int* n_leapfrog, int* accepted, int* depth_out,
                         double* accept_stat, int nthreads) {
    orc_nuts_batch_metric(h, x, r, N, eps, phi, max_depth, seed, iter, particle0, accrej, x_new, r_new, lp_old, lp_new,
                          n_leapfrog, accepted, depth_out, accept_stat, 0, nthreads);
}

/* same with a diagonal metric: x and x_new are in x-space, the transition runs on z = x / scale (scale == NULL: identity) */
void orc_nuts_batch_metric(void* h, const double* x, const double* r, long N, double eps, double phi, int max_depth,
                           uint64_t seed, uint32_t iter, uint64_t particle0, int accrej, double* x_new, double* r_new,
                           double* lp_old, double* lp_new, int* n_leapfrog, int* accepted, int* depth_out,
                           double* accept_stat, const double* scale, int nthreads) {
    const Model* m = (const Model*)h;
    const int D = m->dim;
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (long i = 0; i < N; ++i) {
        Ctx c = {m, D, max_depth, eps, phi, 0.0, seed, particle0 + (uint64_t)i, 0, iter, 0, scale, 0.0, 0.0};
        double lp0, lps;
        int dep;
        if (scale) {
            double z0[MAXD];
            for (int d = 0; d < D; ++d) z0[d] = x[i * D + d] / scale[d];
            nuts_one(&c, z0, r + i * D, x_new + i * D, r_new + i * D, &lp0, &lps, &dep);
            for (int d = 0; d < D; ++d) x_new[i * D + d] = x_new[i * D + d] * scale[d];
        } else {
            nuts_one(&c, x + i * D, r + i * D, x_new + i * D, r_new + i * D, &lp0, &lps, &dep);
        }
        int acc = 1;
        if (accrej) {
            const double *xc = x + i * D, *rc = r + i * D;
            double *xn = x_new + i * D, *rn = r_new + i * D;
            double H1 = lps - (0.5 * dot(rn, rn, D));
            double H0 = lp0 - (0.5 * dot(rc, rc, D));
            double ratio = m_exp(H1 - H0);
            double prob = (ratio < 1.) ? ratio : 1.; /* python min(1., ratio): nan -> 1. */
            double u = stream_uniform(seed, iter, 2u, particle0 + (uint64_t)i, 0);
            int anyinf = 0;
            for (int d = 0; d < D; ++d) anyinf |= isinf(xn[d]);
            if (u > prob || anyinf) {
                acc = 0;
                memcpy(xn, xc, sizeof(double) * D); memcpy(rn, rc, sizeof(double) * D);
                lps = lp0;
            }
        }
        if (lp_old) lp_old[i] = lp0;
        if (lp_new) lp_new[i] = lps;
        if (n_leapfrog) n_leapfrog[i] = (int)c.n_leapfrog;
        if (accepted) accepted[i] = acc;
        if (depth_out) depth_out[i] = dep;
        if (accept_stat) accept_stat[i] = c.n_leapfrog ? c.accept_sum / (double)c.n_leapfrog : 0.0;
    }
}
