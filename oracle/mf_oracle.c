/*
 * oracle/mf_oracle.c -- CPU (fp64, single thread) restatement of the reference's
 * KernelMF / BaselineModel numeric loops.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path
 * (matrix_factorization_b200) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED against the reference run in the build container -- the
 * fixtures under tests/golden/ are produced by oracle/gen_golden.py, which imports the
 * reference package from /root/reference and calls its own njit functions; the
 * test-suite (tests/test_oracle_golden.py) checks every function below against them.
 * The reference ships no tests / golden vectors of its own.
 *
 * Every function cites the reference lines it follows (paths under /root/reference/).
 * Layout conventions: P is n_users x F row-major, Q is n_items x F row-major, ids int32,
 * ratings fp64.  kernel: 0 = linear, 1 = sigmoid, 2 = rbf.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LINEAR 0
#define ORC_SIGMOID 1
#define ORC_RBF 2

static double dot(const double *a, const double *b, int F) {
    double s = 0.0;
    for (int f = 0; f < F; ++f) s += a[f] * b[f];
    return s;
}

/* matrix_factorization/kernels.py:6-105 (sigmoid, kernel_linear, kernel_sigmoid, kernel_rbf) */
double orc_kmf_kernel(int kernel, double mu, double bu, double bi, const double *p,
                      const double *q, int F, double gamma, double a, double c) {
    if (kernel == ORC_LINEAR) {
        return mu + bi + bu + dot(p, q, F); /* kernels.py:41-44 */
    } else if (kernel == ORC_SIGMOID) {
        double x = mu + bu + bi + dot(p, q, F); /* kernels.py:72-75 */
        double s = 1.0 / (1.0 + exp(-x));       /* kernels.py:17 */
        return a + c * s;                       /* kernels.py:77 */
    } else {
        double d2 = 0.0; /* kernels.py:102-104 */
        for (int f = 0; f < F; ++f) {
            double d = p[f] - q[f];
            d2 += d * d;
        }
        return a + c * exp(-gamma * d2);
    }
}

/* One SGD step, in place.  kernels.py:108-180 (linear), :183-262 (sigmoid), :265-327 (rbf). */
void orc_kmf_update(int kernel, int u, int i, double r, double mu, double *bu, double *bi,
                    double *P, double *Q, int F, double lr, double reg, double gamma,
                    double a, double c, int upd_user, int upd_item) {
    double *p = P + (size_t)u * F;
    double *q = Q + (size_t)i * F;
    if (kernel == ORC_LINEAR) {
        double ub = bu[u], ib = bi[i];
        double pred = mu + ib + ub + dot(p, q, F); /* kernels.py:145-150 */
        double err = pred - r;                     /* :153 */
        if (upd_user) bu[u] -= lr * (err + reg * ub); /* :156-157 */
        if (upd_item) bi[i] -= lr * (err + reg * ib); /* :159-160 */
        for (int f = 0; f < F; ++f) {               /* :163-178, both use the OLD p_f, q_f */
            double pf = p[f], qf = q[f];
            if (upd_user) p[f] -= lr * (err * qf + reg * pf);
            if (upd_item) q[f] -= lr * (err * pf + reg * qf);
        }
    } else if (kernel == ORC_SIGMOID) {
        double ub = bu[u], ib = bi[i];
        double x = mu + ub + ib + dot(p, q, F); /* kernels.py:224-226 */
        double s = 1.0 / (1.0 + exp(-x));
        double pred = a + c * s;                /* :228 */
        double err = pred - r;                  /* :231 */
        double D = (s * s) * exp(-x);           /* :234, no factor c */
        if (upd_user) bu[u] -= lr * (err * D + reg * ub); /* :237-239 */
        if (upd_item) bi[i] -= lr * (err * D + reg * ib); /* :241-243 */
        for (int f = 0; f < F; ++f) {                     /* :246-260 */
            double pf = p[f], qf = q[f];
            if (upd_user) p[f] -= lr * (err * (qf * D) + reg * pf);
            if (upd_item) q[f] -= lr * (err * (pf * D) + reg * qf);
        }
    } else {
        double d2 = 0.0; /* kernels.py:301-303: no mu, no biases */
        for (int f = 0; f < F; ++f) {
            double d = p[f] - q[f];
            d2 += d * d;
        }
        double E = exp(-gamma * d2);
        double pred = a + c * E;
        double err = pred - r;      /* :306 */
        double D = 2.0 * E * gamma; /* :309, no factor c */
        for (int f = 0; f < F; ++f) { /* :312-325 */
            double pf = p[f], qf = q[f];
            if (upd_user) p[f] -= lr * (err * (D * (qf - pf)) + reg * pf);
            if (upd_item) q[f] -= lr * (err * (D * (pf - qf)) + reg * qf);
        }
    }
}

/*
 * Same-order replay: apply the reference update rule to the ratings in the
 * order given by order[0..n) (indices into u/i/r), i.e. the body of the rating loop at
 * kernel_matrix_factorization.py:374-425 with the shuffle replaced by an explicit order.
 * order == NULL means 0..n-1.
 */
void orc_kmf_replay(int kernel, const int32_t *u, const int32_t *i, const double *r,
                    const int64_t *order, int64_t n, double mu, double *bu, double *bi,
                    double *P, double *Q, int F, double lr, double reg, double gamma,
                    double min_rating, double max_rating, int upd_user, int upd_item) {
    double a = min_rating, c = max_rating - min_rating;
    for (int64_t k = 0; k < n; ++k) {
        int64_t j = order ? order[k] : k;
        orc_kmf_update(kernel, u[j], i[j], r[j], mu, bu, bi, P, Q, F, lr, reg, gamma, a, c,
                       upd_user, upd_item);
    }
}

/* kernel_matrix_factorization.py:240-317 (_calculate_rmse): unclipped predictions. */
double orc_kmf_rmse(int kernel, const int32_t *u, const int32_t *i, const double *r, int64_t n,
                    double mu, const double *bu, const double *bi, const double *P,
                    const double *Q, int F, double gamma, double min_rating,
                    double max_rating) {
    double a = min_rating, c = max_rating - min_rating, acc = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        double pred = orc_kmf_kernel(kernel, mu, bu[u[k]], bi[i[k]], P + (size_t)u[k] * F,
                                     Q + (size_t)i[k] * F, F, gamma, a, c);
        double e = r[k] - pred;
        acc += e * e;
    }
    return sqrt(acc / (double)n);
}

/* xorshift64* -- the port's own shuffle stream (numba's private MT19937 state used at
 * kernel_matrix_factorization.py:371 cannot be reproduced; any uniform shuffle is the same
 * algorithm statistically). */
static uint64_t rng_next(uint64_t *s) {
    uint64_t x = *s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1DULL;
}

static void shuffle_rows(int32_t *u, int32_t *i, double *r, int64_t n, uint64_t *state) {
    for (int64_t k = n - 1; k > 0; --k) { /* Fisher-Yates over whole rows, as np.random.shuffle(X) */
        int64_t j = (int64_t)(rng_next(state) % (uint64_t)(k + 1));
        int32_t tu = u[k]; u[k] = u[j]; u[j] = tu;
        int32_t ti = i[k]; i[k] = i[j]; i[j] = ti;
        double tr = r[k]; r[k] = r[j]; r[j] = tr;
    }
}

/*
 * kernel_matrix_factorization.py:320-445 (_sgd): per epoch shuffle rows, sequential
 * per-rating update, full RMSE pass.  u/i/r are shuffled in place like X is.
 */
void orc_kmf_sgd(int kernel, int32_t *u, int32_t *i, double *r, int64_t n, double mu,
                 double *bu, double *bi, double *P, double *Q, int F, int n_epochs, double lr,
                 double reg, double gamma, double min_rating, double max_rating, int upd_user,
                 int upd_item, uint64_t seed, double *train_rmse) {
    uint64_t state = seed ? seed : 0x9E3779B97F4A7C15ULL;
    for (int e = 0; e < n_epochs; ++e) {
        shuffle_rows(u, i, r, n, &state);
        orc_kmf_replay(kernel, u, i, r, NULL, n, mu, bu, bi, P, Q, F, lr, reg, gamma,
                       min_rating, max_rating, upd_user, upd_item);
        train_rmse[e] = orc_kmf_rmse(kernel, u, i, r, n, mu, bu, bi, P, Q, F, gamma,
                                     min_rating, max_rating);
    }
}

/* kernel_matrix_factorization.py:448-541 (_predict): -1 ids => bias 0 / zero vector. */
void orc_kmf_predict(int kernel, const int32_t *u, const int32_t *i, int64_t n, double mu,
                     const double *bu, const double *bi, const double *P, const double *Q,
                     int F, double gamma, double min_rating, double max_rating, int bound,
                     double *pred, uint8_t *possible) {
    double a = min_rating, c = max_rating - min_rating;
    double *zero = (double *)calloc((size_t)(F > 0 ? F : 1), sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        int uk = u[k] != -1, ik = i[k] != -1; /* :486-487 */
        double ub = uk ? bu[u[k]] : 0.0, ib = ik ? bi[i[k]] : 0.0;
        const double *p = uk ? P + (size_t)u[k] * F : zero;
        const double *q = ik ? Q + (size_t)i[k] * F : zero;
        double v = orc_kmf_kernel(kernel, mu, ub, ib, p, q, F, gamma, a, c);
        if (bound) { /* :532-536 */
            if (v > max_rating) v = max_rating;
            else if (v < min_rating) v = min_rating;
        }
        pred[k] = v;
        possible[k] = (uint8_t)(uk && ik);
    }
    free(zero);
}

/* ------------------------------ BaselineModel ------------------------------ */

/* baseline_model.py:183-212 (_calculate_rmse) */
double orc_bias_rmse(const int32_t *u, const int32_t *i, const double *r, int64_t n, double mu,
                     const double *bu, const double *bi) {
    double acc = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        double e = r[k] - (mu + bu[u[k]] + bi[i[k]]);
        acc += e * e;
    }
    return sqrt(acc / (double)n);
}

/* Rating loop of baseline_model.py:255-266 in an explicit order (error = rating - pred, +=). */
void orc_bias_replay(const int32_t *u, const int32_t *i, const double *r, const int64_t *order,
                     int64_t n, double mu, double *bu, double *bi, double lr, double reg,
                     int upd_user, int upd_item) {
    for (int64_t k = 0; k < n; ++k) {
        int64_t j = order ? order[k] : k;
        double err = r[j] - (mu + bu[u[j]] + bi[i[j]]);                  /* :259-260 */
        if (upd_user) bu[u[j]] += lr * (err - reg * bu[u[j]]);             /* :263-264 */
        if (upd_item) bi[i[j]] += lr * (err - reg * bi[i[j]]);             /* :265-266 */
    }
}

/* baseline_model.py:215-280 (_sgd) */
void orc_bias_sgd(int32_t *u, int32_t *i, double *r, int64_t n, double mu, double *bu,
                  double *bi, int n_epochs, double lr, double reg, int upd_user, int upd_item,
                  uint64_t seed, double *train_rmse) {
    uint64_t state = seed ? seed : 0x9E3779B97F4A7C15ULL;
    for (int e = 0; e < n_epochs; ++e) {
        shuffle_rows(u, i, r, n, &state);
        orc_bias_replay(u, i, r, NULL, n, mu, bu, bi, lr, reg, upd_user, upd_item);
        train_rmse[e] = orc_bias_rmse(u, i, r, n, mu, bu, bi);
    }
}

/* baseline_model.py:283-362 (_als): counts once, per epoch user pass from zeros, then item
 * pass with the NEW user biases, then RMSE.  Sums run in row order like the reference. */
void orc_bias_als(const int32_t *u, const int32_t *i, const double *r, int64_t n, double mu,
                  double *bu, double *bi, int n_users, int n_items, int n_epochs, double reg,
                  double *train_rmse) {
    double *cu = (double *)calloc((size_t)n_users, sizeof(double));
    double *ci = (double *)calloc((size_t)n_items, sizeof(double));
    for (int64_t k = 0; k < n; ++k) { /* :318-323 */
        cu[u[k]] += 1.0;
        ci[i[k]] += 1.0;
    }
    for (int e = 0; e < n_epochs; ++e) {
        memset(bu, 0, sizeof(double) * (size_t)n_users);              /* :329 */
        for (int64_t k = 0; k < n; ++k) bu[u[k]] += r[k] - mu - bi[i[k]]; /* :332-334 */
        for (int j = 0; j < n_users; ++j) bu[j] = bu[j] / (reg + cu[j]);  /* :337 */
        memset(bi, 0, sizeof(double) * (size_t)n_items);              /* :340 */
        for (int64_t k = 0; k < n; ++k) bi[i[k]] += r[k] - mu - bu[u[k]]; /* :343-345 */
        for (int j = 0; j < n_items; ++j) bi[j] = bi[j] / (reg + ci[j]);  /* :348 */
        train_rmse[e] = orc_bias_rmse(u, i, r, n, mu, bu, bi);        /* :351 */
    }
    free(cu);
    free(ci);
}

/* baseline_model.py:365-417 (_predict) */
void orc_bias_predict(const int32_t *u, const int32_t *i, int64_t n, double mu,
                      const double *bu, const double *bi, double min_rating, double max_rating,
                      int bound, double *pred, uint8_t *possible) {
    for (int64_t k = 0; k < n; ++k) {
        int uk = u[k] != -1, ik = i[k] != -1;
        double v = mu;
        if (uk) v += bu[u[k]];
        if (ik) v += bi[i[k]];
        if (bound) {
            if (v > max_rating) v = max_rating;
            else if (v < min_rating) v = min_rating;
        }
        pred[k] = v;
        possible[k] = (uint8_t)(uk && ik);
    }
}
