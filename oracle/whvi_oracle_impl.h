/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the WHVI hot path.
 *
 * This header is included twice by whvi_oracle.c, once with REAL=float and once
 * with REAL=double.  It restates, in plain C, the algorithms of the reference
 * (ltdung/WHVI) for the batched FWHT and the WHVILinear forward/backward.  It is
 * the checker for the CUDA path; the product never links or calls it.
 *
 * Citations are into /root/reference/.
 */

/* ---- batched FWHT: src/fwht/cpp/fwht.cpp:3-21 -------------------------------
 * The reference clones the input, transposes it and runs
 *     h = 1; while (h < n) { for i in 0..n step 2h: for j in i..i+h:
 *         tmp = x[j]-x[j+h]; x[j] += x[j+h]; x[j+h] = tmp;  }  h *= 2 }
 * on whole columns.  Per row that is the loop below: natural (Sylvester) order,
 * unnormalised, out of place. */
static void FN(fwht_row)(REAL *v, int64_t n)
{
    for (int64_t h = 1; h < n; h *= 2)
        for (int64_t i = 0; i < n; i += 2 * h)
            for (int64_t j = i; j < i + h; ++j) {
                REAL a = v[j], b = v[j + h];
                v[j] = a + b;
                v[j + h] = a - b;
            }
}

void FN(oracle_fwht)(const REAL *in, REAL *out, int64_t rows, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        REAL *o = out + r * n;
        if (o != in + r * n) memcpy(o, in + r * n, (size_t)n * sizeof(REAL));
        FN(fwht_row)(o, n);
    }
}

/* ---- softplus as torch.nn.functional.softplus (beta=1, threshold=20),
 * used by g_sigma: src/weights.py:43-50 ------------------------------------- */
static REAL FN(softplus)(REAL r)
{
    return r > (REAL)20 ? r : (REAL)log1p(exp((double)r));
}
static REAL FN(sigmoid)(REAL r)
{
    return (REAL)(1.0 / (1.0 + exp(-(double)r)));
}

/* ---- reparameterisation: src/weights.py:82-83 and :92-93 --------------------
 * g_s = mu + softplus(rho) * eps_s   (one eps vector per forward call = per MC
 * sample).  dense != 0 selects the superset  g_s = mu + L eps_s  with L a D x D
 * row-major matrix (not in the reference: parity unpinned for dense L). */
void FN(oracle_reparam)(const REAL *mu, const REAL *rho_or_L, const REAL *eps,
                        REAL *g, int64_t S, int64_t D, int dense)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s)
        for (int64_t i = 0; i < D; ++i) {
            if (!dense) {
                g[s * D + i] = mu[i] + FN(softplus)(rho_or_L[i]) * eps[s * D + i];
            } else {
                double acc = 0.0;
                for (int64_t j = 0; j < D; ++j)
                    acc += (double)rho_or_L[i * D + j] * (double)eps[s * D + j];
                g[s * D + i] = (REAL)((double)mu[i] + acc);
            }
        }
}

/* backward of the diagonal reparameterisation:
 * dmu = sum_s dg_s ; drho = (sum_s dg_s * eps_s) * sigmoid(rho) */
void FN(oracle_reparam_bwd)(const REAL *rho, const REAL *eps, const REAL *dg,
                            REAL *dmu, REAL *drho, int64_t S, int64_t D)
{
    for (int64_t i = 0; i < D; ++i) {
        double a = 0.0, b = 0.0;
        for (int64_t s = 0; s < S; ++s) {
            a += (double)dg[s * D + i];
            b += (double)dg[s * D + i] * (double)eps[s * D + i];
        }
        dmu[i] = (REAL)a;
        drho[i] = (REAL)(b * (double)FN(sigmoid)(rho[i]));
    }
}

/* ---- PAPER forward: docstring src/weights.py:77,
 * W = S1 H diag(g) H S2 applied to a row:  y = s1 * H(g_s * H(s2 * x)) (+bias).
 * x is (S,B,D) with sample stride xs (xs == 0: the same (B,D) block is shared by
 * all samples), g is (S,D), y is (S,B,D). */
void FN(oracle_layer_fwd)(const REAL *x, int64_t xs, const REAL *g, const REAL *s1,
                          const REAL *s2, const REAL *bias, REAL *y,
                          int64_t S, int64_t B, int64_t D)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < S * B; ++r) {
        int64_t s = r / B, b = r % B;
        const REAL *xr = x + s * xs + b * D;
        const REAL *gs = g + s * D;
        REAL *t = y + r * D;
        for (int64_t i = 0; i < D; ++i) t[i] = s2[i] * xr[i];
        FN(fwht_row)(t, D);
        for (int64_t i = 0; i < D; ++i) t[i] = gs[i] * t[i];
        FN(fwht_row)(t, D);
        for (int64_t i = 0; i < D; ++i) t[i] = s1[i] * t[i] + (bias ? bias[i] : (REAL)0);
    }
}

/* ---- PAPER backward (verified there against autograd):
 * ds1 = sum dy*t4 ; dt3 = H(s1*dy) ; dg_s = sum_b dt3*t2 ; dt1 = H(g*dt3) ;
 * ds2 = sum dt1*x ; dx = s2*dt1 ; dbias = sum dy.
 * Accumulation is in double regardless of REAL so that the oracle is the more
 * accurate side of every comparison. dx has the (S,B,D) shape even when xs==0
 * (the caller sums over s in that case). */
void FN(oracle_layer_bwd)(const REAL *x, int64_t xs, const REAL *dy, const REAL *g,
                          const REAL *s1, const REAL *s2, REAL *dx, REAL *dg,
                          REAL *ds1, REAL *ds2, REAL *dbias,
                          int64_t S, int64_t B, int64_t D)
{
    double *a1 = (double *)calloc((size_t)D, sizeof(double));
    double *a2 = (double *)calloc((size_t)D, sizeof(double));
    double *ab = (double *)calloc((size_t)D, sizeof(double));
    double *ag = (double *)calloc((size_t)(S * D), sizeof(double));
    REAL *t2 = (REAL *)malloc((size_t)D * sizeof(REAL));
    REAL *t4 = (REAL *)malloc((size_t)D * sizeof(REAL));
    REAL *u = (REAL *)malloc((size_t)D * sizeof(REAL));
    for (int64_t r = 0; r < S * B; ++r) {
        int64_t s = r / B, b = r % B;
        const REAL *xr = x + s * xs + b * D;
        const REAL *dyr = dy + r * D;
        const REAL *gs = g + s * D;
        for (int64_t i = 0; i < D; ++i) t2[i] = s2[i] * xr[i];
        FN(fwht_row)(t2, D);
        for (int64_t i = 0; i < D; ++i) t4[i] = gs[i] * t2[i];
        FN(fwht_row)(t4, D);
        for (int64_t i = 0; i < D; ++i) {
            a1[i] += (double)dyr[i] * (double)t4[i];
            ab[i] += (double)dyr[i];
            u[i] = s1[i] * dyr[i];
        }
        FN(fwht_row)(u, D); /* dt3 */
        for (int64_t i = 0; i < D; ++i) {
            ag[s * D + i] += (double)u[i] * (double)t2[i];
            u[i] = gs[i] * u[i]; /* dt2 */
        }
        FN(fwht_row)(u, D); /* dt1 */
        for (int64_t i = 0; i < D; ++i) {
            a2[i] += (double)u[i] * (double)xr[i];
            dx[r * D + i] = s2[i] * u[i];
        }
    }
    for (int64_t i = 0; i < D; ++i) {
        ds1[i] = (REAL)a1[i];
        ds2[i] = (REAL)a2[i];
        if (dbias) dbias[i] = (REAL)ab[i];
    }
    for (int64_t i = 0; i < S * D; ++i) dg[i] = (REAL)ag[i];
    free(a1); free(a2); free(ab); free(ag); free(t2); free(t4); free(u);
}

/* ---- REFERENCE-AS-WRITTEN weight matrix: src/weights.py:66-73 (w_bar) with
 * matmul_diag_left = row scaling (src/utils.py:4-12) and fwht acting on rows
 * (src/fwht/python/fwht.py:32, :52-55):
 *     w_bar(u) = diag(s1) . fwht_rows( diag(u) . fwht_rows( diag(s2) ) )
 * W is D x D row-major.  Because both FWHTs act on rows with only row scalings
 * in between, H.H = D.I cancels and W = D.diag(s1*u*s2); this
 * function does NOT use that shortcut -- it performs the ops as written. */
void FN(oracle_ref_w_bar)(const REAL *u, const REAL *s1, const REAL *s2, REAL *W, int64_t D)
{
    for (int64_t i = 0; i < D; ++i) {
        REAL *row = W + i * D;
        for (int64_t j = 0; j < D; ++j) row[j] = (i == j) ? s2[i] : (REAL)0; /* torch.diag(s2) */
        FN(fwht_row)(row, D);                                  /* self.fwht(...) */
        for (int64_t j = 0; j < D; ++j) row[j] = u[i] * row[j]; /* matmul_diag_left(u, .) */
        FN(fwht_row)(row, D);                                  /* self.fwht(...) */
        for (int64_t j = 0; j < D; ++j) row[j] = s1[i] * row[j]; /* matmul_diag_left(s1, .) */
    }
}

/* sample_lrt as written, src/weights.py:87-93:  y = h @ (w_bar(mu) + w_bar(sigma*eps)).T
 * for one MC sample; h is (B,D), mu/sig_eps are (D,), y is (B,D). */
void FN(oracle_ref_sample_lrt)(const REAL *h, const REAL *mu, const REAL *sig_eps,
                               const REAL *s1, const REAL *s2, REAL *y,
                               int64_t B, int64_t D)
{
    REAL *W1 = (REAL *)malloc((size_t)(D * D) * sizeof(REAL));
    REAL *W2 = (REAL *)malloc((size_t)(D * D) * sizeof(REAL));
    FN(oracle_ref_w_bar)(mu, s1, s2, W1, D);
    FN(oracle_ref_w_bar)(sig_eps, s1, s2, W2, D);
    for (int64_t i = 0; i < D * D; ++i) W1[i] += W2[i];
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t o = 0; o < D; ++o) {
            double acc = 0.0;
            for (int64_t k = 0; k < D; ++k) acc += (double)h[b * D + k] * (double)W1[o * D + k];
            y[b * D + o] = (REAL)acc;
        }
    free(W1); free(W2);
}

/* PAPER weight matrix W = diag(s1) H diag(g) H diag(s2), dense D x D, for the
 * Column layer (src/weights.py:230-251 takes the first n entries of the
 * flattened sample = row 0) and for dense cross-checks. */
void FN(oracle_paper_weight)(const REAL *g, const REAL *s1, const REAL *s2, REAL *W, int64_t D)
{
    /* column j of W = s1 * H(g * H(s2_j e_j)); build by rows of W^T then transpose */
    REAL *t = (REAL *)malloc((size_t)D * sizeof(REAL));
    for (int64_t j = 0; j < D; ++j) {
        for (int64_t i = 0; i < D; ++i) t[i] = (i == j) ? s2[j] : (REAL)0;
        FN(fwht_row)(t, D);
        for (int64_t i = 0; i < D; ++i) t[i] = g[i] * t[i];
        FN(fwht_row)(t, D);
        for (int64_t i = 0; i < D; ++i) W[i * D + j] = s1[i] * t[i];
    }
    free(t);
}

/* ---- KL: src/utils.py:49-71 called as in src/weights.py:52-64 with
 * mu1 = g_mu, sd1 = softplus(g_rho), mu2 = 0, sd2 = lambda:
 *   0.5*( sum log sd2 - sum log sd1 - d + sum sd1/sd2 + (mu2-mu1).((mu2-mu1)/sd2) )
 * mode 0 = reference (sd used as a variance);
 * mode 1 = statistically consistent (sd1 -> sd1^2 in the log and ratio terms).
 * Also returns d/dmu and d/drho (Appendix A) when the pointers are non-null. */
double FN(oracle_kl)(const REAL *mu, const REAL *rho, double lambda_, int64_t D, int mode,
                     REAL *dmu, REAL *drho)
{
    double s_log = 0.0, s_ratio = 0.0, s_mu = 0.0;
    for (int64_t i = 0; i < D; ++i) {
        double sg = (double)FN(softplus)(rho[i]);
        double v = mode ? sg * sg : sg;
        s_log += log(v);
        s_ratio += v / lambda_;
        s_mu += (double)mu[i] * (double)mu[i] / lambda_;
        if (dmu) dmu[i] = (REAL)((double)mu[i] / lambda_);
        if (drho) {
            double sgm = (double)FN(sigmoid)(rho[i]);
            double dv = mode ? 2.0 * sg : 1.0; /* dv/dsigma */
            drho[i] = (REAL)(0.5 * (1.0 / lambda_ - 1.0 / v) * dv * sgm);
        }
    }
    return 0.5 * ((double)D * log(lambda_) - s_log - (double)D + s_ratio + s_mu);
}

/* ---- MNLL: src/likelihoods.py:18-29.  y (m,n_out), y_hat (m,n_out,n_mc):
 *   -n/(m*n_mc) * sum_{i,b,s} log N(y[b,i] | y_hat[b,i,s], sigma) */
double FN(oracle_mnll)(const REAL *y, const REAL *y_hat, double sigma, int64_t n,
                       int64_t m, int64_t n_out, int64_t n_mc)
{
    double acc = 0.0;
    const double lognorm = -log(sigma) - 0.5 * log(2.0 * M_PI);
    for (int64_t b = 0; b < m; ++b)
        for (int64_t i = 0; i < n_out; ++i)
            for (int64_t s = 0; s < n_mc; ++s) {
                double d = ((double)y[b * n_out + i] - (double)y_hat[(b * n_out + i) * n_mc + s]) / sigma;
                acc += lognorm - 0.5 * d * d;
            }
    return -(double)n / ((double)m * (double)n_mc) * acc;
}
