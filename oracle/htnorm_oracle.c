/* TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's sampling path in plain C99.
 * See htnorm_oracle.h for who may use this and for the parity status (pinned against oracle/_ref).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).  The
 * dense linear algebra the reference delegates to an un-vendored, unpinned Fortran BLAS/LAPACK
 * (src/htnorm_blas.h:9-49; `-lblas -llapack`, Makefile:17) is restated with textbook loops
 * (Netlib semantics: dpotrf/dpotrs lower, dtrmv, dtrsv, dsymv, dgemm), so results agree with the
 * reference+OpenBLAS to rounding (~1e-13 on well-conditioned inputs), not bit-for-bit; the integer
 * RNG streams and the ziggurat decisions ARE bit-for-bit.
 *
 * Build with -O2 -std=c99 -ffp-contract=off (no FMA contraction, as x86-64 gcc -O2 without -mfma).
 */
#include "htnorm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- ziggurat tables: data shared with the product (NumPy's tables, see THIRD_PARTY.md) ---- */
static const uint64_t zig_words[768] = {
#include "../htnorm_b200/csrc/common/zig_tables_u64.inc"
};
static inline uint64_t zig_ki(int i) { return zig_words[i]; }
static inline double zig_wi(int i) { double d; memcpy(&d, &zig_words[256 + i], 8); return d; }
static inline double zig_fi(int i) { double d; memcpy(&d, &zig_words[512 + i], 8); return d; }
/* src/htnorm_ziggurat_constants.h:1-2 */
static const double zig_r = 3.6541528853610087963519472518;
static const double zig_inv_r = 0.27366123732975827203338247596;

/* ---- bit generators ---- */

/* src/htnorm_splitmix64.h:17-24 */
static uint64_t splitmix64(uint64_t* s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* src/htnorm_rng.c:20-35 (pcg32_init x2, src/htnorm_pcg64.h:13-18) and src/htnorm_rng.c:45-60 */
void orc_rng_seed(orc_rng_t* g, int kind, uint64_t seed)
{
    g->kind = kind;
    g->ncalls = 0;
    g->s[0] = splitmix64(&seed);
    g->s[1] = splitmix64(&seed);
    if (kind == ORC_GEN_PCG64) {
        g->s[2] = splitmix64(&seed);
        g->s[3] = splitmix64(&seed);
    } else {
        g->s[2] = g->s[3] = 0;
    }
}

/* src/htnorm_pcg32_minimal.h:19-29 */
static uint32_t pcg32_step(uint64_t* state, uint64_t inc)
{
    uint64_t old = *state;
    *state = old * 6364136223846793005ULL + (inc | 1u);
    uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((0u - rot) & 31u));
}

/* src/htnorm_pcg64.h:23-28 and src/htnorm_xoroshiro128p.h:24-37 */
uint64_t orc_next_u64(orc_rng_t* g)
{
    g->ncalls++;
    if (g->kind == ORC_GEN_PCG64) {
        uint64_t hi = pcg32_step(&g->s[0], g->s[1]);
        uint64_t lo = pcg32_step(&g->s[2], g->s[3]);
        return (hi << 32) | lo;
    }
    uint64_t s0 = g->s[0], s1 = g->s[1];
    uint64_t res = s0 + s1;
    s1 ^= s0;
    g->s[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
    g->s[1] = (s1 << 37) | (s1 >> 27);
    return res;
}

/* src/htnorm_pcg64.h:31-37, src/htnorm_xoroshiro128p.h:40-45 */
double orc_next_double(orc_rng_t* g)
{
    return (double)(orc_next_u64(g) >> 11) * (1.0 / 9007199254740992.0);
}

/* src/htnorm_distributions.c:18-54 */
double orc_std_normal(orc_rng_t* g)
{
    for (;;) {
        uint64_t r = orc_next_u64(g);
        int idx = (int)(r & 0xff);
        r >>= 8;
        int sign = (int)(r & 1);
        uint64_t rabs = (r >> 1) & 0x000fffffffffffffULL;
        double x = (double)rabs * zig_wi(idx);
        if (sign)
            x = -x;
        if (rabs < zig_ki(idx))
            return x;
        if (idx == 0) {
            for (;;) {
                double xx = -zig_inv_r * log(1.0 - orc_next_double(g));
                double yy = -log(1.0 - orc_next_double(g));
                if (yy + yy > xx * xx)
                    return ((rabs >> 8) & 1) ? -(zig_r + xx) : zig_r + xx;
            }
        } else {
            double u = orc_next_double(g);
            if ((zig_fi(idx - 1) - zig_fi(idx)) * u + zig_fi(idx) < exp(-0.5 * x * x))
                return x;
        }
    }
}

static uint64_t seed_of(const uint64_t* seeds, uint64_t seed0, size_t i)
{
    return seeds ? seeds[i] : seed0 + (uint64_t)i;
}

void orc_raw_fill(int kind, const uint64_t* seeds, uint64_t seed0, size_t nstreams, size_t nwords,
                  uint64_t* out)
{
    for (size_t i = 0; i < nstreams; i++) {
        orc_rng_t g;
        orc_rng_seed(&g, kind, seed_of(seeds, seed0, i));
        for (size_t t = 0; t < nwords; t++)
            out[i * nwords + t] = orc_next_u64(&g);
    }
}

/* reverse fill: src/htnorm_distributions.c:12-13 */
static void normal_fill_rev(orc_rng_t* g, size_t n, double* v)
{
    for (size_t i = n; i--;)
        v[i] = orc_std_normal(g);
}

void orc_normal_fill(int kind, const uint64_t* seeds, uint64_t seed0, size_t nstreams, size_t n,
                     double* out, uint32_t* ndraws)
{
    for (size_t i = 0; i < nstreams; i++) {
        orc_rng_t g;
        orc_rng_seed(&g, kind, seed_of(seeds, seed0, i));
        normal_fill_rev(&g, n, out + i * n);
        if (ndraws)
            ndraws[i] = (uint32_t)g.ncalls;
    }
}

/* ---- dense helpers (Netlib semantics of the calls in src/htnorm_blas.h:60-121) ---- */

/* symmetric element as the reference's uplo='L' Fortran calls see a row-major array: the row-major
 * upper triangle */
static inline double sym(const double* a, int n, int i, int j)
{
    return i <= j ? a[(size_t)i * n + j] : a[(size_t)j * n + i];
}

/* dpotrf('L') (src/htnorm_blas.h:111-112).  Pivot test is `ajj <= 0`, false for NaN, so NaN
 * propagates with info = 0 as with OpenBLAS (tests/test_pyhtnorm.py:46-48). */
int orc_potrf(int n, const double* a, double* l)
{
    memset(l, 0, (size_t)n * n * sizeof(double));
    for (int j = 0; j < n; j++) {
        double ajj = sym(a, n, j, j);
        for (int t = 0; t < j; t++)
            ajj -= l[(size_t)j * n + t] * l[(size_t)j * n + t];
        if (ajj <= 0.0)
            return j + 1;
        ajj = sqrt(ajj);
        l[(size_t)j * n + j] = ajj;
        for (int i = j + 1; i < n; i++) {
            double s = sym(a, n, i, j);
            for (int t = 0; t < j; t++)
                s -= l[(size_t)i * n + t] * l[(size_t)j * n + t];
            l[(size_t)i * n + j] = s / ajj;
        }
    }
    return 0;
}

/* Cholesky of a small dense row-major symmetric matrix given in full (lower part used), in place:
 * on exit the lower triangle holds L.  dposv's factorisation step (src/htnorm_blas.h:120-121). */
static int chol_inplace_lower(int n, double* m)
{
    for (int j = 0; j < n; j++) {
        double ajj = m[(size_t)j * n + j];
        for (int t = 0; t < j; t++)
            ajj -= m[(size_t)j * n + t] * m[(size_t)j * n + t];
        if (ajj <= 0.0)
            return j + 1;
        ajj = sqrt(ajj);
        m[(size_t)j * n + j] = ajj;
        for (int i = j + 1; i < n; i++) {
            double s = m[(size_t)i * n + j];
            for (int t = 0; t < j; t++)
                s -= m[(size_t)i * n + t] * m[(size_t)j * n + t];
            m[(size_t)i * n + j] = s / ajj;
        }
    }
    return 0;
}

/* x := L^-1 x  then  x := L^-T x  (dpotrs lower), L row-major lower with leading dimension n */
static void chol_solve(int n, const double* l, double* x)
{
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int t = 0; t < i; t++)
            s -= l[(size_t)i * n + t] * x[t];
        x[i] = s / l[(size_t)i * n + i];
    }
    for (int i = n; i--;) {
        double s = x[i];
        for (int t = i + 1; t < n; t++)
            s -= l[(size_t)t * n + i] * x[t];
        x[i] = s / l[(size_t)i * n + i];
    }
}

/* x := L^-T x  (dtrsv 'L','T','N', src/htnorm_blas.h:106-107) */
static void trsv_lt(int n, const double* l, double* x)
{
    for (int i = n; i--;) {
        double s = x[i];
        for (int t = i + 1; t < n; t++)
            s -= l[(size_t)t * n + i] * x[t];
        x[i] = s / l[(size_t)i * n + i];
    }
}

/* ---- hyperplane-truncated MVN (src/htnorm.c:49-187, src/htnorm_distributions.c:58-88) ---- */

typedef struct {
    int k2, k, diag;
    const double *mean, *cov, *g, *r;
    double* l;      /* k x k lower factor (dense only) */
    double* cov_g;  /* k2 x k: row c = cov * g[c]^T  (src/htnorm.c:136-156, :68-73) */
    double* s;      /* k2 x k2 Cholesky factor of g cov g^T (general path), or the scalar g.cov_g (k2 == 1) */
    int info_cov;   /* dpotrf(cov) */
    int info_s;     /* dposv(g cov g^T) */
} ht_prep_t;

static void ht_prep_free(ht_prep_t* p)
{
    free(p->l);
    free(p->cov_g);
    free(p->s);
}

static int ht_prepare(ht_prep_t* p, int k2, int k, const double* mean, const double* cov,
                      const double* g, const double* r, int diag)
{
    memset(p, 0, sizeof(*p));
    p->k2 = k2; p->k = k; p->diag = diag;
    p->mean = mean; p->cov = cov; p->g = g; p->r = r;
    p->cov_g = malloc((size_t)k2 * k * sizeof(double));
    p->s = malloc((size_t)k2 * k2 * sizeof(double));
    if (!p->cov_g || !p->s)
        return -100;
    if (!diag) {
        p->l = malloc((size_t)k * k * sizeof(double));
        if (!p->l)
            return -100;
        p->info_cov = orc_potrf(k, cov, p->l); /* src/htnorm_distributions.c:75-77 */
        if (p->info_cov)
            return 0;
    }
    /* cov_g[c][j] = sum_i cov(j,i) g[c][i]   (dsymm / dsymv, or the diagonal loops) */
    for (int c = 0; c < k2; c++)
        for (int j = 0; j < k; j++) {
            double acc = 0.0;
            if (diag) {
                acc = cov[(size_t)j * k + j] * g[(size_t)c * k + j];
            } else {
                for (int i = 0; i < k; i++)
                    acc += sym(cov, k, j, i) * g[(size_t)c * k + i];
            }
            p->cov_g[(size_t)c * k + j] = acc;
        }
    /* s[a][c] = sum_j cov_g[a][j] g[c][j]   (dgemm TN, src/htnorm.c:164; ddot, :77) */
    for (int a = 0; a < k2; a++)
        for (int c = 0; c < k2; c++) {
            double acc = 0.0;
            for (int j = 0; j < k; j++)
                acc += p->cov_g[(size_t)a * k + j] * g[(size_t)c * k + j];
            p->s[(size_t)a * k2 + c] = acc;
        }
    if (k2 > 1)
        p->info_s = chol_inplace_lower(k2, p->s); /* dposv, src/htnorm.c:169 */
    return 0;
}

static int ht_draw(const ht_prep_t* p, orc_rng_t* g, double* out, double* gy /* k2 scratch */)
{
    const int k = p->k, k2 = p->k2;
    /* mvn_rand_cov, src/htnorm_distributions.c:58-88 */
    if (p->diag) {
        for (size_t i = k; i--;)
            out[i] = p->mean[i] + sqrt(p->cov[i * k + i]) * orc_std_normal(g);
    } else {
        if (p->info_cov)
            return p->info_cov; /* out untouched, generator not advanced (:78) */
        normal_fill_rev(g, k, out);
        /* dtrmv lower, in place: go bottom-up so inputs are still unmodified */
        for (int i = k; i--;) {
            double s = 0.0;
            for (int t = 0; t <= i; t++)
                s += p->l[(size_t)i * k + t] * out[t];
            out[i] = s;
        }
        for (int i = 0; i < k; i++)
            out[i] += p->mean[i];
    }
    if (k2 == 1) {
        /* hyperplane_truncated_norm_1d_g, src/htnorm.c:49-82 */
        double alpha = 0.0, d = 0.0;
        for (int i = 0; i < k; i++)
            d += p->g[i] * out[i];
        alpha = p->r[0] - d;
        alpha /= p->s[0];
        for (int i = 0; i < k; i++)
            out[i] += alpha * p->cov_g[i];
        return 0;
    }
    /* general path, src/htnorm.c:109-186 */
    for (int c = 0; c < k2; c++) {
        double d = 0.0;
        for (int i = 0; i < k; i++)
            d += p->g[(size_t)c * k + i] * out[i];
        gy[c] = p->r[c] - d;
    }
    if (p->info_s)
        return p->info_s; /* out holds the unprojected y */
    chol_solve(k2, p->s, gy);
    for (int j = 0; j < k; j++) {
        double acc = 0.0;
        for (int c = 0; c < k2; c++)
            acc += p->cov_g[(size_t)c * k + j] * gy[c];
        out[j] += acc;
    }
    return 0;
}

int orc_hyperplane(orc_rng_t* g, int k2, int k, const double* mean, const double* cov,
                   const double* gm, const double* r, int diag, double* out)
{
    ht_prep_t p;
    int rc = ht_prepare(&p, k2, k, mean, cov, gm, r, diag);
    if (!rc) {
        double* gy = malloc((size_t)(k2 > 0 ? k2 : 1) * sizeof(double));
        rc = gy ? ht_draw(&p, g, out, gy) : -100;
        free(gy);
    }
    ht_prep_free(&p);
    return rc;
}

int orc_hyperplane_batch(int kind, const uint64_t* seeds, uint64_t seed0, size_t nsamples,
                         int independent, int k2, int k, const double* mean, const double* cov,
                         const double* gm, const double* r, int diag, double* out, int* info)
{
    double* gy = malloc((size_t)(k2 > 0 ? k2 : 1) * sizeof(double));
    if (!gy)
        return -100;
    int first = 0;
    ht_prep_t p;
    int prepared = 0;
    for (size_t i = 0; i < nsamples; i++) {
        if (independent || !prepared) {
            size_t o = independent ? i : 0;
            if (prepared)
                ht_prep_free(&p);
            int rc = ht_prepare(&p, k2, k, mean + o * k, cov + o * (size_t)k * k,
                                gm + o * (size_t)k2 * k, r + o * k2, diag);
            prepared = 1;
            if (rc) {
                ht_prep_free(&p);
                free(gy);
                return rc;
            }
        }
        orc_rng_t g;
        orc_rng_seed(&g, kind, seed_of(seeds, seed0, i));
        int rc = ht_draw(&p, &g, out + i * (size_t)k, gy);
        if (info)
            info[i] = rc;
        if (rc && !first)
            first = rc;
    }
    if (prepared)
        ht_prep_free(&p);
    free(gy);
    return first;
}

/* ---- structured-precision MVN (src/htnorm.c:190-356, src/htnorm_distributions.c:91-125) ---- */

typedef struct {
    int n, p, a_id, o_id, struct_mean, paper_sign;
    const double *mean, *a, *phi, *omega;
    double* la;   /* p x p factor of A (NORMAL) */
    double* lo;   /* n x n factor of omega (NORMAL) */
    double* x;    /* n x p row-major: row i = A^-1 phi[i]^T  (src/htnorm.c:272-300); NULL for IDENTITY A */
    double* m;    /* n x n: Cholesky factor (lower) of omega^-1 + phi A^-1 phi^T */
    int info_a, info_o, info_m;
} sp_prep_t;

static void sp_prep_free(sp_prep_t* q)
{
    free(q->la);
    free(q->lo);
    free(q->x);
    free(q->m);
}

/* the value every double of an IDENTITY factor holds after memset(factor, 1, ...)
 * (src/htnorm_distributions.c:99) */
static double identity_factor_garbage(void)
{
    double d;
    memset(&d, 1, sizeof(d));
    return d;
}

static int sp_prepare(sp_prep_t* q, int n, int p, const double* mean, const double* a,
                      const double* phi, const double* omega, int struct_mean, int a_id, int o_id,
                      int paper_sign)
{
    memset(q, 0, sizeof(*q));
    q->n = n; q->p = p; q->a_id = a_id; q->o_id = o_id;
    q->struct_mean = struct_mean; q->paper_sign = paper_sign;
    q->mean = mean; q->a = a; q->phi = phi; q->omega = omega;
    q->m = calloc((size_t)n * n, sizeof(double));
    if (!q->m)
        return -100;
    if (a_id == ORC_NORMAL) {
        q->la = malloc((size_t)p * p * sizeof(double));
        if (!q->la)
            return -100;
        q->info_a = orc_potrf(p, a, q->la); /* src/htnorm_distributions.c:114-115 */
        if (q->info_a)
            return 0;
    }
    if (o_id == ORC_NORMAL) {
        q->lo = malloc((size_t)n * n * sizeof(double));
        if (!q->lo)
            return -100;
        q->info_o = orc_potrf(n, omega, q->lo);
        if (q->info_o)
            return 0;
    }
    /* omega^-1 (compute_precision_inverse, src/htnorm.c:190-214): only the part dposv reads is needed */
    if (o_id == ORC_NORMAL) {
        double* col = malloc((size_t)n * sizeof(double));
        if (!col)
            return -100;
        for (int c = 0; c < n; c++) {
            memset(col, 0, (size_t)n * sizeof(double));
            col[c] = 1.0;
            chol_solve(n, q->lo, col);
            for (int r = 0; r < n; r++)
                q->m[(size_t)r * n + c] = col[r];
        }
        free(col);
    } else {
        for (int i = 0; i < n; i++) {
            double d = (o_id == ORC_DIAGONAL) ? omega[(size_t)i * n + i] : identity_factor_garbage();
            q->m[(size_t)i * n + i] = 1 / d;
        }
    }
    /* x = A^-1 phi^T and  m += phi A^-1 phi^T  (src/htnorm.c:261-302) */
    if (a_id != ORC_IDENTITY) {
        q->x = malloc((size_t)n * p * sizeof(double));
        if (!q->x)
            return -100;
        for (int i = 0; i < n; i++) {
            double* row = q->x + (size_t)i * p;
            if (a_id == ORC_DIAGONAL) {
                for (int j = 0; j < p; j++)
                    row[j] = phi[(size_t)i * p + j] / a[(size_t)j * p + j];
            } else {
                memcpy(row, phi + (size_t)i * p, (size_t)p * sizeof(double));
                chol_solve(p, q->la, row);
            }
        }
    }
    const double* xx = (a_id == ORC_IDENTITY) ? phi : q->x;
    for (int r = 0; r < n; r++)
        for (int c = 0; c <= r; c++) {
            double acc = 0.0;
            for (int j = 0; j < p; j++)
                acc += xx[(size_t)r * p + j] * phi[(size_t)c * p + j];
            q->m[(size_t)r * n + c] += acc;
        }
    q->info_m = chol_inplace_lower(n, q->m); /* dposv, src/htnorm.c:330 */
    return 0;
}

/* mvn_rand_prec, src/htnorm_distributions.c:91-125 */
static void prec_draw(orc_rng_t* g, int n, int type, const double* prec, const double* l, double* v)
{
    if (type == ORC_IDENTITY) {
        normal_fill_rev(g, n, v);
    } else if (type == ORC_DIAGONAL) {
        for (size_t i = n; i--;)
            v[i] = orc_std_normal(g) / sqrt(prec[i * n + i]);
    } else {
        normal_fill_rev(g, n, v);
        trsv_lt(n, l, v);
    }
}

static int sp_draw(const sp_prep_t* q, orc_rng_t* g, double* out, double* y1, double* y2)
{
    const int n = q->n, p = q->p;
    if (q->info_a)
        return q->info_a;
    prec_draw(g, p, q->a_id, q->a, q->la, y1);
    if (q->info_o)
        return q->info_o;
    prec_draw(g, n, q->o_id, q->omega, q->lo, y2);
    /* y2 = phi y1 + y2   (src/htnorm.c:304-311) */
    for (int r = 0; r < n; r++) {
        double acc = 0.0;
        for (int j = 0; j < p; j++)
            acc += q->phi[(size_t)r * p + j] * y1[j];
        y2[r] += acc;
    }
    /* sign convention: the reference uses +1 (plain mean) / -1 (structured mean), src/htnorm.c:229,
     * 320,336,343; paper_sign flips both. */
    double coef;
    if (q->struct_mean) {
        for (int r = 0; r < n; r++)
            y2[r] = q->mean[r] - y2[r];
        for (int j = 0; j < p; j++)
            out[j] = y1[j];
        coef = -1.0;
    } else {
        for (int j = 0; j < p; j++)
            out[j] = y1[j] + q->mean[j];
        coef = 1.0;
    }
    if (q->paper_sign)
        coef = -coef;
    if (q->info_m)
        return q->info_m;
    chol_solve(n, q->m, y2);
    const double* xx = (q->a_id == ORC_IDENTITY) ? q->phi : q->x;
    for (int j = 0; j < p; j++) {
        double acc = 0.0;
        for (int r = 0; r < n; r++)
            acc += xx[(size_t)r * p + j] * y2[r];
        out[j] += coef * acc;
    }
    return 0;
}

int orc_structured_batch(int kind, const uint64_t* seeds, uint64_t seed0, size_t nsamples, int n,
                         int p, const double* mean, const double* a, const double* phi,
                         const double* omega, int struct_mean, int a_id, int o_id, int paper_sign,
                         double* out, int* info)
{
    sp_prep_t q;
    int rc = sp_prepare(&q, n, p, mean, a, phi, omega, struct_mean, a_id, o_id, paper_sign);
    double* y1 = malloc((size_t)p * sizeof(double));
    double* y2 = malloc((size_t)n * sizeof(double));
    int first = 0;
    if (!rc && y1 && y2) {
        for (size_t i = 0; i < nsamples; i++) {
            orc_rng_t g;
            orc_rng_seed(&g, kind, seed_of(seeds, seed0, i));
            int r1 = sp_draw(&q, &g, out + i * (size_t)p, y1, y2);
            if (info)
                info[i] = r1;
            if (r1 && !first)
                first = r1;
        }
    } else if (!rc) {
        rc = -100;
    }
    free(y1);
    free(y2);
    sp_prep_free(&q);
    return rc ? rc : first;
}

int orc_structured(orc_rng_t* g, int n, int p, const double* mean, const double* a,
                   const double* phi, const double* omega, int struct_mean, int a_id, int o_id,
                   int paper_sign, double* out)
{
    sp_prep_t q;
    int rc = sp_prepare(&q, n, p, mean, a, phi, omega, struct_mean, a_id, o_id, paper_sign);
    double* y1 = malloc((size_t)p * sizeof(double));
    double* y2 = malloc((size_t)n * sizeof(double));
    if (!rc)
        rc = (y1 && y2) ? sp_draw(&q, g, out, y1, y2) : -100;
    free(y1);
    free(y2);
    sp_prep_free(&q);
    return rc;
}
