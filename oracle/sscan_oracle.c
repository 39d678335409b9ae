/*
 * oracle/sscan_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's Mamba-1 selective scan.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (medical_image_classification_b200)
 * never does.
 *
 * Forward follows selective_scan_ref, reference file
 *   CrossMamba/FusionMamba/mamba_ssm/ops/selective_scan_interface.py:92-158
 * step by step (line numbers quoted at each stage below).  The reference gets
 * its gradients from autograd through that graph; the backward here is the
 * analytic adjoint of the same recurrence (formulas as summarised from
 * selective_scan_bwd_kernel.cuh:278-296,439-453 in SURVEY.md section 2.2),
 * accumulated in REAL precision.
 *
 * Two instantiations: REAL=float (op-for-op like the fp32 reference) and
 * REAL=double (the "exact" recurrence parity reports also quote).
 *
 * Layouts (all dense, row-major, fp32 storage):
 *   u, delta, z, out, dout, du, ddelta, dz : [batch][dim][L]
 *   A, dA                                  : [dim][N]
 *   Bm, Cm, dB, dC                         : [batch][G][N][L]   (G divides dim;
 *        channel d uses group d / (dim/G), interface.py:134-137)
 *   D, delta_bias, dD, ddelta_bias         : [dim]
 *   last_state                             : [batch][dim][N]
 * Pinned against the reference itself: oracle/make_golden.py imports
 * selective_scan_ref from /root/reference and writes tests/golden/sscan_*.npz;
 * tests/test_oracle_golden.py checks this file against those vectors.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#error "compile with -DREAL=float -DSUFFIX=f32 or -DREAL=double -DSUFFIX=f64"
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

static inline REAL r_exp(REAL x) { return sizeof(REAL) == 4 ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static inline REAL r_log1p(REAL x) { return sizeof(REAL) == 4 ? (REAL)log1pf((float)x) : (REAL)log1p((double)x); }

/* F.softplus(x) with beta=1, threshold=20 (interface.py:112-113). */
static inline REAL softplus(REAL x) { return x > (REAL)20 ? x : r_log1p(r_exp(x)); }
static inline REAL sigmoid(REAL x) { return (REAL)1 / ((REAL)1 + r_exp(-x)); }

int FN(sscan_oracle_threads)(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void FN(sscan_oracle_fwd)(int batch, int dim, int L, int N, int G,
                          const float* u, const float* delta, const float* A,
                          const float* Bm, const float* Cm, const float* D,
                          const float* z, const float* delta_bias, int delta_softplus,
                          float* out, float* last_state) {
    const int rpg = dim / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < batch; ++b) {
        for (int d = 0; d < dim; ++d) {
            const int g = d / rpg;
            const float* ur = u + ((size_t)b * dim + d) * L;
            const float* dr = delta + ((size_t)b * dim + d) * L;
            const float* zr = z ? z + ((size_t)b * dim + d) * L : NULL;
            float* orow = out + ((size_t)b * dim + d) * L;
            const float* Bg = Bm + ((size_t)b * G + g) * N * L;
            const float* Cg = Cm + ((size_t)b * G + g) * N * L;
            REAL x[256];
            for (int n = 0; n < N; ++n) x[n] = 0;                    /* :125 zero initial state */
            for (int t = 0; t < L; ++t) {
                REAL dl = (REAL)dr[t];
                if (delta_bias) dl += (REAL)delta_bias[d];           /* :110-111 */
                if (delta_softplus) dl = softplus(dl);               /* :112-113 */
                const REAL ut = (REAL)ur[t];
                REAL y = 0;
                for (int n = 0; n < N; ++n) {
                    const REAL a = r_exp(dl * (REAL)A[(size_t)d * N + n]);      /* :127 deltaA */
                    const REAL bu = dl * (REAL)Bg[(size_t)n * L + t] * ut;      /* :135 deltaB_u */
                    x[n] = a * x[n] + bu;                                      /* :140 */
                    y += x[n] * (REAL)Cg[(size_t)n * L + t];                   /* :147 */
                }
                REAL o = D ? y + ut * (REAL)D[d] : y;                /* :154 */
                if (zr) { REAL zz = (REAL)zr[t]; o = o * (zz * sigmoid(zz)); } /* :155-156 */
                orow[t] = (float)o;
            }
            if (last_state)
                for (int n = 0; n < N; ++n) last_state[((size_t)b * dim + d) * N + n] = (float)x[n]; /* :148-149 */
        }
    }
}

/*
 * Backward.  dB/dC/dA/dD/ddelta_bias are reduced across rows; to stay
 * deterministic each (b, d) row writes private partials that are summed in a
 * fixed order afterwards.
 */
void FN(sscan_oracle_bwd)(int batch, int dim, int L, int N, int G,
                          const float* u, const float* delta, const float* A,
                          const float* Bm, const float* Cm, const float* D,
                          const float* z, const float* delta_bias, int delta_softplus,
                          const float* dout,
                          float* du, float* ddelta, float* dA, float* dB, float* dC,
                          float* dD, float* ddelta_bias, float* dz) {
    const int rpg = dim / G;
    const size_t nrows = (size_t)batch * dim;
    /* per-row partials for the weight grads */
    REAL* pA = (REAL*)calloc(nrows * N, sizeof(REAL));
    REAL* pD = (REAL*)calloc(nrows, sizeof(REAL));
    REAL* pbias = (REAL*)calloc(nrows, sizeof(REAL));
    /* dB/dC accumulate over the rpg channels of a group: one accumulator per (b, g), rows of a
       group are walked sequentially by the same thread so the order is fixed. */
    REAL* accB = (REAL*)calloc((size_t)batch * G * N * L, sizeof(REAL));
    REAL* accC = (REAL*)calloc((size_t)batch * G * N * L, sizeof(REAL));

#pragma omp parallel
    {
        REAL* xs = (REAL*)malloc((size_t)L * N * sizeof(REAL));   /* x_t for all t */
        REAL* as = (REAL*)malloc((size_t)L * N * sizeof(REAL));   /* a_t for all t */
        REAL* dl = (REAL*)malloc((size_t)L * sizeof(REAL));
        REAL* dy = (REAL*)malloc((size_t)L * sizeof(REAL));
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < batch; ++b) {
            for (int g = 0; g < G; ++g) {
                const float* Bg = Bm + ((size_t)b * G + g) * N * L;
                const float* Cg = Cm + ((size_t)b * G + g) * N * L;
                REAL* aB = accB + ((size_t)b * G + g) * N * L;
                REAL* aC = accC + ((size_t)b * G + g) * N * L;
                for (int dd = 0; dd < rpg; ++dd) {
                    const int d = g * rpg + dd;
                    const size_t row = (size_t)b * dim + d;
                    const float* ur = u + row * L;
                    const float* dr = delta + row * L;
                    const float* zr = z ? z + row * L : NULL;
                    const float* gor = dout + row * L;
                    /* forward recompute, keeping every state */
                    for (int t = 0; t < L; ++t) {
                        REAL v = (REAL)dr[t];
                        if (delta_bias) v += (REAL)delta_bias[d];
                        if (delta_softplus) v = softplus(v);
                        dl[t] = v;
                        REAL y = 0;
                        for (int n = 0; n < N; ++n) {
                            const REAL a = r_exp(v * (REAL)A[(size_t)d * N + n]);
                            const REAL xp = t ? xs[(size_t)(t - 1) * N + n] : (REAL)0;
                            const REAL xn = a * xp + v * (REAL)Bg[(size_t)n * L + t] * (REAL)ur[t];
                            as[(size_t)t * N + n] = a;
                            xs[(size_t)t * N + n] = xn;
                            y += xn * (REAL)Cg[(size_t)n * L + t];
                        }
                        REAL go = (REAL)gor[t];
                        if (zr) {
                            const REAL zz = (REAL)zr[t], sg = sigmoid(zz);
                            const REAL o = D ? y + (REAL)ur[t] * (REAL)D[d] : y;
                            if (dz) dz[row * L + t] = (float)(go * o * (sg * ((REAL)1 + zz * ((REAL)1 - sg))));
                            go = go * zz * sg;
                        }
                        dy[t] = go;
                    }
                    /* reverse sweep: gx[n] = d loss / d x_t[n] */
                    REAL gx[256];
                    for (int n = 0; n < N; ++n) gx[n] = 0;
                    REAL sD = 0, sbias = 0;
                    for (int t = L - 1; t >= 0; --t) {
                        const REAL go = dy[t], ut = (REAL)ur[t], v = dl[t];
                        REAL gu = D ? go * (REAL)D[d] : (REAL)0;
                        REAL gdl = 0;
                        sD += go * ut;
                        for (int n = 0; n < N; ++n) {
                            const REAL Bn = (REAL)Bg[(size_t)n * L + t], Cn = (REAL)Cg[(size_t)n * L + t];
                            const REAL a = as[(size_t)t * N + n];
                            const REAL xt = xs[(size_t)t * N + n];
                            const REAL xp = t ? xs[(size_t)(t - 1) * N + n] : (REAL)0;
                            const REAL gxn = gx[n] + go * Cn;           /* total adjoint of x_t[n] */
                            aC[(size_t)n * L + t] += go * xt;
                            aB[(size_t)n * L + t] += gxn * v * ut;
                            gu += gxn * v * Bn;
                            const REAL ga = gxn * xp;                   /* adjoint of a_t[n] */
                            const REAL An = (REAL)A[(size_t)d * N + n];
                            gdl += gxn * Bn * ut + ga * a * An;
                            pA[row * N + n] += ga * a * v;
                            gx[n] = gxn * a;                            /* flows to x_{t-1}[n] */
                        }
                        du[row * L + t] = (float)gu;
                        if (delta_softplus) {
                            REAL raw = (REAL)dr[t];
                            if (delta_bias) raw += (REAL)delta_bias[d];
                            if (!(raw > (REAL)20)) gdl *= sigmoid(raw);
                        }
                        ddelta[row * L + t] = (float)gdl;
                        sbias += gdl;
                    }
                    pD[row] = sD;
                    pbias[row] = sbias;
                }
            }
        }
        free(xs); free(as); free(dl); free(dy);
    }
    for (int d = 0; d < dim; ++d) {
        REAL sD = 0, sb = 0;
        for (int b = 0; b < batch; ++b) { sD += pD[(size_t)b * dim + d]; sb += pbias[(size_t)b * dim + d]; }
        if (dD) dD[d] = (float)sD;
        if (ddelta_bias) ddelta_bias[d] = (float)sb;
        for (int n = 0; n < N; ++n) {
            REAL s = 0;
            for (int b = 0; b < batch; ++b) s += pA[((size_t)b * dim + d) * N + n];
            dA[(size_t)d * N + n] = (float)s;
        }
    }
    const size_t nbc = (size_t)batch * G * N * L;
    for (size_t i = 0; i < nbc; ++i) { dB[i] = (float)accB[i]; dC[i] = (float)accC[i]; }
    free(pA); free(pD); free(pbias); free(accB); free(accC);
}
