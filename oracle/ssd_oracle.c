/*
 * oracle/ssd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, from the definition, of the Mamba-2 SSD operator that the reference calls as
 *   mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D, z, dt_bias, ..., dt_softplus, dt_limit)
 * at SSD/MedSSD.py:361-375 (20 call sites).
 *
 * PARITY UNPINNED: the arithmetic lives in the third-party dependency mamba_ssm==2.2.2
 * (reference README.md:6), which is NOT vendored under /root/reference and cannot be installed
 * here (no network); the reference holds no test, golden vector or fixture for this call.  This
 * file therefore restates the *published* recurrence (Dao & Gu 2024, "Transformers are SSMs",
 * the state-space dual form; mamba_ssm's ssd_combined docstring contract):
 *
 *   dt_t   = clamp(softplus?(dt_raw_t + dt_bias_h), dt_min, dt_max)      softplus: x > 20 ? x : log1p(exp x)
 *   S_t    = exp(dt_t * A_h) * S_{t-1} + dt_t * x_t (outer) B_t          S: [P][N], S_{-1} = initial_states or 0
 *   y_t    = S_t C_t + D_h * x_t                                         (D per head, or per (head, p) if D_has_hdim)
 *   out_t  = y_t * silu(z_t)  if z is given
 *
 * evaluated strictly sequentially in double precision, so it is chunk_size independent (the
 * chunked algorithm is an exact re-association of this recurrence).  The analytic backward is
 * the adjoint of the same recurrence.  Independent cross-checks live in tests/test_oracle_ssd.py
 * (a numpy chunked "minimal SSD" restatement, A=0 prefix-sum known answers, chunk invariance).
 *
 * Layouts (dense, row-major, fp32 storage):
 *   x, z, out, dout, dx, dz : [batch][L][H][P]
 *   dt, ddt                 : [batch][L][H]
 *   A, dt_bias, D, dA, ddt_bias, dD : [H]   (D/dD: [H][P] when D_has_hdim)
 *   Bm, Cm, dB, dC          : [batch][L][G][N]     head h uses group h / (H/G)
 *   initial_states, final_states : [batch][H][P][N]
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double softplus(double x) { return x > 20.0 ? x : log1p(exp(x)); }
static inline double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

static inline double dt_eff(double raw, const float* dt_bias, int h, int dt_softplus, double dt_min, double dt_max) {
    double v = raw + (dt_bias ? (double)dt_bias[h] : 0.0);
    if (dt_softplus) v = softplus(v);
    if (v < dt_min) v = dt_min;
    if (v > dt_max) v = dt_max;
    return v;
}

int ssd_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ssd_oracle_fwd(int batch, int L, int H, int P, int G, int N,
                    const float* x, const float* dt, const float* A, const float* Bm, const float* Cm,
                    const float* D, int D_has_hdim, const float* z, const float* dt_bias, int dt_softplus,
                    double dt_min, double dt_max, const float* initial_states,
                    float* out, float* final_states) {
    const int hpg = H / G;
#pragma omp parallel
    {
        double* S = (double*)malloc((size_t)P * N * sizeof(double));
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < batch; ++b) {
            for (int h = 0; h < H; ++h) {
                const int g = h / hpg;
                if (initial_states)
                    for (int i = 0; i < P * N; ++i) S[i] = initial_states[(((size_t)b * H + h) * P) * N + i];
                else
                    memset(S, 0, (size_t)P * N * sizeof(double));
                for (int t = 0; t < L; ++t) {
                    const double dtv = dt_eff(dt[((size_t)b * L + t) * H + h], dt_bias, h, dt_softplus, dt_min, dt_max);
                    const double a = exp(dtv * (double)A[h]);
                    const float* xr = x + (((size_t)b * L + t) * H + h) * P;
                    const float* Bt = Bm + (((size_t)b * L + t) * G + g) * N;
                    const float* Ct = Cm + (((size_t)b * L + t) * G + g) * N;
                    float* orow = out + (((size_t)b * L + t) * H + h) * P;
                    for (int p = 0; p < P; ++p) {
                        const double xv = xr[p], dx = dtv * xv;
                        double y = 0;
                        double* Sp = S + (size_t)p * N;
                        for (int n = 0; n < N; ++n) {
                            Sp[n] = a * Sp[n] + dx * (double)Bt[n];
                            y += Sp[n] * (double)Ct[n];
                        }
                        if (D) y += xv * (double)(D_has_hdim ? D[(size_t)h * P + p] : D[h]);
                        if (z) { const double zz = z[(((size_t)b * L + t) * H + h) * P + p]; y *= zz * sigmoid(zz); }
                        orow[p] = (float)y;
                    }
                }
                if (final_states)
                    for (int i = 0; i < P * N; ++i) final_states[(((size_t)b * H + h) * P) * N + i] = (float)S[i];
            }
        }
        free(S);
    }
}

/* Backward.  States are kept every CK steps and re-derived inside a block so memory stays bounded. */
#define CK 32
void ssd_oracle_bwd(int batch, int L, int H, int P, int G, int N,
                    const float* x, const float* dt, const float* A, const float* Bm, const float* Cm,
                    const float* D, int D_has_hdim, const float* z, const float* dt_bias, int dt_softplus,
                    double dt_min, double dt_max, const float* initial_states, const float* dout,
                    float* dx, float* ddt, float* dA, float* dB, float* dC, float* dD, float* ddt_bias, float* dz) {
    const int hpg = H / G;
    const size_t PN = (size_t)P * N;
    const int nblk = (L + CK - 1) / CK;
    double* pA = (double*)calloc((size_t)batch * H, sizeof(double));
    double* pbias = (double*)calloc((size_t)batch * H, sizeof(double));
    double* pD = (double*)calloc((size_t)batch * H * (D_has_hdim ? P : 1), sizeof(double));
    double* accB = (double*)calloc((size_t)batch * L * G * N, sizeof(double));
    double* accC = (double*)calloc((size_t)batch * L * G * N, sizeof(double));

#pragma omp parallel
    {
        double* ck = (double*)malloc((size_t)(nblk + 1) * PN * sizeof(double));  /* state entering each block */
        double* blk = (double*)malloc((size_t)(CK + 1) * PN * sizeof(double));   /* states inside a block */
        double* dS = (double*)malloc(PN * sizeof(double));
        double* dyv = (double*)malloc((size_t)P * sizeof(double));
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < batch; ++b) {
            for (int g = 0; g < G; ++g) {
                for (int hh = 0; hh < hpg; ++hh) {
                    const int h = g * hpg + hh;
                    const double Ah = A[h];
                    /* pass 1: block-entry states */
                    double* S = ck;
                    if (initial_states)
                        for (size_t i = 0; i < PN; ++i) S[i] = initial_states[(((size_t)b * H + h) * P) * N + i];
                    else
                        memset(S, 0, PN * sizeof(double));
                    for (int k = 0; k < nblk; ++k) {
                        double* Sn = ck + (size_t)(k + 1) * PN;
                        memcpy(Sn, ck + (size_t)k * PN, PN * sizeof(double));
                        const int t1 = (k + 1) * CK < L ? (k + 1) * CK : L;
                        for (int t = k * CK; t < t1; ++t) {
                            const double dtv = dt_eff(dt[((size_t)b * L + t) * H + h], dt_bias, h, dt_softplus, dt_min, dt_max);
                            const double a = exp(dtv * Ah);
                            const float* xr = x + (((size_t)b * L + t) * H + h) * P;
                            const float* Bt = Bm + (((size_t)b * L + t) * G + g) * N;
                            for (int p = 0; p < P; ++p) {
                                const double dxv = dtv * (double)xr[p];
                                for (int n = 0; n < N; ++n) Sn[(size_t)p * N + n] = a * Sn[(size_t)p * N + n] + dxv * (double)Bt[n];
                            }
                        }
                    }
                    /* pass 2: reverse over blocks */
                    memset(dS, 0, PN * sizeof(double));
                    double sA = 0, sbias = 0;
                    for (int k = nblk - 1; k >= 0; --k) {
                        const int t0 = k * CK, t1 = (k + 1) * CK < L ? (k + 1) * CK : L;
                        memcpy(blk, ck + (size_t)k * PN, PN * sizeof(double));          /* blk[0] = S_{t0-1} */
                        for (int t = t0; t < t1; ++t) {
                            const double dtv = dt_eff(dt[((size_t)b * L + t) * H + h], dt_bias, h, dt_softplus, dt_min, dt_max);
                            const double a = exp(dtv * Ah);
                            const float* xr = x + (((size_t)b * L + t) * H + h) * P;
                            const float* Bt = Bm + (((size_t)b * L + t) * G + g) * N;
                            const double* Sp = blk + (size_t)(t - t0) * PN;
                            double* Sn = blk + (size_t)(t - t0 + 1) * PN;
                            for (int p = 0; p < P; ++p) {
                                const double dxv = dtv * (double)xr[p];
                                for (int n = 0; n < N; ++n) Sn[(size_t)p * N + n] = a * Sp[(size_t)p * N + n] + dxv * (double)Bt[n];
                            }
                        }
                        for (int t = t1 - 1; t >= t0; --t) {
                            const size_t tok = (size_t)b * L + t;
                            const double raw = (double)dt[tok * H + h] + (dt_bias ? (double)dt_bias[h] : 0.0);
                            const double dtv = dt_eff(dt[tok * H + h], dt_bias, h, dt_softplus, dt_min, dt_max);
                            const double a = exp(dtv * Ah);
                            const float* xr = x + (tok * H + h) * P;
                            const float* Bt = Bm + (tok * G + g) * N;
                            const float* Ct = Cm + (tok * G + g) * N;
                            const double* Sp = blk + (size_t)(t - t0) * PN;       /* S_{t-1} */
                            const double* St = blk + (size_t)(t - t0 + 1) * PN;   /* S_t */
                            double* aB = accB + (tok * G + g) * N;
                            double* aC = accC + (tok * G + g) * N;
                            /* dy_t (adjoint of y before the optional gate) */
                            for (int p = 0; p < P; ++p) {
                                double go = dout[(tok * H + h) * P + p];
                                if (z) {
                                    const double zz = z[(tok * H + h) * P + p], sg = sigmoid(zz);
                                    double y = 0;
                                    for (int n = 0; n < N; ++n) y += St[(size_t)p * N + n] * (double)Ct[n];
                                    if (D) y += (double)xr[p] * (double)(D_has_hdim ? D[(size_t)h * P + p] : D[h]);
                                    if (dz) dz[(tok * H + h) * P + p] = (float)(go * y * (sg * (1.0 + zz * (1.0 - sg))));
                                    go *= zz * sg;
                                }
                                dyv[p] = go;
                            }
                            double ga = 0, gdt = 0;
                            for (int p = 0; p < P; ++p) {
                                const double go = dyv[p], xv = xr[p];
                                double gx = 0;
                                for (int n = 0; n < N; ++n) {
                                    const double s = dS[(size_t)p * N + n] + go * (double)Ct[n];  /* adjoint of S_t */
                                    aC[n] += go * St[(size_t)p * N + n];
                                    aB[n] += s * dtv * xv;
                                    gx += s * (double)Bt[n];
                                    ga += s * Sp[(size_t)p * N + n];
                                    dS[(size_t)p * N + n] = s * a;
                                }
                                gdt += gx * xv;
                                double dxo = gx * dtv;
                                if (D) {
                                    const double Dv = D_has_hdim ? D[(size_t)h * P + p] : D[h];
                                    dxo += go * Dv;
                                    pD[((size_t)b * H + h) * (D_has_hdim ? P : 1) + (D_has_hdim ? p : 0)] += go * xv;
                                }
                                dx[(tok * H + h) * P + p] = (float)dxo;
                            }
                            gdt += ga * a * Ah;
                            sA += ga * a * dtv;
                            /* chain through clamp, softplus, bias */
                            double v = raw;
                            if (dt_softplus) v = softplus(raw);
                            if (v < dt_min || v > dt_max) gdt = 0;
                            if (dt_softplus && !(raw > 20.0)) gdt *= sigmoid(raw);
                            ddt[tok * H + h] = (float)gdt;
                            sbias += gdt;
                        }
                    }
                    pA[(size_t)b * H + h] = sA;
                    pbias[(size_t)b * H + h] = sbias;
                }
            }
        }
        free(ck); free(blk); free(dS); free(dyv);
    }
    for (int h = 0; h < H; ++h) {
        double sA = 0, sb = 0;
        for (int b = 0; b < batch; ++b) { sA += pA[(size_t)b * H + h]; sb += pbias[(size_t)b * H + h]; }
        dA[h] = (float)sA;
        if (ddt_bias) ddt_bias[h] = (float)sb;
        if (dD) {
            const int w = D_has_hdim ? P : 1;
            for (int p = 0; p < w; ++p) {
                double s = 0;
                for (int b = 0; b < batch; ++b) s += pD[((size_t)b * H + h) * w + p];
                dD[(size_t)h * w + p] = (float)s;
            }
        }
    }
    const size_t nbc = (size_t)batch * L * G * N;
    for (size_t i = 0; i < nbc; ++i) { dB[i] = (float)accB[i]; dC[i] = (float)accC[i]; }
    free(pA); free(pbias); free(pD); free(accB); free(accC);
}
