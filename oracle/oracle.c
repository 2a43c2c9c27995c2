/*
 * oracle.c -- CPU restatement of the Vivim Temporal-Mamba hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker the CUDA kernels are compared against.  Nothing in the product path
 * (vivim_b200/, causal_conv1d/, mamba_ssm/) may import, link or execute it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * It restates, in plain C with fp64 accumulation and O(L) memory traffic, the algorithms of
 *   - causal_conv1d_ref      /root/reference/causal-conv1d/causal_conv1d/causal_conv1d_interface.py:49-65
 *   - selective_scan_ref     /root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py:86-152
 * and their gradients.  The reference obtains gradients by torch autograd through the refs
 * (O(L^2) traffic for the scan); here they are written out analytically, following the
 * formulas the reference CUDA backward documents in
 *   /root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:279-295,439-453
 *   /root/reference/causal-conv1d/csrc/causal_conv1d_bwd.cu:153-222
 * Parity of this restatement is pinned by tests/test_oracle.py against golden vectors produced
 * by the real Python reference (tests/golden/make_golden.py, run in the build container).
 *
 * All tensors are dense, row-major fp32:  x,u,delta,z,out,dout: (B,D,L);  Bm,Cm: (B,G,N,L);
 * A: (D,N);  Dv, delta_bias, conv bias: (D);  conv weight: (D,K).  Optional pointers may be NULL.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline double sigmoid_d(double v) { return 1.0 / (1.0 + exp(-v)); }

/* F.softplus with the default threshold of 20 (selective_scan_interface.py:107). */
static inline double softplus_d(double v) { return v > 20.0 ? v : log1p(exp(v)); }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ causal depthwise conv1d */

/* out[b,d,t] = act(bias[d] + sum_k w[d,k] * x[b,d,t-(K-1)+k]); samples before t=0 are zero. */
void orc_conv1d_fwd(const float *x, const float *w, const float *bias, float *out,
                    int B, int D, int L, int K, int silu) {
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int d = 0; d < D; ++d) {
            const float *xr = x + ((size_t)b * D + d) * L;
            float *orow = out + ((size_t)b * D + d) * L;
            const float *wr = w + (size_t)d * K;
            const double bv = bias ? (double)bias[d] : 0.0;
            for (int t = 0; t < L; ++t) {
                double acc = bv;
                for (int k = 0; k < K; ++k) {
                    int s = t - (K - 1) + k;
                    if (s >= 0) acc += (double)wr[k] * (double)xr[s];
                }
                orow[t] = (float)(silu ? acc * sigmoid_d(acc) : acc);
            }
        }
    }
}

/* dx (B,D,L), dw (D,K), db (D) (db may be NULL).  dw/db are overwritten, not accumulated. */
void orc_conv1d_bwd(const float *x, const float *w, const float *bias, const float *dout,
                    float *dx, float *dw, float *db,
                    int B, int D, int L, int K, int silu) {
    #pragma omp parallel for schedule(static)
    for (int d = 0; d < D; ++d) {
        const float *wr = w + (size_t)d * K;
        const double bv = bias ? (double)bias[d] : 0.0;
        double dwacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double dbacc = 0.0;
        double *dpre = (double *)malloc(sizeof(double) * (size_t)L);
        for (int b = 0; b < B; ++b) {
            const float *xr = x + ((size_t)b * D + d) * L;
            const float *gr = dout + ((size_t)b * D + d) * L;
            float *dxr = dx + ((size_t)b * D + d) * L;
            for (int t = 0; t < L; ++t) {
                double g = (double)gr[t];
                if (silu) {
                    double pre = bv;
                    for (int k = 0; k < K; ++k) {
                        int s = t - (K - 1) + k;
                        if (s >= 0) pre += (double)wr[k] * (double)xr[s];
                    }
                    double sg = sigmoid_d(pre);
                    g *= sg * (1.0 + pre * (1.0 - sg));
                }
                dpre[t] = g;
                dbacc += g;
                for (int k = 0; k < K; ++k) {
                    int s = t - (K - 1) + k;
                    if (s >= 0) dwacc[k] += (double)xr[s] * g;
                }
            }
            for (int s = 0; s < L; ++s) {
                double acc = 0.0;
                for (int k = 0; k < K; ++k) {
                    int t = s + (K - 1) - k;
                    if (t < L) acc += (double)wr[k] * dpre[t];
                }
                dxr[s] = (float)acc;
            }
        }
        for (int k = 0; k < K; ++k) dw[(size_t)d * K + k] = (float)dwacc[k];
        if (db) db[d] = (float)dbacc;
        free(dpre);
    }
}

/* ------------------------------------------------------------------ selective scan */

/*
 * delta' = softplus?(delta + delta_bias);  h_t = exp(delta'_t A) h_{t-1} + delta'_t B_t u_t;
 * y_t = <C_t, h_t> + D u_t;  out = y;  out_z = y * silu(z).
 * out / out_z / last_state may each be NULL.  G divides D; channel d uses group d / (D/G).
 */
void orc_scan_fwd(const float *u, const float *delta, const float *A,
                  const float *Bm, const float *Cm, const float *Dv, const float *z,
                  const float *delta_bias, float *out, float *out_z, float *last_state,
                  int B, int D, int L, int N, int G, int softplus) {
    const int dpg = D / G;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int d = 0; d < D; ++d) {
            const size_t row = ((size_t)b * D + d) * L;
            const int g = d / dpg;
            const float *Br = Bm + ((size_t)b * G + g) * N * L;
            const float *Cr = Cm + ((size_t)b * G + g) * N * L;
            const float *Ar = A + (size_t)d * N;
            const double db = delta_bias ? (double)delta_bias[d] : 0.0;
            const double Dd = Dv ? (double)Dv[d] : 0.0;
            double h[256];
            for (int n = 0; n < N; ++n) h[n] = 0.0;
            for (int t = 0; t < L; ++t) {
                double dt = (double)delta[row + t] + db;
                if (softplus) dt = softplus_d(dt);
                const double ut = (double)u[row + t];
                double y = 0.0;
                for (int n = 0; n < N; ++n) {
                    h[n] = exp(dt * (double)Ar[n]) * h[n] + dt * (double)Br[(size_t)n * L + t] * ut;
                    y += (double)Cr[(size_t)n * L + t] * h[n];
                }
                y += Dd * ut;
                if (out) out[row + t] = (float)y;
                if (out_z) {
                    double zt = z ? (double)z[row + t] : 0.0;
                    out_z[row + t] = (float)(z ? y * zt * sigmoid_d(zt) : y);
                }
            }
            if (last_state)
                for (int n = 0; n < N; ++n) last_state[((size_t)b * D + d) * N + n] = (float)h[n];
        }
    }
}

/*
 * Gradients of orc_scan_fwd w.r.t. everything, given dout = d(loss)/d(out_z) (or d/d(out) when
 * z == NULL).  du, ddelta, dz: (B,D,L); dA: (D,N); dB, dC: (B,G,N,L); dD, ddelta_bias: (D).
 * All outputs are overwritten.  dz/dD/ddelta_bias may be NULL.
 */
void orc_scan_bwd(const float *u, const float *delta, const float *A,
                  const float *Bm, const float *Cm, const float *Dv, const float *z,
                  const float *delta_bias, const float *dout,
                  float *du, float *ddelta, float *dA, float *dB, float *dC,
                  float *dD, float *ddelta_bias, float *dz,
                  int B, int D, int L, int N, int G, int softplus) {
    const int dpg = D / G;
    const size_t bc_elems = (size_t)B * G * N * L;
    double *dBacc = (double *)calloc(bc_elems, sizeof(double));
    double *dCacc = (double *)calloc(bc_elems, sizeof(double));
    double *dAacc = (double *)calloc((size_t)D * N, sizeof(double));
    double *dDacc = (double *)calloc((size_t)D, sizeof(double));
    double *dbacc = (double *)calloc((size_t)D, sizeof(double));

    /* Parallel over groups-of-channels so that dB/dC rows are never shared between threads. */
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int g = 0; g < G; ++g) {
            double *hist = (double *)malloc(sizeof(double) * (size_t)L * N);
            double *dts = (double *)malloc(sizeof(double) * (size_t)L);
            double *gs = (double *)malloc(sizeof(double) * (size_t)L);
            const float *Br = Bm + ((size_t)b * G + g) * N * L;
            const float *Cr = Cm + ((size_t)b * G + g) * N * L;
            double *dBr = dBacc + ((size_t)b * G + g) * N * L;
            double *dCr = dCacc + ((size_t)b * G + g) * N * L;
            for (int d = g * dpg; d < (g + 1) * dpg; ++d) {
                const size_t row = ((size_t)b * D + d) * L;
                const float *Ar = A + (size_t)d * N;
                const double db = delta_bias ? (double)delta_bias[d] : 0.0;
                const double Dd = Dv ? (double)Dv[d] : 0.0;
                double h[256], adj[256], dA_loc[256];
                double dD_loc = 0.0, dbias_loc = 0.0;
                for (int n = 0; n < N; ++n) dA_loc[n] = 0.0;
                /* forward sweep: keep the whole state trajectory (L x N doubles) */
                for (int n = 0; n < N; ++n) h[n] = 0.0;
                for (int t = 0; t < L; ++t) {
                    double dt = (double)delta[row + t] + db;
                    if (softplus) dt = softplus_d(dt);
                    dts[t] = dt;
                    const double ut = (double)u[row + t];
                    double y = 0.0;
                    for (int n = 0; n < N; ++n) {
                        h[n] = exp(dt * (double)Ar[n]) * h[n] + dt * (double)Br[(size_t)n * L + t] * ut;
                        hist[(size_t)t * N + n] = h[n];
                        y += (double)Cr[(size_t)n * L + t] * h[n];
                    }
                    y += Dd * ut;
                    double go = (double)dout[row + t];
                    if (z) {
                        const double zt = (double)z[row + t];
                        const double sg = sigmoid_d(zt);
                        if (dz) dz[row + t] = (float)(go * y * sg * (1.0 + zt * (1.0 - sg)));
                        go *= zt * sg;
                    }
                    gs[t] = go;
                    dD_loc += go * ut;
                }
                /* reverse sweep: adj[n] = d(loss)/d(h_t[n]) */
                for (int n = 0; n < N; ++n) adj[n] = 0.0;
                for (int t = L - 1; t >= 0; --t) {
                    const double dt = dts[t];
                    const double ut = (double)u[row + t];
                    const double go = gs[t];
                    double du_t = Dd * go, ddt = 0.0;
                    for (int n = 0; n < N; ++n) {
                        const double An = (double)Ar[n];
                        const double Bt = (double)Br[(size_t)n * L + t];
                        const double ht = hist[(size_t)t * N + n];
                        const double a_next = (t + 1 < L) ? exp(dts[t + 1] * An) : 0.0;
                        const double dh = go * (double)Cr[(size_t)n * L + t] + a_next * adj[n];
                        adj[n] = dh;
                        const double ah = ht - dt * Bt * ut; /* = exp(dt A) h_{t-1} */
                        du_t += dh * dt * Bt;
                        ddt += dh * (Bt * ut + An * ah);
                        dA_loc[n] += dh * dt * ah;
                        dBr[(size_t)n * L + t] += dh * dt * ut;
                        dCr[(size_t)n * L + t] += go * ht;
                    }
                    if (softplus) {
                        const double v = (double)delta[row + t] + db;
                        if (v <= 20.0) ddt *= sigmoid_d(v);
                    }
                    du[row + t] = (float)du_t;
                    ddelta[row + t] = (float)ddt;
                    dbias_loc += ddt;
                }
                for (int n = 0; n < N; ++n) {
                    #pragma omp atomic
                    dAacc[(size_t)d * N + n] += dA_loc[n];
                }
                #pragma omp atomic
                dDacc[d] += dD_loc;
                #pragma omp atomic
                dbacc[d] += dbias_loc;
            }
            free(hist); free(dts); free(gs);
        }
    }
    for (size_t i = 0; i < bc_elems; ++i) { dB[i] = (float)dBacc[i]; dC[i] = (float)dCacc[i]; }
    for (size_t i = 0; i < (size_t)D * N; ++i) dA[i] = (float)dAacc[i];
    if (dD) for (int d = 0; d < D; ++d) dD[d] = (float)dDacc[d];
    if (ddelta_bias) for (int d = 0; d < D; ++d) ddelta_bias[d] = (float)dbacc[d];
    free(dBacc); free(dCacc); free(dAacc); free(dDacc); free(dbacc);
}
