/* pea_oracle_c.c -- plain-C restatement of the hot path's arithmetic (TEST INFRASTRUCTURE ONLY).
 *
 * A second, independent oracle next to oracle/pea_oracle.py: scalar loops, double accumulation, no library but libm.
 * It exists so that the PyTorch-based oracle (which the GPU parity tests compare against) is itself pinned by something
 * that shares no code with torch: tests/test_oracle_c.py checks this file against the golden vectors generated from the
 * reference's own classes / loss statements (tests/golden/, made by tests/golden/make_golden.py) and against the Python
 * oracle on seeded ragged cases.  Only tests/, __graft_entry__.build() (compile) and smoke()/bench cpu arms may touch
 * oracle/; the product never does.
 *
 * Each function cites the reference lines it follows (paths relative to the reference tree).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define OC_API __attribute__((visibility("default")))

static double gelu_exact(double x) { return 0.5 * x * (1.0 + erf(x * 0.70710678118654752440)); } /* nn.GELU(): erf form */

/* nn.LayerNorm(in_dim): train_sdxl_zh.py:47,60 / train_sd_zh.py:44,54 -- eps 1e-5, biased variance, affine. */
OC_API void oc_layernorm(const float* x, const float* gamma, const float* beta, double eps, int rows, int D, float* y) {
  for (int r = 0; r < rows; ++r) {
    const float* xr = x + (size_t)r * D;
    double mean = 0.0, var = 0.0;
    for (int i = 0; i < D; ++i) mean += xr[i];
    mean /= D;
    for (int i = 0; i < D; ++i) var += (xr[i] - mean) * (xr[i] - mean);
    var /= D;
    const double rstd = 1.0 / sqrt(var + eps);
    for (int i = 0; i < D; ++i) y[(size_t)r * D + i] = (float)((xr[i] - mean) * rstd * gamma[i] + beta[i]);
  }
}

/* nn.Linear: y = x W^T (+ b), W stored [out, in]. */
OC_API void oc_linear(const float* x, const float* w, const float* bias, int rows, int in, int out, float* y) {
  for (int r = 0; r < rows; ++r)
    for (int o = 0; o < out; ++o) {
      double acc = bias ? (double)bias[o] : 0.0;
      const float* xr = x + (size_t)r * in;
      const float* wr = w + (size_t)o * in;
      for (int i = 0; i < in; ++i) acc += (double)xr[i] * (double)wr[i];
      y[(size_t)r * out + o] = (float)acc;
    }
}

OC_API void oc_gelu(const float* x, long long n, float* y) {
  for (long long i = 0; i < n; ++i) y[i] = (float)gelu_exact((double)x[i]);
}

/* MLP.forward of train_sdxl_zh.py:58-67 (with fc: SDXL / MLP_plus) and train_sd_zh.py:53-56, tests/test_sd_zh.py:41-44
 * (fc_w == NULL: single output).  projector = nlin bias-free Linears with exact GELUs between them; dims[0..nlin] are
 * the layer widths (dims[0] = in_dim).  x2 = fc(GELU(x)); x1 = mean over ALL L tokens of the pre-GELU x (+ input when
 * use_residual).  Returns 0, or -1 if out of memory. */
OC_API int oc_adapter(const float* x, int B, int L, const float* ln_w, const float* ln_b, int nlin, const float* const* ws,
                      const int* dims, const float* fc_w, const float* fc_b, int out1, int use_residual, float* x1,
                      float* x2_or_x) {
  const int rows = B * L;
  int maxd = dims[0];
  for (int i = 1; i <= nlin; ++i)
    if (dims[i] > maxd) maxd = dims[i];
  float* a = (float*)malloc(sizeof(float) * (size_t)rows * maxd);
  float* b = (float*)malloc(sizeof(float) * (size_t)rows * maxd);
  if (!a || !b) { free(a); free(b); return -1; }
  oc_layernorm(x, ln_w, ln_b, 1e-5, rows, dims[0], a);
  for (int i = 0; i < nlin; ++i) {
    oc_linear(a, ws[i], NULL, rows, dims[i], dims[i + 1], b);
    if (i + 1 < nlin) oc_gelu(b, (long long)rows * dims[i + 1], b);
    float* t = a; a = b; b = t;
  }
  const int out = dims[nlin];          /* a = pre-GELU x, [rows, out] */
  if (!fc_w) {
    memcpy(x2_or_x, a, sizeof(float) * (size_t)rows * out);
  } else {
    oc_gelu(a, (long long)rows * out, b);
    oc_linear(b, fc_w, fc_b, rows, out, out1, x2_or_x);
    for (int bi = 0; bi < B; ++bi)
      for (int o = 0; o < out; ++o) {
        double acc = 0.0;
        for (int l = 0; l < L; ++l) {
          acc += a[((size_t)bi * L + l) * out + o];
          if (use_residual) acc += x[((size_t)bi * L + l) * dims[0] + o];   /* ctor asserts in_dim == out_dim */
        }
        x1[(size_t)bi * out + o] = (float)(acc / L);
      }
  }
  free(a); free(b);
  return 0;
}

/* attn2 core as diffusers AttnProcessor2_0 runs it under train_sdxl_zh.py:397,415: per head softmax(q k^T d^-1/2) v,
 * no mask, no dropout, non-causal.  q, o: [B, N, H*d]; k, v: [B, L, H*d] (heads = column slices). */
OC_API void oc_attn2_core(const float* q, const float* k, const float* v, int B, int N, int L, int H, int d, float* o) {
  const int C = H * d;
  const double scale = 1.0 / sqrt((double)d);
  double* p = (double*)malloc(sizeof(double) * (size_t)L);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int n = 0; n < N; ++n) {
        const float* qr = q + ((size_t)b * N + n) * C + h * d;
        double m = -INFINITY, sum = 0.0;
        for (int l = 0; l < L; ++l) {
          const float* kr = k + ((size_t)b * L + l) * C + h * d;
          double s = 0.0;
          for (int i = 0; i < d; ++i) s += (double)qr[i] * (double)kr[i];
          p[l] = s * scale;
          if (p[l] > m) m = p[l];
        }
        for (int l = 0; l < L; ++l) { p[l] = exp(p[l] - m); sum += p[l]; }
        for (int i = 0; i < d; ++i) {
          double acc = 0.0;
          for (int l = 0; l < L; ++l) acc += p[l] * (double)v[((size_t)b * L + l) * C + h * d + i];
          o[((size_t)b * N + n) * C + h * d + i] = (float)(acc / sum);
        }
      }
  free(p);
}

/* Its gradient as autograd forms it: dV = P^T dO; dP = dO V^T; dS = P o (dP - rowsum(P o dP)); dQ = dS K * scale;
 * dK = dS^T Q * scale.  dk / dv are OVERWRITTEN (zeroed here). */
OC_API void oc_attn2_core_bwd(const float* q, const float* k, const float* v, const float* dout, int B, int N, int L, int H, int d,
                              float* dq, float* dk, float* dv) {
  const int C = H * d;
  const double scale = 1.0 / sqrt((double)d);
  double* p = (double*)malloc(sizeof(double) * (size_t)L);
  double* ds = (double*)malloc(sizeof(double) * (size_t)L);
  double* dka = (double*)calloc((size_t)B * L * C, sizeof(double));
  double* dva = (double*)calloc((size_t)B * L * C, sizeof(double));
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int n = 0; n < N; ++n) {
        const float* qr = q + ((size_t)b * N + n) * C + h * d;
        const float* gr = dout + ((size_t)b * N + n) * C + h * d;
        double m = -INFINITY, sum = 0.0, delta = 0.0;
        for (int l = 0; l < L; ++l) {
          const float* kr = k + ((size_t)b * L + l) * C + h * d;
          double s = 0.0;
          for (int i = 0; i < d; ++i) s += (double)qr[i] * (double)kr[i];
          p[l] = s * scale;
          if (p[l] > m) m = p[l];
        }
        for (int l = 0; l < L; ++l) { p[l] = exp(p[l] - m); sum += p[l]; }
        for (int l = 0; l < L; ++l) {
          p[l] /= sum;
          const float* vr = v + ((size_t)b * L + l) * C + h * d;
          double dp = 0.0;
          for (int i = 0; i < d; ++i) dp += (double)gr[i] * (double)vr[i];
          ds[l] = dp;
          delta += p[l] * dp;
        }
        for (int l = 0; l < L; ++l) ds[l] = p[l] * (ds[l] - delta);
        for (int i = 0; i < d; ++i) {
          double acc = 0.0;
          for (int l = 0; l < L; ++l) acc += ds[l] * (double)k[((size_t)b * L + l) * C + h * d + i];
          dq[((size_t)b * N + n) * C + h * d + i] = (float)(acc * scale);
        }
        for (int l = 0; l < L; ++l)
          for (int i = 0; i < d; ++i) {
            dka[((size_t)b * L + l) * C + h * d + i] += ds[l] * (double)qr[i] * scale;
            dva[((size_t)b * L + l) * C + h * d + i] += p[l] * (double)gr[i];
          }
      }
  for (size_t i = 0; i < (size_t)B * L * C; ++i) { dk[i] = (float)dka[i]; dv[i] = (float)dva[i]; }
  free(p); free(ds); free(dka); free(dva);
}

/* mean_b(mean_chw((a - b)^2 * w_b)) and its gradient w.r.t. a, accumulated into da with factor `coef`
 * (the shape of every term of train_sdxl_zh.py:400-441). */
static double masked_mse(const float* a, const float* b, const double* w, int B, long long n, double coef, float* da, int* finite) {
  double total = 0.0;
  for (int bi = 0; bi < B; ++bi) {
    double acc = 0.0;
    for (long long i = 0; i < n; ++i) {
      const double df = (double)a[(size_t)bi * n + i] - (double)b[(size_t)bi * n + i];
      const double t = df * df * w[bi];
      if (finite && !isfinite(t)) *finite = 0;
      acc += t;
      if (da) da[(size_t)bi * n + i] += (float)(coef * 2.0 * df * w[bi] / ((double)n * B));
    }
    total += acc / (double)n;
  }
  return total / B;
}

/* train_sdxl_zh.py:400-441 (7 taps) / train_sd_zh.py:218-272 (9 taps, a tap whose weighted squared error contains a
 * NaN/Inf is skipped when skip_nonfinite):
 *   L0 = mean_b(mean((pred - noise)^2 zh_b));  L1 = mean_b(mean((pred - pred_t)^2 (1 - zh_b)));
 *   L2 = sum_taps mean_b(mean((f_T - f_S)^2 (1 - zh_b)));  L = L0 + L1 + fw L2.
 * out = {L, L0, L1, L2}; dpred and dfs[t] (nullable) receive dL/dpred and dL/df_S (overwritten). */
OC_API void oc_kd_loss(const float* pred, const float* noise, const float* pred_t, int B, long long n0, int ntaps,
                       const float* const* fs, const float* const* ft, const long long* numel, const long long* zh, double fw,
                       int skip_nonfinite, double* out, float* dpred, float* const* dfs) {
  double* w = (double*)malloc(sizeof(double) * (size_t)B);
  double* w1 = (double*)malloc(sizeof(double) * (size_t)B);
  for (int b = 0; b < B; ++b) { w[b] = (double)zh[b]; w1[b] = 1.0 - w[b]; }
  if (dpred) memset(dpred, 0, sizeof(float) * (size_t)B * n0);
  const double l0 = masked_mse(pred, noise, w, B, n0, 1.0, dpred, NULL);
  const double l1 = masked_mse(pred, pred_t, w1, B, n0, 1.0, dpred, NULL);
  double l2 = 0.0;
  for (int t = 0; t < ntaps; ++t) {
    float* g = dfs ? dfs[t] : NULL;
    if (g) memset(g, 0, sizeof(float) * (size_t)B * numel[t]);
    int finite = 1;
    const double term = masked_mse(fs[t], ft[t], w1, B, numel[t], fw, NULL, &finite);
    if (skip_nonfinite && !finite) continue;
    if (g) masked_mse(fs[t], ft[t], w1, B, numel[t], fw, g, NULL);
    l2 += term;
  }
  out[0] = l0 + l1 + fw * l2; out[1] = l0; out[2] = l1; out[3] = l2;
  free(w); free(w1);
}
