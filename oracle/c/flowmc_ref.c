/* CPU restatement (plain C + OpenMP) of flowMC's local-step hot path.  TEST / BASELINE ONLY.
 *
 * Second, independent restatement of the algorithm that oracle/local.py restates in numpy; it
 * must agree with it bit-for-bit on RNG words and accept flags (tests/test_oracle_c.py).  It is
 * also the CPU baseline timed by bench.py (`--impl reference`, kind "port"): the reference
 * itself is Python over jax/XLA, which is not installable in this image, so the closest CPU
 * stand-in is the same algorithm compiled natively and threaded over chains.
 *
 * Follows (paths relative to the flowMC tree):
 *   take_serial_steps   src/flowMC/strategy/take_steps.py:60-144,156-180
 *   mala_step           src/flowMC/resource/kernel/MALA.py:26-89
 *   hmc_step            src/flowMC/resource/kernel/HMC.py:47-50,71-96,98-151
 *   grw_step            src/flowMC/resource/kernel/Gaussian_random_walk.py:25-61
 *   threefry / normal   jax 0.5.0 jax/_src/prng.py, jax/_src/random.py (partitionable threefry),
 *                       XLA ErfInv32 polynomial
 * Unlike the CUDA path it recomputes value_and_grad at the current point every MALA step,
 * exactly as the reference does (MALA.py:59,71-73).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 512

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* o0, uint32_t* o1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int g = 0; g < 5; ++g) {
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl32(x1, R[g & 1][i]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  *o0 = x0;
  *o1 = x1;
}

/* random_bits(key, (n,)) for n counters 0..n-1, written as array loops so the compiler vectorises */
static void bits_vec(uint32_t k0, uint32_t k1, int n, uint32_t* restrict out) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0[MAXD], x1[MAXD];
  for (int j = 0; j < n; ++j) {
    x0[j] = ks[0];
    x1[j] = (uint32_t)j + ks[1];
  }
  for (int g = 0; g < 5; ++g) {
    for (int i = 0; i < 4; ++i) {
      const int r = R[g & 1][i];
      for (int j = 0; j < n; ++j) {
        x0[j] += x1[j];
        x1[j] = (x1[j] << r) | (x1[j] >> (32 - r));
        x1[j] ^= x0[j];
      }
    }
    const uint32_t a = ks[(g + 1) % 3], b = ks[(g + 2) % 3] + (uint32_t)(g + 1);
    for (int j = 0; j < n; ++j) {
      x0[j] += a;
      x1[j] += b;
    }
  }
  for (int j = 0; j < n; ++j) out[j] = x0[j] ^ x1[j];
}

static inline float bits_to_unit(uint32_t b) {
  union { uint32_t u; float f; } v;
  v.u = (b >> 9) | 0x3F800000u;
  return v.f - 1.0f;
}

static inline float erf_inv32(float x) {
  static const float A[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                             -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
  static const float B[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                             -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
  float w = -log1pf(-(x * x));
  const float* c;
  if (w < 5.0f) {
    w = w - 2.5f;
    c = A;
  } else {
    w = sqrtf(w) - 3.0f;
    c = B;
  }
  float p = c[0];
  for (int i = 1; i < 9; ++i) {
    float t = p * w; /* separate multiply and add, like the numpy oracle */
    p = c[i] + t;
  }
  return p * x;
}

static inline float bits_to_normal(uint32_t b) {
  const float lo = -0.99999994f;
  float f = bits_to_unit(b);
  float t = f * 2.0f;
  float u = t + lo;
  if (u < lo) u = lo;
  return 1.41421356237f * erf_inv32(u);
}

/* ---------------------------------------------------------------- targets (oracle/targets.py) */
enum { T_ISO = 0, T_DUALMOON = 1, T_AR1 = 2, T_DENSE = 3, T_ROSEN = 4, T_MIX = 5 };

static float lse2f(float a, float b) {
  float m = a > b ? a : b;
  return m + logf(expf(a - m) + expf(b - m));
}

static float logp_grad(int tgt, const float* data, const float* x, int d, float* g, int want_grad) {
  float lp = 0.0f;
  switch (tgt) {
    case T_ISO: {
      float c = data[0], s = 0.0f;
      for (int j = 0; j < d; ++j) {
        float r = x[j] - data[1 + j];
        s += r * r;
        if (want_grad) g[j] = -2.0f * c * r;
      }
      lp = -c * s;
    } break;
    case T_DUALMOON: {
      float s = 0.0f;
      for (int j = 0; j < d; ++j) {
        float r = x[j] - data[j];
        s += r * r;
      }
      float nrm = sqrtf(s), t = (nrm - 2.0f) / 0.1f;
      float a = (x[0] - 3.0f) / 0.8f, b = (x[0] + 3.0f) / 0.8f;
      float ta = -0.5f * a * a, tb = -0.5f * b * b;
      float c = (x[1] - 3.0f) / 0.6f, e = (x[1] + 3.0f) / 0.6f;
      float tc = -0.5f * c * c, te = -0.5f * e * e;
      float l2 = lse2f(ta, tb), l3 = lse2f(tc, te);
      lp = -(0.5f * t * t - l2 - l3);
      if (want_grad) {
        for (int j = 0; j < d; ++j) g[j] = -(t / 0.1f) * ((x[j] - data[j]) / nrm);
        g[0] += expf(ta - l2) * (-a / 0.8f) + expf(tb - l2) * (-b / 0.8f);
        g[1] += expf(tc - l3) * (-c / 0.6f) + expf(te - l3) * (-e / 0.6f);
      }
    } break;
    case T_AR1: {
      float rho = data[0], a = 1.0f / (1.0f - rho * rho), s = 0.0f;
      for (int j = 0; j < d; ++j) {
        float left = j > 0 ? x[j - 1] : 0.0f, right = j < d - 1 ? x[j + 1] : 0.0f;
        float diag = (j > 0 && j < d - 1) ? 1.0f + rho * rho : 1.0f;
        float px = a * (diag * x[j] - rho * (left + right));
        s += x[j] * px;
        if (want_grad) g[j] = -px;
      }
      lp = -0.5f * s;
    } break;
    case T_DENSE: {
      float s = 0.0f;
      for (int j = 0; j < d; ++j) {
        float px = 0.0f;
        for (int i = 0; i < d; ++i) px += data[(size_t)j * d + i] * x[i];
        s += x[j] * px;
        if (want_grad) g[j] = -px;
      }
      lp = -0.5f * s;
    } break;
    case T_ROSEN: {
      float s = 0.0f;
      if (want_grad) memset(g, 0, sizeof(float) * d);
      for (int j = 0; j < d - 1; ++j) {
        float u = x[j + 1] - x[j] * x[j], v = 1.0f - x[j];
        s += (100.0f * u * u + v * v) / 20.0f;
        if (want_grad) {
          g[j] += (400.0f * x[j] * u + 2.0f * v) / 20.0f;
          g[j + 1] += (-200.0f * u) / 20.0f;
        }
      }
      lp = -s;
    } break;
    case T_MIX: {
      int K = (int)data[0];
      float iv = data[1];
      const float* logw = data + 2;
      const float* mu = data + 2 + K;
      float e[8], m = -INFINITY;
      for (int k = 0; k < K; ++k) {
        float sq = 0.0f;
        for (int j = 0; j < d; ++j) {
          float r = mu[(size_t)k * d + j] - x[j];
          sq += r * r;
        }
        e[k] = logw[k] - 0.5f * iv * sq;
        if (e[k] > m) m = e[k];
      }
      float s = 0.0f;
      for (int k = 0; k < K; ++k) {
        e[k] = expf(e[k] - m);
        s += e[k];
      }
      lp = m + logf(s);
      if (want_grad) {
        for (int j = 0; j < d; ++j) {
          float gj = 0.0f;
          for (int k = 0; k < K; ++k) gj += (e[k] / s) * (mu[(size_t)k * d + j] - x[j]);
          g[j] = iv * gj;
        }
      }
    } break;
  }
  return lp;
}

/* ---------------------------------------------------------------- kernels */
typedef struct {
  int kind;          /* 0 MALA, 1 HMC, 2 GRW */
  float step_size;
  int n_leapfrog;
  const float* chol;   /* [d,d] */
  const float* colsum; /* [d] */
} KernelCfg;

static float mvn_scalar(const float* x, const float* mean, int d, float cov) {
  float yy = 0.0f;
  for (int j = 0; j < d; ++j) {
    float y = x[j] - mean[j];
    yy += y * y;
  }
  return -0.5f * yy / cov - (float)d / 2.0f * (1.8378770664093453f + logf(cov));
}

/* one kernel.kernel() application; dbg (optional) receives {MH log-ratio, log(uniform)} of the accept test */
static int one_step(const KernelCfg* k, int tgt, const float* data, int d, uint32_t s0, uint32_t s1, float* x,
                    float* lp_io, float* dbg) {
  uint32_t key1[2], key2[2], bits[MAXD], a, b;
  float z[MAXD], prop[MAXD], g0[MAXD], g1[MAXD], m0[MAXD], m1[MAXD];
  threefry2x32(s0, s1, 0, 0, &key1[0], &key1[1]);
  threefry2x32(s0, s1, 0, 1, &key2[0], &key2[1]);
  bits_vec(key1[0], key1[1], d, bits);
  for (int j = 0; j < d; ++j) z[j] = bits_to_normal(bits[j]);
  threefry2x32(key2[0], key2[1], 0, 0, &a, &b);
  float u = bits_to_unit(a ^ b);
  if (u < 0.0f) u = 0.0f;
  const float log_u = logf(u);
  int acc = 0;
  if (k->kind == 0) {
    const float dt = k->step_size, dt2 = dt * dt;
    float lp0 = logp_grad(tgt, data, x, d, g0, 1);
    for (int j = 0; j < d; ++j) {
      m0[j] = x[j] + (dt2 * g0[j]) / 2.0f;
      float t = dt * z[j];
      prop[j] = m0[j] + t;
    }
    float lp1 = logp_grad(tgt, data, prop, d, g1, 1);
    for (int j = 0; j < d; ++j) m1[j] = prop[j] + (dt2 * g1[j]) / 2.0f;
    float ratio = lp1 - lp0;
    ratio = ratio - mvn_scalar(prop, m0, d, dt2);
    ratio = ratio + mvn_scalar(x, m1, d, dt2);
    acc = log_u < ratio;
    if (dbg) dbg[0] = ratio;
    if (acc) memcpy(x, prop, sizeof(float) * d);
    *lp_io = acc ? lp1 : lp0;
  } else if (k->kind == 2) {
    for (int j = 0; j < d; ++j) prop[j] = x[j] + z[j] * k->step_size;
    float lp1 = logp_grad(tgt, data, prop, d, g1, 0);
    acc = log_u < (lp1 - *lp_io);
    if (dbg) dbg[0] = lp1 - *lp_io;
    if (acc) {
      memcpy(x, prop, sizeof(float) * d);
      *lp_io = lp1;
    }
  } else {
    float p[MAXD], xs[MAXD];
    const float eps = k->step_size;
    for (int i = 0; i < d; ++i) {
      float s = 0.0f;
      for (int j = 0; j < d; ++j) s += z[j] * k->chol[(size_t)i * d + j];
      p[i] = s;
    }
    float kin = 0.0f;
    for (int j = 0; j < d; ++j) kin += p[j] * p[j] * k->colsum[j];
    const float H = -(*lp_io) + 0.5f * kin;
    memcpy(xs, x, sizeof(float) * d);
    float lp1 = 0.0f;
    for (int it = 0; it < k->n_leapfrog + 2; ++it) {
      const float c0 = it == 0 ? 0.0f : 1.0f;
      const float c1 = (it == 0 || it == k->n_leapfrog + 1) ? 0.5f : 1.0f;
      for (int j = 0; j < d; ++j) xs[j] = xs[j] + eps * c0 * (p[j] * k->colsum[j]);
      lp1 = logp_grad(tgt, data, xs, d, g1, 1);
      for (int j = 0; j < d; ++j) p[j] = p[j] - eps * c1 * (-g1[j]);
    }
    float kin1 = 0.0f;
    for (int j = 0; j < d; ++j) kin1 += p[j] * p[j] * k->colsum[j];
    const float ham = -lp1 + 0.5f * kin1;
    acc = log_u < (H - ham);
    if (dbg) dbg[0] = H - ham;
    if (acc) {
      memcpy(x, xs, sizeof(float) * d);
      *lp_io = lp1;
    }
  }
  if (dbg) dbg[1] = log_u;
  return acc;
}

/* exported ------------------------------------------------------------------------------- */
void ref_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* out) {
  threefry2x32(k0, k1, c0, c1, &out[0], &out[1]);
}

void ref_normal(const uint32_t* key, int n, float* out) {
  for (int base = 0; base < n; base += MAXD) {
    /* counters continue across blocks */
    int m = n - base < MAXD ? n - base : MAXD;
    for (int j = 0; j < m; ++j) {
      uint32_t a, b;
      threefry2x32(key[0], key[1], 0, (uint32_t)(base + j), &a, &b);
      out[base + j] = bits_to_normal(a ^ b);
    }
  }
}

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1: the benchmark arm sets the thread count explicitly */
void ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* take_steps.py:60-144: returns 0 on success.  pos [n, n_out, d], lp/acc [n, n_out], last [n, d]; ratio_out / logu_out
 * (optional, [n, n_out]): the two sides of the accept test of every stored step (near-tie accounting in the tests). */
int ref_take_serial_steps(int kind, int tgt, const float* data, const uint32_t* key, const float* x0, int64_t n,
                          int d, int n_steps, int thinning, int64_t chain_offset, float step_size, int n_leapfrog,
                          const float* chol, const float* colsum, uint32_t* key_out, float* pos, float* lp_out,
                          float* acc_out, float* last, float* ratio_out, float* logu_out) {
  if (d > MAXD || d < 1) return -1;
  KernelCfg cfg = {kind, step_size, n_leapfrog, chol, colsum};
  uint32_t sub[2];
  threefry2x32(key[0], key[1], 0, 0, &key_out[0], &key_out[1]);
  threefry2x32(key[0], key[1], 0, 1, &sub[0], &sub[1]);
  const int n_out = (n_steps + thinning - 1) / thinning;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    float x[MAXD], g[MAXD];
    uint32_t kc[2], nk[2], s[2];
    const uint64_t gi = (uint64_t)(chain_offset + c);
    threefry2x32(sub[0], sub[1], (uint32_t)(gi >> 32), (uint32_t)gi, &kc[0], &kc[1]);
    memcpy(x, x0 + c * d, sizeof(float) * d);
    float lp = logp_grad(tgt, data, x, d, g, 0);
    for (int t = 0; t < n_steps; ++t) {
      threefry2x32(kc[0], kc[1], 0, 0, &nk[0], &nk[1]);
      threefry2x32(kc[0], kc[1], 0, 1, &s[0], &s[1]);
      kc[0] = nk[0];
      kc[1] = nk[1];
      float dbg[2];
      int acc = one_step(&cfg, tgt, data, d, s[0], s[1], x, &lp, dbg);
      if (t % thinning == 0) {
        const int o = t / thinning;
        if (ratio_out) ratio_out[(size_t)c * n_out + o] = dbg[0];
        if (logu_out) logu_out[(size_t)c * n_out + o] = dbg[1];
        if (pos) memcpy(pos + ((size_t)c * n_out + o) * d, x, sizeof(float) * d);
        if (lp_out) lp_out[(size_t)c * n_out + o] = lp;
        if (acc_out) acc_out[(size_t)c * n_out + o] = (float)acc;
        if (o == n_out - 1 && last) memcpy(last + c * d, x, sizeof(float) * d);
      }
    }
  }
  return 0;
}
