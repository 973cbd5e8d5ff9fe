/* CPU oracle, C restatement (TEST INFRASTRUCTURE -- never linked into or called by the product).
 *
 * Plain-C, OpenMP-threaded statement of the reference's QuantLinear arithmetic for the sizes where the
 * numpy oracle (oracle/quip_oracle.py) is too slow, and the host-core baseline that bench.py times
 * ("cpu_baseline" / `--impl reference`; the reference has no CPU implementation of its own ops --
 * register_lib.py registers CUDA-only impls -- so this port IS the CPU statement of that path).
 *
 * Follows:  decode           quip_cuda/origin_order.cu:211-231  (= codebook/e8p12.py:82-103)
 *           weight order     quip_cuda/origin_order.cu:846-856
 *           mm               quip_cuda/origin_order.cu:388-555 (fp16 in, fp32 accumulate, fp16 out)
 *           hadamard         quant.py:50-59 butterfly; fast_hadamard_transform semantics (fp32 internal)
 *           forward chain    qlinear.py:87-115, quant.py:72-88
 * Pinned against oracle/quip_oracle.py (itself pinned on reference-generated goldens) in
 * tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 f16;

int qo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* one code -> 8 packed int8 (quarter units), packed byte order */
static inline uint64_t decode8(uint16_t c, const uint64_t* tab) {
  uint32_t sgn = c & 0xff;
  uint32_t par = __builtin_popcount(sgn) & 1;
  uint64_t s = sgn ^ par;
  uint64_t packed = tab[c >> 8];
  uint64_t d = s * 0x8040201008040201ull;
  d &= 0x8080808080808080ull;
  d >>= 7;
  d *= 252;
  packed ^= d;
  packed |= 0x0101010101010101ull;
  packed -= par * 0x0202020202020202ull;
  return packed;
}

static const int PERM[8] = {0, 2, 1, 3, 4, 6, 5, 7};

/* Qidxs int16 [N, K/8] -> fp16 bits [N, K] */
void qo_decompress_e8p(const uint16_t* q, const uint64_t* tab, uint16_t* out, int64_t N, int64_t K) {
  const int64_t cpr = K / 8;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; n++) {
    for (int64_t j = 0; j < cpr; j++) {
      uint64_t p = decode8(q[n * cpr + j], tab);
      for (int i = 0; i < 8; i++) {
        int8_t b = (int8_t)((p >> (8 * PERM[i])) & 0xff);
        f16 h = (f16)((float)b * 0.25f);
        memcpy(&out[n * K + j * 8 + i], &h, 2);
      }
    }
  }
}

/* y[M,N] (fp32, unrounded) = x[M,K] (fp16 bits) . decode(q)^T, fp32 accumulate; decode-every-call */
void qo_e8p_mm(const uint16_t* xbits, const uint16_t* q, const uint64_t* tab, float* y, int64_t M,
               int64_t N, int64_t K) {
  const int64_t cpr = K / 8;
  float* xf = (float*)malloc(sizeof(float) * M * K);
  for (int64_t i = 0; i < M * K; i++) {
    f16 h;
    memcpy(&h, &xbits[i], 2);
    xf[i] = (float)h;
  }
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; n++) {
    for (int64_t m = 0; m < M; m++) {
      const float* xr = xf + m * K;
      float acc = 0.f;
      for (int64_t j = 0; j < cpr; j++) {
        uint64_t p = decode8(q[n * cpr + j], tab);
        float part = 0.f;
        for (int i = 0; i < 8; i++) {
          int8_t b = (int8_t)((p >> (8 * PERM[i])) & 0xff);
          part += (float)b * xr[j * 8 + i];
        }
        acc += part * 0.25f;
      }
      y[m * N + n] = acc;
    }
  }
  free(xf);
}

/* in-place unnormalised Sylvester FWHT over blocks of length L (power of two), times scale */
void qo_fwht(float* x, int64_t rows, int64_t L, float scale) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; r++) {
    float* v = x + r * L;
    for (int64_t h = 1; h < L; h <<= 1)
      for (int64_t i = 0; i < L; i += 2 * h)
        for (int64_t j = i; j < i + h; j++) {
          float a = v[j], b = v[j + h];
          v[j] = a + b;
          v[j + h] = a - b;
        }
    for (int64_t i = 0; i < L; i++) v[i] *= scale;
  }
}

static inline float r16(float v) { return (float)(f16)v; }

/* (hadK' (x) H_L) over one row of length q = K*L; hk is [K][K] fp32 (already transposed if wanted);
 * rounding to fp16 after the FWHT and after the mix, as the reference's fp16 tensors do. */
static void rotate_row(float* v, float* tmp, int64_t q, int64_t K, const float* hk, float scale) {
  const int64_t L = q / K;
  qo_fwht(v, K, L, scale);
  for (int64_t i = 0; i < q; i++) v[i] = r16(v[i]);
  if (K == 1) return;
  for (int64_t k = 0; k < K; k++)
    for (int64_t c = 0; c < L; c++) {
      float acc = 0.f;
      for (int64_t kp = 0; kp < K; kp++) acc += hk[k * K + kp] * v[kp * L + c];
      tmp[k * L + c] = r16(acc);
    }
  memcpy(v, tmp, sizeof(float) * q);
}

/* Full eval-mode QuantLinear.forward for E8P12 with the reference's fp16 rounding points.
 * x: fp16 bits [M, in]; y: fp32 [M, out] (fp16-representable values).  SU/SV/bias: fp32 or NULL.
 * had_left / had_right: fp32 [K,K] row-major or NULL.  Returns 0. */
int qo_quantlinear_forward_e8p(const uint16_t* xbits, float* y, int64_t M, int64_t in_f, int64_t out_f,
                               int64_t q_in, int64_t q_out, const uint16_t* qidxs, const uint64_t* tab,
                               const float* SU, const float* SV, const float* bias, float wscale,
                               const float* had_left, int64_t K_left, const float* had_right,
                               int64_t K_right) {
  float* xr = (float*)calloc((size_t)M * q_in, sizeof(float));
  float* tmp = (float*)malloc(sizeof(float) * (q_in > q_out ? q_in : q_out));
  float* hlT = NULL;
  if (K_left > 1) {
    hlT = (float*)malloc(sizeof(float) * K_left * K_left);
    for (int64_t a = 0; a < K_left; a++)
      for (int64_t b = 0; b < K_left; b++) hlT[a * K_left + b] = had_left[b * K_left + a];
  }
  uint16_t* xh = (uint16_t*)malloc(sizeof(uint16_t) * M * q_in);
  for (int64_t m = 0; m < M; m++) {
    float* v = xr + m * q_in;
    for (int64_t i = 0; i < in_f; i++) {
      f16 h;
      memcpy(&h, &xbits[m * in_f + i], 2);
      v[i] = SU ? r16((float)h * SU[i]) : (float)h;
    }
    rotate_row(v, tmp, q_in, K_left, hlT, wscale / sqrtf((float)(q_in / K_left)));
    for (int64_t i = 0; i < q_in; i++) {
      f16 h = (f16)v[i];
      memcpy(&xh[m * q_in + i], &h, 2);
    }
  }
  float* o = (float*)malloc(sizeof(float) * M * q_out);
  qo_e8p_mm(xh, qidxs, tab, o, M, q_out, q_in);
  for (int64_t m = 0; m < M; m++) {
    float* v = o + m * q_out;
    for (int64_t i = 0; i < q_out; i++) v[i] = r16(v[i]);
    rotate_row(v, tmp, q_out, K_right, had_right, 1.0f / sqrtf((float)(q_out / K_right)));
    for (int64_t i = 0; i < out_f; i++) {
      float t = v[i];
      if (SV) t = r16(t * SV[i]);
      if (bias) t = r16(t + bias[i]);
      y[m * out_f + i] = t;
    }
  }
  free(o); free(xh); free(hlT); free(tmp); free(xr);
  return 0;
}
