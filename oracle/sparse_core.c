/* sparse_core.c -- sequential CPU oracle of the DMPNN sparse core (plain C, one thread).
 *
 * TEST INFRASTRUCTURE ONLY: never linked into or called by the product (dualmessagepassing_b200).
 * Each function states the algorithm of one C-ABI entry point of include/dmp_b200.h as the obvious
 * sequential loop in edge-id order -- which is what the reference executes through DGL's CPU kernels:
 *   seg_reduce     fn.sum over CSC + nloop/bias adds   SubgraphCountingMatching/models/dmpnn.py:92,131-133,163
 *   edge_update    edge_msg + degree term + eloop      dmpnn.py:112,120-123,142-149 ; UNC model.py:256-259
 *   edge_backward  gSpMM backward gather, d(add)       derived from dmpnn.py:113,121,146
 *   stable_segments DGL COO->CSC: stable counting sort
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction => fp32 results are the IEEE
 * sequence of individually rounded operations, the same sequence the CUDA kernels issue).
 * Parity status: checked against torch CPU index_add / numpy stable argsort in tests/test_oracle_sparse.py;
 * the layer-level chain is pinned to the reference classes by tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EID_MASK 0x7fffffffu

void oracle_seg_reduce(const int32_t* indptr, const uint32_t* eid, const float* w_perm, const float* V,
                       int64_t ldV, int64_t rev_off, const float* base, int64_t ld_base, const float* bias,
                       float* out, int64_t ld_out, int64_t nseg, int64_t H, int mode) {
  /* mode & 16 (SPLIT_BY_REV): forward edges accumulate into out[x, 0:H], reversed edges into out[x, H:2H] */
  float* acc = (float*)malloc(sizeof(float) * (size_t)(H > 0 ? 2 * H : 1));
  const int split = (mode & 16) != 0;
  for (int64_t x = 0; x < nseg; ++x) {
    for (int64_t h = 0; h < 2 * H; ++h) acc[h] = 0.0f;
    for (int32_t j = indptr[x]; j < indptr[x + 1]; ++j) {
      uint32_t r = eid[j] >> 31;
      if (((mode & 4) && r) || ((mode & 8) && !r)) continue; /* ONLY_FWD / ONLY_REV */
      const float* row = V + (int64_t)(eid[j] & EID_MASK) * ldV + (r ? rev_off : 0);
      int neg = (mode & 1) && !r;
      float* a = acc + ((split && r) ? H : 0);
      for (int64_t h = 0; h < H; ++h) {
        float t = row[h];
        if (neg) t = -t;
        if (w_perm) t = t * w_perm[j];
        a[h] = a[h] + t;
      }
    }
    for (int64_t h = 0; h < (split ? 2 * H : H); ++h) {
      float r = (mode & 2) ? -acc[h] : acc[h];
      if (base) r = base[x * ld_base + h] + r;
      if (bias) r = r + bias[h];
      out[x * ld_out + h] = r;
    }
  }
  free(acc);
}

void oracle_edge_update(const int32_t* a32, const int32_t* b32, const float* coef, const float* S, int64_t ldS,
                        const float* P, int64_t ldP, const float* Qd, int64_t ldQd, const float* Qs,
                        int64_t ldQs, const float* ebias, float* out, int64_t ld_out, float* agg,
                        int64_t ld_agg, int64_t E, int64_t H, int order) {
  for (int64_t e = 0; e < E; ++e) {
    const float* qd = Qd + (int64_t)a32[e] * ldQd;
    const float* qs = Qs + (int64_t)b32[e] * ldQs;
    for (int64_t h = 0; h < H; ++h) {
      float msg = qd[h] - qs[h];
      float add = coef[e] * P[e * ldP + h];
      float t = order == 0 ? (S[e * ldS + h] + add) + msg : (S[e * ldS + h] + msg) + add;
      if (ebias) t = t + ebias[h];
      if (agg) agg[e * ld_agg + h] = msg;
      out[e * ld_out + h] = t;
    }
  }
}

void oracle_edge_backward(const int32_t* dst32, const uint8_t* rev, const float* norm, const float* coef,
                          const float* gN, const float* gN_rev, int64_t ld_gN, const float* gE, int64_t ld_gE, float* T,
                          int64_t ldT, int64_t T_rev_off, float* CG, int64_t ldCG, int64_t E, int64_t H) {
  for (int64_t e = 0; e < E; ++e) {
    if (T) {
      int neg = !(rev && rev[e]);
      const float* g = ((neg || !gN_rev) ? gN : gN_rev) + (int64_t)dst32[e] * ld_gN;
      for (int64_t h = 0; h < H; ++h) {
        float x = g[h];
        if (norm) x = x * norm[e];
        T[e * ldT + (neg ? 0 : T_rev_off) + h] = neg ? -x : x;
      }
    }
    if (CG)
      for (int64_t h = 0; h < H; ++h) CG[e * ldCG + h] = coef[e] * gE[e * ld_gE + h];
  }
}

/* stable counting sort of edge ids by key; eid_out[j] = id | (flag << 31) */
void oracle_stable_segments(const int32_t* key, const uint8_t* rev, int64_t N, int64_t E, int32_t* indptr,
                            uint32_t* eid_out) {
  int64_t* pos = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
  int64_t* cur = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
  for (int64_t e = 0; e < E; ++e) pos[key[e] + 1]++;
  for (int64_t x = 0; x < N; ++x) pos[x + 1] += pos[x];
  for (int64_t x = 0; x <= N; ++x) {
    indptr[x] = (int32_t)pos[x];
    cur[x] = pos[x];
  }
  for (int64_t e = 0; e < E; ++e) {
    uint32_t f = (rev && rev[e]) ? 0x80000000u : 0u;
    eid_out[cur[key[e]]++] = (uint32_t)e | f;
  }
  free(pos);
  free(cur);
}

static float act_f(float x, int act, float slope) {
  switch (act) {
    case 1: return x > 0.f ? x : 0.f;
    case 2: return x > 0.f ? x : x * slope;
    case 3: return tanhf(x);
    case 4: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
static float act_g(float x, int act, float slope) {
  switch (act) {
    case 1: return x > 0.f ? 1.f : 0.f;
    case 2: return x > 0.f ? 1.f : slope;
    case 3: { float t = tanhf(x); return 1.f - t * t; }
    case 4: { float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }
    default: return 1.f;
  }
}

static float act_g_out(float y, int act, float slope) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : 0.f;
    case 2: return y > 0.f ? 1.f : slope;
    case 3: return 1.f - y * y;
    case 4: return y * (1.f - y);
    default: return 1.f;
  }
}

/* dmpnn.py:236-241,266-275: out = prev + gate * act(x) */
void oracle_gate_residual(const float* x, int64_t ldx, const float* gate, const float* prev, int64_t ld_prev,
                          float* out, int64_t ld_out, int64_t rows, int64_t H, int act, float slope) {
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t h = 0; h < H; ++h) {
      float y = act_f(x[r * ldx + h], act, slope);
      if (gate) y = y * gate[r];
      if (prev) y = prev[r * ld_prev + h] + y;
      out[r * ld_out + h] = y;
    }
}

void oracle_gate_residual_backward(const float* gout, int64_t ld_gout, const float* x, int64_t ldx,
                                   const float* gate, float* gx, int64_t ld_gx, int64_t rows, int64_t H,
                                   int act, float slope) {
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t h = 0; h < H; ++h) {
      float y = gout[r * ld_gout + h];
      if (gate) y = y * gate[r];
      if ((act & ~16) != 0)
        y = y * ((act & 16) ? act_g_out(x[r * ldx + h], act & ~16, slope) : act_g(x[r * ldx + h], act, slope));
      gx[r * ld_gx + h] = y;
    }
}
