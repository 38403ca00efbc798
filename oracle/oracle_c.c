/*
 * oracle_c.c — scalar C restatement of the reference ray-marcher.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independently written restatement of inr/viewer/brats_rt.slang (klukaszek/
 * MRI-RayTracer); the first is oracle/oracle_torch.py.  The two are cross-checked by
 * tests/test_oracle_c.py; agreement of two independent restatements plus the analytic
 * known-answer tests is what stands in for the golden images the reference does not have
 * (PARITY UNPINNED for the shader arithmetic — see oracle_torch.py's header).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * One IEEE-754 binary32 operation per source operation: compile with -ffp-contract=off
 * (see oracle/Makefile).  Line citations are into inr/viewer/brats_rt.slang.
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle_c.so
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  int32_t W, H;                 /* imageSize */
  float fovY;
  float eye[3], U[3], V[3], Wv[3];
  float volMin[3], voxelSize[3];
  int32_t dims[3];              /* X, Y, Z */
  float stepSize, nearT, farT;
  float bgColor[3];
  int32_t volEnabled[4];
  float volWeight[4];
  float ww, wl, intensityAlpha, gamma;
  int32_t showSeg, showPred;
  float lut[8][4];
  /* extensions */
  int32_t ortho; float orthoHalfHeight; float ertThreshold; int32_t maxSteps;
  int32_t tMode;                /* 0 indexed, 1 accumulate */
  int32_t alphaMode;
  int32_t useTf;                /* 1: 1D LUT, 0: reference intensity TF */
} OParams;

static float lerp1(float a, float b, float t) { return a + t * (b - a); }
static float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* sampleLinear, :60-76 */
static float sample_linear(const float* buf, const int32_t d[3], const float p[3]) {
  float q[3], f[3];
  uint32_t i[3];
  for (int a = 0; a < 3; ++a) {
    q[a] = clampf(p[a], 0.0f, (float)d[a] - 1.001f);         /* :62 */
    float fl = floorf(q[a]);
    i[a] = (uint32_t)fl;                                      /* :63 */
    f[a] = q[a] - fl;                                         /* :64 */
  }
  const uint32_t sY = (uint32_t)d[0], sZ = (uint32_t)d[0] * (uint32_t)d[1];   /* :66 */
  const uint32_t b = i[0] + i[1] * sY + i[2] * sZ;            /* :67 */
  const float c000 = buf[b], c100 = buf[b + 1];
  const float c010 = buf[b + sY], c110 = buf[b + sY + 1];
  const float c001 = buf[b + sZ], c101 = buf[b + sZ + 1];
  const float c011 = buf[b + sZ + sY], c111 = buf[b + sZ + sY + 1];
  return lerp1(lerp1(lerp1(c000, c100, f[0]), lerp1(c010, c110, f[0]), f[1]),
               lerp1(lerp1(c001, c101, f[0]), lerp1(c011, c111, f[0]), f[1]), f[2]);   /* :74-75 */
}

/* sampleLabel, :78-83; round() = half away from zero */
static int32_t sample_label(const int32_t* buf, const int32_t d[3], const float p[3]) {
  uint32_t i[3];
  for (int a = 0; a < 3; ++a) i[a] = (uint32_t)roundf(clampf(p[a], 0.0f, (float)d[a] - 1.0f));
  return buf[i[0] + i[1] * (uint32_t)d[0] + i[2] * (uint32_t)d[0] * (uint32_t)d[1]];
}

static void normalize3(float v[3]) {
  const float n = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  v[0] /= n; v[1] /= n; v[2] /= n;
}

/* one ray; returns number of slots taken, writes rgba[4], T and the clip count */
static int march(const OParams* P, const float* vol, int C, const float* tf, int tfN, const int32_t* labels,
                 const int32_t* preds, int px, int py, float focal, float rgba[4], float* Tout, int* nclip) {
  const float dimx = (float)P->W, dimy = (float)P->H;
  const float uvx = ((float)px + 0.5f) / dimx * 2.0f - 1.0f;   /* :39-40 */
  const float uvy = ((float)py + 0.5f) / dimy * 2.0f - 1.0f;
  const float aspect = dimx / fmaxf(1.0f, dimy);               /* :42 */
  float o[3], d[3];
  if (P->ortho) {                                              /* SURVEY section 8 row A3 */
    const float halfW = aspect * P->orthoHalfHeight;
    const float ax = uvx * halfW, ay = -(uvy * P->orthoHalfHeight);
    for (int a = 0; a < 3; ++a) { o[a] = (P->eye[a] + ax * P->U[a]) + ay * P->V[a]; d[a] = P->Wv[a]; }
  } else {
    float c[3] = { uvx * aspect / focal, -uvy / focal, 1.0f }; /* :43 */
    normalize3(c);
    for (int a = 0; a < 3; ++a) { d[a] = (c[0] * P->U[a] + c[1] * P->V[a]) + c[2] * P->Wv[a]; o[a] = P->eye[a]; }  /* :44 */
    normalize3(d);
  }
  float bmax[3], t0v[3], t1v[3];
  for (int a = 0; a < 3; ++a) {
    bmax[a] = P->volMin[a] + P->voxelSize[a] * (float)P->dims[a];                   /* :93 */
    const float dz = fabsf(d[a]) < 1e-6f ? 1e-6f : d[a];                            /* :96-98 */
    const float rcp = 1.0f / dz;                                                    /* :99 */
    t0v[a] = (P->volMin[a] - o[a]) * rcp;                                           /* :50 */
    t1v[a] = (bmax[a] - o[a]) * rcp;                                                /* :51 */
  }
  const float tmin = fmaxf(fmaxf(fminf(t0v[0], t1v[0]), fminf(t0v[1], t1v[1])), fminf(t0v[2], t1v[2]));
  const float tmax = fminf(fminf(fmaxf(t0v[0], t1v[0]), fmaxf(t0v[1], t1v[1])), fmaxf(t0v[2], t1v[2]));
  rgba[0] = P->bgColor[0]; rgba[1] = P->bgColor[1]; rgba[2] = P->bgColor[2]; rgba[3] = P->alphaMode ? 0.0f : 1.0f;
  *Tout = 1.0f; *nclip = 0;
  if (!(tmax >= fmaxf(tmin, 0.0f))) return 0;                                       /* :56, :102-105 */
  const float t0 = fmaxf(tmin, fmaxf(0.0f, P->nearT));                              /* :107 */
  const float t1 = fminf(tmax, (P->farT > 0.0f) ? P->farT : tmax);                  /* :108 */
  if (t1 <= t0) return 0;                                                           /* :109 */
  const float dt = P->stepSize;
  int n = 0;
  while (t0 + (float)n * dt < t1) ++n;                       /* indexed count: #{k : t0 + k*dt < t1} */
  if (P->maxSteps > 0 && n > P->maxSteps) n = P->maxSteps;
  *nclip = n;
  const float thr = P->ertThreshold != 0.0f ? P->ertThreshold : 0.01f;
  const float lo = P->wl - P->ww * 0.5f;
  float Cc[3] = { P->bgColor[0], P->bgColor[1], P->bgColor[2] };                    /* :111 */
  float T = 1.0f, t = t0;
  int k = 0;
  for (;;) {
    float tk;
    if (P->tMode == 0) { if (!(k < n && T > thr)) break; tk = t0 + (float)k * dt; }
    else { if (!(t < t1 && T > thr && (P->maxSteps == 0 || k < P->maxSteps))) break; tk = t; }   /* :117 */
    float pI[3];
    for (int a = 0; a < 3; ++a) {
      const float p = o[a] + tk * d[a];                                              /* :119 */
      pI[a] = (p - P->volMin[a]) / P->voxelSize[a];                                  /* :120 */
    }
    float v = 0.0f, wsum = 0.0f;
    const size_t nvox = (size_t)P->dims[0] * P->dims[1] * P->dims[2];
    for (int c = 0; c < C && c < 4; ++c)
      if (P->volEnabled[c]) { v += sample_linear(vol + c * nvox, P->dims, pI) * P->volWeight[c]; wsum += P->volWeight[c]; }  /* :125-128 */
    if (wsum > 0.0f) v /= wsum;                                                      /* :130 */
    float val = clampf((v - lo) / P->ww, 0.0f, 1.0f);                                /* :132 */
    if (P->gamma != 1.0f) val = powf(val, P->gamma);                                 /* :133 */
    if (!P->useTf) {
      if (val > 0.0f) {                                                              /* :135-140 */
        const float a = val * P->intensityAlpha;
        const float alpha = 1.0f - expf(-a * dt);
        const float w = alpha * T * val;
        Cc[0] += w; Cc[1] += w; Cc[2] += w;
        T *= (1.0f - alpha);
      }
    } else {                                                                         /* SURVEY row A7 */
      const float u = val * (float)(tfN - 1);
      const float j0f = floorf(u);
      const float fr = u - j0f;
      int j0 = (int)j0f; if (j0 < 0) j0 = 0; if (j0 > tfN - 1) j0 = tfN - 1;
      const int j1 = j0 + 1 < tfN - 1 ? j0 + 1 : tfN - 1;
      float e[4];
      for (int c = 0; c < 4; ++c) e[c] = lerp1(tf[j0 * 4 + c], tf[j1 * 4 + c], fr);
      const float alpha = 1.0f - expf(-e[3] * dt);
      const float w = alpha * T;
      Cc[0] += w * e[0]; Cc[1] += w * e[1]; Cc[2] += w * e[2];
      T *= (1.0f - alpha);
    }
    for (int pass = 0; pass < 2; ++pass) {                                           /* :143-162 */
      const int32_t* lb = pass == 0 ? (P->showSeg ? labels : 0) : (P->showPred ? preds : 0);
      if (!lb) continue;
      const int32_t l = sample_label(lb, P->dims, pI);
      if (l > 0 && l < 8) {
        const float s = pass == 0 ? P->lut[l][3] * dt : P->lut[l][3] * dt * 1.5f;   /* :147, :158 */
        const float alpha = 1.0f - expf(-s);
        const float w = alpha * T;
        Cc[0] += w * P->lut[l][0]; Cc[1] += w * P->lut[l][1]; Cc[2] += w * P->lut[l][2];
        T *= (1.0f - alpha);
      }
    }
    t += dt;                                                                         /* :164 */
    ++k;
  }
  rgba[0] = Cc[0]; rgba[1] = Cc[1]; rgba[2] = Cc[2];
  rgba[3] = P->alphaMode ? 1.0f - T : 1.0f;                                          /* :167 */
  *Tout = T;
  return k;
}

/* Render pixels listed in (px,py)[npix] (or the whole image when px == NULL).
 * out: float[npix][4]; counts: int32[npix][2] = (n_clip, n_taken) or NULL; Tout: float[npix] or NULL.
 * nthreads > 1 splits the pixel list over POSIX threads in chunks of 256 (dynamic). */
#include <pthread.h>
#include <stdatomic.h>

typedef struct {
  const OParams* P; const float* vol; int C; const float* tf; int tfN; const int32_t* labels; const int32_t* preds;
  const int32_t* px; const int32_t* py; int64_t total; float* out; int32_t* counts; float* Tout; float focal;
  atomic_llong* next;
} Job;

static void* worker(void* arg) {
  Job* j = (Job*)arg;
  for (;;) {
    const long long b = atomic_fetch_add(j->next, 256);
    if (b >= j->total) break;
    const long long e = b + 256 < j->total ? b + 256 : j->total;
    for (long long i = b; i < e; ++i) {
      const int x = j->px ? j->px[i] : (int)(i % j->P->W), y = j->px ? j->py[i] : (int)(i / j->P->W);
      float rgba[4], T; int nclip;
      const int taken = march(j->P, j->vol, j->C, j->tf, j->tfN, j->labels, j->preds, x, y, j->focal, rgba, &T, &nclip);
      memcpy(j->out + 4 * i, rgba, sizeof(rgba));
      if (j->counts) { j->counts[2 * i] = nclip; j->counts[2 * i + 1] = taken; }
      if (j->Tout) j->Tout[i] = T;
    }
  }
  return 0;
}

int oracle_c_render(const OParams* P, const float* vol, int C, const float* tf, int tfN, const int32_t* labels,
                    const int32_t* preds, const int32_t* px, const int32_t* py, int64_t npix,
                    float* out, int32_t* counts, float* Tout, int nthreads) {
  atomic_llong next = 0;
  Job j = { P, vol, C, tf, tfN, labels, preds, px, py, px ? npix : (int64_t)P->W * P->H, out, counts, Tout,
            (float)(1.0 / tan(0.5 * (double)P->fovY)),   /* same documented deviation as oracle_torch */
            &next };
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], 0, worker, &j);
  worker(&j);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
  return 0;
}

int oracle_c_sizeof_params(void) { return (int)sizeof(OParams); }
