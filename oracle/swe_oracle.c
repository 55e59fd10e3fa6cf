/* TEST INFRASTRUCTURE ONLY -- see swe_oracle.h. Plain-C restatement of the reference
 * hot path; build with -O2 -ffp-contract=off (no FMA contraction, IEEE fp32 everywhere,
 * like the reference Makefile:3-4). Every function cites the reference lines it follows.
 */
#include "swe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * glibc rand()/srand(): stdlib/random_r.c, TYPE_3 (x^31 + x^3 + 1), deg 31, sep 3.
 * Part of the reference semantics: world.hpp:52,54 and wind.hpp:83 draw from it and
 * SimpleWindErosion.cpp:29,33 seeds it.
 * ---------------------------------------------------------------------------------- */
void swo_srand(swo_rand_state* s, unsigned seed) {
  int32_t word = (int32_t)(seed ? seed : 1u);
  s->r[0] = word;
  for (int i = 1; i < 31; i++) {
    long hi = word / 127773, lo = word % 127773;
    word = (int32_t)(16807 * lo - 2836 * hi);
    if (word < 0) word += 2147483647;
    s->r[i] = word;
  }
  s->f = 3; s->b = 0;
  for (int i = 0; i < 310; i++) (void)swo_rand(s);
}

int swo_rand(swo_rand_state* s) {
  uint32_t v = (uint32_t)s->r[s->f] + (uint32_t)s->r[s->b];
  s->r[s->f] = (int32_t)v;
  if (++s->f >= 31) s->f = 0;
  if (++s->b >= 31) s->b = 0;
  return (int)(v >> 1);
}

/* ------------------------------------------------------------------------------------
 * glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c + e_exp2f_data.c, N = 32 table,
 * degree-3 polynomial in double). wind.hpp:50 calls exp(float) -> std::exp(float) ->
 * expf. Verified EXHAUSTIVELY (all 2^32 inputs) against this container's libm: identical
 * except two inputs, patched below so the restatement equals glibc everywhere.
 * The device kernel restates the same algorithm, so oracle and GPU share one expf.
 * ---------------------------------------------------------------------------------- */
static const uint64_t swo_exp2f_tab[32] = {
  0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
  0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
  0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
  0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
  0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
  0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
  0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
  0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float swo_expf(float x) {
  const double InvLn2N = 0x1.71547652b82fep+0 * 32;
  const double Shift = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32;
  const double C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32;
  const double C2 = 0x1.62e42ff0c52d6p-1 / 32;
  uint32_t ux = f2u(x), abstop = (ux >> 20) & 0x7ff;
  if (abstop >= (0x42b00000u >> 20)) { /* |x| >= 88 or nan */
    if (ux == 0xff800000u) return 0.0f;
    if (abstop >= (0x7f800000u >> 20)) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;   /* overflow */
    if (x < -0x1.9fe368p6f) return 0.0f;      /* underflow */
  }
  /* the two inputs (of 2^32) where this container's libm rounds the other way */
  if (ux == 0x4202422fu) return u2f(0x56fc9f1cu);  /* x =  0x1.04845ep+5 -> 0x1.f93e38p+46 */
  if (ux == 0xc27c65d9u) return u2f(0x11fa2993u);  /* x = -0x1.f8cbb2p+5 -> 0x1.f45326p-92 */
  double z = InvLn2N * (double)x;
  double kd = z + Shift;
  uint64_t ki; memcpy(&ki, &kd, 8);
  kd -= Shift;
  double r = z - kd;
  uint64_t t = swo_exp2f_tab[ki % 32] + (ki << (52 - 5));
  double s; memcpy(&s, &t, 8);
  double zz = C0 * r + C1;
  double r2 = r * r;
  double y = C2 * r + 1.0;
  y = zz * r2 + y;
  y = y * s;
  return (float)y;
}

/* ------------------------------------------------------------------------------------ */
void swo_default_params(swo_params* p) {
  p->max_age = 512; p->boundarylayer = 2.0f; p->suspension = 0.05f;
  p->gravity = (float)0.1;      /* wind.hpp:26 `float Wind::gravity = 0.1;`   */
  p->lrate = 0.1f;
  p->maxdiff = (float)0.005;    /* world.hpp:31 */
  p->settling = (float)1.0;     /* world.hpp:32 */
  p->mapscale = 80;
  /* wind.hpp:13: pspeed = glm::normalize(vec3(1,0,0)) = v * (1/sqrt(dot(v,v))) */
  float inv = 1.0f / sqrtf(1.0f * 1.0f + 0.0f * 0.0f + 0.0f * 0.0f);
  p->pspeed[0] = 1.0f * inv; p->pspeed[1] = 0.0f * inv; p->pspeed[2] = 0.0f * inv;
}

swo_world* swo_create(int nx, int nz) {
  swo_world* w = (swo_world*)calloc(1, sizeof(swo_world));
  if (!w) return NULL;
  w->nx = nx; w->nz = nz;
  for (int i = 0; i < 9; i++) {
    w->plane[i] = (float*)calloc((size_t)nx * nz, sizeof(float));
    if (!w->plane[i]) { swo_destroy(w); return NULL; }
  }
  swo_default_params(&w->p);
  swo_srand(&w->rng, 1);
  return w;
}

void swo_destroy(swo_world* w) {
  if (!w) return;
  for (int i = 0; i < 9; i++) free(w->plane[i]);
  free(w->rock);
  free(w);
}

void swo_set_rock(swo_world* w, const float* rock) {
  free(w->rock);
  w->rock = NULL;
  if (!rock) return;
  const size_t n = (size_t)w->nx * w->nz;
  w->rock = (float*)malloc(n * sizeof(float));
  memcpy(w->rock, rock, n * sizeof(float));
}

/* cellpool.h:372-428 (mapsize = 1: node.pos = 0, tileres = (n,n), lodsize = 1) */
void swo_hill_raw(swo_world* w) {
  const int n = w->nx;
  float* H = w->plane[0];
  for (int i = 1; i < 9; i++) memset(w->plane[i], 0, sizeof(float) * (size_t)w->nx * w->nz);
  /* :381-384  p = vec2(pos)/vec2(tileres); c = vec2(tileres/ivec2(4,2))/vec2(tileres) */
  const float cx = (float)(n / 4) / (float)n, cz = (float)(n / 2) / (float)n;
  for (int x = 0; x < w->nx; x++)
    for (int z = 0; z < w->nz; z++) {
      float px = (float)x / (float)n, pz = (float)z / (float)n;
      float dx = px - cx, dz = pz - cz;
      float d = sqrtf(dx * dx + dz * dz);                  /* glm::length(vec2) */
      /* exp(-d*d*quad::tilesize*0.2): float*float, *int -> float, *0.2 -> double exp */
      H[(size_t)x * w->nz + z] = (float)exp((double)(-d * d * (float)n) * 0.2);
    }
}

void swo_init_gaussian(swo_world* w) {
  swo_hill_raw(w);
  swo_normalize(w);   /* :416-428 */
}

/* ---- glm scalar semantics (see header) ---- */
typedef struct { float x, y, z; } v3;
static inline v3 v3_(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_scale(float s, v3 a) { return v3_(s * a.x, s * a.y, s * a.z); }
static inline v3 v3_scale_r(v3 a, float s) { return v3_(a.x * s, a.y * s, a.z * s); }
static inline float v3_dot(v3 a, v3 b) { v3 t = v3_mul(a, b); return t.x + t.y + t.z; }
static inline float v3_length(v3 a) { return sqrtf(v3_dot(a, a)); }
static inline v3 v3_normalize(v3 a) { return v3_scale_r(a, 1.0f / sqrtf(v3_dot(a, a))); }
static inline v3 v3_cross(v3 x, v3 y) {
  return v3_(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
static inline v3 v3_mix(v3 x, v3 y, float a) {
  return v3_add(v3_scale_r(x, 1.0f - a), v3_scale_r(y, a));
}

static inline int oob(const swo_world* w, int x, int z) { /* cellpool.h:432-438; periodic variant :451-453 */
  if (w->periodic) return 0;
  return x < 0 || z < 0 || x >= w->nx || z >= w->nz;
}
/* cell index; the periodic variant wraps (cellpool.h:102-110: while(p.x < 0) p.x += res.x; ...) */
static inline size_t cidx(const swo_world* w, int x, int z) {
  if (w->periodic) {
    while (x < 0) x += w->nx;
    while (z < 0) z += w->nz;
    while (x >= w->nx) x -= w->nx;
    while (z >= w->nz) z -= w->nz;
  }
  return (size_t)x * w->nz + z;
}
#define HGT(w, x, z) ((w)->plane[0][cidx((w), (x), (z))])

/* quad::_normal, cellpool.h:200-223: up to four triangle normals, in the reference's
 * order; a triangle counts iff its DIAGONAL neighbour is in bounds. */
static v3 normal_v3(const swo_world* w, int x, int z) {
  v3 n = v3_(0.0f, 0.0f, 0.0f);
  const v3 s = v3_(1.0f, (float)w->p.mapscale, 1.0f);
  const float h = HGT(w, x, z);
  if (!oob(w, x + 1, z + 1))
    n = v3_add(n, v3_cross(v3_mul(s, v3_(0.0f, HGT(w, x, z + 1) - h, 1.0f)),
                           v3_mul(s, v3_(1.0f, HGT(w, x + 1, z) - h, 0.0f))));
  if (!oob(w, x - 1, z - 1))
    n = v3_add(n, v3_cross(v3_mul(s, v3_(0.0f, HGT(w, x, z - 1) - h, -1.0f)),
                           v3_mul(s, v3_(-1.0f, HGT(w, x - 1, z) - h, 0.0f))));
  if (!oob(w, x + 1, z - 1))
    n = v3_add(n, v3_cross(v3_mul(s, v3_(1.0f, HGT(w, x + 1, z) - h, 0.0f)),
                           v3_mul(s, v3_(0.0f, HGT(w, x, z - 1) - h, -1.0f))));
  if (!oob(w, x - 1, z + 1))
    n = v3_add(n, v3_cross(v3_mul(s, v3_(-1.0f, HGT(w, x - 1, z) - h, 0.0f)),
                           v3_mul(s, v3_(0.0f, HGT(w, x, z + 1) - h, 1.0f))));
  if (v3_length(n) > 0) n = v3_normalize(n);
  return n;
}

void swo_normal(const swo_world* w, int x, int z, float out[3]) {
  v3 n = normal_v3(w, x, z);
  out[0] = n.x; out[1] = n.y; out[2] = n.z;
}

/* World::cascade, world.hpp:79-153. */
void swo_cascade(swo_world* w, int x, int z) {
  static const int off[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1},
                                {0, 1},   {1, -1}, {1, 0},  {1, 1}};   /* :83-92 */
  struct { int x, z; float h, d; } sn[8], val;
  int num = 0;
  for (int k = 0; k < 8; k++) {                                        /* :105-114 */
    int nx_ = x + off[k][0], nz_ = z + off[k][1];
    if (oob(w, nx_, nz_)) continue;
    float fx = (float)off[k][0], fz = (float)off[k][1];
    sn[num].x = nx_; sn[num].z = nz_;
    sn[num].h = HGT(w, nx_, nz_);
    sn[num].d = sqrtf(fx * fx + fz * fz);                              /* length(vec2(nn)) */
    num++;
  }
  /* :118-120 std::sort, n <= 8 < 16 -> libstdc++ __insertion_sort, comp = a.h > b.h */
  for (int i = 1; i < num; i++) {
    val = sn[i];
    if (val.h > sn[0].h) {
      memmove(&sn[1], &sn[0], sizeof(sn[0]) * (size_t)i);
      sn[0] = val;
    } else {
      int j = i;
      while (val.h > sn[j - 1].h) { sn[j] = sn[j - 1]; j--; }
      sn[j] = val;
    }
  }
  for (int i = 0; i < num; i++) {                                      /* :122-151 */
    float diff = HGT(w, x, z) - sn[i].h;
    if (diff == 0) continue;
    float excess = fabsf(diff) - sn[i].d * w->p.maxdiff * (float)1 /* lodsize */;
    if (excess <= 0) continue;
    float transfer = w->p.settling * excess / 2.0f;
    if (w->rock) { /* layered extension: only sand moves */
      const float sand = (diff > 0) ? HGT(w, x, z) - w->rock[(size_t)x * w->nz + z]
                                    : sn[i].h - w->rock[(size_t)sn[i].x * w->nz + sn[i].z];
      const float cap = sand > 0.0f ? sand : 0.0f;
      if (cap < transfer) transfer = cap;
      if (!(transfer > 0.0f)) continue;
    }
    if (diff > 0) {
      HGT(w, x, z) -= transfer;
      HGT(w, sn[i].x, sn[i].z) += transfer;
    } else {
      HGT(w, x, z) += transfer;
      HGT(w, sn[i].x, sn[i].z) -= transfer;
    }
  }
}

void swo_particle_at(swo_particle* q, float x, float z) { /* wind.hpp:6-12 */
  q->age = 0; q->sediment = 0.0f;
  q->pos[0] = x; q->pos[1] = 0.0f; q->pos[2] = z;
  q->speed[0] = q->speed[1] = q->speed[2] = 0.0f;
}

/* world.hpp:52-54. g++ evaluates constructor arguments right to left, so the LAST
 * component consumes the FIRST draw (probe-verified against oracle/_ref). */
void swo_spawn(swo_world* w, swo_particle* q) {
  int iz = swo_rand(&w->rng) % w->nz;
  int ix = swo_rand(&w->rng) % w->nx;
  int fz = swo_rand(&w->rng) % 1000;
  int fx = swo_rand(&w->rng) % 1000;
  float x = (float)ix, z = (float)iz;
  x += (float)fx / 1000.0f;
  z += (float)fz / 1000.0f;
  swo_particle_at(q, x, z);
}

/* Wind::fly, wind.hpp:28-117. */
int swo_fly(swo_world* w, swo_particle* q) {
  const swo_params* P = &w->p;
  const int ix = (int)roundf(q->pos[0]), iz = (int)roundf(q->pos[2]);   /* :30 */
  if (oob(w, ix, iz)) return 0;                                          /* :32-38 */
  const size_t ind = cidx(w, ix, iz);
  const float h = w->plane[0][ind];
  const v3 n = normal_v3(w, ix, iz);                                     /* :40 */
  v3 pos = v3_(q->pos[0], q->pos[1], q->pos[2]);
  v3 speed = v3_(q->speed[0], q->speed[1], q->speed[2]);
  const v3 pspeed = v3_(P->pspeed[0], P->pspeed[1], P->pspeed[2]);

  if (q->age == 0 || pos.y < h) pos.y = h;                               /* :42-43 */
  q->pos[1] = pos.y;
  if (q->age++ > P->max_age) return 0;                                   /* :45-46 */

  float hfac = swo_expf(-(pos.y - h) / P->boundarylayer);               /* :50 */
  if (hfac < 0) hfac = 0;

  float shadow = v3_dot(v3_normalize(pspeed), n);                        /* :56-59 */
  if (shadow < 0) shadow = 0;
  shadow = 1.0f - shadow;

  speed = v3_add(speed, v3_scale(0.05f, v3_sub(v3_scale(0.1f + 0.9f * shadow, pspeed), speed))); /* :61 */

  if (pos.y > h) speed.y -= P->gravity * q->sediment;                    /* :65-66 */

  float collision = -v3_dot(v3_normalize(speed), n);                     /* :70-71 */
  if (collision < 0) collision = 0;

  v3 rspeed = v3_cross(n, v3_cross(v3_scale(1.0f - collision, speed), n)); /* :75 */

  speed = v3_add(speed, v3_scale(0.9f, v3_sub(v3_scale(shadow, v3_mix(pspeed, rspeed, shadow * hfac)), speed))); /* :79 */

  {                                                                      /* :83 */
    int rz = swo_rand(&w->rng) % 1001;   /* last ctor argument is evaluated first */
    int ry = swo_rand(&w->rng) % 1001;
    int rx = swo_rand(&w->rng) % 1001;
    v3 t = v3_((float)rx - 500.0f, (float)ry - 500.0f, (float)rz - 500.0f);
    t = v3_scale(0.1f * hfac * collision, t);
    t = v3_(t.x / 500.0f, t.y / 500.0f, t.z / 500.0f);
    speed = v3_add(speed, t);
  }

  {                                                                      /* :87 double factor */
    float f = (float)(1.0 - 0.3 * (double)q->sediment);
    speed = v3_(speed.x * f, speed.y * f, speed.z * f);
  }

  pos = v3_add(pos, speed);                                              /* :91 */

  w->plane[6][ind] += speed.x;                                           /* :93-96 */
  w->plane[7][ind] += speed.y;
  w->plane[8][ind] += speed.z;
  w->plane[5][ind] += q->sediment;

  float force = -v3_dot(v3_normalize(speed), n) * v3_length(speed);      /* :100-102 */
  if (force < 0) force = 0;
  float lift = (1.0f - collision) * v3_length(speed);                    /* :104 */
  float capacity = force * hfac + 0.02f * lift * hfac;                   /* :106 */

  float diff = capacity - q->sediment;                                   /* :110-112 */
  w->plane[0][ind] -= P->suspension * diff;
  q->sediment += P->suspension * diff;

  q->pos[0] = pos.x; q->pos[1] = pos.y; q->pos[2] = pos.z;
  q->speed[0] = speed.x; q->speed[1] = speed.y; q->speed[2] = speed.z;

  swo_cascade(w, ix, iz);                                                /* :114 */
  return 1;
}

/* World::erode, world.hpp:36-77 (single tile: mapsize = 1). */
void swo_erode(swo_world* w, int cycles, swo_stats* st) {
  const size_t area = (size_t)w->nx * w->nz;
  for (int k = 5; k < 9; k++) memset(w->plane[k], 0, sizeof(float) * area);     /* :38-44 */
  swo_stats s = {0, 0, 0, 0.0};
  for (int i = 0; i < cycles; i++) {                                              /* :47-66 */
    swo_particle q;
    swo_spawn(w, &q);
    while (swo_fly(w, &q)) s.steps++;
    if (q.age == w->p.max_age + 2) s.exits_age++; else s.exits_oob++;
    s.sediment_lost += (double)q.sediment;
  }
  const float lr = w->p.lrate;                                                    /* :70-76 */
  for (int k = 1; k < 5; k++) {
    float* f = w->plane[k]; const float* t = w->plane[k + 4];
    for (size_t i = 0; i < area; i++) f[i] = (1.0f - lr) * f[i] + lr * t[i];
  }
  if (st) *st = s;
}

void swo_cascade_sweep(swo_world* w, int sweeps) {
  for (int s = 0; s < sweeps; s++)
    for (int c = 0; c < 9; c++) {
      int cx = c / 3, cz = c % 3;
      for (int x = cx; x < w->nx; x += 3)
        for (int z = cz; z < w->nz; z += 3) swo_cascade(w, x, z);
    }
}

float swo_max_excess(const swo_world* w) {
  static const int off[4][2] = {{0, 1}, {1, -1}, {1, 0}, {1, 1}};
  float m = -INFINITY;
  for (int x = 0; x < w->nx; x++)
    for (int z = 0; z < w->nz; z++)
      for (int k = 0; k < 4; k++) {
        int x2 = x + off[k][0], z2 = z + off[k][1];
        if (oob(w, x2, z2)) continue;
        float fx = (float)off[k][0], fz = (float)off[k][1];
        float e = fabsf(HGT(w, x, z) - HGT(w, x2, z2)) - sqrtf(fx * fx + fz * fz) * w->p.maxdiff;
        if (e > m) m = e;
      }
  return m;
}

int swo_trace(swo_world* w, float x, float z, int max_records, float* out) {
  swo_particle q;
  swo_particle_at(&q, x, z);
  int n = 0;
  while (swo_fly(w, &q)) {
    if (n < max_records) {
      float* r = out + (size_t)n * 8;
      r[0] = q.pos[0]; r[1] = q.pos[1]; r[2] = q.pos[2];
      r[3] = q.speed[0]; r[4] = q.speed[1]; r[5] = q.speed[2];
      r[6] = q.sediment; r[7] = (float)q.age;
    }
    n++;
  }
  return n;
}

uint64_t swo_hash_height(const swo_world* w) {
  uint64_t h = 1469598103934665603ull;
  const float* H = w->plane[0];
  for (size_t i = 0; i < (size_t)w->nx * w->nz; i++) {
    h ^= f2u(H[i]); h *= 1099511628211ull;
  }
  return h;
}

/* ------------------------------------------------------------------------------------
 * Host generate, noise stack: quad::map::init's eight layers of
 * FastNoiseLite OpenSimplex2 FBm, cellpool.h:388-411 (commented out in the snapshot, so the
 * layer loop is restated from the comment; the noise itself is source/include/FastNoiseLite.h).
 * FastNoiseLite object state as map::init leaves it: default ctor (FastNoiseLite.h:114-129:
 * seed 1337, 3 octaves, lacunarity 2, gain 0.5, weighted strength 0, fractal bounding
 * 1/1.75f), SetNoiseType(OpenSimplex2), SetFractalType(FBm) (cellpool.h:366-368).
 * ---------------------------------------------------------------------------------- */
/* Lookup<float>::Gradients3D (FastNoiseLite.h:2529-2547) has a closed form: entries 0..59 are
 * five repeats of the twelve cube-edge directions (family = which coordinate is zero, two sign
 * bits), entries 60..63 repeat directions 8, 1, 9, 3. */
static void swo_fnl_gradient(int g, float out[3]) {
  static const int tail[4] = {8, 1, 9, 3};
  const int e = (g >= 60) ? tail[g - 60] : g % 12;
  const int family = e / 4;
  const float s0 = (e & 1) ? -1.0f : 1.0f, s1 = (e & 2) ? -1.0f : 1.0f;
  if (family == 0) { out[0] = 0.0f; out[1] = s0; out[2] = s1; }
  else if (family == 1) { out[0] = s0; out[1] = 0.0f; out[2] = s1; }
  else { out[0] = s0; out[1] = s1; out[2] = 0.0f; }
}

/* GradCoord 3D (FastNoiseLite.h:542-553) on top of Hash (:500-506); int arithmetic wraps. */
static float swo_fnl_gradcoord(int32_t seed, int32_t xp, int32_t yp, int32_t zp, float xd, float yd, float zd) {
  int32_t h = (int32_t)((uint32_t)(seed ^ xp ^ yp ^ zp) * 0x27d4eb2du);
  h ^= h >> 15;                           /* arithmetic shift, as on the reference's int */
  float g[3];
  swo_fnl_gradient((h >> 2) & 63, g);
  return xd * g[0] + yd * g[1] + zd * g[2];
}

#define FNL_PX 501125321
#define FNL_PY 1136930381
#define FNL_PZ 1720413743
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t fnl_round(float f) { return f >= 0 ? (int32_t)(f + 0.5f) : (int32_t)(f - 0.5f); }

/* SingleOpenSimplex2 3D (FastNoiseLite.h:1054-1147): two offset cubic lattices; on each, the
 * nearest lattice point and its neighbour along the axis of largest offset contribute
 * falloff^4 * gradient. */
float swo_fnl_single3(int seed, float x, float y, float z) {
  int32_t li = fnl_round(x), lj = fnl_round(y), lk = fnl_round(z);
  float dx = x - (float)li, dy = y - (float)lj, dz = z - (float)lk;
  int32_t sx = (int32_t)(-1.0f - dx) | 1, sy = (int32_t)(-1.0f - dy) | 1, sz = (int32_t)(-1.0f - dz) | 1;
  float ax = (float)sx * -dx, ay = (float)sy * -dy, az = (float)sz * -dz;
  li = wrap_mul(li, FNL_PX); lj = wrap_mul(lj, FNL_PY); lk = wrap_mul(lk, FNL_PZ);
  float value = 0;
  float fall = (0.6f - dx * dx) - (dy * dy + dz * dz);
  int32_t sd = seed;
  for (int lattice = 0; lattice < 2; lattice++) {
    if (fall > 0) value += (fall * fall) * (fall * fall) * swo_fnl_gradcoord(sd, li, lj, lk, dx, dy, dz);
    /* second vertex: step along the dominant axis (ties: x before y before z, :1100,1106) */
    float fall2 = fall + 1;
    int32_t i2 = li, j2 = lj, k2 = lk;
    float ex = dx, ey = dy, ez = dz;
    if (ax >= ay && ax >= az)      { ex += (float)sx; fall2 -= (float)(sx * 2) * ex; i2 = wrap_add(i2, -wrap_mul(sx, FNL_PX)); }
    else if (ay > ax && ay >= az)  { ey += (float)sy; fall2 -= (float)(sy * 2) * ey; j2 = wrap_add(j2, -wrap_mul(sy, FNL_PY)); }
    else                           { ez += (float)sz; fall2 -= (float)(sz * 2) * ez; k2 = wrap_add(k2, -wrap_mul(sz, FNL_PZ)); }
    if (fall2 > 0) value += (fall2 * fall2) * (fall2 * fall2) * swo_fnl_gradcoord(sd, i2, j2, k2, ex, ey, ez);
    if (lattice == 1) break;
    /* move to the second lattice (:1126-1143) */
    ax = 0.5f - ax; ay = 0.5f - ay; az = 0.5f - az;
    dx = (float)sx * ax; dy = (float)sy * ay; dz = (float)sz * az;
    fall += (0.75f - ax) - (ay + az);
    li = wrap_add(li, (sx >> 1) & FNL_PX); lj = wrap_add(lj, (sy >> 1) & FNL_PY); lk = wrap_add(lk, (sz >> 1) & FNL_PZ);
    sx = -sx; sy = -sy; sz = -sz;
    sd = ~sd;
  }
  return value * 32.69428253173828125f;
}

/* GetNoise(x,y,z) (FastNoiseLite.h:322-340) in the state described above:
 * TransformNoiseCoordinate (:687-721, DefaultOpenSimplex2 rotation) then GenFractalFBm (:866-887). */
float swo_fnl_getnoise(float frequency, float x, float y, float z) {
  x *= frequency; y *= frequency; z *= frequency;
  const float R3 = (float)(2.0 / 3.0);
  const float r = (x + y + z) * R3;
  x = r - x; y = r - y; z = r - z;
  int seed = 1337;
  float sum = 0;
  float amp = 1 / 1.75f;
  for (int o = 0; o < 3; o++) {
    const float noise = swo_fnl_single3(seed++, x, y, z);
    sum += noise * amp;
    amp *= 1.0f + 0.0f * ((noise + 1) * 0.5f - 1.0f);   /* Lerp(1, (noise+1)/2, weightedStrength = 0) */
    x *= 2.0f; y *= 2.0f; z *= 2.0f;
    amp *= 0.5f;
  }
  return sum;
}

/* The layer loop of cellpool.h:392-409: frequency 1,2,4,...; scale 0.6, *= 0.6 (double product
 * rounded to float); p = cell / tile; third coordinate (float)(SEED % 10000). Adds to height. */
void swo_noise_rows(float* H, int x_lo, int x_hi, int nz, int seed, int layers, int tile) {
  /* H = rows [x_lo, x_hi) of the map (a band, so that callers can split the work) */
  float frequency = 1.0f, scale = 0.6f;
  const float zc = (float)(seed % 10000);
  for (int o = 0; o < layers; o++) {
    for (int x = x_lo; x < x_hi; x++)
      for (int z = 0; z < nz; z++) {
        const float px = (float)x / (float)tile, pz = (float)z / (float)tile;
        H[(size_t)(x - x_lo) * nz + z] += scale * swo_fnl_getnoise(frequency, px, pz, zc);
      }
    frequency *= 2;
    scale = (float)((double)scale * 0.6);
  }
}

void swo_noise_stack(swo_world* w, int seed, int layers, int tile) {
  swo_noise_rows(w->plane[0], 0, w->nx, w->nz, seed, layers, tile);
}

/* cellpool.h:416-428: min and max start at 0.0f; height = 0.5*((h-min)/(max-min)). */
void swo_normalize(swo_world* w) {
  float* H = w->plane[0];
  float mn = 0.0f, mx = 0.0f;
  const size_t n = (size_t)w->nx * w->nz;
  for (size_t i = 0; i < n; i++) { mn = (mn < H[i]) ? mn : H[i]; mx = (mx > H[i]) ? mx : H[i]; }
  for (size_t i = 0; i < n; i++) H[i] = (float)(0.5 * (double)((H[i] - mn) / (mx - mn)));
}

/* quad::updatenode, cellpool.h:302-321: the 12-float Vertex (vertexpool.hpp:6-26) of every cell --
 * P = (x, mapscale*h, z), normal, T - P, B - P with T/B the surface at (x+1,z)/(x,z+1); a
 * neighbour outside the map has height 0 (node::height, cellpool.h:257-261). */
void swo_vertices(const swo_world* w, float* out) {
  const float ms = (float)w->p.mapscale;
  for (int x = 0; x < w->nx; x++)
    for (int z = 0; z < w->nz; z++) {
      float* v = out + ((size_t)x * w->nz + z) * 12;
      const float px = (float)x, pz = (float)z;
      const float py = ms * HGT(w, x, z);
      const float ty = ms * (oob(w, x + 1, z) ? 0.0f : HGT(w, x + 1, z));
      const float by = ms * (oob(w, x, z + 1) ? 0.0f : HGT(w, x, z + 1));
      v3 n = normal_v3(w, x, z);
      v[0] = px; v[1] = py; v[2] = pz;
      v[3] = n.x; v[4] = n.y; v[5] = n.z;
      v[6] = (float)(x + 1) - px; v[7] = ty - py; v[8] = pz - pz;
      v[9] = px - px; v[10] = by - py; v[11] = (float)(z + 1) - pz;
    }
}
