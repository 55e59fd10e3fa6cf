/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference hot path.
 *
 * This is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product
 * (simplewinderosion_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * What it restates (all citations relative to /root/reference):
 *   source/wind.hpp:28-117      Wind::fly          -> swo_fly
 *   source/world.hpp:36-77      World::erode       -> swo_erode
 *   source/world.hpp:79-153     World::cascade     -> swo_cascade
 *   source/cellpool.h:200-223   quad::_normal      -> swo_normal
 *   source/cellpool.h:432-476   map::oob/get/height-> inlined bounds tests, index x*nz+z
 *   source/cellpool.h:339-428   map::init (Gaussian hill + normalise) -> swo_init_gaussian, swo_hill_raw, swo_normalize
 *   source/cellpool.h:366-368,388-411 + source/include/FastNoiseLite.h:322-340,500-553,687-721,866-887,1054-1147
 *                               map::init's noise layers (a comment in the snapshot) over FastNoiseLite's
 *                               OpenSimplex2 FBm -> swo_fnl_single3, swo_fnl_getnoise, swo_noise_stack, swo_noise_rows
 *   source/cellpool.h:302-321   quad::updatenode (Vertex records)     -> swo_vertices
 *   glibc 2.39 rand()/srand() (TYPE_3 additive feedback) -> swo_srand / swo_rand
 *   glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c)    -> swo_expf
 *   glm 0.9.9 scalar vec ops (dot/cross/normalize/mix/length/round) -> inlined
 *
 * PARITY PIN: the reference ships no tests or golden vectors. This
 * restatement is pinned against the reference itself compiled headless here
 * (oracle/_ref, built by oracle/Makefile from the unmodified headers) -- bit-exact on
 * heights, trajectories, step counts -- and against fixtures generated from it and
 * committed under tests/golden/ (tests/golden/make_golden.py, make_noise_golden.py,
 * make_vertices_golden.py: the noise against the unmodified FastNoiseLite class, the vertex
 * records against the reference's own updatenode). The glm semantics are an
 * assumption (glm is absent from the container); they are cross-checked by the survey's
 * independent probe fingerprints (tests/test_oracle_ref.py).
 */
#ifndef SWE_ORACLE_H
#define SWE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swo_params {
  int   max_age;        /* Wind::maxAge        wind.hpp:23  = 512   */
  float boundarylayer;  /* Wind::boundarylayer wind.hpp:24  = 2.0   */
  float suspension;     /* Wind::suspension    wind.hpp:25  = 0.05  */
  float gravity;        /* Wind::gravity       wind.hpp:26  = 0.1   */
  float lrate;          /* World::lrate        world.hpp:30 = 0.1   */
  float maxdiff;        /* World::maxdiff      world.hpp:31 = 0.005 */
  float settling;       /* World::settling     world.hpp:32 = 1.0   */
  int   mapscale;       /* quad::mapscale      cellpool.h:184 = 80  */
  float pspeed[3];      /* Wind::pspeed        wind.hpp:13  = normalize(1,0,0) */
} swo_params;

typedef struct swo_rand_state {  /* glibc TYPE_3: 31-word additive feedback ring */
  int32_t r[34];
  int f, b;                      /* front / rear indices into r[3..33] */
} swo_rand_state;

typedef struct swo_world {
  int nx, nz;            /* x = slow index, z = fast index (math.h:11: p.x*s.y + p.y) */
  float* plane[9];       /* SoA mirror of quad::cell (cellpool.h:226-240), same order:
                            height, massflow, momentumx, momentumy, momentumz,
                            massflow_track, momentumx_track, momentumy_track, momentumz_track */
  swo_params p;
  swo_rand_state rng;
  /* Layered-terrain EXTENSION (not in the reference snapshot; restates the product's own
   * specification, parity unpinned): bedrock plane, NULL = the reference's single layer.
   * Only swo_cascade / swo_cascade_sweep honour it: a transfer is capped by the giving
   * cell's sand max(height - rock, 0). */
  float* rock;
  /* Periodic boundary: the variant the reference carries as comments (cellpool.h:96-113 slice::get,
   * :447-465 map::get): oob() is never true and every cell access wraps its indices. 0 = the live
   * closed boundary. Pinned against the reference compiled with those comment blocks swapped in
   * (oracle/Makefile, libswe_ref_p<N>.so). */
  int periodic;
} swo_world;

typedef struct swo_particle {  /* struct Wind, wind.hpp:4-21 */
  int age;
  float sediment;
  float pos[3];
  float speed[3];
} swo_particle;

typedef struct swo_stats {
  long long steps;       /* completed Wind::fly calls (returned true) */
  long long exits_oob;
  long long exits_age;
  double sediment_lost;  /* sum of Wind::sediment at death */
} swo_stats;

void swo_default_params(swo_params* p);
swo_world* swo_create(int nx, int nz);
void swo_destroy(swo_world* w);
void swo_set_rock(swo_world* w, const float* rock);   /* copies nx*nz values; NULL clears */

void swo_srand(swo_rand_state* s, unsigned seed);
int  swo_rand(swo_rand_state* s);
float swo_expf(float x);

/* cellpool.h:372-428 for a single tile of edge nx==nz (Gaussian hill, normalised to [0,0.5]) */
void swo_init_gaussian(swo_world* w);
void swo_hill_raw(swo_world* w);   /* the Gaussian hill of cellpool.h:378-385 before normalisation */

void swo_normal(const swo_world* w, int x, int z, float out[3]);
void swo_spawn(swo_world* w, swo_particle* q);            /* world.hpp:52-54 + Wind ctor */
void swo_particle_at(swo_particle* q, float x, float z);  /* Wind(vec2) */
int  swo_fly(swo_world* w, swo_particle* q);              /* 1 = step completed, 0 = died */
void swo_cascade(swo_world* w, int x, int z);
void swo_erode(swo_world* w, int cycles, swo_stats* st);  /* st may be NULL */
void swo_cascade_sweep(swo_world* w, int sweeps);         /* 9-colour order, c=0..8, cx=c/3, cz=c%3 */
float swo_max_excess(const swo_world* w);                 /* max over pairs of |dh| - d*maxdiff (<=0: relaxed) */
int  swo_trace(swo_world* w, float x, float z, int max_records, float* out8);
uint64_t swo_hash_height(const swo_world* w);             /* FNV-1a-64 over fp32 words */
void swo_vertices(const swo_world* w, float* out12);       /* updatenode, cellpool.h:302-321: 12 floats per cell */
/* host generate, noise stack (cellpool.h:366-368,388-411 + FastNoiseLite.h; see swe_oracle.c) */
float swo_fnl_single3(int seed, float x, float y, float z);            /* SingleOpenSimplex2 3D */
float swo_fnl_getnoise(float frequency, float x, float y, float z);    /* GetNoise, FBm of 3 octaves */
void swo_noise_stack(swo_world* w, int seed, int layers, int tile);    /* height += layers of noise */
void swo_noise_rows(float* rows, int x_lo, int x_hi, int nz, int seed, int layers, int tile);  /* same, one band */
void swo_normalize(swo_world* w);                                      /* cellpool.h:416-428 */

#ifdef __cplusplus
}
#endif
#endif
