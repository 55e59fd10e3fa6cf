// TEST INFRASTRUCTURE ONLY. Headless harness around the UNMODIFIED reference hot-path
// headers (source/world.hpp, source/wind.hpp, source/cellpool.h, source/include/math.h),
// compiled where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libswe_ref_<N>.so  (one library per map edge N, because quad::tilesize is a
// compile-time constant, cellpool.h:186 -- the Makefile rewrites that one constant in a
// scratch copy). Nothing here is linked into or called by the product; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// Stand-ins the reference needs from its absent dependencies (TinyEngine/glm/OpenGL):
//   * oracle/glm_standin.hpp  -- the ~12 glm operations the path uses
//   * Vertex / Vertexpool<Vertex> below -- render memory named by cellpool.h:268,302,339
//   * `using namespace glm/std`, `uint`  -- the reference relies on TinyEngine for these
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <sstream>
#include <vector>

#include "glm_standin.hpp"

using namespace glm;
using namespace std;
typedef unsigned int uint;

// ---- render-memory stub (source/vertexpool.hpp is OpenGL; out of scope) -----------------
// The stub keeps what updatenode hands to fill() (cellpool.h:312-317: position, normal, tangent,
// bitangent for vertex k) so that ref_updatenode can return the reference's own vertex records.
struct Vertex {};
static std::vector<float> g_vertices;   // 12 floats per vertex once ref_updatenode ran
template <typename T> struct Vertexpool {
  std::vector<unsigned int> indices;
  uint dummy = 0;
  void fill(uint*, int k, glm::vec3 p, glm::vec3 n, glm::vec3 t, glm::vec3 b) {
    if ((size_t)(k + 1) * 12 > g_vertices.size()) return;
    float* v = &g_vertices[(size_t)k * 12];
    v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = n.x; v[4] = n.y; v[5] = n.z;
    v[6] = t.x; v[7] = t.y; v[8] = t.z; v[9] = b.x; v[10] = b.y; v[11] = b.z;
  }
  uint* section(int, int, glm::vec3, size_t) { return &dummy; }
  void resize(const uint*, size_t) {}
  void index() { indices.clear(); indices.shrink_to_fit(); }  // keep the harness small in RAM
  void update() {}
  void reserve(int, int) {}
};

#include "world.hpp"  // resolved to the scratch copy next to the size-rewritten cellpool.h

// same globals as SimpleWindErosion.cpp:11-12
mappool::pool<quad::cell> cellpool;
Vertexpool<Vertex> vertexpool;

static std::vector<float> g_initial_height;
static bool g_inited = false;

static inline float* plane_ptr(quad::cell* c, int plane) {
  switch (plane) {
    case 0: return &c->height;
    case 1: return &c->massflow;
    case 2: return &c->momentumx;
    case 3: return &c->momentumy;
    case 4: return &c->momentumz;
    case 5: return &c->massflow_track;
    case 6: return &c->momentumx_track;
    case 7: return &c->momentumy_track;
    default: return &c->momentumz_track;
  }
}

extern "C" {

int ref_size() { return quad::size; }
int ref_cell_bytes() { return (int)sizeof(quad::cell); }

// SimpleWindErosion.cpp:27-38: seed libc rand(), reserve the pool, generate the map.
void ref_init(unsigned seed) {
  std::stringstream sink;  // swallow the generation banner (cellpool.h:360-363)
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  World::SEED = seed;
  srand(seed);
  if (!g_inited) {
    cellpool.reserve(quad::area);
    vertexpool.reserve(quad::tilearea, quad::maparea);
    World::map.init(vertexpool, cellpool, World::SEED);
    g_initial_height.resize(quad::area);
    for (int i = 0; i < quad::area; i++) g_initial_height[i] = cellpool.root.start[i].height;
    g_inited = true;
  } else {
    for (int i = 0; i < quad::area; i++) {
      quad::cell& c = cellpool.root.start[i];
      std::memset(&c, 0, sizeof(c));
      c.height = g_initial_height[i];
    }
  }
  // map.init leaves every plane except height/massflow uninitialised (new cell[]); the app
  // never reads them before the first erode() zeroes the tracks, but the EMA reads
  // momentum{x,y,z}. Zero them so runs are reproducible.
  for (int i = 0; i < quad::area; i++) {
    quad::cell& c = cellpool.root.start[i];
    c.massflow = 0; c.momentumx = 0; c.momentumy = 0; c.momentumz = 0;
    c.massflow_track = 0; c.momentumx_track = 0; c.momentumy_track = 0; c.momentumz_track = 0;
  }
  std::cout.rdbuf(old);
}

void ref_srand(unsigned seed) { srand(seed); }
int ref_rand() { return rand(); }

void ref_set_params(int maxAge, float boundarylayer, float suspension, float gravity,
                    float lrate, float maxdiff, float settling) {
  Wind::maxAge = maxAge; Wind::boundarylayer = boundarylayer;
  Wind::suspension = suspension; Wind::gravity = gravity;
  World::lrate = lrate; World::maxdiff = maxdiff; World::settling = settling;
}

void ref_get_params(float* out7) {
  out7[0] = (float)Wind::maxAge; out7[1] = Wind::boundarylayer; out7[2] = Wind::suspension;
  out7[3] = Wind::gravity; out7[4] = World::lrate; out7[5] = World::maxdiff;
  out7[6] = World::settling;
}

void ref_set_plane(int plane, const float* src) {
  for (int x = 0; x < quad::size; x++)
    for (int z = 0; z < quad::size; z++)
      *plane_ptr(World::map.getCell(ivec2(x, z)), plane) = src[(size_t)x * quad::size + z];
}

void ref_get_plane(int plane, float* dst) {
  for (int x = 0; x < quad::size; x++)
    for (int z = 0; z < quad::size; z++)
      dst[(size_t)x * quad::size + z] = *plane_ptr(World::map.getCell(ivec2(x, z)), plane);
}

// The real thing: world.hpp:36-77.
void ref_erode(int cycles) { World::erode(cycles); }

// Same cycle with counters. Restates ONLY the loop scaffolding of world.hpp:38-76 around
// the reference's own Wind::fly(); stats = {completed steps, oob exits, age exits}.
void ref_erode_counted(int cycles, long long* stats, double* sediment_lost) {
  long long steps = 0, n_oob = 0, n_age = 0;
  double lost = 0.0;
  for (auto& node : World::map.nodes)
    for (auto [cell, pos] : node.s) {
      cell.massflow_track = 0; cell.momentumx_track = 0;
      cell.momentumy_track = 0; cell.momentumz_track = 0;
    }
  for (auto& node : World::map.nodes)
    for (int i = 0; i < cycles; i++) {
      glm::vec2 newpos = node.pos + ivec2(rand() % quad::tileres.x, rand() % quad::tileres.y);
      newpos += vec2(rand() % 1000, rand() % 1000) / 1000.0f;
      Wind wind(newpos);
      while (wind.fly()) steps++;
      if (wind.age == Wind::maxAge + 2) n_age++; else n_oob++;
      lost += (double)wind.sediment;
    }
  for (auto& node : World::map.nodes)
    for (auto [cell, pos] : node.s) {
      cell.massflow = (1.0f - World::lrate) * cell.massflow + World::lrate * cell.massflow_track;
      cell.momentumx = (1.0f - World::lrate) * cell.momentumx + World::lrate * cell.momentumx_track;
      cell.momentumy = (1.0f - World::lrate) * cell.momentumy + World::lrate * cell.momentumy_track;
      cell.momentumz = (1.0f - World::lrate) * cell.momentumz + World::lrate * cell.momentumz_track;
    }
  stats[0] = steps; stats[1] = n_oob; stats[2] = n_age;
  *sediment_lost = lost;
}

// Fly ONE particle spawned at (x,z) with the reference's Wind::fly; record after every
// completed step: pos.xyz, speed.xyz, sediment, age  (8 floats). Returns completed steps.
int ref_trace(float x, float z, int max_records, float* out) {
  Wind wind(glm::vec2(x, z));
  int n = 0;
  while (wind.fly()) {
    if (n < max_records) {
      float* r = out + (size_t)n * 8;
      r[0] = wind.pos.x; r[1] = wind.pos.y; r[2] = wind.pos.z;
      r[3] = wind.speed.x; r[4] = wind.speed.y; r[5] = wind.speed.z;
      r[6] = wind.sediment; r[7] = (float)wind.age;
    }
    n++;
  }
  return n;
}

void ref_cascade(float x, float z) { World::cascade(vec2(x, z)); }

// Whole-map cascade in the conflict-free 3x3 colour order: colour c=0..8,
// cx=c/3, cz=c%3, centres x==cx, z==cz (mod 3), x outer / z inner, calling the reference's
// own World::cascade for every centre.
void ref_cascade_sweep(int sweeps) {
  for (int s = 0; s < sweeps; s++)
    for (int c = 0; c < 9; c++) {
      int cx = c / 3, cz = c % 3;
      for (int x = cx; x < quad::size; x += 3)
        for (int z = cz; z < quad::size; z += 3) World::cascade(vec2(x, z));
    }
}

// World::cascade on every cell in raster order (the survey's CPU cascade-rate probe).
void ref_cascade_raster() {
  for (int x = 0; x < quad::size; x++)
    for (int z = 0; z < quad::size; z++) World::cascade(vec2(x, z));
}

void ref_normal(int x, int z, float* out3) {
  vec3 n = World::map.normal(ivec2(x, z));
  out3[0] = n.x; out3[1] = n.y; out3[2] = n.z;
}

// FNV-1a-64 over the fp32 bit patterns of `height` in pool order.
unsigned long long ref_hash() {
  unsigned long long h = 1469598103934665603ull;
  for (int i = 0; i < quad::area; i++) {
    uint32_t w;
    std::memcpy(&w, &cellpool.root.start[i].height, 4);
    h ^= w; h *= 1099511628211ull;
  }
  return h;
}

double ref_sum_height() {
  double s = 0;
  for (int i = 0; i < quad::area; i++) s += cellpool.root.start[i].height;
  return s;
}

// quad::updatenode (cellpool.h:302-321), the reference's own per-frame vertex update, for every
// node; out = quad::area records of 12 floats (Vertex layout, vertexpool.hpp:6-26).
void ref_updatenode(float* out) {
  g_vertices.assign((size_t)quad::area * 12, 0.0f);
  for (auto& node : World::map.nodes) quad::updatenode(vertexpool, node);
  std::memcpy(out, g_vertices.data(), g_vertices.size() * sizeof(float));
  g_vertices.clear(); g_vertices.shrink_to_fit();
}

// ---- host generate, noise stack -------------------------------------
// The UNMODIFIED FastNoiseLite class (source/include/FastNoiseLite.h, pulled in by
// cellpool.h) configured exactly as map::init configures it (cellpool.h:366-368).
static FastNoiseLite& ref_noise_object() {
  static FastNoiseLite noise;
  static bool set = false;
  if (!set) {
    noise.SetNoiseType(FastNoiseLite::NoiseType_OpenSimplex2);
    noise.SetFractalType(FastNoiseLite::FractalType_FBm);
    set = true;
  }
  return noise;
}

// noise.SetFrequency(frequency); noise.GetNoise(x, y, z)
float ref_get_noise(float frequency, float x, float y, float z) {
  FastNoiseLite& noise = ref_noise_object();
  noise.SetFrequency(frequency);
  return noise.GetNoise(x, y, z);
}

// The layer loop map::init carries as a comment (cellpool.h:392-409), restated around the
// reference's own noise object: height += scale*GetNoise(p.x, p.y, SEED%10000) for `layers`
// layers of doubling frequency, scale 0.6 *= 0.6.
void ref_noise_stack(int SEED, int layers) {
  FastNoiseLite& noise = ref_noise_object();
  for (auto& node : World::map.nodes) {
    float frequency = 1.0f;
    float scale = 0.6f;
    for (int o = 0; o < layers; o++) {
      noise.SetFrequency(frequency);
      for (auto [cell, pos] : node.s) {
        vec2 p = vec2(node.pos + quad::lodsize * pos) / vec2(quad::tileres);
        cell.height += scale * noise.GetNoise(p.x, p.y, (float)(SEED % 10000));
      }
      frequency *= 2;
      scale *= 0.6;
    }
  }
}

// cellpool.h:416-428 (the tail of map::init), verbatim semantics on the current heights.
void ref_normalize() {
  float min = 0.0f;
  float max = 0.0f;
  for (auto& node : World::map.nodes)
    for (auto [cell, pos] : node.s) {
      min = (min < cell.height) ? min : cell.height;
      max = (max > cell.height) ? max : cell.height;
    }
  for (auto& node : World::map.nodes)
    for (auto [cell, pos] : node.s) cell.height = 0.5 * ((cell.height - min) / (max - min));
}

}  // extern "C"
