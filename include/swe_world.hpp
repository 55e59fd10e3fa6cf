// swe_world.hpp -- header-only C++ mirror of the reference's hot-path interface on top of the
// C-ABI (include/swe.h, libswe_b200.so). It re-creates the names a caller of the reference uses:
//
//   World::SEED, World::map, World::lrate, World::maxdiff, World::settling,
//   World::erode(int cycles), World::cascade(vec2 pos)            source/world.hpp:9-32
//   Wind(vec2 pos), Wind::fly(), Wind::{age,sediment,pos,speed,pspeed},
//   Wind::{maxAge,boundarylayer,suspension,gravity}               source/wind.hpp:4-26
//   quad::cell (36-byte AoS), quad::{mapscale,...}                source/cellpool.h:184-240
//   map.oob(p), map.getCell(p), map.height(p), map.normal(p)      source/cellpool.h:432-480
//
// Differences that are part of the design:
//   * The map size is a run-time value (World::init(nx, nz)); the reference fixes it at compile
//     time (quad::tilesize, cellpool.h:186).
//   * The canonical copy of the map lives in HBM. The host-side AoS pool (the reference's
//     cellpool memory layout) is a mirror: World::map.sync_to_device() after the host wrote
//     cells (generation), World::map.sync_from_device() before the host reads them (rendering).
//     World::erode leaves the device copy current; it does not touch the host mirror.
//   * World::mode selects deterministic (bitwise identical to the reference) or parallel.
//   * Errors throw swe::Error (the reference has no error convention).
// Vector types are minimal aggregates; a project that already has glm can pass glm::vec2 /
// glm::ivec2 to the templated accessors (anything with .x and .y).
#ifndef SWE_WORLD_HPP
#define SWE_WORLD_HPP

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "swe.h"

namespace swe {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error("swe: " + m), code(c) {}
};
inline void check(int rc) {
  if (rc != SWE_OK) throw Error(rc, swe_last_error());
}

struct vec2 { float x = 0, y = 0; vec2() = default; vec2(float a, float b) : x(a), y(b) {} };
struct vec3 { float x = 0, y = 0, z = 0; vec3() = default; vec3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct ivec2 { int x = 0, y = 0; ivec2() = default; ivec2(int a, int b) : x(a), y(b) {} };

}  // namespace swe

namespace quad {

const int mapscale = 80;   // cellpool.h:184 (the engine reads it from swe_params.mapscale)

// Raw interleaved cell data, identical layout to the reference's quad::cell (cellpool.h:226-240)
struct cell {
  float height;
  float massflow;
  float momentumx;
  float momentumy;
  float momentumz;
  float massflow_track;
  float momentumx_track;
  float momentumy_track;
  float momentumz_track;
};
static_assert(sizeof(cell) == 36, "quad::cell must match the reference's 36-byte record");

// quad::map (cellpool.h:335-482): bounds tests and accessors; storage = one AoS pool on the
// host (math::flatten order, x*nz + z) mirroring the device-resident planes.
struct map {
  swe_engine* engine = nullptr;
  int nx = 0, nz = 0;
  std::vector<cell> cells;   // host mirror (the reference's cellpool)

  template <typename P> bool oob(const P& p) const {                    // cellpool.h:432-438
    return p.x < 0 || p.y < 0 || p.x >= nx || p.y >= nz;
  }
  template <typename P> cell* getCell(const P& p) {                     // cellpool.h:467-470
    if (oob(p)) return nullptr;
    return &cells[(size_t)p.x * nz + p.y];
  }
  // height / normal of the DEVICE map (current after World::erode without a host sync)
  template <typename P> float height(const P& p) const {                // cellpool.h:472-476
    if (oob(p)) return 0.0f;
    float h = 0.0f;
    swe::check(swe_height(engine, (int)p.x, (int)p.y, &h));
    return h;
  }
  template <typename P> swe::vec3 normal(const P& p) const {            // cellpool.h:478-480
    float n[3];
    swe::check(swe_normal(engine, (int)p.x, (int)p.y, n));
    return swe::vec3(n[0], n[1], n[2]);
  }

  void sync_to_device() { swe::check(swe_upload_cells(engine, cells.data())); }
  void sync_from_device() { swe::check(swe_download_cells(engine, cells.data())); }
  // only the members a renderer reads each frame (SimpleWindErosion.cpp:219-230)
  void sync_render_fields_from_device() {
    for (int pl : {SWE_HEIGHT, SWE_MOMENTUMX, SWE_MOMENTUMY, SWE_MOMENTUMZ})
      swe::check(swe_download_plane(engine, pl, &cells[0].height + pl, sizeof(cell)));
  }

  // quad::map::init (cellpool.h:339-430): zero the cells, add the Gaussian hill (host: it needs
  // libm's double exp), optionally the `noise_layers` layers of FastNoiseLite OpenSimplex2 FBm
  // that the reference carries as a comment (cellpool.h:388-411; generated on the device,
  // bitwise identical to the host library), normalise to [0, 0.5] (cellpool.h:416-428).
  // The reference's active code is noise_layers = 0.
  void init(int SEED, int noise_layers = 0) {
    for (cell& c : cells) c = cell{};
    const int n = nx;
    const float cx = (float)(n / 4) / (float)n, cz = (float)(n / 2) / (float)n;
    for (int x = 0; x < nx; x++)
      for (int z = 0; z < nz; z++) {
        const float dx = (float)x / (float)n - cx, dz = (float)z / (float)n - cz;
        const float d = std::sqrt(dx * dx + dz * dz);
        cells[(size_t)x * nz + z].height = (float)std::exp((double)(-d * d * (float)n) * 0.2);
      }
    sync_to_device();
    if (noise_layers > 0) swe::check(swe_generate_noise(engine, SEED, noise_layers, n));
    float mn = 0.0f, mx = 0.0f;
    swe::check(swe_height_range(engine, &mn, &mx));
    swe::check(swe_normalize_height(engine, mn, mx));
    swe::check(swe_download_plane(engine, SWE_HEIGHT, &cells[0].height, sizeof(cell)));
  }
};

}  // namespace quad

class World {
 public:
  static unsigned int SEED;
  static quad::map map;
  static float lrate, maxdiff, settling;                                 // world.hpp:26-32
  static int mode;                                                       // SWE_MODE_*
  static int device;

  // replaces cellpool.reserve(quad::area) + World::map.init's node construction
  // (SimpleWindErosion.cpp:36-38) with a run-time size
  static void init(int nx, int nz) {
    release();
    swe::check(swe_create(&map.engine, nx, nz, nullptr, device));
    map.nx = nx; map.nz = nz;
    map.cells.assign((size_t)nx * nz, quad::cell{});
    swe::check(swe_srand(map.engine, SEED));
  }
  static void release() {
    if (map.engine) { swe_destroy(map.engine); map.engine = nullptr; }
  }
  static void seed(unsigned int s) {                                     // SimpleWindErosion.cpp:27-34
    SEED = s;
    std::srand(s);
    if (map.engine) swe::check(swe_srand(map.engine, s));
  }
  static void erode(int cycles);                                         // world.hpp:36
  template <typename P> static void cascade(const P& pos) {              // world.hpp:79
    push_params();
    swe::check(swe_cascade(map.engine, (int)pos.x, (int)pos.y));
  }
  static void push_params();   // World::/Wind:: statics -> engine
};

struct Wind {
  template <typename P> explicit Wind(const P& _pos) { swe_wind_init(&s, (float)_pos.x, (float)_pos.y); }  // wind.hpp:6

  int age() const { return s.age; }
  float sediment() const { return s.sediment; }
  swe::vec3 pos() const { return swe::vec3(s.pos[0], s.pos[1], s.pos[2]); }
  swe::vec3 speed() const { return swe::vec3(s.speed[0], s.speed[1], s.speed[2]); }

  static int maxAge;                                                     // wind.hpp:15-18,23-26
  static float boundarylayer, suspension, gravity;
  static swe::vec3 pspeed;                                               // wind.hpp:13 (per instance there, constant)

  bool fly() {                                                           // wind.hpp:28
    World::push_params();
    int alive = 0;
    swe::check(swe_wind_fly(World::map.engine, &s, &alive));
    return alive != 0;
  }
  swe_wind s;
};

#ifndef SWE_WORLD_NO_DEFINITIONS   // define in all but one translation unit
unsigned int World::SEED = 0;
quad::map World::map;
float World::lrate = 0.1f;
float World::maxdiff = (float)0.005;
float World::settling = (float)1.0;
int World::mode = SWE_MODE_PARALLEL;
int World::device = 0;
int Wind::maxAge = 512;
float Wind::boundarylayer = 2.0f;
float Wind::suspension = 0.05f;
float Wind::gravity = (float)0.1;
swe::vec3 Wind::pspeed = swe::vec3(1.0f, 0.0f, 0.0f);

void World::push_params() {
  swe_params p;
  swe_params_default(&p);
  p.max_age = Wind::maxAge; p.boundarylayer = Wind::boundarylayer;
  p.suspension = Wind::suspension; p.gravity = Wind::gravity;
  p.lrate = lrate; p.maxdiff = maxdiff; p.settling = settling;
  p.mapscale = quad::mapscale;
  p.pspeed[0] = Wind::pspeed.x; p.pspeed[1] = Wind::pspeed.y; p.pspeed[2] = Wind::pspeed.z;
  swe::check(swe_set_params(map.engine, &p));
}

void World::erode(int cycles) {
  push_params();
  swe::check(swe_erode(map.engine, cycles, mode));
}
#endif

#endif  // SWE_WORLD_HPP
