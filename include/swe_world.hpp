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
// Source compatibility with the reference's caller (SimpleWindErosion.cpp:11-42,211-231): the names that file uses
// exist here with the same spelling -- `mappool::pool<quad::cell> cellpool; cellpool.reserve(quad::area);`,
// `World world; World::map.init(vertexpool, cellpool, World::SEED);`, `world.erode(8*quad::tilesize)`,
// `for(auto& node: world.map.nodes) updatenode(vertexpool, node);`, `world.map.get(p)->get(p)->momentumx`,
// `quad::{tilesize,tilearea,tileres,mapsize,maparea,size,area,res}`, public `Wind::{age,sediment,pos,speed,pspeed}`
// (tests/test_gpu_cpp_facade.py compiles that frame loop verbatim against this header).
//
// Differences that are part of the design:
//   * The map size is a run-time value: quad::resize(n) before cellpool.reserve / map.init, or World::init(nx, nz);
//     the reference fixes it at compile time (quad::tilesize, cellpool.h:186). One tile (quad::mapsize = 1, the
//     reference's setting).
//   * With World::autosync (default on) World::erode refreshes the members a renderer reads (height, momentum,
//     massflow) in the host cell pool after every call, so unmodified caller code sees current values; switch it
//     off and call map.sync_*_from_device yourself when the host does not read every frame.
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
#include <type_traits>
#include <utility>
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

// minimal aggregates; they convert to any vector type constructible from their components (glm::vec3, ...)
struct vec2 { float x = 0, y = 0; vec2() = default; vec2(float a, float b) : x(a), y(b) {} };
struct vec3 {
  float x = 0, y = 0, z = 0;
  vec3() = default;
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  template <typename V, typename = decltype(V(0.0f, 0.0f, 0.0f))> operator V() const { return V(x, y, z); }
};
struct ivec2 {
  int x = 0, y = 0;
  ivec2() = default;
  explicit ivec2(int a) : x(a), y(a) {}
  ivec2(int a, int b) : x(a), y(b) {}
};

}  // namespace swe

namespace quad {

const int mapscale = 80;   // cellpool.h:184 (the engine reads it from swe_params.mapscale)

// cellpool.h:186-198 with the reference's names; compile-time constants there, run-time values here
// (quad::resize before the map is reserved). One tile covers the map (mapsize = 1, the reference's setting).
inline int tilesize = 512;
inline int tilearea = tilesize * tilesize;
inline swe::ivec2 tileres = swe::ivec2(tilesize);
const int mapsize = 1;
const int maparea = mapsize * mapsize;
inline int size = mapsize * tilesize;
inline int area = maparea * tilearea;
inline swe::ivec2 res = swe::ivec2(size);
const int lodsize = 1;
const int lodarea = lodsize * lodsize;
inline void resize(int n) {
  tilesize = n; tilearea = n * n; tileres = swe::ivec2(n);
  size = mapsize * n; area = maparea * tilearea; res = swe::ivec2(size);
}

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

}  // namespace quad

// mappool::pool (cellpool.h:116-166): one allocation of cells handed out front to back. Only what the reference's
// caller uses is kept: reserve(n) (SimpleWindErosion.cpp:36) and the bump allocation map::init takes its slice from.
namespace mappool {
template <typename T> struct pool {
  std::vector<T> store;
  size_t used = 0;
  void reserve(size_t n) { store.assign(n, T{}); used = 0; }
  T* get(size_t n) {
    if (used + n > store.size()) return nullptr;
    T* p = store.data() + used;
    used += n;
    return p;
  }
  size_t size() const { return store.size(); }
};
}  // namespace mappool

namespace quad {

struct map;

// quad::node (cellpool.h:242-266): one tile of the map; `vertex` is whatever the renderer's vertex pool returned
// for the tile (opaque here). get / height read the host cell pool, normal asks the device.
struct node {
  swe::ivec2 pos = swe::ivec2(0);
  unsigned* vertex = nullptr;
  map* owner = nullptr;
  template <typename P> inline cell* get(const P& p);
  template <typename P> inline bool oob(const P& p) const;
  template <typename P> inline float height(const P& p);
  template <typename P> inline swe::vec3 normal(const P& p) const;
};

// quad::map (cellpool.h:335-482): bounds tests and accessors; storage = one AoS pool on the
// host (math::flatten order, x*nz + z) mirroring the device-resident planes.
struct map {
  swe_engine* engine = nullptr;
  int nx = 0, nz = 0;
  cell* cells = nullptr;      // host mirror (the reference's cellpool memory); nx*nz records
  std::vector<cell> own;      // backing store when no external pool was given
  std::vector<node> nodes;    // cellpool.h:337 (one tile)

  template <typename P> bool oob(const P& p) const {                    // cellpool.h:432-438
    return p.x < 0 || p.y < 0 || p.x >= nx || p.y >= nz;
  }
  template <typename P> node* get(const P& p) {                         // cellpool.h:440-445
    if (oob(p) || nodes.empty()) return nullptr;
    return &nodes[0];
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

  void sync_to_device() { swe::check(swe_upload_cells(engine, cells)); }
  void sync_from_device() { swe::check(swe_download_cells(engine, cells)); }
  // only the members a renderer reads each frame (SimpleWindErosion.cpp:219-230)
  void sync_render_fields_from_device() {
    for (int pl : {SWE_HEIGHT, SWE_MASSFLOW, SWE_MOMENTUMX, SWE_MOMENTUMY, SWE_MOMENTUMZ})
      swe::check(swe_download_plane(engine, pl, &cells[0].height + pl, sizeof(cell)));
  }

  void attach(cell* storage) {   // the map's cells live in `storage` (nx*nz records), e.g. the caller's cellpool
    if (storage) cells = storage;
    else { own.assign((size_t)nx * nz, cell{}); cells = own.data(); }
    nodes.assign(1, node{});
    nodes[0].pos = swe::ivec2(0); nodes[0].owner = this;
  }

  // quad::map::init (cellpool.h:339-430): zero the cells, add the Gaussian hill (host: it needs
  // libm's double exp), optionally the `noise_layers` layers of FastNoiseLite OpenSimplex2 FBm
  // that the reference carries as a comment (cellpool.h:388-411; generated on the device,
  // bitwise identical to the host library), normalise to [0, 0.5] (cellpool.h:416-428).
  // The reference's active code is noise_layers = 0.
  void init(int SEED, int noise_layers = 0) {
    for (size_t i = 0; i < (size_t)nx * nz; i++) cells[i] = cell{};
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

  // The reference's own signature (SimpleWindErosion.cpp:38): World::map.init(vertexpool, cellpool, SEED).
  // Creates the engine for quad::size x quad::size cells if World::init has not, takes the map's cells out of
  // `cellpool` (reserved with quad::area, :36), asks the vertex pool for the tile's section when it has one
  // (cellpool.h:349-351), generates (Gaussian hill) and seeds the engine. Defined after World.
  template <typename VP, typename Pool> void init(VP& vertexpool, Pool& cellpool, int SEED);
};

template <typename P> inline cell* node::get(const P& p) { return owner->getCell(swe::ivec2((int)p.x - pos.x, (int)p.y - pos.y)); }
template <typename P> inline bool node::oob(const P& p) const { return owner->oob(swe::ivec2((int)p.x - pos.x, (int)p.y - pos.y)); }
template <typename P> inline float node::height(const P& p) { cell* c = get(p); return c ? c->height : 0.0f; }   // cellpool.h:256-260
template <typename P> inline swe::vec3 node::normal(const P& p) const { return owner->normal(p); }

// quad::updatenode (cellpool.h:302-321): position / normal / tangent / bitangent of every cell of the tile into
// the renderer's vertex pool, `vertexpool.fill(t.vertex, index, P, n, T - P, B - P)`. The records are computed on the
// device (swe_download_vertices: bitwise what the reference's loop computes) and handed over with the reference's
// call; a renderer that maps its vertex buffer into CUDA uses swe_write_vertices_device and skips the host.
template <typename VP> void updatenode(VP& vertexpool, node& t) {
  map& m = *t.owner;
  static thread_local std::vector<float> rec;
  rec.resize((size_t)m.nx * m.nz * 12);
  swe::check(swe_download_vertices(m.engine, rec.data()));
  for (int x = 0; x < m.nx; x++)
    for (int z = 0; z < m.nz; z++) {
      const float* r = &rec[((size_t)x * m.nz + z) * 12];
      vertexpool.fill(t.vertex, x * m.nz + z, swe::vec3(r[0], r[1], r[2]), swe::vec3(r[3], r[4], r[5]),
                      swe::vec3(r[6], r[7], r[8]), swe::vec3(r[9], r[10], r[11]));
    }
}

}  // namespace quad

using quad::updatenode;

class World {
 public:
  static unsigned int SEED;
  static quad::map map;
  static float lrate, maxdiff, settling;                                 // world.hpp:26-32
  static int mode;                                                       // SWE_MODE_*
  static int device;
  static bool autosync;   // refresh the renderer's members of the host cells after every erode

  // replaces cellpool.reserve(quad::area) + World::map.init's node construction
  // (SimpleWindErosion.cpp:36-38) with a run-time size
  static void init(int nx, int nz, quad::cell* storage = nullptr) {
    release();
    swe::check(swe_create(&map.engine, nx, nz, nullptr, device));
    map.nx = nx; map.nz = nz;
    map.attach(storage);
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
  // north_star's name for the host-side generate step: the reference's map::init on the current map
  static void generate(int noise_layers = 0) { map.init((int)SEED, noise_layers); }
  static void erode(int cycles);                                         // world.hpp:36
  template <typename P> static void cascade(const P& pos) {              // world.hpp:79
    push_params();
    swe::check(swe_cascade(map.engine, (int)pos.x, (int)pos.y));
  }
  static void push_params();   // World::/Wind:: statics -> engine
};

namespace swe { namespace detail {
template <typename VP, typename = void> struct has_section : std::false_type {};
template <typename VP>
struct has_section<VP, decltype((void)std::declval<VP&>().section(0, 0, swe::vec3(), 0))> : std::true_type {};
template <typename VP> unsigned* section_of(VP& vp, std::true_type) {
  return (unsigned*)vp.section(quad::tilearea / quad::lodarea, 0, swe::vec3(0.0f, 0.0f, 0.0f), vp.indices.size());
}
template <typename VP> unsigned* section_of(VP&, std::false_type) { return nullptr; }
}}  // namespace swe::detail

template <typename VP, typename Pool> void quad::map::init(VP& vertexpool, Pool& cellpool, int SEED) {
  quad::cell* storage = cellpool.get((size_t)quad::tilearea / quad::lodarea);            // cellpool.h:351
  if (!storage) throw swe::Error(SWE_ERR_INVALID, "cellpool.reserve(quad::area) must come before map.init");
  World::SEED = (unsigned)SEED;
  if (!engine || nx != quad::size || nz != quad::size) World::init(quad::size, quad::size, storage);
  else attach(storage);
  nodes[0].vertex = swe::detail::section_of(vertexpool, swe::detail::has_section<VP>{});    // cellpool.h:349-350
  swe::check(swe_srand(engine, World::SEED));   // the caller's srand(SEED), SimpleWindErosion.cpp:29,33
  init(SEED, 0);
}

// struct Wind (wind.hpp:4-26): public particle state, one fly() step on the engine's map (deterministic mode).
struct Wind {
  template <typename P> explicit Wind(const P& _pos) : pos((float)_pos.x, 0.0f, (float)_pos.y) {}     // wind.hpp:6

  int age = 0;
  float sediment = 0.0f;
  swe::vec3 pos;
  swe::vec3 speed = swe::vec3(0.0f, 0.0f, 0.0f);
  swe::vec3 pspeed = prevailing;                                          // wind.hpp:13 (a per-instance constant there)

  static int maxAge;                                                     // wind.hpp:15-18,23-26
  static float boundarylayer, suspension, gravity;
  static swe::vec3 prevailing;   // what World::erode's particles get as pspeed

  bool fly() {                                                           // wind.hpp:28
    const swe::vec3 keep = prevailing;
    prevailing = pspeed;
    World::push_params();
    prevailing = keep;
    swe_wind s;
    s.age = age; s.sediment = sediment;
    s.pos[0] = pos.x; s.pos[1] = pos.y; s.pos[2] = pos.z;
    s.speed[0] = speed.x; s.speed[1] = speed.y; s.speed[2] = speed.z;
    int alive = 0;
    swe::check(swe_wind_fly(World::map.engine, &s, &alive));
    age = s.age; sediment = s.sediment;
    pos = swe::vec3(s.pos[0], s.pos[1], s.pos[2]);
    speed = swe::vec3(s.speed[0], s.speed[1], s.speed[2]);
    return alive != 0;
  }
};

#ifndef SWE_WORLD_NO_DEFINITIONS   // define in all but one translation unit
unsigned int World::SEED = 0;
quad::map World::map;
float World::lrate = 0.1f;
float World::maxdiff = (float)0.005;
float World::settling = (float)1.0;
int World::mode = SWE_MODE_PARALLEL;
int World::device = 0;
bool World::autosync = true;
int Wind::maxAge = 512;
float Wind::boundarylayer = 2.0f;
float Wind::suspension = 0.05f;
float Wind::gravity = (float)0.1;
swe::vec3 Wind::prevailing = swe::vec3(1.0f, 0.0f, 0.0f);

void World::push_params() {
  swe_params p;
  swe_params_default(&p);
  p.max_age = Wind::maxAge; p.boundarylayer = Wind::boundarylayer;
  p.suspension = Wind::suspension; p.gravity = Wind::gravity;
  p.lrate = lrate; p.maxdiff = maxdiff; p.settling = settling;
  p.mapscale = quad::mapscale;
  p.pspeed[0] = Wind::prevailing.x; p.pspeed[1] = Wind::prevailing.y; p.pspeed[2] = Wind::prevailing.z;
  swe::check(swe_set_params(map.engine, &p));
}

void World::erode(int cycles) {
  push_params();
  swe::check(swe_erode(map.engine, cycles, mode));
  if (autosync) map.sync_render_fields_from_device();
}
#endif

#endif  // SWE_WORLD_HPP
