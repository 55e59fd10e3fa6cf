// TEST HELPER: source compatibility of include/swe_world.hpp with the reference's caller.
//
// The blocks between "BEGIN/END reference caller lines" are the reference application's own lines
// (/root/reference/SimpleWindErosion.cpp:11-12,26-42 and :211-231) quoted VERBATIM as the thing under test -- they must
// compile and run unchanged against the facade. Everything else in this file stands in for what the reference gets
// from TinyEngine / OpenGL / glm (absent here, out of scope): a recording Vertexpool, Tiny::loop, image::make.
//   g++ -std=c++17 -O2 -Iinclude tests/helpers/frame_loop_compat.cpp -Lsimplewinderosion_b200 -lswe_b200 -o compat
//   ./compat <size> <seed> <frames>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

#include "swe_world.hpp"

// ---- stand-ins for glm / TinyEngine / the OpenGL vertex pool ------------------------------------------------------
namespace glm {
using ivec2 = swe::ivec2;
struct vec3 { float x, y, z; vec3(float a = 0, float b = 0, float c = 0) : x(a), y(b), z(c) {} };
struct vec4 { float x, y, z, w; vec4(double a, double b, double c, double d) : x((float)a), y((float)b), z((float)c), w((float)d) {} };
}  // namespace glm
using namespace glm;
using namespace std;
struct Vertex {   // vertexpool.hpp:6-26
  glm::vec3 position, normal, tangent, bitangent;
  Vertex(glm::vec3 p, glm::vec3 n, glm::vec3 t, glm::vec3 b) : position(p), normal(n), tangent(t), bitangent(b) {}
};
template <typename T> struct Vertexpool {   // vertexpool.hpp: only what the caller lines touch
  std::vector<unsigned> indices;
  std::vector<float> data;                  // 12 floats per vertex, recorded
  std::vector<unsigned> slots;
  void reserve(int k, int n) { data.assign((size_t)k * n * 12, 0.0f); slots.assign(1, 0u); }
  unsigned* section(int, int, swe::vec3, size_t) { return slots.data(); }
  template <typename... A> void fill(unsigned*, int k, A&&... a) {
    T v(std::forward<A>(a)...);
    const float f[12] = {v.position.x, v.position.y, v.position.z, v.normal.x, v.normal.y, v.normal.z,
                         v.tangent.x, v.tangent.y, v.tangent.z, v.bitangent.x, v.bitangent.y, v.bitangent.z};
    for (int i = 0; i < 12; i++) data[(size_t)k * 12 + i] = f[i];
  }
};
static int g_frames = 1;
static bool paused = false;                 // model.hpp:7 (true there until a key is pressed)
namespace Tiny { template <typename F> void loop(F f) { for (int i = 0; i < g_frames; i++) f(); } }
namespace image {
template <typename F> std::vector<glm::vec4> make(F f, const ivec2 res) {
  std::vector<glm::vec4> px;
  for (int x = 0; x < res.x; x++) for (int y = 0; y < res.y; y++) px.push_back(f(ivec2(x, y)));
  return px;
}
}  // namespace image
static struct { std::vector<glm::vec4> last; void raw(std::vector<glm::vec4> p) { last = std::move(p); } } windmap;

// ---- BEGIN reference caller lines (SimpleWindErosion.cpp:11-12) ---------------------------------------------------
mappool::pool<quad::cell> cellpool;
Vertexpool<Vertex> vertexpool;
// ---- END ---------------------------------------------------------------------------------------------------------

int main( int argc, char* args[] ) {
  quad::resize(argc > 1 ? atoi(args[1]) : 128);      // run-time map size (compile-time quad::tilesize in the reference)
  g_frames = argc > 3 ? atoi(args[3]) : 2;
  World::mode = SWE_MODE_DETERMINISTIC;              // bitwise comparable with the oracle
  argc = argc > 2 ? 2 : 1; if (argc == 2) args[1] = args[2];
  try {
// ---- BEGIN reference caller lines (SimpleWindErosion.cpp:26-42) ---------------------------------------------------
  World world;

  if(argc >= 2){
    World::SEED = std::stoi(args[1]);
    srand(std::stoi(args[1]));
  }
  else {
    World::SEED = time(NULL);
    srand(World::SEED);
  }

  cellpool.reserve(quad::area);
  vertexpool.reserve(quad::tilearea, quad::maparea);
  World::map.init(vertexpool, cellpool, World::SEED);

  for(auto& node: world.map.nodes){
    updatenode(vertexpool, node);
  }
// ---- END ---------------------------------------------------------------------------------------------------------

// ---- BEGIN reference caller lines (SimpleWindErosion.cpp:211-231) -------------------------------------------------
  int n = 0;
  Tiny::loop([&](){

    if(paused)
      return;

    world.erode(8*quad::tilesize);

    for(auto& node: world.map.nodes){
      updatenode(vertexpool, node);
    }
    cout<<n++<<endl;

    windmap.raw(image::make([&](const ivec2 p){
      auto node = world.map.get(p);
      auto cell = node->get(p);
      float mx = cell->momentumx;
      float my = cell->momentumy;
      float mz = cell->momentumz;
      return glm::vec4(0.5f*(1.0f+erf(2*mx)), 0.5f*(1.0f+erf(2*my)), 0.5f*(1.0f+erf(2*mz)), 1.0);    }, quad::res));
  });
// ---- END ---------------------------------------------------------------------------------------------------------

  // what the test checks: the engine's state, what the vertex pool received, what the wind map shows
  swe_stats st;
  uint64_t hash = 0;
  swe::check(swe_get_stats(World::map.engine, &st));
  swe::check(swe_hash_height(World::map.engine, &hash));
  const int N = quad::size, cx = N / 4, cz = N / 2;
  const float* v = &vertexpool.data[((size_t)cx * N + cz) * 12];
  const glm::vec4 w = windmap.last[(size_t)cx * N + cz];
  Wind probe(swe::vec2((float)cx, (float)cz));       // public fields, one fly() step
  const bool alive = probe.fly();
  std::printf("steps=%llu hash=%016llx cell_height=%.9g vertex_y=%.9g normal_y=%.9g tangent_x=%.9g wind_r=%.9g "
              "probe_alive=%d probe_age=%d probe_x=%.9g\n",
              (unsigned long long)st.steps, (unsigned long long)hash, world.map.get(ivec2(cx, cz))->get(ivec2(cx, cz))->height,
              v[1], v[4], v[6], w.x, (int)alive, probe.age, probe.pos.x);
  World::release();
  } catch (const swe::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
