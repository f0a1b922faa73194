#include "world.h"

#include "cuda_backend.h"

namespace m4t {

namespace {
World* g_world = nullptr;
std::mutex g_world_mu;
}  // namespace

World::World() {
  env_ = world_env_from_environment();
  ctl_ = std::make_unique<Control>(env_.rank, env_.size, env_.job_id);
  cpu_ = std::make_unique<CpuBackend>(*ctl_);
}

World::~World() {
  cuda_.reset();
  cpu_.reset();
  ctl_.reset();
}

World& World::instance() {
  std::lock_guard<std::mutex> g(g_world_mu);
  if (!g_world) g_world = new World();
  return *g_world;
}

bool World::initialised() {
  std::lock_guard<std::mutex> g(g_world_mu);
  return g_world != nullptr;
}

void World::finalize() {
  std::lock_guard<std::mutex> g(g_world_mu);
  if (g_world) {
    delete g_world;
    g_world = nullptr;
  }
}

void World::init_cuda(int device) {
  std::lock_guard<std::recursive_mutex> g(mu_);
  if (cuda_) return;
  cuda_ = std::make_unique<CudaBackend>(*ctl_, device);
}

}  // namespace m4t
