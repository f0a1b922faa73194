#include "world.h"

#include "cuda_backend.h"

namespace m4t {

namespace {
World* g_world = nullptr;
std::mutex g_world_mu;
}  // namespace

CommContext::CommContext(int rank, int size, const std::string& job_id) : job_id_(job_id) {
  ctl_ = std::make_unique<Control>(rank, size, job_id);
  cpu_ = std::make_unique<CpuBackend>(*ctl_);
}

CommContext::~CommContext() { shutdown(); }

void CommContext::shutdown() {
  if (!ctl_) return;
  // every member reaches this point before anybody unmaps / unlinks shared segments
  ctl_->quiesce(static_cast<double>(env_i64("M4T_EXIT_TIMEOUT_S", 10)));
  cuda_.reset();
  cpu_.reset();
  ctl_.reset();
}

void CommContext::init_cuda(int device, int64_t stage_mb, int64_t symm_mb) {
  if (cuda_) return;
  cuda_ = std::make_unique<CudaBackend>(*ctl_, device, stage_mb, symm_mb);
}

void CommContext::shutdown_cuda() { cuda_.reset(); }

World::World() {
  env_ = world_env_from_environment();
  ctx_ = std::make_shared<CommContext>(env_.rank, env_.size, env_.job_id);
}

World::~World() {
  for (auto& c : children_) c->shutdown();  // creation order (identical on all members)
  children_.clear();
  ctx_->shutdown();
  ctx_.reset();
}

void World::release_child(const CommContext* c) {
  for (auto it = children_.begin(); it != children_.end(); ++it)
    if (it->get() == c) {
      children_.erase(it);
      return;
    }
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
  ctx_->init_cuda(device);
}

}  // namespace m4t
