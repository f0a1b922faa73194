#include "world.h"

#include "cuda_backend.h"

namespace m4t {

namespace {
World* g_world = nullptr;
std::mutex g_world_mu;
std::shared_ptr<NetEngine> g_engine;  // set by set_network() for jobs that span nodes
}  // namespace

CommContext::CommContext(int rank, int size, const std::string& job_id) : job_id_(job_id) {
  ctl_ = std::make_unique<Control>(rank, size, job_id);
  cpu_ = std::make_unique<CpuBackend>(*ctl_);
}

CommContext::CommContext(std::shared_ptr<NetLink> link, const std::string& job_id, int local_rank, int local_size)
    : job_id_(job_id), net_(std::move(link)) {
  netbe_ = std::make_unique<NetBackend>(net_);
  // do all nodes hold `local_size` consecutive ranks?  (every rank checks the same numbers)
  const int P = net_->size(), r = net_->rank();
  bool uniform = local_size > 1 && local_size < P && P % local_size == 0 && r % local_size == local_rank;
  {
    const int64_t mine[2] = {local_size, uniform ? 1 : 0};
    std::vector<int64_t> all(static_cast<size_t>(P) * 2);
    net_->allgather_i64(mine, 2, all.data());
    for (int p = 0; p < P; ++p) uniform = uniform && all[static_cast<size_t>(p) * 2] == local_size && all[static_cast<size_t>(p) * 2 + 1] == 1;
  }
  if (uniform) {
    ranks_per_node_ = local_size;
    // the ranks of this node share a control segment and arenas; one TCP rail per local index connects the nodes
    const int node = r / local_size, nodes = P / local_size;
    ctl_ = std::make_unique<Control>(local_rank, local_size, job_id + "_node" + std::to_string(node));
    {
      // a rank that spins on its node's shared memory keeps its TCP traffic moving (a buffered Isend to another node
      // must not stall because its sender entered a node-local wait)
      std::weak_ptr<NetEngine> eng = net_->engine_ptr();
      ctl_->set_idle_hook([eng] {
        if (auto e = eng.lock()) e->poke();
      });
    }
    cpu_ = std::make_unique<CpuBackend>(*ctl_);
    std::vector<int> rail;
    for (int k = 0; k < nodes; ++k) rail.push_back(net_->members()[static_cast<size_t>(k * local_size + local_rank)]);
    auto rail_link = std::make_shared<NetLink>(net_->engine_ptr(), net_comm_id(job_id + "_rail" + std::to_string(local_rank)),
                                               std::move(rail), node);
    hier_ = std::make_unique<HierBackend>(*cpu_, *netbe_, std::move(rail_link));
  } else {
    // a private one-rank segment keeps the shared-memory classes valid; nothing is exchanged through it
    ctl_ = std::make_unique<Control>(0, 1, job_id + "_n" + std::to_string(net_->engine().rank()));
    cpu_ = std::make_unique<CpuBackend>(*ctl_);
  }
}

CommContext::~CommContext() { shutdown(); }

void CommContext::shutdown() {
  if (!ctl_) return;
  // every member reaches this point before anybody unmaps / unlinks shared segments (or closes its sockets)
  // once one teardown handshake of this process has failed (a peer is gone or stuck) the remaining ones only get a
  // moment: every further wait would just delay the exit of a job that is already broken
  static bool handshake_failed = false;
  const double exit_timeout = handshake_failed ? 0.2 : static_cast<double>(env_i64("M4T_EXIT_TIMEOUT_S", 10));
  if (net_ && !net_->quiesce(exit_timeout)) handshake_failed = true;
  if (!ctl_->quiesce(handshake_failed ? 0.2 : exit_timeout)) handshake_failed = true;
  cuda_.reset();
  hier_.reset();
  netbe_.reset();
  net_.reset();
  cpu_.reset();
  ctl_.reset();
}

void CommContext::barrier() {
  if (net_) net_->barrier();
  else ctl_->barrier();
}

void CommContext::allgather_i64(const int64_t* mine, int k, int64_t* all) {
  if (net_) net_->allgather_i64(mine, k, all);
  else ctl_->allgather_i64(mine, k, all);
}

void CommContext::init_cuda(int device, int64_t stage_mb, int64_t symm_mb) {
  if (cuda_) return;
  M4T_CHECK(!net_, "the NVLink backend needs all ranks on one node; this job spans nodes (CUDA tensors are staged "
                   "through host memory and the TCP transport)");
  cuda_ = std::make_unique<CudaBackend>(*ctl_, device, stage_mb, symm_mb);
}

void CommContext::shutdown_cuda() { cuda_.reset(); }

World::World() {
  env_ = world_env_from_environment();
  if (g_engine) {
    M4T_CHECK(g_engine->rank() == env_.rank && g_engine->size() == env_.size,
              "the TCP mesh was built for rank " << g_engine->rank() << "/" << g_engine->size() << " but the environment says "
                                                 << env_.rank << "/" << env_.size);
    std::vector<int> members(static_cast<size_t>(env_.size));
    for (int p = 0; p < env_.size; ++p) members[static_cast<size_t>(p)] = p;
    auto link = std::make_shared<NetLink>(g_engine, net_comm_id(env_.job_id + "_world"), std::move(members), env_.rank);
    // ranks per node: LOCAL_WORLD_SIZE as the launchers export it (M4T_NET_LOCAL_SIZE overrides: simulated nodes)
    int64_t local_size = env_i64("M4T_NET_LOCAL_SIZE", 0);
    int local_rank = env_.local_rank;
    if (local_size > 0) {
      local_rank = env_.rank % static_cast<int>(local_size);
    } else {
      local_size = env_i64("LOCAL_WORLD_SIZE", 1);
    }
    ctx_ = std::make_shared<CommContext>(std::move(link), env_.job_id, local_rank, static_cast<int>(local_size));
  } else {
    ctx_ = std::make_shared<CommContext>(env_.rank, env_.size, env_.job_id);
  }
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

void World::set_network(std::shared_ptr<NetEngine> engine) {
  std::lock_guard<std::mutex> g(g_world_mu);
  M4T_CHECK(!g_world, "the TCP mesh must be set up before the world communicator exists");
  g_engine = std::move(engine);
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
  g_engine.reset();
}

void World::init_cuda(int device) {
  std::lock_guard<std::recursive_mutex> g(mu_);
  ctx_->init_cuda(device);
}

}  // namespace m4t
