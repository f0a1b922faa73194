#include "plan.h"

#include <algorithm>
#include <numeric>

#include "common.h"

namespace m4t {

void normalize_job(SlabJob& j) {
  // 1. compact away unit loops (keep order, innermost last)
  int64_t n[3], ss[3], ds[3];
  int k = 0;
  for (int i = 0; i < 3; ++i) {
    if (j.n[i] != 1) {
      n[k] = j.n[i];
      ss[k] = j.ss[i];
      ds[k] = j.ds[i];
      ++k;
    }
  }
  // 2. fold innermost loops that simply continue the contiguous run
  while (k > 0 && ss[k - 1] == j.run && ds[k - 1] == j.run) {
    j.run *= n[k - 1];
    --k;
  }
  // 3. fold adjacent outer loops that are jointly linear
  for (int i = k - 2; i >= 0; --i) {
    if (ss[i] == ss[i + 1] * n[i + 1] && ds[i] == ds[i + 1] * n[i + 1]) {
      n[i + 1] *= n[i];
      // remove loop i by shifting the tail down
      for (int m = i; m < k - 1; ++m) {
        n[m] = n[m + 1];
        ss[m] = ss[m + 1];
        ds[m] = ds[m + 1];
      }
      --k;
    }
  }
  // 4. right-align into the 3 slots
  for (int i = 0; i < 3; ++i) {
    j.n[i] = 1;
    j.ss[i] = 0;
    j.ds[i] = 0;
  }
  for (int i = 0; i < k; ++i) {
    j.n[3 - k + i] = n[i];
    j.ss[3 - k + i] = ss[i];
    j.ds[3 - k + i] = ds[i];
  }
}

Axis3 split_axis(const std::vector<int64_t>& shape, int64_t axis) {
  Axis3 a;
  const int64_t nd = static_cast<int64_t>(shape.size());
  M4T_CHECK(axis >= 0 && axis < nd, "axis " << axis << " out of range for a " << nd << "-d tensor");
  for (int64_t i = 0; i < axis; ++i) a.before *= shape[i];
  a.axis = shape[axis];
  for (int64_t i = axis + 1; i < nd; ++i) a.after *= shape[i];
  return a;
}

namespace {
std::vector<int64_t> exclusive_scan(const std::vector<int64_t>& v) {
  std::vector<int64_t> out(v.size() + 1, 0);
  for (size_t i = 0; i < v.size(); ++i) {
    M4T_CHECK(v[i] >= 0, "negative axis length " << v[i] << " on rank " << i);
    out[i + 1] = out[i] + v[i];
  }
  return out;
}
}  // namespace

// Every rank walks its jobs in the same item order, so with jobs sorted by peer all ranks
// would pull from peer 0 at the same time, then from peer 1, ... and share ONE GPU's NVLink
// egress.  Start each rank at itself and continue with rank+1, rank+2, ...: at any moment the
// (reader -> source) pairs form a permutation and every link is busy.
static void rotate_jobs(PullPlan& plan, int rank, int size) {
  std::stable_sort(plan.jobs.begin(), plan.jobs.end(), [&](const SlabJob& a, const SlabJob& b) {
    return (a.peer - rank + size) % size < (b.peer - rank + size) % size;
  });
}

PullPlan plan_gather(int rank, int size, int root, int64_t before, int64_t after,
                     const std::vector<int64_t>& axis_len, bool all) {
  PullPlan plan;
  auto displ = exclusive_scan(axis_len);
  const int64_t total = displ[size];
  const bool i_receive = all || rank == root;
  // who reads my input?  everyone (allgather) or just root
  const bool someone_else_reads_me = all ? (size > 1) : (rank != root);
  plan.stage_elems = someone_else_reads_me ? before * axis_len[rank] * after : 0;
  for (int p = 0; p < size; ++p) {
    bool staged = all ? (size > 1) : (p != root);
    if (staged) plan.max_stage_elems = std::max(plan.max_stage_elems, before * axis_len[p] * after);
  }
  plan.out_elems = i_receive ? before * total * after : 0;
  plan.max_out_elems = before * total * after;
  plan.replicated_output = all;
  if (i_receive) {
    for (int p = 0; p < size; ++p) {
      if (axis_len[p] == 0 || before == 0 || after == 0) continue;
      SlabJob j;
      j.peer = p;
      j.src_off = 0;
      j.dst_off = displ[p] * after;
      j.n[2] = before;
      j.ss[2] = axis_len[p] * after;
      j.ds[2] = total * after;
      j.run = axis_len[p] * after;
      normalize_job(j);
      plan.jobs.push_back(j);
    }
  }
  rotate_jobs(plan, rank, size);
  if (all) {
    plan.before = before;
    plan.after = after;
    plan.axis_len.assign(axis_len.begin(), axis_len.begin() + size);
  }
  return plan;
}

PullPlan plan_scatter(int rank, int size, int root, int64_t before, int64_t after,
                      const std::vector<int64_t>& numelem) {
  PullPlan plan;
  auto displ = exclusive_scan(numelem);
  const int64_t total = displ[size];
  const int64_t root_elems = before * total * after;
  plan.stage_elems = (rank == root && size > 1) ? root_elems : 0;
  plan.max_stage_elems = size > 1 ? root_elems : 0;
  plan.out_elems = before * numelem[rank] * after;
  plan.max_out_elems = before * (*std::max_element(numelem.begin(), numelem.begin() + size)) * after;
  if (plan.out_elems > 0) {
    SlabJob j;
    j.peer = root;
    j.src_off = displ[rank] * after;
    j.dst_off = 0;
    j.n[2] = before;
    j.ss[2] = total * after;
    j.ds[2] = numelem[rank] * after;
    j.run = numelem[rank] * after;
    normalize_job(j);
    plan.jobs.push_back(j);
  }
  return plan;
}

PullPlan plan_alltoall(int rank, int size, const std::vector<int64_t>& shape, int64_t gatheraxis,
                       int64_t scatteraxis, const std::vector<int64_t>& gather_len,
                       const std::vector<int64_t>& numelem) {
  M4T_CHECK(gatheraxis != scatteraxis, "plan_alltoall requires distinct axes");
  PullPlan plan;
  const int64_t nd = static_cast<int64_t>(shape.size());
  M4T_CHECK(gatheraxis >= 0 && gatheraxis < nd && scatteraxis >= 0 && scatteraxis < nd,
            "Alltoall axes (" << gatheraxis << "," << scatteraxis << ") out of range for a " << nd << "-d tensor");
  const int64_t a = std::min(gatheraxis, scatteraxis), b = std::max(gatheraxis, scatteraxis);
  int64_t d0 = 1, d1 = 1, d2 = 1;
  for (int64_t i = 0; i < a; ++i) d0 *= shape[i];
  for (int64_t i = a + 1; i < b; ++i) d1 *= shape[i];
  for (int64_t i = b + 1; i < nd; ++i) d2 *= shape[i];
  auto goff = exclusive_scan(gather_len);
  auto soff = exclusive_scan(numelem);
  const int64_t G = goff[size], S = soff[size];
  const int64_t nr = numelem[rank];
  plan.stage_elems = size > 1 ? d0 * gather_len[rank] * d1 * S * d2 : 0;
  for (int p = 0; p < size; ++p)
    if (size > 1) plan.max_stage_elems = std::max(plan.max_stage_elems, d0 * gather_len[p] * d1 * S * d2);
  plan.out_elems = d0 * G * d1 * nr * d2;
  plan.max_out_elems = d0 * G * d1 * (*std::max_element(numelem.begin(), numelem.begin() + size)) * d2;
  if (plan.out_elems == 0) return plan;
  for (int p = 0; p < size; ++p) {
    const int64_t Gp = gather_len[p];
    if (Gp == 0) continue;
    SlabJob j;
    j.peer = p;
    if (gatheraxis < scatteraxis) {
      // layout [d0, g, d1, s, d2]
      j.src_off = soff[rank] * d2;
      j.dst_off = goff[p] * d1 * nr * d2;
      j.n[0] = d0; j.ss[0] = Gp * d1 * S * d2; j.ds[0] = G * d1 * nr * d2;
      j.n[1] = Gp; j.ss[1] = d1 * S * d2;      j.ds[1] = d1 * nr * d2;
      j.n[2] = d1; j.ss[2] = S * d2;           j.ds[2] = nr * d2;
      j.run = nr * d2;
    } else {
      // layout [d0, s, d1, g, d2]
      j.src_off = soff[rank] * d1 * Gp * d2;
      j.dst_off = goff[p] * d2;
      j.n[0] = d0; j.ss[0] = S * d1 * Gp * d2; j.ds[0] = nr * d1 * G * d2;
      j.n[1] = nr; j.ss[1] = d1 * Gp * d2;     j.ds[1] = d1 * G * d2;
      j.n[2] = d1; j.ss[2] = Gp * d2;          j.ds[2] = G * d2;
      j.run = Gp * d2;
    }
    normalize_job(j);
    plan.jobs.push_back(j);
  }
  rotate_jobs(plan, rank, size);
  return plan;
}

PullPlan plan_repartition(int rank, int size, int64_t before, int64_t after,
                          const std::vector<int64_t>& cur_len, const std::vector<int64_t>& new_len) {
  PullPlan plan;
  auto cur = exclusive_scan(cur_len);
  auto nw = exclusive_scan(new_len);
  M4T_CHECK(cur[size] == nw[size], "Alltoall(same axis): requested lengths sum to " << nw[size]
                                       << " but the global axis has " << cur[size] << " rows");
  plan.stage_elems = size > 1 ? before * cur_len[rank] * after : 0;
  for (int p = 0; p < size; ++p)
    if (size > 1) plan.max_stage_elems = std::max(plan.max_stage_elems, before * cur_len[p] * after);
  plan.out_elems = before * new_len[rank] * after;
  plan.max_out_elems = before * (*std::max_element(new_len.begin(), new_len.begin() + size)) * after;
  if (plan.out_elems == 0) return plan;
  for (int p = 0; p < size; ++p) {
    const int64_t lo = std::max(nw[rank], cur[p]);
    const int64_t hi = std::min(nw[rank + 1], cur[p + 1]);
    if (hi <= lo) continue;
    SlabJob j;
    j.peer = p;
    j.src_off = (lo - cur[p]) * after;
    j.dst_off = (lo - nw[rank]) * after;
    j.n[2] = before;
    j.ss[2] = cur_len[p] * after;
    j.ds[2] = new_len[rank] * after;
    j.run = (hi - lo) * after;
    normalize_job(j);
    plan.jobs.push_back(j);
  }
  rotate_jobs(plan, rank, size);
  return plan;
}

ReducePlan plan_reduce_scatter(int rank, int size, int64_t before, int64_t after,
                               const std::vector<int64_t>& numelem) {
  ReducePlan plan;
  auto off = exclusive_scan(numelem);
  const int64_t total = off[size];
  plan.stage_elems = before * total * after;
  plan.out_elems = before * numelem[rank] * after;
  plan.max_out_elems = before * (*std::max_element(numelem.begin(), numelem.begin() + size)) * after;
  SlabJob& j = plan.box;
  j.peer = -1;
  j.src_off = off[rank] * after;
  j.dst_off = 0;
  j.n[2] = before;
  j.ss[2] = total * after;
  j.ds[2] = numelem[rank] * after;
  j.run = numelem[rank] * after;
  if (plan.out_elems > 0) normalize_job(j);
  plan.before = before;
  plan.after = after;
  plan.numelem.assign(numelem.begin(), numelem.begin() + size);
  return plan;
}

}  // namespace m4t
