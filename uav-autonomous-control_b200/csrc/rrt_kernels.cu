// rrt_kernels.cu -- batched RRT* path planning, fp64, one WARP per mission.
//
// Reference: RRTStar.run (uav_ac/planning/rrt.py:37-79) with its helpers (:120-274).  Tree growth is sequential per
// mission, so a mission owns one warp: the lanes share the O(n) scans of an iteration -- nearest node (:133-138),
// neighbours within 1.5 x step with a collision-free connection (:150-156, slab test :246-274), cost-to-come of every
// neighbour (:168-192) -- and the scalar decisions (steer, insert / re-parent, rewire, best-tree bookkeeping) are taken
// redundantly by all lanes from warp-uniform values.  Nodes are identified by index instead of the reference's
// rounded-coordinate dictionary key; re-parenting an existing key updates its parent in place, which is equivalent to the
// reference's duplicate append (oracle/rrt_np.py documents why).  Random numbers: Philox4x32-10 keyed by
// (seed, global mission index), counter = iteration, so a mission's tree does not depend on batch size or sharding, and
// oracle/rrt_np.py -- pinned on the reference itself by feeding it the same numbers -- reproduces the kernel bit for bit:
// all geometry uses separately rounded fp64 operations in one fixed order.
#include "philox.cuh"
#include "uavb_common.cuh"

namespace uavb {

constexpr int kRrtWarpsPerCta = 1;      // one mission per CTA: a finished mission frees its slot at once (tree sizes vary by an order of magnitude)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double round2(double x) { return __ddiv_rn(rint(__dmul_rn(x, 100.0)), 100.0); }   // np.round(x, 2)

__device__ __forceinline__ double dist3(const double* a, const double* b) {
  const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// Exact segment vs AABB intersection, slab method (rrt.py:246-274).
__device__ __forceinline__ bool segment_hits_cuboid(const double* p, const double* q, const double* box) {
  double t_min = 0.0, t_max = 1.0;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const double d = q[ax] - p[ax], lo = box[2 * ax], hi = box[2 * ax + 1];
    if (fabs(d) < 1e-12) {
      if (p[ax] < lo || p[ax] > hi) return false;
      continue;
    }
    double t_lo = __ddiv_rn(lo - p[ax], d), t_hi = __ddiv_rn(hi - p[ax], d);
    if (t_lo > t_hi) { const double t = t_lo; t_lo = t_hi; t_hi = t; }
    t_min = fmax(t_min, t_lo);
    t_max = fmin(t_max, t_hi);
    if (t_min > t_max) return false;
  }
  return true;
}

__device__ __forceinline__ bool valid_connection(const double* p, const double* q, const double* obstacles, int n_obs) {
  for (int k = 0; k < n_obs; ++k)
    if (segment_hits_cuboid(p, q, obstacles + 6 * k)) return false;
  return true;
}

struct RrtTree {
  double* nodes;   // [cap][3]
  int* parent;     // [cap]
  __device__ __forceinline__ double cost_to_come(int i) const {       // rrt.py:168-179: edge lengths summed towards the start
    double c = 0.0;
    while (i != 0) {
      const int p = parent[i];
      c = __dadd_rn(c, dist3(nodes + 3 * i, nodes + 3 * p));
      i = p;
    }
    return c;
  }
};

struct RrtArgs {
  const double* limits;      // [2][3]
  const double* start;       // [B][3]
  const double* goal;        // [B][3]
  const double* obstacles;   // [n_obs][6]
  double step, epsilon;
  int n_obs, max_iterations, B, max_path;
  unsigned long long seed;
  long long index_base;
  double* nodes; int* parent; int* best_parent; int* nb; double* scratch;   // workspace, each [B][cap] (nodes x3)
  double* path_out; int* path_len_out; double* simple_out; int* simple_len_out; double* cost_out; int* status_out; int* stats_out;
};

__global__ void __launch_bounds__(32 * kRrtWarpsPerCta) rrt_star_kernel(const RrtArgs a) {
  const int lane = threadIdx.x & 31;
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= a.B) return;
  const int cap = a.max_iterations + 1;
  RrtTree t{a.nodes + (size_t)m * cap * 3, a.parent + (size_t)m * cap};
  int* best_parent = a.best_parent + (size_t)m * cap;
  int* nb = a.nb + (size_t)m * cap;
  const double radius = 1.5 * a.step;
  const unsigned long long gi = (unsigned long long)(a.index_base + m);
  const Philox ph{(unsigned)a.seed, (unsigned)(a.seed >> 32)};
  double lw[3], up[3], start[3], goal[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    lw[k] = a.limits[k]; up[k] = a.limits[3 + k];
    start[k] = round2(a.start[3 * m + k]); goal[k] = round2(a.goal[3 * m + k]);      // rrt.py:14-15
  }
  if (lane == 0) { t.nodes[0] = start[0]; t.nodes[1] = start[1]; t.nodes[2] = start[2]; t.parent[0] = -1; }
  __syncwarp();
  int n = 1, goal_idx = -1, best_n = 0, stall = 0, used = 0;
  double best_cost = 1.0 / 0.0;
  bool has_best = false;
  const double stall_limit = (double)a.max_iterations / 10.0;                       // rrt.py:29

  for (int it = 0; it < a.max_iterations; ++it) {
    used = it + 1;
    // ---- sample (rrt.py:120-131)
    unsigned r0[4], r1[4];
    ph.block((unsigned)gi, (unsigned)(gi >> 32), 0x5252u, (unsigned)(2 * it), r0);
    ph.block((unsigned)gi, (unsigned)(gi >> 32), 0x5252u, (unsigned)(2 * it + 1), r1);
    double nw[3];
    if (u01d(r0[0], r0[1]) < a.epsilon) {
      nw[0] = goal[0]; nw[1] = goal[1]; nw[2] = goal[2];
    } else {
      const double u[3] = {u01d(r0[2], r0[3]), u01d(r1[0], r1[1]), u01d(r1[2], r1[3])};
#pragma unroll
      for (int k = 0; k < 3; ++k) nw[k] = round2(__dadd_rn(lw[k], __dmul_rn(up[k] - lw[k], u[k])));
    }
    // ---- nearest node (rrt.py:133-138): first index of the minimum distance
    double dmin = 1.0 / 0.0;
    int imin = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      const double d = dist3(nw, t.nodes + 3 * i);
      if (d < dmin) { dmin = d; imin = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(kFull, dmin, off);
      const int oi = __shfl_xor_sync(kFull, imin, off);
      if (od < dmin || (od == dmin && oi < imin)) { dmin = od; imin = oi; }
    }
    // ---- steer (rrt.py:140-148)
    if (dmin > a.step) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double nk = t.nodes[3 * imin + k];
        nw[k] = round2(__dadd_rn(nk, __ddiv_rn(__dmul_rn(nw[k] - nk, a.step), dmin)));
      }
    }
    // ---- neighbours in node order (rrt.py:150-156) and the index of a node with the same coordinates, if any
    int count = 0, existing = -1;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      bool in = false, same = false;
      if (i < n) {
        const double* q = t.nodes + 3 * i;
        same = q[0] == nw[0] && q[1] == nw[1] && q[2] == nw[2];
        in = dist3(q, nw) <= radius && valid_connection(q, nw, a.obstacles, a.n_obs);
      }
      const unsigned bal = __ballot_sync(kFull, in), eq = __ballot_sync(kFull, same);
      if (in) nb[count + __popc(bal & ((1u << lane) - 1u))] = i;
      count += __popc(bal);
      if (eq && existing < 0) existing = base + __ffs(eq) - 1;
    }
    __syncwarp();
    if (count == 0) continue;
    // ---- best neighbour (rrt.py:181-192): lowest cost-to-come + edge, first in neighbour order
    double cmin = 1.0 / 0.0;
    int jmin = 0x7fffffff;
    for (int j = lane; j < count; j += 32) {
      const int i = nb[j];
      const double c = __dadd_rn(t.cost_to_come(i), dist3(t.nodes + 3 * i, nw));
      if (c < cmin) { cmin = c; jmin = j; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double oc = __shfl_xor_sync(kFull, cmin, off);
      const int oj = __shfl_xor_sync(kFull, jmin, off);
      if (oc < cmin || (oc == cmin && oj < jmin)) { cmin = oc; jmin = oj; }
    }
    const int best = nb[jmin];
    // ---- link the new node (rrt.py:194-213); every lane takes the same decision
    int idx;
    const double* bq = t.nodes + 3 * best;
    if (!(bq[0] == nw[0] && bq[1] == nw[1] && bq[2] == nw[2])) {
      if (existing > 0) {
        const double cur = t.cost_to_come(existing), cand = __dadd_rn(t.cost_to_come(best), dist3(nw, bq));
        if (!(cur <= cand) && lane == 0) t.parent[existing] = best;
        idx = existing;
      } else if (existing == 0) {
        idx = 0;                               // a sample equal to the start: the reference fails on it; skipped
      } else {
        idx = n;
        if (lane == 0) { t.nodes[3 * n] = nw[0]; t.nodes[3 * n + 1] = nw[1]; t.nodes[3 * n + 2] = nw[2]; t.parent[n] = best; }
        if (goal_idx < 0 && nw[0] == goal[0] && nw[1] == goal[1] && nw[2] == goal[2]) goal_idx = n;
        ++n;
      }
    } else {
      idx = existing;
    }
    __syncwarp();
    if (idx <= 0) continue;
    // ---- rewire (rrt.py:215-241).  The reference walks the neighbours in order and an accepted rewire changes the
    // cost-to-come of later neighbours that descend from it, so: evaluate a chunk in parallel, accept the FIRST
    // improving neighbour, restart behind it.
    const double new_cost = t.cost_to_come(idx);
    const int new_parent = t.parent[idx];
    bool rewired = false;
    int j0 = 0;
    while (j0 < count) {
      const int j = j0 + lane;
      bool better = false;
      if (j < count) {
        const int i = nb[j];
        if (i != 0 && i != new_parent) better = __dadd_rn(new_cost, dist3(t.nodes + 3 * i, nw)) < t.cost_to_come(i);
      }
      const unsigned bal = __ballot_sync(kFull, better);
      if (bal == 0) { j0 += 32; continue; }
      const int first = __ffs(bal) - 1;
      if (lane == first) t.parent[nb[j]] = idx;
      rewired = true;
      j0 += first + 1;
      __syncwarp();
    }
    // ---- best-tree bookkeeping (rrt.py:53-72)
    if (goal_idx > 0) {
      // path cost summed from the start like RRTStar.path_cost over the start->goal path (:86-93): record the chain, then add
      int len = 0;
      double cost = 0.0;
      if (lane == 0) {
        for (int i = goal_idx; i != 0; i = t.parent[i]) nb[len++] = i;
        int prev = 0;
        for (int k = len - 1; k >= 0; --k) { cost = __dadd_rn(cost, dist3(t.nodes + 3 * nb[k], t.nodes + 3 * prev)); prev = nb[k]; }
      }
      cost = __shfl_sync(kFull, cost, 0);
      (void)rewired;                           // the reference's "cost increased after rewiring" is a sanity check that cannot fire
      if (cost < best_cost) {
        for (int i = lane; i < n; i += 32) best_parent[i] = t.parent[i];
        best_n = n; best_cost = cost; stall = 0; has_best = true;
      } else {
        ++stall;
      }
      __syncwarp();
      if ((double)stall >= stall_limit) break;
    }
  }

  // ---- result: path of the best tree, start -> goal, and its greedy simplification (rrt.py:97-118)
  if (lane == 0) {
    int status = has_best ? 0 : 1;             // 1 = "No path found" (rrt.py:74-75)
    int len = 0;
    double cost = 0.0;
    if (has_best) {
      for (int i = goal_idx; i != 0; i = best_parent[i]) nb[len++] = i;
      nb[len++] = 0;
      if (len > a.max_path) status = 2;        // path longer than the output buffer
      double* out = a.path_out + (size_t)m * a.max_path * 3;
      for (int k = 0; k < len && k < a.max_path; ++k) {
        const double* q = t.nodes + 3 * nb[len - 1 - k];
        out[3 * k] = q[0]; out[3 * k + 1] = q[1]; out[3 * k + 2] = q[2];
        if (k > 0) cost = __dadd_rn(cost, dist3(q, out + 3 * (k - 1)));
      }
      if (a.simple_out && status == 0) {
        double* so = a.simple_out + (size_t)m * a.max_path * 3;
        int cur = 0, ns = 0;
        so[0] = out[0]; so[1] = out[1]; so[2] = out[2];
        ns = 1;
        while (cur < len - 1) {
          int nxt = len - 1;
          while (nxt > cur + 1 && !valid_connection(out + 3 * cur, out + 3 * nxt, a.obstacles, a.n_obs)) --nxt;
          so[3 * ns] = out[3 * nxt]; so[3 * ns + 1] = out[3 * nxt + 1]; so[3 * ns + 2] = out[3 * nxt + 2];
          ++ns;
          cur = nxt;
        }
        a.simple_len_out[m] = ns;
      } else if (a.simple_len_out) {
        a.simple_len_out[m] = 0;
      }
    } else if (a.simple_len_out) {
      a.simple_len_out[m] = 0;
    }
    a.path_len_out[m] = has_best ? (len < a.max_path ? len : a.max_path) : 0;
    a.cost_out[m] = has_best ? cost : 1.0 / 0.0;
    a.status_out[m] = status;
    if (a.stats_out) { a.stats_out[2 * m] = used; a.stats_out[2 * m + 1] = best_n; }
  }
}

// Batched slab test: hit[i] = segment p[i] -> q[i] intersects ANY of the boxes (RRTStar._is_valid_connection is its negation).
__global__ void __launch_bounds__(128) segment_hits_kernel(const double* __restrict__ p, const double* __restrict__ q, int n,
                                                           const double* __restrict__ boxes, int n_obs, int* __restrict__ hit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  hit[i] = valid_connection(p + 3 * i, q + 3 * i, boxes, n_obs) ? 0 : 1;
}

}  // namespace uavb

using namespace uavb;

extern "C" long long uavb_rrt_workspace_bytes(int B, int max_iterations) {
  if (B < 0 || max_iterations < 1) return -1;
  const long long cap = (long long)max_iterations + 1;
  return (long long)B * cap * (3 * 8 + 4 + 4 + 4 + 8);
}

extern "C" int uavb_rrt_star_f64(const double* space_limits, const double* start, const double* goal, int B, double max_distance,
                                 int max_iterations, const double* obstacles, int n_obs, unsigned long long seed, long long index_base,
                                 void* workspace, double* path_out, int max_path, int* path_len_out, double* simple_path_out,
                                 int* simple_len_out, double* cost_out, int* status_out, int* stats_out, void* stream) {
  UAVB_REQUIRE(space_limits && start && goal && workspace && path_out && path_len_out && cost_out && status_out, "rrt_star: NULL pointer");
  UAVB_REQUIRE(B >= 0 && max_iterations >= 1 && max_path >= 2 && max_distance > 0.0, "rrt_star: B >= 0, max_iterations >= 1, max_path >= 2, max_distance > 0 required");
  UAVB_REQUIRE(n_obs >= 0 && (n_obs == 0 || obstacles), "rrt_star: n_obs > 0 needs obstacles");
  UAVB_REQUIRE((simple_path_out == nullptr) == (simple_len_out == nullptr), "rrt_star: simple_path_out and simple_len_out go together");
  int rc = require_device();
  if (rc) return rc;
  if (B == 0) return UAVB_OK;
  const size_t cap = (size_t)max_iterations + 1;
  RrtArgs a;
  a.limits = space_limits; a.start = start; a.goal = goal; a.obstacles = obstacles;
  a.step = max_distance; a.epsilon = 0.15;                                        // rrt.py:19
  a.n_obs = n_obs; a.max_iterations = max_iterations; a.B = B; a.max_path = max_path;
  a.seed = seed; a.index_base = index_base;
  char* w = static_cast<char*>(workspace);
  a.nodes = reinterpret_cast<double*>(w); w += (size_t)B * cap * 24;
  a.scratch = reinterpret_cast<double*>(w); w += (size_t)B * cap * 8;
  a.parent = reinterpret_cast<int*>(w); w += (size_t)B * cap * 4;
  a.best_parent = reinterpret_cast<int*>(w); w += (size_t)B * cap * 4;
  a.nb = reinterpret_cast<int*>(w);
  a.path_out = path_out; a.path_len_out = path_len_out; a.simple_out = simple_path_out; a.simple_len_out = simple_len_out;
  a.cost_out = cost_out; a.status_out = status_out; a.stats_out = stats_out;
  const int threads = 32 * kRrtWarpsPerCta;
  rrt_star_kernel<<<div_up((long long)B * 32, threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}

extern "C" int uavb_segments_hit_aabbs_f64(const double* p, const double* q, int n, const double* boxes, int n_obs, int* hit_out, void* stream) {
  UAVB_REQUIRE(p && q && hit_out, "segments_hit_aabbs: NULL pointer");
  UAVB_REQUIRE(n >= 0 && n_obs >= 0 && (n_obs == 0 || boxes), "segments_hit_aabbs: n >= 0, n_obs >= 0 (with boxes) required");
  int rc = require_device();
  if (rc) return rc;
  if (n == 0) return UAVB_OK;
  segment_hits_kernel<<<div_up(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p, q, n, boxes, n_obs, hit_out);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}
