/*
 * batch_prep.h — host-side preparation of one scoring batch (pure C++, no CUDA):
 * validate the flat tree, convert it to the device layout, derive ring sizes from
 * the configs and build the config-sorted work list.  Used by capi.cu and by the
 * CPU build of the kernel under tests/emu.
 */
#ifndef REMY_BATCH_PREP_H
#define REMY_BATCH_PREP_H

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "device_layout.h"

namespace remy {

inline uint32_t next_pow2(uint64_t v)
{
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return (uint32_t)std::min<uint64_t>(p, 1u << 30);
}

struct PreparedTree {
  std::vector<Bounds> bounds;     /* [n_nodes][NAX] */
  std::vector<DevNodeMeta> meta;  /* [n_nodes] */
  std::vector<DevLeaf> leaves;    /* [n_leaves] */
  std::vector<double> cuts;       /* per interior node with a table: its cuts, axis after axis */
  std::vector<uint16_t> table;    /* per interior node with a table: first matching child per cell */
  uint32_t axes_mask = 0;
  size_t bytes() const
  {
    return bounds.size() * sizeof(Bounds) + meta.size() * sizeof(DevNodeMeta) + leaves.size() * sizeof(DevLeaf)
           + cuts.size() * sizeof(double) + ((table.size() * sizeof(uint16_t) + 15) / 16) * 16;
  }
};

constexpr size_t MAX_TABLE_CELLS = 4096;   /* per node; larger nodes keep the linear scan */
constexpr size_t MAX_TABLE_TOTAL = 32768;  /* whole tree (64 KB of uint16) */

/* Evaluate the reference's first-match child scan once per cell of the grid induced by the
 * children's bounds (see DevNodeMeta).  For cell index p on an axis with sorted cuts c_0 < ... :
 *   m >= lo  <=>  p > rank(lo)      m < hi  <=>  p <= rank(hi)
 * where rank(v) is v's index among the cuts (-1 for -inf, ncuts for +inf). */
inline void build_cell_table(PreparedTree &t, uint32_t node)
{
  DevNodeMeta &md = t.meta[node];
  std::vector<double> cuts[NAX];
  size_t lo_skip[NAX] = {0}, n_int[NAX] = {0};
  size_t cells = 1;
  for (int a = 0; a < NAX; a++) {
    if (!(t.axes_mask & (1u << a))) continue;
    for (uint32_t c = 0; c < md.count; c++) {
      const Bounds &b = t.bounds[(size_t)(md.first + c) * NAX + a];
      if (b.lo != b.lo || b.hi != b.hi) return;           /* NaN bound: keep the linear scan */
      if (std::isfinite(b.lo)) cuts[a].push_back(b.lo);
      if (std::isfinite(b.hi)) cuts[a].push_back(b.hi);
    }
    std::sort(cuts[a].begin(), cuts[a].end());
    cuts[a].erase(std::unique(cuts[a].begin(), cuts[a].end()), cuts[a].end());
    /* A query that reaches this node lies inside the node's own box (the descent only
     * enters boxes that contain it), so cuts <= lower always count and cuts >= upper never
     * do: only the interior cuts need a comparison on the device. */
    const Bounds &pb = t.bounds[(size_t)node * NAX + a];
    if (pb.lo != pb.lo || pb.hi != pb.hi) return;
    size_t lo_n = 0, hi_n = 0;
    while (lo_n < cuts[a].size() && cuts[a][lo_n] <= pb.lo) lo_n++;
    while (hi_n < cuts[a].size() - lo_n && cuts[a][cuts[a].size() - 1 - hi_n] >= pb.hi) hi_n++;
    lo_skip[a] = lo_n; n_int[a] = cuts[a].size() - lo_n - hi_n;
    if (n_int[a] > 255) return;
    cells *= n_int[a] + 1;
    if (cells > MAX_TABLE_CELLS) return;
  }
  if (md.count >= NO_CELL_CHILD || t.table.size() + cells > MAX_TABLE_TOTAL) return;
  auto rank = [&](int a, double v) -> int {
    if (v == -INFINITY) return -1;
    if (v == INFINITY) return (int)cuts[a].size();
    return (int)(std::lower_bound(cuts[a].begin(), cuts[a].end(), v) - cuts[a].begin());
  };
  md.table_off = (uint32_t)t.table.size();
  md.cuts_off = (uint32_t)t.cuts.size();
  bool bisected = t.axes_mask == 0xFu;
  for (int a = 0; a < NAX; a++) {
    md.ncuts[a] = (uint8_t)n_int[a];
    if ((t.axes_mask & (1u << a)) && n_int[a] != 1) bisected = false;
    t.cuts.insert(t.cuts.end(), cuts[a].begin() + lo_skip[a], cuts[a].begin() + lo_skip[a] + n_int[a]);
  }
  md.ncuts[7] = bisected ? 1 : 0; /* one interior cut on each of the four default axes: the kernel's fast path */
  t.table.resize(t.table.size() + cells, NO_CELL_CHILD);
  std::vector<int> p(NAX, 0);
  for (size_t cell = 0; cell < cells; cell++) {
    size_t rem = cell;
    for (int a = 0; a < NAX; a++) {
      const size_t dim = (t.axes_mask & (1u << a)) ? n_int[a] + 1 : 1;
      p[a] = (int)(rem % dim) + (int)lo_skip[a]; /* number of cuts <= m on this axis */
      rem /= dim;
    }
    for (uint32_t c = 0; c < md.count; c++) {
      bool ok = true;
      for (int a = 0; a < NAX && ok; a++) {
        if (!(t.axes_mask & (1u << a))) continue;
        const Bounds &b = t.bounds[(size_t)(md.first + c) * NAX + a];
        ok = (p[a] > rank(a, b.lo)) && (p[a] <= rank(a, b.hi));
      }
      if (ok) { t.table[md.table_off + cell] = (uint16_t)c; break; }
    }
  }
}

/* Returns "" or a description of what is wrong with the tree. */
inline std::string prepare_tree(const RemyFlatTree *tree, PreparedTree &out)
{
  if (!tree || !tree->nodes || !tree->leaves || tree->n_nodes == 0 || tree->n_leaves == 0) return "empty tree";
  uint32_t mask = 0;
  for (uint32_t i = 0; i < tree->n_nodes; i++) {
    const RemyTreeNode &nd = tree->nodes[i];
    mask |= nd.axis_mask & ((1u << NAX) - 1u);
    if (nd.child_count) {
      /* children must lie after their parent: the descent then always terminates */
      if ((uint64_t)nd.child_begin + nd.child_count > tree->n_nodes || nd.child_begin <= i)
        return "tree node " + std::to_string(i) + " has an invalid child range";
    } else if (nd.leaf_index >= tree->n_leaves) {
      return "tree node " + std::to_string(i) + " has an invalid leaf index";
    }
  }
  out.axes_mask = mask;
  out.bounds.resize((size_t)tree->n_nodes * NAX);
  out.meta.resize(tree->n_nodes);
  out.leaves.resize(tree->n_leaves);
  for (uint32_t i = 0; i < tree->n_nodes; i++) {
    const RemyTreeNode &nd = tree->nodes[i];
    for (int a = 0; a < NAX; a++) {
      Bounds b;
      /* MemoryRange::contains only tests the node's own active axes
       * (reference src/memoryrange.cc:52-58); widening the others lets the kernel
       * test the union of axes for every node. */
      if (nd.axis_mask & (1u << a)) { b.lo = nd.lower[a]; b.hi = nd.upper[a]; }
      else { b.lo = -INFINITY; b.hi = INFINITY; }
      out.bounds[(size_t)i * NAX + a] = b;
    }
    out.meta[i].first = nd.child_count ? nd.child_begin : nd.leaf_index;
    out.meta[i].count = nd.child_count;
    out.meta[i].table_off = NO_TABLE; out.meta[i].cuts_off = 0;
    for (int a = 0; a < 8; a++) out.meta[i].ncuts[a] = 0;
  }
  for (uint32_t i = 0; i < tree->n_nodes; i++)
    if (tree->nodes[i].child_count) build_cell_table(out, i);
  for (uint32_t i = 0; i < tree->n_leaves; i++) {
    out.leaves[i].window_multiple = tree->leaves[i].window_multiple;
    out.leaves[i].intersend = tree->leaves[i].intersend;
    out.leaves[i].window_increment = tree->leaves[i].window_increment;
    out.leaves[i].pad = 0;
  }
  return "";
}

/* One kernel launch: simulations whose num_senders falls in the same bucket.  Shared
 * memory per thread grows with the bucket's largest num_senders, so small-n buckets run
 * with (many) more resident threads per SM than one launch sized for n = 32 would. */
struct PassPlan {
  uint32_t n_max;     /* largest num_senders the pass accepts */
  uint32_t begin, end;/* range of BatchPlan::order */
};

struct BatchPlan {
  uint32_t n_max = 1;       /* largest valid num_senders in the batch */
  uint32_t link_cap = 1;    /* ring entries per slot (power of two) */
  uint32_t delay_cap = 8;
  std::vector<uint32_t> order; /* simulation ids: by bucket, then num_senders, then cost, descending */
  std::vector<PassPlan> passes;
};

constexpr uint32_t LINK_CAP_FIRST_TRY = 16384; /* for limits above this (incl. "infinite"); grown on overflow */

inline uint32_t sender_bucket(uint32_t n)
{
  if (n == 0 || n > REMY_MAX_SENDERS) return 4; /* rejected by the kernel; any pass will do */
  if (n <= 4) return 4;
  if (n <= 8) return 8;
  if (n <= 16) return 16;
  return 32;
}

inline BatchPlan plan_batch(const RemyNetConfig *cfgs, uint32_t n_cfgs, uint32_t n_cands, bool single_pass = false)
{
  BatchPlan p;
  uint64_t lim_max = 0; double bdp_max = 0; bool any_big = false;
  for (uint32_t k = 0; k < n_cfgs; k++) {
    const RemyNetConfig &c = cfgs[k];
    if (c.num_senders >= 1 && c.num_senders <= REMY_MAX_SENDERS) p.n_max = std::max(p.n_max, c.num_senders);
    if (c.buffer_size > LINK_CAP_FIRST_TRY) any_big = true; else lim_max = std::max<uint64_t>(lim_max, c.buffer_size);
    const double bdp = c.link_ppt * c.delay;
    if (c.link_ppt > 0 && c.delay >= 0 && std::isfinite(bdp)) bdp_max = std::max(bdp_max, bdp);
  }
  p.link_cap = next_pow2(std::max<uint64_t>(any_big ? LINK_CAP_FIRST_TRY : 1, std::max<uint64_t>(lim_max, 1) + 1)); /* (+1: the kernel keeps one ring entry in reserve, sim_kernel.cuh REMY_TOP_LOADS) */
  /* the delay line holds at most floor(delay * link_ppt) + 1 packets in flight plus the
   * delivered-but-unread ones of the current tick */
  p.delay_cap = next_pow2((uint64_t)std::min(bdp_max, 1048000.0) + 8);
  const uint32_t n_sims = n_cands * n_cfgs;
  p.order.resize(n_sims);
  for (uint32_t i = 0; i < n_sims; i++) p.order[i] = i;
  std::vector<double> cost(n_cfgs);
  std::vector<uint32_t> bucket(n_cfgs);
  for (uint32_t k = 0; k < n_cfgs; k++) {
    cost[k] = cfgs[k].link_ppt * cfgs[k].simulation_ticks;
    bucket[k] = single_pass ? REMY_MAX_SENDERS : sender_bucket(cfgs[k].num_senders);
  }
  /* instances sorted by config: same num_senders together (same loop trip counts in a
   * warp), longest first (the tail of each pass is made of its shortest simulations) */
  std::stable_sort(p.order.begin(), p.order.end(), [&](uint32_t x, uint32_t y) {
    const uint32_t kx = x % n_cfgs, ky = y % n_cfgs;
    if (bucket[kx] != bucket[ky]) return bucket[kx] > bucket[ky];
    if (cfgs[kx].num_senders != cfgs[ky].num_senders) return cfgs[kx].num_senders > cfgs[ky].num_senders;
    if (cost[kx] != cost[ky]) return cost[kx] > cost[ky];
    return kx < ky;
  });
  for (uint32_t i = 0; i < n_sims;) {
    const uint32_t bk = bucket[p.order[i] % n_cfgs];
    uint32_t j = i;
    while (j < n_sims && bucket[p.order[j] % n_cfgs] == bk) j++;
    PassPlan pp; pp.n_max = single_pass ? std::max(p.n_max, 4u) : bk; pp.begin = i; pp.end = j;
    p.passes.push_back(pp);
    i = j;
  }
  return p;
}

} // namespace remy
#endif
