/*
 * ref_driver.cc — C-ABI shim around the reference's OWN simulator sources.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/remy_oracle.c header).  Compiled by
 * oracle/build_ref.sh together with the reference's unmodified .cc files from
 * /root/reference/src into oracle/_ref/libremy_ref.so (git-ignored).  No
 * reference source is copied into this repository; the only edit is applied by
 * the build script to a build-time copy of memory.cc (the
 * fork's 6-element MAX_MEMORY and a stray assert(false) make the Rat path
 * unrunnable as shipped).
 *
 * Two modes, same output structs as include/remy_b200.h:
 *   mode 0  "white box": builds the Network exactly like
 *           Evaluator<WhiskerTree>::score (src/evaluator.cc:84-98) and steps the
 *           loop of Network::run_simulation (src/network.cc:73-84) itself, so
 *           that loop iterations, packet counters, queue occupancy and the PRNG
 *           end state can be read.  Private members are reached by compiling
 *           the reference headers with `private`/`protected` defined away.
 *   mode 1  "public API": calls the reference's public static
 *           Evaluator<WhiskerTree>::score(tree, seed, {config}, false, ticks)
 *           (src/evaluator.hh:51-55) and fills utility + per-sender
 *           throughput/delay only.  Used to pin mode 0 to the real entry point.
 */
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <future>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>
#include <fcntl.h>
#include <unistd.h>
#include <sys/types.h>

#include "dna.pb.h"
#include "answer.pb.h"
#include "problem.pb.h"
#include "simulationresults.pb.h"
#include <boost/functional/hash.hpp>
#include <boost/accumulators/accumulators.hpp>
#include <Python.h>

#define private public
#define protected public
#include "evaluator.hh"
#include "network.cc"
#include "rat-templates.cc"
#include "fish-templates.cc"
#undef private
#undef protected

#include "../include/remy_b200.h"

namespace {

typedef SenderGang<Rat, TimeSwitchedSender<Rat>> RatGang;
typedef Network<RatGang, RatGang> RatNetwork;
/* The reference instantiates ByteSwitchedSender for its RL sender only (src/unicornevaluator.cc:109,140); the class
 * template itself is generic (src/sendergang.cc:108-138, 256-278), so the checker instantiates it for Rat. */
typedef SenderGang<Rat, ByteSwitchedSender<Rat>> RatFlowGang;
typedef Network<RatFlowGang, RatFlowGang> RatFlowNetwork;

RemyBuffers::Memory memory_dna(const double *v)
{
  RemyBuffers::Memory m;
  m.set_rec_send_ewma(v[0]); m.set_rec_rec_ewma(v[1]); m.set_rtt_ratio(v[2]);
  m.set_slow_rec_rec_ewma(v[3]); m.set_rtt_diff(v[4]); m.set_queueing_delay(v[5]);
  return m;
}

RemyBuffers::MemoryRange range_dna(const RemyTreeNode &nd)
{
  RemyBuffers::MemoryRange r;
  r.mutable_lower()->CopyFrom(memory_dna(nd.lower));
  r.mutable_upper()->CopyFrom(memory_dna(nd.upper));
  for (int a = 0; a < REMY_NUM_AXES; a++)
    if (nd.axis_mask & (1u << a)) r.add_active_axis(a);
  return r;
}

void build_dna(const RemyFlatTree *t, uint32_t node, const RemyLeafOverride *ov, RemyBuffers::WhiskerTree *out)
{
  const RemyTreeNode &nd = t->nodes[node];
  out->mutable_domain()->CopyFrom(range_dna(nd));
  if (nd.child_count == 0) {
    RemyLeafAction a = t->leaves[nd.leaf_index];
    if (ov && ov->leaf_index == nd.leaf_index) {
      a.window_increment = ov->window_increment; a.window_multiple = ov->window_multiple; a.intersend = ov->intersend;
    }
    RemyBuffers::Whisker *w = out->mutable_leaf();
    w->set_window_increment(a.window_increment);
    w->set_window_multiple(a.window_multiple);
    w->set_intersend(a.intersend);
    w->mutable_domain()->CopyFrom(range_dna(nd));
  } else {
    for (uint32_t c = 0; c < nd.child_count; c++) build_dna(t, nd.child_begin + c, ov, out->add_children());
  }
}

/* leaves in depth-first order == RemyTreeNode::leaf_index order */
void collect_counts(const WhiskerTree &t, std::vector<unsigned> &counts)
{
  if (t.is_leaf()) { counts.push_back(t._leaf.front().count()); return; }
  for (const auto &c : t._children) collect_counts(c, counts);
}

NetConfig to_netconfig(const RemyNetConfig &c)
{
  NetConfig x;
  x.mean_on_duration = c.mean_on_duration; x.mean_off_duration = c.mean_off_duration;
  x.num_senders = c.num_senders; x.link_ppt = c.link_ppt; x.delay = c.delay;
  x.buffer_size = c.buffer_size; x.stochastic_loss_rate = c.stochastic_loss_rate;
  x.simulation_ticks = c.simulation_ticks;
  return x;
}

template <class Net>
void run_whitebox_t(WhiskerTree &tree, const RemyNetConfig &c, uint32_t seed,
                    RemySimResult *out, RemySenderResult *snd)
{
  PRNG run_prng(seed);
  const NetConfig x = to_netconfig(c);
  Net net(Rat(tree, false), run_prng, x);
  const double duration = c.simulation_ticks;
  uint64_t iters = 0;
  while (net._tickno < duration) { /* src/network.cc:73-84 */
    net._tickno = std::min(std::min(net._senders.next_event_time(net._tickno),
                                    std::min(net._link.next_event_time(net._tickno),
                                             net._stochastic_loss.next_event_time(net._tickno))),
                           std::min(net._delay.next_event_time(net._tickno),
                                    net._rec.next_event_time(net._tickno)));
    if (net._tickno > duration) break;
    assert(net._tickno < std::numeric_limits<double>::max());
    net.tick();
    iters++;
  }
  memset(out, 0, sizeof *out);
  out->utility = net.senders().utility();
  out->last_tick = net._tickno;
  out->event_ticks = iters;
  const unsigned n = net._senders.gang1_.count_senders();
  for (unsigned i = 0; i < n; i++) {
    const auto &s = net._senders.gang1_._gang[i];
    out->packets_sent += s.sender._packets_sent;
    out->packets_received += s.utility._packets_received;
    if (snd) {
      snd[i].tick_share_sending = s.utility._tick_share_sending;
      snd[i].total_delay = s.utility._total_delay;
      snd[i].packets_received = s.utility._packets_received;
      snd[i].packets_sent = s.sender._packets_sent;
    }
  }
  out->link_queued = (uint32_t)(net._link._buffer.size() + net._link._pending_packet._queue.size());
  uint32_t infl = (uint32_t)net._delay._queue.size();
  for (const auto &v : net._rec._collector) infl += (uint32_t)v.size();
  out->delay_inflight = infl;
  std::ostringstream os; os << run_prng;
  out->prng_state = (uint32_t)std::stoul(os.str());
}

void run_whitebox(WhiskerTree &tree, const RemyNetConfig &c, uint32_t seed, RemySimResult *out, RemySenderResult *snd)
{
  run_whitebox_t<RatNetwork>(tree, c, seed, out, snd);
}

void run_public(WhiskerTree &tree, const RemyNetConfig &c, uint32_t seed,
                RemySimResult *out, RemySenderResult *snd)
{
  std::vector<NetConfig> one { to_netconfig(c) };
  auto outcome = Evaluator<WhiskerTree>::score(tree, seed, one, false, (unsigned int)c.simulation_ticks);
  memset(out, 0, sizeof *out);
  out->utility = outcome.score;
  if (snd) {
    const auto &td = outcome.throughputs_delays.at(0).second;
    for (size_t i = 0; i < td.size(); i++) {
      /* public API exposes only (normalised throughput, average delay) */
      snd[i].tick_share_sending = td[i].first;
      snd[i].total_delay = td[i].second;
      snd[i].packets_received = 0; snd[i].packets_sent = 0;
    }
  }
}

} // namespace

extern "C" int remy_ref_score_batch(const RemyFlatTree *tree,
                                    const RemyLeafOverride *cands, uint32_t n_cands,
                                    const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                    uint32_t sender_stride,
                                    RemySimResult *out_sims, RemySenderResult *out_senders,
                                    uint32_t *out_use_counts, uint32_t n_threads, int mode)
{
  if (!tree || !cfgs || !seeds || !out_sims) return REMY_EINVAL;
  if (!cands) n_cands = 1;
  const uint32_t total = n_cands * n_cfgs;
  if (out_use_counts) memset(out_use_counts, 0, sizeof(uint32_t) * (size_t)n_cands * tree->n_leaves);
  std::atomic<uint32_t> next(0);
  std::mutex mu;
  auto worker = [&]() {
    std::vector<RemySenderResult> tmp(4096);
    for (;;) {
      const uint32_t s = next.fetch_add(1);
      if (s >= total) break;
      const uint32_t c = s / n_cfgs, k = s % n_cfgs;
      const RemyLeafOverride *ov = (cands && cands[c].leaf_index != REMY_NO_OVERRIDE) ? &cands[c] : nullptr;
      RemyBuffers::WhiskerTree dna;
      build_dna(tree, 0, ov, &dna);
      WhiskerTree t(dna);        /* own copy per task, like src/breeder.cc:65 */
      t.reset_counts();          /* src/evaluator.cc:86 */
      if (mode == 0) run_whitebox(t, cfgs[k], seeds[k], &out_sims[s], tmp.data());
      else run_public(t, cfgs[k], seeds[k], &out_sims[s], tmp.data());
      if (out_senders) {
        RemySenderResult *dst = &out_senders[(size_t)s * sender_stride];
        memset(dst, 0, sizeof(RemySenderResult) * sender_stride);
        memcpy(dst, tmp.data(), sizeof(RemySenderResult) * std::min(sender_stride, cfgs[k].num_senders));
      }
      if (out_use_counts) {
        std::vector<unsigned> counts;
        collect_counts(t, counts);
        std::lock_guard<std::mutex> g(mu);
        for (size_t l = 0; l < counts.size() && l < tree->n_leaves; l++)
          out_use_counts[(size_t)c * tree->n_leaves + l] += counts[l];
      }
    }
  };
  if (n_threads <= 1) worker();
  else {
    std::vector<std::thread> th;
    for (uint32_t i = 0; i < n_threads; i++) th.emplace_back(worker);
    for (auto &t : th) t.join();
  }
  return 0;
}

/* The reference's PUBLIC Evaluator<WhiskerTree>::score on SEVERAL configs in one call: one PRNG chained through all of
 * them (src/evaluator.cc:84,92-94).  cfgs[k].simulation_ticks must all be equal (the call takes one run length).
 * out_tput_delay: per config, per sender, (normalised throughput, average delay), sender_stride pairs per config. */
extern "C" int remy_ref_score_chained(const RemyFlatTree *tree, const RemyNetConfig *cfgs, uint32_t n_cfgs, uint32_t prng_seed,
                                      uint32_t sender_stride, double *out_score, double *out_tput_delay, uint32_t *out_use_counts)
{
  if (!tree || !cfgs || !out_score) return REMY_EINVAL;
  RemyBuffers::WhiskerTree dna;
  build_dna(tree, 0, nullptr, &dna);
  WhiskerTree t(dna);
  std::vector<NetConfig> configs;
  for (uint32_t k = 0; k < n_cfgs; k++) configs.push_back(to_netconfig(cfgs[k]));
  auto outcome = Evaluator<WhiskerTree>::score(t, prng_seed, configs, false, (unsigned int)cfgs[0].simulation_ticks);
  *out_score = outcome.score;
  if (out_tput_delay)
    for (uint32_t k = 0; k < n_cfgs; k++) {
      const auto &td = outcome.throughputs_delays.at(k).second;
      for (size_t i = 0; i < td.size() && i < sender_stride; i++) {
        out_tput_delay[((size_t)k * sender_stride + i) * 2] = td[i].first;
        out_tput_delay[((size_t)k * sender_stride + i) * 2 + 1] = td[i].second;
      }
    }
  if (out_use_counts) {
    std::vector<unsigned> counts;
    collect_counts(outcome.used_actions, counts);
    for (size_t l = 0; l < counts.size() && l < tree->n_leaves; l++) out_use_counts[l] = counts[l];
  }
  return 0;
}

/* Rat senders behind the reference's ByteSwitchedSender: cfgs[k].mean_on_duration is the mean flow length in packets
 * (<= 0: unbounded flows).  White-box loop only: no public entry point of the reference builds this network. */
extern "C" int remy_ref_score_batch_flows(const RemyFlatTree *tree,
                                          const RemyLeafOverride *cands, uint32_t n_cands,
                                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                          uint32_t sender_stride,
                                          RemySimResult *out_sims, RemySenderResult *out_senders,
                                          uint32_t *out_use_counts, uint32_t n_threads)
{
  if (!tree || !cfgs || !seeds || !out_sims) return REMY_EINVAL;
  if (!cands) n_cands = 1;
  const uint32_t total = n_cands * n_cfgs;
  if (out_use_counts) memset(out_use_counts, 0, sizeof(uint32_t) * (size_t)n_cands * tree->n_leaves);
  std::atomic<uint32_t> next(0);
  std::mutex mu;
  auto worker = [&]() {
    std::vector<RemySenderResult> tmp(4096);
    for (;;) {
      const uint32_t s = next.fetch_add(1);
      if (s >= total) break;
      const uint32_t c = s / n_cfgs, k = s % n_cfgs;
      const RemyLeafOverride *ov = (cands && cands[c].leaf_index != REMY_NO_OVERRIDE) ? &cands[c] : nullptr;
      RemyBuffers::WhiskerTree dna;
      build_dna(tree, 0, ov, &dna);
      WhiskerTree t(dna);
      t.reset_counts();
      run_whitebox_t<RatFlowNetwork>(t, cfgs[k], seeds[k], &out_sims[s], tmp.data());
      if (out_senders) {
        RemySenderResult *dst = &out_senders[(size_t)s * sender_stride];
        memset(dst, 0, sizeof(RemySenderResult) * sender_stride);
        memcpy(dst, tmp.data(), sizeof(RemySenderResult) * std::min(sender_stride, cfgs[k].num_senders));
      }
      if (out_use_counts) {
        std::vector<unsigned> counts;
        collect_counts(t, counts);
        std::lock_guard<std::mutex> g(mu);
        for (size_t l = 0; l < counts.size() && l < tree->n_leaves; l++)
          out_use_counts[(size_t)c * tree->n_leaves + l] += counts[l];
      }
    }
  };
  if (n_threads <= 1) worker();
  else {
    std::vector<std::thread> th;
    for (uint32_t i = 0; i < n_threads; i++) th.emplace_back(worker);
    for (auto &t : th) t.join();
  }
  return 0;
}

/* The reference's own fan-out SHAPE (src/breeder.cc:57-69): one std::async(std::launch::async, ...) task per
 * candidate - an OS thread each, no pool - that copies the tree, applies its replacement and scores ALL configs
 * serially; the futures are drained in submission order (src/breeder.cc:73-77, 130-134).  Per-simulation seeds as
 * everywhere in this repository.  Wall time = the slowest candidate on however many cores exist. */
extern "C" int remy_ref_score_batch_async(const RemyFlatTree *tree,
                                          const RemyLeafOverride *cands, uint32_t n_cands,
                                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                          RemySimResult *out_sims, int mode)
{
  if (!tree || !cfgs || !seeds || !out_sims) return REMY_EINVAL;
  if (!cands) n_cands = 1;
  std::vector<std::future<double>> scores;
  for (uint32_t c = 0; c < n_cands; c++) {
    scores.emplace_back(std::async(std::launch::async, [=]() {
      const RemyLeafOverride *ov = (cands && cands[c].leaf_index != REMY_NO_OVERRIDE) ? &cands[c] : nullptr;
      RemyBuffers::WhiskerTree dna;
      build_dna(tree, 0, ov, &dna);
      WhiskerTree t(dna);
      t.reset_counts();
      std::vector<RemySenderResult> tmp(4096);
      double score = 0;
      for (uint32_t k = 0; k < n_cfgs; k++) {
        RemySimResult *out = &out_sims[(size_t)c * n_cfgs + k];
        if (mode == 0) run_whitebox(t, cfgs[k], seeds[k], out, tmp.data());
        else run_public(t, cfgs[k], seeds[k], out, tmp.data());
        score += out->utility; /* src/evaluator.cc:96 */
      }
      return score;
    }));
  }
  for (auto &f : scores) f.get();
  return 0;
}

/* libstdc++/glibc arithmetic the reference relies on, exposed so the tests can
 * pin the oracle's restatements against the real library code. */
extern "C" double remy_ref_exponential(uint32_t *state, double rate)
{
  PRNG g(*state);
  Exponential e(rate);
  const double r = e.sample(g);
  std::ostringstream os; os << g; *state = (uint32_t)std::stoul(os.str());
  return r;
}
extern "C" int remy_ref_bernoulli(uint32_t *state, double p)
{
  PRNG g(*state);
  std::bernoulli_distribution d(p);
  const int r = d(g);
  std::ostringstream os; os << g; *state = (uint32_t)std::stoul(os.str());
  return r;
}
extern "C" void remy_ref_shuffle(uint32_t *state, uint32_t *idx, uint32_t n)
{
  PRNG g(*state);
  std::vector<unsigned int> v(idx, idx + n);
  std::shuffle(v.begin(), v.end(), g);
  std::copy(v.begin(), v.end(), idx);
  std::ostringstream os; os << g; *state = (uint32_t)std::stoul(os.str());
}
extern "C" double remy_ref_log(double x) { return std::log(x); }

/* ---- trace == true: per-leaf running medians, and Breeder::apply_best_split ------------- *
 * The reference's own Evaluator::score(..., trace = true, ...), MemoryRange::track / bisect,
 * WhiskerTree::most_used, WhiskerTree(whisker, bisect = true) and replace(src, dst) run here;
 * only the 20-line control flow of Breeder::apply_best_split (src/breeder.cc:15-41, dead in
 * the fork because of its leading assert(false)) is re-stated around them.  Scoring follows
 * the batched evaluator's contract: the static score() once per config with seeds[k], on ONE
 * tree object, so every leaf's accumulators see config 0's lookups, then config 1's, ...;
 * use counts are summed over the calls (each call starts with reset_counts()). */
namespace {

void collect_leaves(WhiskerTree &t, std::vector<Whisker *> &out)
{
  if (t.is_leaf()) { out.push_back(&t._leaf.front()); return; }
  for (auto &c : t._children) collect_leaves(c, out);
}

void flatten_tree(const WhiskerTree &t, std::vector<RemyTreeNode> &nodes, std::vector<RemyLeafAction> &leaves, uint32_t self)
{
  RemyTreeNode nd; memset(&nd, 0, sizeof nd);
  for (unsigned a = 0; a < REMY_NUM_AXES; a++) { nd.lower[a] = t._domain._lower.field(a); nd.upper[a] = t._domain._upper.field(a); }
  for (auto a : t._domain._active_axis) nd.axis_mask |= 1u << (unsigned)a;
  if (t.is_leaf()) {
    const Whisker &w = t._leaf.front();
    nd.child_count = 0; nd.leaf_index = (uint32_t)leaves.size();
    RemyLeafAction la; la.window_multiple = w._window_multiple; la.intersend = w._intersend;
    la.window_increment = w._window_increment; la.generation = w._generation;
    leaves.push_back(la);
    nodes[self] = nd;
    return;
  }
  nd.child_begin = (uint32_t)nodes.size(); nd.child_count = (uint32_t)t._children.size();
  nodes[self] = nd;
  nodes.resize(nodes.size() + t._children.size());
  for (size_t c = 0; c < t._children.size(); c++) flatten_tree(t._children[c], nodes, leaves, nd.child_begin + (uint32_t)c);
}

} // namespace

extern "C" int remy_ref_trace_split(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                    uint32_t generation, int do_split,
                                    double *out_medians /* [n_leaves][REMY_NUM_AXES] */, uint32_t *out_use_counts /* [n_leaves] */,
                                    double *out_score,
                                    RemyTreeNode *out_nodes, uint32_t node_cap, RemyLeafAction *out_leaves, uint32_t leaf_cap,
                                    uint32_t *out_n_nodes, uint32_t *out_n_leaves)
{
  RemyBuffers::WhiskerTree dna;
  build_dna(tree, 0, nullptr, &dna);
  WhiskerTree t(dna);
  {
    std::vector<Whisker *> lv; collect_leaves(t, lv);
    for (size_t l = 0; l < lv.size() && l < tree->n_leaves; l++) lv[l]->_generation = tree->leaves[l].generation;
  }
  std::vector<unsigned> sums(tree->n_leaves, 0);
  double score = 0;
  for (uint32_t k = 0; k < n_cfgs; k++) {
    std::vector<NetConfig> one { to_netconfig(cfgs[k]) };
    auto outcome = Evaluator<WhiskerTree>::score(t, seeds[k], one, true, (unsigned int)cfgs[k].simulation_ticks);
    score += outcome.score;
    std::vector<unsigned> counts; collect_counts(t, counts);
    for (size_t l = 0; l < counts.size(); l++) sums[l] += counts[l];
  }
  std::vector<Whisker *> lv; collect_leaves(t, lv);
  for (size_t l = 0; l < lv.size(); l++) {
    lv[l]->_domain._count = sums[l];
    if (out_use_counts) out_use_counts[l] = sums[l];
    if (out_medians)
      for (unsigned a = 0; a < REMY_NUM_AXES; a++) out_medians[l * REMY_NUM_AXES + a] = boost::accumulators::median(lv[l]->_domain._acc[a]);
  }
  if (out_score) *out_score = score;
  if (do_split) { /* src/breeder.cc:21-40 with outcome.used_actions == a copy of the scored tree */
    WhiskerTree used_actions(t);
    while (1) {
      auto my_action(used_actions.most_used(generation));
      if (!my_action) return -2;
      WhiskerTree bisected_action(*my_action, true);
      if (bisected_action.num_children() == 1) {
        auto mutable_action(*my_action);
        mutable_action.promote(generation + 1);
        if (!used_actions.replace(mutable_action)) return -3;
        continue;
      }
      if (!t.replace(*my_action, bisected_action)) return -4;
      break;
    }
  }
  std::vector<RemyTreeNode> nodes(1); std::vector<RemyLeafAction> leaves;
  flatten_tree(t, nodes, leaves, 0);
  if (out_n_nodes) *out_n_nodes = (uint32_t)nodes.size();
  if (out_n_leaves) *out_n_leaves = (uint32_t)leaves.size();
  if (out_nodes && out_leaves) {
    if (nodes.size() > node_cap || leaves.size() > leaf_cap) return REMY_EINVAL;
    memcpy(out_nodes, nodes.data(), nodes.size() * sizeof(RemyTreeNode));
    memcpy(out_leaves, leaves.data(), leaves.size() * sizeof(RemyLeafAction));
  }
  return 0;
}

/* ---- second sender family: Evaluator<FinTree> / Fish (src/evaluator.cc:105-132) ------------ *
 * Same shapes as remy_ref_score_batch; a leaf's lambda travels in RemyLeafAction.intersend
 * (RemyLeafOverride.intersend for candidates).  mode 0 steps the reference's own
 * Network<SenderGang<Fish, TimeSwitchedSender<Fish>>, ...> exactly like the public score() does
 * (Fish seed = first draw of the run PRNG, :113) so that counters and the PRNG end state can be read;
 * mode 1 calls the public static score(). */

namespace {

typedef SenderGang<Fish, TimeSwitchedSender<Fish>> FishGang;
typedef Network<FishGang, FishGang> FishNetwork;

void build_fin_dna(const RemyFlatTree *t, uint32_t node, const RemyLeafOverride *ov, RemyBuffers::FinTree *out)
{
  const RemyTreeNode &nd = t->nodes[node];
  out->mutable_domain()->CopyFrom(range_dna(nd));
  if (nd.child_count == 0) {
    double lambda = t->leaves[nd.leaf_index].intersend;
    if (ov && ov->leaf_index == nd.leaf_index) lambda = ov->intersend;
    RemyBuffers::Fin *f = out->mutable_leaf();
    f->set_lambda(lambda);
    f->mutable_domain()->CopyFrom(range_dna(nd));
  } else {
    for (uint32_t c = 0; c < nd.child_count; c++) build_fin_dna(t, nd.child_begin + c, ov, out->add_children());
  }
}

void collect_fin_counts(const FinTree &t, std::vector<unsigned> &counts)
{
  if (t.is_leaf()) { counts.push_back(t._leaf.front().count()); return; }
  for (const auto &c : t._children) collect_fin_counts(c, counts);
}

void run_fish_whitebox(FinTree &tree, const RemyNetConfig &c, uint32_t seed, RemySimResult *out, RemySenderResult *snd)
{
  PRNG run_prng(seed);
  unsigned int fish_prng_seed(run_prng());
  const NetConfig x = to_netconfig(c);
  FishNetwork net(Fish(tree, fish_prng_seed, false), run_prng, x);
  const double duration = c.simulation_ticks;
  uint64_t iters = 0;
  while (net._tickno < duration) { /* src/network.cc:73-84 */
    net._tickno = std::min(std::min(net._senders.next_event_time(net._tickno),
                                    std::min(net._link.next_event_time(net._tickno),
                                             net._stochastic_loss.next_event_time(net._tickno))),
                           std::min(net._delay.next_event_time(net._tickno),
                                    net._rec.next_event_time(net._tickno)));
    if (net._tickno > duration) break;
    assert(net._tickno < std::numeric_limits<double>::max());
    net.tick();
    iters++;
  }
  memset(out, 0, sizeof *out);
  out->utility = net.senders().utility();
  out->last_tick = net._tickno;
  out->event_ticks = iters;
  const unsigned n = net._senders.gang1_.count_senders();
  for (unsigned i = 0; i < n; i++) {
    const auto &s = net._senders.gang1_._gang[i];
    out->packets_sent += (unsigned)s.sender._packets_sent;
    out->packets_received += s.utility._packets_received;
    if (snd) {
      snd[i].tick_share_sending = s.utility._tick_share_sending;
      snd[i].total_delay = s.utility._total_delay;
      snd[i].packets_received = s.utility._packets_received;
      snd[i].packets_sent = (unsigned)s.sender._packets_sent;
    }
  }
  out->link_queued = (uint32_t)(net._link._buffer.size() + net._link._pending_packet._queue.size());
  uint32_t infl = (uint32_t)net._delay._queue.size();
  for (const auto &v : net._rec._collector) infl += (uint32_t)v.size();
  out->delay_inflight = infl;
  std::ostringstream os; os << run_prng;
  out->prng_state = (uint32_t)std::stoul(os.str());
}

} // namespace

extern "C" int remy_ref_score_batch_fish(const RemyFlatTree *tree,
                                         const RemyLeafOverride *cands, uint32_t n_cands,
                                         const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                         uint32_t sender_stride,
                                         RemySimResult *out_sims, RemySenderResult *out_senders,
                                         uint32_t *out_use_counts, uint32_t n_threads, int mode)
{
  if (!tree || !cfgs || !seeds || !out_sims) return REMY_EINVAL;
  if (!cands) n_cands = 1;
  const uint32_t total = n_cands * n_cfgs;
  if (out_use_counts) memset(out_use_counts, 0, sizeof(uint32_t) * (size_t)n_cands * tree->n_leaves);
  std::atomic<uint32_t> next(0);
  std::mutex mu;
  auto worker = [&]() {
    std::vector<RemySenderResult> tmp(4096);
    for (;;) {
      const uint32_t s = next.fetch_add(1);
      if (s >= total) break;
      const uint32_t c = s / n_cfgs, k = s % n_cfgs;
      const RemyLeafOverride *ov = (cands && cands[c].leaf_index != REMY_NO_OVERRIDE) ? &cands[c] : nullptr;
      RemyBuffers::FinTree dna;
      build_fin_dna(tree, 0, ov, &dna);
      FinTree t(dna);
      t.reset_counts();
      if (mode == 0) run_fish_whitebox(t, cfgs[k], seeds[k], &out_sims[s], tmp.data());
      else {
        std::vector<NetConfig> one { to_netconfig(cfgs[k]) };
        auto outcome = Evaluator<FinTree>::score(t, seeds[k], one, false, (unsigned int)cfgs[k].simulation_ticks);
        memset(&out_sims[s], 0, sizeof(RemySimResult));
        out_sims[s].utility = outcome.score;
        const auto &td = outcome.throughputs_delays.at(0).second;
        for (size_t i = 0; i < td.size(); i++) { tmp[i].tick_share_sending = td[i].first; tmp[i].total_delay = td[i].second; tmp[i].packets_received = 0; tmp[i].packets_sent = 0; }
      }
      if (out_senders) {
        RemySenderResult *dst = &out_senders[(size_t)s * sender_stride];
        memset(dst, 0, sizeof(RemySenderResult) * sender_stride);
        memcpy(dst, tmp.data(), sizeof(RemySenderResult) * std::min(sender_stride, cfgs[k].num_senders));
      }
      if (out_use_counts) {
        std::vector<unsigned> counts;
        collect_fin_counts(t, counts);
        std::lock_guard<std::mutex> g(mu);
        for (size_t l = 0; l < counts.size() && l < tree->n_leaves; l++)
          out_use_counts[(size_t)c * tree->n_leaves + l] += counts[l];
      }
    }
  };
  if (n_threads <= 1) worker();
  else {
    std::vector<std::thread> th;
    for (uint32_t i = 0; i < n_threads; i++) th.emplace_back(worker);
    for (auto &t : th) t.join();
  }
  return 0;
}
