/*
 * host_capi.cc — thin extern "C" view of the C++ host mirror (remy.hh) so that the Python
 * test-suite can drive it with ctypes.  Not part of the drop-in boundary (that is
 * include/remy_b200.h); exceptions are turned into return codes + remy_host_last_error().
 */
#include "remy.hh"
#include "breeder.hh"
#include "fish.hh"

#include <cstring>
#include <stdexcept>
#include <vector>
#include <string>

using namespace remy_b200;

namespace { thread_local std::string g_err; }

#define GUARD(body) try { body } catch (const std::exception &e) { g_err = e.what(); return -1; }

extern "C" {

const char *remy_host_last_error(void) { return g_err.c_str(); }

void *remy_host_tree_default(void) { return new WhiskerTree(); }
void *remy_host_tree_parse(const void *data, size_t size)
{
  try { return new WhiskerTree(WhiskerTree::parse(data, size)); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void remy_host_tree_free(void *h) { delete (WhiskerTree *)h; }

/* inverse of flatten(): rebuild a WhiskerTree from the POD arrays of include/remy_b200.h */
static WhiskerTree from_flat(const RemyTreeNode *nodes, const RemyLeafAction *leaves, uint32_t i)
{
  WhiskerTree t; t._leaf.clear();
  Memory lo, hi;
  for (unsigned a = 0; a < REMY_NUM_AXES; a++) { lo.f[a] = nodes[i].lower[a]; hi.f[a] = nodes[i].upper[a]; }
  std::vector<int> axes;
  for (int a = 0; a < REMY_NUM_AXES; a++) if (nodes[i].axis_mask & (1u << a)) axes.push_back(a);
  t._domain = MemoryRange(lo, hi, axes);
  if (nodes[i].child_count == 0) {
    const RemyLeafAction &l = leaves[nodes[i].leaf_index];
    Whisker w((unsigned int)l.window_increment, l.window_multiple, l.intersend, t._domain);
    w._window_increment = l.window_increment; w._generation = l.generation;
    t._leaf.push_back(w);
  } else {
    for (uint32_t c = 0; c < nodes[i].child_count; c++) t._children.push_back(from_flat(nodes, leaves, nodes[i].child_begin + c));
  }
  return t;
}
void *remy_host_tree_from_flat(const RemyFlatTree *flat)
{
  try { return new WhiskerTree(from_flat(flat->nodes, flat->leaves, 0)); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

static long copy_out(const std::string &s, char *buf, size_t cap) { if (buf && s.size() <= cap) memcpy(buf, s.data(), s.size()); return (long)s.size(); }
long remy_host_tree_serialize(void *h, char *buf, size_t cap) { GUARD(return copy_out(((WhiskerTree *)h)->DNA(), buf, cap);) }
long remy_host_tree_str(void *h, char *buf, size_t cap) { GUARD(return copy_out(((WhiskerTree *)h)->str(), buf, cap);) }

int remy_host_tree_flatten(void *h, RemyTreeNode *nodes, uint32_t cap_nodes, RemyLeafAction *leaves, uint32_t cap_leaves,
                           uint32_t *n_nodes, uint32_t *n_leaves)
{
  GUARD(
    const FlatTree f = flatten(*(WhiskerTree *)h);
    *n_nodes = (uint32_t)f.nodes.size(); *n_leaves = (uint32_t)f.leaves.size();
    if (nodes && f.nodes.size() <= cap_nodes) memcpy(nodes, f.nodes.data(), f.nodes.size() * sizeof(RemyTreeNode));
    if (leaves && f.leaves.size() <= cap_leaves) memcpy(leaves, f.leaves.data(), f.leaves.size() * sizeof(RemyLeafAction));
    return 0;)
}

/* Whisker::next_generation of leaf `leaf_index` -> overrides; returns the count */
long remy_host_next_generation(void *h, uint32_t leaf_index, int oi, int om, int os, RemyLeafOverride *out, uint32_t cap)
{
  GUARD(
    std::vector<const Whisker *> lv; ((WhiskerTree *)h)->leaves(lv);
    if (leaf_index >= lv.size()) throw std::runtime_error("leaf index out of range");
    const std::vector<Whisker> gen = lv[leaf_index]->next_generation(oi, om, os);
    for (size_t i = 0; i < gen.size() && i < cap; i++) {
      const int idx = ((WhiskerTree *)h)->leaf_index_of(gen[i]);
      out[i].window_multiple = gen[i].window_multiple(); out[i].intersend = gen[i].intersend();
      out[i].window_increment = gen[i].window_increment(); out[i].leaf_index = (uint32_t)idx;
    }
    return (long)gen.size();)
}

/* Evaluator::Evaluator(ConfigRange) expansion: bytes of a ConfigRange / ConfigRangeUnicorn -> NetConfigs */
long remy_host_config_range_expand(const void *data, size_t size, double carefulness, RemyNetConfig *out, uint32_t cap, char *dna_out, size_t dna_cap, long *dna_size)
{
  GUARD(
    const ConfigRange r = ConfigRange::parse(data, size);
    if (dna_size) *dna_size = copy_out(r.DNA(), dna_out, dna_cap);
    Evaluator<WhiskerTree> e(r, 1);
    const unsigned int ticks = r.simulation_ticks * carefulness;
    for (size_t i = 0; i < e.configs().size() && i < cap; i++) out[i] = e.configs()[i].pod((double)ticks);
    return (long)e.configs().size();)
}

/* Evaluator<WhiskerTree>::score(tree, trace=false, carefulness) — needs a GPU */
int remy_host_evaluator_score(void *h, const void *range_bytes, size_t range_size, unsigned int prng_seed, double carefulness,
                              double *score, uint32_t *use_counts, uint32_t n_counts, double *tput_delay, uint32_t td_cap, uint32_t *n_td)
{
  GUARD(
    WhiskerTree &t = *(WhiskerTree *)h;
    Evaluator<WhiskerTree> e(ConfigRange::parse(range_bytes, range_size), prng_seed);
    const auto out = e.score(t, false, carefulness);
    *score = out.score;
    std::vector<const Whisker *> lv; out.used_actions.leaves(lv);
    for (size_t i = 0; i < lv.size() && i < n_counts; i++) use_counts[i] = lv[i]->count();
    uint32_t k = 0;
    for (const auto &cfg : out.throughputs_delays)
      for (const auto &p : cfg.second) { if (k + 2 <= td_cap) { tput_delay[k] = p.first; tput_delay[k + 1] = p.second; } k += 2; }
    if (n_td) *n_td = k;
    return 0;)
}

/* Evaluator::score_candidates for Whisker::next_generation(leaf) — needs a GPU */
long remy_host_score_next_generation(void *h, const void *range_bytes, size_t range_size, unsigned int prng_seed, double carefulness,
                                     uint32_t leaf_index, double *scores, uint32_t cap)
{
  GUARD(
    WhiskerTree &t = *(WhiskerTree *)h;
    std::vector<const Whisker *> lv; t.leaves(lv);
    if (leaf_index >= lv.size()) throw std::runtime_error("leaf index out of range");
    const std::vector<Whisker> gen = lv[leaf_index]->next_generation(true, true, true);
    Evaluator<WhiskerTree> e(ConfigRange::parse(range_bytes, range_size), prng_seed);
    const std::vector<double> s = e.score_candidates(t, gen, carefulness);
    for (size_t i = 0; i < s.size() && i < cap; i++) scores[i] = s[i];
    return (long)s.size();)
}

/* Evaluator::DNA(tree): ProblemBuffers.Problem bytes for a (range, seed, tree) — no GPU needed */
long remy_host_problem_dna(void *h, const void *range_bytes, size_t range_size, unsigned int prng_seed, char *buf, size_t cap)
{
  GUARD(
    Evaluator<WhiskerTree> e(ConfigRange::parse(range_bytes, range_size), prng_seed);
    return copy_out(e.DNA(*(WhiskerTree *)h), buf, cap);)
}

/* parse_problem_and_evaluate(Problem bytes) -> AnswerBuffers.Outcome bytes — needs a GPU */
long remy_host_solve_problem(const void *problem, size_t size, char *buf, size_t cap)
{
  GUARD(
    const auto outcome = Evaluator<WhiskerTree>::parse_problem_and_evaluate(problem, size);
    return copy_out(outcome.DNA(), buf, cap);)
}

/* Outcome(AnswerBuffers::Outcome).DNA() round trip — no GPU needed */
long remy_host_answer_roundtrip(const void *answer, size_t size, char *buf, size_t cap)
{
  GUARD(return copy_out(Evaluator<WhiskerTree>::Outcome::parse(answer, size).DNA(), buf, cap);)
}

/* WhiskerImprover::improve on leaf `leaf_index` of the tree, repeated until the score stops rising
 * (the inner loop of RatBreeder::improve, src/ratbreeder.cc:39-49).  Needs a GPU.  Returns the number of
 * improve() calls; out = {score_to_beat, chosen increment, multiple, intersend, candidates scored, batches}. */
long remy_host_improve_whisker(void *h, const void *range_bytes, size_t range_size, unsigned int prng_seed, uint32_t leaf_index, double *out)
{
  GUARD(
    WhiskerTree &t = *(WhiskerTree *)h;
    std::vector<const Whisker *> lv; t.leaves(lv);
    if (leaf_index >= lv.size()) throw std::runtime_error("leaf index out of range");
    Evaluator<WhiskerTree> e(ConfigRange::parse(range_bytes, range_size), prng_seed);
    auto outcome = e.score(t);
    WhiskerImprover improver(e, t, WhiskerImproverOptions(), outcome.score);
    Whisker w = *lv[leaf_index];
    double score_to_beat = outcome.score;
    long calls = 0;
    for (;;) {
      const double ns = improver.improve(w); calls++;
      if (ns == score_to_beat) break;
      score_to_beat = ns;
    }
    out[0] = score_to_beat; out[1] = w.window_increment(); out[2] = w.window_multiple(); out[3] = w.intersend();
    out[4] = (double)improver.candidates_scored; out[5] = (double)improver.batches;
    return calls;)
}

/* RatBreeder::improve; the tree behind `h` is updated in place */
int remy_host_ratbreeder_improve(void *h, const void *range_bytes, size_t range_size, unsigned int seed, double *score, double *stats)
{
  GUARD(
    WhiskerTree &t = *(WhiskerTree *)h;
    BreederOptions o; o.config_range = ConfigRange::parse(range_bytes, range_size);
    RatBreeder b(o, WhiskerImproverOptions(), seed);
    const auto outcome = b.improve(t);
    *score = outcome.score;
    if (stats) { stats[0] = (double)b.candidates_scored; stats[1] = (double)b.batches; }
    return 0;)
}

/* ---- Fish / FinTree (fish.hh) ---- */
void *remy_host_fintree_default(void) { return new FinTree(); }
void *remy_host_fintree_parse(const void *data, size_t size)
{
  try { return new FinTree(FinTree::parse(data, size)); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void remy_host_fintree_free(void *h) { delete (FinTree *)h; }
static FinTree fin_from_flat(const RemyTreeNode *nodes, const RemyLeafAction *leaves, uint32_t i)
{
  FinTree t; t._leaf.clear();
  Memory lo, hi;
  for (unsigned a = 0; a < REMY_NUM_AXES; a++) { lo.f[a] = nodes[i].lower[a]; hi.f[a] = nodes[i].upper[a]; }
  std::vector<int> axes;
  for (int a = 0; a < REMY_NUM_AXES; a++) if (nodes[i].axis_mask & (1u << a)) axes.push_back(a);
  t._domain = MemoryRange(lo, hi, axes);
  if (nodes[i].child_count == 0) {
    Fin f(leaves[nodes[i].leaf_index].intersend, t._domain);
    f._generation = leaves[nodes[i].leaf_index].generation;
    t._leaf.push_back(f);
  } else {
    for (uint32_t c = 0; c < nodes[i].child_count; c++) t._children.push_back(fin_from_flat(nodes, leaves, nodes[i].child_begin + c));
  }
  return t;
}
void *remy_host_fintree_from_flat(const RemyFlatTree *flat)
{
  try { return new FinTree(fin_from_flat(flat->nodes, flat->leaves, 0)); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
long remy_host_fintree_serialize(void *h, char *buf, size_t cap) { GUARD(return copy_out(((FinTree *)h)->DNA(), buf, cap);) }
long remy_host_fintree_str(void *h, char *buf, size_t cap) { GUARD(return copy_out(((FinTree *)h)->str(), buf, cap);) }
int remy_host_fintree_flatten(void *h, RemyTreeNode *nodes, uint32_t cap_nodes, RemyLeafAction *leaves, uint32_t cap_leaves,
                              uint32_t *n_nodes, uint32_t *n_leaves)
{
  GUARD(
    const FlatTree f = flatten(*(FinTree *)h);
    *n_nodes = (uint32_t)f.nodes.size(); *n_leaves = (uint32_t)f.leaves.size();
    if (nodes && f.nodes.size() <= cap_nodes) memcpy(nodes, f.nodes.data(), f.nodes.size() * sizeof(RemyTreeNode));
    if (leaves && f.leaves.size() <= cap_leaves) memcpy(leaves, f.leaves.data(), f.leaves.size() * sizeof(RemyLeafAction));
    return 0;)
}
/* Fin::next_generation of leaf `leaf_index` -> lambdas; returns the count */
long remy_host_fin_next_generation(void *h, uint32_t leaf_index, double *out, uint32_t cap)
{
  GUARD(
    std::vector<const Fin *> lv; ((FinTree *)h)->leaves(lv);
    if (leaf_index >= lv.size()) throw std::runtime_error("leaf index out of range");
    const std::vector<Fin> gen = lv[leaf_index]->next_generation();
    for (size_t i = 0; i < gen.size() && i < cap; i++) out[i] = gen[i].lambda();
    return (long)gen.size();)
}
/* Evaluator<FinTree>::score(tree, trace = false, carefulness) — needs a GPU */
int remy_host_fish_score(void *h, const void *range_bytes, size_t range_size, unsigned int prng_seed, double carefulness,
                         double *score, uint32_t *use_counts, uint32_t n_counts)
{
  GUARD(
    FinTree &t = *(FinTree *)h;
    Evaluator<FinTree> e(ConfigRange::parse(range_bytes, range_size), prng_seed);
    const auto out = e.score(t, false, carefulness);
    *score = out.score;
    std::vector<const Fin *> lv; out.used_actions.leaves(lv);
    for (size_t i = 0; i < lv.size() && i < n_counts; i++) use_counts[i] = lv[i]->count();
    return 0;)
}
/* FinImprover::improve on leaf `leaf_index`, repeated until the score stops rising (src/fishbreeder.cc:39-49) — needs a GPU.
 * out = {score_to_beat, chosen lambda, candidates scored, batches}; returns the number of improve() calls */
long remy_host_improve_fin(void *h, const void *range_bytes, size_t range_size, unsigned int prng_seed, uint32_t leaf_index, double *out)
{
  GUARD(
    FinTree &t = *(FinTree *)h;
    std::vector<const Fin *> lv; t.leaves(lv);
    if (leaf_index >= lv.size()) throw std::runtime_error("leaf index out of range");
    Evaluator<FinTree> e(ConfigRange::parse(range_bytes, range_size), prng_seed);
    auto outcome = e.score(t);
    FinImprover improver(e, t, outcome.score);
    Fin f = *lv[leaf_index];
    double score_to_beat = outcome.score;
    long calls = 0;
    for (;;) {
      const double ns = improver.improve(f); calls++;
      if (ns == score_to_beat) break;
      score_to_beat = ns;
    }
    out[0] = score_to_beat; out[1] = f.lambda(); out[2] = (double)improver.candidates_scored; out[3] = (double)improver.batches;
    return calls;)
}

/* Evaluator<FinTree>::score(tree, seed, configs, trace = true, ticks) over explicit NetConfigs — needs a GPU */
int remy_host_fish_score_trace(void *h, const RemyNetConfig *cfgs, uint32_t n_cfgs, unsigned int prng_seed, double *score)
{
  GUARD(
    FinTree &t = *(FinTree *)h;
    std::vector<NetConfig> configs;
    for (uint32_t k = 0; k < n_cfgs; k++) {
      NetConfig c; c.mean_on_duration = cfgs[k].mean_on_duration; c.mean_off_duration = cfgs[k].mean_off_duration;
      c.num_senders = cfgs[k].num_senders; c.link_ppt = cfgs[k].link_ppt; c.delay = cfgs[k].delay;
      c.buffer_size = cfgs[k].buffer_size; c.stochastic_loss_rate = cfgs[k].stochastic_loss_rate;
      configs.push_back(c);
    }
    const auto out = Evaluator<FinTree>::score(t, prng_seed, configs, true, n_cfgs ? (unsigned int)cfgs[0].simulation_ticks : 0);
    *score = out.score;
    return 0;)
}
int remy_host_fintree_medians(void *h, double *out, uint32_t n_leaves)
{
  GUARD(
    std::vector<const Fin *> lv; ((FinTree *)h)->leaves(lv);
    for (size_t l = 0; l < lv.size() && l < n_leaves; l++)
      for (unsigned a = 0; a < Memory::datasize; a++) out[l * Memory::datasize + a] = lv[l]->domain()._acc[a].median();
    return 0;)
}
/* Breeder<FinTree>::apply_best_split through FishBreeder — needs a GPU */
int remy_host_fish_apply_best_split(void *h, const void *range_bytes, size_t range_size, unsigned int seed, unsigned int generation)
{
  GUARD(
    BreederOptions o; o.config_range = ConfigRange::parse(range_bytes, range_size);
    FishBreeder b(o, seed);
    b.apply_best_split(*(FinTree *)h, generation);
    return 0;)
}
/* FishBreeder::improve; the tree behind `h` is updated in place — needs a GPU */
int remy_host_fishbreeder_improve(void *h, const void *range_bytes, size_t range_size, unsigned int seed, double *score)
{
  GUARD(
    BreederOptions o; o.config_range = ConfigRange::parse(range_bytes, range_size);
    FishBreeder b(o, seed);
    *score = b.improve(*(FinTree *)h).score;
    return 0;)
}

/* ---- trace == true ---- */

/* P2Median fed with `n` samples — no GPU needed */
double remy_host_p2_median(const double *samples, size_t n)
{
  P2Median acc;
  for (size_t i = 0; i < n; i++) acc(samples[i]);
  return acc.median();
}

/* MemoryRange::track of leaf `leaf_index` with `n` Memory values (6 doubles each) — no GPU needed */
int remy_host_track(void *h, uint32_t leaf_index, const double *mem, size_t n)
{
  GUARD(
    std::vector<Whisker *> lv; ((WhiskerTree *)h)->leaves(lv);
    if (leaf_index >= lv.size()) throw std::runtime_error("leaf index out of range");
    Memory m;
    for (size_t i = 0; i < n; i++) { for (unsigned a = 0; a < Memory::datasize; a++) m.f[a] = mem[i * Memory::datasize + a]; lv[leaf_index]->domain().track(m); }
    return 0;)
}

/* the sink of Evaluator::score(trace = true) on raw device-format records — no GPU needed */
int remy_host_feed_records(void *h, const uint64_t *records, uint64_t n, uint32_t words)
{
  GUARD(return feed_trace_records(*(WhiskerTree *)h, records, n, words);)
}

/* median(_acc[axis]) of every leaf: out[leaf * 6 + axis] — no GPU needed */
int remy_host_medians(void *h, double *out, uint32_t n_leaves)
{
  GUARD(
    std::vector<const Whisker *> lv; ((WhiskerTree *)h)->leaves(lv);
    for (size_t l = 0; l < lv.size() && l < n_leaves; l++)
      for (unsigned a = 0; a < Memory::datasize; a++) out[l * Memory::datasize + a] = lv[l]->domain()._acc[a].median();
    return 0;)
}

/* set the use counts of the leaves (what a scoring run leaves behind) — no GPU needed */
int remy_host_set_counts(void *h, const uint32_t *counts, uint32_t n_leaves)
{
  GUARD(
    std::vector<Whisker *> lv; ((WhiskerTree *)h)->leaves(lv);
    for (size_t l = 0; l < lv.size() && l < n_leaves; l++) lv[l]->_domain._count = counts[l];
    return 0;)
}

/* The loop of Breeder::apply_best_split (src/breeder.cc:23-40) on a tree whose leaves already carry
 * use counts and running medians — no GPU needed */
int remy_host_split_most_used(void *h, unsigned int generation)
{
  GUARD(
    WhiskerTree &tree = *(WhiskerTree *)h;
    WhiskerTree used_actions(tree);
    while (1) {
      const Whisker *my_action = used_actions.most_used(generation);
      if (!my_action) throw std::runtime_error("no whisker was used");
      WhiskerTree bisected_action(*my_action, true);
      if (bisected_action.num_children() == 1) {
        Whisker mutable_action(*my_action);
        mutable_action.promote(generation + 1);
        used_actions.replace(mutable_action);
        continue;
      }
      tree.replace(*my_action, bisected_action);
      break;
    }
    return 0;)
}

/* Evaluator<WhiskerTree>::score(tree, seed, configs, trace = true, ticks) over explicit NetConfigs — needs a GPU.
 * The tree behind `h` keeps the use counts and the running medians. */
int remy_host_score_trace(void *h, const RemyNetConfig *cfgs, uint32_t n_cfgs, unsigned int prng_seed, double *score)
{
  GUARD(
    WhiskerTree &t = *(WhiskerTree *)h;
    std::vector<NetConfig> configs;
    for (uint32_t k = 0; k < n_cfgs; k++) {
      NetConfig c; c.mean_on_duration = cfgs[k].mean_on_duration; c.mean_off_duration = cfgs[k].mean_off_duration;
      c.num_senders = cfgs[k].num_senders; c.link_ppt = cfgs[k].link_ppt; c.delay = cfgs[k].delay;
      c.buffer_size = cfgs[k].buffer_size; c.stochastic_loss_rate = cfgs[k].stochastic_loss_rate;
      configs.push_back(c);
    }
    const auto out = Evaluator<WhiskerTree>::score(t, prng_seed, configs, true, n_cfgs ? (unsigned int)cfgs[0].simulation_ticks : 0);
    *score = out.score;
    return 0;)
}

/* Breeder::apply_best_split through RatBreeder — needs a GPU */
int remy_host_apply_best_split(void *h, const void *range_bytes, size_t range_size, unsigned int seed, unsigned int generation)
{
  GUARD(
    BreederOptions o; o.config_range = ConfigRange::parse(range_bytes, range_size);
    RatBreeder b(o, WhiskerImproverOptions(), seed);
    b.apply_best_split(*(WhiskerTree *)h, generation);
    return 0;)
}

void remy_host_set_chained_prng(int on) { set_chained_prng(on != 0); }

/* Evaluator<WhiskerTree>::seeds: the per-config seeds of one Evaluator (INTEGRATION.md shows the same rule). */
int remy_host_seeds(unsigned int prng_seed, uint32_t *out, uint32_t n)
{
  GUARD(const std::vector<uint32_t> s = Evaluator<WhiskerTree>::seeds(prng_seed, n);
        for (uint32_t i = 0; i < n; i++) out[i] = s[i];
        return 0;)
}

} /* extern "C" */
