/*
 * host_capi.cc — thin extern "C" view of the C++ host mirror (remy.hh) so that the Python
 * test-suite can drive it with ctypes.  Not part of the drop-in boundary (that is
 * include/remy_b200.h); exceptions are turned into return codes + remy_host_last_error().
 */
#include "remy.hh"
#include "breeder.hh"

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

/* RatBreeder::improve (without apply_best_split); the tree behind `h` is updated in place */
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

} /* extern "C" */
