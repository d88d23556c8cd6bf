/* fish.cc — see fish.hh.  file:line citations are relative to /root/reference. */
#include "fish.hh"
#include "pbwire.hh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <stdexcept>
#include <ctime>
#include <unistd.h>

namespace remy_b200 {

/* -------------------------------------------------------------------- Fin -- */
const Fin::OptimizationSettings &Fin::get_optimizer()
{ /* src/fin.hh:49-54 */
  static OptimizationSettings s { { 0.01, 30, 0.01, 3, 4, 5 } };
  return s;
}
std::vector<Fin> Fin::next_generation() const
{ /* src/fin.cc:33-54 */
  std::vector<Fin> ret;
  const auto alts = get_optimizer().lambda.alternatives(_lambda, true);
  printf("Alternatives: lambda %f to %f\n", *std::min_element(alts.begin(), alts.end()), *std::max_element(alts.begin(), alts.end()));
  for (const auto &alt : alts) {
    Fin f(*this);
    f._generation++;
    f._lambda = alt;
    f.round();
    ret.push_back(f);
  }
  return ret;
}
std::vector<Fin> Fin::bisect() const
{
  std::vector<Fin> ret;
  for (const auto &x : _domain.bisect()) { Fin f(*this); f._domain = x; ret.push_back(f); }
  return ret;
}
void Fin::round() { _lambda = (1.0 / 10000.0) * int(10000 * _lambda); }
std::string Fin::str(unsigned int total) const
{ /* src/fin.cc:56-62 */
  char tmp[512];
  snprintf(tmp, 512, "{%s} gen=%u usage=%.4f => (lambda=%f)", _domain.str().c_str(), _generation, double(_domain.count()) / double(total), _lambda);
  return tmp;
}
std::string Fin::DNA() const
{ /* src/fin.cc:22-31 */
  pb::Writer w;
  w.put_double(37, _lambda);
  w.put_message(38, _domain.DNA());
  return w.str();
}
Fin Fin::parse(const void *data, size_t size)
{ /* src/fin.cc:16-20 */
  Fin f; f._lambda = 0;
  pb::Reader rd(data, size); pb::Field g;
  while (rd.next(g)) {
    if (g.number == 37 && g.wire_type == 1) f._lambda = pb::Reader::as_double(g);
    else if (g.number == 38 && g.wire_type == 2) f._domain = MemoryRange::parse(g.data, g.size);
  }
  return f;
}

/* ---------------------------------------------------------------- FinTree -- */
FinTree::FinTree() : _domain(Memory(), MAX_MEMORY(), { 4 /* RTT_DIFF */ }), _leaf(1, Fin(_domain)) {} /* src/fintree.cc:7-12 */
FinTree::FinTree(const Fin &fin, bool bisect) : _domain(fin.domain())
{ /* src/fintree.cc:14-26 */
  if (!bisect) _leaf.push_back(fin);
  else for (const auto &x : fin.bisect()) _children.push_back(FinTree(x, false));
}
const Fin *FinTree::fin(const Memory &m) const
{ /* src/fintree.cc:59-80 */
  if (!_domain.contains(m)) return nullptr;
  if (is_leaf()) return &_leaf[0];
  for (const auto &x : _children) { const Fin *r = x.fin(m); if (r) return r; }
  throw std::runtime_error("FinTree: no child contains the query (reference: assert(false), fintree.cc:78)");
}
const Fin &FinTree::use_fin(const Memory &m, bool track) const
{ /* src/fintree.cc:39-57 */
  const Fin *ret = fin(m);
  if (!ret) throw std::runtime_error("ERROR: No Fin found for " + m.str());
  ret->domain().use();
  if (track) ret->domain().track(m);
  return *ret;
}
void FinTree::reset_counts() { if (is_leaf()) _leaf.front().domain().reset_count(); else for (auto &x : _children) x.reset_counts(); }
void FinTree::promote(unsigned int g) { if (is_leaf()) _leaf.front().promote(g); else for (auto &x : _children) x.promote(g); }
void FinTree::reset_generation() { if (is_leaf()) _leaf.front().demote(0); else for (auto &x : _children) x.reset_generation(); }
const Fin *FinTree::most_used(unsigned int max_generation) const
{ /* src/fintree.cc:82-107 */
  if (is_leaf()) return ((_leaf.front().generation() <= max_generation) && (_leaf.front().count() > 0)) ? &_leaf[0] : nullptr;
  unsigned int count_max = 0;
  const Fin *ret = nullptr;
  for (const auto &x : _children) {
    const Fin *c = x.most_used(max_generation);
    if (c && (c->generation() <= max_generation) && (c->count() >= count_max)) { ret = c; count_max = c->count(); }
  }
  return ret;
}
bool FinTree::replace(const Fin &w)
{ /* src/fintree.cc:135-155 */
  if (!_domain.contains(w.domain().range_median())) return false;
  if (is_leaf()) {
    if (!(w.domain() == _leaf.front().domain())) throw std::runtime_error("FinTree::replace: domain mismatch (reference assert)");
    _leaf.front() = w;
    return true;
  }
  for (auto &x : _children) if (x.replace(w)) return true;
  throw std::runtime_error("FinTree::replace: no child took the fin (reference assert)");
}
bool FinTree::replace(const Fin &src, const FinTree &dst)
{ /* src/fintree.cc:157-178 */
  if (!_domain.contains(src.domain().range_median())) return false;
  if (is_leaf()) {
    if (!(src.domain() == _leaf.front().domain())) throw std::runtime_error("FinTree::replace: domain mismatch (reference assert)");
    const std::string extra = _extra;
    *this = dst;
    _extra = extra;
    return true;
  }
  for (auto &x : _children) if (x.replace(src, dst)) return true;
  throw std::runtime_error("FinTree::replace: no child took the fin (reference assert)");
}
unsigned int FinTree::total_fin_queries() const
{
  if (is_leaf()) return _leaf.front().domain().count();
  unsigned int s = 0;
  for (const auto &x : _children) s += x.total_fin_queries();
  return s;
}
std::string FinTree::str() const { return str(total_fin_queries()); }
std::string FinTree::str(unsigned int total) const
{
  if (is_leaf()) return std::string("[") + _leaf.front().str(total) + "]";
  std::string ret;
  for (const auto &x : _children) ret += x.str(total);
  return ret;
}
void FinTree::leaves(std::vector<const Fin *> &out) const { if (is_leaf()) out.push_back(&_leaf[0]); else for (const auto &x : _children) x.leaves(out); }
void FinTree::leaves(std::vector<Fin *> &out) { if (is_leaf()) out.push_back(&_leaf[0]); else for (auto &x : _children) x.leaves(out); }
int FinTree::leaf_index_of(const Fin &w) const
{ /* follow FinTree::replace's descent; count the leaves passed on the way */
  struct R { static bool go(const FinTree &t, const Memory &q, const Fin &w, int &idx) {
    if (!t._domain.contains(q)) { std::vector<const Fin *> v; t.leaves(v); idx += (int)v.size(); return false; }
    if (t.is_leaf()) { if (w.domain() == t._leaf.front().domain()) return true; idx++; return false; }
    for (const auto &x : t._children) if (go(x, q, w, idx)) return true;
    return false; } };
  int idx = 0;
  return R::go(*this, w.domain().range_median(), w, idx) ? idx : -1;
}
std::string FinTree::DNA() const
{ /* src/fintree.cc:222-240 */
  pb::Writer w;
  w.put_message(90, _domain.DNA());
  if (is_leaf()) w.put_message(92, _leaf.front().DNA());
  else for (const auto &x : _children) w.put_message(91, x.DNA());
  return w.str() + _extra;
}
FinTree FinTree::parse(const void *data, size_t size)
{ /* src/fintree.cc:242-256 */
  FinTree t; t._leaf.clear();
  pb::Reader rd(data, size); pb::Field f;
  bool has_leaf = false;
  while (rd.next(f)) {
    if (f.number == 90 && f.wire_type == 2) t._domain = MemoryRange::parse(f.data, f.size);
    else if (f.number == 91 && f.wire_type == 2) t._children.push_back(parse(f.data, f.size));
    else if (f.number == 92 && f.wire_type == 2) { t._leaf.assign(1, Fin::parse(f.data, f.size)); has_leaf = true; }
    else if (f.number >= 93 && f.number <= 95 && f.wire_type == 2) {
      pb::Writer w; w.put_message(f.number, std::string((const char *)f.data, f.size)); t._extra += w.str();
    }
  }
  if (has_leaf && !t._children.empty()) throw std::runtime_error("FinTree: node has both a leaf and children");
  if (!has_leaf && t._children.empty()) throw std::runtime_error("FinTree: node has neither a leaf nor children");
  return t;
}

FlatTree flatten(const FinTree &root)
{
  FlatTree f;
  std::vector<const Fin *> dfs;
  root.leaves(dfs);
  std::vector<const FinTree *> bfs { &root };
  for (size_t i = 0; i < bfs.size(); i++) { /* breadth-first: the children of a node are contiguous */
    const FinTree *t = bfs[i];
    RemyTreeNode n;
    for (unsigned a = 0; a < REMY_NUM_AXES; a++) { n.lower[a] = t->_domain._lower.f[a]; n.upper[a] = t->_domain._upper.f[a]; }
    n.axis_mask = t->_domain.axis_mask(); n.child_begin = 0; n.child_count = 0; n.leaf_index = 0;
    if (t->is_leaf()) n.leaf_index = (uint32_t)(std::find(dfs.begin(), dfs.end(), &t->_leaf[0]) - dfs.begin());
    else {
      n.child_begin = (uint32_t)bfs.size(); n.child_count = (uint32_t)t->_children.size();
      for (const auto &c : t->_children) bfs.push_back(&c);
    }
    f.nodes.push_back(n);
  }
  for (const Fin *w : dfs) {
    RemyLeafAction a; a.window_multiple = 0; a.intersend = w->lambda(); a.window_increment = 0; a.generation = w->generation();
    f.leaves.push_back(a);
  }
  return f;
}

/* ----------------------------------------------------- Evaluator<FinTree> -- */
Evaluator<FinTree>::Evaluator(const ConfigRange &range, unsigned int prng_seed) : _prng_seed(0), _tick_count(range.simulation_ticks)
{ /* the config expansion and the frozen seed are those of Evaluator<T>::Evaluator (src/evaluator.cc:9-38) */
  const Evaluator<WhiskerTree> e(range, prng_seed);
  _prng_seed = e.prng_seed();
  _configs = e.configs();
}

static void check_fish_sim(const RemySimResult &r)
{
  if (r.error & REMY_ERR_NO_WHISKER) throw std::runtime_error("ERROR: No Fin found (reference: exit(1), fintree.cc:43-46)");
  if (r.error) throw std::runtime_error("simulation failed with error word " + std::to_string(r.error));
}

Evaluator<FinTree>::Outcome Evaluator<FinTree>::score(FinTree &run_fins, unsigned int prng_seed, const std::vector<NetConfig> &configs,
                                                      bool trace, unsigned int ticks_to_run)
{ /* src/evaluator.cc:105-132, one seed per config like Evaluator<WhiskerTree>::score here */
  const FlatTree ft = flatten(run_fins);
  const RemyFlatTree tree = ft.c();
  std::vector<RemyNetConfig> cfgs;
  uint32_t stride = 1;
  for (const auto &x : configs) { cfgs.push_back(x.pod((double)ticks_to_run)); stride = std::max(stride, cfgs.back().num_senders); }
  const std::vector<uint32_t> sd = Evaluator<WhiskerTree>::seeds(prng_seed, cfgs.size());
  std::vector<RemySimResult> sims(cfgs.size());
  std::vector<RemySenderResult> snd(cfgs.size() * (size_t)stride);
  std::vector<uint32_t> counts(ft.leaves.size());
  if (trace) { /* Fish(run_fins, seed, trace) (src/evaluator.cc:122): every lookup also feeds the fin's running medians */
    struct Ctx { std::vector<const MemoryRange *> doms; uint32_t mask; } ctx;
    std::vector<const Fin *> lv;
    static_cast<const FinTree &>(run_fins).leaves(lv);
    for (const Fin *f : lv) ctx.doms.push_back(&f->domain());
    ctx.mask = 0;
    for (const auto &nd : ft.nodes) ctx.mask |= nd.axis_mask;
    auto feed = [](void *user, uint32_t, const uint64_t *rec, uint64_t n, uint32_t words) -> int {
      Ctx *c = static_cast<Ctx *>(user);
      return feed_trace_records(c->doms, c->mask, rec, n, words);
    };
    if (remy_b200_score_trace_fish(&tree, cfgs.data(), sd.data(), (uint32_t)cfgs.size(), stride, sims.data(), snd.data(), counts.data(), feed, &ctx))
      throw std::runtime_error(std::string("remy_b200_score_trace_fish: ") + remy_b200_last_error());
  } else if (remy_b200_score_batch_fish(&tree, nullptr, 1, cfgs.data(), sd.data(), (uint32_t)cfgs.size(), stride, sims.data(), snd.data(), counts.data()))
    throw std::runtime_error(std::string("remy_b200_score_batch_fish: ") + remy_b200_last_error());
  Outcome out;
  for (size_t k = 0; k < cfgs.size(); k++) {
    check_fish_sim(sims[k]);
    out.score += sims[k].utility;
    std::vector<std::pair<double, double>> td;
    for (uint32_t s = 0; s < cfgs[k].num_senders; s++) {
      const RemySenderResult &r = snd[k * stride + s];
      td.emplace_back(r.tick_share_sending == 0 ? 0.0 : double(r.packets_received) / r.tick_share_sending,
                      r.packets_received == 0 ? 0.0 : double(r.total_delay) / double(r.packets_received));
    }
    out.throughputs_delays.emplace_back(configs[k], td);
  }
  std::vector<Fin *> lv;
  run_fins.leaves(lv);
  for (size_t i = 0; i < lv.size(); i++) lv[i]->_domain._count = counts[i];
  out.used_actions = run_fins;
  return out;
}

Evaluator<FinTree>::Outcome Evaluator<FinTree>::score(FinTree &run_actions, bool trace, double carefulness) const
{ /* src/evaluator.cc:196-201 */
  return score(run_actions, _prng_seed, _configs, trace, _tick_count * carefulness);
}

std::vector<double> Evaluator<FinTree>::score_candidates(const FinTree &base, const std::vector<Fin> &replacements, double carefulness) const
{
  const FlatTree ft = flatten(base);
  const RemyFlatTree tree = ft.c();
  const unsigned int ticks = _tick_count * carefulness;
  std::vector<RemyNetConfig> cfgs;
  for (const auto &x : _configs) cfgs.push_back(x.pod((double)ticks));
  const std::vector<uint32_t> sd = Evaluator<WhiskerTree>::seeds(_prng_seed, cfgs.size());
  std::vector<RemyLeafOverride> cands;
  for (const auto &r : replacements) {
    const int idx = base.leaf_index_of(r);
    if (idx < 0) throw std::runtime_error("score_candidates: replacement does not match a fin of the tree");
    RemyLeafOverride o; o.window_multiple = 0; o.intersend = r.lambda(); o.window_increment = 0; o.leaf_index = (uint32_t)idx;
    cands.push_back(o);
  }
  std::vector<RemySimResult> sims(cands.size() * cfgs.size());
  if (remy_b200_score_batch_fish(&tree, cands.data(), (uint32_t)cands.size(), cfgs.data(), sd.data(), (uint32_t)cfgs.size(), 0, sims.data(), nullptr, nullptr))
    throw std::runtime_error(std::string("remy_b200_score_batch_fish: ") + remy_b200_last_error());
  std::vector<double> out(cands.size(), 0.0);
  for (size_t c = 0; c < cands.size(); c++)
    for (size_t k = 0; k < cfgs.size(); k++) { check_fish_sim(sims[c * cfgs.size() + k]); out[c] += sims[c * cfgs.size() + k].utility; }
  return out;
}

/* ------------------------------------------------------------ FinImprover -- */
size_t FinHash::operator()(const Fin &f) const
{
  size_t seed = 0;
  auto mix = [&seed](size_t v) { seed ^= v + 0x9e3779b9 + (seed << 6) + (seed >> 2); };
  mix(std::hash<double>()(f.lambda()));
  for (int a : f.domain()._active_axis) { mix(std::hash<double>()(f.domain()._lower.field(a))); mix(std::hash<double>()(f.domain()._upper.field(a))); }
  return seed;
}
void FinImprover::evaluate_replacements(const std::vector<Fin> &replacements, std::vector<std::pair<bool, double>> &scores, double carefulness)
{ /* src/breeder.cc:53-77 as ONE batch */
  std::vector<Fin> fresh; std::vector<size_t> where;
  scores.assign(replacements.size(), std::make_pair(false, 0.0));
  for (size_t i = 0; i < replacements.size(); i++) {
    auto it = eval_cache_.find(replacements[i]);
    if (it == eval_cache_.end()) { fresh.push_back(replacements[i]); where.push_back(i); }
    else scores[i] = std::make_pair(false, it->second);
  }
  if (fresh.empty()) return;
  const std::vector<double> s = eval_.score_candidates(tree_, fresh, carefulness);
  candidates_scored += fresh.size(); batches++;
  for (size_t k = 0; k < fresh.size(); k++) scores[where[k]] = std::make_pair(true, s[k]);
}
std::vector<Fin> FinImprover::early_bail_out(const std::vector<Fin> &replacements, double carefulness, double quantile_to_keep)
{ /* src/breeder.cc:79-114 */
  std::vector<std::pair<bool, double>> scores;
  evaluate_replacements(replacements, scores, carefulness);
  std::vector<double> raw;
  for (auto &x : scores) raw.push_back(x.second);
  const double lower_bound = std::min(score_to_beat_ * (1 + MAX_PERCENT_ERROR), score_to_beat_ * (1 - MAX_PERCENT_ERROR));
  std::vector<double> sorted(raw);
  std::sort(sorted.begin(), sorted.end(), std::greater<double>());
  const double p = 1 - quantile_to_keep; /* tail_quantile<right>, quantile_probability = p (src/breeder.cc:88-101) */
  const size_t n = (size_t)std::ceil(sorted.size() * (1. - p));
  if (n < 1 || n > sorted.size()) throw std::runtime_error("early_bail_out: empty candidate set");
  const double cutoff = std::max(lower_bound, sorted[n - 1]);
  std::vector<Fin> top;
  for (size_t i = 0; i < replacements.size(); i++) if (raw[i] >= cutoff) top.push_back(replacements[i]);
  return top;
}
double FinImprover::improve(Fin &action_to_improve)
{ /* src/breeder.cc:116-150 */
  auto replacements = action_to_improve.next_generation();
  std::vector<std::pair<bool, double>> scores;
  std::vector<Fin> top = early_bail_out(replacements, 0.1, 0.5);
  evaluate_replacements(top, scores, 1);
  for (size_t i = 0; i < top.size(); i++) {
    if (scores[i].first) eval_cache_.insert(std::make_pair(top[i], scores[i].second));
    if (scores[i].second > score_to_beat_) { score_to_beat_ = scores[i].second; action_to_improve = top[i]; }
  }
  std::cout << "Chose " << action_to_improve.str() << std::endl;
  return score_to_beat_;
}

/* ------------------------------------------------------------ FishBreeder -- */
FishBreeder::FishBreeder(const BreederOptions &o, unsigned int seed) : _options(o), _seed_state(0)
{
  const unsigned int s = seed ? seed : (unsigned int)(time(NULL) ^ getpid());
  _seed_state = s % 2147483647u;
  if (_seed_state == 0) _seed_state = 1;
}
unsigned int FishBreeder::next_seed()
{
  _seed_state = (uint32_t)(((uint64_t)_seed_state * 16807u) % 2147483647u);
  return _seed_state;
}
void FishBreeder::apply_best_split(FinTree &tree, unsigned int generation)
{ /* Breeder<FinTree>::apply_best_split, src/breeder.cc:15-41 (without the fork's stray assert(false)) */
  const Evaluator<FinTree> eval(_options.config_range, next_seed());
  auto outcome(eval.score(tree, true));
  while (1) {
    const Fin *my_action = outcome.used_actions.most_used(generation);
    if (!my_action) throw std::runtime_error("apply_best_split: no fin was used (reference assert, breeder.cc:25)");
    FinTree bisected_action(*my_action, true);
    if (bisected_action.num_children() == 1) {
      printf("Got unbisectable whisker! %s\n", my_action->str().c_str());
      Fin mutable_action(*my_action);
      mutable_action.promote(generation + 1);
      if (!outcome.used_actions.replace(mutable_action)) throw std::runtime_error("apply_best_split: replace failed (reference assert, breeder.cc:33)");
      continue;
    }
    printf("Bisecting === %s === into === %s ===\n", my_action->str().c_str(), bisected_action.str().c_str());
    if (!tree.replace(*my_action, bisected_action)) throw std::runtime_error("apply_best_split: replace failed (reference assert, breeder.cc:38)");
    break;
  }
}
Evaluator<FinTree>::Outcome FishBreeder::improve(FinTree &fins)
{ /* src/fishbreeder.cc:7-72 */
  FinTree input_fintree(fins);
  fins.reset_generation();
  unsigned int generation = 0;
  while (generation < 5) {
    const Evaluator<FinTree> eval(_options.config_range, next_seed());
    auto outcome(eval.score(fins));
    const Fin *most_used_fin_ptr = outcome.used_actions.most_used(generation);
    if (!most_used_fin_ptr) { generation++; fins.promote(generation); continue; }
    FinImprover improver(eval, fins, outcome.score);
    Fin fin_to_improve = *most_used_fin_ptr;
    double score_to_beat = outcome.score;
    while (1) {
      const double new_score = improver.improve(fin_to_improve);
      if (new_score < score_to_beat) throw std::runtime_error("FishBreeder: score decreased (reference assert, fishbreeder.cc:41)");
      if (new_score == score_to_beat) { std::cerr << "Ending search." << std::endl; break; }
      std::cerr << "Score jumps from " << score_to_beat << " to " << new_score << std::endl;
      score_to_beat = new_score;
    }
    candidates_scored += improver.candidates_scored; batches += improver.batches;
    fin_to_improve.demote(generation + 1);
    if (!fins.replace(fin_to_improve)) throw std::runtime_error("FishBreeder: replace failed (reference assert, fishbreeder.cc:55)");
  }
  apply_best_split(fins, generation);
  const Evaluator<FinTree> eval2(_options.config_range, next_seed());
  const auto new_score = eval2.score(fins, false, 10);
  const auto old_score = eval2.score(input_fintree, false, 10);
  if (old_score.score >= new_score.score) {
    fprintf(stderr, "Regression, old=%f, new=%f\n", old_score.score, new_score.score);
    fins = input_fintree;
    return old_score;
  }
  return new_score;
}

} // namespace remy_b200
