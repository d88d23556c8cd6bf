/* breeder.cc — see breeder.hh.  file:line citations are relative to /root/reference. */
#include "breeder.hh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <functional>
#include <iostream>
#include <stdexcept>
#include <unistd.h>

namespace remy_b200 {

size_t WhiskerHash::operator()(const Whisker &w) const
{ /* any hash consistent with Whisker::operator== will do (reference: boost::hash, src/whisker.cc:97-106) */
  size_t seed = 0;
  auto mix = [&seed](size_t v) { seed ^= v + 0x9e3779b9 + (seed << 6) + (seed >> 2); };
  mix(std::hash<int>()(w.window_increment())); mix(std::hash<double>()(w.window_multiple())); mix(std::hash<double>()(w.intersend()));
  for (int a : w.domain()._active_axis) { mix(std::hash<double>()(w.domain()._lower.field(a))); mix(std::hash<double>()(w.domain()._upper.field(a))); }
  return seed;
}

std::vector<Whisker> WhiskerImprover::get_replacements(Whisker &w)
{ /* src/ratbreeder.cc:74-79 */
  return w.next_generation(_options.optimize_window_increment, _options.optimize_window_multiple, _options.optimize_intersend);
}

void WhiskerImprover::evaluate_replacements(const std::vector<Whisker> &replacements, std::vector<std::pair<bool, double>> &scores,
                                            double carefulness)
{ /* src/breeder.cc:53-77: uncached candidates are evaluated, cached ones reuse their score */
  std::vector<Whisker> fresh;
  std::vector<size_t> where;
  scores.assign(replacements.size(), std::make_pair(false, 0.0));
  for (size_t i = 0; i < replacements.size(); i++) {
    auto it = eval_cache_.find(replacements[i]);
    if (it == eval_cache_.end()) { fresh.push_back(replacements[i]); where.push_back(i); }
    else scores[i] = std::make_pair(false, it->second);
  }
  if (fresh.empty()) return;
  const std::vector<double> s = eval_.score_candidates(tree_, fresh, carefulness); /* the whole fan-out: one batch */
  candidates_scored += fresh.size(); batches++;
  for (size_t k = 0; k < fresh.size(); k++) scores[where[k]] = std::make_pair(true, s[k]);
}

std::vector<Whisker> WhiskerImprover::early_bail_out(const std::vector<Whisker> &replacements, double carefulness, double quantile_to_keep)
{ /* src/breeder.cc:79-114 */
  std::vector<std::pair<bool, double>> scores;
  evaluate_replacements(replacements, scores, carefulness);
  std::vector<double> raw_scores;
  for (auto &x : scores) raw_scores.push_back(x.second);
  const double lower_bound = std::min(score_to_beat_ * (1 + MAX_PERCENT_ERROR), score_to_beat_ * (1 - MAX_PERCENT_ERROR));
  /* boost::accumulators tail_quantile<right> with cache_size = N and quantile_probability p:
   * the ceil(N (1 - p))-th largest sample */
  std::vector<double> sorted(raw_scores);
  std::sort(sorted.begin(), sorted.end(), std::greater<double>());
  const double p = 1 - quantile_to_keep;
  const size_t n = (size_t)std::ceil(sorted.size() * (1. - p));
  if (n < 1 || n > sorted.size()) throw std::runtime_error("early_bail_out: empty candidate set");
  const double quantile_bound = sorted[n - 1];
  const double cutoff = std::max(lower_bound, quantile_bound);
  std::vector<Whisker> top;
  for (size_t i = 0; i < replacements.size(); i++) if (raw_scores[i] >= cutoff) top.push_back(replacements[i]);
  return top;
}

double WhiskerImprover::improve(Whisker &action_to_improve)
{ /* src/breeder.cc:116-150 */
  auto replacements = get_replacements(action_to_improve);
  std::vector<std::pair<bool, double>> scores;
  /* 10 % of the simulation time first; keep the better half (and anything within 5 % of the incumbent) */
  std::vector<Whisker> top_replacements = early_bail_out(replacements, 0.1, 0.5);
  evaluate_replacements(top_replacements, scores, 1);
  for (size_t i = 0; i < top_replacements.size(); i++) {
    const Whisker &replacement = top_replacements[i];
    const bool was_new_evaluation = scores[i].first;
    const double score = scores[i].second;
    if (was_new_evaluation) eval_cache_.insert(std::make_pair(replacement, score));
    if (score > score_to_beat_) { score_to_beat_ = score; action_to_improve = replacement; }
  }
  std::cout << "Chose " << action_to_improve.str() << std::endl;
  return score_to_beat_;
}

RatBreeder::RatBreeder(const BreederOptions &o, const WhiskerImproverOptions &w, unsigned int seed)
  : _options(o), _whisker_options(w), _seed_state(0)
{
  const unsigned int s = seed ? seed : (unsigned int)(time(NULL) ^ getpid());
  _seed_state = s % 2147483647u;
  if (_seed_state == 0) _seed_state = 1;
}
unsigned int RatBreeder::next_seed()
{
  _seed_state = (uint32_t)(((uint64_t)_seed_state * 16807u) % 2147483647u);
  return _seed_state;
}

void RatBreeder::apply_best_split(WhiskerTree &tree, unsigned int generation)
{ /* src/breeder.cc:15-41 */
  const Evaluator<WhiskerTree> eval(_options.config_range, next_seed());
  auto outcome(eval.score(tree, true));
  while (1) {
    const Whisker *my_action = outcome.used_actions.most_used(generation);
    if (!my_action) throw std::runtime_error("apply_best_split: no whisker was used (reference assert, breeder.cc:25)");
    WhiskerTree bisected_action(*my_action, true);
    if (bisected_action.num_children() == 1) {
      printf("Got unbisectable whisker! %s\n", my_action->str().c_str());
      Whisker mutable_action(*my_action);
      mutable_action.promote(generation + 1);
      if (!outcome.used_actions.replace(mutable_action)) throw std::runtime_error("apply_best_split: replace failed (reference assert, breeder.cc:33)");
      continue;
    }
    printf("Bisecting === %s === into === %s ===\n", my_action->str().c_str(), bisected_action.str().c_str());
    if (!tree.replace(*my_action, bisected_action)) throw std::runtime_error("apply_best_split: replace failed (reference assert, breeder.cc:38)");
    break;
  }
}

Evaluator<WhiskerTree>::Outcome RatBreeder::improve(WhiskerTree &whiskers)
{ /* src/ratbreeder.cc:7-72 */
  WhiskerTree input_whiskertree(whiskers);
  whiskers.reset_generation();
  unsigned int generation = 0;
  while (generation < 5) {
    const Evaluator<WhiskerTree> eval(_options.config_range, next_seed());
    auto outcome(eval.score(whiskers));
    const Whisker *most_used_whisker_ptr = outcome.used_actions.most_used(generation);
    if (!most_used_whisker_ptr) { generation++; whiskers.promote(generation); continue; }
    WhiskerImprover improver(eval, whiskers, _whisker_options, outcome.score);
    Whisker whisker_to_improve = *most_used_whisker_ptr;
    double score_to_beat = outcome.score;
    while (1) {
      const double new_score = improver.improve(whisker_to_improve);
      if (new_score < score_to_beat) throw std::runtime_error("RatBreeder: score decreased (reference assert, ratbreeder.cc:41)");
      if (new_score == score_to_beat) { std::cerr << "Ending search." << std::endl; break; }
      std::cerr << "Score jumps from " << score_to_beat << " to " << new_score << std::endl;
      score_to_beat = new_score;
    }
    candidates_scored += improver.candidates_scored; batches += improver.batches;
    whisker_to_improve.demote(generation + 1);
    if (!whiskers.replace(whisker_to_improve)) throw std::runtime_error("RatBreeder: replace failed (reference assert, ratbreeder.cc:55)");
  }
  /* Split most used whisker */
  apply_best_split(whiskers, generation);
  const Evaluator<WhiskerTree> eval2(_options.config_range, next_seed());
  const auto new_score = eval2.score(whiskers, false, 10);
  const auto old_score = eval2.score(input_whiskertree, false, 10);
  if (old_score.score >= new_score.score) {
    fprintf(stderr, "Regression, old=%f, new=%f\n", old_score.score, new_score.score);
    whiskers = input_whiskertree;
    return old_score;
  }
  return new_score;
}

} // namespace remy_b200
