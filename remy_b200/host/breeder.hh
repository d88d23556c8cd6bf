/*
 * breeder.hh — the caller of the hot path: Remy's greedy action
 * optimiser on top of the batched evaluator.  Mirrors, with the same names and control flow,
 *   ActionImprover<WhiskerTree, Whisker>   reference src/breeder.hh:16-50, src/breeder.cc:43-150
 *   WhiskerImprover, RatBreeder            reference src/ratbreeder.hh, src/ratbreeder.cc:7-79
 * The one structural change: evaluate_replacements scores all uncached candidates with ONE
 * batch call (Evaluator::score_candidates) instead of one std::async thread per candidate
 * (src/breeder.cc:57-69).
 *
 * Breeder::apply_best_split (src/breeder.cc:15-41) scores the tree with trace == true
 * (remy_b200_score_trace feeds every leaf's P^2 running medians in the reference's order) and
 * bisects the most used whisker at those medians.  In the fork the function starts with a
 * stray `assert(false)` ("FIXME: Just for debugging", src/breeder.cc:18-19), so the fork's own
 * `remy` aborts here; this mirror implements what the code below that line does.
 */
#ifndef REMY_HOST_BREEDER_HH
#define REMY_HOST_BREEDER_HH

#include <unordered_map>
#include <vector>

#include "remy.hh"

namespace remy_b200 {

struct BreederOptions { ConfigRange config_range = ConfigRange(); };

struct WhiskerImproverOptions {
  bool optimize_window_increment = true;
  bool optimize_window_multiple = true;
  bool optimize_intersend = true;
};

struct WhiskerHash { size_t operator()(const Whisker &w) const; };

class WhiskerImprover {
protected:
  const double MAX_PERCENT_ERROR = 0.05;
  const Evaluator<WhiskerTree> eval_;
  WhiskerTree tree_;
  std::unordered_map<Whisker, double, WhiskerHash> eval_cache_ {};
  double score_to_beat_;
  WhiskerImproverOptions _options;

  std::vector<Whisker> get_replacements(Whisker &action_to_improve);
  /* scores[i] = (was_new_evaluation, score) for replacements[i] */
  void evaluate_replacements(const std::vector<Whisker> &replacements, std::vector<std::pair<bool, double>> &scores, double carefulness);
  std::vector<Whisker> early_bail_out(const std::vector<Whisker> &replacements, double carefulness, double quantile_to_keep);

public:
  WhiskerImprover(const Evaluator<WhiskerTree> &evaluator, const WhiskerTree &rat, const WhiskerImproverOptions &options, double score_to_beat)
    : eval_(evaluator), tree_(rat), score_to_beat_(score_to_beat), _options(options) {}
  double improve(Whisker &action_to_improve);
  /* instrumentation for tests/bench: simulations scored so far through the batch API */
  size_t candidates_scored = 0, batches = 0;
};

class RatBreeder {
  BreederOptions _options;
  WhiskerImproverOptions _whisker_options;
  uint32_t _seed_state; /* stands in for global_PRNG() (src/random.cc:7-11): one draw per Evaluator */
  unsigned int next_seed();
public:
  RatBreeder(const BreederOptions &options, const WhiskerImproverOptions &whisker_options, unsigned int seed = 0);
  Evaluator<WhiskerTree>::Outcome improve(WhiskerTree &whiskers);
  void apply_best_split(WhiskerTree &whiskers, unsigned int generation);  /* Breeder<T>::apply_best_split */
  size_t candidates_scored = 0, batches = 0;
};

} // namespace remy_b200
#endif
