/*
 * fish.hh — host mirror of the second sender family's policy types, on top
 * of remy_b200_score_batch_fish.  Same names, argument meaning and error behaviour as
 *   Fin                     reference src/fin.hh:11-57, src/fin.cc
 *   FinTree                 reference src/fintree.hh, src/fintree.cc
 *   Evaluator<FinTree>      reference src/evaluator.hh:14-56, src/evaluator.cc:105-132,148-160
 *   FinImprover, FishBreeder reference src/fishbreeder.{hh,cc} over ActionImprover / Breeder (src/breeder.cc:15-150)
 */
#ifndef REMY_HOST_FISH_HH
#define REMY_HOST_FISH_HH

#include <unordered_map>

#include "remy.hh"
#include "breeder.hh"

namespace remy_b200 {

class Fin {
public:
  unsigned int _generation = 0;
  MemoryRange _domain;
  double _lambda = 5;
  Fin() {}
  Fin(double rate, const MemoryRange &dom) : _domain(dom), _lambda(rate) {}
  explicit Fin(const MemoryRange &dom) : Fin(get_optimizer().lambda.default_value, dom) {}
  const double &lambda() const { return _lambda; }
  const MemoryRange &domain() const { return _domain; }
  const unsigned int &generation() const { return _generation; }
  unsigned int count() const { return _domain.count(); }
  void promote(unsigned int g) { if (g > _generation) _generation = g; }
  void demote(unsigned int g) { _generation = g; }
  std::vector<Fin> next_generation() const;                 /* src/fin.cc:33-54 */
  std::vector<Fin> bisect() const;                          /* Action::bisect<Fin>, src/action.hh:25-33 */
  void round();                                             /* src/fin.cc:64-67 */
  std::string str(unsigned int total = 1) const;
  bool operator==(const Fin &o) const { return _lambda == o._lambda && _domain == o._domain; }
  struct OptimizationSettings { OptimizationSetting<double> lambda; };
  static const OptimizationSettings &get_optimizer();       /* src/fin.hh:49-54 */
  std::string DNA() const;                                  /* RemyBuffers.Fin bytes */
  static Fin parse(const void *data, size_t size);
};

class FinTree {
public:
  MemoryRange _domain;
  std::vector<FinTree> _children;
  std::vector<Fin> _leaf;
  std::string _extra;                                       /* fields 93-95 (config / optimizer / configvector), kept verbatim */
  FinTree();                                                /* one default fin over [0, MAX_MEMORY) on rtt_diff */
  FinTree(const Fin &fin, bool bisect);
  bool is_leaf() const { return !_leaf.empty(); }
  unsigned int num_children() const { return is_leaf() ? 1 : (unsigned)_children.size(); }
  const Fin &use_fin(const Memory &m, bool track) const;    /* throws like the reference exits on a miss */
  bool replace(const Fin &w);
  bool replace(const Fin &src, const FinTree &dst);
  const Fin *most_used(unsigned int max_generation) const;
  void reset_counts();
  void promote(unsigned int generation);
  void reset_generation();
  unsigned int total_fin_queries() const;
  std::string str() const;
  std::string str(unsigned int total) const;
  std::string DNA() const;                                  /* RemyBuffers.FinTree bytes */
  static FinTree parse(const void *data, size_t size);
  void leaves(std::vector<const Fin *> &out) const;
  void leaves(std::vector<Fin *> &out);
  int leaf_index_of(const Fin &w) const;
private:
  const Fin *fin(const Memory &m) const;
};

/* POD view for the C ABI: a leaf's lambda travels in RemyLeafAction.intersend */
FlatTree flatten(const FinTree &tree);

template <> class Evaluator<FinTree> {
public:
  class Outcome {
  public:
    double score = 0;
    std::vector<std::pair<NetConfig, std::vector<std::pair<double, double>>>> throughputs_delays;
    FinTree used_actions;
  };
private:
  unsigned int _prng_seed, _tick_count;
  std::vector<NetConfig> _configs;
public:
  explicit Evaluator(const ConfigRange &range, unsigned int prng_seed = 0);
  const std::vector<NetConfig> &configs() const { return _configs; }
  Outcome score(FinTree &run_actions, bool trace = false, double carefulness = 1) const;
  static Outcome score(FinTree &run_actions, unsigned int prng_seed, const std::vector<NetConfig> &configs,
                       bool trace, unsigned int ticks_to_run);
  /* the batch replacement for ActionImprover::evaluate_replacements (src/breeder.cc:53-77) */
  std::vector<double> score_candidates(const FinTree &tree, const std::vector<Fin> &replacements, double carefulness = 1) const;
};

struct FinHash { size_t operator()(const Fin &f) const; };

class FinImprover {
  const double MAX_PERCENT_ERROR = 0.05;
  const Evaluator<FinTree> eval_;
  FinTree tree_;
  std::unordered_map<Fin, double, FinHash> eval_cache_ {};
  double score_to_beat_;
  void evaluate_replacements(const std::vector<Fin> &replacements, std::vector<std::pair<bool, double>> &scores, double carefulness);
  std::vector<Fin> early_bail_out(const std::vector<Fin> &replacements, double carefulness, double quantile_to_keep);
public:
  FinImprover(const Evaluator<FinTree> &evaluator, const FinTree &fish, double score_to_beat)
    : eval_(evaluator), tree_(fish), score_to_beat_(score_to_beat) {}
  double improve(Fin &action_to_improve);                   /* src/breeder.cc:116-150 */
  size_t candidates_scored = 0, batches = 0;
};

class FishBreeder {
  BreederOptions _options;
  uint32_t _seed_state; /* stands in for global_PRNG() (src/random.cc:7-11): one draw per Evaluator */
  unsigned int next_seed();
public:
  explicit FishBreeder(const BreederOptions &options, unsigned int seed = 0);
  Evaluator<FinTree>::Outcome improve(FinTree &fins);                 /* src/fishbreeder.cc:7-72 */
  void apply_best_split(FinTree &fins, unsigned int generation);      /* Breeder<FinTree>::apply_best_split */
  size_t candidates_scored = 0, batches = 0;
};

} // namespace remy_b200
#endif
