/*
 * remy.hh — C++ host mirror of the reference's scoring API (value types + Evaluator),
 * protobuf-free, on top of the C ABI in include/remy_b200.h.
 *
 * Same names, argument meaning and error behaviour as the reference classes they stand in
 * for (file:line citations are relative to /root/reference):
 *   Range, ConfigRange            src/configrange.hh:5-51   (+ the ConfigRangeUnicorn wire form)
 *   NetConfig                     src/network.hh:16-79
 *   Memory, MemoryRange           src/memory.hh:20-110, src/memoryrange.hh:15-60
 *   Whisker                       src/whisker.hh:11-67, src/whisker.cc:46-95, src/action.hh:12-105
 *   WhiskerTree                   src/whiskertree.hh:11-47, src/whiskertree.cc
 *   Evaluator<WhiskerTree>        src/evaluator.hh:14-56, src/evaluator.cc:9-38,77-103,196-201
 *   trace == true scoring         src/memoryrange.cc:8-41,60-66 (track, bisect), src/action.hh:25-33,
 *                                 src/whiskertree.cc:17-30,159-180, on remy_b200_score_trace
 * FinTree / Fin / Evaluator<FinTree> / FishBreeder live in fish.{hh,cc}; RatBreeder and the improvers in breeder.{hh,cc}.
 * Process-wide switches at the end of this header: flow-switched senders (ByteSwitchedSender), the reference's chained
 * PRNG, the device list.
 */
#ifndef REMY_HOST_REMY_HH
#define REMY_HOST_REMY_HH

#include <cstdint>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "../../include/remy_b200.h"

namespace remy_b200 {

/* ---- src/configrange.hh ---- */
class Range {
public:
  double low = 1, high = 1, incr = 0;
  Range() {}
  Range(double lo, double hi, double inc) : low(lo), high(hi), incr(inc) {}
  bool isOne() const { return (low == high) || (incr == 0); }
  std::string DNA() const;                       /* RemyBuffers.Range bytes */
  static Range parse(const void *data, size_t size);
};

class ConfigRange {
public:
  Range link_ppt, rtt, mean_on_duration, mean_off_duration, num_senders, buffer_size;
  unsigned int simulation_ticks = 1000000;
  Range stochastic_loss_rate { 0, 0, 0 };
  ConfigRange() {}
  /* Accepts RemyBuffers.ConfigRange and ConfigRangeUnicorn bytes (field 77 is a uint32 in
   * the former and a Range in the latter; every file under config/ is the latter). */
  static ConfigRange parse(const void *data, size_t size);
  std::string DNA() const;                       /* RemyBuffers.ConfigRange bytes, protoc field order */
};

/* ---- src/network.hh ---- */
class NetConfig {
public:
  double mean_on_duration = 5000.0, mean_off_duration = 5000.0;
  double num_senders = 8, link_ppt = 1.0, delay = 150;
  double buffer_size = 4294967295.0;
  double stochastic_loss_rate = 0, simulation_ticks = 0;
  NetConfig &set_link_ppt(double v) { link_ppt = v; return *this; }
  NetConfig &set_delay(double v) { delay = v; return *this; }
  NetConfig &set_num_senders(unsigned int n) { num_senders = n; return *this; }
  NetConfig &set_on_duration(double v) { mean_on_duration = v; return *this; }
  NetConfig &set_off_duration(double v) { mean_off_duration = v; return *this; }
  NetConfig &set_buffer_size(unsigned int n) { buffer_size = n; return *this; }
  NetConfig &set_stochastic_loss_rate(double v) { stochastic_loss_rate = v; return *this; }
  NetConfig &set_simulation_ticks(unsigned int n) { simulation_ticks = n; return *this; }
  std::string str() const;
  std::string DNA() const;                         /* RemyBuffers.NetConfig bytes (src/network.hh:57-70) */
  static NetConfig parse(const void *data, size_t size);
  RemyNetConfig pod(double ticks_to_run) const;  /* the conversions of src/network.cc:24-37 */
};

/* ---- src/memory.hh / src/memoryrange.hh ---- */
class Memory {
public:
  static const unsigned int datasize = 6;
  double f[datasize] = { 0, 0, 0, 0, 0, 0 };     /* field(0..5), src/memory.hh:97 */
  const double &field(unsigned int i) const { return f[i]; }
  double &mutable_field(unsigned int i) { return f[i]; }
  std::string str() const;
  std::string str(unsigned int num) const;
  bool operator==(const Memory &o) const { for (unsigned i = 0; i < datasize; i++) if (f[i] != o.f[i]) return false; return true; }
};
const Memory &MAX_MEMORY();

/* boost::accumulators::accumulator_set<double, stats<tag::median>> (src/memoryrange.hh:24-26):
 * Boost's tag::median is the P^2 running estimator with p = 0.5 (p_square_quantile_impl), not an
 * exact median, and its value depends on the order of the samples. */
class P2Median {
  double _h[5], _n[5], _d[5];
  size_t _cnt;
public:
  P2Median();
  void operator()(double sample);
  double median() const { return _h[2]; }
  size_t count() const { return _cnt; }
};

class MemoryRange {
public:
  Memory _lower, _upper;
  std::vector<int> _active_axis { 0, 1, 2, 3 };
  bool _lower_has[Memory::datasize] = { true, true, true, true, true, true };  /* for byte-identical re-serialisation */
  bool _upper_has[Memory::datasize] = { true, true, true, true, true, true };
  bool _axes_explicit = true;
  mutable unsigned int _count = 0;
  mutable std::vector<P2Median> _acc = std::vector<P2Median>(Memory::datasize);
  MemoryRange() : _upper(MAX_MEMORY()) {}
  MemoryRange(const Memory &lo, const Memory &hi, std::vector<int> active = { 0, 1, 2, 3 }) : _lower(lo), _upper(hi), _active_axis(active) {}
  std::vector<MemoryRange> bisect() const;                /* src/memoryrange.cc:8-41 */
  void track(const Memory &q) const;                      /* src/memoryrange.cc:60-66 */
  bool contains(const Memory &q) const;
  Memory range_median() const;
  void use() const { _count++; }
  unsigned int count() const { return _count; }
  void reset_count() const { _count = 0; }
  bool operator==(const MemoryRange &o) const;
  std::string str() const;
  std::string DNA() const;
  static MemoryRange parse(const void *data, size_t size);
  uint32_t axis_mask() const { uint32_t m = 0; for (int a : _active_axis) m |= 1u << a; return m; }
};

/* ---- src/action.hh / src/whisker.hh ---- */
template <typename T> struct OptimizationSetting {
  T min_value, max_value, min_change, max_change, multiplier, default_value;
  bool eligible_value(const T &v) const { return v >= min_value && v <= max_value; }
  std::vector<T> alternatives(const T &value, bool active) const;   /* src/action.hh:62-91 */
};

class Whisker {
public:
  unsigned int _generation = 0;
  MemoryRange _domain;
  int _window_increment = 1;
  double _window_multiple = 1, _intersend = 3;
  Whisker() {}
  Whisker(unsigned int incr, double mult, double intersend, const MemoryRange &dom)
    : _domain(dom), _window_increment((int)incr), _window_multiple(mult), _intersend(intersend) {}
  explicit Whisker(const MemoryRange &dom);       /* default action, src/whisker.hh:20-22 */
  unsigned int window(unsigned int previous_window) const;            /* src/whisker.hh:25 */
  const double &intersend() const { return _intersend; }
  int window_increment() const { return _window_increment; }
  double window_multiple() const { return _window_multiple; }
  const MemoryRange &domain() const { return _domain; }
  const unsigned int &generation() const { return _generation; }
  unsigned int count() const { return _domain.count(); }
  void promote(unsigned int g) { if (g > _generation) _generation = g; }
  void demote(unsigned int g) { _generation = g; }
  std::vector<Whisker> next_generation(bool optimize_window_increment, bool optimize_window_multiple, bool optimize_intersend) const;
  std::vector<Whisker> bisect() const;                    /* Action::bisect<Whisker>, src/action.hh:25-33 */
  void round();
  std::string str(unsigned int total = 1) const;
  bool operator==(const Whisker &o) const { return _window_increment == o._window_increment && _window_multiple == o._window_multiple && _intersend == o._intersend && _domain == o._domain; }
  struct OptimizationSettings { OptimizationSetting<unsigned int> window_increment; OptimizationSetting<double> window_multiple, intersend; };
  static const OptimizationSettings &get_optimizer();                 /* src/whisker.hh:59-66 */
  std::string DNA() const;
  static Whisker parse(const void *data, size_t size);
};

/* ---- src/whiskertree.hh ---- */
class WhiskerTree {
public:
  MemoryRange _domain;
  std::vector<WhiskerTree> _children;
  std::vector<Whisker> _leaf;
  std::string _extra;                             /* fields 4-6 (config / optimizer / configvector), kept verbatim */
  WhiskerTree();                                  /* one default whisker over [0, MAX_MEMORY) */
  WhiskerTree(const Whisker &whisker, bool bisect); /* src/whiskertree.cc:17-30 */
  bool is_leaf() const { return !_leaf.empty(); }
  unsigned int num_children() const { return is_leaf() ? 1 : (unsigned)_children.size(); }
  const Whisker &use_whisker(const Memory &m, bool track) const;      /* exits like the reference on a miss */
  bool replace(const Whisker &w);
  bool replace(const Whisker &src, const WhiskerTree &dst);           /* leaf -> interior node, src/whiskertree.cc:159-180 */
  const Whisker *most_used(unsigned int max_generation) const;
  void reset_counts();
  void promote(unsigned int generation);
  void reset_generation();
  unsigned int total_whisker_queries() const;
  std::string str() const;
  std::string str(unsigned int total) const;
  std::string DNA() const;                        /* RemyBuffers.WhiskerTree bytes */
  static WhiskerTree parse(const void *data, size_t size);
  /* leaves in depth-first (serialisation) order == RemyTreeNode::leaf_index order */
  void leaves(std::vector<const Whisker *> &out) const;
  void leaves(std::vector<Whisker *> &out);
  /* index of the leaf WhiskerTree::replace(w) would overwrite, or -1 */
  int leaf_index_of(const Whisker &w) const;
private:
  const Whisker *whisker(const Memory &m) const;
};

/* Feeds records of remy_b200_score_trace (include/remy_b200.h) to the running medians of the leaves they
 * name: MemoryRange::track for every record, in order (several host threads, one per subset of the
 * independent (leaf, axis) estimators).  Returns non-zero on a malformed record. */
int feed_trace_records(WhiskerTree &tree, const uint64_t *records, uint64_t n_records, uint32_t words_per_record);
/* the same for any tree type: the leaves' domains depth-first + the union of the axes active anywhere in the tree */
int feed_trace_records(const std::vector<const MemoryRange *> &leaf_domains, uint32_t all_axes_mask,
                       const uint64_t *records, uint64_t n_records, uint32_t words_per_record);

/* POD view of a WhiskerTree for the C ABI. */
struct FlatTree {
  std::vector<RemyTreeNode> nodes; std::vector<RemyLeafAction> leaves;
  RemyFlatTree c() const { RemyFlatTree t; t.nodes = nodes.data(); t.leaves = leaves.data(); t.n_nodes = (uint32_t)nodes.size(); t.n_leaves = (uint32_t)leaves.size(); return t; }
};
FlatTree flatten(const WhiskerTree &tree);

/* ---- src/evaluator.hh ---- */
template <typename T> class Evaluator;

template <> class Evaluator<WhiskerTree> {
public:
  class Outcome {
  public:
    double score = 0;
    std::vector<std::pair<NetConfig, std::vector<std::pair<double, double>>>> throughputs_delays;
    WhiskerTree used_actions;
    Outcome() {}
    std::string DNA() const;                              /* AnswerBuffers.Outcome bytes (src/evaluator.cc:162-180) */
    static Outcome parse(const void *data, size_t size);  /* Outcome(AnswerBuffers::Outcome) (src/evaluator.cc:182-194) */
  };
private:
  unsigned int _prng_seed;
  unsigned int _tick_count;
  std::vector<NetConfig> _configs;
public:
  /* freezes a PRNG seed for the life of the Evaluator (src/evaluator.cc:11); pass one for
   * reproducible runs, 0 draws it from time ^ pid like global_PRNG (src/random.cc:7-11) */
  explicit Evaluator(const ConfigRange &range, unsigned int prng_seed = 0);
  const std::vector<NetConfig> &configs() const { return _configs; }
  unsigned int prng_seed() const { return _prng_seed; }
  Outcome score(WhiskerTree &run_actions, bool trace = false, double carefulness = 1) const;
  static Outcome score(WhiskerTree &run_actions, unsigned int prng_seed, const std::vector<NetConfig> &configs,
                       bool trace, unsigned int ticks_to_run);
  /* The data-parallel replacement for ActionImprover::evaluate_replacements
   * (src/breeder.cc:53-77): scores[c] = score of `tree` with replacements[c] applied. */
  std::vector<double> score_candidates(const WhiskerTree &tree, const std::vector<Whisker> &replacements, double carefulness = 1) const;
  /* ProblemBuffers.Problem bytes: settings {prng_seed, tick_count}, configs, whiskers (src/evaluator.cc:40-66) */
  std::string DNA(const WhiskerTree &whiskers) const;
  /* src/evaluator.cc:134-146 */
  static Outcome parse_problem_and_evaluate(const void *problem, size_t size);
  /* per-config seeds of a score() call: prng_seed itself for config 0 (as in the reference), then the
   * successive outputs of minstd_rand0(prng_seed) */
  static std::vector<uint32_t> seeds(unsigned int prng_seed, size_t n);
};

/* Sender switching of every Evaluator<WhiskerTree> of this process: false (default) = TimeSwitchedSender, exponential on
 * and off TIMES, what the reference's Evaluator builds (src/evaluator.cc:92-93); true = ByteSwitchedSender
 * (src/sendergang.cc:108-138, 256-278): mean_on_duration is a mean flow length in PACKETS (<= 0: unbounded), which is what
 * `on` means in the shipped .cfg files under config/.  CLIs: switch=bytes.  trace == true scoring exists for the default only. */
void set_flow_switched(bool on);
bool flow_switched();

/* PRNG chaining of every Evaluator<WhiskerTree> of this process.  false (default): every config of a score() call runs on
 * its own seed (seeds()), so that all simulations of a batch are independent - the batched evaluator.  true: the reference's
 * literal behaviour - ONE PRNG handed from config to config (src/evaluator.cc:84,92-94), i.e. config k starts where config
 * k-1 stopped; a multi-config score() then reproduces the reference's number exactly, at the price of running the configs
 * one after the other (each on a single GPU thread).  For compatibility checks; candidates are still scored unchained. */
void set_chained_prng(bool on);
bool chained_prng();

/* dev= option of the CLIs: "N" binds one CUDA device, "N,M,..." or "all" several - every batch (an
 * evaluator pass, the candidates of an improve step) is then partitioned over them by the library
 * (remy_b200_init_devices), the way the reference spreads an improve step over the cores of one machine
 * (src/breeder.cc:53-77).  Returns the library's return code. */
inline int init_devices(const std::string &spec)
{
  if (spec == "all") return remy_b200_init_devices(nullptr, 0);
  std::vector<int> devs;
  size_t pos = 0;
  while (pos <= spec.size()) {
    const size_t comma = spec.find(',', pos);
    const std::string tok = spec.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
    if (!tok.empty()) devs.push_back(atoi(tok.c_str()));
    if (comma == std::string::npos) break;
    pos = comma + 1;
  }
  if (devs.empty()) devs.push_back(0);
  return remy_b200_init_devices(devs.data(), (int)devs.size());
}

} // namespace remy_b200
#endif
