/*
 * sender-runner — same argv, same stdout as the reference's src/sender-runner.cc:45-149, scoring
 * on the GPU through Evaluator<WhiskerTree> (remy.hh).  scripts/plot.py and tests/ (.pm) of the
 * reference parse this output ("sender: [tp=..., del=...]", src/sender-runner.cc:27-43).
 *
 * Reference arguments: sender=poisson (Fish over a FinTree, src/sender-runner.cc:59-68,80-87,138-141)
 *   if= nsrc= link= rtt= on= off= buf= sloss=
 * Extensions (the reference has none of these): seed=N (default: time ^ pid, src/random.cc:9),
 *   ticks=N (default 1000000, src/sender-runner.cc:57), cf=FILE (a config/ .cfg ConfigRange[Unicorn]
 *   file, mapped as in SURVEY.md Appendix A: every *.low; on <= 0 -> 5000 ms), switch=bytes (ByteSwitchedSender: on= is a
 *   mean flow length in packets, <= 0 unbounded - the meaning `on` has in the shipped .cfg files), dev=N | N,M,... | all (several GPUs share every batch).
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>

#include "fish.hh"
#include "pbwire.hh"
#include "remy.hh"

using namespace remy_b200;
using namespace std;

static string slurp(const string &path)
{
  ifstream f(path, ios::binary);
  if (!f) { perror("open"); exit(1); }
  stringstream ss; ss << f.rdbuf();
  return ss.str();
}

/* protobuf text format of a double: shortest of %.15g / %.17g that round-trips */
static string pb_double(double v)
{
  char buf[64];
  snprintf(buf, sizeof buf, "%.15g", v);
  if (strtod(buf, nullptr) != v) snprintf(buf, sizeof buf, "%.17g", v);
  return buf;
}

struct FieldName { uint32_t number; const char *name; int kind; /* 0 double, 1 varint, 2 Range, 3 OptimizationSetting */ };
static const FieldName CONFIG_RANGE[] = { {71, "link_packets_per_ms", 2}, {72, "rtt", 2}, {73, "num_senders", 2}, {74, "buffer_size", 2},
  {75, "mean_off_duration", 2}, {76, "mean_on_duration", 2}, {77, "simulation_ticks", 1}, {78, "stochastic_loss_rate", 2}, {0, nullptr, 0} };
static const FieldName RANGE[] = { {61, "low", 0}, {62, "high", 0}, {63, "incr", 0}, {0, nullptr, 0} };
static const FieldName OPT_SETTINGS[] = { {51, "window_increment", 3}, {52, "window_multiple", 3}, {53, "intersend", 3}, {54, "lambda", 3}, {0, nullptr, 0} };
static const FieldName OPT_SETTING[] = { {41, "min_value", 0}, {42, "max_value", 0}, {43, "min_change", 0}, {44, "max_change", 0},
  {45, "multiplier", 0}, {46, "default_value", 0}, {0, nullptr, 0} };

/* Message::DebugString() for the two sub-messages print_tree shows (src/sender-runner.cc:15-25) */
static string debug_string(const uint8_t *data, size_t size, const FieldName *schema, int indent)
{
  string out;
  pb::Reader rd(data, size); pb::Field f;
  while (rd.next(f)) {
    const FieldName *fn = schema;
    while (fn->name && fn->number != f.number) fn++;
    if (!fn->name) continue;
    const string pad(indent, ' ');
    if (fn->kind == 0 && f.wire_type == 1) out += pad + fn->name + ": " + pb_double(pb::Reader::as_double(f)) + "\n";
    else if (fn->kind == 1 && f.wire_type == 0) out += pad + fn->name + ": " + to_string(f.varint) + "\n";
    else if (f.wire_type == 2) out += pad + fn->name + " {\n" + debug_string(f.data, f.size, fn->kind == 3 ? OPT_SETTING : RANGE, indent + 2) + pad + "}\n";
  }
  return out;
}

static void print_tree(const string &raw, bool fin_tree)
{
  pb::Reader rd(raw); pb::Field f;
  string config, optimizer; bool has_config = false, has_opt = false;
  const uint32_t f_config = fin_tree ? 93 : 4, f_opt = fin_tree ? 94 : 5; /* protobufs/dna.proto:10-12,24-26 */
  while (rd.next(f)) {
    if (f.number == f_config && f.wire_type == 2) { config = debug_string(f.data, f.size, CONFIG_RANGE, 0); has_config = true; }
    if (f.number == f_opt && f.wire_type == 2) { optimizer = debug_string(f.data, f.size, OPT_SETTINGS, 0); has_opt = true; }
  }
  if (has_config) printf("Prior assumptions:\n%s\n\n", config.c_str());
  if (has_opt) printf("Remy optimization settings:\n%s\n\n", optimizer.c_str());
}

int main(int argc, char *argv[])
{
  WhiskerTree whiskers;
  FinTree fins;
  bool is_poisson = false;
  unsigned int num_senders = 2;
  double link_ppt = 1.0, delay = 100.0, mean_on_duration = 5000.0, mean_off_duration = 5000.0;
  double buffer_size = numeric_limits<unsigned int>::max();
  double stochastic_loss_rate = 0;
  unsigned int simulation_ticks = 1000000, seed = 0;
  string device = "0";

  for (int i = 1; i < argc; i++) {
    string arg(argv[i]);
    if (arg.substr(0, 7) == "sender=" && arg.substr(7) == "poisson") {
      is_poisson = true;
      fprintf(stderr, "Running poisson sender\n");
    }
  }
  /* key=value arguments: the numeric ones are table-driven; each echoes the reference's stderr line */
  struct NumArg { const char *key; double *dst; const char *echo; };
  double nsrc_arg = num_senders;
  const NumArg table[] = {
    { "nsrc=", &nsrc_arg, "Setting num_senders to %.0f\n" },
    { "link=", &link_ppt, "Setting link packets per ms to %f\n" },
    { "rtt=", &delay, "Setting delay to %f ms\n" },
    { "on=", &mean_on_duration, "Setting mean_on_duration to %f ms\n" },
    { "off=", &mean_off_duration, "Setting mean_off_duration to %f ms\n" },
    { "sloss=", &stochastic_loss_rate, "Setting stochastic loss rate to %f\n" },
  };
  for (int i = 1; i < argc; i++) {
    const string arg(argv[i]);
    const size_t eq = arg.find('=');
    if (eq == string::npos) continue;
    const string key = arg.substr(0, eq + 1), val = arg.substr(eq + 1);
    bool numeric = false;
    for (const NumArg &a : table) {
      if (key != a.key) continue;
      *a.dst = (a.dst == &nsrc_arg) ? (double)atoi(val.c_str()) : atof(val.c_str());
      fprintf(stderr, a.echo, *a.dst);
      numeric = true;
    }
    if (numeric) continue;
    if (key == "if=") {
      const string raw = slurp(val);
      try { if (is_poisson) fins = FinTree::parse(raw.data(), raw.size()); else whiskers = WhiskerTree::parse(raw.data(), raw.size()); }
      catch (const exception &e) { fprintf(stderr, "Could not parse %s.\n", val.c_str()); return 1; }
      print_tree(raw, is_poisson);
    } else if (key == "cf=") {
      const string raw = slurp(val);
      const ConfigRange r = ConfigRange::parse(raw.data(), raw.size());
      link_ppt = r.link_ppt.low; delay = r.rtt.low; nsrc_arg = (unsigned int)r.num_senders.low;
      buffer_size = r.buffer_size.low; mean_off_duration = r.mean_off_duration.low;
      /* time-switched senders: a non-positive "on" is a ByteSwitchedSender flow length, not a duration -> the
       * sender-runner default; with switch=bytes (given BEFORE cf=) the file's value is kept: -2 = unbounded flows */
      mean_on_duration = (r.mean_on_duration.low > 0 || remy_b200::flow_switched()) ? r.mean_on_duration.low : 5000.0;
      stochastic_loss_rate = r.stochastic_loss_rate.low;
      if (r.simulation_ticks) simulation_ticks = r.simulation_ticks;
    } else if (key == "buf=") {
      buffer_size = (val == "inf") ? (double)numeric_limits<unsigned int>::max() : (double)atoi(val.c_str());
    } else if (key == "seed=") {
      seed = (unsigned int)strtoul(val.c_str(), nullptr, 10);
    } else if (key == "ticks=") {
      simulation_ticks = (unsigned int)strtoul(val.c_str(), nullptr, 10);
    } else if (key == "switch=") {
      remy_b200::set_flow_switched(val == "bytes");
    } else if (key == "dev=") {
      device = val;
    }
  }
  num_senders = (unsigned int)nsrc_arg;

  ConfigRange configuration_range;
  configuration_range.link_ppt = Range(link_ppt, link_ppt, 0);
  configuration_range.rtt = Range(delay, delay, 0);
  configuration_range.num_senders = Range(num_senders, num_senders, 0);
  configuration_range.mean_on_duration = Range(mean_on_duration, mean_on_duration, 0);
  configuration_range.mean_off_duration = Range(mean_off_duration, mean_off_duration, 0);
  configuration_range.buffer_size = Range(buffer_size, buffer_size, 0);
  configuration_range.stochastic_loss_rate = Range(stochastic_loss_rate, stochastic_loss_rate, 0);
  configuration_range.simulation_ticks = simulation_ticks;

  if (remy_b200::init_devices(device)) { fprintf(stderr, "remy_b200_init: %s\n", remy_b200_last_error()); return 1; }
  /* parse_outcome (src/sender-runner.cc:27-43) */
  auto parse_outcome = [](double score, const vector<pair<NetConfig, vector<pair<double, double>>>> &tds, const string &rules) {
    printf("score = %f\n", score);
    double norm_score = 0;
    for (auto &run : tds) {
      printf("===\nconfig: %s\n", run.first.str().c_str());
      for (auto &x : run.second) {
        printf("sender: [tp=%f, del=%f]\n", x.first / run.first.link_ppt, x.second / run.first.delay);
        norm_score += log2(x.first / run.first.link_ppt) - log2(x.second / run.first.delay);
      }
    }
    printf("normalized_score = %f\n", norm_score);
    printf("Rules: %s\n", rules.c_str());
  };
  try {
    if (is_poisson) {
      Evaluator<FinTree> eval(configuration_range, seed);
      auto outcome = eval.score(fins, false, 10);
      parse_outcome(outcome.score, outcome.throughputs_delays, outcome.used_actions.str());
    } else {
      Evaluator<WhiskerTree> eval(configuration_range, seed);
      auto outcome = eval.score(whiskers, false, 10);
      parse_outcome(outcome.score, outcome.throughputs_delays, outcome.used_actions.str());
    }
  } catch (const exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
