/*
 * remy — the optimiser CLI of the reference (src/remy.cc:20-185) on the GPU path: same key=value
 * arguments (if= of= opt= cf=), same stdout, same checkpoint files (`<of>.<run>` = the WhiskerTree
 * with its config, optimizer and configvector sub-messages).  cf= accepts ConfigRange and
 * ConfigRangeUnicorn files (the reference's own remy silently reads 0 ticks from the latter,
 * SURVEY.md Appendix A).  Every run optimises the whisker actions and then bisects the most used
 * whisker (RatBreeder::improve incl. apply_best_split, see breeder.hh).
 * Extensions: runs=N (stop after N improve() rounds; default: forever, like the reference),
 *             seed=N (default time ^ pid), switch=bytes (flow-switched senders, see sender_runner.cc), dev=N | N,M,... | all (several GPUs share every batch).
 */
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>

#include "breeder.hh"
#include "pbwire.hh"

using namespace remy_b200;
using namespace std;

static void print_range(const Range &range, const string &name)
{ /* src/remy.cc:14-18 */
  printf("Optimizing for %s over [%f : %f : %f]\n", name.c_str(), range.low, range.incr, range.high);
}

static string optimizer_dna()
{ /* Whisker::OptimizationSettings::DNA (src/whisker.hh:47-56): fields 51..53, each six doubles 41..46 */
  const auto &o = Whisker::get_optimizer();
  auto one = [](double mn, double mx, double mc, double xc, double mu, double df) {
    pb::Writer w; w.put_double(41, mn); w.put_double(42, mx); w.put_double(43, mc); w.put_double(44, xc); w.put_double(45, mu); w.put_double(46, df);
    return w.str(); };
  pb::Writer w;
  w.put_message(51, one(o.window_increment.min_value, o.window_increment.max_value, o.window_increment.min_change,
                        o.window_increment.max_change, o.window_increment.multiplier, o.window_increment.default_value));
  w.put_message(52, one(o.window_multiple.min_value, o.window_multiple.max_value, o.window_multiple.min_change,
                        o.window_multiple.max_change, o.window_multiple.multiplier, o.window_multiple.default_value));
  w.put_message(53, one(o.intersend.min_value, o.intersend.max_value, o.intersend.min_change, o.intersend.max_change,
                        o.intersend.multiplier, o.intersend.default_value));
  return w.str();
}

int main(int argc, char *argv[])
{
  WhiskerTree whiskers;
  string output_filename, config_filename;
  BreederOptions options;
  WhiskerImproverOptions whisker_options;
  unsigned int seed = 0;
  long max_runs = -1;
  string device = "0";

  for (int i = 1; i < argc; i++) {
    const string arg(argv[i]);
    const size_t eq = arg.find('=');
    if (eq == string::npos) continue;
    const string key = arg.substr(0, eq + 1), val = arg.substr(eq + 1);
    if (key == "if=" || key == "cf=") {
      ifstream f(val, ios::binary);
      if (!f) { perror(key == "if=" ? "open" : "open config file error"); return 1; }
      stringstream ss; ss << f.rdbuf();
      const string raw = ss.str();
      try {
        if (key == "if=") { whiskers = WhiskerTree::parse(raw.data(), raw.size()); whiskers._extra.clear(); }
        else { options.config_range = ConfigRange::parse(raw.data(), raw.size()); config_filename = val; }
      } catch (const exception &e) {
        fprintf(stderr, key == "if=" ? "Could not parse %s.\n" : "Could not parse input config from file %s. \n", val.c_str());
        return 1;
      }
    } else if (key == "of=") {
      output_filename = val;
    } else if (key == "opt=") {
      whisker_options.optimize_window_increment = whisker_options.optimize_window_multiple = whisker_options.optimize_intersend = false;
      for (char c : val) {
        if (c == 'b') whisker_options.optimize_window_increment = true;
        else if (c == 'm') whisker_options.optimize_window_multiple = true;
        else if (c == 'r') whisker_options.optimize_intersend = true;
        else { fprintf(stderr, "Invalid optimize option: %c\n", c); return 1; }
      }
    } else if (key == "runs=") max_runs = atol(val.c_str());
    else if (key == "seed=") seed = (unsigned int)strtoul(val.c_str(), nullptr, 10);
    else if (key == "dev=") device = val;
    else if (key == "switch=") remy_b200::set_flow_switched(val == "bytes");
  }
  if (config_filename.empty()) {
    fprintf(stderr, "An input configuration protobuf must be provided via the cf= option. \n");
    fprintf(stderr, "You can generate one using './configuration'. \n");
    return 1;
  }
  /* The shipped .cfg files under config/ are ConfigRangeUnicorn files whose field 76 is a flow length in packets for
   * ByteSwitchedSender, -2 meaning unbounded (src/sendergang.cc:128-134).  The Rat path switches senders in time
   * (TimeSwitchedSender): a non-positive "on" would never let the switch loop end, so it is mapped to the
   * sender-runner default of 5000 ms (src/sender-runner.cc:53), like sender_runner.cc does for cf=. */
  if (!remy_b200::flow_switched() && options.config_range.mean_on_duration.low <= 0) {
    fprintf(stderr, "mean_on_duration %f..%f is not a duration (a ByteSwitchedSender flow length?): using 5000 ms\n",
            options.config_range.mean_on_duration.low, options.config_range.mean_on_duration.high);
    options.config_range.mean_on_duration = Range(5000, 5000, 0);
  }
  if (remy_b200::init_devices(device)) { fprintf(stderr, "remy_b200_init: %s\n", remy_b200_last_error()); return 1; }

  RatBreeder breeder(options, whisker_options, seed);
  printf("#######################\n");
  printf("Evaluator simulations will run for %d ticks\n", options.config_range.simulation_ticks);
  printf("Optimizing window increment: %d, window multiple: %d, intersend: %d\n", whisker_options.optimize_window_increment,
         whisker_options.optimize_window_multiple, whisker_options.optimize_intersend);
  print_range(options.config_range.link_ppt, "link packets_per_ms");
  print_range(options.config_range.rtt, "rtt_ms");
  print_range(options.config_range.num_senders, "num_senders");
  print_range(options.config_range.mean_on_duration, "mean_on_duration");
  print_range(options.config_range.mean_off_duration, "mean_off_duration");
  print_range(options.config_range.stochastic_loss_rate, "stochastic_loss_rate");
  if (options.config_range.buffer_size.low != numeric_limits<unsigned int>::max()) print_range(options.config_range.buffer_size, "buffer_size");
  else printf("Optimizing for infinitely sized buffers. \n");
  printf("Initial rules (use if=FILENAME to read from disk): %s\n", whiskers.str().c_str());
  printf("#######################\n");
  if (!output_filename.empty()) printf("Writing to \"%s.N\".\n", output_filename.c_str());
  else printf("Not saving output. Use the of=FILENAME argument to save the results.\n");

  string training_configs; /* RemyBuffers.ConfigVector body: repeated NetConfig, field 81 */
  bool written = false;
  try {
    for (unsigned int run = 0; max_runs < 0 || (long)run < max_runs; run++) {
      auto outcome = breeder.improve(whiskers);
      printf("run = %u, score = %f\n", run, outcome.score);
      printf("whiskers: %s\n", whiskers.str().c_str());
      if (!written) {
        pb::Writer cv;
        for (auto &r : outcome.throughputs_delays) cv.put_message(81, r.first.DNA());
        training_configs = cv.str();
        written = true;
      }
      for (auto &r : outcome.throughputs_delays) {
        printf("===\nconfig: %s\n", r.first.str().c_str());
        for (auto &x : r.second) printf("sender: [tp=%f, del=%f]\n", x.first / r.first.link_ppt, x.second / r.first.delay);
      }
      if (!output_filename.empty()) {
        char of[128];
        snprintf(of, 128, "%s.%d", output_filename.c_str(), run);
        fprintf(stderr, "Writing to \"%s\"... ", of);
        pb::Writer extra;
        extra.put_message(4, options.config_range.DNA());
        extra.put_message(5, optimizer_dna());
        extra.put_message(6, training_configs);
        WhiskerTree remycc(whiskers);
        remycc._extra = extra.str();
        ofstream out(of, ios::binary | ios::trunc);
        if (!out) { perror("open"); return 1; }
        const string bytes = remycc.DNA();
        out.write(bytes.data(), bytes.size());
        if (!out) { fprintf(stderr, "Could not serialize RemyCC.\n"); return 1; }
        fprintf(stderr, "done.\n");
      }
      fflush(NULL);
    }
  } catch (const exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
