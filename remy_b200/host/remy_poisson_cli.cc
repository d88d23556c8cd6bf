/*
 * remy-poisson — the Fish optimiser CLI of the reference (src/remy-poisson.cc) on the GPU path: same
 * key=value arguments (if= of= cf=), same stdout, same checkpoint files (`<of>.<run>` = the FinTree with
 * its config (93), optimizer (94) and configvector (95) sub-messages).  Every run optimises the fins'
 * lambdas and then bisects the most used fin (FishBreeder::improve, fish.hh).
 * Extensions: runs=N (default: forever, like the reference), seed=N (default time ^ pid), dev=N | N,M,... | all (several GPUs share every batch).
 */
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>

#include "fish.hh"
#include "pbwire.hh"

using namespace remy_b200;
using namespace std;

static string optimizer_dna()
{ /* Fin::OptimizationSettings::DNA (src/fin.hh:38-47): field 54, six doubles 41..46 */
  const auto &o = Fin::get_optimizer().lambda;
  pb::Writer s; s.put_double(41, o.min_value); s.put_double(42, o.max_value); s.put_double(43, o.min_change);
  s.put_double(44, o.max_change); s.put_double(45, o.multiplier); s.put_double(46, o.default_value);
  pb::Writer w; w.put_message(54, s.str());
  return w.str();
}

int main(int argc, char *argv[])
{
  FinTree fins;
  string output_filename, config_filename;
  BreederOptions options;
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
        if (key == "if=") { fins = FinTree::parse(raw.data(), raw.size()); fins._extra.clear(); }
        else { options.config_range = ConfigRange::parse(raw.data(), raw.size()); config_filename = val; }
      } catch (const exception &e) {
        fprintf(stderr, key == "if=" ? "Could not parse %s.\n" : "Could not parse input config from file %s. \n", val.c_str());
        return 1;
      }
    } else if (key == "of=") output_filename = val;
    else if (key == "runs=") max_runs = atol(val.c_str());
    else if (key == "seed=") seed = (unsigned int)strtoul(val.c_str(), nullptr, 10);
    else if (key == "dev=") device = val;
  }
  if (config_filename.empty()) {
    fprintf(stderr, "An input configuration protobuf must be provided via the cf= option. \n");
    fprintf(stderr, "You can generate one using './configuration'. \n");
    return 1;
  }
  if (remy_b200::init_devices(device)) { fprintf(stderr, "remy_b200_init: %s\n", remy_b200_last_error()); return 1; }

  FishBreeder breeder(options, seed);
  const ConfigRange &cr = options.config_range;
  printf("#######################\n");
  printf("Evaluator simulations will run for %d ticks\n", cr.simulation_ticks);
  printf("Optimizing for link packets_per_ms in [%f, %f]\n", cr.link_ppt.low, cr.link_ppt.high);
  printf("Optimizing for rtt_ms in [%f, %f]\n", cr.rtt.low, cr.rtt.high);
  printf("Optimizing for num_senders in [%f, %f]\n", cr.num_senders.low, cr.num_senders.high);
  printf("Optimizing for mean_on_duration in [%f, %f], mean_off_duration in [ %f, %f]\n", cr.mean_on_duration.low, cr.mean_on_duration.high,
         cr.mean_off_duration.low, cr.mean_off_duration.high);
  printf("Optimizing for stochastic_loss_rate in [%f, %f]\n", cr.stochastic_loss_rate.low, cr.stochastic_loss_rate.high);
  if (cr.buffer_size.low != numeric_limits<unsigned int>::max()) printf("Optimizing for buffer_size in [%f, %f]\n", cr.buffer_size.low, cr.buffer_size.high);
  else printf("Optimizing for infinitely sized buffers. \n");
  printf("Initial rules (use if=FILENAME to read from disk): %s\n", fins.str().c_str());
  printf("#######################\n");
  if (!output_filename.empty()) printf("Writing to \"%s.N\".\n", output_filename.c_str());
  else printf("Not saving output. Use the of=FILENAME argument to save the results.\n");

  string training_configs; /* RemyBuffers.ConfigVector body: repeated NetConfig, field 81 */
  bool written = false;
  try {
    for (unsigned int run = 0; max_runs < 0 || (long)run < max_runs; run++) {
      auto outcome = breeder.improve(fins);
      printf("run = %u, score = %f\n", run, outcome.score);
      printf("fins: %s\n", fins.str().c_str());
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
        extra.put_message(93, cr.DNA());
        extra.put_message(94, optimizer_dna());
        extra.put_message(95, training_configs);
        FinTree remycc(fins);
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
