/*
 * remy_b200.h — C ABI of the B200-native batched Remy evaluator.
 *
 * This is the drop-in boundary for the one hot path of CN-TU/remy that this
 * repository accelerates: scoring a WhiskerTree (or many single-leaf
 * variations of one tree) over many NetConfigs, i.e. what
 *     Evaluator<WhiskerTree>::score   (reference src/evaluator.cc:77-103)
 * does by running one Network<...>::run_simulation (src/network.cc:63-85) per
 * config, and what ActionImprover::evaluate_replacements
 * (src/breeder.cc:53-77) fans out over std::async threads.
 *
 * Everything here is plain old data: caller-owned host buffers in, caller-owned
 * host buffers out, an int return code, a per-simulation error word.  No C++
 * types, no torch types, no exceptions and no exit() cross this boundary.
 *
 * One *simulation* = one (candidate, config) pair with its own PRNG seed.  The
 * reference chains one PRNG through all configs of a score() call
 * (src/evaluator.cc:84,92-94); the batched evaluator instead takes one seed
 * per config, which is exactly the reference's public static overload
 *     score(tree, seed_k, {config_k}, trace, ticks)   (src/evaluator.hh:51-55)
 * called once per config.  All candidates of a batch share seeds[k] for config
 * k (common random numbers, like the frozen Evaluator seed, evaluator.cc:11).
 */
#ifndef REMY_B200_H
#define REMY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REMY_NUM_AXES 6      /* RemyBuffers::MemoryRange::Axis, protobufs/dna.proto:34-41 */
#define REMY_MAX_SENDERS 32  /* implementation limit of the CUDA path (reference: unbounded) */
#define REMY_NO_OVERRIDE 0xFFFFFFFFu
#define REMY_BUFFER_INFINITE 0xFFFFFFFFu

/* One network scenario.  Mirrors NetConfig (src/network.hh:16-79) after the
 * double->unsigned conversions the Network/SenderGang/Link constructors apply
 * (src/network.cc:24-37, src/link.hh:22-24), plus the run length that
 * Evaluator::score passes to run_simulation (src/evaluator.cc:94). */
typedef struct RemyNetConfig {
  double mean_on_duration;     /* ms; stop distribution rate = 1.0/mean_on  (sendergang.cc:18-19) */
  double mean_off_duration;    /* ms; start distribution rate = 1.0/mean_off */
  double link_ppt;             /* packets per ms; service time = 1.0/link_ppt (link.hh:24) */
  double delay;                /* ms; constant propagation delay (delay.hh:42-51) */
  double stochastic_loss_rate; /* Bernoulli p per packet leaving the link (stochastic-loss.hh:30-35) */
  double simulation_ticks;     /* run_simulation duration (network.cc:63-85); <= 0 is rejected */
  uint32_t num_senders;        /* 1..REMY_MAX_SENDERS */
  uint32_t buffer_size;        /* Link::_limit: 0 = no buffer, 0xFFFFFFFF = infinite (link.hh:26-34) */
} RemyNetConfig;

/* One node of the flattened WhiskerTree (src/whiskertree.hh:11-47).  Node 0 is
 * the root.  The children of a node are stored contiguously, in the order the
 * reference iterates them (first match wins, whiskertree.cc:62-82). */
typedef struct RemyTreeNode {
  double lower[REMY_NUM_AXES]; /* MemoryRange::_lower.field(a) */
  double upper[REMY_NUM_AXES]; /* MemoryRange::_upper.field(a) */
  uint32_t axis_mask;          /* bit a set <=> axis a in MemoryRange::_active_axis (memoryrange.cc:52-58) */
  uint32_t child_begin;        /* index of first child */
  uint32_t child_count;        /* 0 <=> leaf */
  uint32_t leaf_index;         /* leaves are numbered in depth-first (serialisation) order */
} RemyTreeNode;

/* Leaf action (src/whisker.hh:11-26). */
typedef struct RemyLeafAction {
  double window_multiple;
  double intersend;
  int32_t window_increment;
  uint32_t generation;         /* carried for the host; ignored by the simulator */
} RemyLeafAction;

typedef struct RemyFlatTree {
  const RemyTreeNode *nodes;
  const RemyLeafAction *leaves;
  uint32_t n_nodes;
  uint32_t n_leaves;
} RemyFlatTree;

/* A candidate = the base tree with ONE leaf's action replaced
 * (WhiskerTree::replace, whiskertree.cc:137-157).  leaf_index ==
 * REMY_NO_OVERRIDE scores the base tree itself. */
typedef struct RemyLeafOverride {
  double window_multiple;
  double intersend;
  int32_t window_increment;
  uint32_t leaf_index;
} RemyLeafOverride;

/* Per-simulation error bits (reference behaviour in brackets). */
#define REMY_ERR_NO_WHISKER     0x01u /* root domain miss  [fprintf+exit(1), whiskertree.cc:46-49] */
#define REMY_ERR_NO_CHILD       0x02u /* interior node, no child matched [assert(false), whiskertree.cc:80] */
#define REMY_ERR_LINK_OVERFLOW  0x04u /* link queue exceeded this library's ring capacity [no equivalent] */
#define REMY_ERR_DELAY_OVERFLOW 0x08u /* delay line exceeded ring capacity [no equivalent] */
#define REMY_ERR_BAD_CONFIG     0x10u /* num_senders 0 or > REMY_MAX_SENDERS, ticks <= 0, rate <= 0, a negative or NaN mean
                                       * on/off duration or both 0 [the reference's switch loop never ends, sendergang.cc:95-105] */
#define REMY_ERR_ASSERT         0x20u /* a live reference assert would have fired (memory.cc:68-69, delay.hh:47,59) */

typedef struct RemySimResult {
  double utility;            /* SenderGang::utility(): mean over senders (sendergang.cc:280-293) */
  double last_tick;          /* Network::_tickno when the loop ended */
  uint64_t event_ticks;      /* calls of Network::tick() (iterations of network.cc:73-84) */
  uint64_t packets_sent;     /* sum over senders of Rat::_packets_sent */
  uint64_t packets_received; /* sum over senders of Utility::_packets_received */
  uint32_t link_queued;      /* packets in Link::_buffer + in service at the end */
  uint32_t delay_inflight;   /* packets in the Delay queue (+ undelivered mailbox) at the end */
  uint32_t prng_state;       /* final minstd_rand0 state */
  uint32_t error;            /* REMY_ERR_* bits, 0 = ok */
} RemySimResult;

typedef struct RemySenderResult {
  double tick_share_sending; /* Utility::_tick_share_sending (utility.hh:14,19) */
  double total_delay;        /* Utility::_total_delay */
  uint32_t packets_received; /* Utility::_packets_received */
  uint32_t packets_sent;     /* Rat::_packets_sent */
} RemySenderResult;

/* ---- library lifetime ---------------------------------------------------- */

/* Bind the calling process to CUDA device `device` (one process per GPU) and
 * create the library's stream.  Returns 0, or a negative REMY_E* code. */
int remy_b200_init(int device);
/* Bind the calling process to SEVERAL devices (devices == NULL or n_devices <= 0: all visible ones).
 * Every batch created afterwards is partitioned over them: the cost-sorted list of (candidate, config)
 * simulations is dealt round-robin to the devices, each runs its share on its own stream, and the
 * results are gathered into the caller's arrays under the caller's simulation ids - what the
 * reference does with one std::async thread per candidate on the cores of one machine
 * (src/breeder.cc:53-77).  Results do not depend on the number of devices (per-simulation results are
 * returned, not partial sums).  trace == true scoring always runs on the first device of the list.
 * Replaces any earlier binding. */
int remy_b200_init_devices(const int *devices, int n_devices);
/* Number of devices bound by the last successful init (0 before any). */
int remy_b200_device_count(void);
void remy_b200_shutdown(void);
/* Human-readable description of the last failure on this thread ("" if none). */
const char *remy_b200_last_error(void);
/* ABI version of this header; bumped on any layout change. */
uint32_t remy_b200_abi_version(void);
#define REMY_B200_ABI_VERSION 2u

#define REMY_EINVAL   (-1) /* bad argument */
#define REMY_ENODEV   (-2) /* no CUDA device / init not called */
#define REMY_ECUDA    (-3) /* CUDA runtime error, see remy_b200_last_error() */
#define REMY_ENOMEM   (-4)

/* ---- one-shot scoring (host buffers in, host buffers out) ----------------- *
 * Replaces: Evaluator<WhiskerTree>::score (src/evaluator.cc:77-103) when
 * n_cands == 1, and the whole std::async fan-out of
 * ActionImprover::evaluate_replacements (src/breeder.cc:53-77) otherwise.
 *
 * Simulation s = c * n_cfgs + k runs candidate c on config k with seed
 * seeds[k].  cands == NULL means a single candidate: the base tree.
 *
 *   out_sims        [n_cands * n_cfgs]                      required
 *   out_senders     [n_cands * n_cfgs * sender_stride]      optional (NULL)
 *   out_use_counts  [n_cands * tree->n_leaves]              optional (NULL);
 *                   per-leaf use counts summed over the candidate's configs,
 *                   modulo 2^32 like the reference's `unsigned int _count`
 *                   (memoryrange.hh:27,40)
 * sender_stride must be >= max num_senders when out_senders != NULL.
 */
int remy_b200_score_batch(const RemyFlatTree *tree,
                          const RemyLeafOverride *cands, uint32_t n_cands,
                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                          uint32_t sender_stride,
                          RemySimResult *out_sims,
                          RemySenderResult *out_senders,
                          uint32_t *out_use_counts);

/* ---- staged scoring (inputs resident in HBM between runs) ----------------- *
 * create  : copy tree / candidates / configs / seeds to the device, build the
 *           config-sorted work list, size the per-slot queues.
 * run     : launch the simulation kernels on the library stream; returns after
 *           they were enqueued.  If kernel_ms != NULL the call synchronises and
 *           stores the device time of the launches (CUDA events recorded on the
 *           launching stream).
 * fetch   : synchronise and copy results to host buffers (same shapes as above).
 */
typedef struct RemyBatch RemyBatch;

int remy_b200_batch_create(const RemyFlatTree *tree,
                           const RemyLeafOverride *cands, uint32_t n_cands,
                           const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                           uint32_t sender_stride, int want_senders, int want_use_counts,
                           RemyBatch **out_batch);
int remy_b200_batch_run(RemyBatch *batch, float *kernel_ms);
int remy_b200_batch_fetch(RemyBatch *batch, RemySimResult *out_sims,
                          RemySenderResult *out_senders, uint32_t *out_use_counts);
void remy_b200_batch_destroy(RemyBatch *batch);
/* Number of kernel launches performed by the last remy_b200_batch_run (all devices). */
uint32_t remy_b200_batch_launches(const RemyBatch *batch);
/* Number of simulations the last run had to repeat with larger link rings (see REMY_ERR_LINK_OVERFLOW:
 * hidden from the caller as long as the queue fits 2^24 packets). */
uint32_t remy_b200_batch_rerun(const RemyBatch *batch);
/* Number of devices the batch was partitioned over. */
uint32_t remy_b200_batch_devices(const RemyBatch *batch);

/* ---- second sender family: Fish over a FinTree ----------------------------- *
 * Replaces: Evaluator<FinTree>::score (src/evaluator.cc:105-132): the same network with Fish
 * senders (src/fish.cc, src/fish-templates.cc) — Poisson-paced batches of five packets whose rate
 * lambda is looked up in a FinTree (src/fintree.cc) on the rtt_diff signal — instead of Rats.
 * Same arguments and outputs as remy_b200_score_batch.  The tree is the flattened FinTree; a
 * leaf's Fin::_lambda (src/fin.hh:13) travels in RemyLeafAction.intersend and a candidate's in
 * RemyLeafOverride.intersend (window_multiple / window_increment are ignored).  Each
 * simulation's Fish PRNG seed is the first draw of its run PRNG, as in evaluator.cc:113.
 */
int remy_b200_score_batch_fish(const RemyFlatTree *tree,
                               const RemyLeafOverride *cands, uint32_t n_cands,
                               const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                               uint32_t sender_stride,
                               RemySimResult *out_sims,
                               RemySenderResult *out_senders,
                               uint32_t *out_use_counts);

/* Staged form of the above (run / fetch / destroy as for remy_b200_batch_create). */
int remy_b200_batch_create_fish(const RemyFlatTree *tree,
                                const RemyLeafOverride *cands, uint32_t n_cands,
                                const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                uint32_t sender_stride, int want_senders, int want_use_counts,
                                RemyBatch **out_batch);

/* ---- flow-length workloads: Rat senders behind ByteSwitchedSender ---------------------- *
 * Replaces: Network<SenderGang<Rat, ByteSwitchedSender<Rat>>, ...>::run_simulation - the reference's
 * flow-length sender switch (src/sendergang.cc:108-138, 256-278; src/rat-templates.cc:22-26), which the
 * reference itself instantiates only for its RL sender (src/unicornevaluator.cc:109,140) but which is what
 * `on = -2` / `on = 80` mean in the shipped .cfg files under config/.  A sender is switched on after an
 * exponential idle time (mean_off_duration, ms) and switches itself off when it has SENT its flow;
 * RemyNetConfig.mean_on_duration is the mean flow length in PACKETS (length = lrint(ceil(Exp(1/mean))));
 * mean_on_duration <= 0 = unbounded flows (INT_MAX packets); mean_off_duration must be > 0 (REMY_ERR_BAD_CONFIG
 * otherwise: with an idle time of 0 the reference starts flow after flow in one instant and never leaves that tick).
 * Everything else as remy_b200_score_batch.
 * There is no trace == true variant. */
int remy_b200_score_batch_flows(const RemyFlatTree *tree,
                                const RemyLeafOverride *cands, uint32_t n_cands,
                                const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                uint32_t sender_stride,
                                RemySimResult *out_sims,
                                RemySenderResult *out_senders,
                                uint32_t *out_use_counts);
int remy_b200_batch_create_flows(const RemyFlatTree *tree,
                                 const RemyLeafOverride *cands, uint32_t n_cands,
                                 const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                 uint32_t sender_stride, int want_senders, int want_use_counts,
                                 RemyBatch **out_batch);

/* ---- trace == true scoring ------------------------------------------------ *
 * Replaces: Evaluator<WhiskerTree>::score(tree, trace = true) (src/evaluator.cc:77-103 with
 * Rat::_track set, src/rat.cc:8-20), the call Breeder::apply_best_split makes
 * (src/breeder.cc:21) to learn, for every whisker, the running medians of the Memory values
 * it was asked about (MemoryRange::track, src/memoryrange.cc:60-66; used by
 * MemoryRange::bisect, :8-41).  The estimators themselves (Boost's P^2 median, order
 * dependent) live on the host side of this boundary; the device returns what they are fed:
 * one record per WhiskerTree::use_whisker call (src/whiskertree.cc:42-60), in the exact order
 * the reference makes the calls within a simulation.
 *
 *   record = words_per_record 64-bit words: { leaf index,
 *            bit pattern of Memory::field(a) for every axis a that is active anywhere in
 *            the tree, ascending a }          (words_per_record = 1 + number of such axes)
 *
 * `sink` is called on the calling thread, with cfg_index ascending; the records of one
 * config arrive in simulation order, possibly split over several calls (the buffer is only
 * valid during the call).  A non-zero return from the sink aborts with REMY_EINVAL.
 * out_* as in remy_b200_score_batch with n_cands == 1.  If any simulation ends with an error
 * word the results are returned but no records are delivered.
 */
typedef int (*RemyTraceSink)(void *user, uint32_t cfg_index, const uint64_t *records,
                             uint64_t n_records, uint32_t words_per_record);

int remy_b200_score_trace(const RemyFlatTree *tree,
                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                          uint32_t sender_stride,
                          RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts,
                          RemyTraceSink sink, void *user);

/* The same for Fish senders over a FinTree (Evaluator<FinTree>::score(..., trace = true, ...)): a
 * Fish looks its lambda up on every ack batch and at the first send after a reset
 * (src/fish.cc:26-34, src/fish-templates.cc:14-18). */
int remy_b200_score_trace_fish(const RemyFlatTree *tree,
                               const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                               uint32_t sender_stride,
                               RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts,
                               RemyTraceSink sink, void *user);

#ifdef __cplusplus
}
#endif
#endif /* REMY_B200_H */
