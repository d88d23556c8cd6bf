/*
 * sim_kernel.cuh — the sm_100a simulation kernel: one thread per (candidate,
 * config) simulation, persistent blocks pulling work from a cost-sorted list.
 *
 * What it computes is exactly what the reference computes for
 *   Evaluator<WhiskerTree>::score(tree, seed, {config}, false, ticks)
 * (reference src/evaluator.cc:77-103 -> src/network.cc:63-85), but the loop is
 * re-organised for a GPU thread instead of translated:
 *
 *   reference (per event tick)                    this kernel
 *   ---------------------------------------------------------------------------
 *   2x count_active_senders O(n)                  popc of a 32-bit on-mask
 *   switch_senders: n switcher() calls            one compare against a cached
 *     (sendergang.cc:39-45,89-106)                min-switch time; O(n) scan only
 *                                                 on the rare ticks that switch
 *   std::shuffle of a heap-allocated index        PRNG advanced by the exact n/2
 *     vector every tick (sendergang.cc:75-82)     draws (rejections honoured); the
 *                                                 permutation is only materialised
 *                                                 when >= 2 senders transmit in the
 *                                                 same tick (the only case in which
 *                                                 the order is observable)
 *   n TimeSwitchedSender::tick calls              acks taken from the delay ring's
 *     (sendergang.cc:190-205)                     "delivered" window; sends only for
 *                                                 senders whose pacing deadline is due
 *   tick_share += dt/k per sending sender         same additions, same order per
 *     (sendergang.cc:163-173, utility.hh:19)      sender (bit-exact), dt/k hoisted
 *   SenderGang::next_event_time O(n) + Rat O(1)   min(min_switch, min of due times
 *     (sendergang.cc:326-347, rat.cc:50-62)       of window-open senders)
 *   std::deque<Packet> x3, vector<vector<Packet>> two per-slot rings in HBM; the
 *     (link.hh, delay.hh, receiver.hh)            Receiver mailbox is a window of the
 *                                                 delay ring, never copied
 *   one Network::tick per loop iteration          the tick that reads a delivered packet
 *     (network.cc:73-84)                          runs in the pass that delivered it
 *                                                 ("fused tick", see run_sim); the loop has
 *                                                 ONE exit, at its top, so that the warp
 *                                                 re-converges behind every phase
 *
 * Template parameters: COUNT keeps per-leaf use counts, TRACE logs every lookup for the host's
 * running medians (remy_b200_score_trace), FISH swaps the Rat for the Fish sender (FinTree).
 *
 * Every floating-point expression that feeds control flow keeps the
 * reference's association and is compiled without FMA contraction
 * (--fmad=false plus explicit _rn intrinsics where it matters).
 */
#ifndef REMY_SIM_KERNEL_CUH
#define REMY_SIM_KERNEL_CUH

#include <stdint.h>
#include <float.h>
#ifdef REMY_HOST_EMU
/* tests/emu builds this header with g++ to check the restructured event loop
 * against the oracle on machines without a GPU.  Test infrastructure only: the
 * shipped library is always the nvcc build. */
#include "host_emu.h"
#define REMY_NOINLINE __attribute__((noinline))
#else
#include <cuda_runtime.h>
#define REMY_NOINLINE __noinline__
#endif
#include "../../include/remy_b200.h"
#include "exact_log.h"
#include "device_layout.h"

namespace remy {

constexpr uint32_t LCG_M = 2147483647u;
constexpr uint32_t INC_BUF = 8; /* buffered sending-time increments per simulation (shared memory) */

struct KernelArgs {
  /* tree (global copies; staged into shared memory when it fits) */
  const Bounds *bounds;        /* [n_nodes][NAX] */
  const DevNodeMeta *meta;     /* [n_nodes] */
  const DevLeaf *leaves;       /* [n_leaves] */
  const double *cuts; const uint16_t *table; uint32_t n_cuts, n_table; /* cell tables (device_layout.h) */
  uint32_t n_nodes, n_leaves, axes_mask, tree_in_smem;
  /* batch */
  const RemyLeafOverride *cands; /* [n_cands] or nullptr */
  const RemyNetConfig *cfgs;     /* [n_cfgs] */
  const uint32_t *seeds;         /* [n_cfgs] */
  const uint32_t *order;         /* [n_work] simulation ids, cost-sorted */
  uint32_t n_cfgs, n_work;
  uint32_t *work_counter;
  /* per-slot scratch */
  SenderCold *cold;   uint32_t n_max;      /* [slots][n_max] */
  double *share;                           /* [slots][n_max] Utility::_tick_share_sending */
  double *nsw;                             /* [slots][n_max] SwitchedSender::next_switch_tick */
  SendState *sendst;                       /* [slots][n_max] what a transmission needs, compact enough to stay in L2 */
  LinkEnt *link_ring; uint32_t link_cap;   /* [slots][link_cap], power of two */
  DelayEnt *delay_ring; uint32_t delay_cap;/* [slots][delay_cap], power of two */
  uint32_t *slot_counts;                   /* [slots][n_leaves] when counting uses */
  /* outputs */
  RemySimResult *out_sims;                 /* [n_sims] */
  RemySenderResult *out_senders; uint32_t sender_stride; /* optional */
  uint32_t *out_use_counts;                /* [n_cands][n_leaves], optional */
  /* trace == true (TRACE kernels only): one record per WhiskerTree::use_whisker call, in the
   * reference's call order: { leaf, bits of Memory::field(a) for every axis a in axes_mask,
   * ascending }.  Simulation `sim` writes its records from record index trace_off[sim]; with
   * trace_log == nullptr the calls are only counted (out_lookups[sim]) so that the host can lay
   * the log out exactly. */
  uint64_t *trace_log; const uint64_t *trace_off; uint64_t *out_lookups; uint32_t trace_words;
};

/* Non-binding request to bring the line holding p towards the SM: used where an address is
 * known ticks before its data is needed (next ring line, the next sender to transmit). */
__device__ __forceinline__ void prefetch_l2(const void *p)
{
#ifdef REMY_HOST_EMU
  (void)p;
#else
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
__device__ __forceinline__ void prefetch_line(const void *p)
{
#ifdef REMY_HOST_EMU
  (void)p;
#else
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

/* ---------------------------------------------------------------- PRNG ---- */
/* minstd_rand0 step (bits/random.h:1592, 368-372): 16807 x mod (2^31-1) via the
 * Mersenne fold (2^31 == 1 mod m). */
__device__ __forceinline__ uint32_t lcg_next(uint32_t x)
{
  const uint64_t p = (uint64_t)x * 16807u;
  uint32_t r = (uint32_t)(p & 0x7fffffffu) + (uint32_t)(p >> 31);
  return r >= LCG_M ? r - LCG_M : r;
}
__device__ __forceinline__ uint32_t lcg_mul(uint32_t x, uint32_t a) /* a x mod m, a,x < 2^31 */
{
  const uint64_t p = (uint64_t)x * a;
  uint64_t r = (p & 0x7fffffffu) + (p >> 31);          /* < 2^32 */
  r = (r & 0x7fffffffu) + (r >> 31);
  return (uint32_t)(r >= LCG_M ? r - LCG_M : r);
}

/* std::generate_canonical<double,53> for this engine (bits/random.tcc:3349-3383). */
__device__ __forceinline__ double canonical(uint32_t &x)
{
  x = lcg_next(x);
  const double a = (double)(x - 1u);
  x = lcg_next(x);
  const double b = __dmul_rn((double)(x - 1u), 2147483646.0);
  double ret = __ddiv_rn(__dadd_rn(a, b), 0x1.fffffffp+61);
  if (ret >= 1.0) ret = 0x1.fffffffffffffp-1; /* nextafter(1, 0) */
  return ret;
}

/* Exponential::sample (src/exponential.hh:22; bits/random.h:4898-4905). */
__device__ REMY_NOINLINE double exp_sample(uint32_t &x, double rate)
{
  const double u = canonical(x);
  return __ddiv_rn(-remy_exact_log(__dsub_rn(1.0, u)), rate);
}

/* uniform_int_distribution{0,u}, fallback branch (bits/uniform_int_dist.h:333-339). */
__device__ __forceinline__ uint32_t uniform_int(uint32_t &x, uint32_t u)
{
  const uint32_t uerange = u + 1u;
  const uint32_t scaling = 2147483645u / uerange, past = uerange * scaling;
  uint32_t ret;
  do { x = lcg_next(x); ret = x - 1u; } while (ret >= past);
  return ret / scaling;
}

/* The largest remainder 2147483645 mod ((i+1)(i+2)) over all shuffle ranges with
 * n <= 32 is < 1122, so a draw below this bound can never be rejected. */
constexpr uint32_t SHUFFLE_SAFE = 2147483645u - 1122u;

/* Advance the PRNG by exactly the draws std::shuffle makes for n elements
 * (bits/stl_algo.h:3742-3805) without building the permutation.
 *
 * Fast path: the n/2 LCG steps run on a lazily reduced residue (x' = lo31 + hi is congruent
 * to 16807 x mod m but may be >= m; it stays < 2^31 + 2^15, so the next product still fits
 * in 46 bits).  uniform_int can only reject a draw that lies in the top ~1122 values of the
 * range; the running max of the lazily reduced residues over-approximates that (a residue
 * >= m looks large but is really small), so a suspicious tick (p ~ 1e-5) takes the exact path. */
__device__ __forceinline__ uint32_t shuffle_skip(uint32_t x, uint32_t n)
{
  const uint32_t x0 = x;
  const uint32_t draws = n >> 1;
  uint32_t mx = 0;
  for (uint32_t j = 0; j < draws; j++) {
    const uint64_t p = (uint64_t)x * 16807u;
    x = (uint32_t)(p & 0x7fffffffu) + (uint32_t)(p >> 31);
    mx = max(mx, x);
  }
  if (mx - 1u < SHUFFLE_SAFE) return x; /* every residue was canonical (< m) and none rejectable */
  x = x0;
  uint32_t i = 1;
  if ((n & 1u) == 0) { (void)uniform_int(x, 1); i = 2; }
  for (; i < n; i += 2) (void)uniform_int(x, (i + 1) * (i + 2) - 1);
  return x;
}

/* Build the permutation (only when two senders transmit in one tick). */
__device__ REMY_NOINLINE void shuffle_materialise(uint32_t x, uint32_t n, uint8_t *idx)
{
  for (uint32_t i = 0; i < n; i++) idx[i] = (uint8_t)i;
  uint32_t i = 1;
  uint8_t t;
  if ((n & 1u) == 0) {
    const uint32_t d = uniform_int(x, 1);
    t = idx[1]; idx[1] = idx[d]; idx[d] = t;
    i = 2;
  }
  for (; i < n; i += 2) {
    const uint32_t b1 = i + 2;
    const uint32_t v = uniform_int(x, (i + 1) * b1 - 1);
    const uint32_t p0 = v / b1, p1 = v % b1;
    t = idx[i]; idx[i] = idx[p0]; idx[p0] = t;
    t = idx[i + 1]; idx[i + 1] = idx[p1]; idx[p1] = t;
  }
}

/* ------------------------------------------------------------- whiskers --- */
struct TreeView {
  const Bounds *bounds;  const DevNodeMeta *meta; const DevLeaf *leaves;
  const double *cuts; const uint16_t *table;
  uint32_t axes_mask;
};

__device__ __forceinline__ bool node_contains(const TreeView &T, uint32_t node, const double *m)
{
  bool ok = true;
#pragma unroll
  for (int a = 0; a < NAX; a++) {
    if (T.axes_mask & (1u << a)) {
      const Bounds b = T.bounds[node * NAX + a];
      ok = ok && (m[a] >= b.lo) && (m[a] < b.hi);
    }
  }
  return ok;
}

/* WhiskerTree::whisker (src/whiskertree.cc:62-82): first-match descent.
 * Returns the leaf index, or sets err and returns 0xFFFFFFFF. */
__device__ __forceinline__ uint32_t find_leaf(const TreeView &T, const double *m, uint32_t &err)
{
  if (!node_contains(T, 0, m)) { err |= REMY_ERR_NO_WHISKER; return 0xFFFFFFFFu; }
  uint32_t cur = 0;
  for (;;) {
    const DevNodeMeta md = T.meta[cur];
    if (md.count == 0) return md.first;
    uint32_t next = 0xFFFFFFFFu;
    if (md.table_off != NO_TABLE) {
      /* cell lookup: p_a = #{cuts <= m[a]} per axis, then one table read */
      const double *cp = T.cuts + md.cuts_off;
      uint32_t idx = 0, stride = 1;
      if (md.ncuts[7] && T.axes_mask == 0xFu) { /* bisected once on each of the four default axes (Whisker::bisect) */
        idx = (m[0] >= cp[0] ? 1u : 0u) | (m[1] >= cp[1] ? 2u : 0u) | (m[2] >= cp[2] ? 4u : 0u) | (m[3] >= cp[3] ? 8u : 0u);
      } else
#pragma unroll
      for (int a = 0; a < NAX; a++) {
        if (T.axes_mask & (1u << a)) {
          const uint32_t nc = md.ncuts[a];
          if (nc == 1) { /* a bisected axis */
            if (m[a] >= cp[0]) idx += stride;
            stride += stride; cp++;
          } else if (nc) {
            uint32_t p = 0;
            for (uint32_t k = 0; k < nc; k++) p += (m[a] >= cp[k]) ? 1u : 0u;
            idx += p * stride; stride *= nc + 1; cp += nc;
          }
        }
      }
      const uint32_t c = T.table[md.table_off + idx];
      if (c != NO_CELL_CHILD) next = md.first + c;
    } else {
      for (uint32_t c = 0; c < md.count; c++) {
        if (node_contains(T, md.first + c, m)) { next = md.first + c; break; }
      }
    }
    if (next == 0xFFFFFFFFu) { err |= REMY_ERR_NO_CHILD; return 0xFFFFFFFFu; }
    cur = next;
  }
}

/* Whisker::window (src/whisker.hh:25), with the x86-64 cvttsd2si result for
 * out-of-range values that the reference's build produces. */
__device__ __forceinline__ int32_t whisker_window(int32_t prev, double mult, int32_t incr)
{
  const double v = __dadd_rn(__dmul_rn((double)(uint32_t)prev, mult), (double)incr);
  int32_t iv = __double2int_rz(v);
  if (iv == 0x7fffffff && v >= 2147483648.0) iv = (int32_t)0x80000000;
  return min(max(0, iv), 1000000);
}

/* ------------------------------------------------------------- the sim ---- */
/* FISH selects the second sender family: Evaluator<FinTree>::score (src/evaluator.cc:105-132) with
 * Fish senders (src/fish.cc, src/fish-templates.cc) instead of Rats.  A Fish has no window: while on
 * it sends a batch of FISH_BATCH packets whenever its Poisson deadline is due, and every ack batch
 * re-draws that deadline from the sender's own PRNG.  Its state reuses the Rat's record:
 * `intersend` holds Fish::_lambda (0 = not looked up since the reset), `window` the state of
 * Fish::_prng, the pacing deadline (Fish::_next_send_time, 0 = at once) lives in hot_send[];
 * SendState mirrors {packets_sent, prng, lambda}. */
constexpr uint32_t FISH_BATCH = 5; /* Fish::_batch_size (src/fish.cc:20) */

template <bool COUNT, bool TRACE = false, bool FISH = false>
__device__ void run_sim(const KernelArgs &A, const TreeView &T, uint32_t sim,
                        double *inc_buf, double *hot_send, uint32_t HS, /* per-tick arrays, element i at [i * HS] */
                        double *share,                      /* [n_max] Utility::_tick_share_sending, HBM */
                        double *nsw,                        /* [n_max] SwitchedSender::next_switch_tick, HBM */
                        SendState *sendst,                  /* [n_max] mirror of the send fields of `cold` */
                        SenderCold *cold, LinkEnt *lring, DelayEnt *dring, uint32_t *cnt)
{
  const uint32_t cand = sim / A.n_cfgs, cfg_i = sim - cand * A.n_cfgs;
  const RemyNetConfig cfg = A.cfgs[cfg_i];
  RemySimResult res;
  res.utility = 0; res.last_tick = 0; res.event_ticks = 0; res.packets_sent = 0; res.packets_received = 0;
  res.link_queued = 0; res.delay_inflight = 0; res.prng_state = 0; res.error = 0;

  const uint32_t n = cfg.num_senders;
  if (n == 0 || n > A.n_max || n > REMY_MAX_SENDERS || !(cfg.simulation_ticks > 0) || !(cfg.link_ppt > 0)) {
    res.error = REMY_ERR_BAD_CONFIG;
    A.out_sims[sim] = res;
    if (A.out_senders) {
      RemySenderResult z; z.tick_share_sending = 0; z.total_delay = 0; z.packets_received = 0; z.packets_sent = 0;
      for (uint32_t s = 0; s < A.sender_stride; s++) A.out_senders[(size_t)sim * A.sender_stride + s] = z;
    }
    return;
  }

  /* candidate override (WhiskerTree::replace, src/whiskertree.cc:137-157) */
  uint32_t ov_leaf = REMY_NO_OVERRIDE;
  if (A.cands) ov_leaf = A.cands[cand].leaf_index;

  /* MemoryRange::track (src/memoryrange.cc:60-66) of the leaf a lookup returned */
  uint64_t nrec = 0;
  const uint64_t trace_base = (TRACE && A.trace_off) ? A.trace_off[sim] : 0;
  auto track = [&](uint32_t leaf, const double *mem) {
    if (TRACE) {
      if (A.trace_log) {
        uint64_t *r = A.trace_log + (trace_base + nrec) * A.trace_words;
        r[0] = leaf;
        uint32_t k = 1;
#pragma unroll
        for (int a = 0; a < NAX; a++) if (A.axes_mask & (1u << a)) r[k++] = (uint64_t)__double_as_longlong(mem[a]);
      }
      nrec++;
    }
  };

  uint32_t prng = A.seeds[cfg_i] % LCG_M; /* bits/random.tcc:116-126 */
  if (prng == 0) prng = 1;
  uint32_t fish_seed = 0;
  if (FISH) { prng = lcg_next(prng); fish_seed = prng; } /* unsigned int fish_prng_seed( run_prng() ), src/evaluator.cc:113 */
  uint32_t err = 0;

  const double duration = cfg.simulation_ticks;
  /* Per-simulation constants that are only needed on rare paths (switches) or once per packet are
   * re-read from the config array (a broadcast, cache-resident load) instead of living in registers:
   * the event loop keeps ~60 values live and spills would land in its hottest part. */
  const RemyNetConfig *cfgp = &A.cfgs[cfg_i];
  const double delay = cfg.delay;
  const uint32_t limit = cfg.buffer_size;
  const bool lossy = cfg.stochastic_loss_rate > 0.0;
  const double service = __ddiv_rn(1.0, cfg.link_ppt);            /* link.hh:24 */
  const uint32_t lmask = A.link_cap - 1, dmask = A.delay_cap - 1;

  /* SenderGang ctor (sendergang.cc:9-26), Rat ctor (rat.cc:8-20) */
  double min_switch = DBL_MAX;
  for (uint32_t s = 0; s < n; s++) {
    SenderCold c;
#pragma unroll
    for (int a = 0; a < NAX; a++) c.mem[a] = 0;
    c.last_tick_sent = 0; c.last_tick_received = 0; c.min_rtt = 0; c.total_delay = 0;
    c.last_send_time = 0; c.intersend = 0;
    c.reserved = 0;
    const double first_switch = exp_sample(prng, __ddiv_rn(1.0, cfg.mean_off_duration)); /* sendergang.cc:18-19 */
    nsw[s] = first_switch;
    c.packets_sent = 0; c.largest_ack = -1; c.window = FISH ? (int32_t)fish_seed : 0; c.u_received = 0; c.leaf = 0; c.pad = 0;
    cold[s] = c;
    min_switch = fmin(min_switch, first_switch);
    share[s] = 0.0;
    hot_send[s * HS] = DBL_MAX;
  }
  if (COUNT) for (uint32_t l = 0; l < A.n_leaves; l++) cnt[l] = 0;

  double t = 0.0, t_prev = 0.0;
  uint32_t on_mask = 0, open_mask = 0, zero_mask = 0;
  double next_send = DBL_MAX;          /* min of hot_send over open senders */
  bool pending = false;                /* Link::_pending_packet non-empty */
  double pend_release = 0, pend_sent = 0; uint32_t pend_seq = 0, pend_src = 0;
  uint32_t lhead = 0, ltail = 0;       /* Link::_buffer ring */
  uint32_t dcons = 0, ddeliv = 0, dtail = 0; /* delay ring: [dcons,ddeliv) = Receiver mailbox, [ddeliv,dtail) in flight */
  /* the oldest in-flight delay entry and the (single) delivered-but-unread one are kept in
   * registers: the ring is only read when an entry becomes the head */
  double dhead_release = DBL_MAX, dh_sent = 0; uint32_t dh_seq = 0, dh_src = 0;
  /* ... and so is the entry behind it (valid when dtail - ddeliv >= 2), so that the HBM read of
   * a delay-ring entry is issued a whole service time before its first use */
  double dn_release = 0, dn_sent = 0; uint32_t dn_seq = 0, dn_src = 0;
  uint64_t iters = 0;
  /* Sending time is accumulated exactly as the reference does (one rounded addition of
   * dt/k per sender per tick, sendergang.cc:163-173), but the additions are deferred: the
   * per-tick increments are identical for every sender that is on, so up to INC_BUF of them
   * are kept in shared memory and applied, in order, to each such sender at once.  The set
   * of senders cannot change while increments are buffered (a switch flushes first). */
  uint32_t ninc = 0, acc_mask = 0;
  auto flush_share = [&]() {
    uint32_t am = acc_mask;
    /* always INC_BUF additions per sender, from registers: unused slots hold +0.0, and v + 0.0 == v
     * exactly for the non-negative sums kept here */
    for (uint32_t k = ninc; k < INC_BUF; k++) inc_buf[k * HS] = 0.0;
    double inc[INC_BUF];
#pragma unroll
    for (uint32_t k = 0; k < INC_BUF; k++) inc[k] = inc_buf[k * HS];
    while (am) {
      const uint32_t s = __ffs(am) - 1; am &= am - 1;
      double v = share[s];
#pragma unroll
      for (uint32_t k = 0; k < INC_BUF; k++) v = __dadd_rn(v, inc[k]);
      share[s] = v;
    }
    ninc = 0;
  };

  /* One pass of this loop = one reference tick, or two when the tick delivers a packet: the tick
   * that follows a delivery happens at the same time t (Receiver::next_event_time, receiver.hh:34-37)
   * and consists of nothing but the shuffle's draws, the ack and the sends it releases, so it is run
   * in the same pass, right behind the delivery ("fused tick").  Every lane of a warp then reaches
   * the expensive ack code once per pass instead of once every other pass.  For that the link and
   * delay phases of a tick come BEFORE the (single) ack/send code in the pass, which is only
   * equivalent if the tick itself has nothing to send; a tick that may send AND has a departure or
   * a delivery due is therefore split over two passes (`half`): senders first, link + delay (and
   * the fused tick) in the next pass. */
  bool half = false, maybe_sends = false;
  uint32_t prng_before_shuffle = prng;
  /* trace: the lookups of run_senders happen sender by sender in the shuffled order
   * (sendergang.cc:83-86 -> TimeSwitchedSender::tick, :190-205): the ack's lookup, then
   * Rat::send's window-0 lookup, both on the Memory the ack left behind.  The order is
   * observable (the running medians are order dependent), so the permutation is built
   * whenever two senders looked something up in one tick. */
  auto track_senders = [&](uint32_t acked, uint32_t zm) {
    uint32_t lk = acked | zm;
    if (lk) {
      uint8_t perm[REMY_MAX_SENDERS];
      const bool ordered = (lk & (lk - 1)) != 0;
      if (ordered) shuffle_materialise(prng_before_shuffle, n, perm);
      for (uint32_t i = 0; lk; i++) {
        uint32_t s;
        if (ordered) { s = perm[i]; if (!((lk >> s) & 1u)) continue; }
        else s = __ffs(lk) - 1;
        lk &= ~(1u << s);
        const SenderCold c = cold[s];
        if ((acked >> s) & 1u) track(c.leaf, c.mem);
        if ((zm >> s) & 1u) track(c.leaf, c.mem);
      }
    }
  };
  /* Rat::send's lookup when the window is 0 (rat-templates.cc:13-18): the memory has not changed
   * since the sender's last lookup, so it returns the same leaf and the same window/intersend;
   * only the use count moves (and the trace gets a record). */
  auto zero_window_lookups = [&](uint32_t acked) {
    if (FISH) return;
    if (COUNT) {
      uint32_t zm = zero_mask & on_mask;
      while (zm) { const uint32_t s = __ffs(zm) - 1; zm &= zm - 1; cnt[cold[s].leaf]++; }
    }
    if (TRACE) track_senders(acked, zero_mask & on_mask);
  };

  /* The loop is left at ONE place, its top, and nowhere inside a conditional block (an error only
   * sets `err` and lets the pass run out): early exits from inside divergent regions keep the
   * compiler from re-converging the warp behind them, and every phase below then runs once per
   * group of lanes instead of once per pass. */
  for (;;) {
    bool run_link = true;
    /* ---- next event (network.cc:75-78); the Receiver never holds mail here (it is read in the
     * pass that delivers it), except when the run ends on a delivery ---- */
    double tn = fmax(t, fmin(min_switch, next_send)); /* SenderGang::next_event_time */
    if (pending) tn = fmin(tn, pend_release);           /* Link::next_event_time */
    if (ddeliv != dtail) tn = fmin(tn, dhead_release);  /* Delay::next_event_time */
    {
      const bool ended = !half && !(t < duration);
      const bool overshoot = !half && !ended && tn > duration; /* Network::_tickno keeps the overshooting time */
      if (overshoot) t = tn;
      if (err || ended || overshoot) break;
    }
    if (!half) {
      t = tn;
      iters++;
      const bool advanced = t > t_prev;
      uint32_t newly_on = 0;
      bool sends_dirty = false; /* a window or pacing deadline changed in this tick */

      /* ---- 1. switch_senders (sendergang.cc:39-45, 89-106) ---- */
      if (min_switch <= t) {
        sends_dirty = true;
        flush_share();
        const uint32_t k0 = __popc(on_mask);
        double ms = DBL_MAX;
        for (uint32_t s = 0; s < n && !err; s++) {
          double nx = nsw[s];
          if (nx <= t) {
            if (nx != t) err |= REMY_ERR_ASSERT;
            SenderCold c = cold[s];
            bool sending = (on_mask >> s) & 1u;
            double itick = t_prev;       /* SwitchedSender::internal_tick of a sender that was on */
            double shr = share[s];
            do {
              if (sending) {             /* switch_off (sendergang.cc:151-161) */
                if (t > itick) { shr = __dadd_rn(shr, __ddiv_rn(__dsub_rn(t, itick), (double)k0)); itick = t; }
                sending = false;
              } else {                   /* switch_on -> Rat::reset (rat.cc:34-48) */
                sending = true;
  #pragma unroll
                for (int a = 0; a < NAX; a++) c.mem[a] = 0;
                c.last_tick_sent = 0; c.last_tick_received = 0; c.min_rtt = 0;
                c.last_send_time = 0;
                c.largest_ack = (int32_t)(c.packets_sent - 1u);
                if (FISH) {              /* Fish::reset (fish.cc:36-48): no lookup, lambda = 0 until the first send or ack */
                  c.intersend = 0;
                } else {
                  const uint32_t leaf = find_leaf(T, c.mem, err);
                  if (err) break;
                  if (COUNT) cnt[leaf]++;
                  track(leaf, c.mem);
                  DevLeaf w = T.leaves[leaf];
                  if (leaf == ov_leaf) { const RemyLeafOverride o = A.cands[cand]; w.window_multiple = o.window_multiple; w.intersend = o.intersend; w.window_increment = o.window_increment; }
                  c.window = whisker_window(0, w.window_multiple, w.window_increment);
                  c.intersend = w.intersend;
                  c.leaf = leaf;
                }
                itick = t;
              }
              nx = __dadd_rn(nx, exp_sample(prng, __ddiv_rn(1.0, sending ? cfgp->mean_on_duration : cfgp->mean_off_duration)));
            } while (nx <= t);
            nsw[s] = nx;
            cold[s] = c;
            { SendState ss; ss.packets_sent = c.packets_sent; ss.lim = FISH ? c.window : c.largest_ack + 1 + c.window; ss.intersend = c.intersend; sendst[s] = ss; }
            share[s] = shr;
            const uint32_t bit = 1u << s;
            if (sending) {
              on_mask |= bit;
              if (itick == t) newly_on |= bit;
              /* (a sender that is on after its switch tick was reset in it: a Fish then sends at once) */
              const bool open = FISH || (int32_t)c.packets_sent < c.largest_ack + 1 + c.window;
              hot_send[s * HS] = FISH ? 0.0 : (open ? __dadd_rn(c.last_send_time, c.intersend) : DBL_MAX);
              open_mask = open ? (open_mask | bit) : (open_mask & ~bit);
              if (COUNT && !FISH) zero_mask = (c.window == 0) ? (zero_mask | bit) : (zero_mask & ~bit);
            } else {
              on_mask &= ~bit; open_mask &= ~bit;
              hot_send[s * HS] = DBL_MAX;
              if (COUNT) zero_mask &= ~bit;
            }
          }
          ms = fmin(ms, nx);
        }
        min_switch = ms;
      }
      const uint32_t k1 = __popc(on_mask);

      /* ---- 2. run_senders: the shuffle's PRNG draws (sendergang.cc:75-82) ---- */
      prng_before_shuffle = prng;
      prng = shuffle_skip(prng, n);
      zero_window_lookups(0);

      /* ---- 5. accumulate_sending_time_until (sendergang.cc:163-173) ---- */
      if (advanced) {
        uint32_t am = on_mask & ~newly_on;
        if (am) {
          if (ninc && am != acc_mask) flush_share();
          acc_mask = am;
          inc_buf[ninc * HS] = __ddiv_rn(__dsub_rn(t, t_prev), (double)k1);
          ninc++;
        }
      }
      t_prev = t;
      /* flush when the buffer of ANY lane that is here is full: every such lane then flushes in
       * the same pass (at most INC_BUF increments are pending; a lane adds at most one per pass) */
      {
        const uint32_t fullest = __reduce_max_sync(__activemask(), ninc);
        if (fullest >= INC_BUF) { if (ninc) flush_share(); }
        else if (fullest == INC_BUF - 1 && ninc) { /* next pass flushes: start fetching the sums */
          for (uint32_t s0 = 0; s0 < n; s0 += 16) prefetch_line(&share[s0]);
        }
      }

      maybe_sends = sends_dirty || next_send <= t;
      if (maybe_sends && ((pending && pend_release <= t) || (ddeliv != dtail && dhead_release <= t))) run_link = false;
    }

    bool mail = false;
    double m_sent = 0; uint32_t m_seq = 0, m_src = 0; /* the packet delivered in this pass (mail is read in the pass that delivers it) */
    if (run_link) {
      /* ---- 6. Link::tick (link-templates.cc:7-18) -> StochasticLoss (stochastic-loss.hh:21-35) -> Delay::accept ---- */
      if (pending && pend_release <= t) {
        if (pend_release != t) err |= REMY_ERR_ASSERT;
        bool lost;
        if (lossy) lost = canonical(prng) < cfgp->stochastic_loss_rate; /* bernoulli (bits/random.h:3739-3750) */
        else { prng = lcg_mul(prng, 282475249u); lost = false; } /* two draws, 16807^2 */
        if (!lost) {
          if (dtail - dcons >= A.delay_cap) err |= REMY_ERR_DELAY_OVERFLOW;
          else {
          DelayEnt e; e.release = __dadd_rn(t, delay); e.tick_sent = pend_sent; e.seq = pend_seq; e.src = pend_src; e.pad[0] = 0; e.pad[1] = 0;
          dring[dtail & dmask] = e;
          if (ddeliv == dtail) { /* it is the head of the delay line right away */
            dhead_release = e.release; dh_sent = e.tick_sent; dh_seq = e.seq; dh_src = e.src;
          } else if (dtail - ddeliv == 1) { /* ... or the entry right behind the head */
            dn_release = e.release; dn_sent = e.tick_sent; dn_seq = e.seq; dn_src = e.src;
          }
          dtail++;
          }
        }
        pending = false;
      }
      if (!pending && lhead != ltail) {
        const LinkEnt e = lring[lhead & lmask]; lhead++;
        if ((lhead & 7u) == 0 && (ltail - lhead) > 8) prefetch_line(&lring[(lhead + 8) & lmask]); /* next 128-byte line of the queue */
        pending = true; pend_release = __dadd_rn(t, service); pend_sent = e.tick_sent; pend_seq = e.seq; pend_src = e.src;
      }
      /* ---- 7. Delay::tick (delay.hh:53-63) -> Receiver::accept ---- */
      /* (not written as a loop over the register-held head: a loop's condition would wait for
       * the ring read issued below, which exists precisely so that nothing waits for it) */
      if (ddeliv != dtail && dhead_release <= t) {
        if (dhead_release != t) err |= REMY_ERR_ASSERT;
        /* Receiver::accept: the delivered packet's fields move to the mail registers */
        m_sent = dh_sent; m_seq = dh_seq; m_src = dh_src;
        ddeliv++;
        if (ddeliv != dtail) {
          dhead_release = dn_release; dh_sent = dn_sent; dh_seq = dn_seq; dh_src = dn_src;
          if (dhead_release <= t) { /* rare: more entries released at the same instant; walk the ring itself */
            while (ddeliv != dtail) {
              const double rel = dring[ddeliv & dmask].release;
              if (!(rel <= t)) break;
              if (rel != t) err |= REMY_ERR_ASSERT;
              ddeliv++;
            }
            if (ddeliv != dtail) { const DelayEnt h = dring[ddeliv & dmask]; dhead_release = h.release; dh_sent = h.tick_sent; dh_seq = h.seq; dh_src = h.src; }
          }
        }
        if (ddeliv != dtail) {
          prefetch_l2(&cold[dh_src]); /* the next ack's sender record: a service time from now */
          if (dtail - ddeliv >= 2) { /* read the new second entry now; it is not used before the next delivery */
            const DelayEnt h = dring[(ddeliv + 1) & dmask];
            dn_release = h.release; dn_sent = h.tick_sent; dn_seq = h.seq; dn_src = h.src;
            if (((ddeliv + 1) & 3u) == 3u && (dtail - ddeliv) > 6) prefetch_line(&dring[(ddeliv + 5) & dmask]);
          }
        }
      }
      /* the tick that reads the mail: same t, no switch can be due, no time passes */
      mail = (dcons != ddeliv) && (t < duration);
      if (mail) { iters++; prng_before_shuffle = prng; prng = shuffle_skip(prng, n); }
    }

    /* ---- 3. receive_feedback for every mailbox (sendergang.cc:175-188) ---- */
    /* the sender acked in this tick usually transmits in this tick: its send state is handed
     * over in registers instead of being re-read from HBM */
    uint32_t ak_s = 0xFFFFFFFFu, ak_seq = 0; int32_t ak_lim = 0; double ak_isend = 0;
    uint32_t acked = 0; /* senders whose mailbox was read in this tick */
    if (mail) {
      uint32_t done = 0;
      const bool single = (ddeliv - dcons) == 1; /* the usual case: its fields are in m_* */
      for (uint32_t i = dcons; i != ddeliv && !err; i++) {
        const uint32_t s = single ? m_src : dring[i & dmask].src;
        if ((done >> s) & 1u) continue;
        done |= 1u << s;
        SenderCold c = cold[s];
        const int32_t old_ack = c.largest_ack;
        int32_t last_seq = 0;
        for (uint32_t j = i; j != ddeliv; j++) {
          DelayEnt e;
          if (single) { e.tick_sent = m_sent; e.seq = m_seq; e.src = m_src; }
          else e = dring[j & dmask];
          if (e.src != s) continue;
          /* Utility::packets_received (utility.hh:20-27) */
          const double rtt = __dsub_rn(t, e.tick_sent);
          c.u_received++;
          c.total_delay = __dadd_rn(c.total_delay, rtt);
          /* Memory::packets_received (memory.cc:31-80) */
          int32_t outstanding = 1;
          if ((int64_t)e.seq > (int64_t)old_ack) outstanding = (int32_t)((int64_t)e.seq - (int64_t)old_ack);
          if (c.min_rtt <= 0) c.min_rtt = rtt;
          if (c.last_tick_sent == 0 || c.last_tick_received == 0) {
            c.last_tick_sent = e.tick_sent; c.last_tick_received = t; c.min_rtt = rtt;
          } else {
            const double ds = __dsub_rn(e.tick_sent, c.last_tick_sent);
            const double dr = __dsub_rn(t, c.last_tick_received);
            c.mem[0] = __dadd_rn(__dmul_rn(0.875, c.mem[0]), __dmul_rn(0.125, ds));
            c.mem[1] = __dadd_rn(__dmul_rn(0.875, c.mem[1]), __dmul_rn(0.125, dr));
            c.mem[3] = __dadd_rn(__dmul_rn(0.99609375, c.mem[3]), __dmul_rn(0.00390625, dr));
            c.last_tick_sent = e.tick_sent; c.last_tick_received = t;
            c.min_rtt = (rtt < c.min_rtt) ? rtt : c.min_rtt;
            c.mem[2] = __ddiv_rn(rtt, c.min_rtt);
            c.mem[4] = __dsub_rn(rtt, c.min_rtt);
            if (!(c.mem[2] >= 1.0) || !(c.mem[4] >= 0)) err |= REMY_ERR_ASSERT;
            c.mem[5] = __dmul_rn(c.mem[1], (double)outstanding);
          }
          last_seq = (int32_t)e.seq;
        }
        /* Rat::packets_received (rat.cc:22-32) */
        c.largest_ack = max(last_seq, old_ack);
        const uint32_t leaf = find_leaf(T, c.mem, err);
        if (err) break;
        if (COUNT) cnt[leaf]++;
        DevLeaf w = T.leaves[leaf];
        if (leaf == ov_leaf) { const RemyLeafOverride o = A.cands[cand]; w.window_multiple = o.window_multiple; w.intersend = o.intersend; w.window_increment = o.window_increment; }
        double fish_due = 0;
        if (FISH) { /* Fish::packets_received (fish.cc:26-34): _update_lambda, then _update_send_time(_last_send_time) */
          c.intersend = w.intersend;
          uint32_t fp = (uint32_t)c.window;
          const double x = exp_sample(fp, c.intersend), mx = __ddiv_rn(2.0, c.intersend);
          c.window = (int32_t)fp;
          fish_due = __dadd_rn(c.last_send_time, __dmul_rn((double)FISH_BATCH, (mx < x) ? mx : x));
        } else {
          c.window = whisker_window(c.window, w.window_multiple, w.window_increment);
          c.intersend = w.intersend;
        }
        c.leaf = leaf;
        cold[s] = c;
        ak_s = s; ak_seq = c.packets_sent; ak_lim = FISH ? c.window : c.largest_ack + 1 + c.window; ak_isend = c.intersend;
        { SendState ss; ss.packets_sent = ak_seq; ss.lim = ak_lim; ss.intersend = ak_isend; sendst[s] = ss; }
        const uint32_t bit = 1u << s;
        if (on_mask & bit) {
          const bool open = FISH || (int32_t)c.packets_sent < c.largest_ack + 1 + c.window;
          hot_send[s * HS] = FISH ? fish_due : (open ? __dadd_rn(c.last_send_time, c.intersend) : DBL_MAX);
          open_mask = open ? (open_mask | bit) : (open_mask & ~bit);
          if (COUNT && !FISH) zero_mask = (c.window == 0) ? (zero_mask | bit) : (zero_mask & ~bit);
        }
      }
      dcons = ddeliv;
      acked = done;
    }

    if (mail) zero_window_lookups(acked);

    /* ---- 4. Rat::send for every sender whose deadline is due (rat-templates.cc:20-34): the sender phase
     * of the fused tick if mail was read, else of the tick this pass began with (the scan is skipped when
     * no deadline is due and nothing changed: next_send stays valid) ---- */
    if (mail || (!half && maybe_sends)) {
      uint32_t ready = 0, ns_sender = 0xFFFFFFFFu;
      uint32_t first_send = 0; /* Fish that looked their lambda up in this tick's send (trace) */
      double ns = DBL_MAX;
      uint32_t om = open_mask;
      while (om) {
        const uint32_t s = __ffs(om) - 1; om &= om - 1;
        const double st = hot_send[s * HS];
        if (st <= t) ready |= 1u << s; else if (st < ns) { ns = st; ns_sender = s; }
      }
      if (ready) {
        uint8_t perm[REMY_MAX_SENDERS];
        const bool ordered = (ready & (ready - 1)) != 0;
        if (ordered) shuffle_materialise(prng_before_shuffle, n, perm);
        uint32_t left = ready;
        for (uint32_t i = 0; left; i++) {
          uint32_t s;
          if (ordered) { s = perm[i]; if (!((left >> s) & 1u)) continue; }
          else s = __ffs(left) - 1;
          left &= ~(1u << s);
          SenderCold *c = &cold[s];
          uint32_t seq; int32_t lim; double isend;
          if (s == ak_s) { seq = ak_seq; lim = ak_lim; isend = ak_isend; ak_seq = seq + 1; }
          else { const SendState ss = sendst[s]; seq = ss.packets_sent; lim = ss.lim; isend = ss.intersend; }
          if (FISH) { /* Fish::send (fish-templates.cc:8-27) */
            if (isend == 0) { /* no ack since the reset, so the Memory is still all zero: initial lambda */
              double zero[NAX];
#pragma unroll
              for (int a = 0; a < NAX; a++) zero[a] = 0;
              const uint32_t leaf = find_leaf(T, zero, err);
              if (err) break;
              if (COUNT) cnt[leaf]++;
              if (TRACE) { c->leaf = leaf; first_send |= 1u << s; } /* (its stored Memory is the all-zero one of the reset) */
              isend = (leaf == ov_leaf) ? A.cands[cand].intersend : T.leaves[leaf].intersend;
            }
            for (uint32_t b = 0; b < FISH_BATCH; b++, seq++) { /* Link::accept (link.hh:26-34) for each packet of the batch */
              if (!pending) {
                pending = true; pend_release = __dadd_rn(t, service); pend_sent = t; pend_seq = seq; pend_src = s;
              } else if ((ltail - lhead) < limit) {
                if (ltail - lhead >= A.link_cap) { err |= REMY_ERR_LINK_OVERFLOW; break; }
                LinkEnt e; e.tick_sent = t; e.seq = seq; e.src = s;
                lring[ltail & lmask] = e; ltail++;
              }
            }
            if (s == ak_s) ak_seq = seq;
            uint32_t fp = (uint32_t)lim;
            const double x = exp_sample(fp, isend), mx = __ddiv_rn(2.0, isend);
            const double st = __dadd_rn(t, __dmul_rn((double)FISH_BATCH, (mx < x) ? mx : x)); /* _update_send_time(tickno) */
            c->packets_sent = seq; c->last_send_time = t; c->intersend = isend; c->window = (int32_t)fp;
            { SendState ss; ss.packets_sent = seq; ss.lim = (int32_t)fp; ss.intersend = isend; sendst[s] = ss; }
            hot_send[s * HS] = st;
            if (st < ns) { ns = st; ns_sender = s; }
            continue;
          }
          c->packets_sent = seq + 1;
          sendst[s].packets_sent = seq + 1;
          c->last_send_time = t;
          /* Link::accept (link.hh:26-34) */
          if (!pending) {
            pending = true; pend_release = __dadd_rn(t, service); pend_sent = t; pend_seq = seq; pend_src = s;
          } else if ((ltail - lhead) < limit) { /* (a limit of 0 accepts nothing) */
            if (ltail - lhead >= A.link_cap) { err |= REMY_ERR_LINK_OVERFLOW; break; }
            LinkEnt e; e.tick_sent = t; e.seq = seq; e.src = s;
            lring[ltail & lmask] = e; ltail++;
          }
          const bool open = (int32_t)(seq + 1) < lim;
          const double st = open ? __dadd_rn(t, isend) : DBL_MAX;
          hot_send[s * HS] = st;
          if (!open) open_mask &= ~(1u << s);
          if (st < ns) { ns = st; ns_sender = s; }
        }
      }
      next_send = ns;
      /* trace of a Fish tick: per sender, in the shuffled order, the ack's lookup or (never both:
       * an ack sets lambda) the first send's lookup on the all-zero Memory of the reset */
      if (TRACE && FISH) track_senders(acked, first_send);
      if (ns_sender != 0xFFFFFFFFu) prefetch_line(&sendst[ns_sender]); /* what its transmission will read */
    }

    half = !run_link;
  }

  /* ---- results: Utility::utility (utility.hh:46-60), SenderGang::utility (sendergang.cc:280-293) ---- */
  flush_share();
  double total = 0.0;
  uint64_t sent = 0, recv = 0;
  for (uint32_t s = 0; s < n; s++) {
    const SenderCold c = cold[s];
    const double shr = share[s];
    double u;
    if (shr == 0) u = 0.0;
    else if (c.u_received == 0) u = -2147483647.0;
    else {
      const double tput = __ddiv_rn((double)c.u_received, shr);
      const double del = __ddiv_rn(c.total_delay, (double)c.u_received);
      u = __dsub_rn(log2(tput), log2(__ddiv_rn(del, 100.0)));
    }
    total = __dadd_rn(total, u);
    sent += c.packets_sent; recv += c.u_received;
    if (A.out_senders && s < A.sender_stride) {
      RemySenderResult r; r.tick_share_sending = shr; r.total_delay = c.total_delay;
      r.packets_received = c.u_received; r.packets_sent = c.packets_sent;
      A.out_senders[(size_t)sim * A.sender_stride + s] = r;
    }
  }
  if (A.out_senders) {
    RemySenderResult z; z.tick_share_sending = 0; z.total_delay = 0; z.packets_received = 0; z.packets_sent = 0;
    for (uint32_t s = n; s < A.sender_stride; s++) A.out_senders[(size_t)sim * A.sender_stride + s] = z;
  }
  res.utility = __ddiv_rn(total, (double)n);
  res.last_tick = t;
  res.event_ticks = iters;
  res.packets_sent = sent; res.packets_received = recv;
  res.link_queued = (ltail - lhead) + (pending ? 1u : 0u);
  res.delay_inflight = dtail - dcons;
  res.prng_state = prng;
  res.error = err;
  A.out_sims[sim] = res;
  if (TRACE && A.out_lookups) A.out_lookups[sim] = nrec;
  if (COUNT && A.out_use_counts && !(err & (REMY_ERR_LINK_OVERFLOW | REMY_ERR_DELAY_OVERFLOW))) {
    /* (an overflowed simulation is re-run by the host with larger rings) */
    for (uint32_t l = 0; l < A.n_leaves; l++) {
      const uint32_t v = cnt[l];
      if (v) atomicAdd(&A.out_use_counts[(size_t)cand * A.n_leaves + l], v);
    }
  }
}

constexpr int SIM_BLOCK = 128;
#ifndef REMY_MIN_BLOCKS_PER_SM
/* Three resident blocks per SM = 168 registers per thread and no spills.  Four (128 registers, 44 B of
 * spills and re-reads of per-simulation constants) gave 33 % more warps but measured 10 % slower on the
 * final kernel (8.47 vs 9.34 G event ticks/s, 102 400 simulations x 3e4 ticks); two (255 registers) 18 % slower. */
#define REMY_MIN_BLOCKS_PER_SM 3
#endif

#ifndef REMY_HOST_EMU

template <bool COUNT, bool TRACE = false, bool FISH = false>
__global__ void __launch_bounds__(SIM_BLOCK, REMY_MIN_BLOCKS_PER_SM) remy_sim_kernel(const KernelArgs A)
{
  extern __shared__ __align__(16) unsigned char smem[];
  /* layout: [inc_buf: INC_BUF*B doubles][hot_send: n_max*B doubles][tree, if staged] */
  double *inc_buf = reinterpret_cast<double *>(smem) + threadIdx.x;
  double *hot_send = inc_buf + (size_t)INC_BUF * blockDim.x;
  const uint32_t HS = blockDim.x;
  TreeView T;
  T.axes_mask = A.axes_mask;
  if (A.tree_in_smem) {
    unsigned char *p = smem + (size_t)(INC_BUF + A.n_max) * blockDim.x * sizeof(double);
    Bounds *sb = reinterpret_cast<Bounds *>(p);
    DevLeaf *sl = reinterpret_cast<DevLeaf *>(sb + (size_t)A.n_nodes * NAX);
    DevNodeMeta *sm = reinterpret_cast<DevNodeMeta *>(sl + A.n_leaves);
    double *sc = reinterpret_cast<double *>(sm + A.n_nodes);
    uint16_t *st = reinterpret_cast<uint16_t *>(sc + A.n_cuts);
    for (uint32_t i = threadIdx.x; i < A.n_cuts; i += blockDim.x) sc[i] = A.cuts[i];
    for (uint32_t i = threadIdx.x; i < A.n_table; i += blockDim.x) st[i] = A.table[i];
    for (uint32_t i = threadIdx.x; i < A.n_nodes * NAX; i += blockDim.x) sb[i] = A.bounds[i];
    for (uint32_t i = threadIdx.x; i < A.n_leaves; i += blockDim.x) sl[i] = A.leaves[i];
    for (uint32_t i = threadIdx.x; i < A.n_nodes; i += blockDim.x) sm[i] = A.meta[i];
    __syncthreads();
    T.bounds = sb; T.leaves = sl; T.meta = sm; T.cuts = sc; T.table = st;
  } else {
    T.bounds = A.bounds; T.leaves = A.leaves; T.meta = A.meta; T.cuts = A.cuts; T.table = A.table;
  }
  const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  SenderCold *cold = A.cold + slot * A.n_max;
  double *share = A.share + slot * A.n_max;
  double *nsw = A.nsw + slot * A.n_max;
  SendState *sendst = A.sendst + slot * A.n_max;
  LinkEnt *lring = A.link_ring + slot * A.link_cap;
  DelayEnt *dring = A.delay_ring + slot * A.delay_cap;
  uint32_t *cnt = COUNT ? A.slot_counts + slot * A.n_leaves : nullptr;
  for (;;) {
    const uint32_t w = atomicAdd(A.work_counter, 1u);
    if (w >= A.n_work) break;
    run_sim<COUNT, TRACE, FISH>(A, T, A.order[w], inc_buf, hot_send, HS, share, nsw, sendst, cold, lring, dring, cnt);
  }
}

#endif /* !REMY_HOST_EMU */

} // namespace remy
#endif
