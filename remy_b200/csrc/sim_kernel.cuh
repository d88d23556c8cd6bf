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
 *
 * Scheduling notes (read off ptxas' control codes with tools/sass_ctrl.py; the kernel is bound by memory
 * latency at 12 resident warps per SM, so WHERE a load is waited for decides the speed):
 *  - ptxas puts every global load of this kernel on ONE scoreboard, a counter: waiting for any load waits for
 *    all loads in flight, and it drains the counter at loop headers.  Register double-buffering of ring
 *    entries therefore only works if the read is issued where a wait is cheap: the ring reads a pass leaves
 *    behind are issued at the top of the NEXT pass, in front of the shuffle's chain of LCG steps (REMY_TOP_LOADS).
 *  - a value loaded once per simulation and first used inside the event loop costs a scoreboard wait at that
 *    use on every pass (settle()); a rare path that loads into loop-carried registers puts its wait on the
 *    common path behind the join (the rare multi-delivery path consumes its loads itself).
 *  - nothing of the event loop may live in local memory: the PRNG state is passed to the out-of-line log by
 *    value (exp_draw), the per-tick arrays are addressed by shared-window address (TickMem), the tree is read
 *    with ld.shared (kernel template parameter TSM).
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
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#define REMY_NOINLINE __noinline__
#endif
#include "../../include/remy_b200.h"
#include "exact_log.h"
#include "device_layout.h"

namespace remy {

constexpr uint32_t LCG_M = 2147483647u;
constexpr uint32_t INC_BUF = 16;  /* buffered sending-time increments per simulation (shared memory) */
#ifndef REMY_CHAIN
#define REMY_CHAIN 1 /* chained ticks (see run_sim); 0 = one reference tick per pass of the event loop */
#endif
#ifndef REMY_DRING_PF
#define REMY_DRING_PF 0
#endif
#ifndef REMY_SD_REPS
#define REMY_SD_REPS 2 /* times the transmission + link phases are gone through per pass */
#endif
#ifndef REMY_SHUF2X
#define REMY_SHUF2X 1 /* shuffle_skip on the doubled residue (2 dependent instructions per draw instead of 4) */
#endif
#ifndef REMY_TOP_LOADS
#define REMY_TOP_LOADS 1 /* ring reads whose data is needed a pass later are ISSUED at the top of that pass (see run_sim) */
#endif
#ifndef REMY_SD_UNROLL
#define REMY_SD_UNROLL 0 /* 1 = both goes as straight-line code (no loop header, at which ptxas drains every scoreboard): measured 1 % slower */
#endif

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
  const uint32_t *sim_map;       /* nullptr, or [n_sims]: the (candidate, config) index each local simulation id stands for
                                  * (a batch partitioned over several devices: ids, work list and outputs are per device) */
  uint32_t n_cfgs, n_work;
  uint32_t lanes_per_warp;       /* 32, or fewer: only that many lanes of every warp take work (capi.cu, spread lists) */
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
/* Maximum of v over the lanes of the warp that are executing this call together.  Lanes of a warp
 * run different simulations and leave the event loop at different times, so there is no point at
 * which a full-warp mask is known; cooperative groups' coalesced_threads() is the defined way to
 * name "the lanes that are here" and to reduce over exactly them.  The result only steers WHEN the
 * buffered sending-time increments are applied, never what is computed. */
__device__ __forceinline__ uint32_t warp_max_here(uint32_t v)
{
#ifdef REMY_HOST_EMU
  return v;
#else
  const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
  return cooperative_groups::reduce(g, v, cooperative_groups::greater<uint32_t>());
#endif
}

/* See run_sim: make the value of a freshly loaded register "used" at this point of the program (x + (-0.0)
 * is x for every x, including both zeros, infinities and NaN payloads). */
__device__ __forceinline__ double settle(double x)
{
#ifdef REMY_HOST_EMU
  return x;
#else
  double y;
  asm volatile("add.rn.f64 %0, %1, 0d8000000000000000;" : "=d"(y) : "d"(x));
  return y;
#endif
}
__device__ __forceinline__ bool settle_positive(double x) /* x > 0.0, evaluated HERE */
{
#ifdef REMY_HOST_EMU
  return x > 0.0;
#else
  uint32_t r;
  asm volatile("{ .reg .pred p; setp.gt.f64 p, %1, 0d0000000000000000; selp.u32 %0, 1, 0, p; }" : "=r"(r) : "d"(x));
  return r != 0;
#endif
}

/* The two per-tick arrays of a simulation (buffered sending-time increments, pacing deadlines of the senders)
 * live in shared memory, element i of this thread at [i * blockDim.x + threadIdx.x].  They are addressed by
 * their 32-bit shared-window address through ld/st.shared: through C++ pointers the compiler re-derived the
 * window base at every access (S2UR SR_CgaCtaId + ULEA, a scoreboarded special-register read: nine of them
 * per pass of the event loop) and kept two 64-bit generic pointers live. */
struct TickMem {
#ifdef REMY_HOST_EMU
  double *p; uint32_t stride;
  double ld(uint32_t i) const { return p[i * stride]; }
  void st(uint32_t i, double v) const { p[i * stride] = v; }
#else
  uint32_t base; /* shared-window byte address of this thread's element 0 */
  uint32_t stride; /* bytes between consecutive elements */
  __device__ __forceinline__ double ld(uint32_t i) const
  {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + i * stride));
    return v;
  }
  __device__ __forceinline__ void st(uint32_t i, double v) const
  {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + i * stride), "d"(v));
  }
#endif
};

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

/* Exponential::sample (src/exponential.hh:22; bits/random.h:4898-4905).  The out-of-line part takes the
 * engine state BY VALUE: a PRNG whose address is passed to a function that is not inlined lives in local
 * memory for the whole kernel (the event loop then began every pass with a stack load of it and ended every
 * shuffle with a store); the caller advances the state itself by the two draws of generate_canonical. */
__device__ REMY_NOINLINE double exp_draw(uint32_t x, double rate)
{
  const double u = canonical(x);
  return __ddiv_rn(-remy_exact_log(__dsub_rn(1.0, u)), rate);
}
__device__ __forceinline__ double exp_sample(uint32_t &x, double rate)
{
  const double v = exp_draw(x, rate);
  x = lcg_mul(x, 282475249u); /* two steps: 16807^2 */
  return v;
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
#if REMY_SHUF2X
  /* The residue is carried doubled, y = 2 x: the 64-bit product 2 x 16807 = q 2^32 + 2 r splits the 46-bit
   * product q 2^31 + r at a register boundary, so the Mersenne fold x' = q + r is ONE shift-add (y' = lo + 2 hi)
   * instead of mask, funnel shift and add - the chain per draw is a multiply and an add.  q + r < 2^31 except
   * in ~8e-6 of the draws (it is then congruent to a residue below 16807); the doubled value wraps to less than
   * 2 x 16807, which the running minimum catches - those ticks, like the ones with a draw that uniform_int could
   * reject (running maximum), take the exact path.  Otherwise every y / 2 is the canonical residue. */
  {
    uint32_t y = x << 1, mx = 0, mn = 0xFFFFFFFFu;
    for (uint32_t j = 0; j < draws; j++) {
#ifdef REMY_HOST_EMU
      const uint64_t p = (uint64_t)y * 16807u;
      y = (uint32_t)p + 2u * (uint32_t)(p >> 32);
#else
      /* (spelled out: the compiler turns 2 (p >> 32) into (p >> 31) & ~1 - funnel shift, mask, add again) */
      asm("{ .reg .b32 lo, hi; .reg .b64 p;\n\t"
          "mul.wide.u32 p, %1, 16807; mov.b64 {lo, hi}, p;\n\t"
          "shl.b32 hi, hi, 1; add.u32 %0, lo, hi; }" : "=r"(y) : "r"(y));
#endif
      mx = max(mx, y); mn = min(mn, y);
    }
    if (mn >= 2u * 16807u && mx <= 2u * SHUFFLE_SAFE) return y >> 1;
  }
#else
  uint32_t mx = 0;
  for (uint32_t j = 0; j < draws; j++) {
    const uint64_t p = (uint64_t)x * 16807u;
    x = (uint32_t)(p & 0x7fffffffu) + (uint32_t)(p >> 31);
    mx = max(mx, x);
  }
  if (mx - 1u < SHUFFLE_SAFE) return x; /* every residue was canonical (< m) and none rejectable */
#endif
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

/* dt / k for a small integer k (1..32 senders that are on) and dt >= 0, correctly rounded, without
 * the division routine; inv_k = RN(1/k).  With q0 = RN(dt inv_k), the residual r = dt - k q0 is exact
 * in one FMA (it is a multiple of ulp(q0) of magnitude < 2^7 ulp), and q0 + r inv_k differs from dt/k
 * by a relative 2^-105; dt/k with k <= 32 is either a double or at least ulp/64 away from the nearest
 * rounding boundary (k = 2^p o with o odd: a quotient that terminates in binary has o | mantissa and
 * fits 53 bits, so no quotient is an exact midpoint), hence the final rounding gives RN(dt/k): the
 * value of the reference's `(tickno - _internal_tick) / num_sending` (src/sendergang.cc:169-171).
 * Checked against the hardware division for every k on 10^8 random and edge-case dt by
 * tests/test_kernel_emu.py::test_small_divisor_division. */
__device__ __forceinline__ double div_small_int(double dt, double k, double inv_k)
{
  const double q0 = __dmul_rn(dt, inv_k);
  const double r = __fma_rn(-q0, k, dt);
  return __fma_rn(r, inv_k, q0);
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
constexpr uint32_t PEND_STALE = 0xFFFFFFFFu; /* pend_src: the packet in service has not been read from the link ring yet */
constexpr uint32_t FISH_BATCH = 5; /* Fish::_batch_size (src/fish.cc:20) */
/* BYTESW selects ByteSwitchedSender<Rat> (src/sendergang.cc:108-138, 256-278) instead of TimeSwitchedSender<Rat>: a sender
 * is switched ON after an exponential idle time and switches itself OFF when it has sent its flow, whose length in
 * packets is lrint(ceil(Exp(1 / mean_on_duration))) - or INT_MAX when mean_on_duration <= 0, the "-2" of most shipped
 * .cfg files under config/.  The flow's end (Rat::_packets_sent == packets_sent_cap_) is found in the transmission phase,
 * where the idle time to the next flow is drawn from the run's PRNG, in the shuffled sender order.  The cap lives in
 * SenderCold::pad. */

template <bool COUNT, bool TRACE = false, bool FISH = false, bool BYTESW = false>
__device__ void run_sim(const KernelArgs &A, const TreeView &T, uint32_t sim,
                        const TickMem tm, /* per-tick arrays: elements [0, INC_BUF) = increments, INC_BUF + s = deadline of sender s */
                        double *share,                      /* [n_max] Utility::_tick_share_sending, HBM */
                        double *nsw,                        /* [n_max] SwitchedSender::next_switch_tick, HBM */
                        SendState *sendst,                  /* [n_max] mirror of the send fields of `cold` */
                        SenderCold *cold, LinkEnt *lring, DelayEnt *dring, uint32_t *cnt)
{
  static_assert(!(BYTESW && (FISH || TRACE)), "ByteSwitchedSender is built for Rat senders, without the lookup log");
  const uint32_t pair = A.sim_map ? A.sim_map[sim] : sim;
  const uint32_t cand = pair / A.n_cfgs, cfg_i = pair - cand * A.n_cfgs;
  const RemyNetConfig cfg = A.cfgs[cfg_i];
  RemySimResult res;
  res.utility = 0; res.last_tick = 0; res.event_ticks = 0; res.packets_sent = 0; res.packets_received = 0;
  res.link_queued = 0; res.delay_inflight = 0; res.prng_state = 0; res.error = 0;

  const uint32_t n = cfg.num_senders;
  /* (on/off durations: a negative or NaN mean, or both means 0, would keep the switch loop below - and the
   * reference's, src/sendergang.cc:95-105 - from ever ending; on a GPU that hangs the device, so it is refused) */
  /* (flows: mean_on is a length in packets, <= 0 = unbounded; an idle time of 0 lets a sender whose flow fits its window
   * start flow after flow in the same instant - the reference then never leaves that tick) */
  const bool bad_durations = BYTESW ? !(cfg.mean_off_duration > 0)
                                    : (!(cfg.mean_on_duration >= 0) || !(cfg.mean_off_duration >= 0) || !(cfg.mean_on_duration + cfg.mean_off_duration > 0));
  if (n == 0 || n > A.n_max || n > REMY_MAX_SENDERS || !(cfg.simulation_ticks > 0) || !(cfg.link_ppt > 0) || bad_durations) {
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
  /* `delay` and the loss flag are first USED deep inside the event loop.  ptxas tracks a pending load per
   * register, and all global loads of this kernel share one scoreboard: a first use inside the loop puts a
   * wait for that scoreboard there on every pass, which in steady state waits for whatever load was issued
   * last - the link phase's loss test stalled for the send-state prefetch issued just before it.  settle()
   * is an instruction of its own in front of the loop, so the wait happens here, once. */
  const double delay = settle(cfg.delay);
  const uint32_t limit = cfg.buffer_size;
  const bool lossy = settle_positive(cfg.stochastic_loss_rate);
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
    tm.st(INC_BUF + s, DBL_MAX);
  }
  if (COUNT) for (uint32_t l = 0; l < A.n_leaves; l++) cnt[l] = 0;

  double t = 0.0;
  uint32_t on_mask = 0, open_mask = 0, zero_mask = 0;
  double next_send = DBL_MAX;          /* min of hot_send over open senders */
  bool pending = false;                /* Link::_pending_packet non-empty */
  double pend_release = 0, pend_sent = 0; uint32_t pend_seq = 0, pend_src = 0;
  uint32_t lhead = 0, ltail = 0;       /* Link::_buffer ring */
  uint32_t dcons = 0, ddeliv = 0, dtail = 0; /* delay ring: [dcons,ddeliv) = Receiver mailbox, [ddeliv,dtail) in flight */
  /* the oldest in-flight delay entry is kept in registers: the ring is only read when an entry
   * becomes the head */
  double dhead_release = DBL_MAX, dh_sent = 0; uint32_t dh_seq = 0, dh_src = 0;
  /* ... and so is the entry behind it (valid when dtail - ddeliv >= 2), so that the HBM read of
   * a delay-ring entry is issued a whole service time before its first use */
  double dn_release = 0, dn_sent = 0; uint32_t dn_seq = 0, dn_src = 0;
  /* ... and the packet the latest Delay::tick delivered (the first entry of the mailbox), which the
   * senders read in the next tick */
  double m_sent = 0; uint32_t m_seq = 0, m_src = 0;
  uint64_t iters = 0;
  /* Sending time is accumulated exactly as the reference does (one rounded addition of
   * dt/k per sender per tick, sendergang.cc:163-173), but the additions are deferred: the
   * per-tick increments are identical for every sender that is on, so up to INC_BUF of them
   * are kept in shared memory and applied, in order, to each such sender at once.  The set
   * of senders that are on cannot change while increments are buffered (a switch flushes first,
   * and the increment of a switch tick itself is applied at once). */
  uint32_t ninc = 0;
  double inv_on = 0; /* correctly rounded 1 / (number of senders that are on) */
  double k_on = 0;   /* that number itself, as a double: it only changes at switches, and POPC + I2F.F64 in front of
                      * every tick's increment were two scoreboarded operations on the path of all lanes */
  auto flush_share = [&]() {
    /* Four senders at a time: their sums are requested together (one memory latency per group instead of
     * one per sender), and the four addition chains - each strictly in order, as the reference adds -
     * interleave in the FP64 pipe.  The increments are read from shared memory once per group; padding up
     * to a multiple of two holds +0.0 (v + 0.0 == v exactly for the non-negative sums kept here). */
    const uint32_t kmax = (ninc + 1u) & ~1u;
    if (ninc & 1u) tm.st(ninc, 0.0);
    uint32_t am = on_mask;
    while (am) {
      const uint32_t s0 = __ffs(am) - 1; am &= am - 1;
      const bool h1 = am != 0; const uint32_t s1 = h1 ? __ffs(am) - 1 : s0; am &= am - 1;
      const bool h2 = am != 0; const uint32_t s2 = h2 ? __ffs(am) - 1 : s0; am &= am - 1;
      const bool h3 = am != 0; const uint32_t s3 = h3 ? __ffs(am) - 1 : s0; am &= am - 1;
      double v0 = share[s0], v1 = share[s1], v2 = share[s2], v3 = share[s3];
      for (uint32_t k = 0; k < kmax; k += 2) {
        const double i0 = tm.ld(k), i1 = tm.ld((k + 1));
        v0 = __dadd_rn(__dadd_rn(v0, i0), i1); v1 = __dadd_rn(__dadd_rn(v1, i0), i1);
        v2 = __dadd_rn(__dadd_rn(v2, i0), i1); v3 = __dadd_rn(__dadd_rn(v3, i0), i1);
      }
      share[s0] = v0;
      if (h1) share[s1] = v1;
      if (h2) share[s2] = v2;
      if (h3) share[s3] = v3;
    }
    ninc = 0;
  };
  auto share_of = [&](double dt, double k) { return div_small_int(dt, k, inv_on); };

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
  /* A chained tick: the sender phase (up to the transmissions) of a reference tick at time t2 > t in
   * which no switch is due and no mail is waiting - run_senders' shuffle draws (sendergang.cc:75-82)
   * and the sending-time increment (sendergang.cc:163-173).  See the loop below. */
  auto chain_ok = [&]() { return REMY_CHAIN && !err && t < duration && dcons == ddeliv && ninc < INC_BUF; };
  auto begin_chained_tick = [&](double t2) {
    iters++;
    prng_before_shuffle = prng;
    prng = shuffle_skip(prng, n);
    zero_window_lookups(0);
    if (on_mask) { tm.st(ninc, share_of(__dsub_rn(t2, t), k_on)); ninc++; }
    t = t2;
  };

  /* One pass of this loop = up to FOUR reference ticks.  The phases of a tick come in the reference's
   * order - senders (switches, shuffle, acks, transmissions), link, delay (network.cc:54-61) - and each
   * of the three expensive ones exists ONCE in the loop body.  A pass starts with whatever tick is next
   * (network.cc:75-78), the only place where switches and mail are handled.  In front of the
   * transmission, link and delay phases there is a "chain point":
   * if the tick so far is over (nothing else is due at its time) and the NEXT reference tick needs
   * nothing but that phase and the ones behind it - a pacing deadline in front of the transmissions,
   * a departure in front of the link phase, a delivery in front of the delay phase, and no switch -
   * time advances right there (begin_chained_tick) and the pass goes on as that tick.  On a busy link
   * every packet costs one transmission, one departure, one delivery and one ack tick; when they come
   * in that order all four share a pass, and every lane of a warp reaches every phase in most passes,
   * instead of each phase being run for the third of the lanes whose tick happens to need it.
   * Coincident events need no special case: a tick that transmits, has a departure AND a delivery runs
   * its phases in order in one pass.
   *
   * The loop is left at ONE place, its top, and nowhere inside a conditional block (an error only
   * sets `err` and lets the pass run out): early exits from inside divergent regions keep the
   * compiler from re-converging the warp behind them, and every phase below then runs once per
   * group of lanes instead of once per pass. */
  /* the send state of ONE sender is held in registers across passes: the sender acked last, or else the
   * one whose pacing deadline comes next (requested at the end of the transmission phase, a pass or more
   * before its transmission reads it) */
  uint32_t ak_s = 0xFFFFFFFFu, ak_seq = 0; int32_t ak_lim = 0; double ak_isend = 0;
  for (;;) {
    /* ---- next event (network.cc:75-78) ---- */
    const bool mail = dcons != ddeliv;                  /* Receiver::next_event_time: now (receiver.hh:34-37) */
    double tn = fmax(t, fmin(min_switch, next_send));   /* SenderGang::next_event_time */
    if (pending) tn = fmin(tn, pend_release);           /* Link::next_event_time */
    if (ddeliv != dtail) tn = fmin(tn, dhead_release);  /* Delay::next_event_time */
    if (mail) tn = t;
    if (err || !(t < duration)) break;
    if (tn > duration) { t = tn; break; } /* Network::_tickno keeps the overshooting time */
    const double t_old = t;
    t = tn;
    iters++;
    const bool advanced = t > t_old;
    bool maybe_sends = next_send <= t; /* a window or pacing deadline changed in this tick, or a deadline is due */

    /* the buffered sending-time increments are applied when the buffer of ANY lane that is here could
     * overflow in this pass (this tick and the chained ones): every such lane then flushes in the same
     * pass; a switch needs the sums up to date */
    {
      const bool sw = min_switch <= t;
      const uint32_t fullest = warp_max_here(ninc);
      constexpr uint32_t PER_PASS = REMY_CHAIN ? 2 + 2 * REMY_SD_REPS : 1; /* increments a pass can add */
      if ((fullest + PER_PASS > INC_BUF || sw) && ninc) flush_share();
    }

    /* ---- 1. switch_senders (sendergang.cc:39-45, 89-106) ---- */
    if (min_switch <= t) {
      maybe_sends = true;
      ak_s = 0xFFFFFFFFu; /* (the switches rewrite send states) */
      uint32_t newly_on = 0;
      const uint32_t k0 = __popc(on_mask);
      double ms = DBL_MAX;
      for (uint32_t s = 0; s < n && !err; s++) {
        double nx = nsw[s];
        if (nx <= t) {
          if (nx != t) err |= REMY_ERR_ASSERT;
          SenderCold c = cold[s];
          bool sending = (on_mask >> s) & 1u;
          double itick = t_old;        /* SwitchedSender::internal_tick of a sender that was on */
          double shr = share[s];
          do {
            if (BYTESW && sending) { err |= REMY_ERR_ASSERT; break; } /* "never sets a time to switch off" (sendergang.cc:119) */
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
            if (BYTESW) { /* ByteSwitchedSender::switcher (sendergang.cc:108-138): no switch-off time; the length of the flow instead */
              nx = DBL_MAX;
              const double stop_rate = __ddiv_rn(1.0, cfgp->mean_on_duration);
              uint32_t flow = 0x7fffffffu; /* numeric_limits<int>::max(): "(almost) infinite flows" */
              if (stop_rate > 0) flow = (uint32_t)__double2ll_rn(ceil(exp_sample(prng, stop_rate))); /* lrint(ceil(sample)) */
              if (flow == 0) err |= REMY_ERR_ASSERT; /* assert( new_flow_length > 0 ) */
              c.pad += flow;
            } else
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
            tm.st(INC_BUF + s, FISH ? 0.0 : (open ? __dadd_rn(c.last_send_time, c.intersend) : DBL_MAX));
            open_mask = open ? (open_mask | bit) : (open_mask & ~bit);
            if (COUNT && !FISH) zero_mask = (c.window == 0) ? (zero_mask | bit) : (zero_mask & ~bit);
          } else {
            on_mask &= ~bit; open_mask &= ~bit;
            tm.st(INC_BUF + s, DBL_MAX);
            if (COUNT) zero_mask &= ~bit;
          }
        }
        ms = fmin(ms, nx);
      }
      min_switch = ms;
      /* ---- 5. accumulate_sending_time_until (sendergang.cc:163-173) of a switch tick: applied at once to the
       * senders that were on before it and still are (nothing is buffered here: ninc == 0) ---- */
      const uint32_t k1 = __popc(on_mask);
      k_on = (double)k1;
      if (k1) inv_on = __ddiv_rn(1.0, k_on);
      if (advanced) {
        uint32_t am = on_mask & ~newly_on;
        if (am) {
          const double inc = share_of(__dsub_rn(t, t_old), (double)k1);
          while (am) { const uint32_t s = __ffs(am) - 1; am &= am - 1; share[s] = __dadd_rn(share[s], inc); }
        }
      }
    } else if (advanced && on_mask) {
      /* ---- 5. accumulate_sending_time_until, buffered ---- */
      tm.st(ninc, share_of(__dsub_rn(t, t_old), k_on));
      ninc++;
    }

#if REMY_TOP_LOADS
    /* The two ring reads that the previous pass left to this one - the delay-line entry behind the new head,
     * and the packet that moved from the link's queue into service - are issued here, in front of the shuffle's
     * chain of LCG steps (which needs no memory).  Their data is not needed
     * before this pass's delivery / departure; issued where they arose - at the END of the previous pass - they
     * were waited for at whatever scoreboard wait came next, usually at once (ptxas drains the scoreboards at the
     * loop header when it finds loads in flight across the back edge). */
    if (dn_release < 0.0) {
      const DelayEnt h = dring[(ddeliv + 1) & dmask];
      dn_release = h.release; dn_sent = h.tick_sent; dn_seq = h.seq; dn_src = h.src;
    }
    if (pend_src == PEND_STALE) {
      const LinkEnt e = lring[(lhead - 1) & lmask];
      pend_sent = e.tick_sent; pend_seq = e.seq; pend_src = e.src;
    }
#endif

    /* ---- 2. run_senders: the shuffle's PRNG draws (sendergang.cc:75-82) ---- */
    prng_before_shuffle = prng;
    prng = shuffle_skip(prng, n);
    if (!mail) zero_window_lookups(0);

    /* ---- 3. receive_feedback for every mailbox (sendergang.cc:175-188) ---- */
    /* the sender acked in this tick usually transmits in this tick: its send state is handed
     * over in registers instead of being re-read from HBM */
    uint32_t acked = 0; /* senders whose mailbox was read in this tick */
    if (mail) {
      uint32_t done = 0, reopened = 0;
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
          /* Rat::send of this same tick (rat-templates.cc:13-18) looks the whisker up again when the window is 0
           * and applies it to a window of 0: the same leaf (the Memory has not changed), but window(0) = clamp(b)
           * differs from the 0 = clamp(int(w m + b)) above when m < 0 or w m overflowed - hand-made trees only,
           * the optimiser keeps 0 <= m <= 1.  (The lookup itself is counted / traced with the other zero-window
           * lookups of the tick, below.) */
          if (c.window == 0 && ((on_mask >> s) & 1u)) {
            const int32_t w0 = whisker_window(0, w.window_multiple, w.window_increment);
            if (w0 != 0) { c.window = w0; reopened |= 1u << s; }
          }
        }
        c.leaf = leaf;
        cold[s] = c;
        ak_s = s; ak_seq = c.packets_sent; ak_lim = FISH ? c.window : c.largest_ack + 1 + c.window; ak_isend = c.intersend;
        { SendState ss; ss.packets_sent = ak_seq; ss.lim = ak_lim; ss.intersend = ak_isend; sendst[s] = ss; }
        const uint32_t bit = 1u << s;
        if (on_mask & bit) {
          const bool open = FISH || (int32_t)c.packets_sent < c.largest_ack + 1 + c.window;
          tm.st(INC_BUF + s, FISH ? fish_due : (open ? __dadd_rn(c.last_send_time, c.intersend) : DBL_MAX));
          open_mask = open ? (open_mask | bit) : (open_mask & ~bit);
          if (COUNT && !FISH) zero_mask = (c.window == 0 || ((reopened >> s) & 1u)) ? (zero_mask | bit) : (zero_mask & ~bit);
        }
      }
      dcons = ddeliv;
      acked = done;
      maybe_sends = true;
      zero_window_lookups(acked);
      if (COUNT) zero_mask &= ~reopened;
    }

    /* (the transmission and link phases, with their chain points, can be gone through more than once per
     * pass: between two deliveries a busy link sees one departure and about one transmission, in either
     * order) */
#if REMY_SD_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int rep = 0; rep < REMY_SD_REPS; rep++) {
    /* ---- chain point S: nothing to transmit in this tick and nothing else due at its time, and the next
     * reference tick starts with a pacing deadline (a departure or delivery at that same time follows in
     * the phases below) ---- */
    if (!maybe_sends && chain_ok()) {
      const double t2 = next_send;
      if (t2 > t && t2 <= duration && t2 < min_switch && !(pending && pend_release < t2) && !(ddeliv != dtail && dhead_release < t2)) {
        begin_chained_tick(t2);
        maybe_sends = true;
      }
    }

    /* ---- 4. Rat::send for every sender whose deadline is due (rat-templates.cc:20-34) (the scan is skipped
     * when no deadline is due and nothing changed: next_send stays valid) ---- */
    if (maybe_sends) {
      uint32_t ready = 0, ns_sender = 0xFFFFFFFFu;
      uint32_t first_send = 0; /* Fish that looked their lambda up in this tick's send (trace) */
      double ns = DBL_MAX;
      uint32_t om = open_mask;
      while (om) {
        const uint32_t s = __ffs(om) - 1; om &= om - 1;
        const double st = tm.ld(INC_BUF + s);
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
                if (ltail - lhead + REMY_TOP_LOADS >= A.link_cap) { err |= REMY_ERR_LINK_OVERFLOW; break; }
                LinkEnt e; e.tick_sent = t; e.seq = seq; e.src = s;
                lring[ltail & lmask] = e; ltail++;
              }
            }
            uint32_t fp = (uint32_t)lim;
            const double x = exp_sample(fp, isend), mx = __ddiv_rn(2.0, isend);
            if (s == ak_s) { ak_seq = seq; ak_lim = (int32_t)fp; ak_isend = isend; } /* (the phase can come again in this pass) */
            const double st = __dadd_rn(t, __dmul_rn((double)FISH_BATCH, (mx < x) ? mx : x)); /* _update_send_time(tickno) */
            c->packets_sent = seq; c->last_send_time = t; c->intersend = isend; c->window = (int32_t)fp;
            { SendState ss; ss.packets_sent = seq; ss.lim = (int32_t)fp; ss.intersend = isend; sendst[s] = ss; }
            tm.st(INC_BUF + s, st);
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
            if (ltail - lhead + REMY_TOP_LOADS >= A.link_cap) { err |= REMY_ERR_LINK_OVERFLOW; break; }
            LinkEnt e; e.tick_sent = t; e.seq = seq; e.src = s;
            lring[ltail & lmask] = e; ltail++;
          }
          bool open = (int32_t)(seq + 1) < lim;
          if (BYTESW && seq + 1 == c->pad) {
            /* the flow is complete: ByteSwitchedSender::tick (sendergang.cc:270-276) switches the sender off - its sending
             * time up to this tick is already buffered - and draws the idle time to its next flow */
            flush_share();
            const uint32_t bit = 1u << s;
            on_mask &= ~bit; open = false;
            if (COUNT) zero_mask &= ~bit;
            const double nx = __dadd_rn(t, exp_sample(prng, __ddiv_rn(1.0, cfgp->mean_off_duration)));
            nsw[s] = nx;
            min_switch = fmin(min_switch, nx);
            const uint32_t k = __popc(on_mask);
            k_on = (double)k;
            if (k) inv_on = __ddiv_rn(1.0, k_on);
          }
          const double st = open ? __dadd_rn(t, isend) : DBL_MAX;
          tm.st(INC_BUF + s, st);
          if (!open) open_mask &= ~(1u << s);
          if (st < ns) { ns = st; ns_sender = s; }
        }
      }
      next_send = ns;
      /* trace of a Fish tick: per sender, in the shuffled order, the ack's lookup or (never both:
       * an ack sets lambda) the first send's lookup on the all-zero Memory of the reset */
      if (TRACE && FISH) { track_senders(acked, first_send); acked = 0; }
      if (ns_sender != 0xFFFFFFFFu && ns_sender != ak_s) { /* what its transmission will read */
        const SendState ss = sendst[ns_sender];
        ak_s = ns_sender; ak_seq = ss.packets_sent; ak_lim = ss.lim; ak_isend = ss.intersend;
      }
      maybe_sends = false;
    }

    /* ---- chain point D: this tick is over (no departure or delivery is due at its time), and the next
     * reference tick starts with a departure: no switch and no pacing deadline at or before it, no
     * delivery before it ---- */
    if (chain_ok() && pending && pend_release > t) {
      const double t2 = pend_release;
      if (t2 <= duration && t2 < min_switch && t2 < next_send && !(ddeliv != dtail && dhead_release < t2)) begin_chained_tick(t2);
    }

    /* ---- 6. Link::tick (link-templates.cc:7-18) -> StochasticLoss (stochastic-loss.hh:21-35) -> Delay::accept ---- */
    if (pending && pend_release <= t) {
      if (pend_release != t) err |= REMY_ERR_ASSERT;
#if REMY_TOP_LOADS
      if (pend_src == PEND_STALE) { /* (rare: it entered service in this very pass) */
        const LinkEnt e = lring[(lhead - 1) & lmask];
        pend_sent = e.tick_sent; pend_seq = e.seq; pend_src = e.src;
      }
#endif
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
#if REMY_TOP_LOADS
      lhead++; /* (its slot stays untouched until it has been read: the ring keeps one entry in reserve) */
      pending = true; pend_release = __dadd_rn(t, service); pend_src = PEND_STALE;
#else
      const LinkEnt e = lring[lhead & lmask]; lhead++;
      pending = true; pend_release = __dadd_rn(t, service); pend_sent = e.tick_sent; pend_seq = e.seq; pend_src = e.src;
#endif
    }
    } /* rep */

    /* ---- chain point V: no delivery is due in this tick, and the next reference tick is a pure delivery:
     * no switch, pacing deadline or departure at or before the head of the delay line ---- */
    if (chain_ok() && ddeliv != dtail && dhead_release > t) {
      const double t2 = dhead_release;
      if (t2 <= duration && t2 < min_switch && t2 < next_send && !(pending && pend_release <= t2)) begin_chained_tick(t2);
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
          if (ddeliv != dtail) {
            /* (the loaded fields are used - checked - right here so that the wait for this rare load stays on
             * this rare path; left to the first use after the join, it made every delivery wait for the ring
             * read issued below, the one that exists so that nothing has to wait for it) */
            const DelayEnt h = dring[ddeliv & dmask]; dhead_release = h.release; dh_sent = h.tick_sent; dh_seq = h.seq; dh_src = h.src;
            if (!(dh_sent <= t) || dh_src >= n || dh_seq == 0xFFFFFFFFu) err |= REMY_ERR_ASSERT;
          }
        }
      }
      prefetch_l2(&cold[m_src]); /* the ack's sender record: read at the top of the next pass */
#if REMY_DRING_PF
      if (dtail - ddeliv > REMY_DRING_PF) prefetch_l2(&dring[(ddeliv + REMY_DRING_PF) & dmask]); /* the ring entry read a few deliveries from now */
#endif
#if REMY_TOP_LOADS
      if (dtail - ddeliv >= 2) dn_release = -1.0; /* the new second entry: read at the top of the next pass (release times are >= 0) */
#else
      if (ddeliv != dtail) {
        if (dtail - ddeliv >= 2) { /* read the new second entry now; it is not used before the next delivery */
          const DelayEnt h = dring[(ddeliv + 1) & dmask];
          dn_release = h.release; dn_sent = h.tick_sent; dn_seq = h.seq; dn_src = h.src;
        }
      }
#endif
    }
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

#ifndef REMY_SIM_BLOCK
#define REMY_SIM_BLOCK 128
#endif
constexpr int SIM_BLOCK = REMY_SIM_BLOCK;
/* dynamic shared memory a block may ask for (the sm_100 opt-in maximum is 227 KB per block) */
constexpr size_t SIM_SMEM_MAX = 226 * 1024;
#ifndef REMY_MIN_BLOCKS_PER_SM
/* Three resident blocks per SM = 168 registers per thread and no spills.  Four (128 registers, 44 B of
 * spills and re-reads of per-simulation constants) gave 33 % more warps but measured 10 % slower on the
 * final kernel (8.47 vs 9.34 G event ticks/s, 102 400 simulations x 3e4 ticks); two (255 registers) 18 % slower. */
#define REMY_MIN_BLOCKS_PER_SM 3
#endif

#ifndef REMY_HOST_EMU

/* TSM: the tree is staged in shared memory.  A template parameter and not a branch, so that the tree's
 * pointers are shared-memory pointers at compile time and its reads ld.shared (with a run-time choice they were
 * generic loads, which ptxas tracks on the same scoreboard as the global loads of the sender records). */
template <bool COUNT, bool TRACE = false, bool FISH = false, bool BYTESW = false, bool TSM = true>
__global__ void __launch_bounds__(SIM_BLOCK, REMY_MIN_BLOCKS_PER_SM) remy_sim_kernel(const KernelArgs A)
{
  extern __shared__ __align__(16) unsigned char smem[];
  /* layout: [inc_buf: INC_BUF*B doubles][hot_send: n_max*B doubles][tree, if staged] */
  TickMem tm;
  tm.base = (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * (uint32_t)sizeof(double);
  tm.stride = SIM_BLOCK * (uint32_t)sizeof(double); /* (launched with SIM_BLOCK threads: capi.cu) */
  TreeView T;
  T.axes_mask = A.axes_mask;
  if (TSM) {
    unsigned char *p = smem + (size_t)(INC_BUF + A.n_max) * SIM_BLOCK * sizeof(double);
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
  /* (a batch with fewer simulations than resident threads is spread over all warps, see capi.cu: the lanes
   * beyond lanes_per_warp stay idle, so that a warp executes the union of FEWER simulations' code paths) */
  if ((threadIdx.x & 31u) >= A.lanes_per_warp) return;
  for (;;) {
    const uint32_t w = atomicAdd(A.work_counter, 1u);
    if (w >= A.n_work) break;
    run_sim<COUNT, TRACE, FISH, BYTESW>(A, T, A.order[w], tm, share, nsw, sendst, cold, lring, dring, cnt);
  }
}

#endif /* !REMY_HOST_EMU */

} // namespace remy
#endif
