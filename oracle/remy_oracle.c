/*
 * remy_oracle.c — CPU restatement of the reference's scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path; it is
 * never linked into, imported by or called from the shipped library.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it.
 *
 * It restates, literally and without the CUDA kernel's restructuring, what
 *   Evaluator<WhiskerTree>::score(tree, seed, {config}, false, ticks)
 * (and its Fish / ByteSwitchedSender variants, `kind`) does in CN-TU/remy for ONE config (reference src/evaluator.cc:77-103): build a
 * Network, run the next-event loop, read the utility.  Every function cites the
 * reference lines it follows.  Arithmetic that lives outside /root/reference
 * (libstdc++ 13 <random>/<algorithm>, glibc 2.39 log) is restated from the
 * published algorithms: /usr/include/c++/13/bits/{random.h,random.tcc,
 * uniform_int_dist.h,stl_algo.h} and remy_b200/csrc/exact_log.h.
 *
 * Pinning: the reference holds no bit-exact golden vectors for this path
 * (SURVEY.md §4, §8c), so this restatement is pinned against the reference
 * itself: oracle/build_ref.sh compiles the reference's own sources into
 * oracle/_ref/libremy_ref.so, and tests/test_oracle.py (Rat), tests/test_fish.py (Fish),
 * tests/test_flows.py (ByteSwitchedSender) and tests/test_trace.py (trace == true) check that
 * every integer observable and the PRNG end state match and utilities agree bit for bit;
 * tests/golden/ (JSON files) holds vectors generated from that reference build.
 * One piece is "parity unpinned": Boost's P^2 median (no Boost in the image).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/remy_b200.h"
#include "../remy_b200/csrc/exact_log.h"

/* ------------------------------------------------------------------ PRNG -- */
/* PRNG = std::default_random_engine = minstd_rand0 (src/random.hh:6;
 * bits/random.h:1592): x <- 16807 x mod (2^31 - 1). */
#define LCG_M 2147483647u
typedef struct { uint32_t x; } prng_t;

/* linear_congruential_engine::seed (bits/random.tcc:116-126): c == 0, so a seed
 * that is 0 mod m becomes 1. */
static void prng_seed(prng_t *g, uint32_t s)
{
  g->x = s % LCG_M;
  if (g->x == 0) g->x = 1;
}
static inline uint32_t prng_next(prng_t *g)
{
  g->x = (uint32_t)(((uint64_t)g->x * 16807u) % LCG_M);
  return g->x;
}

/* std::generate_canonical<double,53>(minstd_rand0) (bits/random.tcc:3349-3383):
 * range r = 2147483646, two draws, sum in double, divide by (double)(r*r). */
static double prng_canonical(prng_t *g)
{
  double sum = 0.0, tmp = 1.0;
  sum += (double)(prng_next(g) - 1u) * tmp;
  tmp = 2147483646.0;
  sum += (double)(prng_next(g) - 1u) * tmp;
  tmp = 0x1.fffffffp+61; /* (double)(2147483646.0L * 2147483646.0L) */
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}

/* Exponential::sample (src/exponential.hh:22) ->
 * std::exponential_distribution::operator() (bits/random.h:4898-4905). */
static double exp_sample(prng_t *g, double rate)
{
  return -remy_exact_log(1.0 - prng_canonical(g)) / rate;
}

/* std::bernoulli_distribution::operator() (bits/random.h:3739-3750). */
static int bernoulli(prng_t *g, double p) { return prng_canonical(g) < p; }

/* uniform_int_distribution{0,u} on minstd_rand0: the "fallback" downscaling
 * branch (bits/uniform_int_dist.h:333-339), urngrange = 2147483645. */
static uint64_t uniform_int(prng_t *g, uint64_t u)
{
  const uint64_t urngrange = 2147483645u, uerange = u + 1;
  const uint64_t scaling = urngrange / uerange, past = uerange * scaling;
  uint64_t ret;
  do ret = (uint64_t)prng_next(g) - 1u; while (ret >= past);
  return ret / scaling;
}

/* std::shuffle, paired path (bits/stl_algo.h:3742-3805; n <= 46340). */
static void shuffle_indices(prng_t *g, uint32_t *idx, uint32_t n)
{
  if (n == 0) return;
  uint32_t i = 1, t;
  if ((n % 2) == 0) {
    uint32_t d = (uint32_t)uniform_int(g, 1);
    t = idx[1]; idx[1] = idx[d]; idx[d] = t;
    i = 2;
  }
  while (i != n) {
    const uint64_t swap_range = (uint64_t)i + 1;
    const uint64_t b1 = swap_range + 1;
    const uint64_t x = uniform_int(g, swap_range * b1 - 1);
    const uint32_t p0 = (uint32_t)(x / b1), p1 = (uint32_t)(x % b1);
    t = idx[i]; idx[i] = idx[p0]; idx[p0] = t; i++;
    t = idx[i]; idx[i] = idx[p1]; idx[p1] = t; i++;
  }
}

/* --------------------------------------------------------------- packets -- */
/* remy::Packet (src/packet.hh:5-30); flow_id is written but never read on the
 * Rat path (the flow filter is commented out, src/memory.cc:35-38). */
typedef struct { uint32_t src; double tick_sent, tick_received; long seq_num; } pkt_t;

typedef struct { double release; pkt_t p; } dent_t;
typedef struct { dent_t *a; size_t cap, head, len; } dq_t; /* std::deque in Delay (src/delay.hh:16) */
typedef struct { pkt_t *a; size_t cap, head, len; } pq_t;  /* Link::_buffer (src/link.hh:15) */

static void dq_push(dq_t *q, double release, const pkt_t *p)
{
  if (q->len == q->cap) {
    size_t ncap = q->cap ? q->cap * 2 : 64;
    dent_t *na = (dent_t *)malloc(ncap * sizeof(dent_t));
    for (size_t i = 0; i < q->len; i++) na[i] = q->a[(q->head + i) % q->cap];
    free(q->a); q->a = na; q->cap = ncap; q->head = 0;
  }
  dent_t *e = &q->a[(q->head + q->len) % q->cap];
  e->release = release; e->p = *p; q->len++;
}
static void pq_push(pq_t *q, const pkt_t *p)
{
  if (q->len == q->cap) {
    size_t ncap = q->cap ? q->cap * 2 : 64;
    pkt_t *na = (pkt_t *)malloc(ncap * sizeof(pkt_t));
    for (size_t i = 0; i < q->len; i++) na[i] = q->a[(q->head + i) % q->cap];
    free(q->a); q->a = na; q->cap = ncap; q->head = 0;
  }
  q->a[(q->head + q->len) % q->cap] = *p; q->len++;
}

/* ---------------------------------------------------------------- Memory -- */
/* Memory (src/memory.hh:20-41); only fields read on the Rat path are kept.
 * field(a) for a = 0..5 (memory.hh:97). */
typedef struct {
  double f[REMY_NUM_AXES]; /* rec_send_ewma, rec_rec_ewma, rtt_ratio, slow_rec_rec_ewma, rtt_diff, queueing_delay */
  double last_tick_sent, last_tick_received, min_rtt;
} memory_t;

/* Memory::packets_received (src/memory.cc:31-80), one packet. */
static void memory_packet_received(memory_t *m, const pkt_t *x, int largest_ack, uint32_t *err)
{
  const double alpha = 1.0 / 8.0, slow_alpha = 1.0 / 256.0;
  const double rtt = x->tick_received - x->tick_sent;
  int pkt_outstanding = 1;
  if (x->seq_num > largest_ack) pkt_outstanding = (int)(x->seq_num - largest_ack);
  if (m->min_rtt <= 0) m->min_rtt = rtt;
  if (m->last_tick_sent == 0 || m->last_tick_received == 0) {
    m->last_tick_sent = x->tick_sent;
    m->last_tick_received = x->tick_received;
    m->min_rtt = rtt;
  } else {
    m->f[0] = (1 - alpha) * m->f[0] + alpha * (x->tick_sent - m->last_tick_sent);
    m->f[1] = (1 - alpha) * m->f[1] + alpha * (x->tick_received - m->last_tick_received);
    m->f[3] = (1 - slow_alpha) * m->f[3] + slow_alpha * (x->tick_received - m->last_tick_received);
    m->last_tick_sent = x->tick_sent;
    m->last_tick_received = x->tick_received;
    m->min_rtt = rtt < m->min_rtt ? rtt : m->min_rtt; /* std::min(_min_rtt, rtt) */
    m->f[2] = rtt / m->min_rtt;
    m->f[4] = rtt - m->min_rtt;
    if (!(m->f[2] >= 1.0) || !(m->f[4] >= 0)) *err |= REMY_ERR_ASSERT; /* memory.cc:68-69 */
    m->f[5] = m->f[1] * pkt_outstanding;
  }
}

/* ------------------------------------------------------------ whiskers ---- */
typedef struct {
  const RemyFlatTree *tree;
  const RemyLeafOverride *ov; /* may be NULL */
  uint32_t *use_counts;       /* may be NULL; [n_leaves] */
  /* trace == true (Rat::_track, src/rat.hh:22): every use_whisker call reports the leaf it
   * returned and the Memory it was asked about (MemoryRange::track, src/memoryrange.cc:60-66) */
  void (*track)(void *user, uint32_t leaf, const double *mem);
  void *track_user;
} policy_t;

/* MemoryRange::contains (src/memoryrange.cc:52-58). */
static int node_contains(const RemyTreeNode *nd, const memory_t *m)
{
  for (int a = 0; a < REMY_NUM_AXES; a++) {
    if (!(nd->axis_mask & (1u << a))) continue;
    if (!((m->f[a] >= nd->lower[a]) && (m->f[a] < nd->upper[a]))) return 0;
  }
  return 1;
}

/* WhiskerTree::use_whisker / whisker (src/whiskertree.cc:42-82): first-match
 * descent; a use is counted on the returned leaf (action.hh:36). */
static int use_whisker(const policy_t *pol, const memory_t *m, RemyLeafAction *out, uint32_t *err)
{
  const RemyFlatTree *t = pol->tree;
  uint32_t cur = 0;
  if (!node_contains(&t->nodes[0], m)) { *err |= REMY_ERR_NO_WHISKER; return 0; }
  while (t->nodes[cur].child_count) {
    const RemyTreeNode *nd = &t->nodes[cur];
    uint32_t next = UINT32_MAX;
    for (uint32_t c = 0; c < nd->child_count; c++) {
      if (node_contains(&t->nodes[nd->child_begin + c], m)) { next = nd->child_begin + c; break; }
    }
    if (next == UINT32_MAX) { *err |= REMY_ERR_NO_CHILD; return 0; }
    cur = next;
  }
  const uint32_t leaf = t->nodes[cur].leaf_index;
  *out = t->leaves[leaf];
  if (pol->ov && pol->ov->leaf_index == leaf) { /* WhiskerTree::replace, whiskertree.cc:137-157 */
    out->window_increment = pol->ov->window_increment;
    out->window_multiple = pol->ov->window_multiple;
    out->intersend = pol->ov->intersend;
  }
  if (pol->use_counts) pol->use_counts[leaf]++;
  if (pol->track) pol->track(pol->track_user, leaf, m->f); /* whiskertree.cc:55-57 */
  return 1;
}

/* Whisker::window (src/whisker.hh:25).  int(double) out of range is what the
 * reference's x86-64 build does (cvttsd2si -> INT_MIN). */
static int whisker_window(const RemyLeafAction *w, unsigned previous_window)
{
  const double v = previous_window * w->window_multiple + w->window_increment;
  int iv;
  if (v >= 2147483648.0 || v <= -2147483649.0 || v != v) iv = INT_MIN; else iv = (int)v;
  int r = iv > 0 ? iv : 0;
  return r < 1000000 ? r : 1000000;
}

/* ----------------------------------------------------------------- Rat ---- */
typedef struct {
  memory_t memory;
  unsigned packets_sent, packets_received;
  double last_send_time;
  int the_window;
  double intersend_time;
  unsigned flow_id;
  int largest_ack;
} rat_t;

typedef struct { double tick_share_sending; unsigned packets_received; double total_delay; } utility_t;

/* Fish (src/fish.hh:16-36): Poisson-paced sender, the second sender family (Evaluator<FinTree>).
 * Shares memory / packets_sent / flow_id / largest_ack / last_send_time with rat_t. */
typedef struct {
  double next_send_time, lambda, max_intersend;
  prng_t prng;             /* Fish::_prng: every sender's copy starts from the same seed (sendergang.cc:9-26 copies the exemplar) */
} fish_t;
#define FISH_BATCH_SIZE 5u  /* Fish::_batch_size (src/fish.cc:20) */

typedef struct {
  double internal_tick, next_switch_tick;
  rat_t rat;
  fish_t fish;
  utility_t utility;
  int sending;
  unsigned id;
  unsigned packets_sent_cap; /* ByteSwitchedSender::packets_sent_cap_ (src/sendergang.hh:77) */
} sender_t;

typedef struct { pkt_t *a; size_t len, cap; } mailbox_t; /* Receiver::_collector[src] (src/receiver.hh:13) */

typedef struct {
  prng_t prng;
  policy_t pol;
  uint32_t n;
  sender_t *gang;
  double start_rate, stop_rate;
  /* Link (src/link.hh): _pending_packet is a Delay of 1.0/rate */
  dq_t pending; double service; unsigned limit; pq_t buffer;
  /* StochasticLoss buffer (src/stochastic-loss.hh:14) */
  pq_t loss_buf; double loss_tick; double loss_rate;
  /* Delay (src/delay.hh) */
  dq_t delay_q; double delay;
  /* Receiver */
  mailbox_t *collector; uint32_t collector_size;
  double tickno;
  uint32_t err;
  int kind;                /* 0 = Rat over a WhiskerTree, 1 = Fish over a FinTree (leaf lambda in RemyLeafAction.intersend),
                            * 2 = Rat behind ByteSwitchedSender: flows of a random length in packets (mean_on_duration = mean length) */
} net_t;

/* Rat::reset (src/rat.cc:34-48); Memory::reset (src/memory.hh:93). */
static void rat_reset(net_t *N, rat_t *r)
{
  memset(&r->memory, 0, sizeof r->memory);
  r->last_send_time = 0; r->the_window = 0; r->intersend_time = 0;
  r->flow_id++;
  r->largest_ack = (int)(r->packets_sent - 1);
  RemyLeafAction w;
  if (!use_whisker(&N->pol, &r->memory, &w, &N->err)) return;
  r->the_window = whisker_window(&w, (unsigned)r->the_window);
  r->intersend_time = w.intersend;
}

/* Rat::packets_received (src/rat.cc:22-32). */
static void rat_packets_received(net_t *N, rat_t *r, const pkt_t *pk, size_t n)
{
  r->packets_received += (unsigned)n;
  for (size_t i = 0; i < n; i++) memory_packet_received(&r->memory, &pk[i], r->largest_ack, &N->err);
  const int last = (int)pk[n - 1].seq_num;
  r->largest_ack = last > r->largest_ack ? last : r->largest_ack;
  RemyLeafAction w;
  if (!use_whisker(&N->pol, &r->memory, &w, &N->err)) return;
  r->the_window = whisker_window(&w, (unsigned)r->the_window);
  r->intersend_time = w.intersend;
}

/* Rat::next_event_time (src/rat.cc:50-62). */
static double rat_next_event_time(const rat_t *r, double tickno)
{
  if ((int)r->packets_sent < r->largest_ack + 1 + r->the_window) {
    if (r->last_send_time + r->intersend_time <= tickno) return tickno;
    return r->last_send_time + r->intersend_time;
  }
  return DBL_MAX;
}

/* Link::accept (src/link.hh:26-34); Delay::accept (src/delay.hh:42-51). */
static void link_accept(net_t *N, const pkt_t *p, double tickno)
{
  if (N->pending.len == 0) {
    dq_push(&N->pending, tickno + N->service, p);
  } else if (N->limit && N->buffer.len < N->limit) {
    pq_push(&N->buffer, p);
  }
}

/* Rat::send (src/rat-templates.cc:8-35). */
static void rat_send_capped(net_t *N, rat_t *r, unsigned id, double tickno, unsigned packets_sent_cap);
static void rat_send(net_t *N, rat_t *r, unsigned id, double tickno)
{
  rat_send_capped(N, r, id, tickno, 0xFFFFFFFFu); /* default argument: numeric_limits<unsigned int>::max() (src/rat.hh) */
}
static void rat_send_capped(net_t *N, rat_t *r, unsigned id, double tickno, unsigned packets_sent_cap)
{
  if (r->the_window == 0) {
    RemyLeafAction w;
    if (!use_whisker(&N->pol, &r->memory, &w, &N->err)) return;
    r->the_window = whisker_window(&w, (unsigned)r->the_window);
    r->intersend_time = w.intersend;
  }
  if (((int)r->packets_sent < r->largest_ack + 1 + r->the_window)
      && (r->last_send_time + r->intersend_time <= tickno)) {
    if (r->packets_sent >= packets_sent_cap) return; /* "Have we reached the end of the flow for now?" (rat-templates.cc:23-26) */
    pkt_t p; p.src = id; p.tick_sent = tickno; p.tick_received = -1; p.seq_num = r->packets_sent;
    r->packets_sent++;
    link_accept(N, &p, tickno);
    r->last_send_time = tickno;
  }
}

/* ---------------------------------------------------------------- Fish ---- */
/* Fish::_update_lambda / _update_send_time (src/fish.cc:59-70); Exponential::sample with the
 * sender's own PRNG (src/exponential.hh:22). */
static void fish_update_lambda(rat_t *r, fish_t *f, double lambda)
{
  (void)r;
  f->lambda = lambda;
  f->max_intersend = 2.0 / lambda;
}
static void fish_update_send_time(rat_t *r, fish_t *f, double tickno)
{
  r->last_send_time = tickno;
  const double x = exp_sample(&f->prng, f->lambda);
  f->next_send_time = r->last_send_time + FISH_BATCH_SIZE * (x < f->max_intersend ? x : f->max_intersend); /* std::min(sample, max) */
}
/* Fish::reset (src/fish.cc:36-48) */
static void fish_reset(rat_t *r, fish_t *f)
{
  memset(&r->memory, 0, sizeof r->memory);
  r->last_send_time = 0; f->next_send_time = 0; f->lambda = 0; f->max_intersend = 0;
  r->flow_id++;
  r->largest_ack = (int)r->packets_sent - 1;
}
/* Fish::packets_received (src/fish.cc:26-34) */
static void fish_packets_received(net_t *N, rat_t *r, fish_t *f, const pkt_t *pk, size_t n)
{
  r->packets_received += (unsigned)n;
  for (size_t i = 0; i < n; i++) memory_packet_received(&r->memory, &pk[i], r->largest_ack, &N->err);
  const int last = (int)pk[n - 1].seq_num;
  r->largest_ack = last > r->largest_ack ? last : r->largest_ack;
  RemyLeafAction w;
  if (!use_whisker(&N->pol, &r->memory, &w, &N->err)) return; /* FinTree::use_fin (src/fintree.cc:39-57) */
  fish_update_lambda(r, f, w.intersend);
  fish_update_send_time(r, f, r->last_send_time);
}
/* Fish::next_event_time (src/fish.cc:50-57) */
static double fish_next_event_time(const fish_t *f, double tickno)
{
  return f->next_send_time == 0 ? tickno : f->next_send_time;
}
/* Fish::send (src/fish-templates.cc:8-27) */
static void fish_send(net_t *N, rat_t *r, fish_t *f, unsigned id, double tickno)
{
  if (tickno < f->next_send_time) return;
  if (f->lambda == 0) {
    RemyLeafAction w;
    if (!use_whisker(&N->pol, &r->memory, &w, &N->err)) return;
    fish_update_lambda(r, f, w.intersend);
  }
  for (unsigned i = 0; i < FISH_BATCH_SIZE; i++) {
    pkt_t p; p.src = id; p.tick_sent = tickno; p.tick_received = -1; p.seq_num = r->packets_sent;
    r->packets_sent++;
    link_accept(N, &p, tickno);
  }
  fish_update_send_time(r, f, tickno);
}

/* SwitchedSender::accumulate_sending_time_until (src/sendergang.cc:163-173),
 * Utility::sending_duration (src/utility.hh:19). */
static void accumulate(sender_t *s, double tickno, unsigned num_sending)
{
  if (tickno > s->internal_tick) {
    s->utility.tick_share_sending += (tickno - s->internal_tick) / (double)num_sending;
    s->internal_tick = tickno;
  }
}

/* TimeSwitchedSender::switcher (src/sendergang.cc:89-106) with switch_on/off
 * (:140-161). */
static void switcher(net_t *N, sender_t *s, double tickno, unsigned num_sending)
{
  while (s->next_switch_tick <= tickno) {
    if (s->next_switch_tick != tickno) N->err |= REMY_ERR_ASSERT; /* sendergang.cc:98 */
    if (s->sending) { accumulate(s, tickno, num_sending); s->sending = 0; }
    else { s->sending = 1; if (N->kind == 1) fish_reset(&s->rat, &s->fish); else rat_reset(N, &s->rat); s->internal_tick = tickno; }
    s->next_switch_tick += exp_sample(&N->prng, s->sending ? N->stop_rate : N->start_rate);
    if (N->err & (REMY_ERR_NO_WHISKER | REMY_ERR_NO_CHILD)) return;
  }
}

/* ByteSwitchedSender::switcher (src/sendergang.cc:108-138): switched on by the clock, off by the flow's length. */
static void byte_switcher(net_t *N, sender_t *s, double tickno)
{
  while (s->next_switch_tick <= tickno) {
    if (s->next_switch_tick != tickno) N->err |= REMY_ERR_ASSERT;
    if (s->sending) { N->err |= REMY_ERR_ASSERT; return; }       /* "never sets a time to switch off" */
    s->sending = 1; rat_reset(N, &s->rat); s->internal_tick = tickno; /* switch_on (sendergang.cc:140-149) */
    s->next_switch_tick = DBL_MAX;
    unsigned new_flow_length;
    if (N->stop_rate > 0) new_flow_length = (unsigned)lrint(ceil(exp_sample(&N->prng, N->stop_rate)));
    else new_flow_length = (unsigned)INT_MAX;                       /* "Using (almost) infinite flows..." */
    if (!(new_flow_length > 0)) N->err |= REMY_ERR_ASSERT;
    s->packets_sent_cap += new_flow_length;
    if (N->err & (REMY_ERR_NO_WHISKER | REMY_ERR_NO_CHILD)) return;
  }
}

/* TimeSwitchedSender::tick (src/sendergang.cc:190-205) with receive_feedback
 * (:175-188) and Utility::packets_received (src/utility.hh:20-27). */
static void sender_tick(net_t *N, sender_t *s, double tickno, unsigned num_sending)
{
  if (s->id < N->collector_size && N->collector[s->id].len) {
    mailbox_t *mb = &N->collector[s->id];
    s->utility.packets_received += (unsigned)mb->len;
    for (size_t i = 0; i < mb->len; i++) {
      if (!(mb->a[i].tick_received >= mb->a[i].tick_sent)) N->err |= REMY_ERR_ASSERT;
      s->utility.total_delay += mb->a[i].tick_received - mb->a[i].tick_sent;
    }
    if (N->kind == 1) fish_packets_received(N, &s->rat, &s->fish, mb->a, mb->len);
    else rat_packets_received(N, &s->rat, mb->a, mb->len);
    mb->len = 0;
  }
  if (s->sending && N->kind == 2) { /* ByteSwitchedSender::tick (src/sendergang.cc:256-278) */
    if (!(s->rat.packets_sent < s->packets_sent_cap)) N->err |= REMY_ERR_ASSERT;
    rat_send_capped(N, &s->rat, s->id, tickno, s->packets_sent_cap);
    accumulate(s, tickno, num_sending);
    if (s->rat.packets_sent == s->packets_sent_cap) { /* do we need to switch ourselves off? */
      accumulate(s, tickno, num_sending); s->sending = 0; /* switch_off */
      s->next_switch_tick = tickno + exp_sample(&N->prng, N->start_rate);
    }
  } else if (s->sending) {
    if (N->kind == 1) fish_send(N, &s->rat, &s->fish, s->id, tickno);
    else rat_send(N, &s->rat, s->id, tickno);
    accumulate(s, tickno, num_sending);
  }
}

static unsigned count_active(const net_t *N)
{
  unsigned k = 0;
  for (uint32_t i = 0; i < N->n; i++) k += N->gang[i].sending;
  return k;
}

/* Receiver::accept (src/receiver.cc:13-20). */
static void receiver_accept(net_t *N, const pkt_t *p, double tickno)
{
  mailbox_t *mb = &N->collector[p->src];
  if (p->src >= N->collector_size) N->collector_size = p->src + 1; /* autosize, receiver.cc:32-37 */
  if (mb->len == mb->cap) { mb->cap = mb->cap ? mb->cap * 2 : 4; mb->a = (pkt_t *)realloc(mb->a, mb->cap * sizeof(pkt_t)); }
  mb->a[mb->len] = *p; mb->a[mb->len].tick_received = tickno; mb->len++;
}

/* Network::tick (src/network.cc:54-61). */
static void net_tick(net_t *N)
{
  const double t = N->tickno;
  /* SenderGangofGangs::tick (src/sendergangofgangs.cc:44-53); gang 2 is empty. */
  unsigned num_sending = count_active(N);
  for (uint32_t i = 0; i < N->n; i++) {
    if (N->kind == 2) byte_switcher(N, &N->gang[i], t); else switcher(N, &N->gang[i], t, num_sending);
    if (N->err & (REMY_ERR_NO_WHISKER | REMY_ERR_NO_CHILD)) return;
  }
  num_sending = count_active(N);
  /* SenderGang::run_senders (src/sendergang.cc:68-87) */
  uint32_t idx[4096];
  for (uint32_t i = 0; i < N->n; i++) idx[i] = i;
  shuffle_indices(&N->prng, idx, N->n);
  for (uint32_t i = 0; i < N->n; i++) {
    sender_tick(N, &N->gang[idx[i]], t, num_sending);
    if (N->err & (REMY_ERR_NO_WHISKER | REMY_ERR_NO_CHILD)) return;
  }
  /* Link::tick (src/link-templates.cc:7-18): Delay::tick of the in-service
   * packet into StochasticLoss::accept (src/stochastic-loss.hh:30-35) */
  while (N->pending.len && N->pending.a[N->pending.head].release <= t) {
    dent_t *e = &N->pending.a[N->pending.head];
    if (e->release != t) N->err |= REMY_ERR_ASSERT; /* delay.hh:59 */
    if (!bernoulli(&N->prng, N->loss_rate)) { pq_push(&N->loss_buf, &e->p); }
    N->pending.head = (N->pending.head + 1) % N->pending.cap; N->pending.len--;
  }
  if (N->pending.len == 0 && N->buffer.len) {
    dq_push(&N->pending, t + N->service, &N->buffer.a[N->buffer.head]);
    N->buffer.head = (N->buffer.head + 1) % N->buffer.cap; N->buffer.len--;
  }
  /* StochasticLoss::tick (src/stochastic-loss.hh:21-29) -> Delay::accept */
  while (N->loss_buf.len) {
    dq_push(&N->delay_q, t + N->delay, &N->loss_buf.a[N->loss_buf.head]);
    N->loss_buf.head = (N->loss_buf.head + 1) % N->loss_buf.cap; N->loss_buf.len--;
  }
  /* Delay::tick (src/delay.hh:53-63) -> Receiver::accept */
  while (N->delay_q.len && N->delay_q.a[N->delay_q.head].release <= t) {
    dent_t *e = &N->delay_q.a[N->delay_q.head];
    if (e->release != t) N->err |= REMY_ERR_ASSERT;
    receiver_accept(N, &e->p, t);
    N->delay_q.head = (N->delay_q.head + 1) % N->delay_q.cap; N->delay_q.len--;
  }
}

static double dmin(double a, double b) { return b < a ? b : a; } /* std::min */

/* Utility::utility (src/utility.hh:46-60). */
static double utility_value(const utility_t *u)
{
  if (u->tick_share_sending == 0) return 0.0;
  if (u->packets_received == 0) return -INT_MAX;
  const double tput = (double)u->packets_received / u->tick_share_sending;
  const double del = u->total_delay / (double)u->packets_received;
  return log2(tput) - log2(del / 100.0);
}

/* One simulation: Network ctor (src/network.cc:24-37), SenderGang ctor
 * (src/sendergang.cc:9-26), run_simulation (src/network.cc:63-85), then
 * SenderGang::utility (src/sendergang.cc:280-293). */
void remy_oracle_run_sim_traced(const RemyFlatTree *tree, const RemyLeafOverride *ov,
                                const RemyNetConfig *cfg, uint32_t seed,
                                RemySimResult *out, RemySenderResult *out_senders,
                                uint32_t *use_counts,
                                void (*track)(void *, uint32_t, const double *), void *track_user);

void remy_oracle_run_sim(const RemyFlatTree *tree, const RemyLeafOverride *ov,
                         const RemyNetConfig *cfg, uint32_t seed,
                         RemySimResult *out, RemySenderResult *out_senders,
                         uint32_t *use_counts)
{
  remy_oracle_run_sim_traced(tree, ov, cfg, seed, out, out_senders, use_counts, NULL, NULL);
}

void remy_oracle_run_sim_kind(const RemyFlatTree *tree, const RemyLeafOverride *ov,
                              const RemyNetConfig *cfg, uint32_t seed, int kind,
                              RemySimResult *out, RemySenderResult *out_senders,
                              uint32_t *use_counts,
                              void (*track)(void *, uint32_t, const double *), void *track_user);

void remy_oracle_run_sim_traced(const RemyFlatTree *tree, const RemyLeafOverride *ov,
                                const RemyNetConfig *cfg, uint32_t seed,
                                RemySimResult *out, RemySenderResult *out_senders,
                                uint32_t *use_counts,
                                void (*track)(void *, uint32_t, const double *), void *track_user)
{
  remy_oracle_run_sim_kind(tree, ov, cfg, seed, 0, out, out_senders, use_counts, track, track_user);
}

/* kind 0: Evaluator<WhiskerTree>::score (src/evaluator.cc:77-103); kind 1: Evaluator<FinTree>::score
 * (src/evaluator.cc:105-132), where the Fish's own PRNG seed is the first draw of the run's PRNG (:113). */
void remy_oracle_run_sim_kind(const RemyFlatTree *tree, const RemyLeafOverride *ov,
                              const RemyNetConfig *cfg, uint32_t seed, int kind,
                              RemySimResult *out, RemySenderResult *out_senders,
                              uint32_t *use_counts,
                              void (*track)(void *, uint32_t, const double *), void *track_user)
{
  memset(out, 0, sizeof *out);
  if (cfg->num_senders == 0 || cfg->num_senders > 4096 || !(cfg->simulation_ticks > 0)
      || !(cfg->link_ppt > 0)) {
    out->error = REMY_ERR_BAD_CONFIG;
    return;
  }
  net_t N; memset(&N, 0, sizeof N);
  prng_seed(&N.prng, seed);
  N.kind = kind;
  const uint32_t fish_prng_seed = kind == 1 ? prng_next(&N.prng) : 0;
  N.pol.tree = tree;
  N.pol.ov = (ov && ov->leaf_index != REMY_NO_OVERRIDE) ? ov : NULL;
  N.pol.use_counts = use_counts;
  N.pol.track = track; N.pol.track_user = track_user;
  N.n = cfg->num_senders;
  N.gang = (sender_t *)calloc(N.n, sizeof(sender_t));
  N.collector = (mailbox_t *)calloc(N.n, sizeof(mailbox_t));
  N.start_rate = 1.0 / cfg->mean_off_duration;
  N.stop_rate = 1.0 / cfg->mean_on_duration;
  for (uint32_t i = 0; i < N.n; i++) {
    sender_t *s = &N.gang[i];
    s->id = i;
    s->next_switch_tick = exp_sample(&N.prng, N.start_rate);
    s->rat.largest_ack = -1;
    if (kind == 1) prng_seed(&s->fish.prng, fish_prng_seed);
  }
  N.delay = cfg->delay;
  N.service = 1.0 / cfg->link_ppt;
  N.limit = cfg->buffer_size;
  N.loss_rate = cfg->stochastic_loss_rate;

  const double duration = cfg->simulation_ticks;
  uint64_t event_ticks = 0;
  while (N.tickno < duration) {
    /* SenderGang::next_event_time (src/sendergang.cc:334-347) */
    double s_next = DBL_MAX;
    for (uint32_t i = 0; i < N.n; i++) {
      const sender_t *s = &N.gang[i];
      const double ev = dmin(s->next_switch_tick, s->sending ? (kind == 1 ? fish_next_event_time(&s->fish, N.tickno) : rat_next_event_time(&s->rat, N.tickno)) : DBL_MAX);
      if (ev < s_next) s_next = ev;
    }
    const double link_next = N.pending.len ? N.pending.a[N.pending.head].release : DBL_MAX;
    const double loss_next = DBL_MAX; /* the loss buffer is always flushed within the tick */
    const double delay_next = N.delay_q.len ? N.delay_q.a[N.delay_q.head].release : DBL_MAX;
    double rec_next = DBL_MAX; /* Receiver::next_event_time (src/receiver.cc:45-53) */
    for (uint32_t i = 0; i < N.collector_size; i++) if (N.collector[i].len) { rec_next = N.tickno; break; }
    N.tickno = dmin(dmin(s_next, dmin(link_next, loss_next)), dmin(delay_next, rec_next));
    if (N.tickno > duration) break;
    net_tick(&N);
    event_ticks++;
    if (N.err & (REMY_ERR_NO_WHISKER | REMY_ERR_NO_CHILD)) break;
  }

  double total = 0.0;
  for (uint32_t i = 0; i < N.n; i++) {
    const sender_t *s = &N.gang[i];
    total += utility_value(&s->utility);
    out->packets_sent += s->rat.packets_sent;
    out->packets_received += s->utility.packets_received;
    if (out_senders) {
      out_senders[i].tick_share_sending = s->utility.tick_share_sending;
      out_senders[i].total_delay = s->utility.total_delay;
      out_senders[i].packets_received = s->utility.packets_received;
      out_senders[i].packets_sent = s->rat.packets_sent;
    }
  }
  out->utility = total / N.n;
  out->last_tick = N.tickno;
  out->event_ticks = event_ticks;
  out->link_queued = (uint32_t)(N.buffer.len + N.pending.len);
  uint32_t infl = (uint32_t)N.delay_q.len;
  for (uint32_t i = 0; i < N.n; i++) infl += (uint32_t)N.collector[i].len;
  out->delay_inflight = infl;
  out->prng_state = N.prng.x;
  out->error = N.err;

  for (uint32_t i = 0; i < N.n; i++) free(N.collector[i].a);
  free(N.collector); free(N.gang);
  free(N.pending.a); free(N.buffer.a); free(N.loss_buf.a); free(N.delay_q.a);
}

/* ------------------------------------------------------------- batches ---- */
typedef struct {
  const RemyFlatTree *tree; const RemyLeafOverride *cands; uint32_t n_cands;
  const RemyNetConfig *cfgs; const uint32_t *seeds; uint32_t n_cfgs;
  uint32_t sender_stride;
  RemySimResult *out_sims; RemySenderResult *out_senders; uint32_t *out_use_counts;
  uint32_t next; pthread_mutex_t mu;
  int kind;
} batch_t;

static void *batch_worker(void *arg)
{
  batch_t *b = (batch_t *)arg;
  const uint32_t total = b->n_cands * b->n_cfgs;
  uint32_t *local_counts = b->out_use_counts ? (uint32_t *)malloc(sizeof(uint32_t) * b->tree->n_leaves) : NULL;
  RemySenderResult *tmp = (RemySenderResult *)malloc(sizeof(RemySenderResult) * 4096);
  for (;;) {
    pthread_mutex_lock(&b->mu);
    const uint32_t s = b->next++;
    pthread_mutex_unlock(&b->mu);
    if (s >= total) break;
    const uint32_t c = s / b->n_cfgs, k = s % b->n_cfgs;
    if (local_counts) memset(local_counts, 0, sizeof(uint32_t) * b->tree->n_leaves);
    remy_oracle_run_sim_kind(b->tree, b->cands ? &b->cands[c] : NULL, &b->cfgs[k], b->seeds[k], b->kind,
                             &b->out_sims[s], tmp, local_counts, NULL, NULL);
    if (b->out_senders) {
      const uint32_t n = b->cfgs[k].num_senders < b->sender_stride ? b->cfgs[k].num_senders : b->sender_stride;
      RemySenderResult *dst = &b->out_senders[(size_t)s * b->sender_stride];
      memset(dst, 0, sizeof(RemySenderResult) * b->sender_stride);
      if (!(b->out_sims[s].error & REMY_ERR_BAD_CONFIG)) memcpy(dst, tmp, sizeof(RemySenderResult) * n);
    }
    if (local_counts) {
      pthread_mutex_lock(&b->mu);
      for (uint32_t l = 0; l < b->tree->n_leaves; l++)
        b->out_use_counts[(size_t)c * b->tree->n_leaves + l] += local_counts[l];
      pthread_mutex_unlock(&b->mu);
    }
  }
  free(tmp); free(local_counts);
  return NULL;
}

/* Same shapes as remy_b200_score_batch (include/remy_b200.h), run on
 * `n_threads` host threads pulling (candidate, config) items. */
int remy_oracle_score_batch_kind(const RemyFlatTree *tree,
                                 const RemyLeafOverride *cands, uint32_t n_cands,
                                 const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                 uint32_t sender_stride,
                                 RemySimResult *out_sims, RemySenderResult *out_senders,
                                 uint32_t *out_use_counts, uint32_t n_threads, int kind);

int remy_oracle_score_batch(const RemyFlatTree *tree,
                            const RemyLeafOverride *cands, uint32_t n_cands,
                            const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                            uint32_t sender_stride,
                            RemySimResult *out_sims, RemySenderResult *out_senders,
                            uint32_t *out_use_counts, uint32_t n_threads)
{
  return remy_oracle_score_batch_kind(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, out_sims, out_senders,
                                      out_use_counts, n_threads, 0);
}

/* kind 1 = Fish / FinTree: leaf lambda in RemyLeafAction.intersend (and in RemyLeafOverride.intersend) */
int remy_oracle_score_batch_kind(const RemyFlatTree *tree,
                                 const RemyLeafOverride *cands, uint32_t n_cands,
                                 const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                 uint32_t sender_stride,
                                 RemySimResult *out_sims, RemySenderResult *out_senders,
                                 uint32_t *out_use_counts, uint32_t n_threads, int kind)
{
  if (!tree || !cfgs || !seeds || !out_sims) return REMY_EINVAL;
  if (!cands) n_cands = 1;
  batch_t b; memset(&b, 0, sizeof b);
  b.tree = tree; b.cands = cands; b.n_cands = n_cands; b.cfgs = cfgs; b.seeds = seeds; b.n_cfgs = n_cfgs;
  b.sender_stride = sender_stride; b.out_sims = out_sims; b.out_senders = out_senders; b.out_use_counts = out_use_counts;
  b.kind = kind;
  pthread_mutex_init(&b.mu, NULL);
  if (out_use_counts) memset(out_use_counts, 0, sizeof(uint32_t) * (size_t)n_cands * tree->n_leaves);
  if (n_threads <= 1) {
    batch_worker(&b);
  } else {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (uint32_t i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &b);
    for (uint32_t i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
  }
  pthread_mutex_destroy(&b.mu);
  return 0;
}

/* exposed for tests of the arithmetic restatements */
double remy_oracle_exact_log(double x) { return remy_exact_log(x); }
uint32_t remy_oracle_prng_next(uint32_t x) { prng_t g = { x }; return prng_next(&g); }
double remy_oracle_canonical(uint32_t *state) { prng_t g = { *state }; double r = prng_canonical(&g); *state = g.x; return r; }
double remy_oracle_exponential(uint32_t *state, double rate) { prng_t g = { *state }; double r = exp_sample(&g, rate); *state = g.x; return r; }
void remy_oracle_shuffle(uint32_t *state, uint32_t *idx, uint32_t n) { prng_t g = { *state }; shuffle_indices(&g, idx, n); *state = g.x; }

/* ----------------------------------------------------- trace == true ---- */
/* boost::accumulators::accumulator_set<double, stats<tag::median>> as used by
 * MemoryRange::_acc (src/memoryrange.hh:24-26): tag::median is the P^2 running estimator
 * (Jain & Chlamtac 1985) for p = 0.5, NOT an exact median.  Boost is a third-party
 * dependency that is absent from /root/reference and from this image (the reference pins no
 * version, configure.ac:57-59); this restates the published algorithm in the form of
 * boost/accumulators/statistics/p_square_quantile.hpp (1.5x-1.8x: unchanged in that range):
 * first five samples stored and sorted, then per sample: find the cell, shift marker
 * positions, advance desired positions by {0, p/2, p, (1+p)/2, 1}, and move markers 1..3 by
 * +-1 with the piecewise-parabolic formula (linear if it would leave the bracket).
 * "Parity unpinned" for this estimator: no Boost here to run against. */
typedef struct { double h[5], n[5], d[5], inc[5]; uint64_t count; } p2_t;

void remy_oracle_p2_init(p2_t *a)
{
  const double p = 0.5;
  for (int i = 0; i < 5; i++) { a->h[i] = 0; a->n[i] = i + 1.; }
  a->d[0] = 1.; a->d[1] = 1. + 2. * p; a->d[2] = 1. + 4. * p; a->d[3] = 3. + 2. * p; a->d[4] = 5.;
  a->inc[0] = 0.; a->inc[1] = p / 2.; a->inc[2] = p; a->inc[3] = (1. + p) / 2.; a->inc[4] = 1.;
  a->count = 0;
}

void remy_oracle_p2_add(p2_t *a, double x)
{
  const uint64_t cnt = ++a->count;
  if (cnt <= 5) {
    a->h[cnt - 1] = x;
    if (cnt == 5) { /* std::sort of five values */
      for (int i = 1; i < 5; i++) { const double v = a->h[i]; int j = i; while (j > 0 && v < a->h[j - 1]) { a->h[j] = a->h[j - 1]; j--; } a->h[j] = v; }
    }
    return;
  }
  unsigned k;
  if (x < a->h[0]) { a->h[0] = x; k = 1; }
  else if (a->h[4] <= x) { a->h[4] = x; k = 4; }
  else { k = 0; while (k < 5 && !(x < a->h[k])) k++; } /* std::upper_bound */
  for (unsigned i = k; i < 5; i++) a->n[i] += 1.;
  for (unsigned i = 0; i < 5; i++) a->d[i] += a->inc[i];
  for (unsigned i = 1; i <= 3; i++) {
    const double d = a->d[i] - a->n[i];
    const double dp = a->n[i + 1] - a->n[i];
    const double dm = a->n[i - 1] - a->n[i];
    const double hp = (a->h[i + 1] - a->h[i]) / dp;
    const double hm = (a->h[i - 1] - a->h[i]) / dm;
    if ((d >= 1. && dp > 1.) || (d <= -1. && dm < -1.)) {
      const short sign_d = (short)(d / fabs(d));
      const double h = a->h[i] + sign_d / (dp - dm) * ((sign_d - dm) * hp + (dp - sign_d) * hm);
      if (a->h[i - 1] < h && h < a->h[i + 1]) a->h[i] = h;
      else {
        if (d > 0) a->h[i] += hp;
        if (d < 0) a->h[i] -= hm;
      }
      a->n[i] += sign_d;
    }
  }
}

double remy_oracle_p2_median(const p2_t *a) { return a->h[2]; }

/* Raw trace of one simulation: record k = { (double)leaf, mem[0..5] } at out[7k..7k+6], in the
 * order of the use_whisker calls.  Returns the number of calls (records beyond `cap` are
 * counted but not stored). */
typedef struct { double *out; uint64_t cap, n; } trace_log_t;
static void trace_log_cb(void *user, uint32_t leaf, const double *mem)
{
  trace_log_t *t = (trace_log_t *)user;
  if (t->n < t->cap) { double *r = t->out + 7 * t->n; r[0] = (double)leaf; for (int a = 0; a < REMY_NUM_AXES; a++) r[1 + a] = mem[a]; }
  t->n++;
}
uint64_t remy_oracle_trace_sim_kind(const RemyFlatTree *tree, const RemyNetConfig *cfg, uint32_t seed, int kind,
                                    RemySimResult *out_sim, double *out, uint64_t cap)
{
  trace_log_t t = { out, cap, 0 };
  RemySimResult tmp;
  remy_oracle_run_sim_kind(tree, NULL, cfg, seed, kind, out_sim ? out_sim : &tmp, NULL, NULL, trace_log_cb, &t);
  return t.n;
}
uint64_t remy_oracle_trace_sim(const RemyFlatTree *tree, const RemyNetConfig *cfg, uint32_t seed,
                               RemySimResult *out_sim, double *out, uint64_t cap)
{
  return remy_oracle_trace_sim_kind(tree, cfg, seed, 0, out_sim, out, cap);
}

/* Evaluator::score(tree, seed_k, {config_k}, trace = true, ticks) for k = 0..n_cfgs-1 on ONE
 * tree object, i.e. every leaf's accumulators see the lookups of config 0, then config 1, ...
 * (MemoryRange::track feeds _acc[axis] for the leaf's active axes, src/memoryrange.cc:60-66).
 * acc: [n_leaves][REMY_NUM_AXES] estimators, carried in and out (NULL-initialised by the caller
 * with remy_oracle_p2_init). */
typedef struct { const RemyFlatTree *tree; const uint32_t *leaf_axes; p2_t *acc; } trace_acc_t;
static void trace_acc_cb(void *user, uint32_t leaf, const double *mem)
{
  trace_acc_t *t = (trace_acc_t *)user;
  for (int a = 0; a < REMY_NUM_AXES; a++)
    if (t->leaf_axes[leaf] & (1u << a)) remy_oracle_p2_add(&t->acc[(size_t)leaf * REMY_NUM_AXES + a], mem[a]);
}
int remy_oracle_trace_medians_kind(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs, int kind,
                                   RemySimResult *out_sims, uint32_t *out_use_counts, p2_t *acc);
int remy_oracle_trace_medians(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                              RemySimResult *out_sims, uint32_t *out_use_counts, p2_t *acc)
{
  return remy_oracle_trace_medians_kind(tree, cfgs, seeds, n_cfgs, 0, out_sims, out_use_counts, acc);
}
int remy_oracle_trace_medians_kind(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs, int kind,
                                   RemySimResult *out_sims, uint32_t *out_use_counts, p2_t *acc)
{
  uint32_t *leaf_axes = (uint32_t *)calloc(tree->n_leaves, sizeof(uint32_t));
  for (uint32_t i = 0; i < tree->n_nodes; i++)
    if (tree->nodes[i].child_count == 0) leaf_axes[tree->nodes[i].leaf_index] = tree->nodes[i].axis_mask;
  trace_acc_t t = { tree, leaf_axes, acc };
  if (out_use_counts) memset(out_use_counts, 0, sizeof(uint32_t) * tree->n_leaves);
  for (uint32_t k = 0; k < n_cfgs; k++)
    remy_oracle_run_sim_kind(tree, NULL, &cfgs[k], seeds[k], kind, &out_sims[k], NULL, out_use_counts, trace_acc_cb, &t);
  free(leaf_axes);
  return 0;
}
size_t remy_oracle_p2_sizeof(void) { return sizeof(p2_t); }
