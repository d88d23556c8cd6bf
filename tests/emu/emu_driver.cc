/*
 * emu_driver.cc — runs the kernel's run_sim<> (compiled for the host through
 * host_emu.h) over a batch, with the same batch preparation as capi.cu.
 * TEST INFRASTRUCTURE: checks the kernel's restructured loop without a GPU.
 */
#define REMY_HOST_EMU 1
#include "host_emu.h"
#include "../../remy_b200/csrc/sim_kernel.cuh"
#include "../../remy_b200/csrc/batch_prep.h"

#include <vector>
#include <string>

using namespace remy;

static int score_batch_kind(const RemyFlatTree *tree,
                            const RemyLeafOverride *cands, uint32_t n_cands,
                            const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                            uint32_t sender_stride,
                            RemySimResult *out_sims, RemySenderResult *out_senders,
                            uint32_t *out_use_counts, uint32_t link_cap_override, int fish);

extern "C" int remy_emu_score_batch(const RemyFlatTree *tree,
                                    const RemyLeafOverride *cands, uint32_t n_cands,
                                    const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                    uint32_t sender_stride,
                                    RemySimResult *out_sims, RemySenderResult *out_senders,
                                    uint32_t *out_use_counts, uint32_t link_cap_override)
{
  return score_batch_kind(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, out_sims, out_senders, out_use_counts, link_cap_override, 0);
}

/* the Fish / FinTree variant of the kernel (leaf lambda in RemyLeafAction.intersend) */
extern "C" int remy_emu_score_batch_fish(const RemyFlatTree *tree,
                                         const RemyLeafOverride *cands, uint32_t n_cands,
                                         const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                         uint32_t sender_stride,
                                         RemySimResult *out_sims, RemySenderResult *out_senders,
                                         uint32_t *out_use_counts, uint32_t link_cap_override)
{
  return score_batch_kind(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, out_sims, out_senders, out_use_counts, link_cap_override, 1);
}

/* Rat senders behind ByteSwitchedSender (flows of a random length in packets) */
extern "C" int remy_emu_score_batch_flows(const RemyFlatTree *tree,
                                          const RemyLeafOverride *cands, uint32_t n_cands,
                                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                          uint32_t sender_stride,
                                          RemySimResult *out_sims, RemySenderResult *out_senders,
                                          uint32_t *out_use_counts, uint32_t link_cap_override)
{
  return score_batch_kind(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, out_sims, out_senders, out_use_counts, link_cap_override, 2);
}

static int score_batch_kind(const RemyFlatTree *tree,
                            const RemyLeafOverride *cands, uint32_t n_cands,
                            const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                            uint32_t sender_stride,
                            RemySimResult *out_sims, RemySenderResult *out_senders,
                            uint32_t *out_use_counts, uint32_t link_cap_override, int fish)
{
  PreparedTree pt;
  if (!prepare_tree(tree, pt).empty()) return REMY_EINVAL;
  if (!cands) n_cands = 1;
  BatchPlan plan = plan_batch(cfgs, n_cfgs, n_cands);
  if (link_cap_override) plan.link_cap = link_cap_override;
  const uint32_t n_sims = n_cands * n_cfgs;
  std::vector<SenderCold> cold(plan.n_max);
  std::vector<SendState> sendst(plan.n_max);
  std::vector<DelayEnt> dring(plan.delay_cap);
  std::vector<uint32_t> cnt(tree->n_leaves);
  std::vector<double> hot(INC_BUF + plan.n_max), share(plan.n_max), nsw(plan.n_max);
  if (out_use_counts) memset(out_use_counts, 0, sizeof(uint32_t) * (size_t)n_cands * tree->n_leaves);
  uint32_t cap = plan.link_cap;
  std::vector<uint32_t> todo = plan.order;
  for (int round = 0; round < 8 && !todo.empty(); round++) { /* rings grow 16x per round up to 2^24 entries, like the library's re-runs */
    std::vector<LinkEnt> lring(cap);
    KernelArgs A;
    memset(&A, 0, sizeof A);
    A.bounds = pt.bounds.data(); A.meta = pt.meta.data(); A.leaves = pt.leaves.data();
    A.cuts = pt.cuts.data(); A.table = pt.table.data(); A.n_cuts = (uint32_t)pt.cuts.size(); A.n_table = (uint32_t)pt.table.size();
    A.n_nodes = tree->n_nodes; A.n_leaves = tree->n_leaves; A.axes_mask = pt.axes_mask; A.tree_in_smem = 0;
    A.cands = cands; A.cfgs = cfgs; A.seeds = seeds; A.n_cfgs = n_cfgs;
    A.n_max = plan.n_max; A.link_cap = cap; A.delay_cap = plan.delay_cap;
    A.out_sims = out_sims; A.out_senders = out_senders; A.sender_stride = sender_stride; A.out_use_counts = out_use_counts;
    TreeView T; T.bounds = A.bounds; T.meta = A.meta; T.leaves = A.leaves; T.cuts = A.cuts; T.table = A.table; T.axes_mask = A.axes_mask;
    std::vector<uint32_t> redo;
    for (uint32_t sim : todo) {
      if (fish == 2) {
        if (out_use_counts) run_sim<true, false, false, true>(A, T, sim, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), cnt.data());
        else run_sim<false, false, false, true>(A, T, sim, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), nullptr);
      }
      else if (fish) {
        if (out_use_counts) run_sim<true, false, true>(A, T, sim, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), cnt.data());
        else run_sim<false, false, true>(A, T, sim, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), nullptr);
      }
      else if (out_use_counts) run_sim<true>(A, T, sim, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), cnt.data());
      else run_sim<false>(A, T, sim, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), nullptr);
      if (out_sims[sim].error & REMY_ERR_LINK_OVERFLOW) redo.push_back(sim);
    }
    todo.swap(redo);
    if (cap >= (1u << 24)) break;
    cap = (uint32_t)std::min<uint64_t>((uint64_t)cap * 16, 1u << 24);
  }
  return 0;
}

/* trace == true on the CPU build of the kernel: per config k, out_lookups[k] records of
 * trace_words = 1 + popcount(axes) words each, written back to back into `log` (capacity
 * cap_records records; returns REMY_ENOMEM if it does not fit).  Two runs per simulation, like
 * the library: a counting run, then the run that writes. */
static int trace_kind(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                      RemySimResult *out_sims, uint32_t *out_use_counts,
                      uint64_t *log, uint64_t cap_records, uint64_t *out_lookups, uint32_t *out_words, int fish);
extern "C" int remy_emu_trace(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                              RemySimResult *out_sims, uint32_t *out_use_counts,
                              uint64_t *log, uint64_t cap_records, uint64_t *out_lookups, uint32_t *out_words)
{ return trace_kind(tree, cfgs, seeds, n_cfgs, out_sims, out_use_counts, log, cap_records, out_lookups, out_words, 0); }
extern "C" int remy_emu_trace_fish(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                   RemySimResult *out_sims, uint32_t *out_use_counts,
                                   uint64_t *log, uint64_t cap_records, uint64_t *out_lookups, uint32_t *out_words)
{ return trace_kind(tree, cfgs, seeds, n_cfgs, out_sims, out_use_counts, log, cap_records, out_lookups, out_words, 1); }
static int trace_kind(const RemyFlatTree *tree, const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                      RemySimResult *out_sims, uint32_t *out_use_counts,
                      uint64_t *log, uint64_t cap_records, uint64_t *out_lookups, uint32_t *out_words, int fish)
{
  PreparedTree pt;
  if (!prepare_tree(tree, pt).empty()) return REMY_EINVAL;
  BatchPlan plan = plan_batch(cfgs, n_cfgs, 1);
  std::vector<SenderCold> cold(plan.n_max);
  std::vector<SendState> sendst(plan.n_max);
  std::vector<DelayEnt> dring(plan.delay_cap);
  std::vector<uint32_t> cnt(tree->n_leaves);
  std::vector<double> hot(INC_BUF + plan.n_max), share(plan.n_max), nsw(plan.n_max);
  std::vector<uint64_t> offs(n_cfgs, 0);
  std::vector<uint32_t> scratch_counts(tree->n_leaves);
  if (out_use_counts) memset(out_use_counts, 0, sizeof(uint32_t) * tree->n_leaves);
  const uint32_t words = 1 + (uint32_t)__builtin_popcount(pt.axes_mask);
  if (out_words) *out_words = words;
  uint64_t total = 0;
  for (uint32_t k = 0; k < n_cfgs; k++) {
    for (int phase = 0; phase < 2; phase++) {
      uint32_t cap = plan.link_cap;
      for (;;) {
        std::vector<LinkEnt> lring(cap);
        KernelArgs A;
        memset(&A, 0, sizeof A);
        A.bounds = pt.bounds.data(); A.meta = pt.meta.data(); A.leaves = pt.leaves.data();
        A.cuts = pt.cuts.data(); A.table = pt.table.data(); A.n_cuts = (uint32_t)pt.cuts.size(); A.n_table = (uint32_t)pt.table.size();
        A.n_nodes = tree->n_nodes; A.n_leaves = tree->n_leaves; A.axes_mask = pt.axes_mask; A.tree_in_smem = 0;
        A.cands = nullptr; A.cfgs = cfgs; A.seeds = seeds; A.n_cfgs = n_cfgs;
        A.n_max = plan.n_max; A.link_cap = cap; A.delay_cap = plan.delay_cap;
        A.out_sims = out_sims; A.out_senders = nullptr; A.sender_stride = 0;
        A.out_use_counts = phase == 1 ? out_use_counts : scratch_counts.data();
        A.trace_log = phase == 1 ? log : nullptr; A.trace_off = offs.data(); A.out_lookups = out_lookups; A.trace_words = words;
        TreeView T; T.bounds = A.bounds; T.meta = A.meta; T.leaves = A.leaves; T.cuts = A.cuts; T.table = A.table; T.axes_mask = A.axes_mask;
        if (fish) run_sim<true, true, true>(A, T, k, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), cnt.data());
        else run_sim<true, true>(A, T, k, TickMem{hot.data(), 1}, share.data(), nsw.data(), sendst.data(), cold.data(), lring.data(), dring.data(), cnt.data());
        if (!(out_sims[k].error & REMY_ERR_LINK_OVERFLOW) || cap >= (1u << 24)) break;
        cap *= 16;
      }
      if (phase == 0) {
        offs[k] = total;
        if (total + out_lookups[k] > cap_records) return REMY_ENOMEM;
      }
    }
    total += out_lookups[k];
  }
  return 0;
}

/* div_small_int (the kernel's division-free dt / k) against the hardware division: n values of dt per
 * k = 1..32 from a xorshift stream - uniform bit patterns over the positive normal doubles, values
 * shaped like the simulator's (differences of event times) and neighbours of powers of two.  Returns
 * the number of mismatches. */
extern "C" uint64_t remy_emu_check_small_div(uint64_t n, uint64_t seed)
{
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1, bad = 0;
  auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (uint32_t k = 1; k <= 32; k++) {
    const double kd = (double)k, inv = 1.0 / kd;
    for (uint64_t i = 0; i < n; i++) {
      const uint64_t r = next();
      double dt;
      switch (i & 3u) {
        case 0: { uint64_t b = (r & 0x7fffffffffffffffull); if ((b >> 52) == 0x7ff || (b >> 52) < 64) b = (b & 0xfffffffffffffull) | (0x3ffull << 52); memcpy(&dt, &b, 8); break; }
        case 1: dt = (double)(r >> 11) * 0x1p-53 * 1000.0; break;                       /* [0, 1000) ms */
        case 2: dt = (double)((r >> 40) + 1) / (double)(((r >> 8) & 0xffff) + 1); break; /* ratios of small integers */
        default: { uint64_t b = ((uint64_t)(1023 + (int)(r % 40) - 20) << 52); b += (int64_t)((r >> 8) & 7) - 3; memcpy(&dt, &b, 8); break; } /* 2^e +- a few ulp */
      }
      if (remy::div_small_int(dt, kd, inv) != dt / kd) bad++;
    }
  }
  return bad;
}

/* shuffle_skip (the kernel's advance of the PRNG by std::shuffle's draws without building the permutation)
 * against the draw-by-draw uniform_int loop: n_random states per sender count, plus states constructed so
 * that one draw lands among the values uniform_int can reject or among the small residues (where the lazily
 * reduced fast paths must notice that they cannot be trusted).  Returns the number of mismatches. */
extern "C" uint64_t remy_emu_check_shuffle_skip(uint64_t n_random, uint64_t seed)
{
  using namespace remy;
  auto exact = [](uint32_t x, uint32_t n) {
    uint32_t i = 1;
    if ((n & 1u) == 0) { (void)uniform_int(x, 1); i = 2; }
    for (; i < n; i += 2) (void)uniform_int(x, (i + 1) * (i + 2) - 1);
    return x;
  };
  auto modpow = [](uint64_t b, uint64_t e) { uint64_t r = 1; const uint64_t m = 2147483647ull; b %= m; while (e) { if (e & 1) r = r * b % m; b = b * b % m; e >>= 1; } return r; };
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1, bad = 0;
  auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (uint32_t n = 1; n <= REMY_MAX_SENDERS; n++) {
    for (uint64_t k = 0; k < n_random; k++) {
      const uint32_t x = (uint32_t)(next() % 2147483646ull) + 1;
      if (shuffle_skip(x, n) != exact(x, n)) bad++;
    }
    for (uint32_t j = 1; j <= (n >> 1); j++) {
      const uint64_t inv = modpow(modpow(16807, j), 2147483647ull - 2); /* 16807^-j mod m */
      for (uint32_t v = 1; v <= 1300; v++) { /* draw j = m - v */
        const uint32_t x = (uint32_t)((uint64_t)(2147483647u - v) * inv % 2147483647ull);
        if (x && shuffle_skip(x, n) != exact(x, n)) bad++;
      }
      for (uint32_t v = 1; v <= 40000; v += 7) { /* draw j = v */
        const uint32_t x = (uint32_t)((uint64_t)v * inv % 2147483647ull);
        if (x && shuffle_skip(x, n) != exact(x, n)) bad++;
      }
    }
  }
  return bad;
}
