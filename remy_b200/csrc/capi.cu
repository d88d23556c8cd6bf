/*
 * capi.cu — implementation of include/remy_b200.h on top of the sm_100a kernel.
 *
 * Host side of the boundary only: argument checking, flattening the tree into
 * the device layout (with the per-node cell tables), the config-sorted work
 * list (one launch per batch; optionally a launch per sender-count bucket), scratch sizing, kernel launches on
 * the library's streams, and the bounded re-run of simulations whose link
 * queue outgrew the ring they were given.  There is no CPU implementation of
 * the simulator in this library: without a CUDA device every entry point
 * fails with REMY_ENODEV.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "sim_kernel.cuh"
#include "batch_prep.h"

using namespace remy;

/* Builds the list of simulations whose link queue outgrew the ring they were given. */
__global__ void remy_collect_overflow(const RemySimResult *sims, uint32_t n_sims, uint32_t *list, uint32_t *count)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_sims && (sims[i].error & REMY_ERR_LINK_OVERFLOW)) list[atomicAdd(count, 1u)] = i;
}

namespace {

thread_local std::string g_err;
std::mutex g_mu;           /* the library is callable from several host threads; calls serialise */
int g_device = -1;
int g_sms = 0;
constexpr int MAX_PASSES = 4;
cudaStream_t g_stream = nullptr;             /* main stream: copies, events, re-runs */
cudaStream_t g_pass_stream[MAX_PASSES] = {}; /* one per concurrent pass */

int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(REMY_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)

template <class T> struct DevBuf {
  T *p = nullptr; size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept { if (this != &o) { free(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  cudaError_t alloc(size_t count) { free(); n = count; if (!count) return cudaSuccess; return cudaMalloc(&p, count * sizeof(T)); }
  void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { free(); }
};

/* Per-slot working memory of one pass (slot = one resident thread). */
struct Scratch {
  DevBuf<SenderCold> cold; DevBuf<double> share; DevBuf<double> nsw; DevBuf<SendState> sendst; DevBuf<LinkEnt> link; DevBuf<DelayEnt> delay; DevBuf<uint32_t> counts;
  uint32_t slots = 0, n_max = 0, link_cap = 0, delay_cap = 0, count_leaves = 0;
  Scratch() {}
  Scratch(Scratch &&) = default;
  Scratch &operator=(Scratch &&) = default;
  static size_t bytes_per_slot(uint32_t n_max, uint32_t link_cap, uint32_t delay_cap, uint32_t count_leaves)
  {
    return (size_t)n_max * (sizeof(SenderCold) + 2 * sizeof(double) + sizeof(SendState)) + (size_t)link_cap * sizeof(LinkEnt)
           + (size_t)delay_cap * sizeof(DelayEnt) + (size_t)count_leaves * sizeof(uint32_t);
  }
};

struct Pass {
  uint32_t n_max = 0, begin = 0, n_work = 0, grid = 0;
  size_t smem = 0; bool tree_in_smem = false;
  Scratch scratch;
  DevBuf<uint32_t> counter;
  cudaEvent_t done = nullptr;
};

} // namespace

struct RemyBatch {
  uint32_t n_cands = 1, n_cfgs = 0, n_sims = 0, n_leaves = 0, n_nodes = 0;
  uint32_t sender_stride = 0; bool want_senders = false, want_counts = false, has_cands = false;
  uint32_t axes_mask = 0; size_t tree_bytes = 0;
  uint32_t delay_cap = 8;
  /* device */
  DevBuf<Bounds> bounds; DevBuf<DevNodeMeta> meta; DevBuf<DevLeaf> leaves; DevBuf<double> cuts; DevBuf<uint16_t> table;
  DevBuf<RemyLeafOverride> d_cands; DevBuf<RemyNetConfig> d_cfgs; DevBuf<uint32_t> d_seeds;
  DevBuf<uint32_t> order; DevBuf<uint32_t> redo_list; DevBuf<uint32_t> redo_count; DevBuf<uint32_t> redo_counter;
  DevBuf<RemySimResult> out_sims; DevBuf<RemySenderResult> out_senders; DevBuf<uint32_t> out_counts;
  std::vector<std::unique_ptr<Pass>> passes;
  uint32_t launches = 0;
  /* trace == true scoring (remy_b200_score_trace): 0 = off, 1 = counting run, 2 = writing run */
  int trace = 0; uint32_t trace_words = 0;
  bool fish = false; /* Fish senders over a FinTree (remy_b200_score_batch_fish) */
  DevBuf<uint64_t> trace_log, trace_off, out_lookups;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ran = false;
  ~RemyBatch()
  {
    for (auto &p : passes) if (p->done) cudaEventDestroy(p->done);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

namespace {

size_t smem_bytes(const RemyBatch *b, uint32_t n_max, bool with_tree)
{
  return (size_t)(INC_BUF + n_max) * SIM_BLOCK * sizeof(double) + (with_tree ? b->tree_bytes : 0);
}

const void *kernel_fn(const RemyBatch *b)
{
  if (b->trace) return b->fish ? (const void *)remy_sim_kernel<true, true, true> : (const void *)remy_sim_kernel<true, true>;
  if (b->fish) return b->want_counts ? (const void *)remy_sim_kernel<true, false, true> : (const void *)remy_sim_kernel<false, false, true>;
  return b->want_counts ? (const void *)remy_sim_kernel<true> : (const void *)remy_sim_kernel<false>;
}

KernelArgs make_args(RemyBatch *b, const Scratch &s, bool tree_in_smem, const uint32_t *order, uint32_t n_work, uint32_t *counter)
{
  KernelArgs a;
  a.bounds = b->bounds.p; a.meta = b->meta.p; a.leaves = b->leaves.p;
  a.cuts = b->cuts.p; a.table = b->table.p; a.n_cuts = (uint32_t)b->cuts.n; a.n_table = (uint32_t)b->table.n;
  a.n_nodes = b->n_nodes; a.n_leaves = b->n_leaves; a.axes_mask = b->axes_mask; a.tree_in_smem = tree_in_smem;
  a.cands = b->has_cands ? b->d_cands.p : nullptr; a.cfgs = b->d_cfgs.p; a.seeds = b->d_seeds.p;
  a.order = order; a.n_cfgs = b->n_cfgs; a.n_work = n_work; a.work_counter = counter;
  a.cold = s.cold.p; a.share = s.share.p; a.nsw = s.nsw.p; a.sendst = s.sendst.p; a.n_max = s.n_max; a.link_ring = s.link.p; a.link_cap = s.link_cap;
  a.delay_ring = s.delay.p; a.delay_cap = s.delay_cap; a.slot_counts = s.counts.p;
  a.out_sims = b->out_sims.p; a.out_senders = b->want_senders ? b->out_senders.p : nullptr;
  a.sender_stride = b->sender_stride; a.out_use_counts = b->want_counts ? b->out_counts.p : nullptr;
  a.trace_log = b->trace == 2 ? b->trace_log.p : nullptr; a.trace_off = b->trace == 2 ? b->trace_off.p : nullptr;
  a.out_lookups = b->trace ? b->out_lookups.p : nullptr; a.trace_words = b->trace_words;
  return a;
}

/* Working memory of destroyed batches is kept for the next batch of the same geometry (an
 * Evaluator scores batch after batch of identical shape; cudaMalloc/cudaFree of tens of GB
 * would otherwise sit inside every one-shot call). */
std::vector<Scratch> g_scratch_pool;
constexpr size_t SCRATCH_POOL_MAX = 4; /* one batch's worth of passes */

void recycle_scratch(Scratch &&s)
{
  if (!s.cold.p) return;
  if (g_scratch_pool.size() >= SCRATCH_POOL_MAX) g_scratch_pool.erase(g_scratch_pool.begin());
  g_scratch_pool.push_back(std::move(s));
}

int alloc_scratch(RemyBatch *b, Scratch &s, uint32_t slots, uint32_t n_max, uint32_t link_cap, uint32_t delay_cap)
{
  const uint32_t count_leaves = b->want_counts ? b->n_leaves : 0;
  for (size_t i = 0; i < g_scratch_pool.size(); i++) {
    Scratch &c = g_scratch_pool[i];
    if (c.slots == slots && c.n_max == n_max && c.link_cap == link_cap && c.delay_cap == delay_cap && c.count_leaves == count_leaves) {
      s = std::move(c);
      g_scratch_pool.erase(g_scratch_pool.begin() + i);
      return 0;
    }
  }
  s.slots = slots; s.n_max = n_max; s.link_cap = link_cap; s.delay_cap = delay_cap; s.count_leaves = count_leaves;
  for (int attempt = 0; attempt < 2; attempt++) {
    cudaError_t e = s.cold.alloc((size_t)slots * n_max);
    if (e == cudaSuccess) e = s.share.alloc((size_t)slots * n_max);
    if (e == cudaSuccess) e = s.nsw.alloc((size_t)slots * n_max);
    if (e == cudaSuccess) e = s.sendst.alloc((size_t)slots * n_max);
    if (e == cudaSuccess) e = s.link.alloc((size_t)slots * link_cap);
    if (e == cudaSuccess) e = s.delay.alloc((size_t)slots * delay_cap);
    if (e == cudaSuccess && count_leaves) e = s.counts.alloc((size_t)slots * count_leaves);
    if (e == cudaSuccess) return 0;
    cudaGetLastError();
    s.cold.free(); s.share.free(); s.nsw.free(); s.sendst.free(); s.link.free(); s.delay.free(); s.counts.free();
    if (attempt == 0 && !g_scratch_pool.empty()) { g_scratch_pool.clear(); continue; } /* give pooled memory back and retry */
    return fail(REMY_ENOMEM, std::string("scratch allocation failed: ") + cudaGetErrorString(e));
  }
  return fail(REMY_ENOMEM, "scratch allocation failed");
}

/* Shared memory and residency of the kernel for a pass that accepts up to n_max senders. */
int pass_geometry(RemyBatch *b, uint32_t n_max, size_t &smem, bool &tree_in_smem, uint32_t &resident_blocks)
{
  tree_in_smem = smem_bytes(b, n_max, true) <= 160 * 1024 && b->tree_bytes <= 64 * 1024;
  smem = smem_bytes(b, n_max, tree_in_smem);
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel_fn(b), SIM_BLOCK, smem));
  if (per_sm < 1) return fail(REMY_ECUDA, "kernel does not fit on an SM");
  /* tuning knob: fewer resident blocks shrink the working set (more of it stays in L2) at the
   * price of fewer warps to hide latency with */
  if (const char *cap = getenv("REMY_B200_MAX_BLOCKS_PER_SM")) { const int c = atoi(cap); if (c >= 1 && c < per_sm) per_sm = c; }
  resident_blocks = (uint32_t)per_sm * (uint32_t)g_sms;
  return 0;
}

int launch(RemyBatch *b, const Scratch &s, size_t smem, bool tree_in_smem, const uint32_t *order, uint32_t n_work,
           uint32_t grid, uint32_t *counter, cudaStream_t stream)
{
  CU(cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream));
  const KernelArgs a = make_args(b, s, tree_in_smem, order, n_work, counter);
  if (b->trace && b->fish) remy_sim_kernel<true, true, true><<<grid, SIM_BLOCK, smem, stream>>>(a);
  else if (b->trace) remy_sim_kernel<true, true><<<grid, SIM_BLOCK, smem, stream>>>(a);
  else if (b->fish && b->want_counts) remy_sim_kernel<true, false, true><<<grid, SIM_BLOCK, smem, stream>>>(a);
  else if (b->fish) remy_sim_kernel<false, false, true><<<grid, SIM_BLOCK, smem, stream>>>(a);
  else if (b->want_counts) remy_sim_kernel<true><<<grid, SIM_BLOCK, smem, stream>>>(a);
  else remy_sim_kernel<false><<<grid, SIM_BLOCK, smem, stream>>>(a);
  CU(cudaGetLastError());
  b->launches++;
  return 0;
}

} // namespace

extern "C" {

uint32_t remy_b200_abi_version(void) { return REMY_B200_ABI_VERSION; }
const char *remy_b200_last_error(void) { return g_err.c_str(); }

int remy_b200_init(int device)
{
  std::lock_guard<std::mutex> lk(g_mu);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(REMY_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(REMY_EINVAL, "device index out of range");
  if (g_device == device && g_stream) return 0;
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(REMY_ENODEV, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "); this library is built for sm_100a only");
  if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  CU(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  for (int i = 0; i < MAX_PASSES; i++) {
    if (g_pass_stream[i]) { cudaStreamDestroy(g_pass_stream[i]); g_pass_stream[i] = nullptr; }
    CU(cudaStreamCreateWithFlags(&g_pass_stream[i], cudaStreamNonBlocking));
  }
  /* the kernel stages the tree in shared memory: allow the large carve-out */
  CU(cudaFuncSetAttribute((const void *)remy_sim_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU(cudaFuncSetAttribute((const void *)remy_sim_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU(cudaFuncSetAttribute((const void *)remy_sim_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU(cudaFuncSetAttribute((const void *)remy_sim_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU(cudaFuncSetAttribute((const void *)remy_sim_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU(cudaFuncSetAttribute((const void *)remy_sim_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  g_device = device; g_sms = prop.multiProcessorCount;
  g_err.clear();
  return 0;
}

void remy_b200_shutdown(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  g_scratch_pool.clear();
  if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  for (int i = 0; i < MAX_PASSES; i++)
    if (g_pass_stream[i]) { cudaStreamDestroy(g_pass_stream[i]); g_pass_stream[i] = nullptr; }
  g_device = -1;
}

static int create_impl(const RemyFlatTree *tree,
                       const RemyLeafOverride *cands, uint32_t n_cands,
                       const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                       uint32_t sender_stride, int want_senders, int want_use_counts, int trace,
                       RemyBatch **out_batch, bool fish = false)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out_batch) return fail(REMY_EINVAL, "out_batch is NULL");
  *out_batch = nullptr;
  if (g_device < 0 || !g_stream) return fail(REMY_ENODEV, "remy_b200_init has not been called");
  if (!tree || !tree->nodes || !tree->leaves || tree->n_nodes == 0 || tree->n_leaves == 0)
    return fail(REMY_EINVAL, "empty tree");
  if (n_cfgs && (!cfgs || !seeds)) return fail(REMY_EINVAL, "cfgs/seeds is NULL");
  if (!cands) n_cands = 1;
  if ((uint64_t)n_cands * n_cfgs > 0x7fffffffull) return fail(REMY_EINVAL, "too many simulations in one batch");
  CU(cudaSetDevice(g_device));

  std::unique_ptr<RemyBatch> guard(new RemyBatch());
  RemyBatch *b = guard.get();
  b->n_cands = n_cands; b->n_cfgs = n_cfgs; b->n_sims = n_cands * n_cfgs;
  b->has_cands = cands != nullptr;
  b->want_senders = want_senders != 0; b->want_counts = want_use_counts != 0 || trace != 0;
  b->sender_stride = sender_stride;
  b->trace = trace;
  b->fish = fish;

  /* ---- tree -> device layout (validated so the kernel cannot walk off the arrays) ---- */
  PreparedTree pt;
  const std::string terr = prepare_tree(tree, pt);
  if (!terr.empty()) return fail(REMY_EINVAL, terr);
  b->n_nodes = tree->n_nodes; b->n_leaves = tree->n_leaves; b->axes_mask = pt.axes_mask; b->tree_bytes = pt.bytes();
  b->trace_words = 1u + (uint32_t)__builtin_popcount(pt.axes_mask);
  if (trace) CU(b->out_lookups.alloc(b->n_sims));
  if (cands)
    for (uint32_t c = 0; c < n_cands; c++)
      if (cands[c].leaf_index != REMY_NO_OVERRIDE && cands[c].leaf_index >= tree->n_leaves)
        return fail(REMY_EINVAL, "candidate " + std::to_string(c) + " overrides a leaf that does not exist");
  CU(b->bounds.alloc(pt.bounds.size())); CU(b->meta.alloc(pt.meta.size())); CU(b->leaves.alloc(pt.leaves.size()));
  CU(b->cuts.alloc(pt.cuts.size())); CU(b->table.alloc(pt.table.size()));
  CU(cudaMemcpyAsync(b->bounds.p, pt.bounds.data(), pt.bounds.size() * sizeof(Bounds), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(b->meta.p, pt.meta.data(), pt.meta.size() * sizeof(DevNodeMeta), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(b->leaves.p, pt.leaves.data(), pt.leaves.size() * sizeof(DevLeaf), cudaMemcpyHostToDevice, g_stream));
  if (!pt.cuts.empty()) CU(cudaMemcpyAsync(b->cuts.p, pt.cuts.data(), pt.cuts.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  if (!pt.table.empty()) CU(cudaMemcpyAsync(b->table.p, pt.table.data(), pt.table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, g_stream));

  /* ---- configs: ring sizes, config-sorted order, launches ("passes": one by default) ---- */
  /* One launch over one cost-sorted list (largest sender counts first).  Cutting the batch into a
   * launch per sender-count bucket (REMY_B200_SINGLE_PASS=0) buys nothing - residency is bound by
   * registers, not by the per-sender shared memory - and costs a drain per launch: a block leaves
   * only when its slowest simulation has ended, and single simulations run for seconds. */
  const char *sp = getenv("REMY_B200_SINGLE_PASS");
  const BatchPlan plan = plan_batch(cfgs, n_cfgs, n_cands, !(sp && atoi(sp) == 0));
  if (b->want_senders && sender_stride < plan.n_max) return fail(REMY_EINVAL, "sender_stride smaller than the largest num_senders");
  b->delay_cap = plan.delay_cap;
  if (plan.passes.size() > MAX_PASSES) return fail(REMY_EINVAL, "internal: too many passes");

  CU(b->d_cfgs.alloc(n_cfgs)); CU(b->d_seeds.alloc(n_cfgs)); CU(b->order.alloc(b->n_sims));
  CU(b->redo_list.alloc(b->n_sims)); CU(b->redo_count.alloc(1)); CU(b->redo_counter.alloc(1));
  if (n_cfgs) {
    CU(cudaMemcpyAsync(b->d_cfgs.p, cfgs, n_cfgs * sizeof(RemyNetConfig), cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync(b->d_seeds.p, seeds, n_cfgs * sizeof(uint32_t), cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync(b->order.p, plan.order.data(), plan.order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, g_stream));
  }
  if (cands) {
    CU(b->d_cands.alloc(n_cands));
    CU(cudaMemcpyAsync(b->d_cands.p, cands, n_cands * sizeof(RemyLeafOverride), cudaMemcpyHostToDevice, g_stream));
  }
  CU(b->out_sims.alloc(b->n_sims));
  if (b->want_senders) CU(b->out_senders.alloc((size_t)b->n_sims * sender_stride));
  if (b->want_counts) CU(b->out_counts.alloc((size_t)n_cands * b->n_leaves));

  /* ---- launch geometry: persistent blocks, as many as are resident, per pass ---- */
  size_t total_bytes = 0;
  for (const PassPlan &pp : plan.passes) {
    std::unique_ptr<Pass> ps(new Pass());
    ps->n_max = pp.n_max; ps->begin = pp.begin; ps->n_work = pp.end - pp.begin;
    uint32_t resident = 0;
    int rc = pass_geometry(b, pp.n_max, ps->smem, ps->tree_in_smem, resident);
    if (rc) return rc;
    ps->grid = std::max(1u, std::min(resident, (ps->n_work + SIM_BLOCK - 1) / SIM_BLOCK));
    total_bytes += (size_t)ps->grid * SIM_BLOCK * Scratch::bytes_per_slot(pp.n_max, plan.link_cap, plan.delay_cap, b->want_counts ? b->n_leaves : 0);
    b->passes.push_back(std::move(ps));
  }
  /* keep the working memory within a budget: shrink the first-try link ring if needed
   * (overflowing simulations are re-run with larger rings) */
  uint32_t link_cap = plan.link_cap;
  /* tuning knob: first-try ring size for buffers larger than it (overflowing simulations are re-run) */
  if (const char *lc = getenv("REMY_B200_LINK_CAP")) { const long c = atol(lc); if (c >= 16 && (uint32_t)c < link_cap) link_cap = next_pow2((uint64_t)c); }
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  const size_t budget = std::min<size_t>(free_b / 2, (size_t)64 << 30);
  while (total_bytes > budget && link_cap > 1024) {
    link_cap /= 2;
    total_bytes = 0;
    for (auto &ps : b->passes)
      total_bytes += (size_t)ps->grid * SIM_BLOCK * Scratch::bytes_per_slot(ps->n_max, link_cap, plan.delay_cap, b->want_counts ? b->n_leaves : 0);
  }
  for (auto &ps : b->passes) {
    int rc = alloc_scratch(b, ps->scratch, ps->grid * SIM_BLOCK, ps->n_max, link_cap, plan.delay_cap);
    if (rc) return rc;
    CU(ps->counter.alloc(1));
    CU(cudaEventCreateWithFlags(&ps->done, cudaEventDisableTiming));
  }
  CU(cudaEventCreate(&b->ev0)); CU(cudaEventCreate(&b->ev1));
  CU(cudaStreamSynchronize(g_stream));
  *out_batch = guard.release();
  return 0;
}

int remy_b200_batch_create(const RemyFlatTree *tree,
                           const RemyLeafOverride *cands, uint32_t n_cands,
                           const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                           uint32_t sender_stride, int want_senders, int want_use_counts,
                           RemyBatch **out_batch)
{
  return create_impl(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, want_senders, want_use_counts, 0, out_batch);
}

int remy_b200_batch_run(RemyBatch *b, float *kernel_ms)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!b) return fail(REMY_EINVAL, "batch is NULL");
  CU(cudaSetDevice(g_device));
  b->launches = 0;
  if (b->want_counts) CU(cudaMemsetAsync(b->out_counts.p, 0, b->out_counts.n * sizeof(uint32_t), g_stream));
  CU(cudaEventRecord(b->ev0, g_stream));
  if (b->n_sims) {
    /* (with REMY_B200_SINGLE_PASS=0 there are several:) the passes run concurrently on their own streams, largest sender counts first; as a
     * pass drains, the next one's blocks take over the freed SM resources */
    for (size_t i = 0; i < b->passes.size(); i++) {
      Pass &ps = *b->passes[i];
      cudaStream_t st = g_pass_stream[i];
      CU(cudaStreamWaitEvent(st, b->ev0, 0));
      int rc = launch(b, ps.scratch, ps.smem, ps.tree_in_smem, b->order.p + ps.begin, ps.n_work, ps.grid, ps.counter.p, st);
      if (rc) return rc;
      CU(cudaEventRecord(ps.done, st));
      CU(cudaStreamWaitEvent(g_stream, ps.done, 0));
    }
    /* Simulations whose link queue outgrew the ring are re-run with larger rings, inside
     * the timed region.  (The reference's deque is unbounded; a finite ring is this
     * library's artefact, so it is hidden from the caller as long as memory allows.) */
    uint32_t cap = b->passes.empty() ? 1 : b->passes[0]->scratch.link_cap;
    for (int round = 0; round < 4 && cap < (1u << 24); round++) {
      CU(cudaMemsetAsync(b->redo_count.p, 0, sizeof(uint32_t), g_stream));
      remy_collect_overflow<<<(b->n_sims + 255) / 256, 256, 0, g_stream>>>(b->out_sims.p, b->n_sims, b->redo_list.p, b->redo_count.p);
      CU(cudaGetLastError());
      b->launches++;
      uint32_t n_redo = 0;
      CU(cudaMemcpyAsync(&n_redo, b->redo_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, g_stream));
      CU(cudaStreamSynchronize(g_stream));
      if (n_redo == 0) break;
      cap = (uint32_t)std::min<uint64_t>((uint64_t)cap * 16, 1u << 24);
      const size_t per_slot = Scratch::bytes_per_slot(REMY_MAX_SENDERS, cap, b->delay_cap, b->want_counts ? b->n_leaves : 0);
      uint32_t slots = (uint32_t)std::min<size_t>(std::max<size_t>(((size_t)32 << 30) / per_slot, SIM_BLOCK), n_redo);
      slots = ((slots + SIM_BLOCK - 1) / SIM_BLOCK) * SIM_BLOCK;
      Scratch s2;
      int rc = alloc_scratch(b, s2, slots, REMY_MAX_SENDERS, cap, b->delay_cap);
      if (rc) return rc;
      size_t smem = 0; bool tis = false; uint32_t resident = 0;
      rc = pass_geometry(b, REMY_MAX_SENDERS, smem, tis, resident);
      if (rc) return rc;
      rc = launch(b, s2, smem, tis, b->redo_list.p, n_redo, slots / SIM_BLOCK, b->redo_counter.p, g_stream);
      if (rc) return rc;
      CU(cudaStreamSynchronize(g_stream)); /* s2 is freed at the end of this scope */
    }
  }
  CU(cudaEventRecord(b->ev1, g_stream));
  if (kernel_ms) {
    CU(cudaEventSynchronize(b->ev1));
    CU(cudaEventElapsedTime(kernel_ms, b->ev0, b->ev1));
  }
  b->ran = true;
  return 0;
}

int remy_b200_batch_fetch(RemyBatch *b, RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!b || (!out_sims && b->n_sims)) return fail(REMY_EINVAL, "batch/out_sims is NULL");
  if (!b->ran) return fail(REMY_EINVAL, "batch has not been run");
  if (out_senders && !b->want_senders) return fail(REMY_EINVAL, "batch was created without want_senders");
  if (out_use_counts && !b->want_counts) return fail(REMY_EINVAL, "batch was created without want_use_counts");
  CU(cudaSetDevice(g_device));
  if (b->n_sims)
    CU(cudaMemcpyAsync(out_sims, b->out_sims.p, (size_t)b->n_sims * sizeof(RemySimResult), cudaMemcpyDeviceToHost, g_stream));
  if (out_senders && b->n_sims)
    CU(cudaMemcpyAsync(out_senders, b->out_senders.p, (size_t)b->n_sims * b->sender_stride * sizeof(RemySenderResult), cudaMemcpyDeviceToHost, g_stream));
  if (out_use_counts && b->out_counts.n)
    CU(cudaMemcpyAsync(out_use_counts, b->out_counts.p, b->out_counts.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  return 0;
}

uint32_t remy_b200_batch_launches(const RemyBatch *b) { return b ? b->launches : 0; }

void remy_b200_batch_destroy(RemyBatch *b)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!b) return;
  if (g_stream) cudaStreamSynchronize(g_stream);
  for (auto &p : b->passes) recycle_scratch(std::move(p->scratch));
  delete b;
}

int remy_b200_score_batch(const RemyFlatTree *tree,
                          const RemyLeafOverride *cands, uint32_t n_cands,
                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                          uint32_t sender_stride,
                          RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts)
{
  if (!out_sims && n_cfgs) return fail(REMY_EINVAL, "out_sims is NULL");
  RemyBatch *b = nullptr;
  int rc = remy_b200_batch_create(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride,
                                  out_senders != nullptr, out_use_counts != nullptr, &b);
  if (rc) return rc;
  rc = remy_b200_batch_run(b, nullptr);
  if (!rc) rc = remy_b200_batch_fetch(b, out_sims, out_senders, out_use_counts);
  remy_b200_batch_destroy(b);
  return rc;
}

/* ---- Fish senders over a FinTree --------------------------------------------------------- */
int remy_b200_score_batch_fish(const RemyFlatTree *tree,
                               const RemyLeafOverride *cands, uint32_t n_cands,
                               const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                               uint32_t sender_stride,
                               RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts)
{
  if (!out_sims && n_cfgs) return fail(REMY_EINVAL, "out_sims is NULL");
  RemyBatch *b = nullptr;
  int rc = create_impl(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, out_senders != nullptr, out_use_counts != nullptr, 0, &b, true);
  if (rc) return rc;
  rc = remy_b200_batch_run(b, nullptr);
  if (!rc) rc = remy_b200_batch_fetch(b, out_sims, out_senders, out_use_counts);
  remy_b200_batch_destroy(b);
  return rc;
}

/* ---- trace == true scoring ------------------------------------------------------------ */
static int score_trace_impl(const RemyFlatTree *tree,
                            const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                            uint32_t sender_stride,
                            RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts,
                            RemyTraceSink sink, void *user, bool fish);

int remy_b200_score_trace(const RemyFlatTree *tree,
                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                          uint32_t sender_stride,
                          RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts,
                          RemyTraceSink sink, void *user)
{
  return score_trace_impl(tree, cfgs, seeds, n_cfgs, sender_stride, out_sims, out_senders, out_use_counts, sink, user, false);
}

int remy_b200_score_trace_fish(const RemyFlatTree *tree,
                               const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                               uint32_t sender_stride,
                               RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts,
                               RemyTraceSink sink, void *user)
{
  return score_trace_impl(tree, cfgs, seeds, n_cfgs, sender_stride, out_sims, out_senders, out_use_counts, sink, user, true);
}

static int score_trace_impl(const RemyFlatTree *tree,
                            const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                            uint32_t sender_stride,
                            RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts,
                            RemyTraceSink sink, void *user, bool fish)
{
  if (!out_sims && n_cfgs) return fail(REMY_EINVAL, "out_sims is NULL");
  if (!sink) return fail(REMY_EINVAL, "sink is NULL");
  /* 1. counting run: scores, use counts, and the number of lookups of every simulation */
  RemyBatch *b = nullptr;
  int rc = create_impl(tree, nullptr, 1, cfgs, seeds, n_cfgs, sender_stride, out_senders != nullptr, 1, 1, &b, fish);
  if (rc) return rc;
  std::vector<uint64_t> lookups(n_cfgs);
  uint32_t words = b->trace_words;
  rc = remy_b200_batch_run(b, nullptr);
  if (!rc) rc = remy_b200_batch_fetch(b, out_sims, out_senders, out_use_counts);
  if (!rc && n_cfgs) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaError_t e = cudaMemcpy(lookups.data(), b->out_lookups.p, n_cfgs * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(REMY_ECUDA, std::string("cudaMemcpy(lookups): ") + cudaGetErrorString(e));
  }
  remy_b200_batch_destroy(b);
  if (rc) return rc;
  for (uint32_t k = 0; k < n_cfgs; k++)
    if (out_sims[k].error) return 0; /* a failed simulation has no defined trace (the reference would have exited) */

  /* 2. writing runs over consecutive groups of configs whose log fits the budget */
  size_t free_b = 0, total_b = 0;
  { std::lock_guard<std::mutex> lk(g_mu); CU(cudaSetDevice(g_device)); CU(cudaMemGetInfo(&free_b, &total_b)); }
  if (n_cfgs == 0) return 0;
  uint64_t budget = std::min<uint64_t>(free_b / 3, (uint64_t)32 << 30);
  if (const char *e = getenv("REMY_B200_TRACE_LOG_BYTES")) { const long long v = atoll(e); if (v > 0) budget = (uint64_t)v; }
  const uint64_t rec_bytes = (uint64_t)words * sizeof(uint64_t);
  const size_t stage_bytes = (size_t)64 << 20;
  struct Pinned { /* freed on every way out */
    uint64_t *p = nullptr;
    ~Pinned() { if (p) { std::lock_guard<std::mutex> lk(g_mu); cudaFreeHost(p); } }
  } pinned;
  { std::lock_guard<std::mutex> lk(g_mu); CU(cudaMallocHost(&pinned.p, stage_bytes)); }
  uint64_t *const stage = pinned.p;
  const uint64_t stage_recs = stage_bytes / rec_bytes;
  uint32_t g0 = 0;
  while (g0 < n_cfgs && !rc) {
    uint32_t g1 = g0; uint64_t recs = 0;
    std::vector<uint64_t> offs;
    while (g1 < n_cfgs && (g1 == g0 || (recs + lookups[g1]) * rec_bytes <= budget)) { offs.push_back(recs); recs += lookups[g1]; g1++; }
    if (recs * rec_bytes > budget && recs * rec_bytes > ((uint64_t)64 << 30)) { rc = fail(REMY_ENOMEM, "trace of one simulation exceeds the log budget"); break; }
    RemyBatch *g = nullptr;
    rc = create_impl(tree, nullptr, 1, cfgs + g0, seeds + g0, g1 - g0, sender_stride, 0, 1, 2, &g, fish);
    if (rc) break;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      cudaError_t e = g->trace_log.alloc(std::max<uint64_t>(recs, 1) * words);
      if (e == cudaSuccess) e = g->trace_off.alloc(g1 - g0);
      if (e == cudaSuccess) e = cudaMemcpy(g->trace_off.p, offs.data(), offs.size() * sizeof(uint64_t), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { cudaGetLastError(); rc = fail(REMY_ENOMEM, std::string("trace log allocation: ") + cudaGetErrorString(e)); }
    }
    if (!rc) rc = remy_b200_batch_run(g, nullptr);
    if (!rc) {
      std::lock_guard<std::mutex> lk(g_mu);
      cudaError_t e = cudaStreamSynchronize(g_stream);
      /* hand the records to the sink in config order, through a pinned staging buffer */
      for (uint32_t k = g0; k < g1 && e == cudaSuccess && !rc; k++) {
        uint64_t done = 0;
        const uint64_t n = lookups[k], base = offs[k - g0];
        do { /* (a config without lookups still gets one empty call) */
          const uint64_t m = std::min<uint64_t>(n - done, stage_recs);
          if (m) e = cudaMemcpy(stage, g->trace_log.p + (base + done) * words, m * rec_bytes, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) break;
          if (sink(user, k, stage, m, words)) rc = fail(REMY_EINVAL, "trace sink asked to stop");
          done += m;
        } while (done < n && !rc);
      }
      if (e != cudaSuccess) rc = fail(REMY_ECUDA, std::string("trace copy: ") + cudaGetErrorString(e));
    }
    remy_b200_batch_destroy(g);
    g0 = g1;
  }
  return rc;
}

} /* extern "C" */
