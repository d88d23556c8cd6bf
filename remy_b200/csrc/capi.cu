/*
 * capi.cu — implementation of include/remy_b200.h on top of the sm_100a kernel.
 *
 * Host side of the boundary only: argument checking, flattening the tree into
 * the device layout (with the per-node cell tables), the config-sorted work
 * list (one persistent launch per device and batch), the partition of ONE batch
 * over the bound devices, scratch sizing, kernel launches on the library's
 * streams, and the bounded re-run of simulations whose link queue outgrew the
 * ring they were given.  There is no CPU implementation of the simulator in
 * this library: without a CUDA device every entry point fails with REMY_ENODEV.
 *
 * Several devices (remy_b200_init_devices): the reference fans one improve step out
 * over all cores of the machine (src/breeder.cc:53-77); here the cost-sorted list of
 * (candidate, config) simulations is dealt round-robin to the bound GPUs, each runs
 * its share with its own stream and working memory, and the host gathers the
 * per-simulation results into the caller's arrays (scores are then summed per
 * candidate in config order by the caller, like src/evaluator.cc:96, so they do not
 * depend on the number of devices).  Nothing crosses between devices.
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

/* Per-slot working memory of one launch (slot = one resident thread). */
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

/* One bound CUDA device: its stream, recycled working memory and a pinned staging buffer. */
struct DeviceCtx {
  int device = -1, sms = 0;
  cudaStream_t stream = nullptr;
  /* Working memory of destroyed batches is kept for the next batch of the same geometry (an
   * Evaluator scores batch after batch of identical shape; cudaMalloc/cudaFree of tens of GB
   * would otherwise sit inside every one-shot call). */
  std::vector<Scratch> pool;
  void *pinned = nullptr; size_t pinned_bytes = 0; /* staging for results that are scattered on the host (several devices) */
};
constexpr size_t SCRATCH_POOL_MAX = 4;
std::vector<std::unique_ptr<DeviceCtx>> g_ctx; /* devices bound by remy_b200_init / remy_b200_init_devices */

void release_ctx(DeviceCtx &c)
{
  cudaSetDevice(c.device);
  c.pool.clear();
  if (c.pinned) { cudaFreeHost(c.pinned); c.pinned = nullptr; c.pinned_bytes = 0; }
  if (c.stream) { cudaStreamDestroy(c.stream); c.stream = nullptr; }
}

} // namespace

/* The part of a batch that runs on one device.  With one device it is the whole batch and simulation
 * ids are the caller's (candidate * n_cfgs + config); with several, every device numbers ITS
 * simulations 0..n_sims-1 and sim_map translates to the caller's ids (KernelArgs::sim_map). */
struct DevBatch {
  DeviceCtx *ctx = nullptr;
  uint32_t n_cands = 1, n_cfgs = 0, n_sims = 0, n_leaves = 0, n_nodes = 0;
  uint32_t sender_stride = 0; bool want_senders = false, want_counts = false, has_cands = false;
  uint32_t axes_mask = 0; size_t tree_bytes = 0;
  uint32_t delay_cap = 8;
  std::vector<uint32_t> sim_map; /* host copy; empty = the caller's ids */
  /* device */
  DevBuf<Bounds> bounds; DevBuf<DevNodeMeta> meta; DevBuf<DevLeaf> leaves; DevBuf<double> cuts; DevBuf<uint16_t> table;
  DevBuf<RemyLeafOverride> d_cands; DevBuf<RemyNetConfig> d_cfgs; DevBuf<uint32_t> d_seeds; DevBuf<uint32_t> d_sim_map;
  DevBuf<uint32_t> order; DevBuf<uint32_t> redo_list; DevBuf<uint32_t> redo_count; DevBuf<uint32_t> redo_counter;
  DevBuf<RemySimResult> out_sims; DevBuf<RemySenderResult> out_senders; DevBuf<uint32_t> out_counts;
  /* the one launch of the batch: persistent blocks over the cost-sorted work list */
  uint32_t n_max = 0, grid = 0; size_t smem = 0; bool tree_in_smem = false;
  uint32_t n_work = 0; /* entries of the work list */
  uint32_t lanes_per_warp = 32; /* lanes of a warp that take work (fewer when the list is spread, see create_part) */
  Scratch scratch; DevBuf<uint32_t> counter;
  uint32_t launches = 0, redone = 0;
  /* trace == true scoring (remy_b200_score_trace): 0 = off, 1 = counting run, 2 = writing run */
  int trace = 0; uint32_t trace_words = 0;
  bool fish = false;  /* Fish senders over a FinTree (remy_b200_score_batch_fish) */
  bool flows = false; /* Rat senders behind ByteSwitchedSender (remy_b200_score_batch_flows) */
  DevBuf<uint64_t> trace_log, trace_off, out_lookups;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ~DevBatch()
  {
    if (ctx) cudaSetDevice(ctx->device); /* the buffers are freed on the device that owns them */
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

struct RemyBatch {
  std::vector<std::unique_ptr<DevBatch>> parts; /* one per bound device that got work */
  uint32_t n_cands = 1, n_cfgs = 0, n_sims = 0, n_leaves = 0, sender_stride = 0;
  bool want_senders = false, want_counts = false, ran = false;
  uint32_t launches = 0, redone = 0;
};

namespace {

size_t smem_bytes(const DevBatch *b, uint32_t n_max, bool with_tree)
{
  return (size_t)(INC_BUF + n_max) * SIM_BLOCK * sizeof(double) + (with_tree ? b->tree_bytes : 0);
}

/* The kernel variants: use counts, lookup log, Fish senders, flow-switched senders - each built twice, for a
 * tree staged in shared memory (TSM: its reads are ld.shared) and for one read from global memory (a tree
 * too large for the carve-out). */
template <bool TSM> const void *kernel_fn_t(const DevBatch *b)
{
  if (b->trace) return b->fish ? (const void *)remy_sim_kernel<true, true, true, false, TSM> : (const void *)remy_sim_kernel<true, true, false, false, TSM>;
  if (b->fish) return b->want_counts ? (const void *)remy_sim_kernel<true, false, true, false, TSM> : (const void *)remy_sim_kernel<false, false, true, false, TSM>;
  if (b->flows) return b->want_counts ? (const void *)remy_sim_kernel<true, false, false, true, TSM> : (const void *)remy_sim_kernel<false, false, false, true, TSM>;
  return b->want_counts ? (const void *)remy_sim_kernel<true, false, false, false, TSM> : (const void *)remy_sim_kernel<false, false, false, false, TSM>;
}
const void *kernel_fn(const DevBatch *b, bool tree_in_smem) { return tree_in_smem ? kernel_fn_t<true>(b) : kernel_fn_t<false>(b); }

KernelArgs make_args(DevBatch *b, const Scratch &s, bool tree_in_smem, const uint32_t *order, uint32_t n_work, uint32_t *counter)
{
  KernelArgs a;
  a.bounds = b->bounds.p; a.meta = b->meta.p; a.leaves = b->leaves.p;
  a.cuts = b->cuts.p; a.table = b->table.p; a.n_cuts = (uint32_t)b->cuts.n; a.n_table = (uint32_t)b->table.n;
  a.n_nodes = b->n_nodes; a.n_leaves = b->n_leaves; a.axes_mask = b->axes_mask; a.tree_in_smem = tree_in_smem;
  a.cands = b->has_cands ? b->d_cands.p : nullptr; a.cfgs = b->d_cfgs.p; a.seeds = b->d_seeds.p;
  a.order = order; a.sim_map = b->sim_map.empty() ? nullptr : b->d_sim_map.p;
  a.n_cfgs = b->n_cfgs; a.n_work = n_work; a.work_counter = counter; a.lanes_per_warp = (order == b->order.p) ? b->lanes_per_warp : 32; /* (re-runs of overflowed simulations: every lane) */
  a.cold = s.cold.p; a.share = s.share.p; a.nsw = s.nsw.p; a.sendst = s.sendst.p; a.n_max = s.n_max; a.link_ring = s.link.p; a.link_cap = s.link_cap;
  a.delay_ring = s.delay.p; a.delay_cap = s.delay_cap; a.slot_counts = s.counts.p;
  a.out_sims = b->out_sims.p; a.out_senders = b->want_senders ? b->out_senders.p : nullptr;
  a.sender_stride = b->sender_stride; a.out_use_counts = b->want_counts ? b->out_counts.p : nullptr;
  a.trace_log = b->trace == 2 ? b->trace_log.p : nullptr; a.trace_off = b->trace == 2 ? b->trace_off.p : nullptr;
  a.out_lookups = b->trace ? b->out_lookups.p : nullptr; a.trace_words = b->trace_words;
  return a;
}

void recycle_scratch(DeviceCtx *c, Scratch &&s)
{
  if (!s.cold.p) return;
  if (c->pool.size() >= SCRATCH_POOL_MAX) c->pool.erase(c->pool.begin());
  c->pool.push_back(std::move(s));
}

int alloc_scratch(DevBatch *b, Scratch &s, uint32_t slots, uint32_t n_max, uint32_t link_cap, uint32_t delay_cap)
{
  std::vector<Scratch> &pool = b->ctx->pool;
  const uint32_t count_leaves = b->want_counts ? b->n_leaves : 0;
  for (size_t i = 0; i < pool.size(); i++) {
    Scratch &c = pool[i];
    if (c.slots == slots && c.n_max == n_max && c.link_cap == link_cap && c.delay_cap == delay_cap && c.count_leaves == count_leaves) {
      s = std::move(c);
      pool.erase(pool.begin() + i);
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
    if (attempt == 0 && !pool.empty()) { pool.clear(); continue; } /* give pooled memory back and retry */
    return fail(REMY_ENOMEM, std::string("scratch allocation failed: ") + cudaGetErrorString(e));
  }
  return fail(REMY_ENOMEM, "scratch allocation failed");
}

/* Shared memory and residency of the kernel for a launch that accepts up to n_max senders. */
int launch_geometry(DevBatch *b, uint32_t n_max, size_t &smem, bool &tree_in_smem, uint32_t &resident_blocks)
{
  tree_in_smem = smem_bytes(b, n_max, true) <= (SIM_BLOCK > 128 ? SIM_SMEM_MAX : (size_t)160 * 1024) && b->tree_bytes <= 64 * 1024;
  smem = smem_bytes(b, n_max, tree_in_smem);
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel_fn(b, tree_in_smem), SIM_BLOCK, smem));
  if (per_sm < 1) return fail(REMY_ECUDA, "kernel does not fit on an SM");
  /* tuning knob: fewer resident blocks shrink the working set (more of it stays in L2) at the
   * price of fewer warps to hide latency with */
  if (const char *cap = getenv("REMY_B200_MAX_BLOCKS_PER_SM")) { const int c = atoi(cap); if (c >= 1 && c < per_sm) per_sm = c; }
  resident_blocks = (uint32_t)per_sm * (uint32_t)b->ctx->sms;
  return 0;
}

int launch(DevBatch *b, const Scratch &s, size_t smem, bool tree_in_smem, const uint32_t *order, uint32_t n_work,
           uint32_t grid, uint32_t *counter, cudaStream_t stream)
{
  CU(cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream));
  const KernelArgs a = make_args(b, s, tree_in_smem, order, n_work, counter);
  void *params[] = {(void *)&a};
  CU(cudaLaunchKernel(kernel_fn(b, tree_in_smem), dim3(grid), dim3(SIM_BLOCK), params, smem, stream));
  CU(cudaGetLastError());
  b->launches++;
  return 0;
}

int bind_device(int device, int count)
{
  if (device < 0 || device >= count) return fail(REMY_EINVAL, "device index out of range");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(REMY_ENODEV, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "); this library is built for sm_100a only");
  std::unique_ptr<DeviceCtx> c(new DeviceCtx());
  c->device = device; c->sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  /* the kernel stages the tree in shared memory: allow the large carve-out (a per-device attribute) */
  {
    DevBatch probe; /* every variant: walk the selector */
    for (int v = 0; v < 16; v++) {
      probe.trace = (v & 8) ? 1 : 0; probe.fish = (v & 4) != 0; probe.flows = (v & 2) != 0; probe.want_counts = (v & 1) != 0;
      CU(cudaFuncSetAttribute(kernel_fn(&probe, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SIM_SMEM_MAX));
      CU(cudaFuncSetAttribute(kernel_fn(&probe, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SIM_SMEM_MAX));
    }
  }
  g_ctx.push_back(std::move(c));
  return 0;
}

/* Build one device's part of a batch.  `pairs` = the caller's simulation ids this device runs, in
 * work-list (cost-sorted) order; nullptr = the whole batch under the caller's ids, in plan.order. */
int create_part(DeviceCtx *ctx, const RemyFlatTree *tree, const PreparedTree &pt,
                const RemyLeafOverride *cands, uint32_t n_cands,
                const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                uint32_t sender_stride, bool want_senders, bool want_counts, int trace, int kind,
                const BatchPlan &plan, const std::vector<uint32_t> *pairs, std::unique_ptr<DevBatch> &out)
{
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  std::unique_ptr<DevBatch> guard(new DevBatch());
  DevBatch *b = guard.get();
  b->ctx = ctx;
  b->n_cands = n_cands; b->n_cfgs = n_cfgs; b->n_sims = pairs ? (uint32_t)pairs->size() : n_cands * n_cfgs;
  b->has_cands = cands != nullptr;
  b->want_senders = want_senders; b->want_counts = want_counts;
  b->sender_stride = sender_stride;
  b->trace = trace;
  b->fish = kind == 1;
  b->flows = kind == 2;
  b->n_nodes = tree->n_nodes; b->n_leaves = tree->n_leaves; b->axes_mask = pt.axes_mask; b->tree_bytes = pt.bytes();
  b->trace_words = 1u + (uint32_t)__builtin_popcount(pt.axes_mask);
  b->delay_cap = plan.delay_cap;
  b->n_max = std::max(plan.n_max, 4u);
  if (trace) CU(b->out_lookups.alloc(b->n_sims));
  CU(b->bounds.alloc(pt.bounds.size())); CU(b->meta.alloc(pt.meta.size())); CU(b->leaves.alloc(pt.leaves.size()));
  CU(b->cuts.alloc(pt.cuts.size())); CU(b->table.alloc(pt.table.size()));
  CU(cudaMemcpyAsync(b->bounds.p, pt.bounds.data(), pt.bounds.size() * sizeof(Bounds), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(b->meta.p, pt.meta.data(), pt.meta.size() * sizeof(DevNodeMeta), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(b->leaves.p, pt.leaves.data(), pt.leaves.size() * sizeof(DevLeaf), cudaMemcpyHostToDevice, st));
  if (!pt.cuts.empty()) CU(cudaMemcpyAsync(b->cuts.p, pt.cuts.data(), pt.cuts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  if (!pt.table.empty()) CU(cudaMemcpyAsync(b->table.p, pt.table.data(), pt.table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));

  /* ---- configs, seeds, candidates (replicated on every device: a few hundred KB), the work list ---- */
  std::vector<uint32_t> identity;
  const uint32_t *order_h = plan.order.data();
  if (pairs) {
    b->sim_map = *pairs;
    identity.resize(b->n_sims);
    for (uint32_t i = 0; i < b->n_sims; i++) identity[i] = i;
    order_h = identity.data();
    CU(b->d_sim_map.alloc(b->n_sims));
    if (b->n_sims) CU(cudaMemcpyAsync(b->d_sim_map.p, pairs->data(), (size_t)b->n_sims * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  }
  CU(b->d_cfgs.alloc(n_cfgs)); CU(b->d_seeds.alloc(n_cfgs));
  CU(b->redo_list.alloc(b->n_sims)); CU(b->redo_count.alloc(1)); CU(b->redo_counter.alloc(1));
  if (n_cfgs) {
    CU(cudaMemcpyAsync(b->d_cfgs.p, cfgs, n_cfgs * sizeof(RemyNetConfig), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b->d_seeds.p, seeds, n_cfgs * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  }
  if (cands) {
    CU(b->d_cands.alloc(n_cands));
    CU(cudaMemcpyAsync(b->d_cands.p, cands, n_cands * sizeof(RemyLeafOverride), cudaMemcpyHostToDevice, st));
  }
  CU(b->out_sims.alloc(b->n_sims));
  if (b->want_senders) CU(b->out_senders.alloc((size_t)b->n_sims * sender_stride));
  if (b->want_counts) CU(b->out_counts.alloc((size_t)n_cands * b->n_leaves));

  /* ---- launch geometry: persistent blocks, as many as are resident.  One launch over one cost-sorted
   * list (largest sender counts first): cutting the batch into a launch per sender-count bucket buys
   * nothing - residency is bound by registers, not by the per-sender shared memory - and costs a drain
   * per launch (a block leaves only when its slowest simulation has ended). ---- */
  uint32_t resident = 0;
  int rc = launch_geometry(b, b->n_max, b->smem, b->tree_in_smem, resident);
  if (rc) return rc;
  b->grid = std::max(1u, std::min(resident, (b->n_sims + SIM_BLOCK - 1) / SIM_BLOCK));
  b->n_work = b->n_sims;
  std::vector<uint32_t> spread;
  constexpr uint32_t WARPS_PER_BLOCK = SIM_BLOCK / 32;
  b->lanes_per_warp = 32;
  if (b->n_sims && b->n_sims < resident * SIM_BLOCK && !getenv("REMY_B200_NO_SPREAD")) {
    /* Fewer simulations than resident threads: the batch takes as long as its longest simulation, and a warp
     * executes the union of the code paths of its lanes - 32 equally long simulations in one warp run at the
     * pace of all their branches.  The batch is therefore SPREAD over all resident warps: only the first
     * lanes_per_warp lanes of a warp take work, and the list is regrouped so that the simulations a warp pulls
     * together (its lanes' first pull is one aggregated atomic: consecutive entries) come from the cost-sorted
     * list with a stride of the number of warps - one of the longest, one of the next stratum, ... - and the
     * long ones run more and more alone as the short ones of their warp end.  Measured on a B200: a lone
     * evaluator pass (800 simulations) 1.87 s -> 0.86 s, 24 000 simulations 2.05 s -> 1.78 s. */
    const uint32_t warps = resident * WARPS_PER_BLOCK;
    const uint32_t per_warp = (b->n_sims + warps - 1) / warps;            /* 1..32 */
    const uint32_t used = (b->n_sims + per_warp - 1) / per_warp;         /* warps that get work */
    b->lanes_per_warp = per_warp;
    b->grid = (used + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    spread.reserve(b->n_sims);
    for (uint32_t c = 0; c < used; c++)
      for (uint32_t l = 0; l < per_warp; l++)
        if ((uint64_t)l * used + c < b->n_sims) spread.push_back(order_h[l * used + c]);
    order_h = spread.data();
  }
  CU(b->order.alloc(b->n_work));
  if (b->n_work) CU(cudaMemcpyAsync(b->order.p, order_h, (size_t)b->n_work * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  const uint32_t count_leaves = b->want_counts ? b->n_leaves : 0;
  /* keep the working memory within a budget: shrink the first-try link ring if needed
   * (overflowing simulations are re-run with larger rings) */
  uint32_t link_cap = plan.link_cap;
  /* tuning knob: first-try ring size for buffers larger than it (overflowing simulations are re-run) */
  if (const char *lc = getenv("REMY_B200_LINK_CAP")) { const long c = atol(lc); if (c >= 16 && (uint32_t)c < link_cap) link_cap = next_pow2((uint64_t)c); }
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  const size_t budget = std::min<size_t>(free_b / 2, (size_t)64 << 30);
  while ((size_t)b->grid * SIM_BLOCK * Scratch::bytes_per_slot(b->n_max, link_cap, plan.delay_cap, count_leaves) > budget && link_cap > 1024)
    link_cap /= 2;
  rc = alloc_scratch(b, b->scratch, b->grid * SIM_BLOCK, b->n_max, link_cap, plan.delay_cap);
  if (rc) return rc;
  CU(b->counter.alloc(1));
  CU(cudaEventCreate(&b->ev0)); CU(cudaEventCreate(&b->ev1));
  out = std::move(guard);
  return 0;
}

/* Enqueue one part's launch and the scan for overflowed simulations; nothing here waits for the device. */
int run_part_begin(DevBatch *b)
{
  CU(cudaSetDevice(b->ctx->device));
  cudaStream_t st = b->ctx->stream;
  b->launches = 0; b->redone = 0;
  if (b->want_counts) CU(cudaMemsetAsync(b->out_counts.p, 0, b->out_counts.n * sizeof(uint32_t), st));
  CU(cudaEventRecord(b->ev0, st));
  if (b->n_sims) {
    int rc = launch(b, b->scratch, b->smem, b->tree_in_smem, b->order.p, b->n_work, b->grid, b->counter.p, st);
    if (rc) return rc;
    CU(cudaMemsetAsync(b->redo_count.p, 0, sizeof(uint32_t), st));
    remy_collect_overflow<<<(b->n_sims + 255) / 256, 256, 0, st>>>(b->out_sims.p, b->n_sims, b->redo_list.p, b->redo_count.p);
    CU(cudaGetLastError());
    b->launches++;
  }
  return 0;
}

/* Simulations whose link queue outgrew the ring are re-run with larger rings, inside the timed
 * region.  (The reference's deque is unbounded; a finite ring is this library's artefact, so it is
 * hidden from the caller as long as memory allows.) */
int run_part_end(DevBatch *b)
{
  CU(cudaSetDevice(b->ctx->device));
  cudaStream_t st = b->ctx->stream;
  if (b->n_sims) {
    uint32_t cap = b->scratch.link_cap;
    for (int round = 0; round < 4; round++) {
      uint32_t n_redo = 0;
      CU(cudaMemcpyAsync(&n_redo, b->redo_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      if (n_redo == 0 || cap >= (1u << 24)) break;
      if (round == 0) b->redone = n_redo;
      cap = (uint32_t)std::min<uint64_t>((uint64_t)cap * 16, 1u << 24);
      const size_t per_slot = Scratch::bytes_per_slot(REMY_MAX_SENDERS, cap, b->delay_cap, b->want_counts ? b->n_leaves : 0);
      uint32_t slots = (uint32_t)std::min<size_t>(std::max<size_t>(((size_t)32 << 30) / per_slot, SIM_BLOCK), n_redo);
      slots = ((slots + SIM_BLOCK - 1) / SIM_BLOCK) * SIM_BLOCK;
      Scratch s2;
      int rc = alloc_scratch(b, s2, slots, REMY_MAX_SENDERS, cap, b->delay_cap);
      if (rc) return rc;
      size_t smem = 0; bool tis = false; uint32_t resident = 0;
      rc = launch_geometry(b, REMY_MAX_SENDERS, smem, tis, resident);
      if (rc) return rc;
      rc = launch(b, s2, smem, tis, b->redo_list.p, n_redo, slots / SIM_BLOCK, b->redo_counter.p, st);
      if (rc) return rc;
      CU(cudaStreamSynchronize(st)); /* s2 is freed at the end of this scope; redo_list is rebuilt next */
      CU(cudaMemsetAsync(b->redo_count.p, 0, sizeof(uint32_t), st));
      remy_collect_overflow<<<(b->n_sims + 255) / 256, 256, 0, st>>>(b->out_sims.p, b->n_sims, b->redo_list.p, b->redo_count.p);
      CU(cudaGetLastError());
      b->launches++;
    }
  }
  CU(cudaEventRecord(b->ev1, st));
  return 0;
}

} // namespace

extern "C" {

uint32_t remy_b200_abi_version(void) { return REMY_B200_ABI_VERSION; }
const char *remy_b200_last_error(void) { return g_err.c_str(); }

int remy_b200_init_devices(const int *devices, int n_devices)
{
  std::lock_guard<std::mutex> lk(g_mu);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(REMY_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
  std::vector<int> want;
  if (!devices || n_devices <= 0) for (int d = 0; d < count; d++) want.push_back(d); /* all visible devices */
  else want.assign(devices, devices + n_devices);
  for (size_t i = 0; i < want.size(); i++)
    for (size_t j = 0; j < i; j++)
      if (want[i] == want[j]) return fail(REMY_EINVAL, "device listed twice");
  bool same = want.size() == g_ctx.size();
  for (size_t i = 0; same && i < want.size(); i++) same = g_ctx[i]->device == want[i];
  if (same) return 0;
  for (auto &c : g_ctx) release_ctx(*c);
  g_ctx.clear();
  for (int d : want) {
    int rc = bind_device(d, count);
    if (rc) { for (auto &c : g_ctx) release_ctx(*c); g_ctx.clear(); return rc; }
  }
  g_err.clear();
  return 0;
}

int remy_b200_init(int device) { return remy_b200_init_devices(&device, 1); }

int remy_b200_device_count(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  return (int)g_ctx.size();
}

void remy_b200_shutdown(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &c : g_ctx) release_ctx(*c);
  g_ctx.clear();
}

static int create_impl(const RemyFlatTree *tree,
                       const RemyLeafOverride *cands, uint32_t n_cands,
                       const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                       uint32_t sender_stride, int want_senders, int want_use_counts, int trace,
                       RemyBatch **out_batch, int kind = 0 /* 0 Rat, 1 Fish, 2 Rat behind ByteSwitchedSender */)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out_batch) return fail(REMY_EINVAL, "out_batch is NULL");
  if (kind == 2 && trace) return fail(REMY_EINVAL, "no lookup log for flow-switched senders");
  *out_batch = nullptr;
  if (g_ctx.empty()) return fail(REMY_ENODEV, "remy_b200_init has not been called");
  if (!tree || !tree->nodes || !tree->leaves || tree->n_nodes == 0 || tree->n_leaves == 0)
    return fail(REMY_EINVAL, "empty tree");
  if (n_cfgs && (!cfgs || !seeds)) return fail(REMY_EINVAL, "cfgs/seeds is NULL");
  if (!cands) n_cands = 1;
  if ((uint64_t)n_cands * n_cfgs > 0x7fffffffull) return fail(REMY_EINVAL, "too many simulations in one batch");

  /* ---- tree -> device layout (validated so the kernel cannot walk off the arrays) ---- */
  PreparedTree pt;
  const std::string terr = prepare_tree(tree, pt);
  if (!terr.empty()) return fail(REMY_EINVAL, terr);
  if (cands)
    for (uint32_t c = 0; c < n_cands; c++)
      if (cands[c].leaf_index != REMY_NO_OVERRIDE && cands[c].leaf_index >= tree->n_leaves)
        return fail(REMY_EINVAL, "candidate " + std::to_string(c) + " overrides a leaf that does not exist");

  /* ---- ring sizes and the cost-sorted work list ("instances sorted by config") ---- */
  const BatchPlan plan = plan_batch(cfgs, n_cfgs, n_cands, true);
  if (want_senders && sender_stride < plan.n_max) return fail(REMY_EINVAL, "sender_stride smaller than the largest num_senders");

  std::unique_ptr<RemyBatch> guard(new RemyBatch());
  RemyBatch *B = guard.get();
  B->n_cands = n_cands; B->n_cfgs = n_cfgs; B->n_sims = n_cands * n_cfgs; B->n_leaves = tree->n_leaves;
  B->sender_stride = sender_stride;
  B->want_senders = want_senders != 0; B->want_counts = want_use_counts != 0 || trace != 0;

  /* ---- partition: the sorted list is dealt round-robin to the devices, so that each gets the same mix
   * of sender counts and run lengths, and neighbours in a device's list (the lanes of a warp) still share
   * a config.  The lookup log of trace == true scoring is laid out for one device. ---- */
  size_t n_dev = trace ? 1 : g_ctx.size();
  if (B->n_sims < 2 * SIM_BLOCK * n_dev) n_dev = 1; /* not worth spreading */
  for (size_t d = 0; d < n_dev; d++) {
    std::unique_ptr<DevBatch> part;
    std::vector<uint32_t> pairs;
    if (n_dev > 1) for (size_t i = d; i < plan.order.size(); i += n_dev) pairs.push_back(plan.order[i]);
    int rc = create_part(g_ctx[d].get(), tree, pt, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride,
                         B->want_senders, B->want_counts, trace, kind, plan, n_dev > 1 ? &pairs : nullptr, part);
    if (rc) return rc;
    B->parts.push_back(std::move(part));
  }
  for (auto &p : B->parts) { CU(cudaSetDevice(p->ctx->device)); CU(cudaStreamSynchronize(p->ctx->stream)); }
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

int remy_b200_batch_create_fish(const RemyFlatTree *tree,
                                const RemyLeafOverride *cands, uint32_t n_cands,
                                const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                uint32_t sender_stride, int want_senders, int want_use_counts,
                                RemyBatch **out_batch)
{
  return create_impl(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, want_senders, want_use_counts, 0, out_batch, 1);
}

int remy_b200_batch_create_flows(const RemyFlatTree *tree,
                                 const RemyLeafOverride *cands, uint32_t n_cands,
                                 const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                 uint32_t sender_stride, int want_senders, int want_use_counts,
                                 RemyBatch **out_batch)
{
  return create_impl(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, want_senders, want_use_counts, 0, out_batch, 2);
}

int remy_b200_batch_run(RemyBatch *B, float *kernel_ms)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!B) return fail(REMY_EINVAL, "batch is NULL");
  /* every device's launch is enqueued before any of them is waited for */
  for (auto &p : B->parts) { int rc = run_part_begin(p.get()); if (rc) return rc; }
  for (auto &p : B->parts) { int rc = run_part_end(p.get()); if (rc) return rc; }
  B->launches = 0; B->redone = 0;
  for (auto &p : B->parts) { B->launches += p->launches; B->redone += p->redone; }
  if (kernel_ms) {
    float worst = 0;
    for (auto &p : B->parts) { /* device time of the slowest part: CUDA events on its launching stream */
      CU(cudaSetDevice(p->ctx->device));
      CU(cudaEventSynchronize(p->ev1));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
      worst = std::max(worst, ms);
    }
    *kernel_ms = worst;
  }
  B->ran = true;
  return 0;
}

int remy_b200_batch_fetch(RemyBatch *B, RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!B || (!out_sims && B->n_sims)) return fail(REMY_EINVAL, "batch/out_sims is NULL");
  if (!B->ran) return fail(REMY_EINVAL, "batch has not been run");
  if (out_senders && !B->want_senders) return fail(REMY_EINVAL, "batch was created without want_senders");
  if (out_use_counts && !B->want_counts) return fail(REMY_EINVAL, "batch was created without want_use_counts");
  const size_t n_counts = (size_t)B->n_cands * B->n_leaves;
  if (B->parts.size() == 1 && B->parts[0]->sim_map.empty()) { /* one device: straight into the caller's buffers */
    DevBatch *b = B->parts[0].get();
    cudaStream_t st = b->ctx->stream;
    CU(cudaSetDevice(b->ctx->device));
    if (b->n_sims)
      CU(cudaMemcpyAsync(out_sims, b->out_sims.p, (size_t)b->n_sims * sizeof(RemySimResult), cudaMemcpyDeviceToHost, st));
    if (out_senders && b->n_sims)
      CU(cudaMemcpyAsync(out_senders, b->out_senders.p, (size_t)b->n_sims * b->sender_stride * sizeof(RemySenderResult), cudaMemcpyDeviceToHost, st));
    if (out_use_counts && b->out_counts.n)
      CU(cudaMemcpyAsync(out_use_counts, b->out_counts.p, b->out_counts.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
  }
  /* several devices: every part is copied into its device's pinned staging buffer (all copies in flight
   * together), then scattered to the caller's simulation ids; use counts are summed (modulo 2^32, like
   * the reference's unsigned counters) */
  const size_t row = (size_t)B->sender_stride * sizeof(RemySenderResult);
  for (auto &p : B->parts) {
    DevBatch *b = p.get(); DeviceCtx *c = b->ctx;
    CU(cudaSetDevice(c->device));
    const size_t need = (size_t)b->n_sims * sizeof(RemySimResult) + (out_senders ? (size_t)b->n_sims * row : 0) + (out_use_counts ? n_counts * sizeof(uint32_t) : 0);
    if (c->pinned_bytes < need) {
      if (c->pinned) { cudaFreeHost(c->pinned); c->pinned = nullptr; c->pinned_bytes = 0; }
      CU(cudaMallocHost(&c->pinned, need));
      c->pinned_bytes = need;
    }
    unsigned char *dst = (unsigned char *)c->pinned;
    CU(cudaMemcpyAsync(dst, b->out_sims.p, (size_t)b->n_sims * sizeof(RemySimResult), cudaMemcpyDeviceToHost, c->stream));
    dst += (size_t)b->n_sims * sizeof(RemySimResult);
    if (out_senders) { CU(cudaMemcpyAsync(dst, b->out_senders.p, (size_t)b->n_sims * row, cudaMemcpyDeviceToHost, c->stream)); dst += (size_t)b->n_sims * row; }
    if (out_use_counts) CU(cudaMemcpyAsync(dst, b->out_counts.p, n_counts * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  if (out_use_counts) memset(out_use_counts, 0, n_counts * sizeof(uint32_t));
  for (auto &p : B->parts) {
    DevBatch *b = p.get(); DeviceCtx *c = b->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const unsigned char *src = (const unsigned char *)c->pinned;
    const RemySimResult *sims = (const RemySimResult *)src;
    for (uint32_t i = 0; i < b->n_sims; i++) out_sims[b->sim_map[i]] = sims[i];
    src += (size_t)b->n_sims * sizeof(RemySimResult);
    if (out_senders) {
      for (uint32_t i = 0; i < b->n_sims; i++) memcpy((unsigned char *)out_senders + (size_t)b->sim_map[i] * row, src + (size_t)i * row, row);
      src += (size_t)b->n_sims * row;
    }
    if (out_use_counts) {
      const uint32_t *cnt = (const uint32_t *)src;
      for (size_t i = 0; i < n_counts; i++) out_use_counts[i] += cnt[i];
    }
  }
  return 0;
}

uint32_t remy_b200_batch_launches(const RemyBatch *b) { return b ? b->launches : 0; }
uint32_t remy_b200_batch_rerun(const RemyBatch *b) { return b ? b->redone : 0; }
uint32_t remy_b200_batch_devices(const RemyBatch *b) { return b ? (uint32_t)b->parts.size() : 0; }

void remy_b200_batch_destroy(RemyBatch *B)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!B) return;
  for (auto &p : B->parts) {
    cudaSetDevice(p->ctx->device);
    if (p->ctx->stream) cudaStreamSynchronize(p->ctx->stream);
    recycle_scratch(p->ctx, std::move(p->scratch));
  }
  delete B;
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
  int rc = create_impl(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, out_senders != nullptr, out_use_counts != nullptr, 0, &b, 1);
  if (rc) return rc;
  rc = remy_b200_batch_run(b, nullptr);
  if (!rc) rc = remy_b200_batch_fetch(b, out_sims, out_senders, out_use_counts);
  remy_b200_batch_destroy(b);
  return rc;
}

/* ---- Rat senders behind ByteSwitchedSender: flows of a random length in packets ------------- */
int remy_b200_score_batch_flows(const RemyFlatTree *tree,
                                const RemyLeafOverride *cands, uint32_t n_cands,
                                const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                                uint32_t sender_stride,
                                RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts)
{
  if (!out_sims && n_cfgs) return fail(REMY_EINVAL, "out_sims is NULL");
  RemyBatch *b = nullptr;
  int rc = create_impl(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride, out_senders != nullptr, out_use_counts != nullptr, 0, &b, 2);
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
  /* (trace batches live on the first bound device: one tree, one log) */
  /* 1. counting run: scores, use counts, and the number of lookups of every simulation */
  RemyBatch *B = nullptr;
  int rc = create_impl(tree, nullptr, 1, cfgs, seeds, n_cfgs, sender_stride, out_senders != nullptr, 1, 1, &B, fish ? 1 : 0);
  if (rc) return rc;
  std::vector<uint64_t> lookups(n_cfgs);
  const uint32_t words = B->parts[0]->trace_words;
  const int device = B->parts[0]->ctx->device;
  cudaStream_t stream = B->parts[0]->ctx->stream;
  rc = remy_b200_batch_run(B, nullptr);
  if (!rc) rc = remy_b200_batch_fetch(B, out_sims, out_senders, out_use_counts);
  if (!rc && n_cfgs) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaSetDevice(device);
    cudaError_t e = cudaMemcpy(lookups.data(), B->parts[0]->out_lookups.p, n_cfgs * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(REMY_ECUDA, std::string("cudaMemcpy(lookups): ") + cudaGetErrorString(e));
  }
  remy_b200_batch_destroy(B);
  if (rc) return rc;
  for (uint32_t k = 0; k < n_cfgs; k++)
    if (out_sims[k].error) return 0; /* a failed simulation has no defined trace (the reference would have exited) */

  /* 2. writing runs over consecutive groups of configs whose log fits the budget */
  size_t free_b = 0, total_b = 0;
  { std::lock_guard<std::mutex> lk(g_mu); CU(cudaSetDevice(device)); CU(cudaMemGetInfo(&free_b, &total_b)); }
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
    RemyBatch *G = nullptr;
    rc = create_impl(tree, nullptr, 1, cfgs + g0, seeds + g0, g1 - g0, sender_stride, 0, 1, 2, &G, fish ? 1 : 0);
    if (rc) break;
    DevBatch *g = G->parts[0].get();
    {
      std::lock_guard<std::mutex> lk(g_mu);
      cudaSetDevice(device);
      cudaError_t e = g->trace_log.alloc(std::max<uint64_t>(recs, 1) * words);
      if (e == cudaSuccess) e = g->trace_off.alloc(g1 - g0);
      if (e == cudaSuccess) e = cudaMemcpy(g->trace_off.p, offs.data(), offs.size() * sizeof(uint64_t), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { cudaGetLastError(); rc = fail(REMY_ENOMEM, std::string("trace log allocation: ") + cudaGetErrorString(e)); }
    }
    if (!rc) rc = remy_b200_batch_run(G, nullptr);
    if (!rc) {
      std::lock_guard<std::mutex> lk(g_mu);
      cudaSetDevice(device);
      cudaError_t e = cudaStreamSynchronize(stream);
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
    remy_b200_batch_destroy(G);
    g0 = g1;
  }
  return rc;
}

} /* extern "C" */
