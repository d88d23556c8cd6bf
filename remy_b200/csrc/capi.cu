/*
 * capi.cu — implementation of include/remy_b200.h on top of the sm_100a kernel.
 *
 * Host side of the boundary only: argument checking, flattening the tree into
 * the device layout, the cost-sorted work list (instances sorted by config so
 * that neighbouring threads run similar simulations), scratch sizing, kernel
 * launches on the library's stream, and the bounded re-run of simulations whose
 * link queue outgrew the ring they were given.  There is no CPU implementation
 * of the simulator in this library: without a CUDA device every entry point
 * fails with REMY_ENODEV.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
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
cudaStream_t g_stream = nullptr;

int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(REMY_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)

template <class T> struct DevBuf {
  T *p = nullptr; size_t n = 0;
  cudaError_t alloc(size_t count) { free(); n = count; if (!count) return cudaSuccess; return cudaMalloc(&p, count * sizeof(T)); }
  void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { free(); }
};

struct Scratch {
  DevBuf<SenderCold> cold; DevBuf<LinkEnt> link; DevBuf<DelayEnt> delay; DevBuf<uint32_t> counts;
  uint32_t slots = 0, link_cap = 0, delay_cap = 0;
};

} // namespace

struct RemyBatch {
  /* host copies */
  std::vector<RemyNetConfig> cfgs; std::vector<uint32_t> seeds; std::vector<RemyLeafOverride> cands;
  uint32_t n_cands = 1, n_cfgs = 0, n_sims = 0, n_leaves = 0, n_nodes = 0, n_max = 1;
  uint32_t sender_stride = 0; bool want_senders = false, want_counts = false, has_cands = false;
  uint32_t axes_mask = 0; bool tree_in_smem = false; size_t tree_bytes = 0;
  /* device */
  DevBuf<Bounds> bounds; DevBuf<DevNodeMeta> meta; DevBuf<DevLeaf> leaves; DevBuf<double> cuts; DevBuf<uint16_t> table;
  DevBuf<RemyLeafOverride> d_cands; DevBuf<RemyNetConfig> d_cfgs; DevBuf<uint32_t> d_seeds;
  DevBuf<uint32_t> order; DevBuf<uint32_t> work_counter; DevBuf<uint32_t> redo_list; DevBuf<uint32_t> redo_count;
  DevBuf<RemySimResult> out_sims; DevBuf<RemySenderResult> out_senders; DevBuf<uint32_t> out_counts;
  Scratch scratch;
  uint32_t grid = 0; size_t smem = 0;
  uint32_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ran = false;
};

namespace {

size_t smem_bytes(const RemyBatch *b, bool with_tree)
{
  return (size_t)2 * b->n_max * SIM_BLOCK * sizeof(double) + (with_tree ? b->tree_bytes : 0);
}

KernelArgs make_args(RemyBatch *b, const Scratch &s, const uint32_t *order, uint32_t n_work)
{
  KernelArgs a;
  a.bounds = b->bounds.p; a.meta = b->meta.p; a.leaves = b->leaves.p;
  a.cuts = b->cuts.p; a.table = b->table.p; a.n_cuts = (uint32_t)b->cuts.n; a.n_table = (uint32_t)b->table.n;
  a.n_nodes = b->n_nodes; a.n_leaves = b->n_leaves; a.axes_mask = b->axes_mask; a.tree_in_smem = b->tree_in_smem;
  a.cands = b->has_cands ? b->d_cands.p : nullptr; a.cfgs = b->d_cfgs.p; a.seeds = b->d_seeds.p;
  a.order = order; a.n_cfgs = b->n_cfgs; a.n_work = n_work; a.work_counter = b->work_counter.p;
  a.cold = s.cold.p; a.n_max = b->n_max; a.link_ring = s.link.p; a.link_cap = s.link_cap;
  a.delay_ring = s.delay.p; a.delay_cap = s.delay_cap; a.slot_counts = s.counts.p;
  a.out_sims = b->out_sims.p; a.out_senders = b->want_senders ? b->out_senders.p : nullptr;
  a.sender_stride = b->sender_stride; a.out_use_counts = b->want_counts ? b->out_counts.p : nullptr;
  return a;
}

int alloc_scratch(RemyBatch *b, Scratch &s, uint32_t slots, uint32_t link_cap, uint32_t delay_cap)
{
  s.slots = slots; s.link_cap = link_cap; s.delay_cap = delay_cap;
  CU(s.cold.alloc((size_t)slots * b->n_max));
  CU(s.link.alloc((size_t)slots * link_cap));
  CU(s.delay.alloc((size_t)slots * delay_cap));
  if (b->want_counts) CU(s.counts.alloc((size_t)slots * b->n_leaves));
  return 0;
}

int launch(RemyBatch *b, const Scratch &s, const uint32_t *order, uint32_t n_work, uint32_t grid)
{
  CU(cudaMemsetAsync(b->work_counter.p, 0, sizeof(uint32_t), g_stream));
  const KernelArgs a = make_args(b, s, order, n_work);
  if (b->want_counts) remy_sim_kernel<true><<<grid, SIM_BLOCK, b->smem, g_stream>>>(a);
  else remy_sim_kernel<false><<<grid, SIM_BLOCK, b->smem, g_stream>>>(a);
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
  g_device = device; g_sms = prop.multiProcessorCount;
  g_err.clear();
  return 0;
}

void remy_b200_shutdown(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  g_device = -1;
}

int remy_b200_batch_create(const RemyFlatTree *tree,
                           const RemyLeafOverride *cands, uint32_t n_cands,
                           const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                           uint32_t sender_stride, int want_senders, int want_use_counts,
                           RemyBatch **out_batch)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out_batch) return fail(REMY_EINVAL, "out_batch is NULL");
  *out_batch = nullptr;
  if (g_device < 0 || !g_stream) return fail(REMY_ENODEV, "remy_b200_init has not been called");
  if (!tree || !tree->nodes || !tree->leaves || tree->n_nodes == 0 || tree->n_leaves == 0)
    return fail(REMY_EINVAL, "empty tree");
  if (!cfgs || !seeds) return fail(REMY_EINVAL, "cfgs/seeds is NULL");
  if (!cands) n_cands = 1;
  if ((uint64_t)n_cands * n_cfgs > 0x7fffffffull) return fail(REMY_EINVAL, "too many simulations in one batch");
  CU(cudaSetDevice(g_device));

  RemyBatch *b = new RemyBatch();
  std::unique_ptr<RemyBatch> guard(b);
  b->n_cands = n_cands; b->n_cfgs = n_cfgs; b->n_sims = n_cands * n_cfgs;
  b->has_cands = cands != nullptr;
  b->want_senders = want_senders != 0; b->want_counts = want_use_counts != 0;
  b->sender_stride = sender_stride;
  b->cfgs.assign(cfgs, cfgs + n_cfgs); b->seeds.assign(seeds, seeds + n_cfgs);
  if (cands) b->cands.assign(cands, cands + n_cands);

  /* ---- tree -> device layout (validated so the kernel cannot walk off the arrays) ---- */
  PreparedTree pt;
  const std::string terr = prepare_tree(tree, pt);
  if (!terr.empty()) return fail(REMY_EINVAL, terr);
  b->n_nodes = tree->n_nodes; b->n_leaves = tree->n_leaves; b->axes_mask = pt.axes_mask; b->tree_bytes = pt.bytes();
  if (cands)
    for (uint32_t c = 0; c < n_cands; c++)
      if (cands[c].leaf_index != REMY_NO_OVERRIDE && cands[c].leaf_index >= tree->n_leaves)
        return fail(REMY_EINVAL, "candidate " + std::to_string(c) + " overrides a leaf that does not exist");
  CU(b->bounds.alloc(pt.bounds.size())); CU(b->meta.alloc(pt.meta.size())); CU(b->leaves.alloc(pt.leaves.size()));
  CU(cudaMemcpyAsync(b->bounds.p, pt.bounds.data(), pt.bounds.size() * sizeof(Bounds), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(b->meta.p, pt.meta.data(), pt.meta.size() * sizeof(DevNodeMeta), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(b->leaves.p, pt.leaves.data(), pt.leaves.size() * sizeof(DevLeaf), cudaMemcpyHostToDevice, g_stream));
  CU(b->cuts.alloc(pt.cuts.size())); CU(b->table.alloc(pt.table.size()));
  if (!pt.cuts.empty()) CU(cudaMemcpyAsync(b->cuts.p, pt.cuts.data(), pt.cuts.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  if (!pt.table.empty()) CU(cudaMemcpyAsync(b->table.p, pt.table.data(), pt.table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, g_stream));

  /* ---- configs: n_max, ring sizes, cost-sorted order ---- */
  const BatchPlan plan = plan_batch(cfgs, n_cfgs, n_cands);
  b->n_max = plan.n_max;
  if (b->want_senders && sender_stride < plan.n_max) return fail(REMY_EINVAL, "sender_stride smaller than the largest num_senders");
  const uint32_t link_cap = plan.link_cap, delay_cap = plan.delay_cap;
  const std::vector<uint32_t> &order = plan.order;

  CU(b->d_cfgs.alloc(n_cfgs)); CU(b->d_seeds.alloc(n_cfgs)); CU(b->order.alloc(b->n_sims)); CU(b->work_counter.alloc(1)); CU(b->redo_list.alloc(b->n_sims)); CU(b->redo_count.alloc(1));
  CU(cudaMemcpyAsync(b->d_cfgs.p, cfgs, n_cfgs * sizeof(RemyNetConfig), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(b->d_seeds.p, seeds, n_cfgs * sizeof(uint32_t), cudaMemcpyHostToDevice, g_stream));
  CU(cudaMemcpyAsync(b->order.p, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, g_stream));
  if (cands) {
    CU(b->d_cands.alloc(n_cands));
    CU(cudaMemcpyAsync(b->d_cands.p, cands, n_cands * sizeof(RemyLeafOverride), cudaMemcpyHostToDevice, g_stream));
  }
  CU(b->out_sims.alloc(b->n_sims));
  if (b->want_senders) CU(b->out_senders.alloc((size_t)b->n_sims * sender_stride));
  if (b->want_counts) CU(b->out_counts.alloc((size_t)n_cands * b->n_leaves));

  /* ---- launch geometry: persistent blocks, as many as are resident ---- */
  b->tree_in_smem = smem_bytes(b, true) <= 200 * 1024;
  b->smem = smem_bytes(b, b->tree_in_smem);
  const void *kfn = b->want_counts ? (const void *)remy_sim_kernel<true> : (const void *)remy_sim_kernel<false>;
  CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, SIM_BLOCK, b->smem));
  if (per_sm < 1) return fail(REMY_ECUDA, "kernel does not fit on an SM");
  const uint32_t resident = (uint32_t)per_sm * (uint32_t)g_sms;
  b->grid = std::max(1u, std::min(resident, (b->n_sims + SIM_BLOCK - 1) / SIM_BLOCK));
  int rc = alloc_scratch(b, b->scratch, b->grid * SIM_BLOCK, link_cap, delay_cap);
  if (rc) return rc;
  CU(cudaEventCreate(&b->ev0)); CU(cudaEventCreate(&b->ev1));
  CU(cudaStreamSynchronize(g_stream));
  *out_batch = guard.release();
  return 0;
}

int remy_b200_batch_run(RemyBatch *b, float *kernel_ms)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!b) return fail(REMY_EINVAL, "batch is NULL");
  CU(cudaSetDevice(g_device));
  b->launches = 0;
  if (b->want_counts) CU(cudaMemsetAsync(b->out_counts.p, 0, b->out_counts.n * sizeof(uint32_t), g_stream));
  if (kernel_ms) CU(cudaEventRecord(b->ev0, g_stream));
  if (b->n_sims) {
    int rc = launch(b, b->scratch, b->order.p, b->n_sims, b->grid);
    if (rc) return rc;
    /* Simulations whose link queue outgrew the ring are re-run with larger rings, inside
     * the timed region.  (The reference's deque is unbounded; a finite ring is this
     * library's artefact, so it is hidden from the caller as long as memory allows.) */
    uint32_t cap = b->scratch.link_cap;
    for (int round = 0; round < 3 && cap < (1u << 24); round++) {
      CU(cudaMemsetAsync(b->redo_count.p, 0, sizeof(uint32_t), g_stream));
      remy_collect_overflow<<<(b->n_sims + 255) / 256, 256, 0, g_stream>>>(b->out_sims.p, b->n_sims, b->redo_list.p, b->redo_count.p);
      CU(cudaGetLastError());
      b->launches++;
      uint32_t n_redo = 0;
      CU(cudaMemcpyAsync(&n_redo, b->redo_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, g_stream));
      CU(cudaStreamSynchronize(g_stream));
      if (n_redo == 0) break;
      cap = (uint32_t)std::min<uint64_t>((uint64_t)cap * 16, 1u << 24);
      const size_t per_slot = (size_t)cap * sizeof(LinkEnt);
      const size_t budget = (size_t)32 << 30;
      uint32_t slots = (uint32_t)std::min<size_t>(std::max<size_t>(budget / per_slot, SIM_BLOCK), n_redo);
      slots = ((slots + SIM_BLOCK - 1) / SIM_BLOCK) * SIM_BLOCK;
      Scratch s2;
      rc = alloc_scratch(b, s2, slots, cap, b->scratch.delay_cap);
      if (rc) return rc;
      rc = launch(b, s2, b->redo_list.p, n_redo, slots / SIM_BLOCK);
      if (rc) return rc;
      CU(cudaStreamSynchronize(g_stream)); /* s2 is freed at the end of this scope */
    }
  }
  if (kernel_ms) {
    CU(cudaEventRecord(b->ev1, g_stream));
    CU(cudaEventSynchronize(b->ev1));
    CU(cudaEventElapsedTime(kernel_ms, b->ev0, b->ev1));
  }
  b->ran = true;
  return 0;
}

int remy_b200_batch_fetch(RemyBatch *b, RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!b || !out_sims) return fail(REMY_EINVAL, "batch/out_sims is NULL");
  if (!b->ran) return fail(REMY_EINVAL, "batch has not been run");
  if (out_senders && !b->want_senders) return fail(REMY_EINVAL, "batch was created without want_senders");
  if (out_use_counts && !b->want_counts) return fail(REMY_EINVAL, "batch was created without want_use_counts");
  CU(cudaSetDevice(g_device));
  CU(cudaMemcpyAsync(out_sims, b->out_sims.p, (size_t)b->n_sims * sizeof(RemySimResult), cudaMemcpyDeviceToHost, g_stream));
  CU(cudaStreamSynchronize(g_stream));
  if (out_senders)
    CU(cudaMemcpy(out_senders, b->out_senders.p, (size_t)b->n_sims * b->sender_stride * sizeof(RemySenderResult), cudaMemcpyDeviceToHost));
  if (out_use_counts)
    CU(cudaMemcpy(out_use_counts, b->out_counts.p, b->out_counts.n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return 0;
}

uint32_t remy_b200_batch_launches(const RemyBatch *b) { return b ? b->launches : 0; }

void remy_b200_batch_destroy(RemyBatch *b)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!b) return;
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  delete b;
}

int remy_b200_score_batch(const RemyFlatTree *tree,
                          const RemyLeafOverride *cands, uint32_t n_cands,
                          const RemyNetConfig *cfgs, const uint32_t *seeds, uint32_t n_cfgs,
                          uint32_t sender_stride,
                          RemySimResult *out_sims, RemySenderResult *out_senders, uint32_t *out_use_counts)
{
  if (!out_sims) return fail(REMY_EINVAL, "out_sims is NULL");
  RemyBatch *b = nullptr;
  int rc = remy_b200_batch_create(tree, cands, n_cands, cfgs, seeds, n_cfgs, sender_stride,
                                  out_senders != nullptr, out_use_counts != nullptr, &b);
  if (rc) return rc;
  rc = remy_b200_batch_run(b, nullptr);
  if (!rc) rc = remy_b200_batch_fetch(b, out_sims, out_senders, out_use_counts);
  remy_b200_batch_destroy(b);
  return rc;
}

} /* extern "C" */
