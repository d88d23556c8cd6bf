/*
 * device_layout.h — data layout in HBM shared by the kernel (sim_kernel.cuh) and
 * the host code that fills it (batch_prep.h, capi.cu).  Plain structs only.
 */
#ifndef REMY_DEVICE_LAYOUT_H
#define REMY_DEVICE_LAYOUT_H

#include <stdint.h>
#include "../../include/remy_b200.h"

#if defined(__CUDACC__) || defined(REMY_HOST_EMU)
#define REMY_ALIGN(n) __align__(n)
#else
#define REMY_ALIGN(n) alignas(n)
#endif

namespace remy {

constexpr int NAX = REMY_NUM_AXES;

/* Per-(slot, sender) state that is touched only when that sender gets an ack,
 * transmits or switches: one 128-byte line in HBM/L2. */
struct REMY_ALIGN(16) SenderCold {
  double mem[NAX];          /* Memory::field(0..5) (reference src/memory.hh:97) */
  double last_tick_sent;    /* Memory::_last_tick_sent */
  double last_tick_received;
  double min_rtt;
  double total_delay;       /* Utility::_total_delay */
  double reserved;
  uint32_t leaf;            /* leaf returned by the sender's latest lookup */
  uint32_t pad;
  /* bytes 96..127: everything a transmission reads or writes, in one 32-byte sector */
  double last_send_time;    /* Rat::_last_send_time */
  double intersend;         /* Rat::_intersend_time */
  uint32_t packets_sent;    /* Rat::_packets_sent */
  int32_t largest_ack;      /* Rat::_largest_ack */
  int32_t window;           /* Rat::_the_window */
  uint32_t u_received;      /* Utility::_packets_received */
};
static_assert(sizeof(SenderCold) == 128, "SenderCold must be one 128-byte line");

/* What Rat::send needs (reference src/rat-templates.cc:20-34), mirrored out of SenderCold into a
 * 16-byte record: 0.5 KB per simulation, so this array stays L2-resident while the 128-byte cold
 * lines (4 KB per simulation) stream through HBM. */
struct REMY_ALIGN(16) SendState { uint32_t packets_sent; int32_t lim; double intersend; };
static_assert(sizeof(SendState) == 16, "SendState layout");

/* Link::_buffer entry and Delay::_queue entry (reference src/link.hh:15, src/delay.hh:16);
 * flow_id and tick_received are not stored: the first is dead on the Rat path, the second
 * is the tick at which the entry is consumed. */
struct REMY_ALIGN(16) LinkEnt { double tick_sent; uint32_t seq; uint32_t src; };
struct REMY_ALIGN(16) DelayEnt { double release; double tick_sent; uint32_t seq; uint32_t src; uint32_t pad[2]; };
static_assert(sizeof(LinkEnt) == 16 && sizeof(DelayEnt) == 32, "ring entry sizes");

/* Device copy of the tree: bounds[node][axis] = (lower, upper) with axes a node does not
 * test widened to (-inf, +inf); meta[node] = (child_begin or leaf_index, child_count). */
struct REMY_ALIGN(16) Bounds { double lo, hi; };
/* An interior node may carry a *cell table*: the distinct finite bounds of its children on
 * each axis ("cuts") split the axis into ncuts+1 cells; a query's cell index p_a = #{cuts <= m[a]}
 * decides every children's `lower <= m < upper` test at once, so table[cell] holds the first
 * child (in stored order) whose box contains that cell, or NO_CELL_CHILD.  This is exactly the
 * reference's first-match scan (src/whiskertree.cc:62-82) evaluated per cell on the host. */
constexpr uint32_t NO_TABLE = 0xFFFFFFFFu;
constexpr uint16_t NO_CELL_CHILD = 0xFFFFu;
struct DevNodeMeta {
  uint32_t first;       /* child_begin, or leaf_index when count == 0 */
  uint32_t count;       /* number of children */
  uint32_t table_off;   /* offset into the cell-table array, or NO_TABLE (linear scan) */
  uint32_t cuts_off;    /* offset into the cuts array; axes concatenated in axis order */
  uint8_t ncuts[8];     /* cuts per axis (0 for axes the tree never tests) */
};
static_assert(sizeof(DevNodeMeta) == 24, "DevNodeMeta layout");
struct REMY_ALIGN(8) DevLeaf { double window_multiple; double intersend; int32_t window_increment; uint32_t pad; };

} // namespace remy
#endif
