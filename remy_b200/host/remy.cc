/*
 * remy.cc — implementation of remy.hh (host mirror of the reference's scoring API).
 * file:line citations are relative to /root/reference.
 */
#include "remy.hh"
#include "pbwire.hh"

#include <algorithm>
#include <cassert>
#include <cstdio>
#include <ctime>
#include <stdexcept>
#include <thread>
#include <unistd.h>

namespace remy_b200 {

/* ------------------------------------------------------------------ Range -- */
std::string Range::DNA() const
{ /* src/configrange.cc:5-12 sets all three fields */
  pb::Writer w; w.put_double(61, low); w.put_double(62, high); w.put_double(63, incr); return w.str();
}
Range Range::parse(const void *data, size_t size)
{
  Range r(0, 0, 0); /* protobuf getters return 0 for absent fields (src/configrange.hh:25-29) */
  pb::Reader rd(data, size); pb::Field f;
  while (rd.next(f)) {
    if (f.wire_type != 1) continue;
    if (f.number == 61) r.low = pb::Reader::as_double(f);
    else if (f.number == 62) r.high = pb::Reader::as_double(f);
    else if (f.number == 63) r.incr = pb::Reader::as_double(f);
  }
  return r;
}

/* ------------------------------------------------------------ ConfigRange -- */
ConfigRange ConfigRange::parse(const void *data, size_t size)
{ /* src/configrange.cc:25-36; a field that is absent reads as Range(0,0,0) / 0 like protobuf */
  ConfigRange c;
  c.link_ppt = c.rtt = c.mean_on_duration = c.mean_off_duration = c.num_senders = c.buffer_size = Range(0, 0, 0);
  c.simulation_ticks = 0;
  pb::Reader rd(data, size); pb::Field f;
  while (rd.next(f)) {
    if (f.number == 77) {
      if (f.wire_type == 0) c.simulation_ticks = (unsigned int)f.varint;              /* ConfigRange */
      else if (f.wire_type == 2) c.simulation_ticks = (unsigned int)Range::parse(f.data, f.size).low; /* ConfigRangeUnicorn */
      continue;
    }
    if (f.wire_type != 2) continue;
    const Range r = Range::parse(f.data, f.size);
    switch (f.number) {
      case 71: c.link_ppt = r; break;
      case 72: c.rtt = r; break;
      case 73: c.num_senders = r; break;
      case 74: c.buffer_size = r; break;
      case 75: c.mean_off_duration = r; break;
      case 76: c.mean_on_duration = r; break;
      case 78: c.stochastic_loss_rate = r; break;
      default: break; /* 79-82: ConfigRangeUnicorn extras, not used on the Rat path */
    }
  }
  return c;
}
std::string ConfigRange::DNA() const
{ /* src/configrange.cc:38-50, ascending field numbers */
  pb::Writer w;
  w.put_message(71, link_ppt.DNA()); w.put_message(72, rtt.DNA()); w.put_message(73, num_senders.DNA());
  w.put_message(74, buffer_size.DNA()); w.put_message(75, mean_off_duration.DNA()); w.put_message(76, mean_on_duration.DNA());
  w.put_varint(77, simulation_ticks); w.put_message(78, stochastic_loss_rate.DNA());
  return w.str();
}

/* -------------------------------------------------------------- NetConfig -- */
std::string NetConfig::str() const
{ /* src/network.hh:72-78 */
  char tmp[256];
  snprintf(tmp, 256, "mean_on=%f, mean_off=%f, nsrc=%f, link_ppt=%f, delay=%f, buffer_size=%f, stochastic_loss_rate = %f\n",
           mean_on_duration, mean_off_duration, num_senders, link_ppt, delay, buffer_size, stochastic_loss_rate);
  return tmp;
}
std::string NetConfig::DNA() const
{ /* src/network.hh:57-70: all eight fields are set; num_senders / buffer_size / simulation_ticks are uint32 */
  pb::Writer w;
  w.put_double(1, mean_on_duration); w.put_double(2, mean_off_duration); w.put_varint(3, (uint32_t)num_senders);
  w.put_double(4, link_ppt); w.put_double(5, delay); w.put_varint(6, (uint32_t)buffer_size);
  w.put_double(7, stochastic_loss_rate); w.put_varint(8, (uint32_t)simulation_ticks);
  return w.str();
}
NetConfig NetConfig::parse(const void *data, size_t size)
{ /* NetConfig(const RemyBuffers::NetConfig &) (src/network.hh:38-47): absent fields read as 0 */
  NetConfig c;
  c.mean_on_duration = c.mean_off_duration = c.num_senders = c.link_ppt = c.delay = c.buffer_size = c.stochastic_loss_rate = c.simulation_ticks = 0;
  pb::Reader rd(data, size); pb::Field f;
  while (rd.next(f)) {
    switch (f.number) {
      case 1: if (f.wire_type == 1) c.mean_on_duration = pb::Reader::as_double(f); break;
      case 2: if (f.wire_type == 1) c.mean_off_duration = pb::Reader::as_double(f); break;
      case 3: if (f.wire_type == 0) c.num_senders = (uint32_t)f.varint; break;
      case 4: if (f.wire_type == 1) c.link_ppt = pb::Reader::as_double(f); break;
      case 5: if (f.wire_type == 1) c.delay = pb::Reader::as_double(f); break;
      case 6: if (f.wire_type == 0) c.buffer_size = (uint32_t)f.varint; break;
      case 7: if (f.wire_type == 1) c.stochastic_loss_rate = pb::Reader::as_double(f); break;
      case 8: if (f.wire_type == 0) c.simulation_ticks = (uint32_t)f.varint; break;
      default: break;
    }
  }
  return c;
}
RemyNetConfig NetConfig::pod(double ticks_to_run) const
{
  RemyNetConfig c;
  c.mean_on_duration = mean_on_duration; c.mean_off_duration = mean_off_duration; c.link_ppt = link_ppt;
  c.delay = delay; c.stochastic_loss_rate = stochastic_loss_rate; c.simulation_ticks = ticks_to_run;
  c.num_senders = (unsigned int)num_senders;   /* SenderGang ctor takes unsigned int (src/sendergang.hh:131-136) */
  c.buffer_size = (unsigned int)buffer_size;   /* Link ctor takes unsigned int (src/link.hh:22-24) */
  return c;
}

/* ----------------------------------------------------------------- Memory -- */
const Memory &MAX_MEMORY()
{ /* src/memory.cc:115-119 */
  static Memory m = [] { Memory x; for (unsigned i = 0; i < Memory::datasize; i++) x.f[i] = 163840; return x; }();
  return m;
}
std::string Memory::str() const
{ /* src/memory.cc:82-87 */
  char tmp[256];
  snprintf(tmp, 256, "sewma=%f, rewma=%f, rttr=%f, slowrewma=%f, rttd=%f, qdelay=%f", f[0], f[1], f[2], f[3], f[4], f[5]);
  return tmp;
}
std::string Memory::str(unsigned int num) const
{ /* src/memory.cc:89-113 */
  static const char *names[] = { "sewma", "rewma", "rttr", "slowrewma", "rttd", "qdelay" };
  char tmp[50] = "";
  if (num < datasize) snprintf(tmp, 50, "%s=%f ", names[num], f[num]);
  return tmp;
}

/* ------------------------------------------------------------ MemoryRange -- */
P2Median::P2Median() : _cnt(0)
{ /* p_square_quantile_impl's constructor for p = 0.5 */
  const double p = 0.5;
  for (int i = 0; i < 5; i++) { _h[i] = 0; _n[i] = i + 1.; }
  _d[0] = 1.; _d[1] = 1. + 2. * p; _d[2] = 1. + 4. * p; _d[3] = 3. + 2. * p; _d[4] = 5.;
}
void P2Median::operator()(double x)
{
  static const double incr[5] = { 0., 0.5 / 2., 0.5, (1. + 0.5) / 2., 1. };
  if (++_cnt <= 5) {
    _h[_cnt - 1] = x;
    if (_cnt == 5) std::sort(_h, _h + 5);
    return;
  }
  size_t k;
  if (x < _h[0]) { _h[0] = x; k = 1; }
  else if (_h[4] <= x) { _h[4] = x; k = 4; }
  else k = (size_t)(std::upper_bound(_h, _h + 5, x) - _h);
  for (size_t i = k; i < 5; i++) ++_n[i];
  for (size_t i = 0; i < 5; i++) _d[i] += incr[i];
  for (size_t i = 1; i <= 3; i++) {
    const double d = _d[i] - _n[i], dp = _n[i + 1] - _n[i], dm = _n[i - 1] - _n[i];
    const double hp = (_h[i + 1] - _h[i]) / dp, hm = (_h[i - 1] - _h[i]) / dm;
    if ((d >= 1. && dp > 1.) || (d <= -1. && dm < -1.)) {
      const short sign_d = static_cast<short>(d / std::abs(d));
      const double h = _h[i] + sign_d / (dp - dm) * ((sign_d - dm) * hp + (dp - sign_d) * hm);
      if (_h[i - 1] < h && h < _h[i + 1]) _h[i] = h;
      else { if (d > 0) _h[i] += hp; if (d < 0) _h[i] -= hm; }
      _n[i] += sign_d;
    }
  }
}
void MemoryRange::track(const Memory &q) const
{ for (int i : _active_axis) _acc[i](q.field(i)); }
std::vector<MemoryRange> MemoryRange::bisect() const
{ /* src/memoryrange.cc:8-41 */
  std::vector<MemoryRange> ret { *this };
  for (int i : _active_axis) {
    std::vector<MemoryRange> doubled;
    for (const auto &x : ret) {
      Memory ersatz_lower(x._lower), ersatz_upper(x._upper);
      ersatz_lower.mutable_field(i) = ersatz_upper.mutable_field(i) = _acc[i].median();
      if (x._lower == ersatz_upper) /* try range midpoint instead */
        ersatz_lower.mutable_field(i) = ersatz_upper.mutable_field(i) = (x._lower.field(i) + x._upper.field(i)) / 2;
      if (x._lower == ersatz_upper) {
        if (ersatz_lower == x._upper || !(x._lower == ersatz_lower)) throw std::runtime_error("MemoryRange::bisect: degenerate range (reference assert, memoryrange.cc:26-27)");
        doubled.push_back(x); /* cannot double on this axis */
      } else {
        doubled.emplace_back(x._lower, ersatz_upper, x._active_axis);
        doubled.emplace_back(ersatz_lower, x._upper, x._active_axis);
      }
    }
    ret = doubled;
  }
  return ret;
}
bool MemoryRange::contains(const Memory &q) const
{ /* src/memoryrange.cc:52-58 */
  for (int i : _active_axis)
    if (!((q.field(i) >= _lower.field(i)) && (q.field(i) < _upper.field(i)))) return false;
  return true;
}
Memory MemoryRange::range_median() const
{ /* src/memoryrange.cc:43-50 */
  Memory m(_lower);
  for (int i : _active_axis) m.mutable_field(i) = (_lower.field(i) + _upper.field(i)) / 2;
  return m;
}
bool MemoryRange::operator==(const MemoryRange &o) const
{ /* src/memoryrange.cc:68-74 */
  for (int i : _active_axis)
    if (!((_lower.field(i) == o._lower.field(i)) && (_upper.field(i) == o._upper.field(i)))) return false;
  return true;
}
std::string MemoryRange::str() const
{ /* src/memoryrange.cc:76-89 */
  std::string s = "(lo=< ";
  for (int i : _active_axis) s += _lower.str(i);
  s += ">, hi=< ";
  for (int i : _active_axis) s += _upper.str(i);
  s += ">)";
  return s;
}
static std::string memory_dna(const Memory &m, const bool *has)
{
  pb::Writer w;
  for (unsigned i = 0; i < Memory::datasize; i++) if (has[i]) w.put_double(21 + i, m.f[i]);
  return w.str();
}
std::string MemoryRange::DNA() const
{ /* src/memoryrange.cc:91-102 */
  pb::Writer w;
  w.put_message(11, memory_dna(_lower, _lower_has));
  w.put_message(12, memory_dna(_upper, _upper_has));
  if (_axes_explicit) for (int a : _active_axis) w.put_varint(13, (uint64_t)a);
  return w.str();
}
MemoryRange MemoryRange::parse(const void *data, size_t size)
{ /* src/memoryrange.cc:104-119 with Memory(bool, dna) wildcards (src/memory.cc:134-143) */
  MemoryRange r;
  for (unsigned i = 0; i < Memory::datasize; i++) { r._lower.f[i] = 0; r._upper.f[i] = 163840; r._lower_has[i] = r._upper_has[i] = false; }
  r._active_axis.clear();
  pb::Reader rd(data, size); pb::Field f;
  while (rd.next(f)) {
    if ((f.number == 11 || f.number == 12) && f.wire_type == 2) {
      Memory &m = f.number == 11 ? r._lower : r._upper;
      bool *has = f.number == 11 ? r._lower_has : r._upper_has;
      pb::Reader mr(f.data, f.size); pb::Field g;
      while (mr.next(g))
        if (g.number >= 21 && g.number <= 26 && g.wire_type == 1) { m.f[g.number - 21] = pb::Reader::as_double(g); has[g.number - 21] = true; }
    } else if (f.number == 13) {
      if (f.wire_type == 0) r._active_axis.push_back((int)f.varint);
      else if (f.wire_type == 2) { pb::Reader ar(f.data, f.size); while (!ar.done()) r._active_axis.push_back((int)ar.varint()); }
    }
  }
  r._axes_explicit = !r._active_axis.empty();
  if (r._active_axis.empty()) r._active_axis = { 0, 1, 2, 3 }; /* old remys: default axes */
  for (int a : r._active_axis) if (a < 0 || a >= (int)Memory::datasize) throw std::runtime_error("MemoryRange: unknown axis");
  return r;
}

/* ---------------------------------------------------------------- Whisker -- */
template <typename T> std::vector<T> OptimizationSetting<T>::alternatives(const T &value, bool active) const
{ /* src/action.hh:62-91 */
  if (!eligible_value(value))
    throw std::runtime_error("Ineligible value: " + std::to_string(value) + " is not between " + std::to_string(min_value) + " and " + std::to_string(max_value));
  std::vector<T> ret(1, value);
  if (!active) return ret;
  for (T proposed_change = min_change; proposed_change <= max_change; proposed_change *= multiplier) {
    const T up = value + proposed_change, down = value - proposed_change;
    if (eligible_value(up)) ret.push_back(up);
    if (eligible_value(down)) ret.push_back(down);
  }
  return ret;
}
template struct OptimizationSetting<unsigned int>;
template struct OptimizationSetting<double>;

const Whisker::OptimizationSettings &Whisker::get_optimizer()
{ /* src/whisker.hh:59-66 */
  static OptimizationSettings s { { 0, 256, 1, 32, 4, 1 }, { 0, 1, 0.01, 0.5, 4, 1 }, { 0.25, 3, 0.05, 1, 4, 3 } };
  return s;
}
Whisker::Whisker(const MemoryRange &dom)
  : _domain(dom), _window_increment((int)get_optimizer().window_increment.default_value),
    _window_multiple(get_optimizer().window_multiple.default_value), _intersend(get_optimizer().intersend.default_value) {}
unsigned int Whisker::window(unsigned int previous_window) const
{ return std::min(std::max(0, int(previous_window * _window_multiple + _window_increment)), 1000000); }
void Whisker::round()
{ /* src/whisker.cc:91-95 */
  _window_multiple = (1.0 / 10000.0) * int(10000 * _window_multiple);
  _intersend = (1.0 / 10000.0) * int(10000 * _intersend);
}
std::vector<Whisker> Whisker::bisect() const
{ /* Action::bisect<Whisker>, src/action.hh:25-33 */
  std::vector<Whisker> ret;
  for (const auto &x : _domain.bisect()) { Whisker new_action(*this); new_action._domain = x; ret.push_back(new_action); }
  return ret;
}
std::vector<Whisker> Whisker::next_generation(bool oi, bool om, bool os) const
{ /* src/whisker.cc:46-81 (without its printf) */
  std::vector<Whisker> ret;
  const auto wi = get_optimizer().window_increment.alternatives((unsigned int)_window_increment, oi);
  const auto wm = get_optimizer().window_multiple.alternatives(_window_multiple, om);
  const auto is = get_optimizer().intersend.alternatives(_intersend, os);
  for (const auto &a : wi) for (const auto &b : wm) for (const auto &c : is) {
    Whisker w(*this);
    w._generation++;
    w._window_increment = (int)a; w._window_multiple = b; w._intersend = c;
    w.round();
    ret.push_back(w);
  }
  return ret;
}
std::string Whisker::str(unsigned int total) const
{ /* src/whisker.cc:83-89 */
  char tmp[512];
  snprintf(tmp, 512, "{%s} gen=%u usage=%.4f => (win=%d + %f * win, intersend=%f)", _domain.str().c_str(), _generation,
           double(_domain.count()) / double(total), _window_increment, _window_multiple, _intersend);
  return tmp;
}
std::string Whisker::DNA() const
{ /* src/whisker.cc:34-44 */
  pb::Writer w;
  w.put_sint32(31, _window_increment); w.put_double(32, _window_multiple); w.put_double(33, _intersend);
  w.put_message(34, _domain.DNA());
  return w.str();
}
Whisker Whisker::parse(const void *data, size_t size)
{ /* src/whisker.cc:26-32 */
  Whisker w; w._window_increment = 0; w._window_multiple = 0; w._intersend = 0;
  pb::Reader rd(data, size); pb::Field f;
  while (rd.next(f)) {
    if (f.number == 31 && f.wire_type == 0) w._window_increment = pb::Reader::as_sint32(f);
    else if (f.number == 32 && f.wire_type == 1) w._window_multiple = pb::Reader::as_double(f);
    else if (f.number == 33 && f.wire_type == 1) w._intersend = pb::Reader::as_double(f);
    else if (f.number == 34 && f.wire_type == 2) w._domain = MemoryRange::parse(f.data, f.size);
  }
  return w;
}

/* ------------------------------------------------------------ WhiskerTree -- */
WhiskerTree::WhiskerTree() : _domain(Memory(), MAX_MEMORY()), _leaf(1, Whisker(_domain)) {} /* src/whiskertree.cc:10-15 */

const Whisker *WhiskerTree::whisker(const Memory &m) const
{ /* src/whiskertree.cc:62-82 */
  if (!_domain.contains(m)) return nullptr;
  if (is_leaf()) return &_leaf[0];
  for (const auto &x : _children) { const Whisker *r = x.whisker(m); if (r) return r; }
  throw std::runtime_error("WhiskerTree: no child contains the query (reference: assert(false), whiskertree.cc:80)");
}
const Whisker &WhiskerTree::use_whisker(const Memory &m, bool track) const
{ /* src/whiskertree.cc:42-60 */
  const Whisker *ret = whisker(m);
  if (!ret) throw std::runtime_error("ERROR: No whisker found for " + m.str());
  ret->domain().use();
  if (track) ret->domain().track(m);
  return *ret;
}
WhiskerTree::WhiskerTree(const Whisker &whisker, bool bisect) : _domain(whisker.domain())
{ /* src/whiskertree.cc:17-30 */
  if (!bisect) _leaf.push_back(whisker);
  else for (const auto &x : whisker.bisect()) _children.push_back(WhiskerTree(x, false));
}
bool WhiskerTree::replace(const Whisker &src, const WhiskerTree &dst)
{ /* src/whiskertree.cc:159-180 */
  if (!_domain.contains(src.domain().range_median())) return false;
  if (is_leaf()) {
    if (!(src.domain() == _leaf.front().domain())) throw std::runtime_error("WhiskerTree::replace: domain mismatch (reference assert)");
    const std::string extra = _extra;
    *this = dst; /* convert from leaf to interior node */
    _extra = extra;
    return true;
  }
  for (auto &x : _children) if (x.replace(src, dst)) return true;
  throw std::runtime_error("WhiskerTree::replace: no child took the whisker (reference assert)");
}
void WhiskerTree::reset_counts()
{ if (is_leaf()) _leaf.front().domain().reset_count(); else for (auto &x : _children) x.reset_counts(); }
void WhiskerTree::promote(unsigned int g)
{ if (is_leaf()) _leaf.front().promote(g); else for (auto &x : _children) x.promote(g); }
void WhiskerTree::reset_generation()
{ if (is_leaf()) _leaf.front().demote(0); else for (auto &x : _children) x.reset_generation(); }
const Whisker *WhiskerTree::most_used(unsigned int max_generation) const
{ /* src/whiskertree.cc:84-109 */
  if (is_leaf()) {
    if ((_leaf.front().generation() <= max_generation) && (_leaf.front().count() > 0)) return &_leaf[0];
    return nullptr;
  }
  unsigned int count_max = 0;
  const Whisker *ret = nullptr;
  for (const auto &x : _children) {
    const Whisker *c = x.most_used(max_generation);
    if (c && (c->generation() <= max_generation) && (c->count() >= count_max)) { ret = c; count_max = c->count(); }
  }
  return ret;
}
bool WhiskerTree::replace(const Whisker &w)
{ /* src/whiskertree.cc:137-157 */
  if (!_domain.contains(w.domain().range_median())) return false;
  if (is_leaf()) {
    if (!(w.domain() == _leaf.front().domain())) throw std::runtime_error("WhiskerTree::replace: domain mismatch (reference assert)");
    _leaf.front() = w;
    return true;
  }
  for (auto &x : _children) if (x.replace(w)) return true;
  throw std::runtime_error("WhiskerTree::replace: no child took the whisker (reference assert)");
}
unsigned int WhiskerTree::total_whisker_queries() const
{
  if (is_leaf()) return _leaf.front().domain().count();
  unsigned int s = 0;
  for (const auto &x : _children) s += x.total_whisker_queries();
  return s;
}
std::string WhiskerTree::str() const { return str(total_whisker_queries()); }
std::string WhiskerTree::str(unsigned int total) const
{ /* src/whiskertree.cc:196-212 */
  if (is_leaf()) return std::string("[") + _leaf.front().str(total) + "]";
  std::string ret;
  for (const auto &x : _children) ret += x.str(total);
  return ret;
}
void WhiskerTree::leaves(std::vector<const Whisker *> &out) const
{ if (is_leaf()) out.push_back(&_leaf[0]); else for (const auto &x : _children) x.leaves(out); }
void WhiskerTree::leaves(std::vector<Whisker *> &out)
{ if (is_leaf()) out.push_back(&_leaf[0]); else for (auto &x : _children) x.leaves(out); }
int WhiskerTree::leaf_index_of(const Whisker &w) const
{
  /* follow WhiskerTree::replace's descent; count leaves passed on the way */
  struct R { static bool go(const WhiskerTree &t, const Memory &q, const Whisker &w, int &idx) {
    if (!t._domain.contains(q)) { std::vector<const Whisker *> v; t.leaves(v); idx += (int)v.size(); return false; }
    if (t.is_leaf()) { if (w.domain() == t._leaf.front().domain()) return true; idx++; return false; }
    for (const auto &x : t._children) if (go(x, q, w, idx)) return true;
    return false; } };
  int idx = 0;
  return R::go(*this, w.domain().range_median(), w, idx) ? idx : -1;
}
std::string WhiskerTree::DNA() const
{ /* src/whiskertree.cc:234-252 */
  pb::Writer w;
  w.put_message(1, _domain.DNA());
  if (is_leaf()) w.put_message(3, _leaf.front().DNA());
  else for (const auto &x : _children) w.put_message(2, x.DNA());
  std::string s = w.str();
  s += _extra;
  return s;
}
WhiskerTree WhiskerTree::parse(const void *data, size_t size)
{ /* src/whiskertree.cc:254-268 */
  WhiskerTree t; t._leaf.clear();
  pb::Reader rd(data, size); pb::Field f;
  bool has_leaf = false;
  while (rd.next(f)) {
    if (f.number == 1 && f.wire_type == 2) t._domain = MemoryRange::parse(f.data, f.size);
    else if (f.number == 2 && f.wire_type == 2) t._children.push_back(parse(f.data, f.size));
    else if (f.number == 3 && f.wire_type == 2) { t._leaf.assign(1, Whisker::parse(f.data, f.size)); has_leaf = true; }
    else if (f.number >= 4 && f.number <= 6 && f.wire_type == 2) {
      pb::Writer w; w.put_message(f.number, std::string((const char *)f.data, f.size)); t._extra += w.str();
    }
  }
  if (has_leaf && !t._children.empty()) throw std::runtime_error("WhiskerTree: node has both a leaf and children");
  if (!has_leaf && t._children.empty()) throw std::runtime_error("WhiskerTree: node has neither a leaf nor children");
  return t;
}

FlatTree flatten(const WhiskerTree &root)
{
  FlatTree f;
  std::vector<const Whisker *> dfs;
  root.leaves(dfs);
  std::vector<const WhiskerTree *> bfs { &root };
  for (size_t i = 0; i < bfs.size(); i++) { /* breadth-first: the children of a node are contiguous */
    const WhiskerTree *t = bfs[i];
    RemyTreeNode n;
    for (unsigned a = 0; a < REMY_NUM_AXES; a++) { n.lower[a] = t->_domain._lower.f[a]; n.upper[a] = t->_domain._upper.f[a]; }
    n.axis_mask = t->_domain.axis_mask(); n.child_begin = 0; n.child_count = 0; n.leaf_index = 0;
    if (t->is_leaf()) {
      n.leaf_index = (uint32_t)(std::find(dfs.begin(), dfs.end(), &t->_leaf[0]) - dfs.begin());
    } else {
      n.child_begin = (uint32_t)bfs.size(); n.child_count = (uint32_t)t->_children.size();
      for (const auto &c : t->_children) bfs.push_back(&c);
    }
    f.nodes.push_back(n);
  }
  for (const Whisker *w : dfs) {
    RemyLeafAction a; a.window_multiple = w->window_multiple(); a.intersend = w->intersend();
    a.window_increment = w->window_increment(); a.generation = w->generation();
    f.leaves.push_back(a);
  }
  return f;
}

/* -------------------------------------------------------------- Evaluator -- */
static bool g_flow_switched = false;
void set_flow_switched(bool on) { g_flow_switched = on; }
bool flow_switched() { return g_flow_switched; }

static bool g_chained_prng = false;
void set_chained_prng(bool on) { g_chained_prng = on; }
bool chained_prng() { return g_chained_prng; }

static uint32_t lcg(uint32_t x) { return (uint32_t)(((uint64_t)x * 16807u) % 2147483647u); }

std::vector<uint32_t> Evaluator<WhiskerTree>::seeds(unsigned int prng_seed, size_t n)
{
  /* config 0 runs on PRNG(prng_seed) itself — exactly what the reference does (src/evaluator.cc:84),
   * so a single-config score() is bit-identical to the reference's; config k > 0 gets the k-th
   * output of that engine instead of wherever config k-1 happened to leave it */
  std::vector<uint32_t> s(n);
  uint32_t x = prng_seed % 2147483647u;
  if (x == 0) x = 1;
  for (size_t i = 0; i < n; i++) {
    if (i == 0) { s[0] = prng_seed; continue; }
    x = lcg(x); s[i] = x;
  }
  return s;
}

Evaluator<WhiskerTree>::Evaluator(const ConfigRange &range, unsigned int prng_seed)
  : _prng_seed(prng_seed ? prng_seed : (unsigned int)(time(NULL) ^ getpid())), _tick_count(range.simulation_ticks), _configs()
{
  /* every point in the cube of configs: the loops of src/evaluator.cc:16-37, verbatim in
   * structure because the floating-point loop counters decide which points exist */
  for (double link_ppt = range.link_ppt.low; link_ppt <= range.link_ppt.high; link_ppt += range.link_ppt.incr) {
    for (double rtt = range.rtt.low; rtt <= range.rtt.high; rtt += range.rtt.incr) {
      for (unsigned int senders = range.num_senders.low; senders <= range.num_senders.high; senders += range.num_senders.incr) {
        for (double on = range.mean_on_duration.low; on <= range.mean_on_duration.high; on += range.mean_on_duration.incr) {
          for (double off = range.mean_off_duration.low; off <= range.mean_off_duration.high; off += range.mean_off_duration.incr) {
            for (double buffer_size = range.buffer_size.low; buffer_size <= range.buffer_size.high; buffer_size += range.buffer_size.incr) {
              for (double loss_rate = range.stochastic_loss_rate.low; loss_rate <= range.stochastic_loss_rate.high; loss_rate += range.stochastic_loss_rate.incr) {
                _configs.push_back(NetConfig().set_link_ppt(link_ppt).set_delay(rtt).set_num_senders(senders).set_on_duration(on)
                                     .set_off_duration(off).set_buffer_size(buffer_size).set_stochastic_loss_rate(loss_rate));
                if (range.stochastic_loss_rate.isOne()) break;
              }
              if (range.buffer_size.isOne()) break;
            }
            if (range.mean_off_duration.isOne()) break;
          }
          if (range.mean_on_duration.isOne()) break;
        }
        if (range.num_senders.isOne()) break;
      }
      if (range.rtt.isOne()) break;
    }
    if (range.link_ppt.isOne()) break;
  }
}

static void check_sim(const RemySimResult &r)
{
  if (r.error & REMY_ERR_NO_WHISKER) throw std::runtime_error("ERROR: No whisker found (reference: exit(1), whiskertree.cc:46-49)");
  if (r.error) throw std::runtime_error("simulation failed with error word " + std::to_string(r.error));
}

namespace {
/* What remy_b200_score_trace's records are fed into: MemoryRange::track of the leaf each lookup
 * returned (src/whiskertree.cc:55-57), in the order the device reports them (config by config,
 * and within a config in the reference's call order). */
struct TrackSink {
  std::vector<const MemoryRange *> leaves; /* the domain of every leaf, depth-first */
  std::vector<int> axes; /* axis of record word 1, 2, ... */
  std::vector<uint32_t> leaf_axes; /* active-axis mask of every leaf */
  /* One estimator per (leaf, axis), each fed in record order: the pairs are independent of each
   * other, so a chunk is fed by several host threads, each owning a fixed subset of the pairs. */
  void feed_range(const uint64_t *rec, uint64_t n, uint32_t words, unsigned part, unsigned parts) const
  {
    const size_t na = axes.size();
    for (uint64_t i = 0; i < n; i++, rec += words) {
      const uint64_t l = rec[0];
      const uint32_t mask = leaf_axes[l];
      for (size_t k = 0; k < na; k++) {
        if (!(mask & (1u << axes[k])) || (l * na + k) % parts != part) continue;
        double v; memcpy(&v, &rec[1 + k], sizeof v);
        leaves[l]->_acc[axes[k]](v);
      }
    }
  }
  static int feed(void *user, uint32_t, const uint64_t *rec, uint64_t n, uint32_t words)
  {
    TrackSink *t = static_cast<TrackSink *>(user);
    if (words != t->axes.size() + 1) return 1;
    for (uint64_t i = 0; i < n; i++) if (rec[i * words] >= t->leaves.size()) return 1;
    unsigned parts = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (n < 65536) parts = 1;
    if (parts == 1) { t->feed_range(rec, n, words, 0, 1); return 0; }
    std::vector<std::thread> th;
    for (unsigned p = 0; p < parts; p++) th.emplace_back([=]() { t->feed_range(rec, n, words, p, parts); });
    for (auto &x : th) x.join();
    return 0;
  }
};
TrackSink make_sink(const std::vector<const MemoryRange *> &leaf_domains, uint32_t all_axes_mask)
{
  TrackSink sink;
  sink.leaves = leaf_domains;
  for (const MemoryRange *d : leaf_domains) sink.leaf_axes.push_back(d->axis_mask());
  for (int a = 0; a < REMY_NUM_AXES; a++) if (all_axes_mask & (1u << a)) sink.axes.push_back(a);
  return sink;
}
TrackSink make_sink(WhiskerTree &tree)
{
  std::vector<const Whisker *> lv;
  static_cast<const WhiskerTree &>(tree).leaves(lv);
  std::vector<const MemoryRange *> doms;
  for (const Whisker *w : lv) doms.push_back(&w->domain());
  uint32_t mask = 0;
  std::vector<const WhiskerTree *> stack { &tree }; /* the union over ALL nodes, as in the device layout */
  while (!stack.empty()) { const WhiskerTree *t = stack.back(); stack.pop_back(); mask |= t->_domain.axis_mask(); for (const auto &c : t->_children) stack.push_back(&c); }
  return make_sink(doms, mask);
}
} // namespace

int feed_trace_records(WhiskerTree &tree, const uint64_t *records, uint64_t n_records, uint32_t words_per_record)
{
  TrackSink sink = make_sink(tree);
  return TrackSink::feed(&sink, 0, records, n_records, words_per_record);
}

int feed_trace_records(const std::vector<const MemoryRange *> &leaf_domains, uint32_t all_axes_mask,
                       const uint64_t *records, uint64_t n_records, uint32_t words_per_record)
{
  TrackSink sink = make_sink(leaf_domains, all_axes_mask);
  return TrackSink::feed(&sink, 0, records, n_records, words_per_record);
}

Evaluator<WhiskerTree>::Outcome Evaluator<WhiskerTree>::score(WhiskerTree &run_whiskers, unsigned int prng_seed,
    const std::vector<NetConfig> &configs, bool trace, unsigned int ticks_to_run)
{
  const FlatTree ft = flatten(run_whiskers);
  const RemyFlatTree tree = ft.c();
  std::vector<RemyNetConfig> cfgs;
  uint32_t stride = 1;
  for (const auto &x : configs) { cfgs.push_back(x.pod((double)ticks_to_run)); stride = std::max(stride, cfgs.back().num_senders); }
  const std::vector<uint32_t> sd = seeds(prng_seed, cfgs.size());
  std::vector<RemySimResult> sims(cfgs.size());
  std::vector<RemySenderResult> snd(cfgs.size() * (size_t)stride);
  std::vector<uint32_t> counts(ft.leaves.size());
  if (trace && g_flow_switched) throw std::runtime_error("trace == true scoring is not available for flow-switched senders");
  if (trace) { /* Rat(run_whiskers, trace) (src/evaluator.cc:93): every lookup also feeds the leaf's running medians */
    TrackSink sink = make_sink(run_whiskers);
    if (remy_b200_score_trace(&tree, cfgs.data(), sd.data(), (uint32_t)cfgs.size(), stride,
                              sims.data(), snd.data(), counts.data(), TrackSink::feed, &sink))
      throw std::runtime_error(std::string("remy_b200_score_trace: ") + remy_b200_last_error());
  } else if (g_chained_prng) {
    /* the reference's one PRNG (src/evaluator.cc:84): config k is seeded with the state config k-1 left behind
     * (a minstd_rand0 state in [1, 2^31-2] seeds the engine to exactly that state) */
    uint32_t seed = prng_seed;
    std::vector<uint32_t> c1(ft.leaves.size());
    for (size_t k = 0; k < cfgs.size(); k++) {
      if ((g_flow_switched ? remy_b200_score_batch_flows : remy_b200_score_batch)(&tree, nullptr, 1, &cfgs[k], &seed, 1, stride,
                                                                                   &sims[k], &snd[k * (size_t)stride], c1.data()))
        throw std::runtime_error(std::string("remy_b200_score_batch: ") + remy_b200_last_error());
      for (size_t l = 0; l < c1.size(); l++) counts[l] += c1[l];
      seed = sims[k].prng_state;
    }
  } else if ((g_flow_switched ? remy_b200_score_batch_flows : remy_b200_score_batch)(&tree, nullptr, 1, cfgs.data(), sd.data(), (uint32_t)cfgs.size(), stride,
                            sims.data(), snd.data(), counts.data()))
    throw std::runtime_error(std::string("remy_b200_score_batch: ") + remy_b200_last_error());
  Outcome out;
  for (size_t k = 0; k < cfgs.size(); k++) {
    check_sim(sims[k]);
    out.score += sims[k].utility;                                            /* src/evaluator.cc:96 */
    std::vector<std::pair<double, double>> td;
    for (uint32_t s = 0; s < cfgs[k].num_senders; s++) {                     /* src/utility.hh:30-44 */
      const RemySenderResult &r = snd[k * stride + s];
      td.emplace_back(r.tick_share_sending == 0 ? 0.0 : double(r.packets_received) / r.tick_share_sending,
                      r.packets_received == 0 ? 0.0 : double(r.total_delay) / double(r.packets_received));
    }
    out.throughputs_delays.emplace_back(configs[k], td);
  }
  /* reset_counts() + the run's uses (src/evaluator.cc:86,100): on the caller's tree and on the copy */
  std::vector<Whisker *> lv;
  run_whiskers.leaves(lv);
  for (size_t i = 0; i < lv.size(); i++) lv[i]->_domain._count = counts[i];
  out.used_actions = run_whiskers;
  return out;
}

Evaluator<WhiskerTree>::Outcome Evaluator<WhiskerTree>::score(WhiskerTree &run_actions, bool trace, double carefulness) const
{ /* src/evaluator.cc:196-201 */
  return score(run_actions, _prng_seed, _configs, trace, _tick_count * carefulness);
}

std::string Evaluator<WhiskerTree>::DNA(const WhiskerTree &whiskers) const
{ /* src/evaluator.cc:40-66 */
  pb::Writer settings; settings.put_varint(11, _prng_seed); settings.put_varint(12, _tick_count);
  pb::Writer w;
  w.put_message(1, settings.str());
  for (const auto &x : _configs) w.put_message(2, x.DNA());
  w.put_message(3, whiskers.DNA());
  return w.str();
}

Evaluator<WhiskerTree>::Outcome Evaluator<WhiskerTree>::parse_problem_and_evaluate(const void *problem, size_t size)
{ /* src/evaluator.cc:134-146 */
  std::vector<NetConfig> configs;
  WhiskerTree run_whiskers;
  unsigned int prng_seed = 0, tick_count = 0;
  bool have_tree = false;
  pb::Reader rd(problem, size); pb::Field f;
  while (rd.next(f)) {
    if (f.wire_type != 2) continue;
    if (f.number == 1) {
      pb::Reader sr(f.data, f.size); pb::Field g;
      while (sr.next(g)) { if (g.number == 11 && g.wire_type == 0) prng_seed = (unsigned int)g.varint; else if (g.number == 12 && g.wire_type == 0) tick_count = (unsigned int)g.varint; }
    } else if (f.number == 2) configs.push_back(NetConfig::parse(f.data, f.size));
    else if (f.number == 3) { run_whiskers = WhiskerTree::parse(f.data, f.size); have_tree = true; }
  }
  if (!have_tree) throw std::runtime_error("Problem has no whiskers");
  return score(run_whiskers, prng_seed, configs, false, tick_count);
}

std::string Evaluator<WhiskerTree>::Outcome::DNA() const
{ /* src/evaluator.cc:162-180 */
  pb::Writer w;
  for (const auto &run : throughputs_delays) {
    pb::Writer td;
    td.put_message(21, run.first.DNA());
    for (const auto &x : run.second) { pb::Writer r; r.put_double(11, x.first); r.put_double(12, x.second); td.put_message(22, r.str()); }
    w.put_message(32, td.str());
  }
  w.put_double(33, score);
  return w.str();
}

Evaluator<WhiskerTree>::Outcome Evaluator<WhiskerTree>::Outcome::parse(const void *data, size_t size)
{ /* src/evaluator.cc:182-194 */
  Outcome o;
  pb::Reader rd(data, size); pb::Field f;
  while (rd.next(f)) {
    if (f.number == 33 && f.wire_type == 1) o.score = pb::Reader::as_double(f);
    else if (f.number == 32 && f.wire_type == 2) {
      NetConfig cfg; std::vector<std::pair<double, double>> td;
      pb::Reader tr(f.data, f.size); pb::Field g;
      while (tr.next(g)) {
        if (g.number == 21 && g.wire_type == 2) cfg = NetConfig::parse(g.data, g.size);
        else if (g.number == 22 && g.wire_type == 2) {
          double tp = 0, dl = 0;
          pb::Reader rr(g.data, g.size); pb::Field h;
          while (rr.next(h)) { if (h.number == 11 && h.wire_type == 1) tp = pb::Reader::as_double(h); else if (h.number == 12 && h.wire_type == 1) dl = pb::Reader::as_double(h); }
          td.emplace_back(tp, dl);
        }
      }
      o.throughputs_delays.emplace_back(cfg, td);
    }
  }
  return o;
}

std::vector<double> Evaluator<WhiskerTree>::score_candidates(const WhiskerTree &base, const std::vector<Whisker> &replacements,
                                                             double carefulness) const
{
  const FlatTree ft = flatten(base);
  const RemyFlatTree tree = ft.c();
  const unsigned int ticks_to_run = _tick_count * carefulness;
  std::vector<RemyNetConfig> cfgs;
  for (const auto &x : _configs) cfgs.push_back(x.pod((double)ticks_to_run));
  const std::vector<uint32_t> sd = seeds(_prng_seed, cfgs.size());
  std::vector<RemyLeafOverride> cands;
  for (const auto &r : replacements) {
    const int leaf = base.leaf_index_of(r);
    if (leaf < 0) throw std::runtime_error("score_candidates: replacement matches no leaf (reference: assert(found_replacement), breeder.cc:66-67)");
    RemyLeafOverride o; o.window_multiple = r.window_multiple(); o.intersend = r.intersend();
    o.window_increment = r.window_increment(); o.leaf_index = (uint32_t)leaf;
    cands.push_back(o);
  }
  std::vector<RemySimResult> sims(cands.size() * cfgs.size());
  if (!cands.empty() && (g_flow_switched ? remy_b200_score_batch_flows : remy_b200_score_batch)(&tree, cands.data(), (uint32_t)cands.size(), cfgs.data(), sd.data(), (uint32_t)cfgs.size(), 0,
                                              sims.data(), nullptr, nullptr))
    throw std::runtime_error(std::string("remy_b200_score_batch: ") + remy_b200_last_error());
  std::vector<double> scores(cands.size(), 0.0);
  for (size_t c = 0; c < cands.size(); c++)
    for (size_t k = 0; k < cfgs.size(); k++) { check_sim(sims[c * cfgs.size() + k]); scores[c] += sims[c * cfgs.size() + k].utility; }
  return scores;
}

} // namespace remy_b200
