/*
 * pbwire.hh — minimal proto2 wire-format reader/writer (no protoc/libprotobuf in the image).
 * Enough for the reference's protobufs/dna.proto, problem.proto and answer.proto: varint,
 * 64-bit, length-delimited, 32-bit; unknown fields are skipped; writers emit fields in the
 * order the caller asks for (ascending field numbers give byte-identical files to protoc's).
 */
#ifndef REMY_HOST_PBWIRE_HH
#define REMY_HOST_PBWIRE_HH
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace remy_b200 {
namespace pb {

struct Field { uint32_t number; uint32_t wire_type; uint64_t varint; const uint8_t *data; size_t size; };

class Reader {
  const uint8_t *p_, *end_;
public:
  Reader(const void *data, size_t size) : p_((const uint8_t *)data), end_((const uint8_t *)data + size) {}
  explicit Reader(const std::string &s) : Reader(s.data(), s.size()) {}
  bool done() const { return p_ >= end_; }
  uint64_t varint()
  {
    uint64_t v = 0; int shift = 0;
    for (;;) {
      if (p_ >= end_ || shift > 63) throw std::runtime_error("protobuf: truncated varint");
      const uint8_t b = *p_++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
  }
  bool next(Field &f)
  {
    if (done()) return false;
    const uint64_t key = varint();
    f.number = (uint32_t)(key >> 3); f.wire_type = (uint32_t)(key & 7); f.varint = 0; f.data = nullptr; f.size = 0;
    switch (f.wire_type) {
      case 0: f.varint = varint(); break;
      case 1: need(8); f.data = p_; f.size = 8; p_ += 8; break;
      case 2: { const uint64_t n = varint(); need(n); f.data = p_; f.size = (size_t)n; p_ += n; break; }
      case 5: need(4); f.data = p_; f.size = 4; p_ += 4; break;
      default: throw std::runtime_error("protobuf: unsupported wire type");
    }
    return true;
  }
  static double as_double(const Field &f) { double d; if (f.size != 8) throw std::runtime_error("protobuf: expected 64-bit field"); memcpy(&d, f.data, 8); return d; }
  static int32_t as_sint32(const Field &f) { const uint32_t v = (uint32_t)f.varint; return (int32_t)((v >> 1) ^ (~(v & 1) + 1)); }
private:
  void need(uint64_t n) { if ((uint64_t)(end_ - p_) < n) throw std::runtime_error("protobuf: truncated field"); }
};

class Writer {
  std::string out_;
  void key(uint32_t number, uint32_t wt) { raw_varint(((uint64_t)number << 3) | wt); }
public:
  void raw_varint(uint64_t v) { while (v >= 0x80) { out_.push_back((char)(v | 0x80)); v >>= 7; } out_.push_back((char)v); }
  void put_double(uint32_t number, double d) { key(number, 1); char b[8]; memcpy(b, &d, 8); out_.append(b, 8); }
  void put_varint(uint32_t number, uint64_t v) { key(number, 0); raw_varint(v); }
  void put_sint32(uint32_t number, int32_t v) { put_varint(number, (uint32_t)((v << 1) ^ (v >> 31))); }
  void put_message(uint32_t number, const std::string &body) { key(number, 2); raw_varint(body.size()); out_ += body; }
  const std::string &str() const { return out_; }
};

} // namespace pb
} // namespace remy_b200
#endif
