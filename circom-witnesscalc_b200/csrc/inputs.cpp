#include "inputs.hpp"

#include <string.h>

namespace gw {

namespace {
struct J {
  const char* p; const char* end;
  void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
  [[noreturn]] void fail(const char* what) { throw Error(std::string("Failed to parse inputs: invalid JSON: ") + what); }
  char peek() { ws(); if (p >= end) fail("unexpected end"); return *p; }
  void expect(char c) { if (peek() != c) fail("unexpected character"); p++; }
  std::string str() {
    expect('"');
    std::string s;
    while (true) {
      if (p >= end) fail("unterminated string");
      char c = *p++;
      if (c == '"') break;
      if (c == '\\') {
        if (p >= end) fail("bad escape");
        char e = *p++;
        switch (e) {
          case '"': s += '"'; break; case '\\': s += '\\'; break; case '/': s += '/'; break;
          case 'b': s += '\b'; break; case 'f': s += '\f'; break; case 'n': s += '\n'; break;
          case 'r': s += '\r'; break; case 't': s += '\t'; break;
          case 'u': {
            if (end - p < 4) fail("bad \\u escape");
            unsigned cp = 0;
            for (int i = 0; i < 4; i++) {
              char h = *p++; cp <<= 4;
              if (h >= '0' && h <= '9') cp |= h - '0'; else if (h >= 'a' && h <= 'f') cp |= h - 'a' + 10;
              else if (h >= 'A' && h <= 'F') cp |= h - 'A' + 10; else fail("bad \\u escape");
            }
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
            else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: fail("bad escape");
        }
      } else s += c;
    }
    return s;
  }
  // a JSON number; accepted as a value only if it is a non-negative integer that fits u64 (lib.rs:211-214)
  bool number(uint64_t* out) {
    ws();
    const char* s = p;
    bool neg = false, integral = true;
    if (p < end && *p == '-') { neg = true; p++; }
    if (p >= end || *p < '0' || *p > '9') fail("bad number");
    while (p < end && *p >= '0' && *p <= '9') p++;
    if (p < end && *p == '.') { integral = false; p++; while (p < end && *p >= '0' && *p <= '9') p++; }
    if (p < end && (*p == 'e' || *p == 'E')) { integral = false; p++; if (p < end && (*p == '+' || *p == '-')) p++; while (p < end && *p >= '0' && *p <= '9') p++; }
    if (neg || !integral) return false;
    uint64_t v = 0;
    for (const char* q = s; q < p; q++) {
      uint64_t d = (uint64_t)(*q - '0');
      if (v > (UINT64_MAX - d) / 10) return false;
      v = v * 10 + d;
    }
    *out = v;
    return true;
  }
  void skip_value() {   // only used to give a precise error for unsupported value kinds
    char c = peek();
    if (c == '{') { p++; if (peek() == '}') { p++; return; } while (true) { str(); expect(':'); skip_value(); if (peek() == ',') { p++; continue; } expect('}'); return; } }
    if (c == '[') { p++; if (peek() == ']') { p++; return; } while (true) { skip_value(); if (peek() == ',') { p++; continue; } expect(']'); return; } }
    if (c == '"') { str(); return; }
    if (c == 't' && end - p >= 4 && !memcmp(p, "true", 4)) { p += 4; return; }
    if (c == 'f' && end - p >= 5 && !memcmp(p, "false", 5)) { p += 5; return; }
    if (c == 'n' && end - p >= 4 && !memcmp(p, "null", 4)) { p += 4; return; }
    uint64_t v; number(&v);
  }
};

U256 parse_scalar(J& j, const std::string& key, bool in_array) {
  char c = j.peek();
  if (c == '"') {
    std::string s = j.str();
    U256 v;
    if (!u256_parse_dec(s, &v)) throw Error("Failed to calculate witness: InputFieldNumberParseError(\"" + s + "\")");
    return v;
  }
  if (c == '-' || (c >= '0' && c <= '9')) {
    uint64_t v;
    if (!j.number(&v)) throw Error("Failed to calculate witness: InputsUnmarshal(\"signal value is not a positive integer\")");
    return u256_from_u64(v);
  }
  if (in_array) throw Error("Failed to calculate witness: InputsUnmarshal(\"inputs must be a string: " + key + "\")");
  throw Error("Failed to calculate witness: InputsUnmarshal(\"value for key " + key +
              " must be an a number as a string, as a number of an array of strings of numbers\")");
}
}  // namespace

InputList deserialize_inputs(const char* json, size_t len) {
  J j{json, json + len};
  if (j.peek() != '{') {
    j.skip_value();
    throw Error("Failed to calculate witness: InputsUnmarshal(\"inputs must be an object\")");
  }
  j.p++;
  InputList out;
  if (j.peek() == '}') { j.p++; }
  else {
    while (true) {
      std::string key = j.str();
      j.expect(':');
      std::vector<U256> vals;
      if (j.peek() == '[') {
        j.p++;
        if (j.peek() == ']') j.p++;
        else while (true) {
          vals.push_back(parse_scalar(j, key, true));
          if (j.peek() == ',') { j.p++; continue; }
          j.expect(']');
          break;
        }
      } else vals.push_back(parse_scalar(j, key, false));
      // serde_json::Map keeps the last value of a duplicated key
      bool replaced = false;
      for (auto& kv : out) if (kv.first == key) { kv.second = vals; replaced = true; }
      if (!replaced) out.emplace_back(key, std::move(vals));
      if (j.peek() == ',') { j.p++; continue; }
      j.expect('}');
      break;
    }
  }
  j.ws();
  if (j.p != j.end) j.fail("trailing characters");
  return out;
}

std::vector<U256> build_inputs_buffer(const Graph& g, const InputList& inputs) {
  std::vector<U256> buf(g.inputs_size, u256_from_u64(0));
  buf[0] = u256_from_u64(1);
  for (auto& kv : inputs) {
    auto it = g.inputs.find(kv.first);
    if (it == g.inputs.end()) throw Error("Failed to calculate witness: unknown input signal " + kv.first);
    uint32_t off = it->second.first, ln = it->second.second;
    if (ln != kv.second.size()) throw Error("Failed to calculate witness: Invalid input length for " + kv.first);
    if ((size_t)off + ln > buf.size()) throw Error("Failed to calculate witness: input " + kv.first + " out of range");
    for (uint32_t i = 0; i < ln; i++) buf[off + i] = kv.second[i];
  }
  return buf;
}

}  // namespace gw
