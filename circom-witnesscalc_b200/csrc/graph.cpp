#include "graph.hpp"

#include <string.h>

#include <algorithm>

namespace gw {

const U256 BN254_M = {{0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}};

static const char kMagic[] = "wtns.graph.001";   // storage.rs:16
static const size_t kMagicLen = 14;

// ---- small big-int helpers (host side only: loading constants / parsing inputs) --------------------
static bool geq(const U256& a, const U256& b) { return !(a < b); }
static void sub_in_place(U256& a, const U256& b) {
  int64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (int64_t)a.l[i] - b.l[i]; a.l[i] = (uint32_t)c; c >>= 32; }
}
static void reduce_mod_m(U256& a) { while (geq(a, BN254_M)) sub_in_place(a, BN254_M); }
// a = (a * 256 + byte) mod M for a < M
static void shl8_add_mod(U256& a, uint8_t byte) {
  for (int k = 0; k < 8; k++) {                    // eight doublings keep every intermediate < 2M < 2^256
    uint32_t carry = 0;
    for (int i = 0; i < 8; i++) { uint32_t nc = a.l[i] >> 31; a.l[i] = (a.l[i] << 1) | carry; carry = nc; }
    reduce_mod_m(a);
  }
  uint64_t c = byte;
  for (int i = 0; i < 8 && c; i++) { c += a.l[i]; a.l[i] = (uint32_t)c; c >>= 32; }
  reduce_mod_m(a);
}

U256 u256_from_le_bytes_mod_order(const uint8_t* p, size_t n) {
  U256 r; memset(&r, 0, sizeof r);
  if (n <= 32) {
    uint8_t buf[32] = {0};
    memcpy(buf, p, n);
    for (int i = 0; i < 8; i++) r.l[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) | ((uint32_t)buf[4 * i + 3] << 24);
    reduce_mod_m(r);
    return r;
  }
  for (size_t i = n; i-- > 0;) shl8_add_mod(r, p[i]);   // most significant byte first
  return r;
}

U256 u256_from_u64(uint64_t v) {
  U256 r; memset(&r, 0, sizeof r); r.l[0] = (uint32_t)v; r.l[1] = (uint32_t)(v >> 32); return r;
}

// decimal digits only (no underscores): nine digits at a time on 64-bit limbs -- 9 x 4 multiply-adds for a field element
// instead of 77 x 8 (the batch input path parses 170 of these per authV2 record)
bool u256_parse_dec_digits(const char* s, size_t n, U256* out) {
  uint64_t r[4] = {0, 0, 0, 0};
  size_t i = 0;
  while (i < n) {
    const size_t take = (n - i) % 9 ? (n - i) % 9 : 9;      // a short first chunk, then whole chunks of nine
    uint64_t chunk = 0, scale = 1;
    for (size_t k = 0; k < take; k++) {
      const char ch = s[i + k];
      if (ch < '0' || ch > '9') return false;
      chunk = chunk * 10 + (uint64_t)(ch - '0'); scale *= 10;
    }
    unsigned __int128 c = chunk;
    for (int q = 0; q < 4; q++) { c += (unsigned __int128)r[q] * scale; r[q] = (uint64_t)c; c >>= 64; }
    if (c) return false;                           // does not fit 256 bits
    i += take;
  }
  for (int q = 0; q < 4; q++) { out->l[2 * q] = (uint32_t)r[q]; out->l[2 * q + 1] = (uint32_t)(r[q] >> 32); }
  return true;
}

bool u256_parse_dec(const std::string& s, U256* out) {
  if (s.find('_') == std::string::npos) return u256_parse_dec_digits(s.data(), s.size(), out);
  U256 r; memset(&r, 0, sizeof r);
  for (char ch : s) {
    if (ch == '_') continue;                       // ruint skips underscores
    if (ch < '0' || ch > '9') return false;
    uint64_t c = (uint64_t)(ch - '0');
    for (int i = 0; i < 8; i++) { c += (uint64_t)r.l[i] * 10u; r.l[i] = (uint32_t)c; c >>= 32; }
    if (c) return false;                           // does not fit 256 bits
  }
  *out = r;
  return true;
}

size_t Graph::n_ops() const {
  size_t n = 0;
  for (const Node& nd : nodes) n += (nd.kind >= N_UNO);
  return n;
}

// ---- protobuf wire reader ---------------------------------------------------------------------------
namespace {
struct Reader {
  const uint8_t* p; const uint8_t* end;
  bool eof() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0; int shift = 0;
    for (;;) {
      if (p >= end) throw Error("graph: unexpected EOF in varint");
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
      if (shift > 63) throw Error("graph: varint too long");
    }
  }
  Reader sub() {
    uint64_t n = varint();
    if (n > (uint64_t)(end - p)) throw Error("graph: unexpected EOF in length-delimited field");
    Reader r{p, p + n}; p += n; return r;
  }
  void skip(uint32_t wt) {
    switch (wt) {
      case 0: varint(); break;
      case 1: if (end - p < 8) throw Error("graph: unexpected EOF"); p += 8; break;
      case 2: sub(); break;
      case 5: if (end - p < 4) throw Error("graph: unexpected EOF"); p += 4; break;
      default: throw Error("graph: unsupported protobuf wire type");
    }
  }
};

// read up to nmax uint32 scalar fields (numbered 1..nmax) of a flat message
void read_uints(Reader r, uint32_t* out, int nmax) {
  for (int i = 0; i <= nmax; i++) out[i] = 0;
  while (!r.eof()) {
    uint64_t key = r.varint(); uint32_t fno = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (wt == 0 && fno >= 1 && (int)fno <= nmax) out[fno] = (uint32_t)r.varint();
    else r.skip(wt);
  }
}

void put_varint(std::vector<uint8_t>& b, uint64_t v) {
  while (v >= 0x80) { b.push_back((uint8_t)(v | 0x80)); v >>= 7; }
  b.push_back((uint8_t)v);
}
void put_kv(std::vector<uint8_t>& b, uint32_t fno, uint64_t v) { if (v) { put_varint(b, fno << 3); put_varint(b, v); } }
void put_ld(std::vector<uint8_t>& b, uint32_t fno, const std::vector<uint8_t>& payload) {
  put_varint(b, (fno << 3) | 2); put_varint(b, payload.size()); b.insert(b.end(), payload.begin(), payload.end());
}
}  // namespace

// upper bound of the input buffer length (slots of 32 bytes): 2^28 slots = 8 GiB per input set is beyond any circuit
static const uint64_t kMaxInputs = 1ull << 28;

Graph deserialize_witnesscalc_graph(const uint8_t* data, size_t len) {
  if (len < kMagicLen + 8) throw Error("graph: file too short");
  if (memcmp(data, kMagic, kMagicLen) != 0) throw Error("graph: Invalid magic");
  uint64_t n_nodes = 0;
  for (int i = 0; i < 8; i++) n_nodes |= (uint64_t)data[kMagicLen + i] << (8 * i);
  if (n_nodes > len) throw Error("graph: node count larger than file");
  Reader file{data + kMagicLen + 8, data + len};
  Graph g;
  g.nodes.reserve(n_nodes);
  std::map<U256, uint32_t> const_ix;
  for (uint64_t i = 0; i < n_nodes; i++) {
    Reader msg = file.sub();
    bool have = false;
    Node nd{};
    while (!msg.eof()) {
      uint64_t key = msg.varint(); uint32_t fno = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
      if (wt != 2 || fno < 1 || fno > 5) { msg.skip(wt); continue; }
      Reader inner = msg.sub();
      uint32_t f[5];
      have = true;
      memset(&nd, 0, sizeof nd);
      switch (fno) {
        case 1: read_uints(inner, f, 1); nd.kind = N_INPUT; nd.a = f[1]; break;
        case 2: {
          bool got = false; U256 v{};
          while (!inner.eof()) {
            uint64_t k2 = inner.varint();
            if ((k2 >> 3) == 1 && (k2 & 7) == 2) {
              Reader big = inner.sub();
              got = true; memset(&v, 0, sizeof v);
              while (!big.eof()) {
                uint64_t k3 = big.varint();
                if ((k3 >> 3) == 1 && (k3 & 7) == 2) { Reader bytes = big.sub(); v = u256_from_le_bytes_mod_order(bytes.p, (size_t)(bytes.end - bytes.p)); }
                else big.skip((uint32_t)(k3 & 7));
              }
            } else inner.skip((uint32_t)(k2 & 7));
          }
          if (!got) throw Error("graph: ConstantNode without value");
          nd.kind = N_CONST;
          auto it = const_ix.find(v);
          if (it == const_ix.end()) { it = const_ix.emplace(v, (uint32_t)g.constants.size()).first; g.constants.push_back(v); }
          nd.a = it->second;
          break;
        }
        case 3: read_uints(inner, f, 2); nd.kind = N_UNO; nd.op = (uint8_t)f[1]; nd.a = f[2];
          if (f[1] > 3) throw Error("graph: unknown UnoOp"); break;
        case 4: read_uints(inner, f, 3); nd.kind = N_DUO; nd.op = (uint8_t)f[1]; nd.a = f[2]; nd.b = f[3];
          if (f[1] > 19) throw Error("graph: unknown DuoOp"); break;
        case 5: read_uints(inner, f, 4); nd.kind = N_TRES; nd.op = (uint8_t)f[1]; nd.a = f[2]; nd.b = f[3]; nd.c = f[4];
          if (f[1] > 0) throw Error("graph: unknown TresOp"); break;
      }
    }
    if (!have) throw Error("graph: empty Node message");
    uint32_t self = (uint32_t)g.nodes.size();
    if (nd.kind == N_UNO && nd.a >= self) throw Error("graph: forward operand reference");
    if (nd.kind == N_DUO && (nd.a >= self || nd.b >= self)) throw Error("graph: forward operand reference");
    if (nd.kind == N_TRES && (nd.a >= self || nd.b >= self || nd.c >= self)) throw Error("graph: forward operand reference");
    g.nodes.push_back(nd);
  }
  Reader meta = file.sub();
  while (!meta.eof()) {
    uint64_t key = meta.varint(); uint32_t fno = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (fno == 1 && wt == 2) { Reader pk = meta.sub(); while (!pk.eof()) g.witness_signals.push_back((uint32_t)pk.varint()); }
    else if (fno == 1 && wt == 0) g.witness_signals.push_back((uint32_t)meta.varint());
    else if (fno == 2 && wt == 2) {
      Reader entry = meta.sub();
      std::string name; uint32_t f[3] = {0, 0, 0};
      while (!entry.eof()) {
        uint64_t k2 = entry.varint();
        if ((k2 >> 3) == 1 && (k2 & 7) == 2) { Reader s = entry.sub(); name.assign((const char*)s.p, (size_t)(s.end - s.p)); }
        else if ((k2 >> 3) == 2 && (k2 & 7) == 2) read_uints(entry.sub(), f, 2);
        else entry.skip((uint32_t)(k2 & 7));
      }
      g.inputs[name] = std::make_pair(f[1], f[2]);
    } else meta.skip(wt);
  }
  for (uint32_t s : g.witness_signals) if (s >= g.nodes.size()) throw Error("graph: witness signal out of range");
  // get_inputs_size, lib.rs:138-152: (max Input idx in the first contiguous run of Input nodes) + 1.
  // Deviation: also cover every Input node and every mapped slot so nothing can index out of range.
  // Computed in 64 bits and bounded: an Input index or a mapped extent near 2^32 must not wrap the buffer size.
  uint64_t mx = 0;
  for (const Node& nd : g.nodes) if (nd.kind == N_INPUT) mx = std::max<uint64_t>(mx, nd.a);
  for (auto& kv : g.inputs) if (kv.second.second) mx = std::max<uint64_t>(mx, (uint64_t)kv.second.first + kv.second.second - 1);
  if (mx + 1 > kMaxInputs) throw Error("graph: input index or input signal extent out of range");
  g.inputs_size = (uint32_t)(mx + 1);
  return g;
}

std::vector<uint8_t> serialize_witnesscalc_graph(const Graph& g) {
  std::vector<uint8_t> out(kMagic, kMagic + kMagicLen);
  uint64_t n = g.nodes.size();
  for (int i = 0; i < 8; i++) out.push_back((uint8_t)(n >> (8 * i)));
  for (const Node& nd : g.nodes) {
    std::vector<uint8_t> inner, msg;
    uint32_t fno = 0;
    switch (nd.kind) {
      case N_INPUT: put_kv(inner, 1, nd.a); fno = 1; break;
      case N_CONST: {
        const U256& v = g.constants.at(nd.a);
        std::vector<uint8_t> le(32);
        for (int i = 0; i < 32; i++) le[i] = (uint8_t)(v.l[i / 4] >> (8 * (i % 4)));
        while (le.size() > 1 && le.back() == 0) le.pop_back();     // num-bigint to_bytes_le: minimal, zero -> [0]
        std::vector<uint8_t> big; put_ld(big, 1, le); put_ld(inner, 1, big); fno = 2; break;
      }
      case N_UNO: put_kv(inner, 1, nd.op); put_kv(inner, 2, nd.a); fno = 3; break;
      case N_DUO: put_kv(inner, 1, nd.op); put_kv(inner, 2, nd.a); put_kv(inner, 3, nd.b); fno = 4; break;
      case N_TRES: put_kv(inner, 1, nd.op); put_kv(inner, 2, nd.a); put_kv(inner, 3, nd.b); put_kv(inner, 4, nd.c); fno = 5; break;
      default: throw Error("graph: bad node kind");
    }
    put_ld(msg, fno, inner);
    put_varint(out, msg.size());
    out.insert(out.end(), msg.begin(), msg.end());
  }
  uint64_t meta_off = out.size();
  std::vector<uint8_t> meta;
  if (!g.witness_signals.empty()) {
    std::vector<uint8_t> packed;
    for (uint32_t s : g.witness_signals) put_varint(packed, s);
    put_ld(meta, 1, packed);
  }
  for (auto& kv : g.inputs) {
    std::vector<uint8_t> sd, entry;
    put_kv(sd, 1, kv.second.first); put_kv(sd, 2, kv.second.second);
    put_ld(entry, 1, std::vector<uint8_t>(kv.first.begin(), kv.first.end()));
    put_ld(entry, 2, sd);
    put_ld(meta, 2, entry);
  }
  put_varint(out, meta.size());
  out.insert(out.end(), meta.begin(), meta.end());
  for (int i = 0; i < 8; i++) out.push_back((uint8_t)(meta_off >> (8 * i)));
  return out;
}

}  // namespace gw
