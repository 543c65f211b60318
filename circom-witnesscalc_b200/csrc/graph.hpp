// Witness-graph model and the wtns.graph.001 file codec.
//
// Mirrors, for the drop-in boundary, /root/reference/src/graph.rs:236-245 (enum Node),
// src/storage.rs:214-249 (deserialize_witnesscalc_graph), :137-183 (serialize_witnesscalc_graph) and
// the wire schema protos/messages.proto:1-85.  The protobuf wire format is decoded by hand (no
// protoc/prost in this build).  Unlike the reference, malformed files raise gw::Error instead of
// panicking, and operand indices are validated (reference: graph.rs:343-356, call commented out).
#pragma once
#include <stdint.h>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace gw {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

struct U256 {
  uint32_t l[8];
  bool operator==(const U256& o) const { for (int i = 0; i < 8; i++) if (l[i] != o.l[i]) return false; return true; }
  bool operator<(const U256& o) const { for (int i = 7; i >= 0; i--) if (l[i] != o.l[i]) return l[i] < o.l[i]; return false; }
};
U256 u256_from_le_bytes_mod_order(const uint8_t* p, size_t n);   // Fr::from_le_bytes_mod_order, storage.rs:28
U256 u256_from_u64(uint64_t v);
bool u256_parse_dec(const std::string& s, U256* out);            // U256::from_str_radix(s, 10), lib.rs:208
bool u256_parse_dec_digits(const char* s, size_t n, U256* out);  // the same for a run of decimal digits (no underscores)
extern const U256 BN254_M;

enum NodeKind : uint8_t { N_INPUT = 0, N_CONST = 1, N_UNO = 2, N_DUO = 3, N_TRES = 4 };

struct Node {
  uint8_t kind;
  uint8_t op;          // DuoOp / UnoOp / TresOp number of protos/messages.proto
  uint32_t a, b, c;    // operand node indices; a = input index for N_INPUT, constant-table index for N_CONST
};

struct Graph {
  std::vector<Node> nodes;
  std::vector<U256> constants;                                   // canonical values of N_CONST nodes
  std::vector<uint32_t> witness_signals;                         // node index per witness position
  std::map<std::string, std::pair<uint32_t, uint32_t>> inputs;   // name -> (offset, len), InputSignalsInfo lib.rs:19
  uint32_t inputs_size = 1;                                      // get_inputs_size, lib.rs:138-152 (>= every mapped slot)
  size_t n_ops() const;                                          // Op + UnoOp + TresOp nodes
};

Graph deserialize_witnesscalc_graph(const uint8_t* data, size_t len);
std::vector<uint8_t> serialize_witnesscalc_graph(const Graph& g);

}  // namespace gw
