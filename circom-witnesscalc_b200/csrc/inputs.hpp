// Inputs JSON -> flat input buffer.
// Mirrors /root/reference/src/lib.rs:195-247 (deserialize_inputs), :154-181 (populate_inputs,
// get_inputs_buffer).  Errors the reference reports as Error::InputsUnmarshal /
// InputFieldNumberParseError keep their wording; cases where the reference panics (invalid JSON,
// unknown key, wrong length) are reported as gw::Error instead.
#pragma once
#include "graph.hpp"

namespace gw {

typedef std::vector<std::pair<std::string, std::vector<U256>>> InputList;   // insertion order kept

InputList deserialize_inputs(const char* json, size_t len);
// buffer of g.inputs_size values, slot 0 = 1, unmentioned slots = 0 (lib.rs:177-181)
std::vector<U256> build_inputs_buffer(const Graph& g, const InputList& inputs);

}  // namespace gw
