// .wtns (snarkjs witness file, version 2) framing.
// Mirrors /root/reference/src/lib.rs:114-123 (wtns_from_witness; layout of the wtns-file 0.1.5 crate):
//   "wtns" u32(2) u32(2) | u32(1) u64(40) u32(32) M[32 LE] u32(W) | u32(2) u64(32 W) W x 32 B LE
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace gw {
static const size_t WTNS_HEADER_BYTES = 76;
inline size_t wtns_size(size_t n_witness) { return WTNS_HEADER_BYTES + 32 * n_witness; }
void wtns_write_header(uint8_t* dst, uint32_t n_witness);   // writes the 76 bytes preceding the values
}  // namespace gw
