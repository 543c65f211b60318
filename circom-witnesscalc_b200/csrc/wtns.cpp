#include "wtns.hpp"

#include <string.h>

#include "graph.hpp"

namespace gw {
static void put32(uint8_t*& p, uint32_t v) { for (int i = 0; i < 4; i++) *p++ = (uint8_t)(v >> (8 * i)); }
static void put64(uint8_t*& p, uint64_t v) { for (int i = 0; i < 8; i++) *p++ = (uint8_t)(v >> (8 * i)); }
void wtns_write_header(uint8_t* dst, uint32_t n_witness) {
  uint8_t* p = dst;
  memcpy(p, "wtns", 4); p += 4;
  put32(p, 2); put32(p, 2);
  put32(p, 1); put64(p, 40);
  put32(p, 32);
  for (int i = 0; i < 8; i++) put32(p, BN254_M.l[i]);
  put32(p, n_witness);
  put32(p, 2); put64(p, 32ull * n_witness);
}
}  // namespace gw
