// calc-witness <graph.bin> <inputs.json> <witness.wtns>
// Same CLI as /root/reference/src/bin/calc-witness.rs:13-49 (usage text, exit code, the two stdout
// lines, timer started after the files are read), evaluated on the GPU through the C ABI.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/graph_witness.h"

static bool read_file(const char* path, std::vector<char>& out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  out.resize((size_t)n);
  size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n;
}

int main(int argc, char** argv) {
  if (argc != 4) {
    fprintf(stderr, "Usage: %s <graph.bin> <inputs.json> <witness.wtns>\n", argv[0]);
    return 1;
  }
  std::vector<char> inputs, graph;
  if (!read_file(argv[2], inputs)) { fprintf(stderr, "Failed to read input file\n"); return 101; }
  if (!read_file(argv[1], graph)) { fprintf(stderr, "Failed to read graph file\n"); return 101; }
  inputs.push_back('\0');
  auto t0 = std::chrono::steady_clock::now();
  void* wtns = nullptr; size_t wtns_len = 0;
  gw_status_t status; status.code = OK; status.error_msg = nullptr;
  int r = gw_calc_witness(inputs.data(), graph.data(), graph.size(), &wtns, &wtns_len, &status);
  if (r != 0) {
    fprintf(stderr, "Error: %s\n", status.error_msg ? status.error_msg : "unknown");
    gw_free_status(&status);
    return 101;
  }
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  printf("Witness generated in: %.6fms\n", ms);
  FILE* f = fopen(argv[3], "wb");
  if (!f || fwrite(wtns, 1, wtns_len, f) != wtns_len) { fprintf(stderr, "Failed to write %s\n", argv[3]); return 101; }
  fclose(f);
  free(wtns);
  printf("witness saved to %s\n", argv[3]);
  return 0;
}
