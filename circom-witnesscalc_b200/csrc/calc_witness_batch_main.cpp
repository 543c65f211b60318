// calc-witness-batch <graph.bin> <inputs.jsonl> <out_dir> [n_gpus]
// Batch companion of calc-witness (/root/reference/src/bin/calc-witness.rs:13-49): the graph is loaded once, every
// line of <inputs.jsonl> (or every element of a top-level JSON array) is one inputs object, and set i is written to
// <out_dir>/<i as 8 digits>.wtns, byte-identical to what `calc-witness` writes for the same inputs.  The work is
// done by the C ABI: gw_inputs_parse_batch (multi-threaded JSON -> packed rows), gw_calc_witness_batch_wtns
// (GPU evaluation, .wtns images framed by the pitched device-to-host copy).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
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
static int fail(gw_status_t* st, const char* what) {
  fprintf(stderr, "Error: %s: %s\n", what, st->error_msg ? st->error_msg : "unknown");
  gw_free_status(st);
  return 101;
}

int main(int argc, char** argv) {
  if (argc != 4 && argc != 5) {
    fprintf(stderr, "Usage: %s <graph.bin> <inputs.jsonl> <out_dir> [n_gpus]\n", argv[0]);
    return 1;
  }
  const int n_gpus = argc == 5 ? atoi(argv[4]) : 1;
  std::vector<char> text, graph_bytes;
  if (!read_file(argv[2], text)) { fprintf(stderr, "Failed to read input file\n"); return 101; }
  if (!read_file(argv[1], graph_bytes)) { fprintf(stderr, "Failed to read graph file\n"); return 101; }
  mkdir(argv[3], 0777);
  gw_status_t st; st.code = OK; st.error_msg = nullptr;
  gw_graph_t* g = nullptr;
  auto t0 = std::chrono::steady_clock::now();
  if (gw_graph_load(graph_bytes.data(), graph_bytes.size(), &g, &st)) return fail(&st, "graph");
  gw_graph_info_t info; gw_graph_info(g, &info);
  uint8_t* inputs = nullptr; size_t n = 0;
  if (gw_inputs_parse_batch(g, text.data(), text.size(), 0, &inputs, &n, &st)) return fail(&st, "inputs");
  auto t1 = std::chrono::steady_clock::now();
  const size_t fsz = gw_wtns_file_size(g), in_b = (size_t)info.n_inputs * 32;
  const size_t chunk = std::max<size_t>(1, std::min<size_t>(n, ((size_t)4 << 30) / fsz));   // <= 4 GiB of images at a time
  std::vector<uint8_t> files(chunk * fsz);
  double ms_eval = 0;
  for (size_t lo = 0; lo < n; lo += chunk) {
    const size_t nb = std::min(chunk, n - lo);
    auto a = std::chrono::steady_clock::now();
    if (gw_calc_witness_batch_wtns(g, inputs + lo * in_b, nb, files.data(), fsz, nullptr, n_gpus, &st)) return fail(&st, "witness");
    ms_eval += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
    for (size_t i = 0; i < nb; i++) {
      char path[4096];
      snprintf(path, sizeof path, "%s/%08zu.wtns", argv[3], lo + i);
      FILE* f = fopen(path, "wb");
      if (!f || fwrite(files.data() + i * fsz, 1, fsz, f) != fsz) { fprintf(stderr, "Failed to write %s\n", path); return 101; }
      fclose(f);
    }
  }
  double ms_load = std::chrono::duration<double, std::milli>(t1 - t0).count();
  printf("%zu input sets parsed, graph loaded in: %.3fms\n", n, ms_load);
  printf("Witnesses generated in: %.3fms (%.1f witnesses/s)\n", ms_eval, n ? n / (ms_eval * 1e-3) : 0.0);
  printf("witnesses saved to %s/\n", argv[3]);
  free(inputs);
  gw_graph_free(g);
  return 0;
}
