/* C ABI of the B200 witness evaluator.
 *
 * Part 1 is the reference's header verbatim in meaning: /root/reference/include/graph_witness.h:7-28
 * (GW_ERROR_CODE, gw_status_t, gw_calc_witness, gw_free_status), implemented by the reference in
 * src/lib.rs:28-111.  examples/calc_witness.c of the reference compiles against this file unchanged.
 * Part 2 is the batch extension: load a graph once, evaluate many input sets.
 *
 * Every entry point is backed by CUDA kernels; there is no CPU fallback: without a usable GPU the
 * calls return 1 with an error message.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#ifndef RUST_GRAPH_WITNESS_H
#define RUST_GRAPH_WITNESS_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- part 1: drop-in (reference include/graph_witness.h:7-28) ---------------------------------- */

typedef enum {
  OK = 0,
  ERROR = 1
} GW_ERROR_CODE;

typedef struct {
  GW_ERROR_CODE code;
  char *error_msg;
} gw_status_t;

/* inputs: NUL-terminated UTF-8 JSON; graph_data: a wtns.graph.001 file; on success *wtns_data is a
 * malloc'ed .wtns file of *wtns_len bytes that the caller frees.  Returns 0 on success, 1 on error
 * (status->code = ERROR, status->error_msg malloc'ed).  Deviation from src/lib.rs:106-108: on
 * success status is {OK, NULL} (the reference writes {ERROR, "test error"}) and nothing is printed. */
int
gw_calc_witness(const char *inputs,
                const void *graph_data, const size_t graph_data_len,
                void **wtns_data, size_t *wtns_len,
                const gw_status_t *status);

/* reference include/graph_witness.h:23-28 defines this function in the header; kept, but static
 * inline so that the header can be included from more than one translation unit. */
static inline void
gw_free_status(gw_status_t *status) {
  if (status->error_msg != NULL) {
    free(status->error_msg);
  }
}

/* ---- part 2: batch extension ---------------------------------------------------------------------- */

typedef struct gw_graph gw_graph_t;

typedef struct {
  uint64_t n_nodes;       /* nodes in the graph file */
  uint64_t n_ops;         /* Op + UnoOp + TresOp nodes (unit of the node-ops/s metric) */
  uint32_t n_inputs;      /* I: input buffer length including slot 0 (the constant 1) */
  uint32_t n_witness;     /* W: witness length */
  uint32_t n_input_signals;
  uint32_t n_instrs;      /* device instructions per witness */
  uint32_t n_regs;        /* per-witness registers in shared memory */
  uint32_t n_spill;       /* per-witness spill slots in HBM */
  uint64_t n_mul;         /* live Mul nodes of the graph (field multiplications per witness, algorithmic) */
  uint64_t n_div;         /* live Div nodes of the graph (field divisions per witness, algorithmic) */
  uint64_t n_spill_ld;    /* spill loads per witness */
  uint64_t n_spill_st;    /* spill stores per witness */
  /* what the plan compiler made of it (per witness) */
  uint32_t n_slots;       /* 16-byte program slots */
  uint32_t n_dot;         /* fused linear combinations (one Montgomery reduction each) */
  uint32_t n_dot_mac;     /* value x constant products inside them */
  uint32_t n_mul_instr;   /* value x value multiplications (Barrett) */
  uint32_t n_inversions;  /* modular inversions after Div batching */
  uint32_t threads;       /* threads per CTA */
  uint32_t sets_per_thread; /* input sets evaluated per thread */
  uint32_t n_narrow_instr; /* instructions computed on int64 (values the plan compiler proves to be small, isa.h F_NARROW) */
  /* bit-sliced plan (Boolean graphs: 32 input sets per machine word, one 3-input LUT per lane) */
  uint32_t bit_eligible;  /* 1: batch launches of this graph run bit-sliced (input sets that break the bit contract: generic kernel) */
  uint32_t bit_luts;      /* LUT instructions per group of 32 input sets */
  uint32_t bit_steps;     /* steps of 32 LUTs */
  uint32_t bit_wide;      /* witness values assembled from several planes (Bits2Num sums, field inputs passed through) */
} gw_graph_info_t;

/* replaces storage::deserialize_witnesscalc_graph (src/storage.rs:214-249) + upload; parse once */
int gw_graph_load(const void *graph_data, size_t graph_data_len, gw_graph_t **graph, gw_status_t *status);
void gw_graph_free(gw_graph_t *graph);
int gw_graph_info(const gw_graph_t *graph, gw_graph_info_t *info);
/* i-th input signal of the graph's input map (InputSignalsInfo, src/lib.rs:19); name is owned by the graph */
int gw_graph_input_signal(const gw_graph_t *graph, uint32_t i, const char **name, uint32_t *offset, uint32_t *len);

/* calc_witness (src/lib.rs:125-136) + wtns_from_witness (:114-123) on a pre-loaded graph */
int gw_graph_calc_witness(gw_graph_t *graph, const char *inputs_json, void **wtns_data, size_t *wtns_len,
                          gw_status_t *status);

/* graph::evaluate (src/graph.rs:367-391) for n_sets independent input sets, HOST buffers.
 *   inputs : n_sets x n_inputs x 32 bytes, little-endian 256-bit values, row-major; slot 0 of every
 *            row is ignored (it is the constant 1); values >= M are reduced mod M like Fr::new.
 *   witness: n_sets x n_witness x 32 bytes, canonical little-endian (the .wtns payload of each set).
 *   flags  : optional (may be NULL) n_sets x uint32: per-set bits for cases where the reference
 *            panics (1 Shl overflow, 2 Bor/Bxor == M, 4 Pow, 8 Id, 16 Lnot/Bnot).
 *   n_gpus : the sets are split into n_gpus contiguous shards, one per device, no collective. */
int gw_calc_witness_batch(gw_graph_t *graph, const uint8_t *inputs, size_t n_sets, uint8_t *witness,
                          uint32_t *flags, int n_gpus, gw_status_t *status);

/* same on devices first_device .. first_device + n_gpus - 1: for hosts that run one process per GPU (each process
 * passes its own device) or that keep some devices for other work. */
int gw_calc_witness_batch_on(gw_graph_t *graph, int first_device, const uint8_t *inputs, size_t n_sets,
                             uint8_t *witness, uint32_t *flags, int n_gpus, gw_status_t *status);

/* same, buffers already resident on CUDA device `device`; enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = default stream) and asynchronous with respect to the host. */
int gw_calc_witness_batch_device(gw_graph_t *graph, int device, const void *d_inputs, size_t n_sets,
                                 void *d_witness, uint32_t *d_flags, void *cuda_stream, gw_status_t *status);

/* Calls on one graph and device may come from any thread and any stream: the library runs the kernels of one graph
 * on one device one after the other (they share a per-device scratch area), in the order the calls were made. */

/* single-witness latency mode (BASELINE config 5): ONE input set (n_inputs x 32 B, host), parallel across the nodes
 * of the graph's dependency levels; witness is n_witness x 32 B (host).  *kernel_ms (optional) receives the device time
 * of the kernels alone.
 *   - Boolean graphs (SHA-256, Num2Bits, comparators; gw_graph_info_t.bit_luts > 0) run their bit-sliced plan with
 *     one warp: a step = up to 32 independent LUT nodes, one per lane.  An input set that breaks the plan's contract
 *     (an input that must be 0 or 1 and is not) is evaluated by the generic kernel below in the same call.
 *   - everything else: one CTA; every warp walks its own stream of packets (one instruction per lane), packets wait on
 *     progress counters of the other warps; the long operations (Div, Pow, Idiv, Mod) have warps of their own.
 * Fails with "latency plan: graph is too wide ..." when the values alive at once do not fit the shared memory of one
 * SM.  gw_calc_witness / gw_graph_calc_witness use this mode for their single witness and fall back to the throughput
 * kernel with a batch of one in that case. */
int gw_calc_witness_latency(gw_graph_t *graph, int device, const uint8_t *inputs, uint8_t *witness, uint32_t *flags,
                            float *kernel_ms, gw_status_t *status);

/* ---- batch input path (replaces deserialize_inputs + populate_inputs, src/lib.rs:154-247, for many
 * input sets).  `text` is JSON Lines (one inputs object per non-empty line) or one top-level JSON array of such
 * objects; it is parsed by n_threads host threads (0 = all) into *inputs = a malloc'ed n_sets x n_inputs x 32 B
 * buffer in the layout gw_calc_witness_batch takes (caller frees).  Per-record semantics and error texts are those
 * of gw_calc_witness; the error message names the 1-based record. */
int gw_inputs_parse_batch(const gw_graph_t *graph, const char *text, size_t text_len, int n_threads,
                          uint8_t **inputs, size_t *n_sets, gw_status_t *status);

/* ---- batch output path ----
 * .wtns framing: like gw_calc_witness_batch, but set i is delivered as a complete .wtns file image
 * (wtns_from_witness, src/lib.rs:114-123: 76-byte header + n_witness x 32 B) at files + i * file_pitch;
 * file_pitch >= gw_wtns_file_size(graph) (bytes between consecutive images; pad as you like).  The payload rows
 * are placed by the device-to-host DMA itself (pitched copy), the headers are written by the host. */
size_t gw_wtns_file_size(const gw_graph_t *graph);
int gw_calc_witness_batch_wtns(gw_graph_t *graph, const uint8_t *inputs, size_t n_sets, uint8_t *files, size_t file_pitch,
                               uint32_t *flags, int n_gpus, gw_status_t *status);
/* Selected signals only (e.g. the public outputs): a new graph handle whose witness consists of the given witness
 * positions of `graph`, in the given order (repeats allowed).  Everything the selection does not depend on is dead
 * code for the device program, and only n_positions x 32 B per set cross PCIe.  All entry points work on the result;
 * free it with gw_graph_free. */
int gw_graph_select(const gw_graph_t *graph, const uint32_t *positions, size_t n_positions, gw_graph_t **selected,
                    gw_status_t *status);

/* Streaming (SURVEY 8f-2 "streaming D2H"): like gw_calc_witness_batch_on, but the witnesses are handed to `fn` chunk by
 * chunk as they land in a ring of three pinned host buffers per GPU that the library owns, so a batch whose witnesses
 * do not fit host memory (262 144 authV2 input sets = 806 GB) runs end to end at its real size.  `fn` is called with
 *   rows  : n_sets x row_bytes bytes (row_bytes = n_witness x 32), the witnesses of input sets first_set .. first_set + n_sets - 1,
 *   flags : n_sets x uint32 (the per-set bits of gw_calc_witness_batch),
 * valid only until it returns; returning nonzero stops the stream (the call then fails with an error).  One worker
 * thread per GPU calls it: chunks of one GPU arrive in order, chunks of different GPUs may arrive concurrently.
 * While `fn` looks at chunk k, chunk k + 1 is being copied and chunk k + 2 computed.  chunk_sets = 0 lets the library
 * choose (GW_STREAM_CHUNK_MB of pinned memory per ring slot, default 8192).  The worker threads are pinned to the CPUs of
 * their GPU's NUMA node before the ring is allocated (GW_NUMA=0 turns that off). */
typedef int (*gw_witness_chunk_fn)(void *user, int device, size_t first_set, size_t n_sets, const uint8_t *rows,
                                   size_t row_bytes, const uint32_t *flags);
int gw_calc_witness_batch_stream(gw_graph_t *graph, int first_device, int n_gpus, const uint8_t *inputs, size_t n_sets,
                                 size_t chunk_sets, gw_witness_chunk_fn fn, void *user, gw_status_t *status);

/* CUDA device used by the single-witness entry points gw_calc_witness / gw_graph_calc_witness (default: the GW_DEVICE
 * environment variable, else device 0 of CUDA_VISIBLE_DEVICES).  Returns 1 if the device does not exist. */
int gw_set_device(int device);

/* writes the 76-byte .wtns header for n_witness values (src/lib.rs:114-123) */
void gw_wtns_header(uint32_t n_witness, uint8_t *dst76);

/* diagnostics */
int gw_device_count(void);
/* integer-pipe microbenchmark: executed multiply-add instructions per second on `device`;
 * which = 0 IMAD.WIDE.U32 carry rows (32x32+64, the form the field multiplier uses), 1 mad.lo.u32,
 * 2 mad.hi.u32, 3 add.u32 */
double gw_microbench_imad(int device, int which);

#ifdef __cplusplus
}
#endif

#endif /* RUST_GRAPH_WITNESS_H */
