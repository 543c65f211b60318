#!/usr/bin/env python3
"""Headline benchmark: batched witness generation for circuit9_authV2 (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (graph::evaluate for every input set) over the whole batch held
by a rank: --batch input sets (default 262 144, BASELINE config 4), evaluated chunk by chunk because a
full authV2 witness is 3.08 MB (W = 96 259 values) and 262 144 of them do not fit HBM; every chunk's
witnesses are written to HBM in full.  Weak scaling: every rank owns its own --batch input sets, there
is no collective on the data path (witnesses are independent), only the timing barrier / max-reduce.

Printed JSON line (rank 0): metric witnesses/s (+ node-ops/s in `extra`), `value` kernel-only with
inputs resident in HBM, `e2e` through the C ABI with pinned HOST buffers (H2D + kernel + D2H inside the
timed region), `roofline` (integer pipe, the binding one for authV2) and `roofline_hbm`, `cpu_baseline`
(C restatement of the reference algorithm on the host cores, bounded sample), clocks, launches.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CIRCUIT_DEFAULT = "circuit9_authV2"
M = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def shard_range(n_units: int, rank: int, world: int):
    """contiguous shard [lo, hi) of n_units for `rank` (used for strong-scaling runs and by the tests)"""
    return n_units * rank // world, n_units * (rank + 1) // world


def synth_inputs(name: str, n_sets: int, n_inputs: int, input_map: dict, seed: int) -> np.ndarray:
    """Synthetic input sets, uint8 [n_sets, I, 32]: set 0 = the reference's own fixture
    (a valid proof request for authV2), the others uniform in [0, M) (bits for SHA-256 and for the
    *NoAux flags of authV2)."""
    from tests import util
    rng = np.random.default_rng(seed)
    vals = util.random_field_batch(rng, (n_sets, n_inputs))
    if "sha256" in name:
        vals[:] = 0
        vals[:, :, 0] = rng.integers(0, 2, size=(n_sets, n_inputs), dtype=np.uint64)
    for key in ("authClaimNonRevMtpNoAux", "gistMtpNoAux"):
        if key in input_map:
            off, ln = input_map[key]
            vals[:, off:off + ln, :] = 0
            vals[:, off:off + ln, 0] = rng.integers(0, 2, size=(n_sets, ln), dtype=np.uint64)
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    try:
        from oracle import pyoracle as po
        fixture = po.deserialize_inputs(util.golden_inputs(name))
        row = [1] + [0] * (n_inputs - 1)
        for k, v in fixture.items():
            off, ln = input_map[k]
            row[off:off + ln] = v
        vals[0] = np.frombuffer(util.pack_u256(row), dtype=np.uint64).reshape(n_inputs, 4)
    except FileNotFoundError:
        pass
    return vals.view(np.uint8).reshape(n_sets, n_inputs, 32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(name, graph_bytes, n_inputs, input_map, n_sets, steps, warmup, seed):
    """C restatement of the reference algorithm (oracle/ref_eval.c), all host threads"""
    from oracle import cref
    cg = cref.CGraph(graph_bytes)
    cores = os.cpu_count() or 1
    inp = synth_inputs(name, n_sets, n_inputs, input_map, seed)
    for _ in range(warmup):
        cg.evaluate_batch(inp[:max(cores, 1)], cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        cg.evaluate_batch(inp, cores)
    dt = (time.perf_counter() - t0) / steps
    return n_sets / dt, dt, cores, cg.n_ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--circuit", default=CIRCUIT_DEFAULT)
    ap.add_argument("--batch", type=int, default=262144, help="input sets per GPU and step")
    ap.add_argument("--chunk", type=int, default=0, help="input sets per kernel launch (0 = auto)")
    ap.add_argument("--unique", type=int, default=16384, help="distinct synthetic input sets (tiled to --batch)")
    ap.add_argument("--e2e-sets", type=int, default=16384)
    ap.add_argument("--cpu-sample", type=int, default=0, help="input sets for the cpu_baseline leg (0 = 96 x cores)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from tests import util
    graph_bytes = util.golden_graph(a.circuit)
    unit = "witnesses/s"
    metric = f"{a.circuit} witnesses/sec (batched witness generation, graph::evaluate)"

    if a.impl == "reference":
        # the reference's own CPU implementation of the path; the Rust binary cannot be built in this image
        # (no cargo/rustc), so this is its C restatement with the same cost structure, on all host threads.
        if rank != 0:
            return
        from oracle import cref
        cg = cref.CGraph(graph_bytes)
        from oracle import pyoracle as po
        _, _, imap = po.deserialize_graph(graph_bytes)
        cores = os.cpu_count() or 1
        n_sets = a.cpu_sample or 32 * cores
        val, dt, cores, n_ops = cpu_reference_run(a.circuit, graph_bytes, cg.n_inputs, imap, n_sets, a.steps, min(a.warmup, 1), 9)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery limbs)", "data": "synthetic",
            "config": {"workload": f"{a.circuit}: {n_sets} input sets per step (bounded sample of the {a.batch}-set batch)",
                       "graph_nodes": cg.n_nodes, "witness_len": cg.n_witness},
            "extra": {"node_ops_per_s": val * n_ops},
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port",
                             "sample": f"{n_sets} input sets x {a.steps} steps, oracle/ref_eval.c, one witness per thread"},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # native libraries (NCCL prints its version banner) write to fd 1: keep stdout clean for the ONE JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    import torch
    import torch.distributed as dist
    cwc = importlib.import_module("circom-witnesscalc_b200")
    if not torch.cuda.is_available() or cwc.device_count() < 1:
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the witness evaluator has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)     # plumbing only: barrier + max of the step time

    g = cwc.Graph(graph_bytes)
    info = g.info
    I, W = g.n_inputs, g.n_witness
    B = a.batch
    free_b, _ = torch.cuda.mem_get_info()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    wave = sms * 32                                          # one CTA per SM, whole warps: chunk = sms x T input sets
    quad = sms * 128                                         # ... and the same number of warps on each of the 4 schedulers of an SM
    # HBM budget: the chunk's witnesses (W x 32 B each) next to the inputs of the whole batch and the spill area
    # (n_spill x 32 B for each of the sms x 512 threads a launch can have)
    reserve = B * I * 32 + info["n_spill"] * 32 * sms * 512 + (3 << 30)
    max_chunk = int(max(wave, (free_b - reserve) // (W * 32)))
    if a.chunk:
        chunk = a.chunk
    elif B <= max_chunk:
        chunk = B
    elif max_chunk >= quad:
        chunk = min(max_chunk // quad * quad, sms * 512)     # kernel time steps with ceil(warps per SM / 4): fill whole quads
    else:
        chunk = max_chunk // wave * wave                     # whole CTAs on every SM: no ragged last wave
    n_chunks = (B + chunk - 1) // chunk
    n_unique = min(a.unique, B)
    host_in = synth_inputs(a.circuit, n_unique, I, g.input_signals, seed=9 + rank)
    d_unique = torch.from_numpy(host_in.reshape(n_unique, I * 32)).to(dev)
    d_in = d_unique.repeat((B + n_unique - 1) // n_unique, 1)[:B].contiguous()
    d_out = torch.empty((chunk, W * 32), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    chunks = [(lo, min(lo + chunk, B)) for lo in range(0, B, chunk)]

    def step():
        for lo, hi in chunks:
            g.calc_witness_batch_device(local_rank, d_in[lo:hi].data_ptr(), hi - lo, d_out.data_ptr(), None, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    value = world * B / (ms_step * 1e-3)
    launches = a.steps * len(chunks)
    kernel_ms = ms_total / launches            # launches are back to back on one stream

    # parity spot check of what the timed region produced (last chunk) against the C oracle
    verified = None
    if rank == 0:
        from oracle import cref
        cg = cref.CGraph(graph_bytes)
        lo, hi = chunks[-1]
        rows = sorted({0, (hi - lo) // 2, hi - lo - 1})
        got = d_out[rows].cpu().numpy().reshape(len(rows), W, 32)
        src = d_in[[lo + r for r in rows]].cpu().numpy().reshape(len(rows), I, 32)
        verified = bool((cg.evaluate_batch(src, min(len(rows), os.cpu_count() or 1)) == got).all())

    # end to end through the C ABI with HOST buffers (pinned), H2D + kernel + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        # per-rank sample: the ranks of one box share its host memory (pinned) and its PCIe root, so the sample is
        # divided among them and bounded by a third of the memory that is available right now
        n_e = min(max(a.e2e_sets // world, 4096), B)
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable:")][0]
            n_e = max(256, min(n_e, int(avail / 3 / world / (W * 32))))
        except (OSError, IndexError, ValueError):
            pass
        del d_out
        torch.cuda.empty_cache()
        h_in = torch.from_numpy(host_in[:min(n_unique, n_e)].reshape(-1, I * 32)).repeat((n_e + n_unique - 1) // n_unique, 1)[:n_e].contiguous().pin_memory()
        h_out = torch.empty((n_e, W * 32), dtype=torch.uint8).pin_memory()
        g.calc_witness_batch_ptr(h_in.data_ptr(), min(n_e, 512), h_out.data_ptr(), first_device=local_rank)          # warm-up (allocates staging)
        g.calc_witness_batch_ptr(h_in.data_ptr(), n_e, h_out.data_ptr(), first_device=local_rank)
        barrier()
        reps = 2
        t0 = time.perf_counter()
        for _ in range(reps):
            g.calc_witness_batch_ptr(h_in.data_ptr(), n_e, h_out.data_ptr(), first_device=local_rank)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n_e / float(tt.item()), "unit": unit, "h2d_bytes_per_step": n_e * I * 32,
               "d2h_bytes_per_step": n_e * W * 32, "sets_per_step": n_e,
               "api": "gw_calc_witness_batch (pinned host buffers, chunked double-buffered H2D/kernel/D2H)"}
        if rank == 0 and verified:
            verified = bool((h_out[0].numpy() == got[0].reshape(-1)).all()) if rows[0] == 0 and chunks[-1][0] % n_unique == 0 else verified

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # rooflines.  Integer pipe: algorithmic IMADs (264 per field multiplication on 8x32-bit
    # limbs; a Div costs `mul_per_div` multiplications with the inversion actually used) against the
    # IMAD rate measured on this GPU by the library's microbenchmark.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    imad_lo = cwc.microbench_imad(local_rank, 1)
    imad_wide = cwc.microbench_imad(local_rank, 0)
    # Div = one multiplication + one safegcd inversion: 20 rounds x 92 signed 32x32->64 multiply-adds
    # (update_de 56 + update_fg 36) = 3680 IMAD-equivalents; the 600 divsteps themselves use no multiplier.
    mul_per_div = 1.0 + 20 * 92 * 2 / 264.0
    mul_equiv = info["n_mul"] + info["n_div"] * mul_per_div
    imad_per_witness = 264.0 * mul_equiv
    per_launch_sets = B / len(chunks)
    achieved_imad = imad_per_witness * per_launch_sets / (kernel_ms * 1e-3)
    bytes_per_witness = 32.0 * (I - 1 + W)
    achieved_gbs = bytes_per_witness * per_launch_sets / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed one-pass ncu capture (bytes per input set x sets per launch)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(a.circuit)
        if tj:
            traffic, traffic_src = tj["dram_bytes_per_set"] * per_launch_sets, tj["source"]
    except Exception:
        pass
    roofline = {"bound": "imad", "achieved": achieved_imad / 1e12, "peak": imad_lo / 1e12, "unit": "TIMAD/s",
                "frac": achieved_imad / imad_lo, "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 32.0 * (I - 1 + W) * per_launch_sets,
                "note": "algorithmic 264 IMAD per field mul x (n_mul + n_div*mul_per_div); peak = mad.lo.u32 rate measured "
                        "live by gw_microbench_imad; IMAD.WIDE.U32 carry-row rate also measured",
                "imad_wide_peak_Tops": imad_wide / 1e12, "mul_per_div": mul_per_div}
    roofline_hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                    "traffic": traffic, "peak_source": hbm_src,
                    "note": "algorithmic bytes = 32*(I-1+W) per witness (inputs read + witness written)"}

    cpu = None
    if not a.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        n_sets = a.cpu_sample or 96 * cores
        v, dt, cores, _ = cpu_reference_run(a.circuit, graph_bytes, I, g.input_signals, n_sets, 1, 1, 9)
        cpu = {"value": v, "unit": unit, "cores": cores, "kind": "port",
               "sample": f"{n_sets} input sets of the same synthetic batch, oracle/ref_eval.c (C restatement of the reference "
                         f"algorithm, graph parsed once), one witness per thread, {dt:.2f} s"}

    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (8x32-bit limbs, canonical BN254 scalars)", "data": "synthetic",
        "config": {"workload": f"{a.circuit} (iden3 authV2(40,64)): {B} input sets per GPU and step, {len(chunks)} launches: "
                               f"{len(chunks) - 1} x {chunk} + {chunks[-1][1] - chunks[-1][0]} sets", "graph_nodes": info["n_nodes"], "node_ops": info["n_ops"], "inputs_len": I,
                   "witness_len": W, "unique_input_sets": n_unique,
                   "l2": f"every launch writes {chunk * W * 32 / 1e9:.1f} GB of witness (>> 126 MB L2), no flush needed",
                   "regs_per_witness": info["n_regs"], "spill_slots": info["n_spill"]},
        "extra": {"node_ops_per_s": value * info["n_ops"], "field_mul_per_s": value * info["n_mul"],
                  "kernel_ms_per_launch": kernel_ms, "verified_against_oracle": verified},
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
    }
    if world > 1:
        dist.destroy_process_group()
    emit(line)


if __name__ == "__main__":
    main()
