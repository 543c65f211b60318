"""In-tree build of the CUDA library and the calc-witness CLI (nvcc cross-compiles sm_100a without a GPU).

  python circom-witnesscalc_b200/build.py [--force] [--verbose]

Outputs (git-ignored, shipped to the GPU box by gpurun):
  circom-witnesscalc_b200/lib/libcircom_witnesscalc.so   the C-ABI library of include/graph_witness.h
  circom-witnesscalc_b200/bin/calc-witness               CLI with the reference's argv contract
  circom-witnesscalc_b200/bin/calc-witness-batch         batch companion: <graph.bin> <inputs.jsonl> <out_dir> [n_gpus]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libcircom_witnesscalc.so")
# the same library with -DGW_PROFILING: per-packet clocks, timing experiments and the fault injection that the watchdog
# test needs (tests/test_gpu_parity.py); loaded through GW_LIB_PATH only
LIB_PROF = os.path.join(HERE, "lib_variants", "libcwc_prof.so")
BIN = os.path.join(HERE, "bin", "calc-witness")
BIN_BATCH = os.path.join(HERE, "bin", "calc-witness-batch")
# the reference's own C embedding example (examples/calc_witness.c), compiled UNCHANGED from where it lies against this
# library's header: the acceptance test of the drop-in C ABI.  Only built where /root/reference exists (here, not on the
# GPU box, which uses the prebuilt binary); no reference source is copied into the repo.
REF_EXAMPLE_SRC = "/root/reference/examples/calc_witness.c"
BIN_REF_EXAMPLE = os.path.join(HERE, "bin", "ref-example-calc-witness")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["engine.cu", "graph.cpp", "plan.cpp", "bitplan.cpp", "inputs.cpp", "wtns.cpp", "capi.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "graph_witness.h")]
    if force or _newer(LIB, deps):
        cmd = [NVCC, "-shared", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas,-Wno-misleading-indentation", "-std=c++17", "-O3", "-lineinfo", *ARCH,
               "-Xptxas", "-v" if verbose else "-O3", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if force or _newer(LIB_PROF, deps):
        os.makedirs(os.path.dirname(LIB_PROF), exist_ok=True)
        cmd = [NVCC, "-shared", "-Xcompiler", "-fPIC,-O3", "-std=c++17", "-O3", "-lineinfo", *ARCH, "-DGW_PROFILING",
               "-o", LIB_PROF] + [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.check_call(cmd)
    if force or _newer(BIN, [LIB, os.path.join(CSRC, "calc_witness_main.cpp")]):
        cmd = ["g++", "-O2", "-std=c++17", "-o", BIN, os.path.join(CSRC, "calc_witness_main.cpp"),
               "-L" + os.path.dirname(LIB), "-lcircom_witnesscalc", "-Wl,-rpath,$ORIGIN/../lib"]
        subprocess.check_call(cmd)
    if force or _newer(BIN_BATCH, [LIB, os.path.join(CSRC, "calc_witness_batch_main.cpp")]):
        cmd = ["g++", "-O2", "-std=c++17", "-o", BIN_BATCH, os.path.join(CSRC, "calc_witness_batch_main.cpp"),
               "-L" + os.path.dirname(LIB), "-lcircom_witnesscalc", "-Wl,-rpath,$ORIGIN/../lib"]
        subprocess.check_call(cmd)
    header = os.path.join(HERE, "..", "include", "graph_witness.h")
    if os.path.exists(REF_EXAMPLE_SRC) and (force or _newer(BIN_REF_EXAMPLE, [LIB, REF_EXAMPLE_SRC, header])):
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            # the example includes "../include/graph_witness.h": give it OUR header at that relative path
            os.makedirs(os.path.join(tmp, "examples")); os.makedirs(os.path.join(tmp, "include"))
            os.symlink(REF_EXAMPLE_SRC, os.path.join(tmp, "examples", "calc_witness.c"))
            os.symlink(os.path.abspath(header), os.path.join(tmp, "include", "graph_witness.h"))
            subprocess.check_call(["gcc", "-w", "-o", BIN_REF_EXAMPLE, os.path.join(tmp, "examples", "calc_witness.c"),
                                   "-L" + os.path.dirname(LIB), "-lcircom_witnesscalc", "-Wl,-rpath,$ORIGIN/../lib"])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
