"""Shared helpers of the test-suite: golden fixtures, random graphs, ctypes bindings of the
test-only host libraries (tests/csrc) and of the product library."""
import ctypes
import json
import lzma
import os
import random
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILD = os.path.join(ROOT, "tests", "_build")
CSRC = os.path.join(ROOT, "circom-witnesscalc_b200", "csrc")

from oracle import pyoracle as po  # noqa: E402

M = po.M


# ---- fixtures -------------------------------------------------------------------------------------
def manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_graph(name) -> bytes:
    return lzma.decompress(open(os.path.join(GOLDEN, "graphs", name + ".bin.xz"), "rb").read())


def golden_inputs(name) -> str:
    return open(os.path.join(GOLDEN, "inputs", name + "_inputs.json")).read()


def golden_wtns(name) -> bytes:
    return lzma.decompress(open(os.path.join(GOLDEN, "wtns", name + ".wtns.xz"), "rb").read())


def kat():
    """reference-held known answers per golden graph (tests/golden/kat.json, written by tools/gen_graphs.py):
    [{inputs: {name: value | [values]}, expOut: {main signal: value}, source}]"""
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


def kat_cases():
    return [(g, i) for g, v in sorted(kat().items()) for i in range(len(v))]


def kat_check(witness, input_map, n_inputs_row, case, out_names):
    """expOut of a circom_tester assertOut: main OUTPUT signals sit at witness positions 1.. in declaration order
    (out_names), main INPUT signals after them in input-buffer order."""
    n_out = len(out_names)
    for k, v in case["expOut"].items():
        if k in out_names:
            assert witness[1 + out_names.index(k)] == int(v), (k, case.get("desc"))
        else:
            off, ln = input_map[k]
            assert ln == 1 and witness[n_out + off] == int(v), (k, case.get("desc"))


KAT_OUTPUTS = {"authV2_32_32": ["userID"], "poseidon2": ["out"], "poseidon3": ["out"], "poseidon5": ["out"]}


# ---- packing ----------------------------------------------------------------------------------------
def pack_u256(vals) -> bytes:
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def unpack_u256(buf) -> list:
    buf = bytes(buf)
    return [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]


def random_field_batch(rng: np.random.Generator, shape):
    """uniform values in [0, M) as a uint64 array [..., 4] (little-endian limbs) by rejection sampling"""
    n = int(np.prod(shape))
    out = np.empty((n, 4), dtype=np.uint64)
    top = M >> 192
    filled = 0
    while filled < n:
        cand = rng.integers(0, 1 << 64, size=(2 * (n - filled) + 16, 4), dtype=np.uint64)
        cand[:, 3] &= np.uint64((1 << 62) - 1)      # 254 bits
        hi = cand[:, 3]
        ok = hi < np.uint64(top)                      # strictly below the top limb of M: always < M
        eq = hi == np.uint64(top)
        if eq.any():
            for i in np.nonzero(eq)[0]:
                v = sum(int(cand[i, k]) << (64 * k) for k in range(4))
                ok[i] = v < M
        good = cand[ok]
        take = min(len(good), n - filled)
        out[filled:filled + take] = good[:take]
        filled += take
    return out.reshape(tuple(shape) + (4,))


def limbs_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


# ---- random graphs ----------------------------------------------------------------------------------
EDGE = [0, 1, 2, 3, M - 1, M - 2, M >> 1, (M >> 1) + 1, (M >> 1) - 1, 1 << 253, (1 << 253) - 1, (1 << 254) - 1 - M,
        0xFFFFFFFF, 1 << 32, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 253, 254, 255, 256, 31, 32, 33, 64]


def random_value(rnd: random.Random):
    r = rnd.random()
    if r < 0.3:
        return rnd.choice(EDGE) % M
    if r < 0.45:
        return rnd.randrange(1 << rnd.randrange(1, 254))
    return rnd.randrange(M)


def random_graph(rnd: random.Random, n_inputs=6, n_ops=200, ops=None, n_consts=12, uno_ext=True):
    """A graph in the layout build-circuit produces: Input run first, then constants and ops.
    Returns (nodes, witness_signals, input_map)."""
    duo = ops if ops is not None else list(range(20))
    nodes = [(po.K_INPUT, i) for i in range(n_inputs + 1)]
    for _ in range(n_consts):
        nodes.append((po.K_CONST, random_value(rnd)))
    small = [len(nodes) + i for i in range(6)]
    for v in (0, 1, 5, 31, 32, 200):
        nodes.append((po.K_CONST, v))
    for _ in range(n_ops):
        n = len(nodes)
        pick = lambda: rnd.randrange(n) if rnd.random() < 0.7 else rnd.randrange(max(0, n - 8), n)
        r = rnd.random()
        if r < 0.08:
            # Neg mostly; Id / Lnot / Bnot (north_star's unary ops; the reference snapshot has Neg and Id only,
            # graph.rs:175-178, and Id is unimplemented! at run time) are drawn too when `uno_ext`
            nodes.append((po.K_UNO, rnd.choice((0, 0, 1, 2, 3)) if uno_ext else 0, pick()))
        elif r < 0.16:
            nodes.append((po.K_TRES, 0, pick(), pick(), pick()))
        else:
            op = rnd.choice(duo)
            a, b = pick(), pick()
            if op in (15, 16) and rnd.random() < 0.7:
                b = rnd.choice(small)
            if op == 0 and rnd.random() < 0.2:
                b = a
            nodes.append((po.K_DUO, op, a, b))
    n = len(nodes)
    wit = [0] + [rnd.randrange(n) for _ in range(min(n, 40))] + list(range(n - 10, n))
    return nodes, wit, {"x": (1, n_inputs)}


def poseidon_like_graph(rnd: random.Random, t=3, rounds=8, n_inputs=6):
    """Random graph with the shapes of circomlib's Poseidon (poseidon.circom: Ark, Sigma, Mix, MixS) and an occasional
    division: S-box chains x^2, x^4, x^5 (operand order and witness order varied), `lc += c * in[j]` sums,
    `in[i] + in[0] * c` updates, constants added before the S-box.  Exercises OP_POW5, the OP_DOT shapes, the
    Mul(const, x + c) folding, lane chains and split linear combinations.  Returns (nodes, witness_signals, input_map)."""
    nodes = [(po.K_INPUT, i) for i in range(n_inputs + 1)]

    def const(v):
        nodes.append((po.K_CONST, v))
        return len(nodes) - 1

    def duo(op, a, b):
        nodes.append((po.K_DUO, po.DUO[op], a, b))
        return len(nodes) - 1

    st = [1 + rnd.randrange(n_inputs) for _ in range(t)]
    wit = [0]
    for _ in range(rounds):
        full = rnd.random() < 0.3
        st = [duo("Add", s, const(rnd.randrange(M))) if rnd.random() < 0.8 else s for s in st]
        for j in range(t if full else 1):
            x = st[j]
            x2 = duo("Mul", x, x)
            x4 = duo("Mul", x2, x2)
            x5 = duo("Mul", x4, x) if rnd.random() < 0.8 else duo("Mul", x, x4)
            outs = [x2, x4, x5]
            if rnd.random() < 0.3:
                rnd.shuffle(outs)
            wit += [o for o in outs if rnd.random() < 0.9]
            st[j] = x5
        new = []
        if full or rnd.random() < 0.3:
            for _i in range(t):
                lc = None
                for j in range(t):
                    c = const(rnd.randrange(M))
                    term = duo("Mul", c, st[j]) if rnd.random() < 0.5 else duo("Mul", st[j], c)
                    lc = term if lc is None else duo("Add", lc, term)
                new.append(lc)
        else:
            lc = None
            for j in range(t):
                term = duo("Mul", const(rnd.randrange(M)), st[j])
                lc = term if lc is None else duo("Add", lc, term)
            new.append(lc)
            for i in range(1, t):
                new.append(duo("Add" if rnd.random() < 0.8 else "Sub", st[i], duo("Mul", st[0], const(rnd.randrange(M)))))
        st = new
        wit += [s for s in st if rnd.random() < 0.8]
        if rnd.random() < 0.15:
            d = duo("Div", st[0], st[-1])
            wit.append(d)
            st[0] = d
    wit += st
    return nodes, wit, {"x": (1, n_inputs)}


# ---- test-only host libraries -------------------------------------------------------------------------
def _build(name, sources, extra=()):
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, name)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [s for s in sources]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", out,
                               *sources, *extra])
    return out


_libs = {}


def field_host_lib(emulate_ptx=False):
    """csrc/field.cuh on the host.  emulate_ptx=True compiles the DEVICE formulations (PTX carry chains, even/odd
    accumulators) with an emulated carry flag instead of the plain host loops."""
    key = "field_ptx" if emulate_ptx else "field"
    if key not in _libs:
        _libs[key] = ctypes.CDLL(_build("libfieldhost_ptx.so" if emulate_ptx else "libfieldhost.so",
                                        [os.path.join(ROOT, "tests/csrc/field_host_api.cpp")],
                                        extra=("-DGW_EMULATE_PTX",) if emulate_ptx else ()))
    return _libs[key]


def sim_lib():
    if "sim" not in _libs:
        src = [os.path.join(ROOT, "tests/csrc/plan_host_sim.cpp")] + [os.path.join(CSRC, f) for f in
                                                                      ("graph.cpp", "plan.cpp", "bitplan.cpp", "inputs.cpp", "wtns.cpp")]
        L = ctypes.CDLL(_build("libgwsim.so", src))
        L.sim_load.restype = ctypes.c_void_p
        L.sim_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
        L.sim_load2.restype = ctypes.c_void_p
        L.sim_load2.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        L.sim_free.argtypes = [ctypes.c_void_p]
        L.sim_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.sim_info2.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.sim_eval.restype = ctypes.c_int64
        L.sim_eval.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.sim_eval_latency.restype = ctypes.c_int64
        L.sim_eval_latency.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64)]
        L.sim_eval_latency2.restype = ctypes.c_int64
        L.sim_eval_latency2.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint32)]
        L.sim_inputs.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.sim_reserialize.restype = ctypes.c_size_t
        L.sim_reserialize.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.sim_bit_load.restype = ctypes.c_void_p
        L.sim_bit_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        L.sim_bit_free.argtypes = [ctypes.c_void_p]
        L.sim_bit_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.sim_bit_eval.restype = ctypes.c_int64
        L.sim_bit_eval.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p]
        _libs["sim"] = L
    return _libs["sim"]


class BitSimGraph:
    """bit-sliced plan (csrc/bitplan.cpp) on the host simulator: groups of up to 32 input sets"""

    def __init__(self, data: bytes, max_support=0, merge=True):
        self.L = sim_lib()
        err = ctypes.create_string_buffer(512)
        self.h = self.L.sim_bit_load(data, len(data), max_support, int(merge), err, 512)
        if not self.h:
            raise ValueError(err.value.decode())
        info = (ctypes.c_uint64 * 15)()
        self.L.sim_bit_info(self.h, info)
        keys = ["eligible", "n_steps", "n_slots", "n_luts", "n_levels", "n_bit", "n_tt", "n_bv", "n_full_adders", "n_merged", "n_inputs_checked", "n_consts",
                "has_field_inputs", "n_wide", "plane_stride"]
        self.info = dict(zip(keys, [int(x) for x in info]))
        self.reason = err.value.decode()
        nodes, wit, _ = po.deserialize_graph(data)
        self.I = 1 + max([n[1] for n in nodes if n[0] == po.K_INPUT] + [0])
        self.W = len(wit)

    def eval(self, rows):
        """rows: list (<= 32) of input lists -> (list of witness lists or None where the bit contract fails, ok mask)"""
        n = len(rows)
        out = ctypes.create_string_buffer(32 * self.W * n)
        ok = self.L.sim_bit_eval(self.h, b"".join(pack_u256(r) for r in rows), n, out)
        assert ok >= 0, "malformed bit plan"
        res = [unpack_u256(out.raw[32 * self.W * b:32 * self.W * (b + 1)]) if (ok >> b) & 1 else None for b in range(n)]
        return res, int(ok)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sim_bit_free(self.h)
            self.h = None


class SimGraph:
    """ctypes wrapper over tests/csrc/plan_host_sim.cpp"""

    def __init__(self, data: bytes, n_regs=24, fuse=True, div_batch=0, narrow=True, pow5=True):
        self.L = sim_lib()
        err = ctypes.create_string_buffer(512)
        self.h = self.L.sim_load2(data, len(data), n_regs, int(bool(fuse)) | (0 if narrow else 2) | (0 if pow5 else 4) | (int(div_batch) << 8), err, 512)
        if not self.h:
            raise ValueError(err.value.decode())
        info = (ctypes.c_uint64 * 19)()
        self.L.sim_info(self.h, info)
        keys = ["n_nodes", "I", "W", "n_instrs", "n_regs", "n_spill", "spill_ld", "spill_st", "max_live", "n_consts",
                "live_ops", "graph_ops", "n_dot", "n_dot_mac", "inversions", "div_nodes", "slots", "n_mul", "n_addsub"]
        self.info = dict(zip(keys, [int(x) for x in info]))
        info2 = (ctypes.c_uint64 * 4)()
        self.L.sim_info2(self.h, info2)
        self.info["narrow_instrs"], self.info["widen"], self.info["n_spill_narrow"] = int(info2[0]), int(info2[1]), int(info2[2])
        self.info["pow5"] = int(info2[3])

    def eval(self, inputs):
        """inputs: list of I ints -> (witness list, status bits)"""
        assert len(inputs) == self.info["I"]
        out = ctypes.create_string_buffer(32 * self.info["W"])
        st = self.L.sim_eval(self.h, pack_u256(inputs), out)
        assert st >= 0, "malformed plan"
        return unpack_u256(out.raw), int(st)

    def eval_latency(self, inputs, mode=0, n_warps=0, n_slow_warps=0, slow_levels=0, split_dot=True, fuse=True, packet_slots=0, chain=True,
                     dataflow=False, sbox_links=True):
        """latency-mode plan on the host simulator -> (witness list, info dict).  Level plan: mode 0 / 1 = the slow-warp
        jobs run as early / as late as the protocol allows.  dataflow=True: per-warp packet streams with wait vectors;
        mode 0 / 1 / >= 2 = the next runnable warp is the lowest / the highest / a pseudo-random one
        (tests/csrc/plan_host_sim.cpp)."""
        out = ctypes.create_string_buffer(32 * self.info["W"])
        o8 = (ctypes.c_uint64 * 10)()
        opts = (ctypes.c_uint32 * 9)(n_warps, n_slow_warps, slow_levels, int(split_dot), int(fuse), packet_slots, int(chain), int(dataflow), int(sbox_links))
        rc = self.L.sim_eval_latency2(self.h, pack_u256(inputs), out, o8, mode, opts)
        assert rc == 0, f"latency plan failed ({rc})"
        keys = ["n_levels", "n_slots", "n_instrs", "max_width", "status", "est_cycles", "n_split", "slow_levels", "n_chained", "n_sbox_links"]
        if dataflow:
            keys[6], keys[8] = "n_rows", "n_waits"
        return unpack_u256(out.raw), dict(zip(keys, [int(x) for x in o8]))

    def inputs_from_json(self, js: str):
        buf = ctypes.create_string_buffer(32 * self.info["I"])
        err = ctypes.create_string_buffer(512)
        r = self.L.sim_inputs(self.h, js.encode(), buf, err, 512)
        if r:
            raise ValueError(err.value.decode())
        return unpack_u256(buf.raw)

    def reserialize(self) -> bytes:
        n = self.L.sim_reserialize(self.h, None, 0)
        buf = ctypes.create_string_buffer(n)
        self.L.sim_reserialize(self.h, buf, n)
        return buf.raw

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sim_free(self.h)
            self.h = None
