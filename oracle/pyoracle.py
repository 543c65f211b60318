"""CPU oracle (Python big-int restatement) of the circom-witnesscalc hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA library under
``circom-witnesscalc_b200/``) may import or call this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg do, and only
as the checker.

What is restated (citations are relative to /root/reference):

* ``evaluate``                -> src/graph.rs:367-391
* ``eval_duo``                -> src/graph.rs:102-144 (Operation::eval_fr), helpers
                                 shl :621-635, shr :637-672, bit_and/or/xor :674-717,
                                 u_lt/u_gt/u_lte/u_gte :723-769, halfM :720
* ``eval_uno``                -> src/graph.rs:188-197
* ``eval_tres``               -> src/graph.rs:221-225
* ``M``                       -> src/field.rs:3-4
* ``deserialize_graph`` / ``serialize_graph`` -> src/storage.rs:137-183, 214-249,
                                 wire schema protos/messages.proto:1-85
* ``deserialize_inputs``      -> src/lib.rs:195-247
* ``calc_witness``            -> src/lib.rs:125-181
* ``wtns_from_witness``       -> src/lib.rs:114-123 (+ wtns-file 0.1.5 layout)

The arithmetic itself lives in crates that are not vendored in the reference
tree (ark-ff 0.4.2 / ark-bn254 0.4.0 Fr, ruint 1.12.3 U256, prost 0.13.3,
wtns-file 0.1.5, Cargo.lock).  Field and integer operations are mathematically
determined, so they are restated from the math on Python integers (all values
are kept CANONICAL, i.e. in [0, M); the reference keeps Montgomery form
internally but every observable result is canonical).

Parity pin: the reference cannot be compiled here (no cargo/rustc), so this
oracle is pinned against the reference's own unit-test vectors
(src/graph.rs:779-883: shl, Div, Idiv, Mod, u_gte; src/lib.rs:258-280 JSON
forms; src/storage.rs:345-465 codec round trips) and against circuit-level
known answers (Poseidon / SHA-256 test vectors of circomlib) in
tests/test_oracle_kat.py.  Ops with no in-tree vector (Mul/Add/Sub/Neg, Shr,
Band/Bor/Bxor, Eq/Neq, Lt/Gt/Leq, Land/Lor, TernCond) are "parity unpinned
in-tree": their only pin is the circuit-level KATs.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, Iterable, List, Sequence, Tuple

# src/field.rs:3-4
M = 21888242871839275222246405745257275088548364400416034343698204186575808495617
# src/graph.rs:720
HALF_M = 10944121435919637611123202872628637544274182200208017171849102093287904247808
MASK256 = (1 << 256) - 1
MASK254 = (1 << 254) - 1

# protos/messages.proto:5-26
DUO_OPS = ["Mul", "Div", "Add", "Sub", "Pow", "Idiv", "Mod", "Eq", "Neq", "Lt", "Gt",
           "Leq", "Geq", "Land", "Lor", "Shl", "Shr", "Bor", "Band", "Bxor"]
DUO = {n: i for i, n in enumerate(DUO_OPS)}
# protos/messages.proto:28-31  (Lnot/Bnot do not exist in this snapshot; they are
# an extension of the new implementation, numbered after the reference's values)
UNO_OPS = ["Neg", "Id", "Lnot", "Bnot"]
UNO = {n: i for i, n in enumerate(UNO_OPS)}
# protos/messages.proto:33-35
TRES_OPS = ["TernCond"]
TRES = {n: i for i, n in enumerate(TRES_OPS)}

# Node kinds.  A node is a tuple:
#   (K_INPUT, idx) | (K_CONST, value) | (K_UNO, op, a) | (K_DUO, op, a, b) | (K_TRES, op, a, b, c)
K_INPUT, K_CONST, K_UNO, K_DUO, K_TRES = 0, 1, 2, 3, 4

GRAPH_MAGIC = b"wtns.graph.001"  # src/storage.rs:16


class ReferenceUndefined(Exception):
    """The reference panics / is unimplemented for these operands."""


# --------------------------------------------------------------------------
# per-op semantics (canonical domain)
# --------------------------------------------------------------------------

def _neg_sign(x: int) -> bool:
    return x > HALF_M


def eval_duo(op: int, a: int, b: int, undefined: str = "raise") -> int:
    """Operation::eval_fr, src/graph.rs:102-144.  a, b, result in [0, M)."""
    if op == 0:   # Mul :105
        return a * b % M
    if op == 1:   # Div :109
        return 0 if b == 0 else a * pow(b, -1, M) % M
    if op == 2:   # Add :110
        return (a + b) % M
    if op == 3:   # Sub :111
        return (a - b) % M
    if op == 4:   # Pow: unimplemented! at runtime (:141-142); build-time meaning :79
        if undefined == "raise":
            raise ReferenceUndefined("Pow")
        return pow(a, b, M)
    if op == 5:   # Idiv :112-116
        return 0 if b == 0 else a // b
    if op == 6:   # Mod :117-121
        return 0 if b == 0 else a % b
    if op == 7:   # Eq :122-125
        return int(a == b)
    if op == 8:   # Neq :126-129
        return int(a != b)
    if op in (9, 10, 11, 12):  # Lt Gt Leq Geq :130-133 -> :723-769
        an, bn = _neg_sign(a), _neg_sign(b)
        if an != bn:
            # different signs: the negative one is the smaller
            return int(an) if op in (9, 11) else int(bn)
        return int({9: a < b, 10: a > b, 11: a <= b, 12: a >= b}[op])
    if op == 13:  # Land :134
        return int(a != 0 and b != 0)
    if op == 14:  # Lor :135
        return int(a != 0 or b != 0)
    if op == 15:  # Shl :136 -> :621-635
        if b == 0:
            return a
        if b >= 254:
            return 0
        r = (a << b) & MASK256          # BigInt::muln truncates to 256 bits
        if r >= M:                      # from_bigint(..).unwrap() panics (:634)
            if undefined == "raise":
                raise ReferenceUndefined("Shl overflow")
            r = (a << b) & MASK254      # circom semantics
            while r >= M:
                r -= M
        return r
    if op == 16:  # Shr :137 -> :637-672
        if b == 0:
            return a
        if b >= 254:
            return 0
        return a >> b
    if op in (17, 18, 19):  # Bor Band Bxor :138-140 -> :674-717
        d = (a | b) if op == 17 else (a & b) if op == 18 else (a ^ b)
        if d > M:
            d -= M
        if d == M:                      # from_bigint(M).unwrap() panics
            if undefined == "raise":
                raise ReferenceUndefined("bitwise result == M")
            d = 0
        return d
    raise ValueError(f"bad duo op {op}")


def eval_uno(op: int, a: int, undefined: str = "raise") -> int:
    """UnoOperation::eval_fr, src/graph.rs:188-197."""
    if op == 0:   # Neg :190-194
        return 0 if a == 0 else M - a
    if undefined == "raise":
        raise ReferenceUndefined(UNO_OPS[op])   # Id: unimplemented! (:195)
    if op == 1:   # Id
        return a
    if op == 2:   # Lnot (extension; circom: !a)
        return int(a == 0)
    if op == 3:   # Bnot (extension; circom: (~a & mask254) mod M)
        r = (~a) & MASK254
        while r >= M:
            r -= M
        return r
    raise ValueError(f"bad uno op {op}")


def eval_tres(op: int, a: int, b: int, c: int) -> int:
    """TresOperation::eval_fr, src/graph.rs:221-225."""
    if op != 0:
        raise ValueError(f"bad tres op {op}")
    return c if a == 0 else b


def evaluate(nodes: Sequence[tuple], inputs: Sequence[int], outputs: Sequence[int],
             undefined: str = "raise", return_values: bool = False):
    """graph::evaluate, src/graph.rs:367-391.

    ``inputs`` are raw 256-bit integers (reduced mod M on load like Fr::new, :376);
    constants are reduced mod M as well (storage.rs:28).  Returns canonical ints.
    """
    values: List[int] = []
    ap = values.append
    for node in nodes:
        k = node[0]
        if k == K_DUO:
            ap(eval_duo(node[1], values[node[2]], values[node[3]], undefined))
        elif k == K_CONST:
            ap(node[1] % M)
        elif k == K_INPUT:
            ap(inputs[node[1]] % M)
        elif k == K_UNO:
            ap(eval_uno(node[1], values[node[2]], undefined))
        elif k == K_TRES:
            ap(eval_tres(node[1], values[node[2]], values[node[3]], values[node[4]]))
        else:
            raise ValueError("bad node kind")
    out = [values[i] for i in outputs]
    if return_values:
        return out, values
    return out


# --------------------------------------------------------------------------
# protobuf wire helpers (proto3, hand-rolled: no protoc in the image)
# --------------------------------------------------------------------------

def _put_varint(buf: bytearray, v: int) -> None:
    while v >= 0x80:
        buf.append((v & 0x7F) | 0x80)
        v >>= 7
    buf.append(v)


def _get_varint(data: bytes, pos: int) -> Tuple[int, int]:
    shift = 0
    v = 0
    while True:
        if pos >= len(data):
            raise EOFError("truncated varint")
        b = data[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _fields(data: bytes) -> Iterable[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) of one message."""
    pos = 0
    n = len(data)
    while pos < n:
        key, pos = _get_varint(data, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(data, pos)
        elif wt == 2:
            ln, pos = _get_varint(data, pos)
            if pos + ln > n:
                raise EOFError("truncated field")
            v = data[pos:pos + ln]
            pos += ln
        elif wt == 1:
            v = data[pos:pos + 8]
            pos += 8
        elif wt == 5:
            v = data[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _uint_fields(data: bytes, nmax: int) -> List[int]:
    out = [0] * (nmax + 1)
    for fno, wt, v in _fields(data):
        if wt == 0 and fno <= nmax:
            out[fno] = v & 0xFFFFFFFF
    return out


def _kv(buf: bytearray, fno: int, v: int) -> None:
    if v:                       # proto3: zero-valued scalars are omitted
        _put_varint(buf, fno << 3)
        _put_varint(buf, v)


def _ld(buf: bytearray, fno: int, payload: bytes) -> None:
    _put_varint(buf, (fno << 3) | 2)
    _put_varint(buf, len(payload))
    buf += payload


def encode_node(node: tuple) -> bytes:
    """proto Node message (without the length prefix). messages.proto:37-75."""
    k = node[0]
    inner = bytearray()
    if k == K_INPUT:
        _kv(inner, 1, node[1])
        fno = 1
    elif k == K_CONST:
        v = node[1] % M
        le = v.to_bytes(max(1, (v.bit_length() + 7) // 8), "little")  # num-bigint to_bytes_le
        big = bytearray()
        _ld(big, 1, le)
        _ld(inner, 1, bytes(big))
        fno = 2
    elif k == K_UNO:
        _kv(inner, 1, node[1]); _kv(inner, 2, node[2])
        fno = 3
    elif k == K_DUO:
        _kv(inner, 1, node[1]); _kv(inner, 2, node[2]); _kv(inner, 3, node[3])
        fno = 4
    elif k == K_TRES:
        _kv(inner, 1, node[1]); _kv(inner, 2, node[2]); _kv(inner, 3, node[3]); _kv(inner, 4, node[4])
        fno = 5
    else:
        raise ValueError("bad node kind")
    out = bytearray()
    _ld(out, fno, bytes(inner))
    return bytes(out)


def decode_node(msg: bytes) -> tuple:
    """From<proto::Node> for Node, src/storage.rs:20-48."""
    node = None
    for fno, wt, v in _fields(msg):
        if wt != 2:
            continue
        if fno == 1:
            f = _uint_fields(v, 1)
            node = (K_INPUT, f[1])
        elif fno == 2:
            val = None
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    val = 0
                    for f3, w3, v3 in _fields(v2):
                        if f3 == 1 and w3 == 2:
                            val = int.from_bytes(v3, "little")
            if val is None:
                raise ValueError("ConstantNode without value")  # reference: unwrap panic
            node = (K_CONST, val % M)   # Fr::from_le_bytes_mod_order, storage.rs:28
        elif fno == 3:
            f = _uint_fields(v, 2)
            if f[1] >= len(UNO_OPS):
                raise ValueError("unknown UnoOp")
            node = (K_UNO, f[1], f[2])
        elif fno == 4:
            f = _uint_fields(v, 3)
            if f[1] >= len(DUO_OPS):
                raise ValueError("unknown DuoOp")
            node = (K_DUO, f[1], f[2], f[3])
        elif fno == 5:
            f = _uint_fields(v, 4)
            if f[1] >= len(TRES_OPS):
                raise ValueError("unknown TresOp")
            node = (K_TRES, f[1], f[2], f[3], f[4])
    if node is None:
        raise ValueError("empty Node message")   # reference: value.node.unwrap() panic
    return node


def serialize_graph(nodes: Sequence[tuple], witness_signals: Sequence[int],
                    input_signals: Dict[str, Tuple[int, int]]) -> bytes:
    """serialize_witnesscalc_graph, src/storage.rs:137-183."""
    out = bytearray(GRAPH_MAGIC)
    out += struct.pack("<Q", len(nodes))
    for node in nodes:
        msg = encode_node(node)
        _put_varint(out, len(msg))
        out += msg
    meta_off = len(out)
    meta = bytearray()
    if witness_signals:
        packed = bytearray()
        for s in witness_signals:
            _put_varint(packed, s)
        _ld(meta, 1, bytes(packed))
    for name, (off, ln) in input_signals.items():
        sd = bytearray()
        _kv(sd, 1, off); _kv(sd, 2, ln)
        entry = bytearray()
        _ld(entry, 1, name.encode("utf-8"))
        _ld(entry, 2, bytes(sd))
        _ld(meta, 2, bytes(entry))
    _put_varint(out, len(meta))
    out += meta
    out += struct.pack("<Q", meta_off)
    return bytes(out)


def deserialize_graph(data: bytes):
    """deserialize_witnesscalc_graph, src/storage.rs:214-249.

    Returns (nodes, witness_signals, input_signals{name: (offset, len)}).
    """
    if len(data) < len(GRAPH_MAGIC) + 8:
        raise EOFError("graph file too short")
    if data[:len(GRAPH_MAGIC)] != GRAPH_MAGIC:
        raise ValueError("Invalid magic")
    pos = len(GRAPH_MAGIC)
    (n,) = struct.unpack_from("<Q", data, pos)
    pos += 8
    nodes = []
    for _ in range(n):
        ln, pos = _get_varint(data, pos)
        if pos + ln > len(data):
            raise EOFError("Unexpected EOF")
        nodes.append(decode_node(data[pos:pos + ln]))
        pos += ln
    ln, pos = _get_varint(data, pos)
    if pos + ln > len(data):
        raise EOFError("Unexpected EOF")
    meta = data[pos:pos + ln]
    witness: List[int] = []
    inputs: Dict[str, Tuple[int, int]] = {}
    for fno, wt, v in _fields(meta):
        if fno == 1 and wt == 2:       # packed
            p = 0
            while p < len(v):
                x, p = _get_varint(v, p)
                witness.append(x)
        elif fno == 1 and wt == 0:     # unpacked
            witness.append(v)
        elif fno == 2 and wt == 2:
            name, off, sl = "", 0, 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    name = bytes(v2).decode("utf-8")
                elif f2 == 2 and w2 == 2:
                    f = _uint_fields(v2, 2)
                    off, sl = f[1], f[2]
            inputs[name] = (off, sl)
    return nodes, witness, inputs


# --------------------------------------------------------------------------
# API glue: inputs JSON, input buffer, .wtns
# --------------------------------------------------------------------------

class InputsError(Exception):
    pass


def _parse_u256_dec(s: str) -> int:
    # ruint U256::from_str_radix(s, 10): digits only (underscores are skipped by ruint),
    # error on overflow past 2^256.
    if not isinstance(s, str):
        raise InputsError("not a string")
    t = s.replace("_", "")
    if t == "" or not all("0" <= ch <= "9" for ch in t):
        raise InputsError(f"invalid digit in {s!r}")
    v = int(t)
    if v >> 256:
        raise InputsError("number does not fit 256 bits")
    return v


def _parse_scalar(v) -> int:
    if isinstance(v, bool):
        raise InputsError("bool is not a signal value")
    if isinstance(v, str):
        return _parse_u256_dec(v)
    if isinstance(v, int):
        if v < 0 or v >= 1 << 64:
            raise InputsError("signal value is not a positive integer")
        return v
    raise InputsError("inputs must be a string")


def deserialize_inputs(data) -> Dict[str, List[int]]:
    """deserialize_inputs, src/lib.rs:195-247."""
    if isinstance(data, (bytes, bytearray)):
        data = data.decode("utf-8")
    v = json.loads(data)
    if not isinstance(v, dict):
        raise InputsError("inputs must be an object")
    out: Dict[str, List[int]] = {}
    for k, val in v.items():
        if isinstance(val, list):
            out[k] = [_parse_scalar(x) for x in val]   # nested arrays rejected (:231-233)
        elif isinstance(val, (str, int)) and not isinstance(val, bool):
            out[k] = [_parse_scalar(val)]
        else:
            raise InputsError(f"value for key {k} must be a number, a string or an array of those")
    return out


def get_inputs_size(nodes: Sequence[tuple]) -> int:
    """src/lib.rs:138-152."""
    start = False
    mx = 0
    for node in nodes:
        if node[0] == K_INPUT:
            mx = max(mx, node[1])
            start = True
        elif start:
            break
    return mx + 1


def build_input_buffer(nodes, input_map, inputs: Dict[str, List[int]]) -> List[int]:
    """get_inputs_buffer + populate_inputs, src/lib.rs:154-181."""
    buf = [0] * get_inputs_size(nodes)
    buf[0] = 1
    for key, vals in inputs.items():
        if key not in input_map:
            raise InputsError(f"unknown input signal {key}")        # reference: panic
        off, ln = input_map[key]
        if ln != len(vals):
            raise InputsError(f"Invalid input length for {key}")    # reference: panic
        if off + ln > len(buf):
            raise InputsError(f"input {key} out of range")          # reference: panic
        buf[off:off + ln] = vals
    return buf


def calc_witness(inputs_json, graph_data: bytes, undefined: str = "raise") -> List[int]:
    """calc_witness, src/lib.rs:125-136."""
    inputs = deserialize_inputs(inputs_json)
    nodes, signals, input_map = deserialize_graph(graph_data)
    buf = build_input_buffer(nodes, input_map, inputs)
    return evaluate(nodes, buf, signals, undefined)


def wtns_from_witness(witness: Sequence[int]) -> bytes:
    """wtns_from_witness, src/lib.rs:114-123; wtns-file 0.1.5 / snarkjs layout, version 2."""
    out = bytearray(b"wtns")
    out += struct.pack("<II", 2, 2)
    out += struct.pack("<IQ", 1, 40)
    out += struct.pack("<I", 32) + M.to_bytes(32, "little") + struct.pack("<I", len(witness))
    out += struct.pack("<IQ", 2, 32 * len(witness))
    for w in witness:
        out += int(w).to_bytes(32, "little")
    return bytes(out)


def parse_wtns(data: bytes) -> List[int]:
    assert data[:4] == b"wtns"
    ver, nsec = struct.unpack_from("<II", data, 4)
    assert (ver, nsec) == (2, 2)
    sid, ssz = struct.unpack_from("<IQ", data, 12)
    assert (sid, ssz) == (1, 40)
    n8 = struct.unpack_from("<I", data, 24)[0]
    assert n8 == 32 and int.from_bytes(data[28:60], "little") == M
    nw = struct.unpack_from("<I", data, 60)[0]
    sid, ssz = struct.unpack_from("<IQ", data, 64)
    assert (sid, ssz) == (2, 32 * nw)
    return [int.from_bytes(data[76 + 32 * i:108 + 32 * i], "little") for i in range(nw)]
