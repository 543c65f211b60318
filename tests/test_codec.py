"""Graph file codec (wtns.graph.001) and .wtns framing: golden bytes derived by hand from
protos/messages.proto, the round trips of src/storage.rs:345-465, an independent decode with the
google.protobuf runtime, and the C++ codec of the product library against the Python one."""
import struct

import pytest

from tests import util
from tests.util import M, po


def test_golden_bytes_from_the_proto_schema():
    # Node{input{idx=1}}: field 1 (LEN) -> 0A 02 ; InputNode.idx=1 -> 08 01
    assert po.encode_node((po.K_INPUT, 1)) == bytes([0x0A, 0x02, 0x08, 0x01])
    # proto3 omits zero scalars: Input(0) is an empty InputNode
    assert po.encode_node((po.K_INPUT, 0)) == bytes([0x0A, 0x00])
    # Constant 1: Node.constant(2) > ConstantNode.value(1) > BigUInt.valueLE(1) = 01
    assert po.encode_node((po.K_CONST, 1)) == bytes([0x12, 0x05, 0x0A, 0x03, 0x0A, 0x01, 0x01])
    # zero constant is one 00 byte (num-bigint to_bytes_le)
    assert po.encode_node((po.K_CONST, 0)) == bytes([0x12, 0x05, 0x0A, 0x03, 0x0A, 0x01, 0x00])
    # UnoOp Id(=1) a=4
    assert po.encode_node((po.K_UNO, 1, 4)) == bytes([0x1A, 0x04, 0x08, 0x01, 0x10, 0x04])
    # DuoOp Mul(=0, omitted) a=5 b=6
    assert po.encode_node((po.K_DUO, 0, 5, 6)) == bytes([0x22, 0x04, 0x10, 0x05, 0x18, 0x06])
    # TresOp TernCond(=0) a=7 b=8 c=9
    assert po.encode_node((po.K_TRES, 0, 7, 8, 9)) == bytes([0x2A, 0x06, 0x10, 0x07, 0x18, 0x08, 0x20, 0x09])
    # indices above 127 use multi-byte varints
    assert po.encode_node((po.K_DUO, 16, 300, 1)) == bytes([0x22, 0x07, 0x08, 0x10, 0x10, 0xAC, 0x02, 0x18, 0x01])


def test_file_layout_and_round_trip_like_storage_rs():
    # src/storage.rs:421-465 test_deserialize_inputs (operands rewritten to be backward references)
    nodes = [(po.K_INPUT, 0), (po.K_CONST, 1), (po.K_UNO, 1, 0), (po.K_DUO, 0, 1, 2), (po.K_TRES, 0, 1, 2, 3)]
    wit = [4, 1]
    imap = {"sig1": (1, 3), "sig2": (5, 1)}
    data = po.serialize_graph(nodes, wit, imap)
    assert data[:14] == b"wtns.graph.001"
    assert struct.unpack_from("<Q", data, 14)[0] == 5             # u64 node count (storage.rs:145,228)
    n2, w2, i2 = po.deserialize_graph(data)
    assert (n2, w2, i2) == (nodes, wit, imap)
    # trailing u64 = offset of the metadata record (storage.rs:180)
    off = struct.unpack_from("<Q", data, len(data) - 8)[0]
    ln, p = po._get_varint(data, off)
    assert off + (p - off) + ln == len(data) - 8
    # the C++ codec parses it and re-serialises to something that decodes identically
    g = util.SimGraph(data)
    assert po.deserialize_graph(g.reserialize()) == (nodes, wit, imap)


def test_read_message_framing_like_storage_rs():
    # src/storage.rs:317-342: two length-delimited Input nodes back to back
    buf = bytearray()
    for idx in (1, 2):
        m = po.encode_node((po.K_INPUT, idx))
        po._put_varint(buf, len(m))
        buf += m
    pos = 0
    got = []
    for _ in range(2):
        ln, pos = po._get_varint(bytes(buf), pos)
        got.append(po.decode_node(bytes(buf[pos:pos + ln])))
        pos += ln
    assert got == [(po.K_INPUT, 1), (po.K_INPUT, 2)] and pos == len(buf)


def test_constants_reduced_mod_m_and_decoder_leniency():
    # constants >= M are reduced on load (Fr::from_le_bytes_mod_order, storage.rs:28)
    raw = bytearray(b"wtns.graph.001") + struct.pack("<Q", 1)
    big = (M + 5).to_bytes(32, "little")
    inner = bytes([0x0A, 0x22, 0x0A, 0x20]) + big
    msg = bytes([0x12, len(inner)]) + inner
    raw += bytes([len(msg)]) + msg
    meta = bytes([0x08, 0x00])                                   # witnessSignals unpacked encoding: [0]
    raw += bytes([len(meta)]) + meta + struct.pack("<Q", len(raw))
    nodes, wit, _ = po.deserialize_graph(bytes(raw))
    assert nodes == [(po.K_CONST, 5)] and wit == [0]
    g = util.SimGraph(bytes(raw))
    assert g.eval([1])[0] == [5]


def test_independent_decode_with_protobuf_runtime():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="m.proto", package="t", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=1)
            if tname:
                f.type_name = ".t." + tname
        return m
    T = descriptor_pb2.FieldDescriptorProto
    msg("BigUInt", [("valueLE", 1, T.TYPE_BYTES, None)])
    msg("InputNode", [("idx", 1, T.TYPE_UINT32, None)])
    msg("ConstantNode", [("value", 1, T.TYPE_MESSAGE, "BigUInt")])
    msg("UnoOpNode", [("op", 1, T.TYPE_UINT32, None), ("aIdx", 2, T.TYPE_UINT32, None)])
    msg("DuoOpNode", [("op", 1, T.TYPE_UINT32, None), ("aIdx", 2, T.TYPE_UINT32, None), ("bIdx", 3, T.TYPE_UINT32, None)])
    msg("TresOpNode", [("op", 1, T.TYPE_UINT32, None), ("aIdx", 2, T.TYPE_UINT32, None), ("bIdx", 3, T.TYPE_UINT32, None),
                       ("cIdx", 4, T.TYPE_UINT32, None)])
    msg("Node", [("input", 1, T.TYPE_MESSAGE, "InputNode"), ("constant", 2, T.TYPE_MESSAGE, "ConstantNode"),
                 ("unoOp", 3, T.TYPE_MESSAGE, "UnoOpNode"), ("duoOp", 4, T.TYPE_MESSAGE, "DuoOpNode"),
                 ("tresOp", 5, T.TYPE_MESSAGE, "TresOpNode")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    Node = message_factory.GetMessageClass(pool.FindMessageTypeByName("t.Node"))
    data = util.golden_graph("circuit2")
    nodes, _, _ = po.deserialize_graph(data)
    pos = 22
    for want in nodes:
        ln, pos = po._get_varint(data, pos)
        n = Node()
        n.ParseFromString(data[pos:pos + ln])
        pos += ln
        if want[0] == po.K_INPUT:
            assert n.HasField("input") and n.input.idx == want[1]
        elif want[0] == po.K_CONST:
            assert int.from_bytes(n.constant.value.valueLE, "little") == want[1]
        elif want[0] == po.K_UNO:
            assert (n.unoOp.op, n.unoOp.aIdx) == want[1:]
        elif want[0] == po.K_DUO:
            assert (n.duoOp.op, n.duoOp.aIdx, n.duoOp.bIdx) == want[1:]
        else:
            assert (n.tresOp.op, n.tresOp.aIdx, n.tresOp.bIdx, n.tresOp.cIdx) == want[1:]
        # and the protobuf runtime's own encoding equals ours
        assert n.SerializeToString() == po.encode_node(want)


@pytest.mark.parametrize("name", ["circuit5_poseidon", "circuit11_key_expansion", "circuit9_authV2"])
def test_cpp_codec_round_trip_on_golden_graphs(name):
    data = util.golden_graph(name)
    g = util.SimGraph(data)
    assert po.deserialize_graph(g.reserialize()) == po.deserialize_graph(data)


def test_malformed_graphs_are_errors_not_crashes():
    good = util.golden_graph("circuit2")
    bad_magic = b"wtns.graph.002" + good[14:]
    fwd = po.serialize_graph([(po.K_INPUT, 0), (po.K_DUO, 0, 0, 5)], [0], {})     # forward operand reference
    bad_op = po.serialize_graph([(po.K_INPUT, 0), (po.K_DUO, 25, 0, 0)], [0], {})
    bad_wit = po.serialize_graph([(po.K_INPUT, 0)], [3], {})
    for blob in (b"", b"short", bad_magic, good[:40], good[:len(good) // 2], fwd, bad_op, bad_wit):
        with pytest.raises(ValueError):
            util.SimGraph(blob)
    for blob in (bad_magic, good[:40]):
        with pytest.raises((ValueError, EOFError)):
            po.deserialize_graph(blob)


def test_wtns_layout():
    w = [1, 2, M - 1]
    data = po.wtns_from_witness(w)
    assert len(data) == 76 + 32 * 3
    assert data[:4] == b"wtns" and struct.unpack_from("<II", data, 4) == (2, 2)       # version forced to 2 (lib.rs:118)
    assert struct.unpack_from("<IQI", data, 12) == (1, 40, 32)
    assert int.from_bytes(data[28:60], "little") == M and struct.unpack_from("<I", data, 60)[0] == 3
    assert struct.unpack_from("<IQ", data, 64) == (2, 96)
    assert po.parse_wtns(data) == w
    import ctypes
    hdr = ctypes.create_string_buffer(76)
    util.sim_lib().sim_wtns_header(3, hdr)
    assert hdr.raw == data[:76]
