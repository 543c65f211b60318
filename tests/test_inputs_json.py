"""inputs JSON -> input buffer: src/lib.rs:195-247 (deserialize_inputs), :154-181; reference test lib.rs:258-280."""
import pytest

from tests import util
from tests.util import M, po

DOC = '''
    {
        "key1": ["123", "456", 100500],
        "key2": "789",
        "key3": 123123
    }
'''


def _graph():
    nodes = [(po.K_INPUT, i) for i in range(6)] + [(po.K_DUO, 2, 1, 5)]
    return po.serialize_graph(nodes, [0, 1, 2, 3, 4, 5, 6], {"key1": (1, 3), "key2": (4, 1), "key3": (5, 1)})


def test_reference_vector_lib_rs_258():
    want = {"key1": [123, 456, 100500], "key2": [789], "key3": [123123]}
    assert po.deserialize_inputs(DOC) == want
    g = util.SimGraph(_graph())
    assert g.inputs_from_json(DOC) == [1, 123, 456, 100500, 789, 123123]


def test_missing_keys_stay_zero_and_slot0_is_one():
    g = util.SimGraph(_graph())
    assert g.inputs_from_json('{"key2": "5"}') == [1, 0, 0, 0, 5, 0]
    nodes, _, imap = po.deserialize_graph(_graph())
    assert po.build_input_buffer(nodes, imap, {"key2": [5]}) == [1, 0, 0, 0, 5, 0]


def test_big_values_and_reduction():
    g = util.SimGraph(_graph())
    big = (1 << 256) - 1
    buf = g.inputs_from_json('{"key2": "%d", "key3": 18446744073709551615}' % big)
    assert buf[4] == big and buf[5] == (1 << 64) - 1
    w, _ = g.eval(buf)
    assert w[4] == big % M                       # Fr::new reduces inputs >= M (graph.rs:376)


@pytest.mark.parametrize("doc", [
    '[1, 2]',                                  # inputs must be an object
    '{"key2": -1}', '{"key2": 1.5}',           # not a positive integer (lib.rs:211-214)
    '{"key1": [["1"], "2", "3"]}',             # nested arrays rejected (lib.rs:231-233)
    '{"key2": true}', '{"key2": null}', '{"key2": {"a": 1}}',
    '{"key2": "12x"}', '{"key2": "-5"}',       # U256::from_str_radix errors
    '{"key2": "%d"}' % (1 << 256),             # does not fit 256 bits
    '{"key2": "1"', '{"key2" "1"}', '',        # invalid JSON (reference: panic)
    '{"nope": "1"}',                           # unknown key (reference: HashMap index panic, lib.rs:158)
    '{"key1": ["1", "2"]}',                    # wrong length (reference: panic, lib.rs:159-161)
])
def test_rejected_documents(doc):
    g = util.SimGraph(_graph())
    with pytest.raises(ValueError):
        g.inputs_from_json(doc)
    nodes, _, imap = po.deserialize_graph(_graph())
    with pytest.raises(Exception):
        po.build_input_buffer(nodes, imap, po.deserialize_inputs(doc))


def test_string_escapes_and_whitespace():
    g = util.SimGraph(po.serialize_graph([(po.K_INPUT, 0), (po.K_INPUT, 1)], [1], {'a"b': (1, 1)}))
    assert g.inputs_from_json(' {\n "a\\"b" :\t"7" }\n') == [1, 7]
    assert g.inputs_from_json('{"a\\u0022b": ["7"]}') == [1, 7]


# ---- batch input path (gw_inputs_parse_batch): product library, host-only code, no GPU needed ---------------------
def _cwc():
    import importlib
    import subprocess
    import os
    subprocess.check_call(["python", os.path.join(util.ROOT, "circom-witnesscalc_b200", "build.py")])
    return importlib.import_module("circom-witnesscalc_b200")


def test_batch_parse_jsonl_and_array_match_single_record_semantics():
    import json
    import random
    cwc = _cwc()
    g = cwc.Graph(_graph())
    nodes, _, imap = po.deserialize_graph(_graph())
    rnd = random.Random(3)
    recs = []
    for i in range(1000):
        r = {}
        if rnd.random() < 0.9:
            r["key1"] = [str(rnd.randrange(1 << 256)), rnd.randrange(1 << 64), str(rnd.randrange(M))]
        if rnd.random() < 0.7:
            r["key2"] = str(rnd.randrange(M)) if rnd.random() < 0.5 else rnd.randrange(1 << 40)
        if rnd.random() < 0.5:
            r["key3"] = [str(i)]
        recs.append(r)
    want = [po.build_input_buffer(nodes, imap, po.deserialize_inputs(json.dumps(r))) for r in recs]
    jsonl = "\n".join(json.dumps(r) for r in recs) + "\n\n  \n"
    for text in (jsonl, "\r\n".join(json.dumps(r, indent=None) for r in recs), json.dumps(recs, indent=1)):
        for nt in (1, 4, 0):
            arr = g.parse_inputs_batch(text, nt)
            assert arr.shape == (1000, 6, 32)
            got = [util.unpack_u256(arr[i].tobytes()) for i in range(0, 1000, 37)]
            assert got == [want[i] for i in range(0, 1000, 37)]
    assert g.parse_inputs_batch("").shape == (0, 6, 32) and g.parse_inputs_batch("[]").shape == (0, 6, 32)


def test_batch_parse_errors_name_the_record():
    cwc = _cwc()
    g = cwc.Graph(_graph())
    good = '{"key2": "5"}'
    for bad, what in (('{"zzz": "1"}', "unknown input signal"), ('{"key1": ["1"]}', "Invalid input length"),
                      ('{"key2": -1}', "not a positive integer"), ('{"key2": "1"', "invalid JSON"),
                      ('{"key2": "12x"}', "InputFieldNumberParseError")):
        text = "\n".join([good] * 70 + [bad] + [good] * 5)
        with pytest.raises(cwc.WitnessCalcError, match="input set 71: .*" + what):
            g.parse_inputs_batch(text, 3)
    with pytest.raises(cwc.WitnessCalcError, match="unterminated array"):
        g.parse_inputs_batch('[{"key2": "5"}, {"key2": "6"}')


def test_graph_select_compiles_a_pruned_plan_without_gpu():
    cwc = _cwc()
    g = cwc.Graph(util.golden_graph("circuit9_authV2"))
    sel = g.select([0, 1, 2, 1])
    assert sel.n_witness == 4 and sel.n_inputs == g.n_inputs and sel.input_signals == g.input_signals
    assert sel.info["n_instrs"] < g.info["n_instrs"]          # what the selection does not need is dead code
    assert sel.wtns_file_size() == 76 + 4 * 32 and g.wtns_file_size() == 76 + 32 * g.n_witness
    with pytest.raises(cwc.WitnessCalcError, match="out of range"):
        g.select([g.n_witness])
