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
