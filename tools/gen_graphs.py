#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference's test circuits.

Runs only where /root/reference exists (this container); the GPU box uses the
committed outputs:

  tests/golden/graphs/<circuit>.bin.xz     wtns.graph.001 file (xz), built by tools/circom_frontend.py
                                           from /root/reference/test_circuits/<circuit>.circom
  tests/golden/inputs/<circuit>_inputs.json copy of the reference's input fixture
  tests/golden/wtns/<circuit>.wtns(.xz)     .wtns the Python oracle produces for that input
  tests/golden/manifest.json               sizes, op histograms, sha256 of every artefact, and the
                                           result of re-checking every `===` of the circom sources

Extra graphs built from tiny generated circom files: poseidon2 (the circomlib Poseidon KAT arity).
"""
import hashlib
import json
import lzma
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po           # noqa: E402
from tools import circom_frontend as cf     # noqa: E402

REF = "/root/reference"
LIB = REF + "/test_deps/circomlib/circuits"
CIRCUITS = ["circuit1", "circuit2", "circuit3", "circuit4", "circuit5_poseidon", "circuit6_num2bits",
            "circuit7_poseidon4", "circuit8_sha256_512", "circuit9_authV2", "circuit11_key_expansion"]
EXTRA = {
    "poseidon2": ('pragma circom 2.0.0;\ninclude "poseidon.circom";\ncomponent main = Poseidon(2);\n',
                  {"inputs": ["1", "2"]}),
}


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def main():
    gdir = os.path.join(ROOT, "tests/golden/graphs")
    idir = os.path.join(ROOT, "tests/golden/inputs")
    wdir = os.path.join(ROOT, "tests/golden/wtns")
    for d in (gdir, idir, wdir):
        os.makedirs(d, exist_ok=True)
    manifest = {}
    tmp = tempfile.mkdtemp()
    jobs = [(c, f"{REF}/test_circuits/{c}.circom", f"{REF}/test_circuits/{c}_inputs.json") for c in CIRCUITS]
    for name, (src, inputs) in EXTRA.items():
        p = os.path.join(tmp, name + ".circom")
        open(p, "w").write(src)
        ip = os.path.join(tmp, name + "_inputs.json")
        json.dump(inputs, open(ip, "w"))
        jobs.append((name, p, ip))
    for name, path, inputs_path in jobs:
        inputs_json = open(inputs_path).read()
        chk = po.deserialize_inputs(inputs_json)
        data, nodes, wit, imap, st = cf.build_graph(path, [LIB], chk)
        assert st["constraints_violated"] == 0, (name, st["violations"])
        # round trip through the codec and evaluate with the oracle
        witness = po.calc_witness(inputs_json, data)
        wtns = po.wtns_from_witness(witness)
        with open(os.path.join(gdir, name + ".bin.xz"), "wb") as f:
            f.write(lzma.compress(data, preset=9 | lzma.PRESET_EXTREME))
        shutil.copyfile(inputs_path, os.path.join(idir, name + "_inputs.json"))
        with open(os.path.join(wdir, name + ".wtns.xz"), "wb") as f:
            f.write(lzma.compress(wtns, preset=9))
        st.pop("violations", None)
        manifest[name] = {
            "graph_sha256": sha(data), "graph_bytes": len(data),
            "wtns_sha256": sha(wtns), "wtns_bytes": len(wtns),
            "n_nodes": len(nodes), "n_witness": len(wit),
            "n_inputs": 1 + sum(ln for _, ln in imap.values()),
            "stats": st,
        }
        print(name, json.dumps(manifest[name]["stats"], default=str))
    with open(os.path.join(ROOT, "tests/golden/manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True, default=str)


if __name__ == "__main__":
    main()
