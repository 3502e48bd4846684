"""Regenerates tests/golden/jsonio_expected.json from the reference's only golden fixture that touches this
repo's "next" row f1 (restart / moment snapshots): utils/iocore/jsonio_expected.json, the metadata file that the
reference's jsonio_test must reproduce (utils/iocore/unittest.py:63-76).  Stored parsed and re-serialised (key order
kept), so tests on the GPU box -- where /root/reference does not exist -- can diff against it.

    python tests/golden/make_jsonio_fixture.py [/root/reference]
"""
import json
import os
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
with open(os.path.join(ref, "utils", "iocore", "jsonio_expected.json")) as f:
    root = json.load(f)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jsonio_expected.json")
with open(out, "w") as f:
    json.dump(root, f, indent=1)
print(out, {k: list(v) for k, v in root.items()})
