"""Regenerates the golden fixtures under tests/golden/ from the reference tree.

Run in the dev container only (needs /root/reference):  python tests/golden/make_goldens.py

* lattice.ark.txt / lattice.char.ark.txt: the reference's own example lattices
  (kwsbin2/egs/), the inputs of its README worked examples.  They are data
  fixtures, the only golden vectors the reference holds for the hot path.
* README_goldens.json: the expected stdout of the five README commands; checked
  here against the README text so a typo in the JSON cannot pass silently.
"""
import json
import os
import shutil

REF = "/root/reference/kwsbin2"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    for f in ("lattice.ark.txt", "lattice.char.ark.txt"):
        shutil.copyfile(os.path.join(REF, "egs", f), os.path.join(HERE, f))
    readme = open(os.path.join(REF, "README.md")).read()
    g = json.load(open(os.path.join(HERE, "README_goldens.json")))
    for k, v in g.items():
        if k.startswith("_") or k == "state_times":
            continue
        assert v in readme, "golden %s does not appear verbatim in the reference README" % k
    print("fixtures refreshed; %d goldens verified against README.md" % 5)
