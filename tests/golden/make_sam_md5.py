#!/usr/bin/env python
"""Regenerates tests/golden/sam_md5.json: md5 of the UNMODIFIED reference's SAM (oracle/_ref/bwa mem, @PG removed)
on the seeded datasets of tests/test_oracle_pin.py:CASES.  Run from the repo root on a box that has oracle/_ref."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import conftest
from test_oracle_pin import CASES
d = conftest.DataSets()
out = {}
for name, ref, kw, extra in CASES:
    fa, fqs = d.reads(ref, **kw)
    out[name] = conftest.md5(conftest.ref_sam(extra + ["-K", "100000000", "-t", "4", fa] + fqs))
json.dump(out, open(os.path.join(os.path.dirname(__file__), "sam_md5.json"), "w"), indent=1, sort_keys=True)
print(out)
