#!/usr/bin/env python3
"""Imports the reference's four preset files (the only fixtures it ships, SURVEY.md section 2
row 6) into presets/ as data fixtures.  The values are unchanged (every number is written with
repr(), which round-trips a double exactly); the files are re-serialised compactly with sorted
keys.  Run in the authoring container, where /root/reference exists; the GPU box only ever sees
the committed copies."""
import json
import os
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/cuda-native"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "presets")
os.makedirs(DST, exist_ok=True)
for name in ("settings", "littlecells", "eater", "pulser"):
    with open(os.path.join(SRC, name + ".json")) as f:
        data = json.load(f)
    with open(os.path.join(DST, name + ".json"), "w") as f:
        json.dump(data, f, sort_keys=True, separators=(",", ":"))
        f.write("\n")
    print(name, len(data), "keys")
