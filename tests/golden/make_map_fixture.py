#!/usr/bin/env python3
"""Writes tests/golden/two_rooms.map — the map fixture of tests/test_map_compiler_cpu.py as a file, for the reference-side
emitter tools/rust_golden/src/bin/map_golden.rs (which a maintainer with a Rust toolchain runs once; its output,
tests/golden/reference_map_compiler.json, pins the C++ map compiler against the reference's)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from test_map_compiler_cpu import two_rooms_map  # noqa: E402

if __name__ == "__main__":
    with open(os.path.join(HERE, "two_rooms.map"), "w", encoding="utf-8") as f:
        f.write(two_rooms_map())
