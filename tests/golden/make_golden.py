"""Regenerates tests/golden/oracle_kat.json: sha256 digests of oracle outputs on seeded inputs.

PARITY UNPINNED: these are self-generated known answers of the CPU restatement (the reference has no tests or golden
vectors); they pin the restatement against compiler / platform drift and let the GPU tests cross-check
the oracle built on the GPU box.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from test_oracle_cpu import kat_cases  # noqa: E402

if __name__ == "__main__":
    digests = kat_cases(oracle)
    json.dump({"generator": "tests/golden/make_golden.py", "reference": "Panzerschrek/Square-Wheel @4820f17 (restated)", "digests": digests},
              open(os.path.join(HERE, "oracle_kat.json"), "w"), indent=1, sort_keys=True)
    print(f"wrote {len(digests)} digests")
