"""Differential fuzz of the device expression semantics against the oracle (tools/where_fuzz.py): seeded random Filter
conditions and aggregate operands - overflow promotion, cross-type collation, 4-valued logic, BETWEEN / IN / IS tests -
through the generated scan kernel on synthetic documents.  tests/test_tail_fuzz.py does the same for the host evaluator
of the operator's tail on the CPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_filters_and_operands_against_the_oracle():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "where_fuzz.py"), "14", "1"], capture_output=True, text=True, timeout=300)
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert p.returncode == 0 and line["rounds"] == 14 and not line["failures"], line
