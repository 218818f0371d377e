"""BASELINE.json configs 2-5 on device-generated columns (tools/workloads.py), checked through size-independent
properties: every group of the result equals torch reductions (bincount / index_add_ / scatter_reduce_ / a bitmap of the
DISTINCT pairs) over the very same generated chunks - exact for counts, integer sums, min / max, group sets and DISTINCT
sets; float64 SUM / AVG within 1e-12 of exact integer-scaled totals (BASELINE.json north_star tolerance).  tools/full_size.py
and bench.py run the same code at 10 M / 60 M / 200 M / 1 B rows; here the row counts are scaled to keep the suite short
while still crossing every structure the full sizes use (front cache overflow, DISTINCT bitmap, direct tables)."""
import json
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

TOOLS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")


@pytest.fixture(scope="module")
def full_size():
    sys.path.insert(0, TOOLS)
    import query_b200 as q
    q.init(0)
    import full_size as mod
    return mod


@pytest.mark.parametrize("config,scale", [("config2", 0.2), ("config3", 0.02), ("config4", 0.02), ("config5", 0.02)])
def test_baseline_config_against_torch_reductions(full_size, config, scale, capsys, monkeypatch):
    if config == "config4":
        monkeypatch.setenv("N1GPU_SET_PASSES", "2")  # the sliced bitmap of the 200 M-row run, at this size
    full_size.run(config, scale=scale)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["config"] == config and line["groups"] > 0 and "check" in line
