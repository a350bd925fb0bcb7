"""CPU: the reference arm of bench.py (`--impl reference`: the UNMODIFIED reference from oracle/_ref timed on the host
cores) prints the JSON line the driver expects, and the non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

DRIVER = os.path.join(ROOT, "oracle", "_ref", "gf_ref_driver")


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/gf_ref_driver not built (needs the reference sources)")
def test_reference_arm_json_line():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "assembled_elements_per_s" and line["unit"] == "elements/s" and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["vs_baseline"] is None and "workload" in line["config"]
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/gf_ref_driver not built (needs the reference sources)")
def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout[-500:], out.stderr[-500:])
