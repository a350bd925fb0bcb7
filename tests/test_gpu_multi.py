"""GPU, two or more devices: the REAL multi-GPU run (one process per GPU, NCCL over NVLink).  Every rank assembles its
element block; after the halo exchange -- the library's own ncclSend / ncclRecv group (csrc/comm.cu) and, for comparison,
torch.distributed P2P -- its owned column slab must be the single-GPU CSC bit for bit in pattern and to 1e-12 in values,
and its residual slice likewise.  Skipped on a one-GPU box (tests/test_gpu_halo.py plays the ranks on one device there)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    (3, [4, 4, 6], "PK", 2, 3, 4, "elast", [1.3, 0.7]),
    (3, [5, 5, 6], "PK", 1, 1, 2, "laplace", [2.0]),
    (3, [2, 2, 4], "QK", 2, 3, 6, "svk", [1.0, 1.0]),
]


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s%d" % (c[6], c[2], c[3]))
def test_owned_slabs_over_nccl_match_the_single_gpu_assembly(case):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    port = 29500 + os.getpid() % 2000
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "_multi_worker.py"),
                          json.dumps(case)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    lines = [json.loads(l[6:]) for l in out.stdout.splitlines() if l.startswith("MULTI ")]
    assert len(lines) == world, out.stdout[-2000:]
    assert any(r[mode]["sends"] for r in lines for mode in ("library", "torch")), "no rank sends anything: nothing exercised"
    for r in lines:
        for mode in ("library", "torch"):
            res = r[mode]
            assert res["cols_ok"] and res["rows_ok"], (mode, r)
            if res["own"][1] > res["own"][0]:
                assert res["rel_K"] < 1e-12 and res["rel_R"] < 1e-12, (mode, r)
        # the two transports move the same bytes: identical slabs
        if "slab_checksum" in r["library"]:
            assert r["library"]["slab_checksum"] == r["torch"]["slab_checksum"], r
