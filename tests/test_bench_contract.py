"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the reference's own
AMICI/CVODES path on the host cores) prints ONE JSON line with the keys the driver reads, and the
counting models of the roofline give the documented figures on known inputs."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line(oracle_available):
    env = dict(os.environ, OMP_NUM_THREADS="1")  # as torchrun exports it: the arm must not depend on it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "boehm", "--conditions", "48"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference"
    # every host core is used whatever OMP_NUM_THREADS says (round 1: torchrun's OMP_NUM_THREADS=1 made the
    # reference arm single-threaded at N > 1 while it reported all cores)
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert "%d OpenMP threads" % line["cpu_baseline"]["cores"] in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--workload", "boehm", "--conditions", "8"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_counts_follow_the_survey_formulas():
    sys.path.insert(0, ROOT)
    import bench
    from parpe_b200.tables import STAT_NAMES
    stats = np.zeros((2, 20), dtype=np.int64)
    vals = dict(nst=100, nfe=220, nsetups=15, nje=3, nni=110, nniS=105, nfSe=945, qsum=400)
    for k, v in vals.items():
        stats[:, STAT_NAMES.index(k)] = v
    ops = dict(xdot=46, J=32, dxdotdp=48, nnzJ=21)
    nx, npl = 8, 9
    flops, mean = bench.algorithmic_flops(stats, ops, nx, npl)
    q = 4.0
    F = (220 * 46 + 3 * 32 + 15 * (2.0 / 3.0 * nx ** 3 + nx ** 2) + 110 * 2 * nx ** 2 + 105 * npl * 2 * nx ** 2
         + 945 * (2 * 21 + 48 / npl))
    V = 100 * (q + 1) * (q + 2) * nx * (1 + npl) + 100 * 2 * (q + 1) * nx * (1 + npl) + (110 + 105 * npl) * 6 * nx
    assert abs(flops - 2 * (F + V)) <= 1e-9 * flops
    assert mean["nst"] == 100
