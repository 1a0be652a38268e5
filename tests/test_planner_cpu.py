"""Host logic of the convolution planner (dynmm_b200/csrc/conv_plan.cuh) without a GPU: tools/plan_report.cu compiles
the planner with tensor-map encoding stubbed out (-DDYNMM_PLAN_DRYRUN) and plans every convolution of the
ESANet-R34-NBt1D forward.  Invariants: every shape plans, the tiles cover all output pixels and channels, the
pipeline has at least two stages inside the 227 KiB budget (113 KiB when two CTAs share an SM), and the implied
tensor-pipe time is never below the ideal."""
import csv
import io
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def plan_report(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("plan") / "plan_report")
    subprocess.run([NVCC, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-DDYNMM_PLAN_DRYRUN", "-I", os.path.join(ROOT, "dynmm_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tools", "plan_report.cu")], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("batch,active", [(8, 8), (8, 3), (8, 0), (1, 1), (32, 32), (5, 2)])
def test_every_layer_plans_and_covers_its_output(plan_report, batch, active):
    r = subprocess.run([plan_report, str(batch), str(active), "--csv"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    rows = list(csv.DictReader(io.StringIO(r.stdout)))
    assert len(rows) >= (30 if active == 0 else 47)
    for row in rows:
        n, h, w, cout = (int(row[k]) for k in ("n", "h", "w", "cout"))
        kh, kw = int(row["kh"]), int(row["kw"])
        name = row["layer"]
        sh = 2 if ("3x1 s2" in name or "1x1 s2" in name) else 1
        sw = 2 if ("1x3 s2" in name or "1x1 s2" in name) else 1
        h_out, w_out = (h + 2 * (kh // 2) - kh) // sh + 1, (w + 2 * (kw // 2) - kw) // sw + 1
        b1, b2, bn, tile_n, c_tiles = (int(row[k]) for k in ("b1", "b2", "bn", "tile_n", "c_tiles"))
        assert b1 * b2 * bn <= 128, name
        assert tile_n % 16 == 0 and 16 <= tile_n <= 256 and tile_n * c_tiles >= cout, name
        tiles = int(row["tiles"])
        # the boxes tile the (d1, d2, sample) lattice: at least the pixel count, whatever the orientation
        assert tiles * b1 * b2 * bn >= n * h_out * w_out * c_tiles, name
        assert tiles % c_tiles == 0, name
        two = int(row["two_per_sm"])
        assert int(row["stages"]) >= 2, name
        assert float(row["smem_kb"]) <= (113.0 if two else 227.0), name
        assert int(row["ctas"]) == (296 if two else 148)
        assert int(row["rounds"]) == -(-tiles // int(row["ctas"])), name
        assert float(row["efficiency"]) <= 1.0 + 1e-9, name
        if two:                 # only the large resident-weight C = 64 layers share an SM
            assert tile_n == 64 and int(row["resident"]) == 1 and row["mode"] == "halo", name


def test_report_totals_are_consistent(plan_report):
    out = subprocess.run([plan_report, "8", "8"], capture_output=True, text=True, check=True).stdout
    total = out.strip().splitlines()[-1]
    assert total.startswith("TOTAL 177 launches") and "tiling efficiency" in total
    eff = float(total.rsplit(" ", 1)[1])
    assert 0.3 < eff <= 1.0
