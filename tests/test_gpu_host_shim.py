"""The C++ host mirror (CLDPC-shaped shim + sweep driver) against the Python/C-ABI path."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import llrgen

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "mod-interleaveavx_multithreads-faid_b200" / "host"
N, M, K = 17664, 3072, 14592


@pytest.fixture(scope="module")
def host_bins(engine_lib):
    subprocess.run(["make", "-C", str(HOST)], check=True, capture_output=True)
    return HOST


@pytest.mark.parametrize("method", [2, 4])
def test_csimulate_style_loop_matches_c_abi(host_bins, tmp_path, method):
    import ldpc_b200
    prof = tmp_path / "Profile.txt"
    prof.write_text((ROOT / "tests/golden/Profile_shipped.txt").read_text().replace("DecodeMethod: 2", f"DecodeMethod: {method}"))
    cw = llrgen.golden_codeword()
    (tmp_path / "cw.txt").write_text("".join(str(int(b)) for b in cw))
    eb, seed, blocks = 3.6, 77, 6
    r = subprocess.run([str(HOST / "run_like_csimulate"), str(prof), str(tmp_path / "cw.txt"), str(eb), str(seed), str(blocks)],
                       capture_output=True, text=True, check=True)
    got = [int(x) for x in r.stdout.split()]
    cfg = ldpc_b200.read_profile(prof)
    with ldpc_b200.Decoder(cfg) as dec:
        c = dec.simulate(eb, seed, 0, blocks, codeword=cw)
    assert got == [int(c[0]), int(c[1]), int(c[2]), int(c[3])]
    assert got[0] == 32 * blocks and 0 < got[1] < got[0]


def test_sweep_driver_runs_and_reports(host_bins, tmp_path):
    prof = tmp_path / "Profile.txt"
    prof.write_text((ROOT / "tests/golden/Profile_shipped.txt").read_text().replace("StartSNR: 3", "StartSNR: 3.4").replace("EndSNR: 5", "EndSNR: 3.65"))
    r = subprocess.run([str(HOST / "ldpc_sim"), str(prof), "--max-frames", "3200", "--groups-per-round", "50"],
                       capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith("Eb/N0") and len(lines) == 4  # 3.4, 3.5, 3.6
    fers = [float(l.split("\t")[4]) for l in lines[1:]]
    assert all(0 <= f <= 1 for f in fers) and fers[0] >= fers[-1]
