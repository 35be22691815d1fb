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


def test_sweep_driver_writes_reference_files_and_error_dumps(host_bins, tmp_path):
    """--out-dir: Result.txt / Temp.txt / demod.txt / iterCount.txt in the reference's formats (main.cpp:194-227,
    CSimulate.cpp:171-179) and, with the collect flag, the error-frame dumps of CLDPC.cpp:4877-4991 obtained by
    replaying the failing rounds through the step-wise C-ABI (the replay must find exactly the counted error frames,
    otherwise the driver exits non-zero)."""
    prof = tmp_path / "Profile.txt"
    prof.write_text((ROOT / "tests/golden/Profile_shipped.txt").read_text()
                    .replace("DecodeMethod: 2", "DecodeMethod: 4").replace("StartSNR: 3", "StartSNR: 3.7").replace("EndSNR: 5", "EndSNR: 3.85"))
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([str(HOST / "ldpc_sim"), str(prof), "--max-frames", "640", "--groups-per-round", "10", "--out-dir", str(out),
                        "--collect"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in (out / "Result.txt").read_text().strip().splitlines()]
    assert len(rows) == 2 and len(rows[0]) == 8  # Eb/N0 3.7 and 3.8; the reference's eight columns
    stdout_rows = [l.split("\t") for l in r.stdout.strip().splitlines()[1:]]
    for a, b in zip(rows, stdout_rows):
        assert [int(a[1]), int(a[2]), int(a[3])] == [int(b[1]), int(b[2]), int(b[3])]
    n_err = sum(int(a[2]) for a in rows)
    assert (out / "errorindex.txt").read_text().count("ErrorFrame:") == n_err
    assert (out / "errordecode.txt").read_text().count("Decodedbits=[") == n_err
    assert (out / "errorfloat.txt").read_text().count("ErrorChar=[") == n_err
    assert "lastPhilox" in (out / "Temp.txt").read_text()
    assert (out / "iterCount.txt").read_text().count("Eb/N0:") == 2 and (out / "demod.txt").exists()
