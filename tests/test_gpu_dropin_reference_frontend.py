"""Drop-in evidence with the reference's OWN objects.

oracle/_ref/dropin_csimulate links the reference's CTool / CChannel / CModulate / CLDPC translation units (compiled
unmodified from /root/reference in the dev container; the binary travels to the GPU box) and runs the body of
CSimulate::Run (CSimulate.cpp:103-169) with the reference's 3-LCG channel, seed 101, three times:

  ref    the reference's CLDPC::Decode*()
  cabi   the one-call patch: ldpc_b200_decode(gpu, ldpc->fixInput, ldpc->decodedBits, 1, &BFiter, &its, nullptr)
  shim   `CLDPC` replaced by the class CLDPC_B200 (host/CLDPC_b200.h compiled against the reference's headers)

Per 32-frame block ErrorFrame / ErrorBits / LT3ErrBitFrame / BFiter and a hash of decodedBits must be identical.
"""
import subprocess
from pathlib import Path

import pytest

import llrgen

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "oracle" / "_ref" / "dropin_csimulate"

# (method, modType, InterleaveModType, Factor_1, Factor_2, scale, Eb/N0, MaxIteration)
CASES = [
    (0, 2, 1, 26, 26, 13, 3.6, 6),     # BASELINE config 0: NMS 26/26
    (1, 2, 1, 1, 6, 13, 3.7, 6),
    (2, 2, 1, 1, 6, 13, 3.6, 6),       # config 1: FAID3 + DTBF
    (3, 2, 2, 1, 6, 13, 3.6, 6),
    (4, 4, 4, 1, 6, 13, 7.3, 6),       # config 3: OMS + DTBF, 16-QAM
    (4, 6, 1, 1, 6, 13, 12.5, 6),      #           ... and 64-QAM
    (4, 6, 6, 1, 6, 13, 12.4, 6),      #           ... with the bit interleaver
    (5, 2, 1, 1, 6, 12.5, 3.6, 6),     # config 2: hybrid + 2B1C, scale 12.5
    (5, 2, 1, 1, 6, 12.5, 3.5, 12),
]


def _profile(method, mod, il, f1, f2, scale, max_iter):
    return (f"Simulation parameter\nStartSNR: 3\nSNRPass: 0.1\nEndSNR: 5\nDecodeMethod: {method}\nMaxIteration: {max_iter}\n"
            f"Modulation Parameter:\nmodType: {mod}\nInterleaveModType: {il}\nNMS  Factor:\nFactor_1: {f1}\nFactor_2: {f2}\n"
            f"noFrames: 32\nscale: {scale}\nMatrix Factor\nFileName: 50GPON-CP12\nZ: 256\n")


@pytest.mark.parametrize("method,mod,il,f1,f2,scale,ebn0,max_iter", CASES)
@pytest.mark.parametrize("codeword", ["golden", "zero"])
def test_reference_front_end_with_gpu_decoder(engine_lib, tmp_path, method, mod, il, f1, f2, scale, ebn0, max_iter, codeword):
    if not BIN.exists():
        pytest.skip("oracle/_ref/dropin_csimulate not built (needs /root/reference at build time: `make -C oracle ref dropin`)")
    if codeword == "zero" and (method, mod) not in ((0, 2), (5, 2)):
        pytest.skip("the shipped all-zero CodeWord_sym is exercised on two configurations")
    (tmp_path / "Profile.txt").write_text(_profile(method, mod, il, f1, f2, scale, max_iter))
    cw_arg = "zero"
    if codeword == "golden":
        (tmp_path / "cw.txt").write_text("".join(str(int(b)) for b in llrgen.golden_codeword()))
        cw_arg = "cw.txt"
    blocks = 8
    r = subprocess.run([str(BIN), str(method), str(ebn0), str(blocks), cw_arg], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = {}
    for line in r.stdout.split("\n"):
        f = line.split()
        if len(f) == 7:
            rows.setdefault(f[0], []).append(tuple(f[2:]))
    assert set(rows) == {"ref", "cabi", "shim"} and all(len(v) == blocks for v in rows.values())
    assert rows["cabi"] == rows["ref"] and rows["shim"] == rows["ref"]
    # the operating points are chosen so that the comparison is not vacuous: some frames fail, not all
    ef = sum(int(x[0]) for x in rows["ref"])
    assert 0 < ef < 32 * blocks, ef
    if method in (3, 4):
        assert any(int(x[3]) > 0 for x in rows["ref"]), "BF iterations should be exercised"
