"""CPU: the plain-C oracle against the committed golden vectors (produced by the reference's own code) and the
reference's only in-tree known-answer vector, the "50G PON NS NP" codeword (Codeword.h:6-460)."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import golden_util as gu
import llrgen

ROOT = Path(__file__).resolve().parent.parent
N, M, K = 17664, 3072, 14592


def test_golden_codeword_pins_H_and_encoder(oracle):
    cw = llrgen.golden_codeword()
    assert int(cw.sum()) == 8759
    assert oracle.syndrome_weight(cw) == 0
    bad = cw.copy()
    bad[1234] ^= 1
    assert oracle.syndrome_weight(bad) == 6  # block column 4 has weight 6
    assert (oracle.encode_frame(cw[:K]) == cw).all()
    rng = np.random.default_rng(1)
    for _ in range(3):
        c = oracle.encode_frame(rng.integers(0, 2, K, dtype=np.int8))
        assert oracle.syndrome_weight(c) == 0


@pytest.mark.parametrize("name", gu.CASE_NAMES)
def test_oracle_decoders_match_reference_dumps(oracle, name):
    c = gu.case(name)
    cfg = gu.apply(oracle.default_config(c["method"], c["lut"]), c)
    dec, infos = oracle.decode(cfg, c["fix"])
    assert int((dec != c["dec"]).sum()) == 0
    for g, info in enumerate(infos):
        if c["bf"][g] >= 0:  # the reference returns BFiter only for methods 3 and 4
            assert info.bf_iters == c["bf"][g]
        if c["its"][g] >= 0:
            assert info.iters_executed == c["its"][g]
            if c["method"] != 0:
                log = np.array([list(r) for r in info.errsum_log])[: info.errsum_n]
                assert (log == c["errsum"][g][: info.errsum_n]).all()


def test_oracle_chain_matches_reference_hashes(oracle):
    """LCG + Box-Muller channel, max-log demapper (double-precision subtraction), de-interleaver, 4-bit quantiser."""
    hashes = json.loads((ROOT / "tests" / "golden" / "chain_hashes.json").read_text())
    cw = llrgen.golden_codeword()
    tx = np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]).astype(np.int8)
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for key, ref in hashes.items():
        mod, il = int(key[3]), int(key.split("il")[1])
        mod_sym = oracle.modulate(tx, mod, il)
        sd = np.float32(oracle.sigma(ref["ebn0"], mod) / np.sqrt(2))
        sym, _ = oracle.awgn(mod_sym, sd, [ref["seed"]] * 3)
        demod, deint = oracle.demodulate(sym, mod, il)
        fix = oracle.quantize(deint, ref["scale"])
        assert np.bincount(fix.astype(int) + 7, minlength=15).tolist() == ref["fix_hist"], key
        assert (h(sym), h(demod), h(deint), h(fix)) == (ref["symbols"], ref["demod"], ref["deint"], ref["fix"]), key


def test_quantiser_edge_cases(oracle):
    x = np.array([0.0, 7 / 13, 0.53846157, -0.5384616, 1e30, -1e30, np.inf, -np.inf, np.nan, 3e9], dtype=np.float32)
    q = oracle.quantize(x, 13.0)
    assert q.tolist() == [0, 7, 7, -7, -7, -7, -7, -7, -7, -7]  # incl. cvttps "integer indefinite" -> -7 (checked against oracle/_ref)


def test_calc_errors_counts_info_bits_only(oracle):
    info = np.zeros(32 * K, dtype=np.int8)
    dec = np.zeros((32, N), dtype=np.int8)
    dec[0, 5] = 1
    dec[1, [1, 2, 3]] = 1
    dec[2, K + 7] = 1  # parity-only error: located by the reference but never counted (CLDPC.cpp:4858-4868)
    st = oracle.calc_errors(info, dec.reshape(-1))
    assert st.tolist() == [2, 4, 1]
