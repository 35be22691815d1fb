"""GPU parity over RANDOM configurations: the engine must implement the same parametrised algorithm as the oracle, not
only its shipped operating points.  Every draw changes the constants the reference fixes at compile time or reads from
Profile.txt -- factors, LUTs (non-monotone and different per column-weight class, which forces the general FAID kernel),
error-floor / offset thresholds, bit-flipping constants, iteration limits, tail puncturing -- and decodes the same seeded
LLR groups with both.  The oracle's own pinning to the reference is in tests/test_oracle_vs_reference.py."""
import numpy as np
import pytest

import llrgen

pytestmark = pytest.mark.gpu


def _randomise(cfgs, method, rng):
    """apply one random draw to all config objects in `cfgs` (engine + oracle structs have the same fields)"""
    draw = {}
    draw["max_iteration"] = int(rng.integers(1, 13))
    if method == 0:
        draw["factor_1"], draw["factor_2"] = int(rng.integers(8, 40)), int(rng.integers(8, 40))
    elif method in (1, 3, 4):
        f1 = int(rng.integers(0, 4))
        draw["factor_1"], draw["factor_2"] = f1, int(rng.integers(f1 + 1, 8))
        draw["oms_floor_err_count"] = int(rng.integers(0, 256))
        draw["oms_floor_iter_thresh"] = int(rng.integers(-1, 8))
        if rng.random() < 0.3:
            draw["oms_mode"], draw["oms_offset"] = 0, int(rng.integers(0, 3))
    else:
        draw["ef_elimination"] = int(rng.integers(0, 3))
        draw["ef_floor_err_count"] = int(rng.integers(0, 128))
        draw["ef_floor_iter_thresh"] = int(rng.integers(-1, 8))
    draw["puncture_tail"] = int(rng.choice([0, 384, 100, 1000]))
    if method >= 2:
        draw["bf_max_iter"] = int(rng.integers(0, 14))
        draw["dtbf_L0"], draw["dtbf_L1"] = int(rng.integers(0, 6)), int(rng.integers(0, 4))
        draw["dtbf_delta"], draw["dtbf_alpha"] = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        draw["hard2_threshold"] = int(rng.integers(1, 32))
    luts = None
    if method in (2, 5):
        style = rng.integers(0, 3)
        if style == 0:      # monotone, same for all classes: the fast kernel
            base = np.sort(rng.integers(0, 8, size=(2, 6, 1, 8)), axis=-1).repeat(4, axis=2)
        elif style == 1:    # monotone but different per class: general kernel
            base = np.sort(rng.integers(0, 8, size=(2, 6, 4, 8)), axis=-1)
        else:               # arbitrary
            base = rng.integers(0, 8, size=(2, 6, 4, 8))
        luts = base.astype(np.int8)
    for c in cfgs:
        for k, v in draw.items():
            setattr(c, k, v)
        if luts is not None:
            for it in range(6):
                for w in range(4):
                    for a in range(8):
                        c.v2c_lut[it][w][a] = int(luts[0, it, w, a])
                        c.v2c_lut_ef[it][w][a] = int(luts[1, it, w, a])
    return draw


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_random_configurations_match_oracle(oracle, engine_lib, method):
    import ldpc_b200
    rng = np.random.default_rng(9000 + method)
    fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, seed=500 + 7 * method + i)[0] for i, eb in enumerate((3.2, 3.7, 4.3))]
                         + [llrgen.sparse_error_groups(1, 7, 6, seed=method)])
    for trial in range(8):
        cfg = ldpc_b200.default_config(method, -1)
        ocfg = oracle.default_config(method, -1)
        draw = _randomise((cfg, ocfg), method, rng)
        with ldpc_b200.Decoder(cfg) as dec:
            out, info = dec.decode(fix, want_info=True)
        ref, infos = oracle.decode(ocfg, fix)
        nd = int((out != ref).sum())
        assert nd == 0, f"method {method} trial {trial}: {nd} differing bits with {draw}"
        assert [i.bf_iters for i in infos] == list(info["bf_iters"]), draw
        assert [i.iters_executed for i in infos] == list(info["its_per_group"]), draw
