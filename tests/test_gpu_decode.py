"""GPU parity: the CUDA decoders (through the C-ABI) against the CPU oracle on the same seeded LLRs."""
import numpy as np
import pytest

import llrgen

pytestmark = pytest.mark.gpu

N, K = 17664, 14592
CASES = [
    # (method, lut, ebn0 points)
    (0, -1, (3.0, 3.6, 4.4)),
    (1, -1, (3.0, 3.6, 4.4)),
    (2, 0, (3.0, 3.6, 4.4)),
    (2, 1, (3.3, 3.8)),
    (2, 2, (3.3, 3.8)),
    (3, -1, (3.0, 3.6, 4.4)),
    (4, -1, (3.0, 3.6, 4.4)),
    (5, 3, (3.0, 3.6, 4.4)),
]


@pytest.mark.parametrize("method,lut,ebs", CASES)
def test_decode_matches_oracle(oracle, engine_lib, method, lut, ebs):
    import ldpc_b200
    cfg = ldpc_b200.default_config(method, lut)
    ocfg = oracle.default_config(method, lut)
    assert bytes(cfg)[: -8 * 4] == bytes(ocfg)[: -8 * 4]  # everything but the execution fields
    fix = np.concatenate([llrgen.qpsk_llr_groups(2, eb, scale=cfg.scale, seed=100 * method + i)[0] for i, eb in enumerate(ebs)])
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(fix, want_info=True)
    ref, infos = oracle.decode(ocfg, fix)
    nd = int((out != ref).sum())
    assert nd == 0, f"{nd} differing bits"
    assert [i.bf_iters for i in infos] == list(info["bf_iters"])
    assert [i.iters_executed for i in infos] == list(info["its_per_group"])
    assert (np.array([list(i.conv_iter) for i in infos]) == info["conv_iter"]).all()
