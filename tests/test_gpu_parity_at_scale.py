"""Parity gate (iii) of SURVEY.md section 8c at full size: for every DecodeMethod, >= 1000 groups of 32 frames spanning
three Eb/N0 points (below / at / above the waterfall) are decoded on the GPU through the C-ABI and by the REFERENCE
ITSELF (oracle/_ref: the reference's own translation units compiled unmodified, travelling to the GPU box as a
prebuilt .so).  decodedBits and the returned BF-iteration counts must be identical for every group.

The LLRs come from the engine's own fused producer (golden codeword + Philox AWGN), so the test also exercises
generate -> decode on device buffers.  Where oracle/_ref is not available (no AVX-512 host, or not built) the
plain-C oracle checks a bounded subset instead, which keeps the test meaningful but slower per group.
"""
import numpy as np
import pytest

import llrgen

pytestmark = pytest.mark.gpu
N, K = 17664, 14592
EBN0 = (3.0, 3.6, 4.2)
GROUPS_PER_POINT = 336  # 3 x 336 = 1008 groups = 32,256 frames per method


def _reference(variant):
    import pyoracle
    try:
        if pyoracle.ref_available(variant):
            return pyoracle.Ref(variant)
    except OSError:
        pass
    return None


def _tx_group():
    cw = llrgen.golden_codeword()
    return np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]).astype(np.int8)


@pytest.mark.parametrize("method,lut,variant,groups_per_point", [
    (0, -1, "faid3", GROUPS_PER_POINT),
    (1, -1, "faid3", GROUPS_PER_POINT),
    (2, 0, "faid3", GROUPS_PER_POINT),
    (3, -1, "faid3", GROUPS_PER_POINT),
    (4, -1, "faid3", GROUPS_PER_POINT),
    (5, 3, "faid3", GROUPS_PER_POINT),
    (2, 1, "faid32", 64),
    (2, 2, "faid2", 64),
])
def test_thousand_groups_bit_exact_vs_reference(oracle, engine_lib, method, lut, variant, groups_per_point):
    import torch

    import ldpc_b200
    cfg = ldpc_b200.default_config(method, lut)
    ref = _reference(variant)
    if ref is None:
        groups_per_point = 4  # plain-C oracle: ~0.2-0.5 s per group
    G = groups_per_point
    tx = torch.from_numpy(_tx_group()).cuda().repeat(G, 1)
    bad = []
    with ldpc_b200.Decoder(cfg) as dec:
        for i, eb in enumerate(EBN0):
            fix = dec.generate(tx, eb, 7000 + method, i * G * 32, G)
            out, info = dec.decode(fix, want_info=True)
            fix_h = fix.cpu().numpy()
            out_h = out.cpu().numpy()
            if ref is not None:
                exp, bfs = ref.decode(ldpc_b200.default_config(method, lut), fix_h)
            else:
                exp, infos = oracle.decode(oracle.default_config(method, lut), fix_h)
                bfs = [x.bf_iters for x in infos]
            diff_groups = np.nonzero((out_h != exp).any(axis=1))[0]
            if diff_groups.size:
                bad.append((eb, "bits", diff_groups[:5].tolist(), int((out_h != exp).sum())))
            if method in (3, 4):  # the reference returns BFiter only from Decode_OMSBF / Decode_OMS_DTBF
                bf_bad = np.nonzero(np.asarray(bfs) != info["bf_iters"])[0]
                if bf_bad.size:
                    bad.append((eb, "bf_iters", bf_bad[:5].tolist()))
            # sanity of the workload itself: the three points straddle the waterfall
            fer = float((out_h.reshape(G * 32, N)[:, :K] != llrgen.golden_codeword()[None, :K]).any(axis=1).mean())
            if i == 0:
                fer_low = fer
        assert fer <= fer_low, "FER must not rise from 3.0 dB to 4.2 dB"
    assert not bad, bad
