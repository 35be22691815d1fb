"""Parity gate (v) of SURVEY.md section 8c: with its NATIVE noise source (Philox4x32-10 + Box-Muller on the SFU) the
engine's FER / BER must fall inside the 95 % confidence interval of the reference chain run with the reference's own
3-LCG + Box-Muller channel (CChannel.cpp:71-97) on the box's CPU.

Reference side: oracle/_ref (CModulate + CChannel + CLDPC compiled unmodified) driven like CSimulate::Run, golden
codeword, thread-0 seed 101 (CSimulate.cpp:11); falls back to the plain-C oracle chain when the reference build is not
available.  GPU side: ldpc_b200_simulate on 50x more frames, so its own sampling error is negligible next to the CI.
"""
import numpy as np
import pytest

import llrgen

pytestmark = pytest.mark.gpu
N, K = 17664, 14592


def _reference_fer_ber(oracle, method, mod, il, eb, blocks):
    import pyoracle
    cfg = oracle.default_config(method, -1)
    cfg.mod_type, cfg.interleave_mod_type = mod, il
    cw = llrgen.golden_codeword()
    sigma = oracle.sigma(eb, mod)
    ef = ebits = 0
    if pyoracle.ref_available("faid3"):
        ref = pyoracle.Ref("faid3")
        sim = pyoracle.RefSim(ref, cfg, seed=101)
        sim.set_codeword(cw)
        for _ in range(blocks):
            sim.noise_block(sigma, cfg.scale)
            _, st, _ = sim.decode_and_count(method)
            ef += int(st[0])
            ebits += int(st[1])
    else:
        tx = np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]).astype(np.int8)
        modseq = oracle.modulate(tx, mod, il)
        state = np.array([101, 101, 101], dtype=np.uint64)
        info = np.tile(cw[:K], 32).astype(np.int8)
        for _ in range(blocks):
            sym, state = oracle.awgn(modseq, np.float32(sigma / np.sqrt(2)), state)
            _, deint = oracle.demodulate(sym, mod, il)
            dec, _ = oracle.decode(cfg, oracle.quantize(deint, cfg.scale)[None, :])
            st = oracle.calc_errors(info, dec[0])
            ef += int(st[0])
            ebits += int(st[1])
    n = 32 * blocks
    return ef / n, ebits / (n * K), n


@pytest.mark.parametrize("method,mod,il,eb,blocks", [
    (0, 2, 1, 3.6, 60),    # NMS, waterfall (FER ~ 0.7)
    (2, 2, 1, 3.5, 60),    # FAID3 + DTBF
    (4, 4, 4, 8.0, 40),    # OMS + DTBF, 16-QAM with bit interleaving
])
def test_native_rng_fer_inside_reference_confidence_interval(oracle, engine_lib, method, mod, il, eb, blocks):
    import ldpc_b200
    p_ref, ber_ref, n_ref = _reference_fer_ber(oracle, method, mod, il, eb, blocks)
    cfg = ldpc_b200.default_config(method, -1)
    cfg.mod_type, cfg.interleave_mod_type = mod, il
    G = 50 * blocks
    with ldpc_b200.Decoder(cfg) as dec:
        c = dec.simulate(eb, 12345, 0, G, codeword=llrgen.golden_codeword())
    p_gpu = float(c[1]) / float(c[0])
    ber_gpu = float(c[2]) / (float(c[0]) * K)
    # 95 % CI of the reference estimate around the (much better known) GPU value; the frames of one reference block are
    # independent Bernoulli trials
    half = 1.96 * np.sqrt(max(p_gpu * (1.0 - p_gpu), 1e-6) / n_ref) + 1.0 / n_ref
    assert abs(p_ref - p_gpu) <= half, f"FER reference {p_ref:.4f} ({n_ref} frames) vs engine {p_gpu:.4f} ({int(c[0])} frames), CI half width {half:.4f}"
    # BER: errors come in bursts of failed frames; bound the ratio of per-failed-frame error counts instead
    if p_ref > 0 and p_gpu > 0:
        per_frame_ref, per_frame_gpu = ber_ref / p_ref, ber_gpu / p_gpu
        assert 0.6 < per_frame_ref / per_frame_gpu < 1.6, (ber_ref, ber_gpu)
