"""GPU parity of the frame-generation / scoring kernels against the CPU oracle."""
import numpy as np
import pytest

import llrgen

pytestmark = pytest.mark.gpu
N, M, K = 17664, 3072, 14592


def _cfg(method=0, mod=2, il=1):
    import ldpc_b200
    cfg = ldpc_b200.default_config(method, -1)
    cfg.mod_type, cfg.interleave_mod_type = mod, il
    return cfg


def test_quantize_bit_exact(oracle, engine_lib):
    import ldpc_b200
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1 << 20) * 0.6).astype(np.float32)
    # exact integer boundaries, huge values, NaN / inf (cvttps "integer indefinite" -> -7)
    x[:16] = np.array([0.0, -0.0, 7 / 13, -7 / 13, 1 / 13, -1 / 13, 1e30, -1e30, np.inf, -np.inf, np.nan, 0.5384615, 0.53846157, -0.5384616, 3e9, -3e9], dtype=np.float32)
    with ldpc_b200.Decoder(_cfg()) as dec:
        for scale in (13.0, 12.5, 14.0):
            assert (dec.quantize(x, scale) == oracle.quantize(x, scale)).all()


@pytest.mark.parametrize("mod,il", [(2, 1), (2, 2), (4, 1), (4, 4), (6, 1), (6, 6), (8, 1), (8, 8)])
def test_demap_matches_oracle(oracle, engine_lib, mod, il):
    import ldpc_b200
    rng = np.random.default_rng(mod * 10 + il)
    sym = (rng.standard_normal(2 * 32 * N // mod) * 0.7).astype(np.float32)
    with ldpc_b200.Decoder(_cfg(mod=mod, il=il)) as dec:
        llr, fix = dec.demap(sym[None, :])
    demod, deint = oracle.demodulate(sym, mod, il)
    # the double-precision subtraction is reproduced, so floats are bit-identical, not just within 1e-5
    assert np.array_equal(llr.reshape(-1).view(np.uint32), deint.view(np.uint32))
    assert (fix.reshape(-1) == oracle.quantize(deint, 13.0)).all()


@pytest.mark.parametrize("mod,il", [(2, 1), (4, 4), (6, 1), (6, 3), (8, 1), (8, 4)])
def test_generate_noiseless_mapping_and_noise_stats(oracle, engine_lib, mod, il):
    """Mapping/interleaving exactness (symbol means equal the oracle's constellation points) and noise statistics."""
    import ldpc_b200
    rng = np.random.default_rng(7)
    info = rng.integers(0, 2, (32 * K,), dtype=np.int8)
    tx = oracle.encode_group(info)
    ref_sym = oracle.modulate(tx, mod, il)
    cfg = _cfg(mod=mod, il=il)
    with ldpc_b200.Decoder(cfg) as dec:
        fix, sym = dec.generate(tx[None, :], 60.0, 11, 0, 1, want_symbols=True)  # 60 dB: noise ~1e-3
        assert np.abs(sym.reshape(-1) - ref_sym).max() < 1e-2
        # and the demapped, quantised LLRs equal the oracle's on those very symbols
        _, deint = oracle.demodulate(sym.reshape(-1), mod, il)
        assert (fix.reshape(-1) == oracle.quantize(deint, cfg.scale)).all()
        eb = 4.0
        fix2, sym2 = dec.generate(tx[None, :], eb, 11, 0, 1, want_symbols=True)
        noise = sym2.reshape(-1) - ref_sym
        sd = ldpc_b200.ebn0_sigma(cfg, eb) / np.sqrt(2.0)
        n = noise.size
        assert abs(noise.mean()) < 5 * sd / np.sqrt(n)
        assert abs(noise.std() / sd - 1) < 5 / np.sqrt(2 * n)
        # independence of the call partitioning: frames 32..63 of a 2-group call == a call starting at frame 32
        f_a = dec.generate(np.concatenate([tx, tx])[None, :].reshape(2, -1), eb, 11, 0, 2)
        f_b = dec.generate(tx[None, :], eb, 11, 32, 1)
        assert (f_a[1] == f_b[0]).all() and (f_a[0] == fix2[0]).all()


def test_encode_matches_oracle_and_golden(oracle, engine_lib):
    import ldpc_b200
    rng = np.random.default_rng(3)
    info = rng.integers(0, 2, (2, 32 * K), dtype=np.int8)
    info[0, :K] = llrgen.golden_codeword()[:K]
    with ldpc_b200.Decoder(_cfg()) as dec:
        tx = dec.encode(info)
    for g in range(2):
        assert (tx[g] == oracle.encode_group(info[g])).all()
    cw = np.concatenate([tx[0, :K], tx[0, 32 * K: 32 * K + M]])
    assert (cw == llrgen.golden_codeword()).all()


def test_count_errors_matches_oracle(oracle, engine_lib):
    import ldpc_b200
    rng = np.random.default_rng(4)
    info = rng.integers(0, 2, (3, 32 * K), dtype=np.int8)
    dec_bits = np.zeros((3, 32, N), dtype=np.int8)
    dec_bits[:, :, :K] = info.reshape(3, 32, K)
    dec_bits[:, :, K:] = rng.integers(0, 2, (3, 32, M))  # parity mismatches are not counted
    flips = [(0, 0, [5]), (0, 3, [7, 9]), (1, 1, [1, 2, 3]), (2, 31, list(range(100, 400)))]
    for g, f, idx in flips:
        dec_bits[g, f, idx] ^= 1
    with ldpc_b200.Decoder(_cfg()) as dec:
        c = dec.count_errors(info, dec_bits.reshape(3, -1))
    exp = sum(oracle.calc_errors(info[g], dec_bits[g].reshape(-1)) for g in range(3))
    assert c[0] == 96 and (c[1], c[2], c[3]) == tuple(int(x) for x in exp)


@pytest.mark.parametrize("method", [0, 2, 5])
def test_simulate_equals_stepwise_pipeline(oracle, engine_lib, method):
    """The fused on-device round gives exactly the counters of generate -> decode -> count through the API,
    and those agree with the oracle decoding the very same LLRs."""
    import ldpc_b200
    cfg = ldpc_b200.default_config(method, -1)
    cw = llrgen.golden_codeword()
    eb, seed, G = 3.7, 99, 3
    with ldpc_b200.Decoder(cfg) as dec:
        c_sim = dec.simulate(eb, seed, 1000, G, codeword=cw)
        tx = np.tile(np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]), (G, 1)).astype(np.int8)
        fix = dec.generate(tx, eb, seed, 1000, G)
        out, info = dec.decode(fix, want_info=True)
        c_step = dec.count_errors(np.tile(cw[:K], (G, 32)).astype(np.int8), out)
        # random info bits path (device encoder)
        c_rand = dec.simulate(eb, seed, 0, G)
    assert tuple(c_sim[:4]) == tuple(c_step[:4])
    ref, _ = oracle.decode(oracle.default_config(method, -1), fix)
    assert (ref == out).all()
    assert c_sim[4] == G and c_sim[5] == info["its_per_group"].sum()
    assert c_rand[0] == 32 * G and c_rand[1] <= 32 * G


@pytest.mark.parametrize("method,mod,il,eb", [(0, 2, 1, 3.6), (4, 2, 2, 3.3), (4, 4, 4, 8.0), (1, 6, 1, 12.0), (2, 6, 3, 12.5), (0, 8, 1, 17.0)])
def test_fused_producer_equals_separate_kernels(engine_lib, monkeypatch, method, mod, il, eb):
    """the producer fused into the decoder's loader (default) and the separate generate_kernel ->
    LLR buffer -> decode path (LDPC_B200_NO_FUSED_PRODUCER) count exactly the same errors and iterations, for a
    fixed codeword and for random info bits through the device encoder."""
    import ldpc_b200
    cfg = _cfg(method=method, mod=mod, il=il)
    cw = llrgen.golden_codeword()
    G = 5
    res = []
    for no_fuse in (False, True):
        if no_fuse:
            monkeypatch.setenv("LDPC_B200_NO_FUSED_PRODUCER", "1")
        else:
            monkeypatch.delenv("LDPC_B200_NO_FUSED_PRODUCER", raising=False)
        with ldpc_b200.Decoder(cfg) as dec:
            a = dec.simulate(eb, 4242, 77, G, codeword=cw).copy()
            b = dec.simulate(eb, 4242, 96 + 32 * G, G).copy()
        res.append((a, b))
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all()
    assert res[0][0][0] == 32 * G
    if (method, mod, il) == (0, 2, 1):
        assert 0 < res[0][0][1] < 32 * G, "operating point should give a mix of good and bad frames"


@pytest.mark.parametrize("mod", [2, 4, 6, 8])
def test_generate_fast_path_equals_general_path(engine_lib, mod):
    """InterleaveModType == 1: the word-wise producer kernel (generate_i1_kernel, also what the fused decoder loader
    runs) and the general per-bit kernel (taken when the noisy symbols are requested) emit identical LLRs."""
    import ldpc_b200
    rng = np.random.default_rng(mod)
    cfg = _cfg(mod=mod, il=1)
    G = 3
    tx = rng.integers(0, 2, (G, 32 * N), dtype=np.int8)
    with ldpc_b200.Decoder(cfg) as dec:
        fast = dec.generate(tx, 6.0, 31, 5, G)
        general, _sym = dec.generate(tx, 6.0, 31, 5, G, want_symbols=True)
    assert (fast == general).all()
    assert np.abs(fast).max() <= 7 and len(np.unique(fast)) > 8


@pytest.mark.parametrize("bits", [1, 2, 3, 5, 6])
def test_quantiser_variants_bit_exact(oracle, engine_lib, bits):
    """ldpc_b200_quantize_bits and the producer's configurable quantiser (config.quant_bits) against the oracle."""
    import ldpc_b200
    rng = np.random.default_rng(bits)
    x = (rng.standard_normal(1 << 18) * 1.5).astype(np.float32)
    x[:16] = np.array([0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 1e30, -1e30, np.inf, -np.inf, np.nan, 3e9, -3e9, 0.49999997], dtype=np.float32)
    cfg = _cfg()
    cfg.quant_bits = bits
    with ldpc_b200.Decoder(cfg) as dec:
        for scale in (1.0, 13.0, 5.8):
            assert (dec.quantize(x, scale, bits=bits) == oracle.quantize(x, scale, bits)).all()
        sym = (rng.standard_normal(2 * 32 * N // 2) * 0.7).astype(np.float32)
        llr, fix = dec.demap(sym[None, :])
        assert (fix.reshape(-1) == oracle.quantize(llr.reshape(-1), cfg.scale, bits)).all()


@pytest.mark.parametrize("reuse", [1, 4, 50])
def test_simulate_codeword_reuse_is_chunk_independent(engine_lib, reuse):
    """Random-info rounds encode once per `codeword_reuse` consecutive groups (CSimulate.cpp:103-117: one Encode() per
    50 noise blocks).  Which codeword and which noise a frame gets depends only on its global index, so splitting a
    round into calls / chunks at arbitrary group boundaries cannot change the counters."""
    import ldpc_b200
    cfg = _cfg(method=1)
    cfg.codeword_reuse = reuse
    cfg.chunk_groups = 4
    with ldpc_b200.Decoder(cfg) as dec:
        whole = dec.simulate(3.5, 9, 32 * 3, 11).copy()
        parts = np.zeros_like(whole)
        for g0, n in ((0, 2), (2, 5), (7, 4)):
            dec.simulate(3.5, 9, 32 * (3 + g0), n, counters=parts)
    assert (whole == parts).all()
    assert whole[0] == 32 * 11 and 0 < whole[1] < 32 * 11


def test_encoder_both_launch_shapes_agree(oracle, engine_lib):
    """encode_group_kernel runs one CTA per group for many groups and 12 CTAs per group (one per parity block row) for
    few groups; both must give the oracle's codewords."""
    import ldpc_b200
    rng = np.random.default_rng(12)
    info = rng.integers(0, 2, (66, 32 * K), dtype=np.int8)
    with ldpc_b200.Decoder(_cfg()) as dec:
        many = dec.encode(info)            # 66 groups: throughput shape
        few = dec.encode(info[[0, 65]])    # 2 groups: latency shape
    assert (many[[0, 65]] == few).all()
    assert (few[0] == oracle.encode_group(info[0])).all() and (few[1] == oracle.encode_group(info[65])).all()


def _bpsk_cfg(method=1):
    cfg = _cfg(method=method, mod=1, il=1)
    return cfg


def test_bpsk_generate_matches_oracle(oracle, engine_lib):
    """modType 1 (CSimulate.cpp:121-124, CModulate.cpp:363-370): x = 2b - 1 on the two-region buffer, real AWGN of the
    full sigma, the quantiser straight on the received amplitude."""
    import ldpc_b200
    rng = np.random.default_rng(17)
    tx = np.stack([oracle.encode_group(rng.integers(0, 2, 32 * K, dtype=np.int8)) for _ in range(2)])
    x = np.stack([oracle.bpsk_modulate(t) for t in tx])
    cfg = _bpsk_cfg()
    with ldpc_b200.Decoder(cfg) as dec:
        fix, sym = dec.generate(tx, 80.0, 3, 0, 2, want_symbols=True)   # 80 dB: noise ~1e-4
        assert sym.shape == (2, 32 * N) and np.abs(sym - x).max() < 1e-2
        assert (fix == np.where(x > 0, 7, -7)).all()
        eb = 4.0
        fix2, sym2 = dec.generate(tx, eb, 3, 0, 2, want_symbols=True)
        # the quantised LLRs are the oracle's quantiser on those very amplitudes, bit for bit
        assert (fix2.reshape(-1) == oracle.quantize(sym2.reshape(-1), cfg.scale)).all()
        noise = (sym2 - x).reshape(-1)
        sd = oracle.sigma(eb, 1)
        assert abs(ldpc_b200.ebn0_sigma(cfg, eb) - sd) < 1e-7
        n = noise.size
        assert abs(noise.mean()) < 5 * sd / np.sqrt(n)
        assert abs(noise.std() / sd - 1) < 5 / np.sqrt(2 * n)
        assert abs(np.corrcoef(noise[0::2], noise[1::2])[0, 1]) < 5 / np.sqrt(n / 2)  # the two normals of one Philox call
        # same frames whatever the call partitioning; device pointers too
        f_b = dec.generate(tx[1:], eb, 3, 32, 1)
        assert (f_b[0] == fix2[1]).all()
        import torch
        f_d = dec.generate(torch.from_numpy(tx).cuda(), eb, 3, 0, 2)
        assert (f_d.cpu().numpy() == fix2).all()
        # other quantiser widths through the same kernel
    cfg6 = _bpsk_cfg()
    cfg6.quant_bits = 6
    with ldpc_b200.Decoder(cfg6) as dec:
        fix6, sym6 = dec.generate(tx[:1], 4.0, 3, 0, 1, want_symbols=True)
        assert (fix6.reshape(-1) == oracle.quantize(sym6.reshape(-1), cfg6.scale, 6)).all()
        with pytest.raises(Exception):
            dec.demap(sym6)   # BPSK has no demapper in the reference


@pytest.mark.parametrize("method,reuse,g0,G", [(1, 0, 48, 9), (4, 4, 3, 11), (2, 1, 0, 3), (0, 50, 0, 4)])
def test_bpsk_simulate_equals_stepwise_pipeline(oracle, engine_lib, method, reuse, g0, G):
    """Random-info BPSK rounds: codeword reuse (default 50 = one Encode() per 50 noise blocks, CSimulate.cpp:103-117) must
    pick each group's transmitted bits from the shared codeword group, exactly as the QPSK / QAM producers do.  The counters
    of the on-device round equal gen_msg_seq -> encode -> generate -> decode -> count_errors through the API, and the decoded
    bits equal the oracle's on the same LLRs."""
    import ldpc_b200
    cfg = _bpsk_cfg(method)
    cfg.codeword_reuse = reuse
    cfg.chunk_groups = 4
    eb, seed = 3.9, 77
    r = reuse if reuse else 50
    with ldpc_b200.Decoder(cfg) as dec:
        c_sim = dec.simulate(eb, seed, 32 * g0, G).copy()
        cw_groups = sorted({(g0 + g) // r for g in range(G)})
        info_of = {c: dec.gen_msg_seq(seed, c * r * 32, 1)[0] for c in cw_groups}
        tx_of = {c: dec.encode(info_of[c][None, :])[0] for c in cw_groups}
        info = np.stack([info_of[(g0 + g) // r] for g in range(G)])
        tx = np.stack([tx_of[(g0 + g) // r] for g in range(G)])
        fix = dec.generate(tx, eb, seed, 32 * g0, G)
        out, inf = dec.decode(fix, want_info=True)
        c_step = dec.count_errors(info, out)
    assert tuple(c_sim[:4]) == tuple(c_step[:4]), (c_sim[:6], c_step[:6])
    assert c_sim[0] == 32 * G and c_sim[1] < 32 * G, "garbage transmitted bits would fail (nearly) every frame"
    assert c_sim[5] == inf["its_per_group"].sum()
    for c in cw_groups:
        assert (tx_of[c] == oracle.encode_group(info_of[c])).all()
    ref, _ = oracle.decode(oracle.default_config(method, -1), fix)
    assert (ref == out).all()
    if reuse != 1 and len(cw_groups) > 1:
        assert (tx_of[cw_groups[0]] != tx_of[cw_groups[1]]).any()


def test_misaligned_device_pointers_are_rejected(engine_lib):
    """Entry points that use vector accesses on caller-supplied DEVICE pointers return EINVAL for a misaligned view instead
    of faulting; simulate() wants whole groups."""
    import ctypes as C
    import torch
    import ldpc_b200
    with ldpc_b200.Decoder(_cfg()) as dec:
        lib = dec.lib
        buf = torch.zeros(32 * N + 64, dtype=torch.int8, device="cuda")
        info = torch.zeros(32 * K + 64, dtype=torch.int8, device="cuda")
        cnt = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
        rc = lib.ldpc_b200_count_errors(dec.h, C.c_void_p(info.data_ptr() + 4), C.c_void_p(buf.data_ptr()), 1, C.c_void_p(cnt.ctypes.data))
        assert rc == -1 and b"aligned" in lib.ldpc_b200_last_error()
        rc = lib.ldpc_b200_count_errors(dec.h, C.c_void_p(info.data_ptr()), C.c_void_p(buf.data_ptr() + 8), 1, C.c_void_p(cnt.ctypes.data))
        assert rc == -1
        rc = lib.ldpc_b200_gen_msg_seq(dec.h, 1, 0, 1, C.c_void_p(info.data_ptr() + 1))
        assert rc == -1
        assert lib.ldpc_b200_count_errors(dec.h, C.c_void_p(info.data_ptr()), C.c_void_p(buf.data_ptr()), 1, C.c_void_p(cnt.ctypes.data)) == 0
        rc = lib.ldpc_b200_simulate(dec.h, None, 3.5, 1, 17, 1, C.c_void_p(cnt.ctypes.data))
        assert rc == -1 and b"multiple of 32" in lib.ldpc_b200_last_error()
        # the context is still healthy
        assert dec.simulate(3.5, 1, 32, 1)[0] == 32
