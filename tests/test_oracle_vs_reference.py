"""The plain-C oracle against the reference's own translation units (oracle/_ref/*.so, compiled unmodified).
Skipped where the reference build is absent; the committed fixtures (test_oracle_golden.py) cover that case."""
import numpy as np
import pytest

import llrgen

pyoracle = pytest.importorskip("pyoracle")
pytestmark = pytest.mark.skipif(not pyoracle.ref_available("faid3"), reason="oracle/_ref not built")
N, M, K = 17664, 3072, 14592


@pytest.fixture(scope="module")
def refs():
    return {v: pyoracle.Ref(v) for v in ("faid3", "faid2", "faid32", "instr", "oms0", "ef2") if pyoracle.ref_available(v)}


@pytest.mark.parametrize("method,lut,variant", [(0, -1, "faid3"), (1, -1, "faid3"), (2, 0, "faid3"), (2, 1, "faid32"),
                                                (2, 2, "faid2"), (3, -1, "faid3"), (4, -1, "faid3"), (5, 3, "faid3")])
def test_decoders_on_fresh_noise(oracle, refs, method, lut, variant):
    cfg = oracle.default_config(method, lut)
    fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, scale=cfg.scale, seed=1000 + 10 * method + i)[0]
                          for i, eb in enumerate((3.2, 3.5, 3.9))])
    dec_r, bf_r = refs[variant].decode(cfg, fix)
    dec_o, infos = oracle.decode(cfg, fix)
    assert int((dec_r != dec_o).sum()) == 0
    if method in (3, 4):
        assert bf_r == [i.bf_iters for i in infos]


def test_all_zero_codeword_and_extreme_inputs(oracle, refs):
    """All-zero codeword (the shipped CodeWord_sym), saturated and all-zero LLR inputs."""
    zero_cw = np.zeros(N, dtype=np.int8)
    fix, _ = llrgen.qpsk_llr_groups(1, 3.4, seed=5, codeword=zero_cw)
    extremes = np.stack([fix[0], np.zeros(32 * N, np.int8), np.full(32 * N, 7, np.int8), np.full(32 * N, -7, np.int8)])
    for method in range(6):
        cfg = oracle.default_config(method)
        dec_r, _ = refs["faid3"].decode(cfg, extremes)
        dec_o, _ = oracle.decode(cfg, extremes)
        assert int((dec_r != dec_o).sum()) == 0, method


def test_iteration_counts_match_instrumented_reference(oracle, refs):
    for method in (1, 2, 4, 5):
        cfg = oracle.default_config(method)
        cfg.max_iteration = 10
        fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, scale=cfg.scale, seed=7 + i)[0] for i, eb in enumerate((3.4, 4.0, 4.6))])
        _, _, its, logs = refs["instr"].decode(cfg, fix, want_iters=True)
        _, infos = oracle.decode(cfg, fix)
        assert its == [i.iters_executed for i in infos]
        for g, info in enumerate(infos):
            log = np.array([list(r) for r in info.errsum_log])[: info.errsum_n]
            assert (log == logs[g][: info.errsum_n]).all()
            # per-frame convergence iteration = first logged iteration whose error_sum is zero for that lane
            full = np.vstack([log, np.zeros((1, 32), np.uint8)]) if info.iters_executed < cfg.max_iteration else log
            for f in range(32):
                z = np.nonzero(full[:, f] == 0)[0]
                assert info.conv_iter[f] == (int(z[0]) if z.size else -1)


@pytest.mark.parametrize("mod,il,eb", [(2, 1, 3.6), (4, 4, 8.0), (6, 1, 12.0), (6, 6, 12.0), (8, 1, 17.0), (8, 8, 17.0)])
def test_full_chain_and_error_counting(oracle, refs, mod, il, eb):
    cfg = oracle.default_config(4)
    cfg.mod_type, cfg.interleave_mod_type = mod, il
    sim = pyoracle.RefSim(refs["faid3"], cfg, seed=107)
    cw = llrgen.golden_codeword()
    inp, outb, modseq = sim.set_codeword(cw)
    assert (oracle.modulate(outb, mod, il) == modseq).all()
    st0 = sim.rng_state()
    sigma = oracle.sigma(eb, mod)
    sym, demod, deint, fix = sim.noise_block(sigma, cfg.scale)
    sym_o, st1 = oracle.awgn(modseq, np.float32(sigma / np.sqrt(2)), st0)
    assert (sym_o == sym).all() and (st1 == sim.rng_state()).all()
    d_o, di_o = oracle.demodulate(sym, mod, il)
    assert (d_o == demod).all() and (di_o == deint).all()
    assert (oracle.quantize(deint, cfg.scale) == fix).all()
    dec, stats, bf = sim.decode_and_count(4)
    dec_o, infos = oracle.decode(cfg, fix)
    assert (dec_o[0] == dec).all() and infos[0].bf_iters == bf
    assert (oracle.calc_errors(inp, dec) == stats).all()


def test_transposes(oracle, refs):
    rng = np.random.default_rng(0)
    import ctypes as C
    def aligned(n):  # the reference's transposes use aligned 256-bit loads/stores (vec_malloc'd buffers, CTool.cpp:578-586)
        raw = np.empty(n + 64, dtype=np.int8)
        off = (-raw.ctypes.data) % 64
        return raw[off:off + n]
    a = aligned(32 * 96); got = aligned(32 * 96); exp = aligned(32 * 96)
    a[:] = rng.integers(-7, 8, 32 * 96, dtype=np.int8)
    refs["faid3"].lib.ref_transpose(a.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), 96)
    oracle.lib.ldpc_oracle_transpose(a.ctypes.data_as(C.c_void_p), exp.ctypes.data_as(C.c_void_p), 96)
    assert (got == exp).all()
    refs["faid3"].lib.ref_itranspose(a.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), 96)
    oracle.lib.ldpc_oracle_itranspose(a.ctypes.data_as(C.c_void_p), exp.ctypes.data_as(C.c_void_p), 96)
    assert (got == exp).all()


@pytest.mark.parametrize("bits", [1, 2, 3, 4, 5, 6])
def test_quantiser_variants(oracle, refs, bits):
    """float2LimitChar_{1..6}bit (CLDPC.cpp:4385-4770): rounding mode, asymmetric clamps, "integer indefinite" inputs."""
    rng = np.random.default_rng(bits)
    x = (rng.standard_normal(1 << 16) * 1.5).astype(np.float32)
    x[:20] = np.array([0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.49999997, -0.49999997, 1e30, -1e30, np.inf, -np.inf, np.nan,
                       3e9, -3e9, 1 / 13, -1 / 13, 31.5 / 13], dtype=np.float32)
    for scale in (1.0, 13.0, 5.8):
        assert (oracle.quantize(x, scale, bits) == refs["faid3"].quantize(x, scale, bits)).all(), (bits, scale)


@pytest.mark.parametrize("method", [1, 3, 4])
def test_simple_oms_mode(oracle, refs, method):
    """OMS_MODE 0 (CDecoder_OMS.cpp:3,383-385; reference built with that one #define changed): cste = min - offset,
    negative for min = 0."""
    if "oms0" not in refs:
        pytest.skip("oracle/_ref/libldpc_ref_oms0.so not built")
    cfg = oracle.default_config(method)
    cfg.oms_mode = 0
    fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, scale=cfg.scale, seed=50 + 3 * method + i)[0] for i, eb in enumerate((3.3, 3.9))])
    dec_r, bf_r = refs["oms0"].decode(cfg, fix)
    dec_o, infos = oracle.decode(cfg, fix)
    assert int((dec_r != dec_o).sum()) == 0
    if method in (3, 4):
        assert bf_r == [i.bf_iters for i in infos]
    # and it is a different decoder from the shipped selective mode
    dec_sel, _ = oracle.decode(oracle.default_config(method), fix)
    assert (dec_sel != dec_o).any()


def test_decode1_generic_init_is_no_tail_puncture(oracle, refs):
    """CLDPC::Decode1 (CLDPC.cpp:2303-4383, never called by CSimulate) = NMS without the hard-coded zeroing of the last
    384 code bits when _PunctureBits = _ShortenBits = 0 (as shipped): reproduced by puncture_tail = 0."""
    cfg = oracle.default_config(0)
    fix, _ = llrgen.qpsk_llr_groups(2, 3.5, seed=77)
    dec_r, _ = refs["faid3"].decode(cfg, fix, method_override=101)
    cfg.puncture_tail = 0
    dec_o, _ = oracle.decode(cfg, fix)
    assert int((dec_r != dec_o).sum()) == 0


EF2_CASES = [(2, 7, 4, 2), (1, 7, 5, 3), (1, 5, 6, 2), (1, 7, 6, 6)]  # (|LLR| of good bits, of flipped bits, flips per frame, MaxIteration)


@pytest.mark.parametrize("mag_ok,mag_bad,n_flip,max_iter", EF2_CASES)
def test_erasure_mode_ef_elimination_2(oracle, refs, mag_ok, mag_bad, n_flip, max_iter):
    """EF_ELIMINATION 2 (CDecoder_FAID.cpp:6,200-203,623-628,673-680; reference built with that one #define changed):
    error-floor LUT with floor_err_count 20 plus erasure of regular VNs whose three checks are all unsatisfied.  The
    inputs put a handful of strongly wrong bits into otherwise clean frames, so that the erasures fire."""
    if "ef2" not in refs:
        pytest.skip("oracle/_ref/libldpc_ref_ef2.so not built")
    cfg = oracle.default_config(2, 0)
    cfg.max_iteration = max_iter
    cfg.ef_elimination, cfg.ef_floor_err_count, cfg.ef_floor_iter_thresh = 2, 20, 6
    fix = np.concatenate([llrgen.sparse_error_groups(mag_ok, mag_bad, n_flip, seed=1), llrgen.qpsk_llr_groups(1, 4.0, seed=9)[0]])
    dec_r, _ = refs["ef2"].decode(cfg, fix)
    dec_o, _ = oracle.decode(cfg, fix)
    assert int((dec_r != dec_o).sum()) == 0
    cfg.ef_elimination = 1
    dec_1, _ = oracle.decode(cfg, fix)
    assert (dec_1 != dec_o).any(), "the erasures must change something on these inputs"
    # the hybrid decoder compiled with the same define has the thresholds of mode 2 but no erasure (CDecoder_FAID_2B1C.cpp:120-123)
    cfg5 = oracle.default_config(5, -1)
    cfg5.max_iteration = max_iter
    cfg5.ef_elimination, cfg5.ef_floor_err_count, cfg5.ef_floor_iter_thresh = 2, 20, 6
    dec_r5, _ = refs["ef2"].decode(cfg5, fix)
    dec_o5, _ = oracle.decode(cfg5, fix)
    assert int((dec_r5 != dec_o5).sum()) == 0


def test_full_int8_range_inputs(oracle, refs):
    """LLRs outside [-31, 31] (no quantiser of the reference produces them, but fixInput is a public int8 buffer): the
    reference's 8-bit saturating arithmetic decides, and the oracle reproduces it.  Also pins the equivalence the CUDA
    loader relies on: clamping the channel values to [-31, +39] (min-sum family) / [-31, +31] (FAID family) changes nothing,
    while clamping the min-sum family at +31 does (v is not clamped above, CLDPC.cpp:330)."""
    rng = np.random.default_rng(5)
    fix, _ = llrgen.qpsk_llr_groups(1, 3.8, seed=9)
    scaled = np.clip(fix.astype(np.int32) * rng.integers(1, 19, size=fix.shape), -128, 127).astype(np.int8)
    rnd = rng.integers(-128, 128, size=fix.shape).astype(np.int8)
    x = np.concatenate([scaled, rnd])
    nms_differs = False
    for method in range(6):
        cfg = oracle.default_config(method)
        dec_r, bf_r = refs["faid3"].decode(cfg, x)
        dec_o, infos = oracle.decode(cfg, x)
        assert int((dec_r != dec_o).sum()) == 0, method
        if method in (3, 4):
            assert bf_r == [i.bf_iters for i in infos]
        lo, hi = (-31, 39) if method in (0, 1, 3, 4) else (-31, 31)
        dec_c, infos_c = oracle.decode(cfg, np.clip(x, lo, hi))
        assert int((dec_c != dec_o).sum()) == 0, method
        assert [i.bf_iters for i in infos] == [i.bf_iters for i in infos_c]
        assert [i.iters_executed for i in infos] == [i.iters_executed for i in infos_c]
        if method == 0:
            nms_differs = bool((oracle.decode(cfg, np.clip(x, -31, 31))[0] != dec_o).any())
    assert nms_differs


def test_bpsk_mapping_and_receive(oracle, refs):
    """BPSK (modType 1): CModulate::BPSKModulation (CModulate.cpp:363-370) and the receive side of CSimulate.cpp:121-124
    (float2LimitChar_4bit straight on the noisy amplitudes).  Only the MKL MT2203 noise stream itself is unpinned."""
    cfg = oracle.default_config(1)
    cfg.mod_type = 1
    sim = pyoracle.RefSim(refs["faid3"], cfg, seed=101)
    rng = np.random.default_rng(2)
    tx = oracle.encode_group(rng.integers(0, 2, 32 * K, dtype=np.int8))
    x_ref = sim.bpsk_modulate(tx)
    x_o = oracle.bpsk_modulate(tx)
    assert (x_ref == x_o).all() and set(np.unique(x_o)) == {-1.0, 1.0}
    noisy = (x_o + rng.standard_normal(x_o.size).astype(np.float32) * oracle.sigma(4.0, 1)).astype(np.float32)
    assert (sim.bpsk_receive(noisy, cfg.scale) == oracle.quantize(noisy, cfg.scale)).all()
    assert abs(oracle.sigma(4.0, 1) - 1.0 / np.sqrt(2 * 0.8444444 * 10 ** 0.4)) < 1e-6
