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


@pytest.fixture(params=[False, True], ids=["fastlut", "generallut"])
def faid_general_path(request, monkeypatch):
    """FAID methods run either the monotone-LUT kernel (default for every LUT set the reference ships) or, with
    LDPC_B200_NO_FAID_FAST set when the handle is created, the general per-edge-LUT kernel."""
    if request.param:
        monkeypatch.setenv("LDPC_B200_NO_FAID_FAST", "1")
    return request.param


@pytest.mark.parametrize("method,lut,ebs", CASES)
def test_decode_matches_oracle(oracle, engine_lib, method, lut, ebs, faid_general_path):
    import ldpc_b200
    if faid_general_path and method not in (2, 5):
        pytest.skip("only the FAID methods have two kernels")
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


import golden_util as gu  # noqa: E402


@pytest.mark.parametrize("name", gu.CASE_NAMES)
def test_decode_matches_reference_golden_vectors(engine_lib, name):
    """The CUDA path against the REFERENCE's own dumped outputs (tests/golden/decode_vectors.npz)."""
    import ldpc_b200
    c = gu.case(name)
    cfg = gu.apply(ldpc_b200.default_config(c["method"], c["lut"]), c)
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(c["fix"], want_info=True)
        # the native packed-nibble / packed-bit interface gives the same bits
        hp = dec.decode_packed(ldpc_b200.pack_llr(c["fix"]))
    assert int((out != c["dec"]).sum()) == 0
    assert (ldpc_b200.unpack_hard(hp).reshape(3, -1) == out).all()
    for g in range(3):
        if c["bf"][g] >= 0:
            assert info["bf_iters"][g] == c["bf"][g]
        if c["its"][g] >= 0:
            assert info["its_per_group"][g] == c["its"][g]


def test_ragged_sizes_and_chunking(oracle, engine_lib):
    """n_groups = 0, 1, a count that is not a multiple of the chunk, device pointers and host pointers."""
    import torch
    import ldpc_b200
    cfg = ldpc_b200.default_config(4, -1)
    cfg.chunk_groups = 2
    cfg.n_streams = 3
    fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, seed=40 + i)[0] for i, eb in enumerate((3.0, 3.4, 3.8, 4.2, 4.6))])
    ref, infos = oracle.decode(oracle.default_config(4, -1), fix)
    with ldpc_b200.Decoder(cfg) as dec:
        assert dec.decode(fix[:0]).shape[0] == 0
        out1 = dec.decode(fix[:1])
        out5, info = dec.decode(fix, want_info=True)
        d_out = dec.decode(torch.from_numpy(fix).cuda())
    assert (out1[0] == ref[0]).all() and (out5 == ref).all() and (d_out.cpu().numpy() == ref).all()
    assert list(info["bf_iters"]) == [i.bf_iters for i in infos]


def test_full_size_properties(engine_lib):
    """BASELINE-size batch (1024 groups): decoding is deterministic, independent of how the batch is chunked, a
    noiseless codeword is a fixed point, and every decoded frame that claims convergence satisfies H."""
    import torch
    import ldpc_b200
    G = 1024
    cfg = ldpc_b200.default_config(2, 0)
    cw = llrgen.golden_codeword()
    tx = torch.from_numpy(np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]).astype(np.int8)).cuda().repeat(64, 1)
    cfg.chunk_groups = 1024
    with ldpc_b200.Decoder(cfg) as dec:
        fix = torch.cat([dec.generate(tx, 3.9, 5, g0 * 32, 64) for g0 in range(0, G, 64)])
        a, info = dec.decode(fix, want_info=True)
        clean = dec.generate(tx[:1], 60.0, 5, 0, 1)
        assert (dec.decode(clean).cpu().numpy().reshape(32, N) == cw).all()
    cfg.chunk_groups = 96
    with ldpc_b200.Decoder(cfg) as dec:
        b = dec.decode(fix)
    assert torch.equal(a, b)
    a = a.cpu().numpy().reshape(G * 32, N)
    conv = info["conv_iter"].reshape(-1) >= 0
    ok = (a[:, :K] == cw[:K]).all(1)
    assert ok.mean() > 0.9
    # frames whose min-sum iterations reported a zero syndrome and that needed no BF are codewords: check a sample
    import pyoracle
    orc = pyoracle.Oracle()
    for f in np.nonzero(conv)[0][:40]:
        if info["bf_iters"][f // 32] == 0:
            assert orc.syndrome_weight(a[f]) == 0


@pytest.mark.parametrize("method", [1, 4])
def test_simple_oms_mode_matches_oracle(oracle, engine_lib, method):
    """OMS_MODE 0 as a run-time option (config.oms_mode = 0)."""
    import ldpc_b200
    cfg = ldpc_b200.default_config(method, -1)
    cfg.oms_mode = 0
    ocfg = oracle.default_config(method, -1)
    ocfg.oms_mode = 0
    fix = np.concatenate([llrgen.qpsk_llr_groups(2, eb, scale=cfg.scale, seed=900 + i)[0] for i, eb in enumerate((3.3, 3.9))])
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(fix, want_info=True)
    ref, infos = oracle.decode(ocfg, fix)
    assert int((out != ref).sum()) == 0
    assert [i.bf_iters for i in infos] == list(info["bf_iters"])


def test_no_tail_puncture_is_decode1(oracle, engine_lib):
    """puncture_tail = 0 gives the generic initialisation of CLDPC::Decode1 (CLDPC.cpp:2303-2353) for the shipped
    _PunctureBits = _ShortenBits = 0: the last 384 code bits keep their channel LLRs."""
    import ldpc_b200
    cfg = ldpc_b200.default_config(0, -1)
    cfg.puncture_tail = 0
    ocfg = oracle.default_config(0, -1)
    ocfg.puncture_tail = 0
    fix, _ = llrgen.qpsk_llr_groups(2, 3.5, seed=77)
    with ldpc_b200.Decoder(cfg) as dec:
        out = dec.decode(fix)
    ref, _ = oracle.decode(ocfg, fix)
    assert int((out != ref).sum()) == 0
    ref384, _ = oracle.decode(oracle.default_config(0, -1), fix)
    assert (ref384 != ref).any()


@pytest.mark.parametrize("method", [2, 3, 4, 5])
def test_generic_bf_stage_matches_oracle(oracle, engine_lib, monkeypatch, method):
    """The table-driven bit-flipping stage (taken for regular_col_weight != 3 or alpha > 1; forced here with
    LDPC_B200_NO_FAST_BF) gives the same bits and BF iteration counts as the unrolled one, i.e. as the oracle."""
    import ldpc_b200
    monkeypatch.setenv("LDPC_B200_NO_FAST_BF", "1")
    cfg = ldpc_b200.default_config(method, -1)
    fix = np.concatenate([llrgen.qpsk_llr_groups(2, eb, scale=cfg.scale, seed=300 + 7 * method + i)[0] for i, eb in enumerate((3.2, 3.8))])
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(fix, want_info=True)
    ref, infos = oracle.decode(oracle.default_config(method, -1), fix)
    assert int((out != ref).sum()) == 0
    assert [i.bf_iters for i in infos] == list(info["bf_iters"])


def test_bf_with_other_alpha_uses_generic_stage(oracle, engine_lib):
    """dtbf_alpha = 2 is outside the unrolled stage's domain: the engine must fall back to the generic one (and agree
    with the oracle)."""
    import ldpc_b200
    cfg = ldpc_b200.default_config(4, -1)
    cfg.dtbf_alpha = 2
    ocfg = oracle.default_config(4, -1)
    ocfg.dtbf_alpha = 2
    fix, _ = llrgen.qpsk_llr_groups(2, 3.4, scale=cfg.scale, seed=41)
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(fix, want_info=True)
    ref, infos = oracle.decode(ocfg, fix)
    assert int((out != ref).sum()) == 0
    assert [i.bf_iters for i in infos] == list(info["bf_iters"])


def test_two_handles_with_different_luts_coexist(oracle, engine_lib):
    """The LUT tables travel with each launch (kernel parameter bank), not in a device-global symbol: handles with
    different LUT sets, alive at the same time on one device, must not disturb each other."""
    import ldpc_b200
    fix, _ = llrgen.qpsk_llr_groups(2, 3.6, seed=8)
    a = ldpc_b200.Decoder(ldpc_b200.default_config(2, 0))   # FAID3
    b = ldpc_b200.Decoder(ldpc_b200.default_config(2, 2))   # FAID2, created later
    try:
        out_a = a.decode(fix)
        out_b = b.decode(fix)
        out_a2 = a.decode(fix)
    finally:
        a.close()
        b.close()
    ref_a, _ = oracle.decode(oracle.default_config(2, 0), fix)
    ref_b, _ = oracle.decode(oracle.default_config(2, 2), fix)
    assert (out_a == ref_a).all() and (out_a2 == ref_a).all() and (out_b == ref_b).all()
    assert (ref_a != ref_b).any()


def test_nms_direct_output_equals_finalize_path(engine_lib, monkeypatch):
    """DecodeMethod 0 writes decodedBits / packed decisions straight from the decode kernel when no per-group outputs
    are requested; with them (or with LDPC_B200_NO_DIRECT_OUTPUT) finalize_kernel formats the output.  Same bits."""
    import ldpc_b200
    fix, _ = llrgen.qpsk_llr_groups(3, 3.5, seed=21)
    cfg = ldpc_b200.default_config(0, -1)
    with ldpc_b200.Decoder(cfg) as dec:
        direct = dec.decode(fix)
        direct_packed = dec.decode_packed(ldpc_b200.pack_llr(fix))
        via_finalize, info = dec.decode(fix, want_info=True)
        assert dec.last_timing()[1] == 2  # decode + finalize launches
        dec.decode(fix)
        assert dec.last_timing()[1] == 1  # decode only
    monkeypatch.setenv("LDPC_B200_NO_DIRECT_OUTPUT", "1")
    with ldpc_b200.Decoder(cfg) as dec:
        forced = dec.decode(fix)
    assert (direct == via_finalize).all() and (forced == direct).all()
    assert (ldpc_b200.unpack_hard(direct_packed).reshape(3, -1) == direct).all()
    assert list(info["its_per_group"]) == [6, 6, 6] and (info["conv_iter"] == -1).all()


@pytest.mark.parametrize("method,max_iter", [(0, 1), (1, 1), (2, 64), (4, 20)])
def test_iteration_limits(oracle, engine_lib, method, max_iter):
    """MaxIteration = 1 (a single pass, the early-stop bookkeeping still has to hold) and the maximum of 64."""
    import ldpc_b200
    cfg = ldpc_b200.default_config(method, -1)
    cfg.max_iteration = max_iter
    ocfg = oracle.default_config(method, -1)
    ocfg.max_iteration = max_iter
    fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, scale=cfg.scale, seed=70 + i)[0] for i, eb in enumerate((3.3, 4.3))])
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(fix, want_info=True)
    ref, infos = oracle.decode(ocfg, fix)
    assert int((out != ref).sum()) == 0
    assert [i.iters_executed for i in infos] == list(info["its_per_group"])
    assert [i.bf_iters for i in infos] == list(info["bf_iters"])


def test_runtime_setters_and_argument_errors(oracle, engine_lib):
    """ldpc_b200_set_factors (the reference re-reads Factor_1/2 from Profile.txt in every call, CLDPC.cpp:216-222) and the
    negative status codes for unusable arguments."""
    import ctypes as C
    import ldpc_b200
    fix, _ = llrgen.qpsk_llr_groups(1, 3.5, seed=3)
    cfg = ldpc_b200.default_config(0, -1)
    with ldpc_b200.Decoder(cfg) as dec:
        dec.set_factors(22, 29)
        out = dec.decode(fix)
        lib = dec.lib
        bad = np.zeros(32 * N + 8, dtype=np.int8)
        res = np.zeros(32 * N + 32, dtype=np.int8)
        rc = lib.ldpc_b200_decode(dec.h, C.c_void_p(bad.ctypes.data + 1), C.c_void_p(res.ctypes.data), 1, None, None, None)
        assert rc == -1 and b"aligned" in lib.ldpc_b200_last_error()
        assert lib.ldpc_b200_decode(dec.h, None, C.c_void_p(res.ctypes.data), 1, None, None, None) == -1
        assert lib.ldpc_b200_set_factors(dec.h, 1, 2) == 0  # still a valid NMS configuration
        assert lib.ldpc_b200_set_max_iteration(dec.h, 65) == -1
    ocfg = oracle.default_config(0, -1)
    ocfg.factor_1, ocfg.factor_2 = 22, 29
    ref, _ = oracle.decode(ocfg, fix)
    assert int((out != ref).sum()) == 0


def test_erasure_mode_matches_oracle(oracle, engine_lib):
    """EF_ELIMINATION 2 (CDecoder_FAID.cpp:673-680; oracle pinned to the reference compiled with that define in
    tests/test_oracle_vs_reference.py): groups with a few strongly wrong bits, where the erasures fire, and noisy groups;
    the hybrid decoder with the same setting only changes its thresholds."""
    import ldpc_b200
    fix = np.concatenate([llrgen.sparse_error_groups(mo, mb, nf, seed=3 + i) for i, (mo, mb, nf) in
                          enumerate([(2, 7, 4), (1, 7, 5), (1, 5, 6), (1, 7, 6), (2, 6, 10), (3, 7, 12)])]
                         + [llrgen.qpsk_llr_groups(2, 3.7, seed=41)[0]])
    for method, lut, max_iter in [(2, 0, 2), (2, 0, 3), (2, 0, 6), (2, 1, 15), (5, 3, 3), (5, 3, 15)]:
        cfg = ldpc_b200.default_config(method, lut)
        ocfg = oracle.default_config(method, lut)
        for c in (cfg, ocfg):
            c.max_iteration = max_iter
            c.ef_elimination, c.ef_floor_err_count, c.ef_floor_iter_thresh = 2, 20, 6
        with ldpc_b200.Decoder(cfg) as dec:
            out, info = dec.decode(fix, want_info=True)
        ref, infos = oracle.decode(ocfg, fix)
        assert int((out != ref).sum()) == 0, (method, lut, max_iter)
        assert [i.bf_iters for i in infos] == list(info["bf_iters"])
        assert [i.iters_executed for i in infos] == list(info["its_per_group"])
        if method == 2 and max_iter <= 3:
            ocfg.ef_elimination = 1
            assert (oracle.decode(ocfg, fix)[0] != ref).any(), "the erasures must change something on these inputs"
    cfg = ldpc_b200.default_config(2, 0)
    cfg.ef_elimination, cfg.regular_col_weight = 2, 6
    with pytest.raises(Exception):
        ldpc_b200.Decoder(cfg)


@pytest.mark.parametrize("method", [0, 2, 5])
def test_host_staging_variants_are_identical(oracle, engine_lib, monkeypatch, method):
    """ldpc_b200_decode with HOST buffers: direct PCIe copies of the caller's arrays vs bit-packed decisions expanded by the
    host threads (default) vs nibble-packed LLRs on top; pageable and pinned arrays; several chunks per call; a chunk whose
    LLRs do not fit a nibble (6-bit quantiser range) falls back to bytes."""
    import ldpc_b200
    fix = np.concatenate([llrgen.qpsk_llr_groups(7, 3.5, seed=60 + method)[0]])
    wide = fix.copy()
    wide[3, ::97] = np.where(wide[3, ::97] > 0, 21, -19).astype(np.int8)  # group 3 no longer fits 4 bits
    results = {}
    for name, env in (("direct", {"LDPC_B200_HOST_THREADS": "0"}), ("stage_out", {"LDPC_B200_STAGE_OUT": "1", "LDPC_B200_STAGE_IN": "0", "LDPC_B200_HOST_THREADS": "5"}),
                      ("stage_in_out", {"LDPC_B200_STAGE_IN": "1", "LDPC_B200_STAGE_OUT": "1", "LDPC_B200_HOST_THREADS": "3"}),
                      ("default", {})):
        for k in ("LDPC_B200_HOST_THREADS", "LDPC_B200_STAGE_OUT", "LDPC_B200_STAGE_IN"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        cfg = ldpc_b200.default_config(method, -1)
        cfg.chunk_groups, cfg.n_streams = 2, 3
        with ldpc_b200.Decoder(cfg) as dec:
            st = dec.host_staging()
            out, info = dec.decode(fix, want_info=True)
            out_w = dec.decode(wide)
            pin_in, pin_out = ldpc_b200.PinnedArray(fix.shape, np.int8), ldpc_b200.PinnedArray(fix.shape, np.int8)
            pin_in.array[:] = fix
            dec.decode(pin_in.array, pin_out.array)
            st2 = dec.host_staging()
            assert (pin_out.array == out).all()
        if name == "direct":
            assert st["threads"] == 0 and not st["stage_out"]
            assert (st2["last_h2d_bytes"], st2["last_d2h_bytes"]) == (fix.size, fix.size)
        elif name == "stage_out":
            assert st["threads"] == 5 and st["stage_out"] and not st["stage_in"]
            assert (st2["last_h2d_bytes"], st2["last_d2h_bytes"]) == (fix.size, fix.size // 8)
        elif name == "stage_in_out":
            assert (st2["last_h2d_bytes"], st2["last_d2h_bytes"]) == (fix.size // 2, fix.size // 8)
        else:  # default: decisions as bits with >= 4 host threads, LLRs as nibbles when the process has >= 8 cores to itself
            assert (not st["stage_in"] or st["stage_out"]) and (st["threads"] >= 1) == st["stage_out"]
        results[name] = (out, out_w, list(info["bf_iters"]))
    ref, _ = oracle.decode(oracle.default_config(method, -1), fix)
    ref_w, _ = oracle.decode(oracle.default_config(method, -1), wide)
    for name, (out, out_w, bf) in results.items():
        assert (out == ref).all(), name
        assert (out_w == ref_w).all(), name
        assert bf == results["direct"][2]


def test_multi_rank_default_stages_only_the_decisions(oracle, engine_lib, monkeypatch):
    """Several ranks on one host (LOCAL_WORLD_SIZE, as torchrun sets it): each handle takes its share of the host threads, the
    decisions return as bits and fixInput is copied as it is (packing would double the host memory traffic of the input)."""
    import os
    import ldpc_b200
    for k in ("LDPC_B200_HOST_THREADS", "LDPC_B200_STAGE_OUT", "LDPC_B200_STAGE_IN", "LDPC_B200_HYBRID"):
        monkeypatch.delenv(k, raising=False)
    cores = len(os.sched_getaffinity(0))
    if cores < 8:
        pytest.skip("needs >= 8 host CPUs to give two ranks 4 threads each")
    fix = llrgen.qpsk_llr_groups(6, 3.6, seed=77)[0]
    ref, _ = oracle.decode(oracle.default_config(0, -1), fix)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    cfg = ldpc_b200.default_config(0, -1)
    cfg.chunk_groups, cfg.n_streams = 2, 3
    with ldpc_b200.Decoder(cfg) as dec:
        st = dec.host_staging()
        assert st["stage_out"] and not st["stage_in"] and 1 <= st["threads"] <= max(1, cores // 2)
        out = dec.decode(fix)
        st2 = dec.host_staging()
        assert (st2["last_h2d_bytes"], st2["last_d2h_bytes"]) == (fix.size, fix.size // 8)
        assert (out == ref).all()
    monkeypatch.setenv("LOCAL_WORLD_SIZE", str(cores))  # one thread per rank: everything is copied as it is
    with ldpc_b200.Decoder(cfg) as dec:
        st = dec.host_staging()
        assert not st["stage_out"] and not st["stage_in"]
        assert (dec.decode(fix) == ref).all()


def test_device_input_alignment_paths(engine_lib):
    """LLRs resident on the device at 4-, 8- and 16-byte aligned addresses (the loader reads words): same decisions."""
    import torch
    import ldpc_b200
    fix = llrgen.qpsk_llr_groups(3, 3.6, seed=77)[0]
    for method in (0, 1):
        cfg = ldpc_b200.default_config(method, -1)
        with ldpc_b200.Decoder(cfg) as dec:
            outs = []
            for off in (0, 4, 8, 16):
                buf = torch.zeros(fix.size + 64, dtype=torch.int8, device="cuda")
                view = buf[off: off + fix.size]
                view.copy_(torch.from_numpy(fix.reshape(-1)))
                assert view.data_ptr() % 16 == off % 16
                outs.append(dec.decode(view.view(3, -1)).cpu().numpy())
            for o in outs[1:]:
                assert (o == outs[0]).all()


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_full_int8_range(oracle, engine_lib, method, faid_general_path):
    """fixInput values outside [-31, 31]: the engine reproduces the reference's 8-bit saturating behaviour (the oracle is
    pinned to the compiled reference on the same inputs, tests/test_oracle_vs_reference.py::test_full_int8_range_inputs)."""
    import ldpc_b200
    if faid_general_path and method not in (2, 5):
        pytest.skip("only the FAID methods have two kernels")
    rng = np.random.default_rng(50 + method)
    fix, _ = llrgen.qpsk_llr_groups(2, 3.8, seed=9)
    scaled = np.clip(fix.astype(np.int32) * rng.integers(1, 19, size=fix.shape), -128, 127).astype(np.int8)
    rnd = rng.integers(-128, 128, size=fix.shape).astype(np.int8)
    edge = np.tile(np.array([-128, -127, -39, -32, -31, 31, 32, 38, 39, 40, 126, 127], dtype=np.int8), 32 * N // 12)[None, :]
    x = np.concatenate([scaled, rnd, edge])
    cfg = ldpc_b200.default_config(method, -1)
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(x, want_info=True)
        import torch
        d_out = dec.decode(torch.from_numpy(x).cuda()).cpu().numpy()
    ref, infos = oracle.decode(oracle.default_config(method, -1), x)
    assert int((out != ref).sum()) == 0 and (d_out == ref).all()
    assert [i.bf_iters for i in infos] == list(info["bf_iters"])
    assert [i.iters_executed for i in infos] == list(info["its_per_group"])


def test_handles_of_different_methods_coexist(oracle, engine_lib):
    """finalize_kernel's dynamic shared memory limit is a per-device function attribute: a later handle with a smaller need
    (NMS / OMS) must not lower it under an earlier one that runs the 2B1C stage (224 KB)."""
    import ldpc_b200
    fix, _ = llrgen.qpsk_llr_groups(2, 3.4, scale=12.5, seed=8)
    a = ldpc_b200.Decoder(ldpc_b200.default_config(5, -1))
    b = ldpc_b200.Decoder(ldpc_b200.default_config(1, -1))   # created later, needs 70 KB only
    c = ldpc_b200.Decoder(ldpc_b200.default_config(0, -1))
    try:
        out_b, _ = b.decode(fix, want_info=True)
        out_c, _ = c.decode(fix, want_info=True)
        out_a, info_a = a.decode(fix, want_info=True)
    finally:
        a.close(); b.close(); c.close()
    for m, out in ((5, out_a), (1, out_b), (0, out_c)):
        ref, infos = oracle.decode(oracle.default_config(m, -1), fix)
        assert (out == ref).all(), m
    assert max(info_a["bf_iters"]) > 0, "the point should exercise the 2B1C stage"


def test_max_iteration_beyond_64_and_runtime_lowering(oracle, engine_lib):
    """MaxIteration has no cap in the reference (an int from Profile.txt); here the bound is 1000.  ldpc_b200_set_max_iteration
    may move freely below the creation-time value (the scratch is sized for that one)."""
    import ldpc_b200
    fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, seed=90 + i)[0] for i, eb in enumerate((2.9, 3.5))])
    for method, mi in ((1, 100), (0, 80)):
        cfg = ldpc_b200.default_config(method, -1)
        cfg.max_iteration = mi
        ocfg = oracle.default_config(method, -1)
        with ldpc_b200.Decoder(cfg) as dec:
            for cur in (mi, 3, 70, mi):
                assert dec.lib.ldpc_b200_set_max_iteration(dec.h, cur) == 0
                ocfg.max_iteration = cur
                out, info = dec.decode(fix, want_info=True)
                ref, infos = oracle.decode(ocfg, fix)
                assert (out == ref).all(), (method, cur)
                assert [i.iters_executed for i in infos] == list(info["its_per_group"])
                assert (np.array([list(i.conv_iter) for i in infos]) == info["conv_iter"]).all()
            assert dec.lib.ldpc_b200_set_max_iteration(dec.h, mi + 1) == -1
    cfg.max_iteration = 1001
    with pytest.raises(Exception):
        ldpc_b200.Decoder(cfg)


@pytest.mark.parametrize("method,out_bits", [(0, None), (4, None), (0, "0"), (4, "0")])
def test_hybrid_host_path_routes_chunks_both_ways(oracle, engine_lib, monkeypatch, method, out_bits):
    """Pinned caller arrays + staging on + enough chunks: the library packs the LLRs of some chunks on the host threads and lets
    the copy engine move the others as they are, concurrently (ldpc_b200_last_routing); decisions of every chunk come back as
    bits (default) or, with LDPC_B200_HYBRID_OUT_BITS=0, as bytes for the directly copied chunks.  Bits and per-group outputs
    must not depend on the route a chunk took.  LDPC_B200_HYBRID = number of direct slots (default 1); 0 stages everything."""
    import ldpc_b200
    G = 26
    fix = np.concatenate([llrgen.qpsk_llr_groups(G // 2, eb, seed=500 + method + i)[0] for i, eb in enumerate((3.4, 3.9))])
    ref, infos = oracle.decode(oracle.default_config(method, -1), fix)
    for k in ("LDPC_B200_HOST_THREADS", "LDPC_B200_STAGE_OUT", "LDPC_B200_STAGE_IN", "LDPC_B200_HYBRID", "LDPC_B200_HYBRID_OUT_BITS",
              "LDPC_B200_HYBRID_CHUNK"):
        monkeypatch.delenv(k, raising=False)
    if out_bits is not None:
        monkeypatch.setenv("LDPC_B200_HYBRID_OUT_BITS", out_bits)
    monkeypatch.setenv("LDPC_B200_HYBRID", "2")
    monkeypatch.setenv("LDPC_B200_HOST_THREADS", "4")
    monkeypatch.setenv("LDPC_B200_STAGE_IN", "1")
    monkeypatch.setenv("LDPC_B200_STAGE_OUT", "1")
    cfg = ldpc_b200.default_config(method, -1)
    cfg.chunk_groups, cfg.n_streams = 2, 3
    pin_in, pin_out = ldpc_b200.PinnedArray(fix.shape, np.int8), ldpc_b200.PinnedArray(fix.shape, np.int8)
    pin_in.array[:] = fix
    with ldpc_b200.Decoder(cfg) as dec:
        for rep in range(3):
            pin_out.array[:] = 9
            out, info = dec.decode(pin_in.array, pin_out.array, want_info=True)
            r = dec.last_routing()
            assert r["staged_chunks"] + r["direct_chunks"] == G // 2 and r["direct_chunks"] >= 2 and r["staged_chunks"] >= 1, r
            assert (out == ref).all()
            assert [i.bf_iters for i in infos] == list(info["bf_iters"])
            assert [i.iters_executed for i in infos] == list(info["its_per_group"])
        # pageable arrays cannot be copied directly: everything is staged
        out_p = dec.decode(fix)
        assert dec.last_routing()["direct_chunks"] == 0 and (out_p == ref).all()
    monkeypatch.setenv("LDPC_B200_HYBRID", "0")
    with ldpc_b200.Decoder(cfg) as dec:
        out = dec.decode(pin_in.array, pin_out.array)
        assert dec.last_routing() == {"staged_chunks": G // 2, "direct_chunks": 0} and (out == ref).all()


def test_pageable_arrays_page_locked_on_first_use(oracle, engine_lib):
    """LDPC_B200_HOST_REGISTER=1 (read once per process, hence a fresh interpreter): plain numpy arrays are page-locked by
    the first call and reused by address; bits equal the oracle's with staging on and off, repeated calls, a second buffer pair."""
    import subprocess, sys, os, json
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = r'''
import sys, os, json
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(%r, p))
import numpy as np, ctypes as C
import ldpc_b200, llrgen, pyoracle
orc = pyoracle.Oracle()
fix = np.concatenate([llrgen.qpsk_llr_groups(3, eb, seed=81 + i)[0] for i, eb in enumerate((3.4, 4.0))])
ref, _ = orc.decode(orc.default_config(0, -1), fix)
res = []
for threads in ("0", "4"):
    os.environ["LDPC_B200_HOST_THREADS"] = threads
    cfg = ldpc_b200.default_config(0, -1); cfg.chunk_groups, cfg.n_streams = 2, 3
    with ldpc_b200.Decoder(cfg) as dec:
        a, b = fix.copy(), np.empty_like(fix)
        for rep in range(3):
            b[:] = 7
            dec.decode(a, b)
            res.append(bool((b == ref).all()))
        a2, b2 = fix[::-1].copy(), np.empty_like(fix)
        dec.decode(a2, b2)
        res.append(bool((b2 == ref[::-1]).all()))
        attr_ok = True
print("REG " + json.dumps(res))
''' % str(root)
    env = dict(os.environ, LDPC_B200_HOST_REGISTER="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("REG ")][-1][4:])
    assert len(res) == 8 and all(res)
