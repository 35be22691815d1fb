"""CPU-side checks of the drop-in boundary: the shared library loads, exports every declared symbol, the
configuration defaults equal the reference's shipped constants, the Profile.txt parser follows ReadProfile."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(engine_lib):
    from ldpc_b200 import abi
    header = (ROOT / "include" / "ldpc_b200.h").read_text()
    declared = set(re.findall(r"LDPC_B200_API\s+[\w\s\*]+?\b(ldpc_b200_\w+)\(", header))
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    for name in declared:
        assert getattr(engine_lib, name) is not None


def test_struct_layout_matches_header(engine_lib):
    from ldpc_b200 import abi
    import ldpc_b200
    cfg = ldpc_b200.default_config(0)
    assert cfg.struct_size == C.sizeof(abi.Config)
    assert cfg.abi_version == abi.ABI_VERSION


@pytest.mark.parametrize("method", range(6))
@pytest.mark.parametrize("lut", [-1, 0, 1, 2, 3])
def test_default_config_equals_oracle_constants(engine_lib, oracle, method, lut):
    """Two independent restatements of the reference's #defines (product vs oracle) must agree byte for byte."""
    import ldpc_b200
    a, b = ldpc_b200.default_config(method, lut), oracle.default_config(method, lut)
    assert bytes(a) == bytes(b)


def test_no_device_is_a_loud_error_not_a_fallback(engine_lib):
    import torch
    import ldpc_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ldpc_b200.LdpcError) as e:
        ldpc_b200.Decoder(ldpc_b200.default_config(0))
    assert e.value.code == ldpc_b200.abi.ENODEV


def test_read_profile_token_order(engine_lib, tmp_path):
    """Same positional parser as ReadProfile (CTool.cpp:597-616): labels are ignored, order is everything."""
    import ldpc_b200
    p = tmp_path / "Profile.txt"
    p.write_text("Simulation parameter\nStartSNR: 2.5\nSNRPass: 0.25\nEndSNR: 4\nDecodeMethod: 5\nMaxIteration: 9\n"
                 "Modulation Parameter:\nmodType: 4\nInterleaveModType: 4\nNMS  Factor:\nFactor_1: 2\nFactor_2: 5\n"
                 "noFrames: 32\nscale: 12.5\nMatrix Factor\nFileName: 50GPON-CP12\nZ: 256\n")
    cfg = ldpc_b200.read_profile(p)
    assert (cfg.snr_start, cfg.snr_pass, cfg.snr_end) == (2.5, 0.25, 4.0)
    assert (cfg.decode_method, cfg.max_iteration, cfg.mod_type, cfg.interleave_mod_type) == (5, 9, 4, 4)
    assert (cfg.factor_1, cfg.factor_2, cfg.nb_frames, cfg.scale, cfg.Z) == (2, 5, 32, 12.5, 256)
    # method-dependent constants follow DecodeMethod
    assert (cfg.bf_mode, cfg.bf_max_iter, cfg.dtbf_L0, cfg.ef_elimination) == (ldpc_b200.BF_2B1C, 10, 100, 1)
    shipped = ldpc_b200.read_profile(ROOT / "tests" / "golden" / "Profile_shipped.txt")
    assert (shipped.decode_method, shipped.max_iteration, shipped.factor_1, shipped.factor_2, shipped.scale) == (2, 6, 1, 6, 13.0)
    with pytest.raises(ldpc_b200.LdpcError):
        ldpc_b200.read_profile(tmp_path / "missing.txt")


def test_validation_rejects_unsupported(engine_lib):
    import ldpc_b200
    lib = engine_lib
    for field, val in (("nb_frames", 16), ("Z", 128), ("max_iteration", 1001), ("max_iteration", -1), ("mod_type", 3), ("interleave_mod_type", 5)):
        cfg = ldpc_b200.default_config(1)
        setattr(cfg, field, val)
        h = C.c_void_p()
        assert lib.ldpc_b200_create(C.byref(cfg), C.byref(h)) == ldpc_b200.abi.EINVAL, field


def test_pack_helpers_roundtrip():
    import ldpc_b200
    rng = np.random.default_rng(0)
    fix = rng.integers(-7, 8, (2, 32 * 17664), dtype=np.int8)
    p = ldpc_b200.pack_llr(fix)
    assert p.shape == (64, 17664 // 2)
    lo = ((p & 0xF).astype(np.int8) ^ 8) - 8
    assert (lo[0, : 14592 // 2] == fix[0, :14592][0::2]).all()
    bits = rng.integers(0, 2, (3, 17664), dtype=np.uint8)
    words = np.packbits(bits, axis=1, bitorder="little").view(np.uint32)
    assert (ldpc_b200.unpack_hard(words) == bits).all()
