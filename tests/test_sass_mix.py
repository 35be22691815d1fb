"""Static instruction mix of the shipped kernels (tools/sass_mix.py, lib/sass_mix.json): the numbers bench.py turns into the pipe
roofline must come from the library that is loaded, cover every kernel kind, and stay in a sane range -- a loop-detection slip
(ptxas gives the iteration loop several back edges) once made an OMS iteration look 25 % shorter than it is."""
import hashlib
import json
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "mod-interleaveavx_multithreads-faid_b200" / "lib" / "libldpc_b200.so"
KINDS = ("NMS", "NMS_general_scale", "OMS", "FAID", "FAID_EF", "FAID_M", "FAID_EF_M", "FAID_ER")


@pytest.fixture(scope="module")
def mix(engine_lib):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not installed")
    sys.path.insert(0, str(ROOT / "tools"))
    import sass_mix
    return sass_mix.mix_of(LIB)


def test_every_kind_has_an_iteration_loop_of_plausible_size(mix):
    kinds = mix["kinds"]
    assert set(KINDS) <= set(kinds)
    for k in KINDS:
        v = kinds[k]
        pe = v["per_edge"]
        # 275 edges per iteration and row thread; the leanest kernel (NMS) needs ~20 instructions per edge, the general FAID ~45
        assert 18.0 < v["per_edge_total"] < 50.0, (k, v["per_edge_total"])
        assert 8.0 < pe["alu"] < 30.0 and 5.0 < pe["fma"] < 16.0 and 2.0 < pe["lsu"] < 8.0, (k, pe)
        assert abs(sum(pe.values()) - v["per_edge_total"]) < 0.01
        assert v["loop_instructions"] == round(v["per_edge_total"] * 275)
    # the kinds with a start-of-iteration syndrome walk all 275 edges twice: they cannot be cheaper than NMS
    assert kinds["OMS"]["per_edge_total"] > kinds["NMS"]["per_edge_total"] + 3.0
    assert kinds["FAID_M"]["per_edge"]["alu"] < kinds["FAID"]["per_edge"]["alu"]  # the monotone-LUT path is the cheaper one
    assert kinds["NMS"]["per_edge"]["alu"] <= kinds["NMS_general_scale"]["per_edge"]["alu"]


def test_json_next_to_the_library_belongs_to_it(engine_lib, mix):
    p = LIB.parent / "sass_mix.json"
    assert p.exists(), "build.py writes lib/sass_mix.json after linking"
    d = json.loads(p.read_text())
    assert d["lib_sha256"] == hashlib.sha256(LIB.read_bytes()).hexdigest()
    for k in KINDS:
        assert d["kinds"][k]["per_edge"] == mix["kinds"][k]["per_edge"], k
