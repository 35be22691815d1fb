"""Own bounds checks in place of compute-sanitizer (closed on this GPU pool): a library built with -DLDPC_DEBUG_BOUNDS=1
range-checks, on the device, every APP / message-word access of the layer code, the snapshot / hard-decision / LLR indices and
the finalize kernel's slices.  All six DecodeMethods (both FAID kernels, erasure kind, generic BF stage) run small cases with
it, must report zero violations AND still be bit-exact against the oracle; a deliberately corrupted launch must be reported."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "mod-interleaveavx_multithreads-faid_b200"
DBG_LIB = ROOT / "build" / "variants" / "bounds.so"

WORKER = r"""
import json, os, sys
root = %(root)r
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(root, p))
import numpy as np
import ldpc_b200, llrgen, pyoracle
orc = pyoracle.Oracle()
res = {}
cw = llrgen.golden_codeword()
cases = [(0, -1, {}), (1, -1, {}), (2, 0, {}), (2, 2, {"LDPC_B200_NO_FAID_FAST": "1"}), (3, -1, {}), (4, -1, {}), (4, -1, {"LDPC_B200_NO_FAST_BF": "1"}),
         (5, 3, {}), (5, 3, {"LDPC_B200_NO_FAID_FAST": "1"}), ("er", 0, {})]
for method, lut, env in cases:
    for k in ("LDPC_B200_NO_FAID_FAST", "LDPC_B200_NO_FAST_BF"):
        os.environ.pop(k, None)
    os.environ.update(env)
    m = 2 if method == "er" else method
    cfg, ocfg = ldpc_b200.default_config(m, lut), orc.default_config(m, lut)
    if method == "er":
        for c in (cfg, ocfg):
            c.ef_elimination, c.ef_floor_err_count, c.ef_floor_iter_thresh = 2, 20, 6
    cfg.chunk_groups, cfg.n_streams = 2, 2
    fix = np.concatenate([llrgen.qpsk_llr_groups(1, eb, scale=cfg.scale, seed=700 + 3 * m + i)[0] for i, eb in enumerate((3.2, 3.7, 4.3))])
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(fix, want_info=True)
        packed = dec.decode_packed(ldpc_b200.pack_llr(fix))
        c = dec.simulate(3.6, 5, 0, 3, codeword=cw)
        c2 = dec.simulate(3.6, 5, 0, 3)
        rec = dec.debug_bounds()
    ref, infos = orc.decode(ocfg, fix)
    res[f"{method}/{lut}/{sorted(env)}"] = {"rec": rec, "exact": bool((out == ref).all() and (ldpc_b200.unpack_hard(packed).reshape(3, -1) == ref).all()),
                                            "bf": [int(x) for x in info["bf_iters"]] == [i.bf_iters for i in infos], "frames": int(c[0] + c2[0])}
# negative control: one out-of-range offset goes through the checker
os.environ["LDPC_B200_DEBUG_FAULT"] = "1"
with ldpc_b200.Decoder(ldpc_b200.default_config(0, -1)) as dec:
    dec.decode(fix)
    res["fault"] = dec.debug_bounds()
print("BOUNDS " + json.dumps(res))
"""


def test_debug_bounds_build_reports_no_violation(engine_lib, tmp_path):
    import ldpc_b200
    r = subprocess.run([sys.executable, str(PKG / "build.py"), f"--out={DBG_LIB}", "-DLDPC_DEBUG_BOUNDS=1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    script = tmp_path / "bounds_worker.py"
    script.write_text(WORKER % {"root": str(ROOT)})
    env = dict(os.environ, LDPC_B200_LIB=str(DBG_LIB))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("BOUNDS ")][-1][len("BOUNDS "):])
    fault = res.pop("fault")
    assert fault["violations"] == 1 and fault["first"] >> 32 == 1 and (fault["first"] & 0xFFFFFFFF) == 17664 * 4, fault  # code DBG_APP, one launch
    assert len(res) == 10
    for name, v in res.items():
        assert v["rec"]["compiled_in"] is True, name
        assert v["rec"]["violations"] == 0, (name, v["rec"])
        assert v["exact"] and v["bf"] and v["frames"] == 2 * 96, (name, v)
    # the shipped library carries no checks and says so
    with ldpc_b200.Decoder(ldpc_b200.default_config(0, -1)) as dec:
        assert dec.debug_bounds() == {"compiled_in": False, "violations": 0, "first": 0}
