"""Where the reference tree is present (the build container, not the GPU box) re-run the reference's own code
(oracle/refrun/run_reference.py) and check that it reproduces the committed fixture bit for bit, from reference
files with the recorded SHA-256.  Skipped where /root/reference does not exist."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("MSI_REFERENCE_ROOT", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "geometry")), reason="reference tree not present")
def test_committed_fixture_is_what_the_reference_code_returns(tmp_path):
    out = str(tmp_path / "rerun.npz")
    env = dict(os.environ, MSI_REFRUN_OUT=out)
    subprocess.run([sys.executable, "-m", "oracle.refrun.run_reference"], cwd=ROOT, env=env, check=True,
                   stdout=subprocess.DEVNULL, timeout=600)
    a = np.load(os.path.join(ROOT, "tests", "golden", "reference_run.npz"))
    b = np.load(out)
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        if k == "meta_json":
            continue
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
    ma, mb = (json.loads(bytes(z["meta_json"]).decode()) for z in (a, b))
    assert ma["full"] == mb["full"]
    for f, digest in ma["file_sha256"].items():
        assert hashlib.sha256(open(os.path.join(REFERENCE, f), "rb").read()).hexdigest() == digest, f
