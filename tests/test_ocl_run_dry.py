"""oracle/refbuild/ocl_run.py needs an OpenCL runtime (the GPU box has NVIDIA's); this dry run swaps the
ctypes binding for a stand-in that does nothing, so the control flow — scene preparation, per-spec program
selection, comparison, digests, the JSON log — stays exercised in the CPU suite."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref",
                                                                "ocl_program.bin")),
                                reason="oracle/_ref/ocl_program.bin not built (no reference tree)")


class FakeCL:
    device_info = {"name": "fake", "version": "none"}

    def __init__(self, loader):
        self.built = []

    def open(self):
        pass

    def build(self, source, options):
        assert source.startswith("#define STOCHASTIC_FACTOR") and "__kernel void generateThresholds" in source
        self.built.append(options)
        return {k: k for k in ("generateThresholds", "sortThresholds", "renderThresholds")}

    def buffer(self, nbytes, host=None, flags=1):
        return ctypes.c_void_p(1)

    def release(self, m):
        pass

    def set_args(self, kernel, args):
        for a in args:
            if not isinstance(a, ctypes.c_void_p):
                np.ascontiguousarray(a)

    def launch(self, kernel, n_tiles, threads):
        assert n_tiles > 0 and threads in (64, 256)

    def finish(self):
        pass

    def read(self, m, arr):
        pass


def test_dry_run(tmp_path, monkeypatch):
    from oracle.refbuild import ocl_run
    out = tmp_path / "ocl.json"
    monkeypatch.setattr(ocl_run, "CL", FakeCL)
    monkeypatch.setattr(ocl_run, "T0", time.time(), raising=False)
    monkeypatch.setattr(sys, "argv", ["ocl_run.py", "--out", str(out), "--scenes", "small", "--variants", "reference,strict"])
    ocl_run.main()
    log = json.loads(out.read_text())
    assert "FAILED" not in [s["msg"] for s in log["steps"]]
    for variant in ("reference", "strict"):
        results = log["results"][variant]
        assert "fuzzy_circles_small_tiles" in results and "tiny_square" in results
        r = results["tiny_square"]
        assert r["pixels"] == 256 and r["digest_equals_oracle"] is False      # the stand-in renders nothing
        assert set(log["hashes"][variant]["tiny_square"]) == {"sha256", "thresholds", "counts_sha256", "bits_sha256"}
