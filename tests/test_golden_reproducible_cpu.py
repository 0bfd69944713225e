"""CPU only, and only where the reference checkout exists (the build container; skipped on the GPU box): the committed
fixtures are what the REAL reference code produces today - regenerate every family into a scratch directory with
oracle/gen_golden.py and compare array by array."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'local_utils')), reason='reference checkout not present')


@pytest.mark.parametrize('family,files', [
    ('kmeans', ['kmeans_blobs_demo.npz', 'kmeans_small.npz', 'kmeans_empty_cluster.npz']),
    ('constrained', ['kmeans_constrained.npz']),
    ('hungarian', ['hungarian.npz']),
    ('naming', ['naming_small.npz']),
    ('eval', ['eval_small.npz']),
])
def test_fixture_family_regenerates_bit_identically(tmp_path, golden_dir, family, files):
    r = subprocess.run([sys.executable, '-m', 'oracle.gen_golden', '--ref', REF, '--only', family, '--out', str(tmp_path)],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for f in files:
        new, old = np.load(tmp_path / f), np.load(os.path.join(golden_dir, f))
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            same = np.array_equal(old[k], new[k], equal_nan=True) if old[k].dtype.kind == 'f' else np.array_equal(old[k], new[k])
            assert old[k].dtype == new[k].dtype and same, (f, k)
