"""GPU: libmdzcuda through its C ABI against the golden fixtures produced by the
unmodified reference cmdline (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

import golden_util as G
import mdz_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", G.names())
def test_cuda_reproduces_reference_raw(name):
    meta, raw, rgb = G.load(name)
    view, info = G.view_of(meta)
    try:
        got = mdz_b200.render(view)
    except mdz_b200.MdzCudaError as ex:
        if "GMP" in str(ex):
            pytest.skip(str(ex))
        raise
    bad = int((got != raw).sum())
    assert bad == 0, "%d of %d pixels differ" % (bad, raw.size)
