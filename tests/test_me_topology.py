"""ocg_me_topology (host code of the product library, no GPU needed) must equal
the tables the reference builds for its own encoder context: mb_maps luma
offsets (state.c:300-330) and the cneighbors lists (encode.c:967-1048),
including frames whose macro-block column/row count is odd (partial super
blocks)."""
import ctypes as C

import numpy as np
import pytest

import megen
import support as S
from theora_b200 import abi

pytestmark = pytest.mark.skipif(not S.ref_available("c"), reason="needs oracle/_ref")

SIZES = [(16, 16, 0), (32, 16, 0), (64, 64, 0), (176, 144, 0), (352, 288, 0), (80, 48, 0), (1920, 1088, 0),
         (208, 112, 2), (144, 176, 3), (48, 240, 0)]


@pytest.mark.parametrize("size", SIZES)
def test_topology_matches_reference(size):
    fw, fh, fmt = size
    R = megen.bind_ref_me(S.ref("c"))
    L = abi.lib()
    g = S.make_geometry(fw, fh, fmt, 6)
    h = R.refh_me_open(fw, fh, fmt)
    assert h
    try:
        n = R.refh_me_nmbs(h)
        assert n == L.ocg_me_nmbs(C.byref(g))
        assert R.refh_me_frame_size(h) == g.ref_frame_sz
        want = np.zeros(n, abi.ME_TOPO_DTYPE)
        got = np.zeros(n, abi.ME_TOPO_DTYPE)
        R.refh_me_topology(h, want.ctypes.data)
        assert L.ocg_me_topology(C.byref(g), got.ctypes.data) == 0
        for f in ("valid", "ncn", "frag_off", "cn"):
            assert np.array_equal(got[f], want[f]), f
        # what the wave-front relies on: neighbours precede the macro block in coding order
        for i in np.nonzero(want["valid"])[0]:
            assert all(want["cn"][i][k] < i for k in range(want["ncn"][i]))
    finally:
        R.refh_me_close(h)
