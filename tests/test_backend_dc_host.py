"""The back-end's restated DC un-prediction (ocg_host_dc_unpredict_mcu_plane, table-driven) against the
reference's oc_dec_dc_unpredict_mcu_plane_c (decode.c:1392) on random planes: both run inside the
integrated library on the same fake decoder context (ocg_host_dc_selftest); every DC value and both
fragment counts must agree, MCU by MCU."""
import ctypes as C

import pytest

import th_streams as streams

pytestmark = pytest.mark.skipif(not streams.available(), reason="needs the integrated build")


@pytest.mark.parametrize("dims", [(240, 136, 8), (120, 68, 4), (7, 5, 4), (1, 9, 4), (33, 1, 4), (480, 270, 8), (2, 2, 1)])
def test_restated_dc_unprediction_equals_reference(dims):
    L = streams.lib()
    L.ocg_host_dc_selftest.restype = C.c_long
    L.ocg_host_dc_selftest.argtypes = [C.c_int] * 3 + [C.c_uint, C.c_int, C.c_int]
    nh, nv, mcu = dims
    for pct in (0, 5, 30, 70, 100):
        for mixed in (0, 1):
            for seed in range(3):
                assert L.ocg_host_dc_selftest(nh, nv, mcu, seed, pct, mixed) == 0, (pct, mixed, seed)
