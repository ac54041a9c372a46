"""GPU path against the committed golden vectors: golden Theora packets decoded
through th_decode_packetin with the B200 back-end must hash to what the
unmodified reference produced; encoder kernels must reproduce the reference's
fDCT / quantiser / metric outputs."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import support as S
from theora_b200 import abi
import th_streams as streams

pytestmark = pytest.mark.gpu

U = np.load(os.path.join(S.GOLDEN_DIR, "units.npz"))
G = np.load(os.path.join(S.GOLDEN_DIR, "streams.npz"))
STREAM_NAMES = sorted(k[:-5] for k in G.files if k.endswith("_blob"))


@pytest.mark.skipif(not streams.available(), reason="integrated build not present")
@pytest.mark.parametrize("name", STREAM_NAMES)
def test_golden_stream_decodes_bit_exact_on_gpu(name):
    blob = G[name + "_blob"].tobytes()
    want = G[name + "_hashes"]
    L = streams.lib()
    L.ocg_backend_set_mode(streams.BACKEND_GPU)
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    sh = L.refh_stream_from_blob(buf, len(blob))
    d = L.refh_dec_open(sh)
    assert d, abi.lib().ocg_last_error()
    for i in range(len(want)):
        assert L.refh_dec_next(d) >= 0
        h = (C.c_uint64 * 3)()
        L.refh_dec_hash(d, h)
        assert [int(x) for x in h] == [int(x) for x in want[i]], (name, i)
    L.refh_dec_close(d)
    L.refh_stream_free(sh)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_encoder_kernels_against_golden():
    L = abi.lib()
    st = torch.cuda.current_stream().cuda_stream
    # metrics: blocks are packed 8x8 (stride 8); build one frame holding them all
    blocks = U["met_blocks"]
    n = len(blocks)
    src, r1, r2 = (_dev(blocks[:, k]) for k in range(3))
    refs = torch.cat([r1.reshape(-1), r2.reshape(-1)])
    fr = np.zeros(n, S.ENC_FRAG_DTYPE)
    fr["src_off"] = np.arange(n) * 64
    fr["ref_off0"] = np.arange(n) * 64
    fr["ref_off1"] = S.INT32_MIN
    fr2 = fr.copy()
    fr2["ref_off1"] = n * 64 + np.arange(n) * 64
    want = U["met_out"]
    cases = [(0, fr, 0, None), (0, fr2, 1, None), (1, fr, 2, 3), (1, fr2, 4, 5), (2, fr, 6, 7), (3, fr, 8, None),
             (4, fr, 9, None)]
    pad = torch.zeros(64, dtype=torch.uint8, device="cuda")
    srcp = torch.cat([src.reshape(-1), pad])
    refp = torch.cat([refs, pad])
    for metric, frs, col, dccol in cases:
        dfr = _dev(frs.view(np.int32).reshape(n, 4))
        ov = torch.zeros(n, dtype=torch.int32, device="cuda")
        odc = torch.zeros(n, dtype=torch.int32, device="cuda")
        abi.check(L.ocg_enc_metrics_batch(metric, srcp.data_ptr(), refp.data_ptr(), 8, dfr.data_ptr(), n,
                                          ov.data_ptr(), odc.data_ptr(), st))
        torch.cuda.synchronize()
        assert np.array_equal(ov.cpu().numpy().astype(np.int64), want[:, col]), metric
        if dccol is not None:
            assert np.array_equal(odc.cpu().numpy().astype(np.int64), want[:, dccol]), metric
    # fDCT + quantise: residual = src - 128 with src chosen so that the residual equals the golden input
    x = U["fdct_x"]
    keep = np.nonzero((x.min(axis=1) >= -128) & (x.max(axis=1) <= 127))[0]
    assert len(keep) >= 4
    m = len(keep)
    srcb = (x[keep].astype(np.int32) + 128).astype(np.uint8)
    dsrc = torch.cat([_dev(srcb).reshape(-1), pad])
    fr = np.zeros(m, S.ENC_FRAG_DTYPE)
    fr["src_off"] = np.arange(m) * 64
    fr["ref_off0"] = S.INT32_MIN
    fr["ref_off1"] = S.INT32_MIN
    for j, i in enumerate(keep.tolist()):
        deq = np.tile(U["q_deq"][i], (18, 1))
        enq = np.tile(U["q_enq"][i], (18, 1))
        od = torch.zeros((1, 64), dtype=torch.int16, device="cuda")
        oq = torch.zeros((1, 64), dtype=torch.int16, device="cuda")
        onz = torch.zeros(1, dtype=torch.int32, device="cuda")
        one = _dev(fr[j:j + 1].view(np.int32).reshape(1, 4))
        ddeq, denq = _dev(deq.view(np.int16)), _dev(enq)  # keep alive across the launch
        abi.check(L.ocg_enc_fdct_quant_batch(dsrc.data_ptr(), dsrc.data_ptr(), 8, one.data_ptr(), 1,
                                             ddeq.data_ptr(), denq.data_ptr(),
                                             od.data_ptr(), oq.data_ptr(), onz.data_ptr(), st))
        torch.cuda.synchronize()
        assert np.array_equal(od.cpu().numpy()[0], U["fdct_y"][i])
        assert np.array_equal(oq.cpu().numpy()[0], U["q_out"][i])
        assert int(onz.item()) == int(U["q_last"][i])
