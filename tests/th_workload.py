"""TEST / BENCH PLUMBING (not part of the product package).  Synthetic benchmark workload: a deterministic
1080p/2160p 4:2:0 Theora stream produced by the reference ENCODER host code in
the integrated library (C kernels; input synthesis, never timed), GOP segments
encoded in parallel and cached under /tmp.  Used by bench.py and the tests."""
import concurrent.futures as cf
import ctypes as C
import hashlib
import os


CACHE_DIR = os.environ.get("THEORA_B200_CACHE", "/tmp/theora_b200_cache")


def _key(*a):
    return hashlib.sha1(repr(a).encode()).hexdigest()[:16]


def synth_stream(w=1920, h=1080, nframes=300, quality=32, kf=64, speed=1, noise_shift=30, seed=12345,
                 threads=None, lib=None, use_cache=True):
    """Returns the serialised stream (3 headers + nframes data packets) as bytes.

    A keyframe is forced every `kf` frames, so each GOP is encoded by an
    independent encoder instance on its own thread and the data packets are
    concatenated behind the first segment's headers (all segments share the
    same headers).  The result is a valid stream with ceil(nframes/kf) intra
    frames; it is not byte-identical to a serial encode (rate-control history
    differs), which does not matter for a decode workload."""
    if lib is None:
        import th_streams as streams
        L = streams.lib()
    else:
        L = lib
    path = os.path.join(CACHE_DIR, "synth_%s.ogs" % _key(w, h, nframes, quality, kf, speed, noise_shift, seed))
    if use_cache and os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    segs = [(f0, min(kf, nframes - f0)) for f0 in range(0, nframes, kf)]
    threads = threads or min(len(segs), os.cpu_count() or 1)

    def enc(seg):
        hnd = L.refh_encode_synth(w, h, seg[0], seg[1], quality, kf, speed, noise_shift, seed)
        assert hnd, "encoder failed"
        return hnd

    # stream synthesis is tooling: keep the reference's host encoder even for
    # intra-only streams (the device encoder is what tests/bench measure)
    integrated = lib is None
    if integrated:
        L.ocg_backend_set_enc_mode(1)  # OCG_ENC_HOST
    try:
        with cf.ThreadPoolExecutor(max_workers=threads) as ex:
            handles = list(ex.map(enc, segs))
    finally:
        if integrated:
            L.ocg_backend_set_enc_mode(0)  # OCG_ENC_AUTO
    first = handles[0]
    for hnd in handles[1:]:
        L.refh_stream_append_data(first, hnd)
        L.refh_stream_free(hnd)
    n = L.refh_stream_blob_size(first)
    buf = (C.c_uint8 * n)()
    assert L.refh_stream_to_blob(first, buf, n) == n
    L.refh_stream_free(first)
    blob = bytes(buf)
    if use_cache:
        os.makedirs(CACHE_DIR, exist_ok=True)
        tmp = path + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(blob)
        os.replace(tmp, path)
    return blob


def algorithmic_bytes(work):
    """SURVEY.md 8(d) per-unit figures: recon intra 192 B, inter 256 B, uncoded
    copy 128 B; loop filter 128 B per fragment of plane area; returns
    (recon_stage_bytes, loop_filter_bytes)."""
    if work is None:
        return 0, 0
    coded = work.coded_mask
    intra = int((work.recs["refi"] == 2).sum())
    ncoded = int(coded.sum())
    inter = ncoded - intra
    recon = intra * 192 + inter * 256 + (len(work.recs) - ncoded) * 128
    lf = len(work.recs) * 128 if work.lf_limit else 0
    return recon, lf
