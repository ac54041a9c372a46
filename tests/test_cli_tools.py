"""The Ogg / YUV4MPEG2 command-line tools (tools/cli: the jobs of the reference's
examples/encoder_example.c and examples/dump_video.c, libogg replaced by
ogg_lite.c).  CPU part = BASELINE configs[0]: a 64x64 2-frame synthetic .ogv
through the reference's C path.  The container writer is cross-checked by an
independent page parser written here (RFC 3533: capture pattern, CRC-32 with
generator 0x04c11db7, lacing) and, where OpenCV's bundled FFmpeg is present, by
FFmpeg's own Ogg demuxer + Theora decoder; the GPU part checks that the tools
linked against the B200 back-end write byte-identical files."""
import os
import struct
import subprocess

import numpy as np
import pytest

import support as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "cli", "bin")


def have(*names):
    return all(os.path.exists(os.path.join(BIN, n)) for n in names)


def write_y4m(path, w, h, n, seed=1, chroma="420jpeg"):
    rng = np.random.default_rng(seed)
    hd, vd = (1, 1) if chroma.startswith("420") else ((1, 0) if chroma.startswith("422") else (0, 0))
    cw, ch = (w + hd) >> hd, (h + vd) >> vd
    yy, xx = np.mgrid[0:h, 0:w]
    frames = []
    with open(path, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C%s\n" % (w, h, chroma)).encode())
        for t in range(n):
            y = ((2 * (xx + 3 * t) + (yy + t)) & 255) // 2 + 60 * ((((xx + 3 * t) >> 4) ^ ((yy + t) >> 4)) & 1)
            y = np.clip(y + rng.integers(0, 4, size=y.shape), 0, 255).astype(np.uint8)
            cb = np.full((ch, cw), 128, np.uint8) + (np.arange(cw, dtype=np.uint8)[None, :] >> 3 & 15)
            cr = np.full((ch, cw), 128, np.uint8) - (np.arange(ch, dtype=np.uint8)[:, None] >> 3 & 15)
            f.write(b"FRAME\n" + y.tobytes() + cb.tobytes() + cr.tobytes())
            frames.append((y, cb, cr))
    return frames


def read_y4m(path):
    with open(path, "rb") as f:
        hdr = f.readline().decode().split()
        assert hdr[0] == "YUV4MPEG2"
        kv = {t[0]: t[1:] for t in hdr[1:]}
        w, h, c = int(kv["W"]), int(kv["H"]), kv.get("C", "420jpeg")
        hd, vd = (1, 1) if c.startswith("420") else ((1, 0) if c.startswith("422") else (0, 0))
        cw, ch = (w + hd) >> hd, (h + vd) >> vd
        frames = []
        while True:
            line = f.readline()
            if not line:
                break
            assert line == b"FRAME\n"
            buf = f.read(w * h + 2 * cw * ch)
            assert len(buf) == w * h + 2 * cw * ch
            frames.append(np.frombuffer(buf, np.uint8))
    return w, h, frames


def ogg_crc(data):
    tab = []
    for i in range(256):
        r = i << 24
        for _ in range(8):
            r = ((r << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if r & 0x80000000 else (r << 1) & 0xFFFFFFFF
        tab.append(r)
    crc = 0
    for b in data:
        crc = ((crc << 8) & 0xFFFFFFFF) ^ tab[((crc >> 24) & 0xFF) ^ b]
    return crc


def parse_ogg(blob):
    """Independent RFC 3533 page walk: returns (packets, pages) and checks every page."""
    pos, packets, cur, pages, seq = 0, [], b"", [], 0
    serial = None
    while pos < len(blob):
        assert blob[pos:pos + 4] == b"OggS", "lost sync at %d" % pos
        ver, flags, gp, ser, pageno, crc, nsegs = struct.unpack_from("<BBqIIIB", blob, pos + 4)
        assert ver == 0
        lac = blob[pos + 27:pos + 27 + nsegs]
        blen = sum(lac)
        page = bytearray(blob[pos:pos + 27 + nsegs + blen])
        page[22:26] = b"\0\0\0\0"
        assert ogg_crc(bytes(page)) == crc, "bad CRC on page %d" % pageno
        serial = ser if serial is None else serial
        assert ser == serial and pageno == seq
        assert bool(flags & 2) == (seq == 0), "BOS flag only on the first page"
        assert bool(flags & 1) == (len(cur) > 0), "continued flag must match an open packet"
        seq += 1
        body = blob[pos + 27 + nsegs:pos + 27 + nsegs + blen]
        o, done_here = 0, 0
        for l in lac:
            cur += body[o:o + l]
            o += l
            if l < 255:
                packets.append(cur)
                cur = b""
                done_here += 1
        pages.append({"flags": flags, "granulepos": gp, "packets_ended": done_here})
        assert (gp == -1) == (done_here == 0)
        pos += 27 + nsegs + blen
    assert cur == b"" and pages[-1]["flags"] & 4, "stream must end with a complete packet on an EOS page"
    return packets, pages


def run(*cmd):
    p = subprocess.run(list(cmd), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, "%s\n%s" % (" ".join(cmd), p.stderr)
    return p.stderr


CASES = [(64, 64, 2, "420jpeg", 8.0, 64), (100, 70, 5, "420jpeg", 5.0, 3), (352, 288, 6, "420jpeg", 6.0, 64),
         (96, 80, 4, "422jpeg", 7.0, 64), (80, 48, 4, "444", 4.0, 2)]


@pytest.mark.skipif(not have("ref_encoder_example", "ref_dump_video"), reason="tools/cli not built (needs the reference)")
@pytest.mark.parametrize("case", CASES)
def test_reference_path_roundtrip_and_container(case, tmp_path):
    w, h, n, chroma, q, kf = case
    y4m, ogv, out = str(tmp_path / "in.y4m"), str(tmp_path / "a.ogv"), str(tmp_path / "out.y4m")
    src = write_y4m(y4m, w, h, n, chroma=chroma)
    run(os.path.join(BIN, "ref_encoder_example"), "-o", ogv, "-v", str(q), "-k", str(kf), y4m)
    blob = open(ogv, "rb").read()
    packets, pages = parse_ogg(blob)
    assert len(packets) == 3 + n and packets[0][:7] == b"\x80theora" and packets[1][:7] == b"\x81theora"
    assert pages[0]["packets_ended"] == 1, "the identification header sits alone on the first page"
    run(os.path.join(BIN, "ref_dump_video"), "-c", "-o", out, ogv)
    ow, oh, frames = read_y4m(out)
    assert (ow, oh, len(frames)) == (w, h, n)
    # lossy, but it must be THIS video: luma PSNR against the source
    for t in range(n):
        y = frames[t][:w * h].reshape(h, w).astype(np.float64)
        mse = np.mean((y - src[t][0].astype(np.float64)) ** 2)
        assert 10 * np.log10(255 ** 2 / max(mse, 1e-9)) > 26, "frame %d does not resemble the input" % t
    # the same packets through the library harness (no container) give the same pictures
    R = S.ref("c")
    import ctypes as C
    hb = np.frombuffer(S.Stream.encode(R, 64, 64, 1).to_bytes()[:16], np.uint32)  # harness blob header layout
    blob2 = np.array([hb[0], len(packets), hb[2], hb[3]], np.uint32).tobytes() + \
        np.array([len(p) for p in packets], np.uint32).tobytes() + b"".join(packets)
    st = S.Stream.from_bytes(R, blob2)
    dec = S.Decoder(R, st)
    hd, vd = (1, 1) if chroma.startswith("420") else ((1, 0) if chroma.startswith("422") else (0, 0))
    for t in range(n):
        assert dec.next() >= 0
        full = dec.frame()
        fw, fh = dec.fw, dec.fh
        yp = full[:fw * fh].reshape(fh, fw)[dec.py:dec.py + h, dec.px:dec.px + w]
        assert np.array_equal(yp.ravel(), frames[t][:w * h]), "frame %d: tool output differs from the library's" % t
    dec.close()
    st.free()


@pytest.mark.skipif(not have("ref_encoder_example"), reason="tools/cli not built (needs the reference)")
def test_ffmpeg_reads_our_ogg_files(tmp_path):
    """Third opinion on the muxer: OpenCV's bundled FFmpeg demuxes the pages (it verifies page CRCs) and its
    own Theora decoder reproduces the pictures (BGR output, so a resemblance check, not a bit-exact one)."""
    cv2 = pytest.importorskip("cv2")
    w, h, n = 176, 144, 8
    y4m, ogv, out = str(tmp_path / "in.y4m"), str(tmp_path / "a.ogv"), str(tmp_path / "o.y4m")
    write_y4m(y4m, w, h, n)
    run(os.path.join(BIN, "ref_encoder_example"), "-o", ogv, "-v", "8", "-k", "4", y4m)
    run(os.path.join(BIN, "ref_dump_video"), "-c", "-o", out, ogv)
    _, _, ours = read_y4m(out)
    cap = cv2.VideoCapture(ogv)
    if not cap.isOpened():
        pytest.skip("this OpenCV build cannot open Ogg/Theora")
    got = 0
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        assert bgr.shape[:2] == (h, w)
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY).astype(np.float64)
        y = ours[got][:w * h].reshape(h, w).astype(np.float64)
        cand = [np.mean(np.abs(gray - y)), np.mean(np.abs(gray - np.clip((y - 16) * 255 / 219, 0, 255)))]
        assert min(cand) < 6.0, "frame %d decoded by FFmpeg does not match ours (%r)" % (got, cand)
        got += 1
    assert got == n


@pytest.mark.gpu
@pytest.mark.skipif(not have("ref_encoder_example", "ref_dump_video", "ocg_dump_video", "ocg_encoder_example"),
                    reason="tools/cli not built")
@pytest.mark.parametrize("case", CASES + [(1920, 1080, 3, "420jpeg", 5.0, 64)])
def test_gpu_tools_write_identical_files(case, tmp_path):
    """BASELINE configs[0] on the device: dump_video through the B200 back-end == through the reference;
    with post-processing too; and the intra-only device encoder writes the same .ogv bytes."""
    w, h, n, chroma, q, kf = case
    y4m, ogv = str(tmp_path / "in.y4m"), str(tmp_path / "a.ogv")
    write_y4m(y4m, w, h, n, chroma=chroma)
    run(os.path.join(BIN, "ref_encoder_example"), "-o", ogv, "-v", str(q), "-k", str(kf), y4m)
    for extra in ([], ["-c"], ["-p", "6"]):
        a, b, z = str(tmp_path / "ref.y4m"), str(tmp_path / "gpu.y4m"), str(tmp_path / "gpu_z.y4m")
        run(os.path.join(BIN, "ref_dump_video"), *extra, "-o", a, ogv)
        run(os.path.join(BIN, "ocg_dump_video"), *extra, "-o", b, ogv)
        assert open(a, "rb").read() == open(b, "rb").read(), "dump_video output differs (%r)" % (extra,)
        # zero-copy hand-off: rows written straight out of the page-locked buffer the device filled
        run(os.path.join(BIN, "ocg_dump_video"), "-z", *extra, "-o", z, ogv)
        assert open(a, "rb").read() == open(z, "rb").read(), "zero-copy dump_video output differs (%r)" % (extra,)
    if chroma == "420jpeg" and w <= 352:
        a, b = str(tmp_path / "ref1.ogv"), str(tmp_path / "gpu1.ogv")
        run(os.path.join(BIN, "ref_encoder_example"), "-o", a, "-v", str(q), "-k", "1", y4m)
        run(os.path.join(BIN, "ocg_encoder_example"), "-o", b, "-v", str(q), "-k", "1", y4m)
        assert open(a, "rb").read() == open(b, "rb").read(), "intra-only .ogv differs from the reference encoder's"


@pytest.mark.skipif(not have("ref_encoder_example", "ref_dump_video"), reason="tools/cli not built (needs the reference)")
def test_large_packets_span_pages_and_damage_is_survived(tmp_path):
    """A key frame larger than one page's 255 x 255 bytes must continue on the following pages (RFC 3533
    section 5: continued-packet flag, granule position -1 on pages where no packet ends); and the reader
    must drop a page with a bad CRC, resynchronise, and carry on with the packets it can still complete."""
    w, h, n = 640, 480, 4
    y4m, ogv, out = str(tmp_path / "in.y4m"), str(tmp_path / "a.ogv"), str(tmp_path / "out.y4m")
    rng = np.random.default_rng(9)
    with open(y4m, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h)).encode())
        for t in range(n):
            f.write(b"FRAME\n" + rng.integers(0, 256, size=w * h * 3 // 2, dtype=np.uint8).tobytes())  # incompressible
    run(os.path.join(BIN, "ref_encoder_example"), "-o", ogv, "-v", "10", "-k", "2", y4m)
    blob = open(ogv, "rb").read()
    packets, pages = parse_ogg(blob)
    assert len(packets) == 3 + n
    assert max(len(p) for p in packets) > 255 * 255, "the test needs a packet that cannot fit one page"
    assert any(pg["flags"] & 1 for pg in pages) and any(pg["granulepos"] == -1 for pg in pages)
    run(os.path.join(BIN, "ref_dump_video"), "-o", out, ogv)
    _, _, frames = read_y4m(out)
    assert len(frames) == n
    # flip one byte in the middle of the LAST data packet's pages: that frame is lost, the rest decodes
    bad = bytearray(blob)
    bad[len(bad) - 2000] ^= 0x55
    open(ogv, "wb").write(bytes(bad))
    log = run(os.path.join(BIN, "ref_dump_video"), "-o", out, ogv)
    assert "1 bad CRC" in log
    _, _, frames2 = read_y4m(out)
    assert 1 <= len(frames2) < n
    for a, b in zip(frames2, frames):
        assert np.array_equal(a, b)


def _page(serial, pageno, flags, granule, packet):
    """One Ogg page (RFC 3533 section 6) holding one complete packet shorter than 255 bytes."""
    hdr = struct.pack("<4sBBqIIIB", b"OggS", 0, flags, granule, serial, pageno, 0, 1) + bytes([len(packet)])
    crc = ogg_crc(hdr + packet)
    return hdr[:22] + struct.pack("<I", crc) + hdr[26:] + packet


@pytest.mark.skipif(not have("ref_encoder_example", "ref_dump_video"), reason="tools/cli not built (needs the reference)")
def test_reader_follows_the_theora_stream_of_a_multiplexed_file_and_resyncs_inside_a_damaged_page(tmp_path):
    """(1) A multiplexed file begins with one beginning-of-stream page per logical stream; the reader must follow
    the one whose first packet is a Theora identification header even if another stream's comes first (the
    reference's dump_video probes every stream, examples/dump_video.c:330-370).  (2) When a page's SEGMENT TABLE
    is damaged the page length read from it is garbage: the reader restarts the capture-pattern search right
    behind the failed page's first byte instead of skipping what the header claimed (ogg_sync_pageseek does the
    same), so the intact pages that follow are not swallowed."""
    w, h, n = 176, 144, 6
    y4m, ogv, out = str(tmp_path / "in.y4m"), str(tmp_path / "a.ogv"), str(tmp_path / "out.y4m")
    write_y4m(y4m, w, h, n)
    run(os.path.join(BIN, "ref_encoder_example"), "-o", ogv, "-v", "6", "-k", "3", y4m)
    blob = open(ogv, "rb").read()
    run(os.path.join(BIN, "ref_dump_video"), "-o", out, ogv)
    _, _, want = read_y4m(out)
    assert len(want) == n
    # (1) another logical stream's BOS page in front, one of its data pages somewhere behind the headers
    other_bos = _page(0x1234, 0, 2, 0, b"\x01vorbis" + bytes(23))
    other_data = _page(0x1234, 1, 0, 0, bytes(40))
    first_len = 27 + 1 + blob[27]  # the Theora BOS page: one packet, one lacing value
    open(ogv, "wb").write(other_bos + blob[:first_len] + other_data + blob[first_len:])
    run(os.path.join(BIN, "ref_dump_video"), "-o", out, ogv)
    _, _, got = read_y4m(out)
    assert len(got) == n and all(np.array_equal(a, b) for a, b in zip(got, want))
    # (2) damage the segment COUNT of a late page so that it claims far more than it holds
    packets, pages = parse_ogg(blob)
    pos, starts = 0, []
    while pos < len(blob):
        nsegs = blob[pos + 26]
        starts.append(pos)
        pos += 27 + nsegs + sum(blob[pos + 27:pos + 27 + nsegs])
    victim = starts[-3]
    bad = bytearray(blob)
    bad[victim + 26] = 255
    open(ogv, "wb").write(bytes(bad))
    log = run(os.path.join(BIN, "ref_dump_video"), "-o", out, ogv)
    assert "bad CRC" in log
    _, _, got = read_y4m(out)
    # only the frames carried by the damaged page may be missing: the pages behind it must still be read
    assert len(got) >= n - 3 and len(got) < n
