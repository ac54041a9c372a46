"""Diagnostic: first packet at which the device-backed encoder's output differs from the reference's.
python tools/diag_enc.py NFRAMES SPEED [W H]"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import support as S, th_streams as streams

def packets(lib, nf, speed, w, h):
    lib.refh_stream_packet_data.restype = C.c_void_p
    lib.refh_stream_packet_data.argtypes = [C.c_void_p, C.c_int]
    lib.refh_stream_packet_size.restype = C.c_long
    lib.refh_stream_packet_size.argtypes = [C.c_void_p, C.c_int]
    lib.refh_stream_npackets.argtypes = [C.c_void_p]
    sh = lib.refh_encode_synth(w, h, 0, nf, 32, 64, speed, 30, 12345)
    n = lib.refh_stream_npackets(sh)
    out = []
    for i in range(n):
        sz = lib.refh_stream_packet_size(sh, i)
        out.append(C.string_at(lib.refh_stream_packet_data(sh, i), sz))
    return out

nf = int(sys.argv[1]); speed = int(sys.argv[2])
w, h = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
R = S.ref("asm" if S.ref_available("asm") else "c")
G = streams.lib()
a = packets(R, nf, speed, w, h); b = packets(G, nf, speed, w, h)
print("packets", len(a), len(b))
for i, (x, y) in enumerate(zip(a, b)):
    if x != y:
        print("first difference at packet", i, "(frame %d)" % (i - 3), len(x), len(y))
        k = next((j for j in range(min(len(x), len(y))) if x[j] != y[j]), None)
        print("  first differing byte", k)
        break
else:
    print("identical")
