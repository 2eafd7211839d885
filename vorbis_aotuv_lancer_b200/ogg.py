"""Ogg pages around the packets of one stream (csrc/oggframe.c): what the reference's outer loop
(examples/encoder_example.c:141-176,224-235) gets from libogg."""
import ctypes as C

from .lib import L, check

L.vb200_oggmux_data.restype = C.POINTER(C.c_ubyte)
L.vb200_oggmux_pages.restype = C.c_long
L.vb200_ogg_crc.restype = C.c_uint32
L.vb200_ogg_crc.argtypes = [C.c_char_p, C.c_size_t]


def crc(data):
    return int(L.vb200_ogg_crc(bytes(data), len(data)))


def mux(headers, packets, serial=0x42323030):
    """headers: the three Vorbis header packets; packets: [(bytes, granulepos, eos)] -> the .ogg byte stream."""
    h = C.c_void_p()
    check(L.vb200_oggmux_create(C.c_int32(serial), C.byref(h)), "vb200_oggmux_create")
    try:
        for hp in headers:
            check(L.vb200_oggmux_packet(h, bytes(hp), C.c_long(len(hp)), C.c_int64(0), 0), "vb200_oggmux_packet")
        check(L.vb200_oggmux_flush(h), "vb200_oggmux_flush")         # audio starts on a fresh page
        for data, gp, eos in packets:
            check(L.vb200_oggmux_packet(h, bytes(data), C.c_long(len(data)), C.c_int64(gp), int(eos)), "vb200_oggmux_packet")
        n = C.c_size_t()
        p = L.vb200_oggmux_data(h, C.byref(n))
        return C.string_at(p, n.value)
    finally:
        L.vb200_oggmux_destroy(h)
