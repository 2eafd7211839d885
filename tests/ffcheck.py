"""Test helper: an independent, third-party reader of what the product writes.

FFmpeg's libavformat / libavcodec travel with this image inside the opencv wheel (libavformat 62 = FFmpeg 8.0).
Its Ogg demultiplexer is its own implementation (libavformat/oggdec.c, not libogg) and verifies every page's CRC;
its Vorbis decoder (libavcodec/vorbisdec.c) is a from-scratch implementation of the Vorbis I specification, not
libvorbis.  Neither shares code with the reference or with this repository, so a stream they accept and decode to
the reference decoder's PCM is evidence beyond "our own demultiplexer reads our own pages".  Test infrastructure
only; nothing in the package imports it.  Struct offsets used below are those of FFmpeg 6-8 (x86-64) and every one of
them is validated against a value known in advance before it is trusted (codec type/id, sample format)."""
import ctypes as C
import glob
import os
import tempfile

import numpy as np

_LIBS = None
AV_CODEC_ID_VORBIS = 0x15005
AVMEDIA_TYPE_AUDIO = 1
AV_SAMPLE_FMT_FLTP = 8
AVERROR_EOF = -0x20464F45          # -MKTAG('E','O','F',' ')
AVERROR_EAGAIN = -11


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    found = None
    try:
        import importlib.util
        spec = importlib.util.find_spec("cv2")
        base = os.path.dirname(os.path.dirname(spec.origin)) if spec and spec.origin else None
        for name in ("opencv_python_headless.libs", "opencv_python.libs", "opencv_contrib_python_headless.libs"):
            d = os.path.join(base, name) if base else None
            if d and os.path.isdir(d):
                found = d
                break
    except Exception:
        found = None
    if not found:
        _LIBS = False
        return _LIBS
    libs = {}
    try:
        for n in ("libdrm", "libcrypto", "libssl", "libpng16", "libaom", "libvpx", "libavutil", "libswresample",
                  "libavcodec", "libavformat"):
            hits = sorted(glob.glob(os.path.join(found, n + "-*.so*")))
            if hits:
                libs[n] = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
        av, ac, af = libs["libavutil"], libs["libavcodec"], libs["libavformat"]
    except (OSError, KeyError):
        _LIBS = False
        return _LIBS
    af.avformat_version.restype = C.c_uint
    if not 59 <= (af.avformat_version() >> 16) <= 62:
        _LIBS = False
        return _LIBS
    ac.avcodec_find_decoder.restype = C.c_void_p
    ac.avcodec_find_decoder.argtypes = [C.c_int]
    af.av_find_input_format.restype = C.c_void_p
    if not ac.avcodec_find_decoder(AV_CODEC_ID_VORBIS) or not af.av_find_input_format(b"ogg"):
        _LIBS = False
        return _LIBS
    ac.avcodec_alloc_context3.restype = C.c_void_p
    ac.avcodec_alloc_context3.argtypes = [C.c_void_p]
    ac.avcodec_parameters_to_context.argtypes = [C.c_void_p, C.c_void_p]
    ac.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    ac.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    ac.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    ac.avcodec_free_context.argtypes = [C.c_void_p]
    ac.av_packet_alloc.restype = C.c_void_p
    ac.av_packet_unref.argtypes = [C.c_void_p]
    ac.av_packet_free.argtypes = [C.c_void_p]
    av.av_frame_alloc.restype = C.c_void_p
    av.av_frame_unref.argtypes = [C.c_void_p]
    av.av_frame_free.argtypes = [C.c_void_p]
    av.av_log_set_level.argtypes = [C.c_int]
    af.avformat_open_input.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
    af.avformat_find_stream_info.argtypes = [C.c_void_p, C.c_void_p]
    af.av_read_frame.argtypes = [C.c_void_p, C.c_void_p]
    af.avformat_close_input.argtypes = [C.c_void_p]
    av.av_log_set_level(16)         # AV_LOG_ERROR
    _LIBS = (av, ac, af)
    return _LIBS


def available():
    return bool(_load())


class AVPacketHead(C.Structure):
    # libavcodec/packet.h: the leading fields of AVPacket, unchanged since FFmpeg 4
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p),
                ("size", C.c_int), ("stream_index", C.c_int), ("flags", C.c_int)]


def _u32(addr):
    return C.c_uint32.from_address(addr).value


def _ptr(addr):
    return C.c_void_p.from_address(addr).value


def read_ogg(stream_bytes, channels, decode=True):
    """Demultiplex (and decode) an .ogg byte string with FFmpeg.
    -> dict(packets=[(bytes, pts)], pcm=float32 [channels][n] or None).  Raises on any FFmpeg error, including a
    page whose CRC does not verify (FFmpeg drops such a page and resynchronises: the packet list then differs)."""
    libs = _load()
    assert libs, "FFmpeg libraries not present"
    av, ac, af = libs
    fd, path = tempfile.mkstemp(suffix=".ogg")
    os.write(fd, stream_bytes)
    os.close(fd)
    ctx = C.c_void_p()
    cctx = pkt = frame = None
    try:
        rc = af.avformat_open_input(C.byref(ctx), path.encode(), None, None)
        assert rc == 0, "avformat_open_input: %d" % rc
        rc = af.avformat_find_stream_info(ctx, None)
        assert rc >= 0, "avformat_find_stream_info: %d" % rc
        # AVFormatContext: av_class, iformat, oformat, priv_data, pb (5 pointers), ctx_flags, nb_streams, streams
        assert _u32(ctx.value + 44) == 1, "expected one logical stream"
        st0 = _ptr(_ptr(ctx.value + 48))
        par = None
        for off in (16, 8):          # AVStream: [av_class,] index, id, codecpar
            cand = _ptr(st0 + off)
            if cand and _u32(cand) == AVMEDIA_TYPE_AUDIO and _u32(cand + 4) == AV_CODEC_ID_VORBIS:
                par = cand
                break
        assert par, "AVStream.codecpar not found (or the stream was not recognised as Vorbis audio)"
        if decode:
            dec = ac.avcodec_find_decoder(AV_CODEC_ID_VORBIS)
            cctx = ac.avcodec_alloc_context3(dec)
            assert ac.avcodec_parameters_to_context(cctx, par) >= 0
            rc = ac.avcodec_open2(cctx, dec, None)
            assert rc == 0, "avcodec_open2 (the Vorbis headers were rejected): %d" % rc
            frame = av.av_frame_alloc()
        pkt = ac.av_packet_alloc()
        packets, chunks = [], []

        def drain():
            while True:
                r = ac.avcodec_receive_frame(cctx, frame)
                if r in (AVERROR_EAGAIN, AVERROR_EOF):
                    return
                assert r == 0, "avcodec_receive_frame: %d" % r
                n, fmt = C.c_int.from_address(frame + 112).value, C.c_int.from_address(frame + 116).value
                assert fmt == AV_SAMPLE_FMT_FLTP and 0 < n <= 8192, (fmt, n)
                ext = _ptr(frame + 96)
                chunk = np.empty((channels, n), np.float32)
                for c in range(channels):
                    chunk[c] = np.ctypeslib.as_array((C.c_float * n).from_address(_ptr(ext + 8 * c)))
                chunks.append(chunk)
                av.av_frame_unref(frame)

        while True:
            rc = af.av_read_frame(ctx, pkt)
            if rc == AVERROR_EOF:
                break
            assert rc == 0, "av_read_frame: %d" % rc
            p = AVPacketHead.from_address(pkt)
            packets.append((C.string_at(p.data, p.size), p.pts))
            if decode:
                rc = ac.avcodec_send_packet(cctx, pkt)
                assert rc == 0, "avcodec_send_packet (packet %d rejected): %d" % (len(packets) - 1, rc)
                drain()
            ac.av_packet_unref(pkt)
        if decode:
            ac.avcodec_send_packet(cctx, None)
            drain()
        pcm = np.concatenate(chunks, axis=1) if chunks else (np.zeros((channels, 0), np.float32) if decode else None)
        return dict(packets=packets, pcm=pcm)
    finally:
        if frame:
            f = C.c_void_p(frame)
            av.av_frame_free(C.byref(f))
        if pkt:
            p = C.c_void_p(pkt)
            ac.av_packet_free(C.byref(p))
        if cctx:
            c = C.c_void_p(cctx)
            ac.avcodec_free_context(C.byref(c))
        if ctx:
            af.avformat_close_input(C.byref(ctx))
        os.unlink(path)
