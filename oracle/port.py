"""TEST INFRASTRUCTURE -- ctypes binding of the CPU oracle (oracle/port/*.c -> liboracle.so).

The oracle is a checker: only tests/, smoke() and bench.py's cpu_baseline leg may import it.
Tables come from the same setup blob the product loads, parsed in Python.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "liboracle.so")
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


class OrcPsy(C.Structure):
    _fields_ = [(k, C.c_int) for k in "n blockflag firstoc shiftoc eighth total_octave_lines rate".split()] + \
        [("m3n", C.c_int * 3)] + \
        [(k, C.c_int) for k in "tonecomp_endp min_nn_lp tonefix_end n25p n33p n75p noisewindowfixed normal_p "
                               "normal_start normal_partition partition".split()] + \
        [(k, C.c_float) for k in "m_val tonecomp_thres ath_adjatt ath_maxatt".split()] + \
        [("tone_masteratt", C.c_float * 3)] + \
        [(k, C.c_float) for k in "tone_abs_limit max_curve_dB noisemaxsupp".split()] + \
        [("normal_thresh", C.c_double)] + \
        [("ath", fp), ("octave", ip), ("bark", ip), ("tonecurves", fp), ("noiseoffset", fp), ("ntfix", fp),
         ("compand", fp), ("compand_hi", fp), ("stn_compand", ip), ("freq_bfn128", ip), ("freq_bfn256", ip)]


class OrcFloor(C.Structure):
    _fields_ = [(k, C.c_int) for k in "posts n quant_q mult".split()] + \
        [(k, C.c_float) for k in "maxover maxunder maxerr twofitweight twofitatten".split()] + \
        [(k, ip) for k in "postlist sorted forward reverse lonb hinb".split()]


class OrcBook(C.Structure):
    _fields_ = [(k, C.c_int) for k in "dim entries quantvals minval delta".split()] + [("len", C.POINTER(C.c_uint8))]


class OrcResidue(C.Structure):
    _fields_ = [(k, C.c_int) for k in "type begin end grouping partitions stages phrase_dim groupbook".split()] + \
        [(k, ip) for k in "secondstages classmetric1 classmetric2 partbooks".split()] + [("books", C.POINTER(OrcBook))]


class OrcCouple(C.Structure):
    _fields_ = [("coupling_steps", C.c_int), ("mag", ip), ("ang", ip), ("pointlimit", ip), ("prepointamp", ip),
                ("postpointamp", ip), ("stereo_th", C.POINTER(C.c_double)), ("stereo_thx", C.POINTER(C.c_double)),
                ("fromdB", fp)]


def _fptr(a):
    return a.ctypes.data_as(fp)


def _iptr(a):
    return a.ctypes.data_as(ip)


class Oracle:
    """The CPU oracle bound to one setup (tables = vorbis_aotuv_lancer_b200.blob.parse(raw))."""

    def __init__(self, tables):
        if not os.path.exists(LIBPATH):
            build()
        self.L = L = C.CDLL(LIBPATH)
        self.t = t = {k: np.ascontiguousarray(v) for k, v in tables.items()}
        L.orc_todB.restype = C.c_float
        L.orc_postnoise.restype = C.c_float
        L.orc_loudnoise_fix.restype = C.c_float
        cfg = t["cfg.i32"]
        self.channels, self.rate, self.bs = int(cfg[0]), int(cfg[1]), (int(cfg[2]), int(cfg[3]))
        self.modes = t["cfg.modes"].reshape(-1, 2)
        self.psy = []
        for k in range(int(cfg[8])):
            v, f = t["psy%d.i32" % k], t["psy%d.f32" % k]
            p = OrcPsy()
            (p.n, p.blockflag, p.firstoc, p.shiftoc, p.eighth, p.total_octave_lines, p.rate) = [int(x) for x in v[:7]]
            p.m3n[:] = [int(x) for x in v[7:10]]
            (p.tonecomp_endp, p.min_nn_lp, p.tonefix_end, p.n25p, p.n33p, p.n75p, p.noisewindowfixed, p.normal_p,
             p.normal_start, p.normal_partition) = [int(x) for x in v[10:20]]
            p.partition = p.normal_partition if p.normal_p else 16
            p.m_val, p.tonecomp_thres, p.ath_adjatt, p.ath_maxatt = [float(x) for x in f[:4]]
            p.tone_masteratt[:] = [float(x) for x in f[4:7]]
            p.tone_abs_limit, p.max_curve_dB, p.noisemaxsupp = [float(x) for x in f[7:10]]
            p.normal_thresh = float(t["psy%d.f64" % k][0])
            p.ath, p.octave, p.bark = _fptr(t["psy%d.ath" % k]), _iptr(t["psy%d.octave" % k]), _iptr(t["psy%d.bark" % k])
            p.tonecurves, p.noiseoffset = _fptr(t["psy%d.tonecurves" % k]), _fptr(t["psy%d.noiseoffset" % k])
            p.ntfix, p.compand, p.compand_hi = _fptr(t["psy%d.ntfix" % k]), _fptr(t["psy%d.compand" % k]), _fptr(t["psy%d.compand_hi" % k])
            p.stn_compand, p.freq_bfn128, p.freq_bfn256 = _iptr(t["aotuv.stn_compand"]), _iptr(t["aotuv.freq_bfn128"]), _iptr(t["aotuv.freq_bfn256"])
            self.psy.append(p)
        self.floor = []
        for k in range(int(cfg[9])):
            v = t["flr%d.i32" % k]
            fl = OrcFloor()
            if v[0] == 1:
                f = t["flr%d.f32" % k]
                fl.posts, fl.n, fl.quant_q, fl.mult = int(v[1]), int(v[2]), int(v[3]), int(v[4])
                fl.maxover, fl.maxunder, fl.maxerr, fl.twofitweight, fl.twofitatten = [float(x) for x in f[:5]]
                for nm, key in (("postlist", "postlist"), ("sorted", "sorted"), ("forward", "forward"),
                                ("reverse", "reverse"), ("lonb", "lonb"), ("hinb", "hinb")):
                    setattr(fl, nm, _iptr(t["flr%d.%s" % (k, key)]))
            self.floor.append(fl)
        nb = int(cfg[11])
        self.books = (OrcBook * max(nb, 1))()
        for k in range(nb):
            v = t["book%d.i32" % k]
            b = self.books[k]
            b.dim, b.entries, b.quantvals, b.minval, b.delta = [int(x) for x in v[:5]]
            b.len = t["book%d.len" % k].ctypes.data_as(C.POINTER(C.c_uint8))
        self.residue = []
        for k in range(int(cfg[10])):
            v = t["res%d.i32" % k]
            r = OrcResidue()
            r.type, r.begin, r.end, r.grouping, r.partitions = [int(x) for x in v[:5]]
            r.groupbook, r.stages, r.phrase_dim = int(v[6]), int(v[7]), int(v[8])
            r.secondstages, r.classmetric1 = _iptr(t["res%d.secondstages" % k]), _iptr(t["res%d.classmetric1" % k])
            r.classmetric2, r.partbooks = _iptr(t["res%d.classmetric2" % k]), _iptr(t["res%d.partbooks" % k])
            r.books = C.cast(self.books, C.POINTER(OrcBook))
            self.residue.append(r)
        self.couple_cfg = []
        for k in range(int(cfg[12])):
            c = OrcCouple()
            c.coupling_steps = int(t["map%d.i32" % k][1])
            c.mag, c.ang = _iptr(t["map%d.coupling_mag" % k]), _iptr(t["map%d.coupling_ang" % k])
            c.pointlimit, c.prepointamp, c.postpointamp = _iptr(t["psyg.pointlimit"]), _iptr(t["psyg.prepointamp"]), _iptr(t["psyg.postpointamp"])
            c.stereo_th = t["aotuv.stereo_th"].ctypes.data_as(C.POINTER(C.c_double))
            c.stereo_thx = t["aotuv.stereo_thx"].ctypes.data_as(C.POINTER(C.c_double))
            c.fromdB = _fptr(t["fromdB"])
            self.couple_cfg.append(c)

    # ------------------------------------------------------------ front end
    def window(self, pcm, W, lW, nW):
        d = np.array(pcm, np.float32)
        lW, nW = (lW, nW) if W else (0, 0)
        self.L.orc_window(_fptr(d), len(d), self.bs[lW], self.bs[nW], _fptr(self.t["win%d" % lW]), _fptr(self.t["win%d" % nW]))
        return d

    def mdct_forward(self, windowed):
        w = np.ascontiguousarray(windowed, np.float32)
        out = np.empty(len(w) // 2, np.float32)
        self.L.orc_mdct_forward(_fptr(w), _fptr(out), len(w))
        return out

    def mdct_forward_exact(self, windowed):
        """the reference's own float operations in stage-parallel form: bit-exact against mdct_forward"""
        w = np.ascontiguousarray(windowed, np.float32)
        out = np.empty(len(w) // 2, np.float32)
        self.L.orc_mdct_forward_exact(_fptr(w), _fptr(out), len(w))
        return out

    def logfft_exact(self, windowed):
        """spectrum on the operation-by-operation restatement of drft_forward: bit-exact against the reference"""
        w = np.ascontiguousarray(windowed, np.float32)
        out = np.empty(len(w) // 2, np.float32)
        amp = C.c_float()
        if self.L.orc_logfft_exact(_fptr(w), len(w), _fptr(out), C.byref(amp)):
            raise ValueError("block size with factors other than 2 and 4")
        return out, amp.value

    def logfft(self, windowed):
        w = np.ascontiguousarray(windowed, np.float32)
        out = np.empty(len(w) // 2, np.float32)
        amp = C.c_float()
        self.L.orc_logfft(_fptr(w), len(w), _fptr(out), C.byref(amp))
        return out, amp.value

    def logmdct(self, mdct):
        m = np.ascontiguousarray(mdct, np.float32)
        out = np.empty_like(m)
        self.L.orc_logmdct(_fptr(m), len(m), _fptr(out))
        return out

    def postnoise(self, pcm, mode, lw_mode):
        p = np.ascontiguousarray(pcm, np.float32)
        return self.L.orc_postnoise(_fptr(p), len(p), mode, lw_mode)

    def loudnoise_fix(self, psy, level, logmdct, block_mode, lW_block_mode):
        m = np.ascontiguousarray(logmdct, np.float32)
        return self.L.orc_loudnoise_fix(C.byref(self.psy[psy]), C.c_float(level), _fptr(m), block_mode, lW_block_mode)

    # ------------------------------------------------------------ psy
    def noisemask(self, psy, level, logmdct, lastmdct, poste, block_mode):
        p = self.psy[psy]
        lm, la = np.ascontiguousarray(logmdct, np.float32), np.ascontiguousarray(lastmdct, np.float32)
        ep, no, npk = np.empty(p.n, np.float32), np.empty(p.n, np.float32), np.empty(p.n // p.partition, np.float32)
        self.L.orc_noisemask(C.byref(p), C.c_float(level), _fptr(lm), _fptr(la), _fptr(ep), _fptr(npk), _fptr(no),
                             C.c_float(poste), block_mode)
        return no, ep, npk

    def tonemask(self, psy, logfft, global_specmax, local_specmax):
        p = self.psy[psy]
        lf = np.ascontiguousarray(logfft, np.float32)
        tone = np.empty(p.n, np.float32)
        self.L.orc_tonemask(C.byref(p), _fptr(lf), _fptr(tone), C.c_float(global_specmax), C.c_float(local_specmax))
        return tone

    def offset_and_mix(self, psy, noise, tone, sel, managed, mdct, logmdct, lastmdct, tempmdct, low_compand, npeak,
                       end_block, block_mode, nW, lW_block_mode, lW_no, impadnum):
        p = self.psy[psy]
        noise, tone = np.ascontiguousarray(noise, np.float32), np.ascontiguousarray(tone, np.float32)
        mdct, logmdct = np.array(mdct, np.float32), np.ascontiguousarray(logmdct, np.float32)
        last, temp, npk = np.array(lastmdct, np.float32), np.array(tempmdct, np.float32), np.array(npeak, np.float32)
        mask = np.empty(p.n, np.float32)
        self.L.orc_offset_and_mix(C.byref(p), _fptr(noise), _fptr(tone), sel, managed, _fptr(mask), _fptr(mdct),
                                  _fptr(logmdct), _fptr(last), _fptr(temp), C.c_float(low_compand), _fptr(npk),
                                  end_block, block_mode, nW, lW_block_mode, lW_no, impadnum)
        return mask, mdct, last, temp, npk

    # ------------------------------------------------------------ floor
    def floor1_fit(self, floor, logmdct, logmask):
        f = self.floor[floor]
        posts = np.zeros(65, np.int32)
        ok = self.L.orc_floor1_fit(C.byref(f), _fptr(np.ascontiguousarray(logmdct, np.float32)),
                                   _fptr(np.ascontiguousarray(logmask, np.float32)), _iptr(posts))
        return posts[:f.posts].copy() if ok else None

    def floor1_encode(self, floor, posts, n2):
        f = self.floor[floor]
        vals, ilog = np.zeros(65, np.int32), np.zeros(n2, np.int32)
        post = np.zeros(65, np.int32)
        if posts is not None:
            post[:f.posts] = posts
        nz = self.L.orc_floor1_encode(C.byref(f), _iptr(post), int(posts is not None), _iptr(vals), _iptr(ilog), n2)
        return nz, post[:f.posts].copy(), vals[:f.posts].copy(), ilog

    # ------------------------------------------------------------ coupling / residue
    @staticmethod
    def _rows(arr, ctype):
        ptrs = (C.POINTER(ctype) * len(arr))()
        for i, a in enumerate(arr):
            ptrs[i] = a.ctypes.data_as(C.POINTER(ctype))
        return ptrs

    def couple(self, psy, mapping, blobno, mdct, epeak, npeak, iwork, nonzero, sliding_lowpass, lowpassr):
        ch = len(mdct)
        mdct = [np.ascontiguousarray(x, np.float32) for x in mdct]
        epeak = [np.ascontiguousarray(x, np.float32) for x in epeak]
        npeak = [np.array(x, np.float32) for x in npeak]
        iwork = [np.array(x, np.int32) for x in iwork]
        nz = np.array(nonzero, np.int32)
        self.L.orc_couple(C.byref(self.psy[psy]), C.byref(self.couple_cfg[mapping]), blobno, self._rows(mdct, C.c_float),
                          self._rows(epeak, C.c_float), self._rows(npeak, C.c_float), self._rows(iwork, C.c_int),
                          _iptr(nz), sliding_lowpass, ch, lowpassr)
        return np.stack(iwork), nz, np.stack(npeak)

    def residue_class(self, resnum, rows, nonzero):
        r = self.residue[resnum]
        ch = len(rows)
        rows = [np.ascontiguousarray(x, np.int32) for x in rows]
        partvals = (r.end - r.begin) // r.grouping
        pw = np.zeros((ch, partvals), np.int32)
        nz = np.ascontiguousarray(nonzero, np.int32)
        n = self.L.orc_residue_class(C.byref(r), self._rows(rows, C.c_int), _iptr(nz), ch, _iptr(pw))
        return pw[:n].copy() if n else None

    def residue_forward(self, resnum, rows, nonzero, partword, max_syms=20000):
        r = self.residue[resnum]
        ch = len(rows)
        rows = [np.array(x, np.int32) for x in rows]
        nz = np.ascontiguousarray(nonzero, np.int32)
        pw = np.ascontiguousarray(partword, np.int32)
        syms = np.zeros((max_syms, 2), np.int32)
        n = self.L.orc_residue_forward(C.byref(r), self._rows(rows, C.c_int), _iptr(nz), ch, _iptr(pw), len(rows[0]),
                                       _iptr(syms), max_syms)
        return syms[:n].copy(), np.stack(rows)
