"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ELECWAVEL_DEFAULT = 0.019866  # param.cpp:86


class OracleCfg(C.Structure):
    _fields_ = [("NumberPixels", C.c_int), ("maxDisplaceCenter", C.c_int),
                ("GridSpaceCenter", C.c_int), ("writeAngles", C.c_int), ("tousepsf", C.c_int),
                ("doquater", C.c_int), ("shiftX", C.c_int), ("shiftY", C.c_int),
                ("pixelSize", C.c_float), ("Ntotpi", C.c_float), ("volu", C.c_float),
                ("sigmaPriorbctf", C.c_float), ("sigmaPriordefo", C.c_float),
                ("Priordefcent", C.c_float), ("sigmaPrioramp", C.c_float),
                ("Priorampcent", C.c_float)]


PROB_MAP_DTYPE = np.dtype([("Total", "<f8"), ("Constoadd", "<f8"), ("cent_x", "<i4"),
                           ("cent_y", "<i4"), ("orient", "<i4"), ("conv", "<i4"),
                           ("norm", "<f4"), ("mu", "<f4")])
PROB_ANGLE_DTYPE = np.dtype([("forAngles", "<f8"), ("ConstAngle", "<f8")])
assert PROB_MAP_DTYPE.itemsize == 40 and PROB_ANGLE_DTYPE.itemsize == 16


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("bioem_oracle.cpp", "fft_core.hpp")]
    if force or not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path)
                                                for s in src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return path


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        fp = C.POINTER(C.c_float)
        L.oracle_volu.restype = C.c_float
        L.oracle_volu.argtypes = [C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float,
                                  C.c_float, C.c_float, C.c_float, C.c_float]
        L.oracle_voluang_list.restype = C.c_float
        L.oracle_voluang_list.argtypes = [C.c_int, C.c_float]
        L.oracle_model_prepare.restype = C.c_float
        L.oracle_model_prepare.argtypes = [fp, C.c_int, C.c_int]
        L.oracle_calc_logpro.restype = C.c_double
        L.oracle_calc_logpro.argtypes = [C.POINTER(OracleCfg)] + [C.c_float] * 8
        L.oracle_final_logprob.restype = C.c_double
        L.oracle_final_logprob.argtypes = [C.POINTER(OracleCfg), C.c_double, C.c_double]
        L.oracle_ctf_table.restype = C.c_int
        L.oracle_ctf_table.argtypes = [C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int,
                                       C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
                                       C.c_int, fp, fp, fp]
        L.oracle_projection.argtypes = [C.POINTER(OracleCfg), fp, C.c_int, C.c_float, fp, fp, fp]
        L.oracle_convolve.argtypes = [C.c_int, fp, fp, fp, fp, fp]
        L.oracle_cross_correlation.argtypes = [C.c_int, fp, fp, fp]
        L.oracle_particle_prepare.argtypes = [fp, C.c_int, C.c_int, fp, fp, fp]
        L.oracle_normalise_map.argtypes = [fp, C.c_int]
        L.oracle_defocus_to_phase.argtypes = [C.c_float, C.c_float, C.c_float, fp, fp, fp, fp]
        L.oracle_run.argtypes = [C.POINTER(OracleCfg), fp, C.c_int, C.c_float, fp, C.c_int,
                                 C.c_int, C.c_int, fp, fp, C.c_int, fp, fp, fp, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, fp, fp]
        L.oracle_fft_r2c.argtypes = [C.c_int, fp, fp]
        L.oracle_fft_c2r.argtypes = [C.c_int, fp, fp]
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def set_fft_double(flag: bool) -> None:
    lib().oracle_set_fft(int(flag))


class Prepared:
    """Everything the hot path consumes, derived the way the reference derives it
    (readParameters / CalculateRefCTF / readModel / RefMap.precalculate)."""

    def __init__(self, case, model, quats, particles, normalise_particles=None,
                 priorMod: float = 1.0, elecwavel: float = ELECWAVEL_DEFAULT):
        L = lib()
        n = case.n_pixels
        self.case = case
        self.N = n
        self.F = n * (n // 2 + 1)
        ctf = case.ctf
        f32 = np.float32
        sig_b = f32(ctf.get("SIGMA_PRIOR_B_CTF", 100.0))
        sig_d = C.c_float(ctf.get("SIGMA_PRIOR_DEFOCUS", 2.0))
        cen_d = C.c_float(ctf.get("PRIOR_DEFOCUS_CENTER", 3.0))
        sig_a = f32(ctf.get("SIGMA_PRIOR_AMP_CTF", 0.5))
        cen_a = f32(ctf.get("PRIOR_AMP_CTF_CENTER", 0.0))
        self.use_psf = "PSF_PHASE" in ctf
        grids = np.zeros(3, dtype=np.float32)
        if not self.use_psf:
            d0, d1, nd = ctf["CTF_DEFOCUS"]
            b0, b1, nb = ctf["CTF_B_ENV"]
            a0, a1, na = ctf["CTF_AMPLITUDE"]
            p0, p1 = C.c_float(), C.c_float()
            L.oracle_defocus_to_phase(f32(d0), f32(d1), f32(elecwavel), C.byref(p0), C.byref(p1),
                                      C.byref(cen_d), C.byref(sig_d))
        else:  # USE_PSF: phase / envelope grids are given directly (param.cpp:330-385)
            p0, p1, nd = ctf["PSF_PHASE"]
            b0, b1, nb = ctf["PSF_ENVELOPE"]
            a0, a1, na = ctf["PSF_AMPLITUDE"]
            p0, p1 = C.c_float(p0), C.c_float(p1)
        self.C = int(na * nd * nb)
        self.refCTF = np.zeros((self.C, self.F, 2), dtype=np.float32)
        self.CtfParam = np.zeros((self.C, 3), dtype=np.float32)
        c = L.oracle_ctf_table(n, f32(case.pixel_size), int(self.use_psf), f32(a0), f32(a1), int(na), p0, p1,
                               int(nd), f32(b0), f32(b1), int(nb), _fp(self.refCTF),
                               _fp(self.CtfParam), _fp(grids))
        assert c == self.C
        self.quats = np.ascontiguousarray(quats, dtype=np.float32)
        self.O = self.quats.shape[0]
        voluang = L.oracle_voluang_list(self.O, f32(priorMod))
        volu = L.oracle_volu(voluang, case.grid_space, f32(case.pixel_size), case.max_disp,
                             int(na), grids[2], grids[1], sig_b, sig_d, sig_a)
        self.cfg = OracleCfg(n, case.max_disp, case.grid_space, case.write_angles, int(self.use_psf), 1, 0, 0,
                             f32(case.pixel_size), f32(n * n), volu, sig_b, sig_d, cen_d, sig_a,
                             cen_a)
        self.pts = np.ascontiguousarray(model, dtype=np.float32).copy()
        self.A = self.pts.shape[0]
        self.NormDen = L.oracle_model_prepare(_fp(self.pts), self.A, 1)
        if normalise_particles is None:
            normalise_particles = case.particle_format == "mrc"
        self.maps = np.ascontiguousarray(particles, dtype=np.float32).copy()
        self.M = self.maps.shape[0]
        if normalise_particles:
            for m in range(self.M):
                L.oracle_normalise_map(_fp(self.maps[m]), n)
        self.RefMapsFFT = np.zeros((self.M, self.F, 2), dtype=np.float32)
        self.sumRef = np.zeros(self.M, dtype=np.float32)
        self.sumsqRef = np.zeros(self.M, dtype=np.float32)
        L.oracle_particle_prepare(_fp(self.maps), self.M, n, _fp(self.RefMapsFFT),
                                  _fp(self.sumRef), _fp(self.sumsqRef))
        self.D = L.oracle_num_displacements(n, case.max_disp, case.grid_space)

    # -- single stages ------------------------------------------------------
    def projection(self, o: int, want_real: bool = False):
        out = np.zeros((self.F, 2), dtype=np.float32)
        real = np.zeros((self.N, self.N), dtype=np.float32) if want_real else None
        lib().oracle_projection(C.byref(self.cfg), _fp(self.pts), self.A, self.NormDen,
                                _fp(self.quats[o]), _fp(out), _fp(real))
        return (out, real) if want_real else out

    def convolve(self, proj, c: int):
        out = np.zeros((self.F, 2), dtype=np.float32)
        s, ss = C.c_float(), C.c_float()
        lib().oracle_convolve(self.N, _fp(proj), _fp(self.refCTF[c]), _fp(out), C.byref(s),
                              C.byref(ss))
        return out, s.value, ss.value

    def cross_correlation(self, conv, m: int):
        out = np.zeros((self.N, self.N), dtype=np.float32)
        lib().oracle_cross_correlation(self.N, _fp(conv), _fp(self.RefMapsFFT[m]), _fp(out))
        return out

    def logpro(self, c: int, sumC, sumsqC, value, m: int) -> float:
        amp, pha, env = (float(v) for v in self.CtfParam[c])
        return lib().oracle_calc_logpro(C.byref(self.cfg), amp, pha, env, sumC, sumsqC, value,
                                        float(self.sumRef[m]), float(self.sumsqRef[m]))

    def logpro_at(self, m: int, o: int, c: int, cent_x: int, cent_y: int) -> float:
        """float-narrowed logpro of one (image, orientation, ctf, reported centre)
        — the quantity calProb compares (bioem_algorithm.h:84-96)."""
        conv, s, ss = self.convolve(self.projection(o), c)
        cc = self.cross_correlation(conv, m)
        dx, dy = -cent_x, -cent_y
        v = np.float32(cc[dx % self.N, dy % self.N]) / np.float32(self.N * self.N)
        return float(np.float32(self.logpro(c, s, ss, float(v), m)))

    # -- whole run ------------------------------------------------------------
    def run(self, o_begin: int = 0, o_end: int | None = None, trace_image: int = -1):
        o_end = self.O if o_end is None else o_end
        pm = np.zeros(self.M, dtype=PROB_MAP_DTYPE)
        pa = np.zeros((self.O, self.M), dtype=PROB_ANGLE_DTYPE) if self.cfg.writeAngles else None
        second = np.zeros(self.M, dtype=np.float64)
        tl = tv = None
        if trace_image >= 0:
            tl = np.zeros((o_end - o_begin, self.C, self.D), dtype=np.float32)
            tv = np.zeros_like(tl)
        lib().oracle_run(C.byref(self.cfg), _fp(self.pts), self.A, self.NormDen, _fp(self.quats),
                         self.O, o_begin, o_end, _fp(self.refCTF), _fp(self.CtfParam), self.C,
                         _fp(self.RefMapsFFT), _fp(self.sumRef), _fp(self.sumsqRef), self.M,
                         pm.ctypes.data, pa.ctypes.data if pa is not None else None,
                         second.ctypes.data, trace_image, _fp(tl), _fp(tv))
        res = dict(prob=pm, angle=pa, second=second)
        if trace_image >= 0:
            res["trace_logpro"] = tl
            res["trace_value"] = tv
        return res

    def final_logprob(self, total, const) -> float:
        return lib().oracle_final_logprob(C.byref(self.cfg), float(total), float(const))
