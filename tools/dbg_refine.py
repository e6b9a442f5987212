#!/usr/bin/env python3
"""Diagnostic: does re-evaluating a particle's winning (orientation, CTF) reproduce the record's logpro bit for bit?"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from bioem_b200 import api  # noqa: E402
from bioem_b200.cases import build_case  # noqa: E402

cd = build_case("cfg2", n_particles=64, n_orient=40)
hi, parts = api.inputs_for_case(cd)
f32 = np.float32


def host_lp(e, o, c, m):
    v = e.debug_correlation(o, c, m).astype(np.float32).ravel()
    _, sC, ssC = e.debug_convolved(o, c)
    _, sR, ssR = e.debug_particle(m)
    sC, ssC, sR, ssR = f32(sC), f32(ssC), f32(sR), f32(ssR)
    Nt = f32(hi.cfg.Ntotpi)
    fe = Nt * (ssR * ssC - v * v) + (f32(2.0) * sR * sC) * v - (ssR * sC) * sC - (sR * sR) * ssC
    fl = ssC * Nt - sC * sC
    amp, pha, env = (float(x) for x in hi.CtfParam[c][:3])
    g = hi.cfg
    prior = (env * env / 2. / g.sigmaPriorbctf / g.sigmaPriorbctf - (pha - g.Priordefcent) ** 2 / 2. / g.sigmaPriordefo ** 2
             - (amp - g.Priorampcent) ** 2 / 2. / g.sigmaPrioramp ** 2)
    a = (3.0 - float(Nt)) * 0.5
    bterm = (float(Nt) * 0.5 - 2.0) * np.log(float(Nt - f32(2.0)) * float(fl)) - prior
    lp = (a * np.log(fe.astype(np.float64)) + bterm).astype(np.float32)
    return lp, fe


for og in ("1", "2"):
    os.environ["BIOEM_B200_OG"] = og
    e = api.Engine(hi.cfg)
    e.upload_all(hi, parts)
    os.environ["BIOEM_B200_NO_EXACT_ARGMAX"] = "1"
    e.run()
    raw, _ = e.download()
    del os.environ["BIOEM_B200_NO_EXACT_ARGMAX"]
    e.reset()
    e.run()
    pm, _ = e.download()
    print("OG", og, "exact pass (records, corrected, disagreed):", e.exact_argmax_info(), flush=True)
    bad = 0
    for m in range(parts.shape[0]):
        o, c = int(pm[m]["orient"]), int(pm[m]["conv"])
        lp, fe = host_lp(e, o, c, m)
        # the same likelihood again through the regular path, alone in its launch
        e.reset()
        e.run(o, o + 1)
        os.environ["BIOEM_B200_NO_EXACT_ARGMAX"] = "1"
        alone, _ = e.download()
        del os.environ["BIOEM_B200_NO_EXACT_ARGMAX"]
        ok_host = float(lp.max()) == pm[m]["Constoadd"]
        ok_alone = alone[m]["Constoadd"] == pm[m]["Constoadd"] and alone[m]["conv"] == c
        if not (ok_host and ok_alone):
            bad += 1
            print(f"  m={m} o={o} c={c}: record {pm[m]['Constoadd']!r} host-from-debug {float(lp.max())!r} alone {alone[m]['Constoadd']!r} "
                  f"(alone conv {alone[m]['conv']}) raw==refined lin: {raw[m]['cent_x'] == pm[m]['cent_x'] and raw[m]['cent_y'] == pm[m]['cent_y']}")
    print("  mismatching images:", bad, "of", parts.shape[0], flush=True)
    e.close()
