"""Parsers for the reference's text outputs (Output_Probabilities, ANG_PROB;
reference bioem.cpp:1046-1374).  Used by tests and tools on both sides."""
from __future__ import annotations

import re

import numpy as np


def parse_output_probabilities(path: str, quaternions: bool = True) -> dict:
    """Returns dict of arrays: logp, const, angles [M,4], amp, defocus, env,
    cent_x, cent_y, norm, mu.  Images for which the reference wrote 'Warning'
    lines get NaN logp."""
    rows = {}
    with open(path) as f:
        for line in f:
            if line.startswith("RefMap:"):
                tok = line.split()
                m = int(tok[1])
                r = rows.setdefault(m, {})
                if tok[2] == "LogProb:":
                    r["logp"] = float(tok[3])
                    r["const"] = float(tok[5])
                elif tok[2] == "Maximizing":
                    vals = [t for t in tok[4:] if not t.startswith("[")]
                    v = [float(x) for x in vals]
                    na = 4 if quaternions else 3
                    if len(v) == na + 7 + 1:
                        r["maxlogp"] = v[0]
                        v = v[1:]
                    r["angles"] = v[:na] + [0.0] * (4 - na)
                    r["amp"], r["defocus"], r["env"] = v[na:na + 3]
                    r["cent_x"], r["cent_y"] = int(v[na + 3]), int(v[na + 4])
                    r["norm"], r["mu"] = v[na + 5], v[na + 6]
    n = max(rows) + 1 if rows else 0
    out = dict(logp=np.full(n, np.nan), const=np.full(n, np.nan), angles=np.zeros((n, 4)),
               amp=np.zeros(n), defocus=np.zeros(n), env=np.zeros(n),
               cent_x=np.zeros(n, dtype=int), cent_y=np.zeros(n, dtype=int), norm=np.zeros(n),
               mu=np.zeros(n))
    for m, r in rows.items():
        for k, v in r.items():
            if k in out:
                out[k][m] = v
    return out


def parse_ang_prob(path: str, quaternions: bool = True) -> list[list[dict]]:
    """Per image, the list of kept orientations in file order."""
    res: dict[int, list] = {}
    na = 4 if quaternions else 3
    with open(path) as f:
        for line in f:
            if "Separated:" not in line:
                continue
            tok = line.split()
            m = int(tok[0])
            ang = [float(x) for x in tok[1:1 + na]]
            logp = float(tok[1 + na])
            sep = tok.index("Separated:")
            res.setdefault(m, []).append(dict(angles=ang, logp=logp, log_for=float(tok[sep + 1]),
                                              const=float(tok[sep + 2]), add=float(tok[sep + 3])))
    return [res[m] for m in sorted(res)]


_PROB_RE = re.compile(r"Prob: iRefMap (\d+), iOrient (\d+), iConv (\d+), disx (-?\d+), disy (-?\d+), "
                      r"address -, value (\S+), logpro (\S+)")


def parse_debug_prob(text: str) -> np.ndarray:
    """Rows (iRefMap, iOrient, iConv, disx, disy, value, logpro) of the reference's
    -DDEBUG_PROB stream (bioem_algorithm.h:88-92), in emission order."""
    rows = [[float(x) for x in m.groups()] for m in _PROB_RE.finditer(text)]
    return np.asarray(rows, dtype=np.float64).reshape(-1, 7)
