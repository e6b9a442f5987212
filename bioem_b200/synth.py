"""Synthetic inputs for the BioEM likelihood path (numpy only, seeded).

Writers produce the reference's on-disk formats so the same files feed the
reference binary, the oracle and the B200 path:

* model text  ``x y z radius density``           (reference model.cpp:419-601)
* parameter file keywords                        (reference param.cpp:121-527)
* particle text stack ``PARTICLE`` + ``%8d%8d%16.8f`` rows, 32 chars + newline
                                                 (reference map.cpp:268-414)
* particle MRC mode-2 stack (1024-byte header; the reader transposes and
  normalises each image)                          (reference map.cpp:663-853)
* orientation list, fixed 12-character columns    (reference param.cpp:1213-1327)

The particle generator uses its own small numpy projector (atoms splatted as
projected spheres, CTF applied in Fourier space) — it only has to make
plausible images; it is not the checker and not the product path.
"""
from __future__ import annotations

import os
import struct

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

PRODUCTION_GRID = dict(  # reference doc/index.rst:2098-2116
    CTF_B_ENV=(2.0, 300.0, 4),
    CTF_DEFOCUS=(0.5, 4.5, 8),
    CTF_AMPLITUDE=(0.1, 0.1, 1),
    SIGMA_PRIOR_B_CTF=50.0,
    SIGMA_PRIOR_DEFOCUS=0.4,
    PRIOR_DEFOCUS_CENTER=2.8,
)


def load_quaternions(n: int) -> np.ndarray:
    """float32 [n,4] = (q0,q1,q2,q3) of the reference's QUATERNION_LIST_<n>_Orient
    exactly as the reference's fixed-width parser reads them.
    Fixtures made by tools/make_quaternion_fixtures.py."""
    path = os.path.join(_DATA, f"quat_{n}.npy")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no orientation fixture for {n} orientations: {path}")
    return np.load(path)


def make_model(n_points: int = 1000, seed: int = 1, sigma: float = 18.0,
               rmax: float = 45.0) -> np.ndarray:
    """[A,5] float64 rows x y z radius density (pseudo-residues)."""
    rng = np.random.default_rng(seed)
    pts = []
    while len(pts) < n_points:
        p = rng.normal(0.0, sigma, size=(2 * n_points, 3))
        p = p[np.linalg.norm(p, axis=1) <= rmax]
        pts.extend(p.tolist())
    pos = np.asarray(pts[:n_points])
    rad = rng.uniform(2.25, 3.4, size=n_points)
    den = rng.uniform(40.0, 108.0, size=n_points)
    # round to the precision the text file carries, so file == array
    m = np.column_stack([pos, rad, den])
    return np.round(m, 4)


def make_volume(edge: int = 48, seed: int = 2) -> np.ndarray:
    """float32 [nx,ny,nz] density map for the MRC-volume model (cfg 4): two Gaussian blobs on a
    5 % noise floor, clipped at 0."""
    rng = np.random.default_rng(seed)
    g = (np.arange(edge) + 0.5) / edge - 0.5
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    v = np.exp(-((x + 0.12) ** 2 + (y - 0.05) ** 2 + (z + 0.08) ** 2) / (2 * 0.11 ** 2))
    v += 0.8 * np.exp(-((x - 0.16) ** 2 + (y + 0.1) ** 2 + (z - 0.12) ** 2) / (2 * 0.08 ** 2))
    v += 0.05 * rng.normal(size=v.shape)
    return np.clip(v, 0.0, None).astype(np.float32)


def volume_to_points(vol: np.ndarray, px: float) -> np.ndarray:
    """[A,5] rows x y z radius density exactly as the reference's --ReadModelMRC reader makes them
    (model.cpp:375-396): one point per voxel, file order (first index outermost), 1-based indices,
    position (i - nx/2.0)*px, radius 2*px, density = voxel."""
    nx, ny, nz = vol.shape
    i, j, k = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    pos = np.stack([(i - nx / 2.0) * px, (j - ny / 2.0) * px, (k - nz / 2.0) * px], axis=-1).reshape(-1, 3)
    a = pos.shape[0]
    return np.column_stack([pos, np.full(a, 2.0 * px), vol.reshape(-1).astype(np.float64)])


def write_volume_mrc(path: str, vol: np.ndarray) -> None:
    """Mode-2 MRC volume in the element order the reference's reader consumes (model.cpp:375-381: the first
    index is the outermost loop over the file's floats)."""
    nx, ny, nz = vol.shape
    hdr = bytearray(1024)
    struct.pack_into("<4i", hdr, 0, nx, ny, nz, 2)
    struct.pack_into("<3i", hdr, 28, nx, ny, nz)
    struct.pack_into("<3f", hdr, 40, float(nx), float(ny), float(nz))
    struct.pack_into("<3f", hdr, 52, 90.0, 90.0, 90.0)
    struct.pack_into("<3i", hdr, 64, 1, 2, 3)
    hdr[208:212] = b"MAP "
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.ascontiguousarray(vol, dtype="<f4").tobytes())


def write_model_text(path: str, model: np.ndarray) -> None:
    with open(path, "w") as f:
        for x, y, z, r, d in model:
            f.write(f"{x:.4f} {y:.4f} {z:.4f} {r:.4f} {d:.4f}\n")


def write_param_file(path: str, n_pixels: int, pixel_size: float, max_disp: int,
                     grid_space: int, ctf: dict, use_quaternions: bool = True,
                     write_prob_angles: int = 0, extra: list[str] | None = None) -> None:
    """ctf: dict with CTF_B_ENV / CTF_DEFOCUS / CTF_AMPLITUDE triples and
    optional prior keywords (see PRODUCTION_GRID)."""
    lines = [
        f"PIXEL_SIZE {pixel_size}",
        f"NUMBER_PIXELS {n_pixels}",
        f"DISPLACE_CENTER {max_disp} {grid_space}",
    ]
    if use_quaternions:
        lines.append("USE_QUATERNIONS")
    for k in ("CTF_B_ENV", "CTF_DEFOCUS", "CTF_AMPLITUDE", "PSF_ENVELOPE", "PSF_PHASE",
              "PSF_AMPLITUDE"):
        if k in ctf:
            a, b, n = ctf[k]
            lines.append(f"{k} {a} {b} {int(n)}")
    for k in ("SIGMA_PRIOR_B_CTF", "SIGMA_PRIOR_DEFOCUS", "PRIOR_DEFOCUS_CENTER",
              "SIGMA_PRIOR_AMP_CTF", "PRIOR_AMP_CTF_CENTER"):
        if k in ctf:
            lines.append(f"{k} {ctf[k]}")
    if write_prob_angles:
        lines.append(f"WRITE_PROB_ANGLES {int(write_prob_angles)}")
    if extra:
        lines.extend(extra)
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def write_orientation_list(path: str, q: np.ndarray) -> None:
    """Fixed 12-char columns; chosen so that the reference's fixed-width parse
    returns exactly the float32 values in q (checked)."""
    q = np.asarray(q, dtype=np.float32)
    with open(path, "w") as f:
        f.write(f"{q.shape[0]:<12d}\n")
        for row in q:
            cols = []
            for v in row:
                s = f"{float(v):12.8f}"
                if np.float32(float(s)) != v:
                    s = f"{float(v):12.9f}"[:12] if v >= 0 else f"{float(v):12.9f}"
                    if len(s) != 12 or np.float32(float(s)) != v:
                        s = f"{float(v):12.5e}"
                        assert len(s) == 12 and np.float32(float(s)) == v, (v, s)
                cols.append(s)
            f.write("".join(cols) + "\n")


def parse_orientation_list(path: str, ncol: int = 4) -> np.ndarray:
    """The reference's parser: first line = count in the first 12 chars, then
    fixed 12-character columns read with sscanf("%f") (param.cpp:1221-1327)."""
    with open(path) as f:
        lines = f.read().split("\n")
    n = int(lines[0][:12].split()[0])
    out = np.zeros((n, 4), dtype=np.float32)
    r = 0
    for line in lines[1:]:
        if r >= n and line.strip() == "":
            continue
        if line == "" and r < n:
            continue
        for k in range(ncol):
            tok = line[12 * k:12 * k + 12].split()
            out[r, k] = np.float32(float(tok[0]))
        r += 1
    assert r == n, (r, n)
    return out


def write_particles_text(path: str, imgs: np.ndarray) -> None:
    m, n, _ = imgs.shape
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    ii = ii.ravel()
    jj = jj.ravel()
    with open(path, "w") as f:
        for k in range(m):
            f.write(f"PARTICLE {k + 1}\n")
            v = imgs[k].ravel()
            f.write("".join(f"{a:8d}{b:8d}{c:16.8f}\n" for a, b, c in zip(ii, jj, v)))


def write_particles_mrc(path: str, imgs: np.ndarray) -> None:
    """Mode-2 stack.  The reference stores file element (section s, row j,
    column i) at maps[s][i*N+j] (map.cpp:815-826), so write the transpose to
    make imgs[s][i][j] come back as maps[s][i*N+j]."""
    m, n, _ = imgs.shape
    hdr = bytearray(1024)
    struct.pack_into("<4i", hdr, 0, n, n, m, 2)
    struct.pack_into("<3i", hdr, 28, n, n, m)
    struct.pack_into("<3f", hdr, 40, float(n), float(n), float(m))
    struct.pack_into("<3f", hdr, 52, 90.0, 90.0, 90.0)
    struct.pack_into("<3i", hdr, 64, 1, 2, 3)
    hdr[208:212] = b"MAP "
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.ascontiguousarray(np.transpose(imgs, (0, 2, 1)), dtype="<f4").tobytes())


def quat_to_rot(q: np.ndarray) -> np.ndarray:
    """Rotation matrix in the reference's convention (bioem.cpp:1638-1646)."""
    q0, q1, q2, q3 = [float(x) for x in q]
    return np.array([
        [1 - 2 * q1 * q1 - 2 * q2 * q2, 2 * (q0 * q1 + q2 * q3), 2 * (q0 * q2 - q1 * q3)],
        [2 * (q0 * q1 - q2 * q3), 1 - 2 * q0 * q0 - 2 * q2 * q2, 2 * (q1 * q2 + q0 * q3)],
        [2 * (q0 * q2 + q1 * q3), 2 * (q1 * q2 - q0 * q3), 1 - 2 * q0 * q0 - 2 * q1 * q1],
    ])


def project_numpy(model: np.ndarray, q: np.ndarray, n: int, px: float) -> np.ndarray:
    """Small stand-alone projector for making particles (sphere footprints, vectorised)."""
    rot = quat_to_rot(q)
    p = model[:, :3] @ rot.T
    img = np.zeros((n, n), dtype=np.float64)
    ci = np.floor(p[:, 0] / px + n / 2.0 + 0.5).astype(int)
    cj = np.floor(p[:, 1] / px + n / 2.0 + 0.5).astype(int)
    r = model[:, 3]
    d = model[:, 4]
    small = r <= px
    ok = small & (ci >= 0) & (ci < n) & (cj >= 0) & (cj < n)
    np.add.at(img, (ci[ok], cj[ok]), d[ok])
    ir = (r / px).astype(int) + 1
    big = (~small) & (ci >= ir) & (cj >= ir) & (ci < n - ir) & (cj < n - ir)
    if big.any():
        irmax = int(ir[big].max())
        o = np.arange(-irmax, irmax + 1)
        di, dj = np.meshgrid(o, o, indexing="ij")
        di = di.ravel()[None, :]
        dj = dj.ravel()[None, :]
        rb, db, cib, cjb, irb = r[big, None], d[big, None], ci[big, None], cj[big, None], ir[big, None]
        dist = (di * di + dj * dj) * px * px
        inside = (dist < rb * rb) & (np.abs(di) <= irb) & (np.abs(dj) <= irb)
        w = 2.0 * np.sqrt(np.maximum(rb * rb - dist, 0.0)) * db * 3.0 / (4 * np.pi * rb ** 3) * px * px
        ii = np.broadcast_to(cib + di, w.shape)[inside]
        jj = np.broadcast_to(cjb + dj, w.shape)[inside]
        np.add.at(img, (ii, jj), w[inside])
    return img


def ctf_table_numpy(n: int, px: float, amp: float, phase: float, env: float) -> np.ndarray:
    """Smooth radial CTF on the full n x (n/2+1) half-spectrum (no mirror-row quirk;
    only used to make particles)."""
    ky = np.fft.fftfreq(n, d=1.0) * n
    kx = np.arange(n // 2 + 1)
    radsq = (ky[:, None] ** 2 + kx[None, :] ** 2) / n / n / px / px
    c = np.exp(-env * radsq / 2) * (-amp * np.cos(phase * radsq / 2)
                                    - np.sqrt(1 - amp * amp) * np.sin(phase * radsq / 2))
    return c / c[0, 0]


def make_particles(model: np.ndarray, quats: np.ndarray, n: int, px: float, m: int,
                   max_disp: int, ctf_params: np.ndarray, snr: float = 0.1,
                   seed: int = 100, normalise: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """Returns (imgs float32 [m,n,n], truth int [m,4] = orientation, ctf, dx, dy).
    ctf_params: [C,3] rows (amp, phase, env)."""
    imgs = np.zeros((m, n, n), dtype=np.float32)
    truth = np.zeros((m, 4), dtype=np.int64)
    for k in range(m):
        rng = np.random.default_rng(seed + k)
        o = int(rng.integers(0, quats.shape[0]))
        c = int(rng.integers(0, ctf_params.shape[0]))
        h = max(max_disp // 2, 0)
        dx, dy = (int(v) for v in rng.integers(-h, h + 1, size=2))
        proj = project_numpy(model, quats[o], n, px)
        amp, pha, env = (float(v) for v in ctf_params[c])
        f = np.fft.rfft2(proj) * ctf_table_numpy(n, px, amp, pha, env)
        sig = np.fft.irfft2(f, s=(n, n))
        sig = np.roll(sig, (dx, dy), axis=(0, 1))
        sig = sig * rng.uniform(0.5, 2.0) + rng.normal(0.0, 1.0)
        noise = rng.normal(0.0, np.sqrt(sig.var() / snr), size=(n, n))
        im = sig + noise
        if normalise:
            im = (im - im.mean()) / im.std()
        imgs[k] = im.astype(np.float32)
        truth[k] = (o, c, dx, dy)
    return imgs, truth


def make_particles_bulk(model: np.ndarray, quats: np.ndarray, n: int, px: float, m: int, max_disp: int,
                        ctf_params: np.ndarray, n_distinct: int = 256, snr: float = 0.1, seed: int = 100,
                        normalise: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """Large stacks for the throughput workloads (cfg 3-5): n_distinct clean signals (orientation, CTF,
    shift drawn like make_particles), every particle = one of them with its own scale, offset and white
    noise at the given SNR.  Same statistics per image as make_particles at a fraction of the time."""
    nd = min(n_distinct, m)
    sig = np.zeros((nd, n, n), dtype=np.float32)
    tr = np.zeros((nd, 4), dtype=np.int64)
    for k in range(nd):
        rng = np.random.default_rng(seed + k)
        o = int(rng.integers(0, quats.shape[0]))
        c = int(rng.integers(0, ctf_params.shape[0]))
        h = max(max_disp // 2, 0)
        dx, dy = (int(v) for v in rng.integers(-h, h + 1, size=2))
        proj = project_numpy(model, quats[o], n, px)
        amp, pha, env = (float(v) for v in ctf_params[c])
        f = np.fft.rfft2(proj) * ctf_table_numpy(n, px, amp, pha, env)
        s_ = np.roll(np.fft.irfft2(f, s=(n, n)), (dx, dy), axis=(0, 1))
        sig[k] = s_.astype(np.float32)
        tr[k] = (o, c, dx, dy)
    imgs = np.empty((m, n, n), dtype=np.float32)
    truth = np.zeros((m, 4), dtype=np.int64)
    rng = np.random.default_rng(seed + 7919)
    sd = sig.reshape(nd, -1).std(axis=1)
    for k0 in range(0, m, 256):
        k1 = min(m, k0 + 256)
        idx = np.arange(k0, k1) % nd
        scale = rng.uniform(0.5, 2.0, size=(k1 - k0, 1, 1)).astype(np.float32)
        off = rng.normal(0.0, 1.0, size=(k1 - k0, 1, 1)).astype(np.float32)
        noise = rng.standard_normal(size=(k1 - k0, n, n), dtype=np.float32)
        noise *= (scale[:, 0, 0] * sd[idx] / np.sqrt(snr)).astype(np.float32)[:, None, None]
        im = sig[idx] * scale + off + noise
        if normalise:
            im -= im.mean(axis=(1, 2), keepdims=True)
            im /= im.std(axis=(1, 2), keepdims=True)
        imgs[k0:k1] = im
        truth[k0:k1] = tr[idx]
    return imgs, truth


def ctf_grid_params(ctf: dict, elecwavel: float = 0.019866) -> np.ndarray:
    """[C,3] float32 (amp, phase, env) in the reference's enumeration order
    (amp outer, phase, envelope inner; param.cpp:1365-1436) — for particle making."""
    a0, a1, na = ctf["CTF_AMPLITUDE"]
    d0, d1, nd = ctf["CTF_DEFOCUS"]
    b0, b1, nb = ctf["CTF_B_ENV"]
    p0 = d0 * np.pi * 2.0 * 10000 * elecwavel
    p1 = d1 * np.pi * 2.0 * 10000 * elecwavel
    ga = (a1 - a0) / na if na != 1 else a0
    gp = (p1 - p0) / nd if nd != 1 else p0
    gb = (b1 - b0) / nb if nb != 1 else b0
    out = []
    for ia in range(int(na)):
        for ip in range(int(nd)):
            for ib in range(int(nb)):
                out.append((ia * ga + a0, ip * gp + p0, ib * gb + b0))
    return np.asarray(out, dtype=np.float32)
