"""Named synthetic cases (SURVEY §8 size table / BASELINE.json configs) and a
builder that writes them in the reference's file formats."""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import synth

CFG1_CTF = dict(CTF_DEFOCUS=(1.0, 4.0, 3), CTF_B_ENV=(2.0, 300.0, 1), CTF_AMPLITUDE=(0.1, 0.1, 1))
# USE_PSF (real-space point-spread functions): phase / envelope in 1/A^2 (reference doc, PSF keywords)
PSF_GRID = dict(PSF_PHASE=(0.004, 0.016, 3), PSF_ENVELOPE=(0.006, 0.018, 2), PSF_AMPLITUDE=(0.1, 0.1, 1),
                SIGMA_PRIOR_B_CTF=200.0, SIGMA_PRIOR_DEFOCUS=8000.0, PRIOR_DEFOCUS_CENTER=3000.0)
DENSE_CTF = dict(CTF_DEFOCUS=(0.5, 4.5, 16), CTF_B_ENV=(2.0, 300.0, 8),
                 CTF_AMPLITUDE=(0.05, 0.25, 2), SIGMA_PRIOR_B_CTF=50.0,
                 SIGMA_PRIOR_DEFOCUS=0.4, PRIOR_DEFOCUS_CENTER=2.8)


@dataclass
class Case:
    name: str
    n_pixels: int
    pixel_size: float
    n_atoms: int
    n_particles: int
    quat_list: int            # which shipped list (576 / 4608 / 36864)
    n_orient: int             # how many leading rows of it are used
    ctf: dict = field(default_factory=dict)
    max_disp: int = 10
    grid_space: int = 1
    write_angles: int = 0
    model_sigma: float = 18.0
    model_rmax: float = 45.0
    particle_format: str = "text"   # "text" | "mrc"
    extra: tuple = ()               # further parameter-file lines (optional keywords)
    angle_priors: bool = False      # PRIOR_ANGLES: a fifth column (log prior) in the orientation list
    euler_grid: tuple = ()          # (GRIDPOINTS_ALPHA, GRIDPOINTS_BETA): Euler-angle grid instead of a quaternion list
    small_radii: bool = False       # every second model point gets a radius below the pixel size (point branch)
    voxel_model: int = 0            # edge of an MRC density volume read with --ReadModelMRC (0: text model)
    bulk_particles: int = 0         # > 0: large stack made from that many distinct clean signals (throughput workloads)

    @property
    def use_psf(self) -> bool:
        return "PSF_PHASE" in self.ctf

    @property
    def n_ctf(self) -> int:
        if self.use_psf:
            return int(self.ctf["PSF_AMPLITUDE"][2] * self.ctf["PSF_PHASE"][2] * self.ctf["PSF_ENVELOPE"][2])
        return int(self.ctf["CTF_AMPLITUDE"][2] * self.ctf["CTF_DEFOCUS"][2] * self.ctf["CTF_B_ENV"][2])

    @property
    def likelihoods(self) -> int:
        return self.n_orient * self.n_ctf * self.n_particles


CASES = {
    # tiny known-answer cases (oracle / reference finish in well under a second)
    "toy32": Case("toy32", 32, 1.5, 60, 3, 576, 24, CFG1_CTF, 4, 1, write_angles=3,
                  model_sigma=5.0, model_rmax=12.0),
    "toy32psf": Case("toy32psf", 32, 1.5, 60, 3, 576, 16, PSF_GRID, 4, 1, model_sigma=5.0, model_rmax=12.0),
    # edge windows: a single displacement, and the largest window the image allows
    "toy32d0": Case("toy32d0", 32, 1.5, 60, 3, 576, 12, CFG1_CTF, 0, 1, model_sigma=5.0, model_rmax=12.0),
    "toy32full": Case("toy32full", 32, 1.5, 60, 3, 576, 6, CFG1_CTF, 15, 1, model_sigma=5.0, model_rmax=12.0),
    # optional keywords of the parameter file in one case: sphere-footprint shift, model prior, no centring,
    # another electron wavelength, amplitude prior, per-orientation priors from the orientation list
    "toy32opts": Case("toy32opts", 32, 1.5, 60, 3, 576, 20, CFG1_CTF, 4, 1, write_angles=4, model_sigma=5.0,
                      model_rmax=12.0, extra=("SHIFT_X 1", "SHIFT_Y -1", "PRIOR_MODEL 0.25", "NO_CENTEROFMASS",
                                              "ELECTRON_WAVELENGTH 0.0251", "SIGMA_PRIOR_AMP_CTF 0.3",
                                              "PRIOR_AMP_CTF_CENTER 0.1", "SIGMA_PRIOR_DEFOCUS 1.5"),
                      angle_priors=True),
    # Euler-angle grid generated from the parameter file (ZXZ rotation path, alpha/beta/gamma output header)
    "toy32euler": Case("toy32euler", 32, 1.5, 60, 3, 576, 16, CFG1_CTF, 4, 1, write_angles=3, model_sigma=5.0,
                       model_rmax=12.0, euler_grid=(4, 3)),
    # both rasterisation branches of createProjection: radius <= pixel size adds the density to one pixel
    # (bioem.cpp:1719-1739), larger radii the sphere footprint (:1745-1801)
    "toy32pts": Case("toy32pts", 32, 1.5, 60, 3, 576, 16, CFG1_CTF, 4, 1, model_sigma=5.0, model_rmax=12.0,
                     small_radii=True),
    # a model wider than the frame: points whose pixel (or sphere footprint) leaves the image are skipped with
    # a warning, in both rasterisation branches (bioem.cpp:1724-1734,1756-1780)
    "toy32clip": Case("toy32clip", 32, 1.5, 60, 3, 576, 16, CFG1_CTF, 4, 1, model_sigma=13.0, model_rmax=30.0,
                      small_radii=True),
    # all three CTF grid axes longer than one: pins the enumeration order amplitude / defocus / envelope
    "toy32amp": Case("toy32amp", 32, 1.5, 60, 3, 576, 12, dict(CTF_DEFOCUS=(1.0, 4.0, 3), CTF_B_ENV=(2.0, 300.0, 2),
                                                                CTF_AMPLITUDE=(0.05, 0.25, 2)), 4, 1,
                     model_sigma=5.0, model_rmax=12.0),
    "toy36g2": Case("toy36g2", 36, 1.5, 60, 4, 576, 16, CFG1_CTF, 6, 2, model_sigma=6.0,
                    model_rmax=14.0, particle_format="mrc"),
    "toy64": Case("toy64", 64, 1.5, 200, 5, 576, 32, synth.PRODUCTION_GRID, 10, 1,
                  model_sigma=10.0, model_rmax=26.0, particle_format="mrc"),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "cfg1": Case("cfg1", 128, 1.5, 1000, 10, 576, 576, CFG1_CTF, 10, 1),
    # reduced 224-pixel case with the production grid (parity at the headline size)
    "cfg2_slice": Case("cfg2_slice", 224, 1.0, 1000, 6, 4608, 8, synth.PRODUCTION_GRID, 40, 1,
                       particle_format="mrc"),
    # BASELINE.json configs[1]: the headline workload
    "cfg2": Case("cfg2", 224, 1.0, 1000, 1000, 4608, 4608, synth.PRODUCTION_GRID, 40, 1,
                 particle_format="mrc"),
    "cfg3": Case("cfg3", 224, 1.0, 1000, 10000, 36864, 36864, synth.PRODUCTION_GRID, 40, 1,
                 particle_format="mrc", bulk_particles=1000),
    # DISPLACE_CENTER whose spacing does not divide the maximum displacement: Algo 1 enumerates 0,2,4 and
    # N-5,N-3,N-1 (3 + 3 points per axis), not a symmetric set (quirk Q3, bioem_algorithm.h:156-197)
    # image edges without a fused FFT kernel (direct-DFT path of the library): odd, odd with a displacement grid and
    # WRITE_PROB_ANGLES, even with a prime factor above 7
    "toy33": Case("toy33", 33, 1.5, 60, 3, 576, 12, CFG1_CTF, 4, 1, model_sigma=5.0, model_rmax=12.0),
    "toy35g2": Case("toy35g2", 35, 1.5, 60, 3, 576, 10, CFG1_CTF, 5, 2, write_angles=2, model_sigma=5.0, model_rmax=12.0),
    "toy26": Case("toy26", 26, 1.5, 60, 3, 576, 12, CFG1_CTF, 3, 1, model_sigma=4.0, model_rmax=10.0),
    "toy32g2odd": Case("toy32g2odd", 32, 1.5, 60, 3, 576, 12, CFG1_CTF, 5, 2, model_sigma=5.0, model_rmax=12.0),
    "toy32g3": Case("toy32g3", 32, 1.5, 60, 3, 576, 12, CFG1_CTF, 7, 3, write_angles=2, model_sigma=5.0, model_rmax=12.0),
    # BASELINE.json configs[3] at its named shape: 48^3 MRC density volume (110,592 points of radius 2 px) read with
    # --ReadModelMRC, 360 x 360 pixels, dense CTF grid 16 x 8 x 2 = 256; a slice of particles / orientations so
    # that the unmodified reference finishes in seconds
    "cfg4_voxel_slice": Case("cfg4_voxel_slice", 360, 2.0, 0, 3, 4608, 2, DENSE_CTF, 40, 1, particle_format="mrc",
                             voxel_model=48),
    "cfg4": Case("cfg4", 360, 2.0, 0, 2000, 4608, 4608, DENSE_CTF, 40, 1, particle_format="mrc", voxel_model=48,
                 bulk_particles=64),
    "cfg5": Case("cfg5", 224, 1.0, 1000, 5000, 4608, 4608, synth.PRODUCTION_GRID, 40, 1, write_angles=10,
                 particle_format="mrc", bulk_particles=1000),
    "cfg4_slice": Case("cfg4_slice", 360, 2.0, 1000, 4, 4608, 4, DENSE_CTF, 40, 1,
                       model_sigma=40.0, model_rmax=110.0, particle_format="mrc"),
    "cfg5_slice": Case("cfg5_slice", 224, 1.0, 1000, 5, 4608, 12, synth.PRODUCTION_GRID, 40, 1,
                       write_angles=10, particle_format="mrc"),
}


@dataclass
class CaseData:
    case: Case
    model: np.ndarray        # [A,5] float64 (text-file precision)
    quats: np.ndarray        # [O,4] float32
    particles: np.ndarray    # [M,N,N] float32 as the reader will see them
    truth: np.ndarray        # [M,4]
    paths: dict


def build_case(name_or_case, outdir: str | None = None, n_particles: int | None = None,
               n_orient: int | None = None) -> CaseData:
    c = CASES[name_or_case] if isinstance(name_or_case, str) else name_or_case
    if n_particles is not None or n_orient is not None:
        c = Case(**{**c.__dict__, "n_particles": n_particles or c.n_particles,
                    "n_orient": n_orient or c.n_orient})
    vol = None
    if c.voxel_model:
        vol = synth.make_volume(c.voxel_model, seed=2)
        model = synth.volume_to_points(vol, c.pixel_size)
    else:
        model = synth.make_model(c.n_atoms, seed=1, sigma=c.model_sigma, rmax=c.model_rmax)
    if c.small_radii:
        model[::2, 3] = np.round(model[::2, 3] * 0.3, 4)
    quats = synth.load_quaternions(c.quat_list)[:c.n_orient].copy()
    ctfp = synth.ctf_grid_params(CFG1_CTF if c.use_psf else c.ctf)
    if c.bulk_particles:
        imgs, truth = synth.make_particles_bulk(model, quats, c.n_pixels, c.pixel_size, c.n_particles, c.max_disp, ctfp,
                                                n_distinct=c.bulk_particles, snr=0.1, seed=100,
                                                normalise=(c.particle_format == "mrc"))
    else:
        imgs, truth = synth.make_particles(model, quats, c.n_pixels, c.pixel_size, c.n_particles,
                                           c.max_disp, ctfp, snr=0.1, seed=100,
                                           normalise=(c.particle_format == "mrc"))
    if c.particle_format == "text":
        # the text format carries 8 decimals: make array == file
        imgs = np.round(imgs.astype(np.float64), 8).astype(np.float32)
    paths = {}
    if outdir is not None:
        os.makedirs(outdir, exist_ok=True)
        paths = dict(model=os.path.join(outdir, "model.mrc" if c.voxel_model else "model.txt"),
                     param=os.path.join(outdir, "param.txt"),
                     orient=os.path.join(outdir, "orient.txt"),
                     particles=os.path.join(outdir, "particles.mrc" if c.particle_format == "mrc"
                                            else "particles.txt"))
        if c.voxel_model:
            synth.write_volume_mrc(paths["model"], vol)
        else:
            synth.write_model_text(paths["model"], model)
        extra = list(c.extra) + (["USE_PSF", "WRITE_CTF_PARAM 1"] if c.use_psf else []) + \
            (["PRIOR_ANGLES"] if c.angle_priors else []) + \
            ([f"GRIDPOINTS_ALPHA {c.euler_grid[0]}", f"GRIDPOINTS_BETA {c.euler_grid[1]}"] if c.euler_grid else [])
        synth.write_param_file(paths["param"], c.n_pixels, c.pixel_size, c.max_disp, c.grid_space,
                               c.ctf, not c.euler_grid, c.write_angles, extra=extra or None)
        if c.angle_priors:  # log priors -0.5, -0.25, 0, ... per orientation (exact in the 12-column format)
            pri = (-0.5 + 0.25 * (np.arange(len(quats)) % 5)).astype(np.float32)
            synth.write_orientation_list(paths["orient"], np.concatenate([quats, pri[:, None]], axis=1))
        else:
            synth.write_orientation_list(paths["orient"], quats)
        if c.particle_format == "mrc":
            synth.write_particles_mrc(paths["particles"], imgs)
        else:
            synth.write_particles_text(paths["particles"], imgs)
    return CaseData(c, model, quats, imgs, truth, paths)


def reference_cli(cd: CaseData, outfile: str = "Output_Probabilities") -> list[str]:
    """Command-line arguments (after the binary name) in the reference's CLI
    (bioem.cpp:193-224) for a built case."""
    a = ["--Modelfile", cd.paths["model"], "--Particlesfile", cd.paths["particles"],
         "--Inputfile", cd.paths["param"], "--OutputFile", outfile]
    if not cd.case.euler_grid:
        a += ["--ReadOrientation", cd.paths["orient"]]
    if cd.case.particle_format == "mrc":
        a.append("--ReadMRC")
    if cd.case.voxel_model:
        a.append("--ReadModelMRC")
    return a
