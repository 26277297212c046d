"""Deterministic synthetic trajectories for tests and benchmarks.

A frame is  x_f = R_f . (base + sum_m a[m,f] mode_m + eps_f) + t_f  with
  base   : Gaussian globule, radius of gyration 2.2 * N^0.38 Angstrom
  mode_m : 10 fixed random unit displacement fields, a[m,.] an AR(1) series
           (rho = 0.999, stationary per-atom RMS 3/sqrt(m+1) Angstrom)
  eps_f  : iid N(0, 0.3 Angstrom) thermal noise
  R_f    : random walk of small rotations (2 degrees / frame)
  t_f    : N(0, 5 Angstrom) translation
cast to float32 as a DCD/XTC/HDF5 file would hold it.  Neighbouring frames end
up 0.4-0.6 Angstrom apart and distant ones 3-8 Angstrom (the regime where
Ga + Gb - 2*lambda cancels).  Writers for the PDB and DCD layouts the reference
reads (src/pdb.cpp:66-141, src/dcd.cpp:175-261,329-371, src/dcdwriter.cpp:60-96)
are included so that the native front-end can be exercised on real files.
"""
import struct

import numpy as np

SEED0 = 20261019
N_MODES = 10


def _rot_from_rotvec(v):
    """Rodrigues: (K,3) rotation vectors -> (K,3,3)."""
    th = np.linalg.norm(v, axis=1)
    k = v / np.where(th > 0, th, 1.0)[:, None]
    K = np.zeros((len(v), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def _cumulative_rotations(steps):
    """Prefix product R_f = r_f . r_{f-1} ... r_0 of (F,3,3) by recursive doubling."""
    R = steps.copy()
    F = len(R)
    shift = 1
    while shift < F:
        R[shift:] = R[shift:] @ R[:-shift]
        shift *= 2
    # re-orthonormalise (drift after log2(F) products is ~1e-15, harmless)
    return R


def base_cloud(n, rng):
    rg = 2.2 * n ** 0.38
    return rng.normal(size=(n, 3)) * (rg / np.sqrt(3.0))


class Dynamics:
    """Frame generator for `n` tracked sites; frames are produced in chunks so
    that C2/C4/C5-sized sets never need more than one chunk of float64."""

    def __init__(self, n, seed, nframes):
        self.n, self.F = n, nframes
        rng = np.random.default_rng(seed)
        self.base = base_cloud(n, rng)
        modes = rng.normal(size=(N_MODES, n, 3))
        modes /= np.sqrt((modes ** 2).sum(axis=(1, 2), keepdims=True) / n)  # per-atom RMS 1
        self.modes = modes
        rho = 0.999
        sig = 3.0 / np.sqrt(np.arange(N_MODES) + 1.0)
        drive = rng.normal(size=(nframes, N_MODES)) * np.sqrt(1 - rho * rho)
        drive[0] = rng.normal(size=N_MODES)
        try:
            from scipy.signal import lfilter
            a = lfilter([1.0], [1.0, -rho], drive, axis=0)
        except Exception:  # pragma: no cover
            a = np.empty_like(drive); a[0] = drive[0]
            for f in range(1, nframes):
                a[f] = rho * a[f - 1] + drive[f]
        self.amp = a * sig[None, :]
        steps = _rot_from_rotvec(rng.normal(size=(nframes, 3)) * np.deg2rad(2.0) / np.sqrt(3.0))
        steps[0] = _rot_from_rotvec(rng.normal(size=(1, 3)) * 2.0)[0]
        self.rot = _cumulative_rotations(steps)
        self.trans = rng.normal(size=(nframes, 3)) * 5.0
        self.noise_seed = seed ^ 0x5EED

    def internal(self, lo, hi):
        """Un-rotated internal coordinates (float64) for frames [lo, hi)."""
        rng = np.random.default_rng([self.noise_seed, lo])
        x = self.base[None] + np.tensordot(self.amp[lo:hi], self.modes, axes=(1, 0))
        x += rng.normal(size=x.shape) * 0.3
        return x

    def frames(self, lo=0, hi=None, extra=None):
        """(hi-lo, n [+ len(extra)], 3) float32.  `extra` = (parent_index, offset)
        rides rigidly on the tracked sites (non-selected atoms of a rich model)."""
        hi = self.F if hi is None else hi
        x = self.internal(lo, hi)
        if extra is not None:
            parent, off = extra
            x = np.concatenate([x, x[:, parent] + off[None]], axis=1)
        x = np.einsum("fij,fnj->fni", self.rot[lo:hi], x) + self.trans[lo:hi, None, :]
        return x.astype(np.float32)


def synth_frames(F, N, seed=SEED0, chunk=4096):
    """(F, N, 3) float32 coordinates of the selected atoms only."""
    dyn = Dynamics(N, seed, F)
    out = np.empty((F, N, 3), dtype=np.float32)
    for lo in range(0, F, chunk):
        out[lo:lo + chunk] = dyn.frames(lo, min(F, lo + chunk))
    return out


# ---------------------------------------------------------------------------
# Models.  rich: residues ALA/A/PROT with atoms N CA C O CB HA + 64 TIP3/SOLV
# waters; lean: CA atoms with ~2 % interleaved decoy hydrogens.

RICH_ATOMS = ("N", "CA", "C", "O", "CB", "HA")


def make_model(n_res, flavour="rich", n_water=64):
    """List of dicts(name, resname, resid, chain, segid, track) in file order;
    `track` = index of the tracked site the atom rides on."""
    atoms = []
    if flavour == "rich":
        for r in range(n_res):
            for nm in RICH_ATOMS:
                atoms.append(dict(name=nm, resname="ALA", resid=r + 1, chain="A", segid="PROT", track=r))
        for w in range(n_water):
            for nm in ("OH2", "H1", "H2"):
                atoms.append(dict(name=nm, resname="TIP3", resid=n_res + w + 1, chain="W",
                                  segid="SOLV", track=(w * 7) % n_res))
    elif flavour == "lean":
        for r in range(n_res):
            atoms.append(dict(name="CA", resname="ALA", resid=r + 1, chain="A", segid="PROT", track=r))
            if r % 50 == 49:
                atoms.append(dict(name="HA", resname="ALA", resid=r + 1, chain="A", segid="PROT", track=r))
    else:
        raise ValueError(flavour)
    return atoms


def synth_system(n_res, F, seed=SEED0, flavour="rich"):
    """(atoms, coords[F, natoms, 3] float32) with atoms in model-file order."""
    atoms = make_model(n_res, flavour)
    dyn = Dynamics(n_res, seed, F)
    rng = np.random.default_rng(seed + 7)
    track = np.array([a["track"] for a in atoms])
    is_site = np.array([a["name"] == "CA" for a in atoms])
    off = rng.normal(size=(len(atoms), 3)) * 0.8
    off[is_site] = 0.0
    # tracked sites first, then everything as "extra", then reorder to file order
    x = dyn.frames(0, F, extra=(track, off))
    return atoms, np.ascontiguousarray(x[:, n_res:, :])


def _h36(width, value):
    """hybrid-36 encoder (src/utils.cpp:259-316 decodes it)."""
    if value < 10 ** width:
        return "%*d" % (width, value)
    digits_u = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
    value -= 10 ** width
    if value < 26 * 36 ** (width - 1):
        value += 10 * 36 ** (width - 1)
        s = ""
        while value:
            s = digits_u[value % 36] + s
            value //= 36
        return s
    raise ValueError("value out of hybrid-36 range")


def write_pdb(path, atoms, xyz):
    """Fixed-column ATOM records as src/pdb.cpp:66-141 parses them."""
    with open(path, "w") as fh:
        fh.write("REMARK synthetic model for loos_b200 tests\n")
        for i, (a, p) in enumerate(zip(atoms, xyz)):
            name = a["name"]
            name4 = (" " + name).ljust(4) if len(name) < 4 else name[:4]
            fh.write("ATOM  %5s %4s %-4s%1s%4s    %8.3f%8.3f%8.3f%6.2f%6.2f      %-4s%2s\n" % (
                _h36(5, i + 1), name4, a["resname"], a["chain"], _h36(4, a["resid"]),
                p[0], p[1], p[2], 1.0, 0.0, a["segid"], name[0].rjust(2)))
        fh.write("END\n")


def write_dcd(path, coords, with_crystal=False, big_endian=False, header_nframes=None):
    """CHARMM/NAMD DCD as src/dcd.cpp:175-261,329-371 reads it (layout written by
    src/dcdwriter.cpp:60-96,134-148): 84-byte CORD header record, title record,
    natoms record, then per frame [48-byte crystal record] X, Y, Z float records."""
    coords = np.asarray(coords, dtype=np.float32)
    F, n, _ = coords.shape
    e = ">" if big_endian else "<"
    with open(path, "wb") as fh:
        icntrl = [0] * 20
        icntrl[0] = F if header_nframes is None else header_nframes
        icntrl[1] = 1; icntrl[2] = 1; icntrl[3] = F
        icntrl[10] = 1 if with_crystal else 0
        icntrl[19] = 27
        body = b"CORD" + struct.pack(e + "9i", *icntrl[:9]) + struct.pack(e + "f", 1.0) + \
            struct.pack(e + "10i", *icntrl[10:])
        assert len(body) == 84
        fh.write(struct.pack(e + "i", 84) + body + struct.pack(e + "i", 84))
        title = b"synthetic trajectory for loos_b200".ljust(80)
        fh.write(struct.pack(e + "i", 84) + struct.pack(e + "i", 1) + title + struct.pack(e + "i", 84))
        fh.write(struct.pack(e + "i", 4) + struct.pack(e + "i", n) + struct.pack(e + "i", 4))
        rec = struct.pack(e + "i", 4 * n)
        xtal = struct.pack(e + "i", 48) + struct.pack(e + "6d", 100, 90, 100, 90, 90, 100) + struct.pack(e + "i", 48)
        fdt = np.dtype(np.float32).newbyteorder(e)
        for f in range(F):
            if with_crystal:
                fh.write(xtal)
            for k in range(3):
                fh.write(rec); fh.write(coords[f, :, k].astype(fdt).tobytes()); fh.write(rec)


def synth_frames_torch(F, N, seed=SEED0, device="cuda", chunk=8192):
    """Same model as synth_frames(), generated with torch on `device` (different
    random stream; used for the large benchmark workloads).  (F, N, 3) float32."""
    import math

    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    dd = dict(device=device, dtype=torch.float64)
    rg = 2.2 * N ** 0.38
    base = torch.randn((N, 3), generator=g, **dd) * (rg / math.sqrt(3.0))
    modes = torch.randn((N_MODES, N, 3), generator=g, **dd)
    modes = modes / torch.sqrt((modes ** 2).sum(dim=(1, 2), keepdim=True) / N)
    rho = 0.999
    sig = 3.0 / torch.sqrt(torch.arange(N_MODES, **dd) + 1.0)
    # AR(1) via blocked scan: a_f = rho^(f-f0) a_f0 + sum rho^(f-k) e_k, done in blocks of B
    B = 1024
    nb = (F + B - 1) // B
    e = torch.randn((nb * B, N_MODES), generator=g, **dd) * math.sqrt(1 - rho * rho)
    e[0] = torch.randn((N_MODES,), generator=g, **dd)
    pw = rho ** torch.arange(B, **dd)
    # lower-triangular Toeplitz of powers
    idx = torch.arange(B, device=device)
    Tm = torch.where(idx[:, None] >= idx[None, :], rho ** (idx[:, None] - idx[None, :]).to(torch.float64),
                     torch.zeros((), **dd))
    loc = torch.einsum("ij,bjm->bim", Tm, e.view(nb, B, N_MODES))
    amp = torch.empty((nb, B, N_MODES), **dd)
    carry = torch.zeros((N_MODES,), **dd)
    for b in range(nb):
        amp[b] = loc[b] + (pw * rho)[:, None] * carry[None, :]
        carry = amp[b, -1]
    amp = amp.view(nb * B, N_MODES)[:F] * sig[None, :]
    # rotation random walk: cumulative product by recursive doubling
    v = torch.randn((F, 3), generator=g, **dd) * (math.radians(2.0) / math.sqrt(3.0))
    v[0] = torch.randn((3,), generator=g, **dd) * 2.0
    th = v.norm(dim=1)
    k = v / torch.where(th > 0, th, torch.ones_like(th))[:, None]
    K = torch.zeros((F, 3, 3), **dd)
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    R = torch.eye(3, **dd)[None] + torch.sin(th)[:, None, None] * K + (1 - torch.cos(th))[:, None, None] * (K @ K)
    shift = 1
    while shift < F:
        R = torch.cat([R[:shift], R[shift:] @ R[:-shift]], dim=0)
        shift *= 2
    trans = torch.randn((F, 3), generator=g, **dd) * 5.0
    out = torch.empty((F, N, 3), device=device, dtype=torch.float32)
    for lo in range(0, F, chunk):
        hi = min(F, lo + chunk)
        x = base[None] + torch.einsum("fm,mnj->fnj", amp[lo:hi], modes)
        x = x + torch.randn(x.shape, generator=g, **dd) * 0.3
        x = torch.einsum("fij,fnj->fni", R[lo:hi], x) + trans[lo:hi, None, :]
        out[lo:hi] = x.to(torch.float32)
    return out
