"""synth.py — seeded synthetic inputs for the hot path.

Minimal NumPy restatements of the reference's *geometry providers* — only their
3 x N / 3 x I outputs cross the DAS boundary:

    TransducerArray.positions      src/TransducerArray.m:95-99
    TransducerMatrix.positions     src/TransducerMatrix.m:130-150
    ScanCartesian.getImagingGrid   src/ScanCartesian.m:126-145   (order 'ZXY' => I1 = z fastest, :11)
    UltrasoundSystem.DAS pos_args  src/UltrasoundSystem.m:3341-3351
    Transducer.xdcImpulse          src/Transducer.m:901-925, :1124-1127 (complex Gaussian pulse)
    Waveform.conv / .time          src/Waveform.m:384-433, :482-486

plus the BASELINE.json configs C1..C5 as named presets.  No files, no network.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

C0 = 1540.0


def linear_array(numel: int, pitch: float) -> np.ndarray:
    """3 x N element positions of a TransducerArray (src/TransducerArray.m:95-99)."""
    w = (numel - 1) * pitch
    p = np.zeros((3, numel))
    p[0] = np.linspace(-w / 2, w / 2, numel)
    return p


def matrix_array(nx: int, ny: int, pitch: float) -> np.ndarray:
    """3 x (nx*ny) element positions of a TransducerMatrix, x fastest (src/TransducerMatrix.m:130-150)."""
    x = np.linspace(-(nx - 1) * pitch / 2, (nx - 1) * pitch / 2, nx)
    y = np.linspace(-(ny - 1) * pitch / 2, (ny - 1) * pitch / 2, ny)
    X, Y = np.meshgrid(x, y, indexing="ij")
    p = np.zeros((3, nx * ny))
    p[0], p[1] = X.reshape(-1, order="F"), Y.reshape(-1, order="F")
    return p


def scan_cartesian(x, z, y=(0.0,)) -> np.ndarray:
    """3 x I1 x I2 x I3 pixel positions, order 'ZXY' (src/ScanCartesian.m:126-145)."""
    x, y, z = (np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (x, y, z))
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    return np.stack([X, Y, Z], 0)


def noise_cube(T, N, M, F=1, seed=0, dtype=np.complex64, pad=4) -> np.ndarray:
    """x = (randn + i randn)/sqrt(2), first/last `pad` samples zeroed."""
    rng = np.random.Generator(np.random.PCG64(seed))
    shape = (T, N, M) if F == 1 else (T, N, M, F)
    x = np.empty(shape, dtype=dtype, order="F")
    # generate trace-by-trace blocks to bound temporary memory at the 1 GB headline size
    flat = x.reshape(T, -1, order="F")
    step = 4096
    for j in range(0, flat.shape[1], step):
        k = min(step, flat.shape[1] - j)
        blk = rng.standard_normal((k, T, 2), dtype=np.float32) * np.float32(np.sqrt(0.5))
        flat[:, j:j + k] = (blk[..., 0] + 1j * blk[..., 1]).T
    if pad:
        flat[:pad] = 0
        flat[T - pad:] = 0
    return x


def gauspuls_cutoff(fc, bw, bwr=-6.0, tpr=-80.0) -> float:
    """MATLAB gauspuls('cutoff', fc, bw, bwr, tpr) as used at src/Transducer.m:917."""
    r = 10.0 ** (bwr / 20.0)
    fv = -(bw * bw * fc * fc) / (8.0 * np.log(r))
    tv = 1.0 / (4.0 * np.pi * np.pi * fv)
    delta = 10.0 ** (tpr / 20.0)
    return float(np.sqrt(-2.0 * tv * np.log(delta)))


def xdc_impulse(fc, bw_frac, bwr=-6.0):
    """Complex Gaussian pulse of Transducer.cgauspulsfun (src/Transducer.m:1124-1127) and its cutoff tc."""
    isig = (4 * np.pi * np.pi * (-bw_frac * bw_frac * fc * fc / (8 * np.log(10.0 ** (bwr / 20.0))))) / 2
    tc = gauspuls_cutoff(fc, bw_frac, bwr, -80.0)
    return (lambda t: np.exp(-t * t * isig) * np.exp(2j * np.pi * fc * t)), tc


def greens_kernel(fc, bw_frac, fs):
    """kern = samples of conv(rx.impulse, conv(tx.impulse, delta)) at fs, and wv.t0 (src/UltrasoundSystem.m:584-588).

    Waveform.conv resamples both at fs and contracts (src/Waveform.m:425-427); .time = floor(t0 fs):ceil(tend fs).
    """
    f, tc = xdc_impulse(fc, bw_frac)
    t0, tend = -2 * tc, 2 * tc
    k = np.arange(np.floor(t0 * fs), np.ceil(tend * fs) + 1) / fs
    t = np.arange(np.floor(t0 * fs), np.ceil(tend * fs) + 1) / fs

    def samp(tt):  # Waveform.sample: zero outside [t0, tend] of the single impulse
        return np.where((tt >= -tc) & (tt <= tc), f(tt), 0)

    kern = samp(t[:, None] - k[None, :]) @ samp(k)
    return kern.astype(np.complex128), float(t0), float(tend)


@dataclass
class DasProblem:
    """Positional/option arguments of one das_spec call (kern/das_spec.m:1)."""
    name: str
    Pi: np.ndarray
    Pr: np.ndarray
    Pv: np.ndarray
    Nv: np.ndarray
    T: int
    fs: float
    t0: float
    c0: float
    opts: tuple = ()
    interp: str = "cubic"
    meta: dict = field(default_factory=dict)

    @property
    def N(self): return self.Pr.shape[1]
    @property
    def M(self): return max(self.Pv.shape[1], self.Nv.shape[1])
    @property
    def Isz(self): return tuple(self.Pi.shape[1:])
    @property
    def I(self): return int(np.prod(self.Isz))

    def args(self, x, dtype=np.float32):
        """(Pi, Pr, Pv, Nv, x, t0, fs, c, *opts) ready for das_spec."""
        f = lambda a: np.asarray(a, dtype=dtype)
        return (f(self.Pi), f(self.Pr), f(self.Pv), f(self.Nv), x, dtype(self.t0), dtype(self.fs), dtype(self.c0),
                *self.opts, "interp", self.interp)

    def bytes_alg(self, taps=None, Bs=8) -> int:
        """Algorithmic bytes of SURVEY.md §8d: I*N*M*k*B_s + I*B_s."""
        k = taps or {"nearest": 1, "linear": 2, "cubic": 4, "lanczos3": 4}[self.interp]
        return self.I * self.N * self.M * k * Bs + self.I * Bs


def config_c1(nz=128, nx=128, T=2048) -> DasProblem:
    """C1: L11-5v (128 el, pitch 0.3 mm, fc 7.25 MHz), 1 plane wave, 128x128, linear (src/TransducerArray.m:339-349)."""
    fc = 7.25e6
    Pr = linear_array(128, 0.3e-3)
    Pi = scan_cartesian(np.linspace(-19.05e-3, 19.05e-3, nx), np.linspace(5e-3, 45e-3, nz))
    return DasProblem("C1", Pi, Pr, np.zeros((3, 1)), np.array([[0.0], [0.0], [1.0]]), T, 4 * fc, 0.0, C0,
                      ("plane-waves",), "linear", {"fc": fc})


def config_c2(nz=1024, nx=1024, N=256, M=256, T=2048, interp="cubic") -> DasProblem:
    """C2 headline: 256-el L12-5v geometry, 256 focused transmits at z = 30 mm, 1024x1024, cubic fp32."""
    fc = 7.5e6
    pitch = float(np.frombuffer(bytes.fromhex("3f29992e39cf2ea7"), dtype=">f8")[0])  # src/TransducerArray.m:371
    Pr = linear_array(N, pitch * 256 / N if N != 256 else pitch)
    xe = linear_array(M, pitch * 256 / M if M != 256 else pitch)[0]
    Pv = np.stack([xe, np.zeros(M), np.full(M, 30e-3)], 0)   # foci [x_el(m); 0; 30 mm]
    Nv = Pv / np.linalg.norm(Pv, 2)                          # nf ./ norm(nf): matrix 2-norm (src/UltrasoundSystem.m:3349-3350)
    Pi = scan_cartesian(np.linspace(-25e-3, 25e-3, nx), np.linspace(1e-3, 51e-3, nz))
    return DasProblem("C2", Pi, Pr, Pv, Nv, T, 4 * fc, 0.0, C0, (), interp, {"fc": fc})


def config_c3(nz=512, nx=512, N=192, M=128, T=2048, interp="cubic") -> DasProblem:
    """C3: L12-3v (192 el, pitch 0.2 mm), 128 plane waves -16..16 deg, 512x512, baseband IQ."""
    fc = 7.5e6
    Pr = linear_array(N, 0.2e-3)
    th = np.deg2rad(np.linspace(-16, 16, M))
    Nv = np.stack([np.sin(th), np.zeros(M), np.cos(th)], 0)  # src/SequenceRadial.m:150
    Pi = scan_cartesian(np.linspace(-19e-3, 19e-3, nx), np.linspace(2e-3, 40e-3, nz))
    return DasProblem("C3", Pi, Pr, np.zeros((3, 1)), Nv, T, 4 * fc, 0.0, C0, ("plane-waves", "modulation", fc),
                      interp, {"fc": fc})


def config_c4(n=256, T=2048, interp="cubic") -> DasProblem:
    """C4: PO192O 32x32 matrix (pitch 0.3 mm), 64 diverging waves (8x8 virtual sources at z = -10 mm), n^3 voxels."""
    fc = 3.47e6
    Pr = matrix_array(32, 32, 0.3e-3)
    g = np.linspace(-4e-3, 4e-3, 8)
    X, Y = np.meshgrid(g, g, indexing="ij")
    Pv = np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F"), np.full(64, -10e-3)], 0)
    Nv = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, 64))
    ax = np.linspace(-5e-3, 5e-3, n)
    Pi = scan_cartesian(ax, np.linspace(5e-3, 45e-3, n), ax)
    return DasProblem("C4", Pi, Pr, Pv, Nv, T, 4 * fc, 0.0, C0, ("diverging-waves",), interp, {"fc": fc})


def config_c5_das(nz=1024, nx=1024, N=256, T=2048) -> DasProblem:
    """C5 DAS leg: FSA (each element transmits), 'diverging-waves' (src/UltrasoundSystem.m:3341-3344)."""
    p = config_c2(nz, nx, N, N, T)
    Pv = p.Pr.copy()
    Nv = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, N))
    return DasProblem("C5", p.Pi, p.Pr, Pv, Nv, T, p.fs, 0.0, C0, ("diverging-waves",), "cubic", dict(p.meta))
