"""Seeded synthetic baseband generators shared by tests, smoke() and bench.py."""
import numpy as np

SEED = 0xD5B5


def caspsr_bytes(ndat, seed=SEED, sigma=20.0):
    """CASPSR 8-bit real dual-pol: byte 8*(i/4) + 4*p + i%4 (two's complement int8)."""
    rng = np.random.default_rng(seed)
    x = np.clip(np.rint(rng.standard_normal((2, ndat)) * sigma), -128, 127).astype(np.int8)
    raw = np.empty((ndat // 4, 2, 4), np.int8)
    raw[:, 0, :] = x[0].reshape(-1, 4)
    raw[:, 1, :] = x[1].reshape(-1, 4)
    return raw.reshape(-1).view(np.uint8)


def generic8_bytes(ndat, nchan, npol, ndim, seed=SEED, sigma=20.0):
    rng = np.random.default_rng(seed)
    x = np.clip(np.rint(rng.standard_normal(ndat * nchan * npol * ndim) * sigma), -128, 127).astype(np.int8)
    return x.view(np.uint8)


def meerkat_bytes(ndat, nchan, npol, seed=SEED, sigma=20.0):
    rng = np.random.default_rng(seed)
    x = np.clip(np.rint(rng.standard_normal(ndat * nchan * npol * 2) * sigma), -128, 127).astype(np.int8)
    return x.view(np.uint8)


def uwb_bytes(ndat, npol, seed=SEED, sigma=2000.0):
    rng = np.random.default_rng(seed)
    x = np.clip(np.rint(rng.standard_normal(ndat * npol * 2) * sigma), -32768, 32767).astype(np.int16)
    return (x.view(np.uint16) ^ np.uint16(0x8000)).view(np.uint8)


def relerr(a, b):
    """max |a-b| normalised to the RMS of b (north_star's parity metric)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    rms = np.sqrt(np.mean(b * b))
    return float(np.max(np.abs(a - b)) / rms) if rms > 0 else float(np.max(np.abs(a - b)))


def twobit_bytes(ndat, npol=2, seed=0, sigma=1.0, threshold=0.9674):
    """CPSR2-style 2-bit OffsetBinary stream: Gaussian noise digitised at +-threshold*sigma_nominal
    (codes 0..3 = -hi, -lo, +lo, +hi), four samples per byte most-significant first, polarisations
    interleaved byte by byte.  `sigma` scales the input power (1 = nominal)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((npol, ndat)) * sigma
    code = np.where(x < -threshold, 0, np.where(x < 0, 1, np.where(x < threshold, 2, 3))).astype(np.uint8)
    c = code.reshape(npol, ndat // 4, 4)
    byte = (c[:, :, 0] << 6) | (c[:, :, 1] << 4) | (c[:, :, 2] << 2) | c[:, :, 3]
    return np.ascontiguousarray(byte.T).reshape(-1).astype(np.uint8)     # [ndat/4][npol]
