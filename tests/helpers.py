"""Shared helpers of the test-suite (not collected)."""
import numpy as np

from brutefir_b200.formats import unpack_block


def rng_cbuf(rng, n, dtype, scale=1.0):
    return (rng.standard_normal(n) * scale).astype(dtype)


def unpack_run(raw_blocks, bfs, L):
    """uint8[nb, bytes] -> float64[ch, nb*L]."""
    vals = np.stack([unpack_block(raw_blocks[b], bfs, L) for b in range(raw_blocks.shape[0])])
    return vals.transpose(1, 0, 2).reshape(len(bfs), -1)


def blocked_to_complex(r):
    """Reference blocked layout (fftw_convfuns.h:25-42) -> complex bins X_0..X_{N/2}."""
    n = r.size
    m = n // 2
    c = r.reshape(-1, 8)
    x = np.zeros(m + 1, np.complex128)
    x[:m] = (c[:, :4] + 1j * c[:, 4:]).reshape(-1)
    x[m] = r[4]
    x[0] = r[0]
    return x


def hc_to_complex(hc):
    n = hc.size
    m = n // 2
    x = np.zeros(m + 1, np.complex128)
    x[0], x[m] = hc[0], hc[m]
    x[1:m] = hc[1:m] + 1j * hc[n - 1:m:-1]
    return x


def ulp_tol(dtype, n_fft, magnitude):
    """Tolerance for comparing two independent FFTs of size n_fft in `dtype` on data of the given
    magnitude: a few eps * sqrt(log2 n) * magnitude."""
    eps = np.finfo(dtype).eps
    return 8.0 * eps * np.sqrt(np.log2(n_fft)) * magnitude
