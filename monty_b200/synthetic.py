"""Seed-defined synthetic inputs: per-stream sampler parameters and
density observations from a counter hash, identical wherever they are made --
numpy on the host (the parity tests, the CPU baseline) or torch on the device
(the full-size density workload, which would not fit a host round trip).

    h(i, k) = int_to_real<double>(splitmix64(0x9E3779B97F4A7C15 * (2 (8 i + k) + 1) ^ 0xC0FFEE))

Pure numpy / torch; nothing here touches the kernels or the oracle.
"""
import numpy as np


def counter_hash(i, k):
    """h(i, k) in (0, 1) for stream / observation index i (array) and slot k."""
    i = np.asarray(i, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(0x9E3779B97F4A7C15) * (np.uint64(2) * (i * np.uint64(8) + np.uint64(k)) + np.uint64(1))) ^ np.uint64(0xC0FFEE)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 1.1102230246251565e-16 + 1.1102230246251565e-16 / 2.0


def counter_hash_torch(idx, k):
    """The same hash on a torch int64 index tensor (two's complement == mod 2^64)."""
    import torch
    m64 = (1 << 64) - 1

    def c(v):
        v &= m64
        return v - (1 << 64) if v >= (1 << 63) else v
    z = (idx * 8 + k) * 2 + 1
    z = z * c(0x9E3779B97F4A7C15)
    z = z ^ c(0xC0FFEE)
    z = z + c(0x9E3779B97F4A7C15)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * c(0xBF58476D1CE4E5B9)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * c(0x94D049BB133111EB)
    z = z ^ ((z >> 31) & ((1 << 33) - 1))
    top = (z >> 11) & ((1 << 53) - 1)
    return top.to(torch.float64) * 1.1102230246251565e-16 + 1.1102230246251565e-16 / 2.0


N_PARAMS = {"real": 0, "uniform": 2, "exponential_rate": 1, "normal_ziggurat": 2,
            "normal_box_muller": 2, "normal_polar": 2, "binomial": 2, "poisson": 1,
            "gamma_scale": 2, "beta": 2, "negative_binomial_prob": 2,
            "hypergeometric": 3, "truncated_normal": 4}


def sampler_params(dist, n, first_stream=0, per_stream_normal=False):
    """Benchmark parameters of `dist` for streams [first_stream, first_stream + n):
    scalars where the BASELINE configs use scalars (C1 / C2), per-stream vectors
    spanning every regime of the sampler for C3 / C4."""
    idx = np.arange(first_stream, first_stream + n)
    h = lambda k: counter_hash(idx, k)  # noqa: E731
    if dist == "real":
        return []
    if dist in ("normal_ziggurat", "normal_box_muller", "normal_polar"):
        if per_stream_normal:
            return [h(0) * 10 - 5, 0.5 + h(1)]
        return [np.array([0.0]), np.array([1.0])]
    if dist == "uniform":
        return [np.array([0.0]), np.array([1.0])]
    if dist == "exponential_rate":
        return [np.array([1.0])]
    if dist == "binomial":
        return [np.floor(10 ** (6 * h(0))), h(1)]
    if dist == "poisson":
        lam = 10 ** (-1 + 5 * h(2))
        return [np.where(idx % 128 == 0, 1e8 * (1 + h(0)), lam)]
    if dist == "gamma_scale":
        return [10 ** (-1 + 2.5 * h(0)), 0.5 + h(1)]
    if dist == "beta":
        return [0.2 + 20 * h(0), 0.2 + 20 * h(1)]
    if dist == "negative_binomial_prob":
        return [0.5 + 100 * h(0), 0.05 + 0.9 * h(1)]
    if dist == "hypergeometric":
        n1 = np.floor(2000 * h(0))
        n2 = np.floor(2000 * h(1))
        return [n1, n2, np.floor((n1 + n2) * h(2))]
    if dist == "truncated_normal":
        mn = -3 + 5 * h(0)
        mx = mn + 0.1 + 4 * h(1)
        mx = np.where(idx % 8 == 1, np.inf, mx)
        mn = np.where(idx % 16 == 0, -np.inf, mn)
        return [np.zeros(n), np.ones(n), mn, mx]
    raise ValueError(dist)


DENSITY_BYTES_PER_OBS = {"poisson": 12, "negative_binomial_mu": 20, "beta_binomial_prob": 24}


def density_inputs_torch(name, n, first, dev, chunk=1 << 25):
    """Config C5 inputs for observations [first, first + n), generated on the
    device in chunks: x int32; poisson: lambda; negative_binomial_mu: size, mu;
    beta_binomial_prob: size (int32), prob, rho."""
    import torch
    x = torch.empty(n, dtype=torch.int32, device=dev)
    if name == "poisson":
        pars = [torch.empty(n, dtype=torch.float64, device=dev)]
    elif name == "negative_binomial_mu":
        pars = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(2)]
    elif name == "beta_binomial_prob":
        pars = [torch.empty(n, dtype=torch.int32, device=dev)] + \
               [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(2)]
    else:
        raise ValueError(name)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = torch.arange(first + lo, first + hi, dtype=torch.int64, device=dev)
        h = lambda k: counter_hash_torch(idx, k)  # noqa: E731
        xs = torch.floor(200 * h(0) ** 2).to(torch.int32)
        x[lo:hi] = xs
        if name == "poisson":
            pars[0][lo:hi] = 0.5 + 100 * h(1)
        elif name == "negative_binomial_mu":
            pars[0][lo:hi] = 0.5 + 50 * h(1)
            pars[1][lo:hi] = 0.1 + 100 * h(2)
        else:
            pars[0][lo:hi] = (xs + torch.floor(100 * h(1)).to(torch.int32)).to(torch.int32)
            pars[1][lo:hi] = 0.05 + 0.9 * h(2)
            pars[2][lo:hi] = 0.01 + 0.49 * h(3)
    return x, pars


def density_inputs_numpy(name, n, first=0):
    """The same inputs on the host (for the oracle / CPU baseline)."""
    idx = np.arange(first, first + n)
    h = lambda k: counter_hash(idx, k)  # noqa: E731
    x = np.floor(200 * h(0) ** 2).astype(np.int32)
    if name == "poisson":
        return x, [0.5 + 100 * h(1)]
    if name == "negative_binomial_mu":
        return x, [0.5 + 50 * h(1), 0.1 + 100 * h(2)]
    if name == "beta_binomial_prob":
        size = (x + np.floor(100 * h(1)).astype(np.int32)).astype(np.int32)
        return x, [size, 0.05 + 0.9 * h(2), 0.01 + 0.49 * h(3)]
    raise ValueError(name)
