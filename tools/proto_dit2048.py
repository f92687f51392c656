#!/usr/bin/env python
"""Executable specification of csrc/zplt_fft2048_kernels.cu: the same index arithmetic (slot ownership, the two exchanges
through an 8-byte image with pencil-local swizzle, the slot permutation bo, the final decimation-in-time combine), thread by
thread in numpy, checked against numpy.fft.  Run: python tools/proto_dit2048.py"""
import numpy as np

M, R3 = 64, 4
NT = 16


def at(a):
    return a ^ ((a >> 2) & 1)


def dft(v, n):  # unnormalised backward DFT of length n along the last axis
    return np.fft.ifft(v, axis=-1) * n


def fft1024_split(x):
    """x: 1024 complex values of one pencil.  Returns (X as owned by the threads, bo per slot): regs[b][e] = X[bo(b) + 64 e]."""
    N = 1024
    W = np.exp(2j * np.pi * np.arange(N) / N)
    regs = np.array([[x[b + M * e] for e in range(16)] for b in range(M)])  # regs[b][e]
    # pass 1: radix 16 over stride M, twiddle W^(b k)
    regs = dft(regs, 16)
    for b in range(M):
        regs[b] *= W[(b * np.arange(16)) % N]
    S = np.zeros(N + 2)
    new = np.zeros_like(regs)
    for part in ("real", "imag"):
        S[:] = np.nan
        for b in range(M):
            for k in range(16):
                S[at(k * M + b)] = getattr(regs[b, k], part)
        for b in range(M):
            k1, i = divmod(b, R3)
            base = k1 * 16 * R3
            vals = np.array([S[at(base + n * R3 + i)] for n in range(16)])
            assert not np.isnan(vals).any()
            new[b] = new[b] + (vals if part == "real" else 1j * vals)
    regs = new
    # pass 2: radix 16 over stride R3, twiddle W^(16 i k)
    regs = dft(regs, 16)
    for b in range(M):
        i = b % R3
        regs[b] *= W[(16 * i * np.arange(16)) % N]
    new = np.zeros_like(regs)
    for part in ("real", "imag"):
        S[:] = np.nan
        for b in range(M):
            k1, i = divmod(b, R3)
            base = k1 * 16 * R3
            for k in range(16):
                S[at(base + k * R3 + i)] = getattr(regs[b, k], part)
        for b in range(M):
            k1, i = divmod(b, R3)
            base = k1 * 16 * R3
            vals = np.array([S[at(base + (i + R3 * j) * R3 + n)] for j in range(16 // R3) for n in range(R3)])
            assert not np.isnan(vals).any()
            new[b] = new[b] + (vals if part == "real" else 1j * vals)
    regs = new
    # pass 3: radix R3 butterflies on consecutive groups, then slot order
    out = np.zeros_like(regs)
    NB = 16 // R3
    for b in range(M):
        g = dft(regs[b].reshape(NB, R3), R3)  # group j = regs[j*R3 .. j*R3+R3-1]
        for j in range(NB):
            for k in range(R3):
                out[b, j + NB * k] = g[j, k]
    bo = np.array([(b // R3) + 16 * (b % R3) for b in range(M)])
    return out, bo


def main():
    rng = np.random.RandomState(1)
    # the 1024-point core
    x = rng.standard_normal(1024) + 1j * rng.standard_normal(1024)
    regs, bo = fft1024_split(x)
    want = dft(x, 1024)
    got = np.zeros(1024, dtype=complex)
    for b in range(M):
        for e in range(16):
            got[bo[b] + M * e] = regs[b, e]
    err = np.abs(got - want).max() / np.abs(want).max()
    print("1024-point split-exchange core: max relative error", err)
    assert err < 1e-12 and sorted(bo) == list(range(M))
    # the 2048-point decimation in time
    x = rng.standard_normal(2048) + 1j * rng.standard_normal(2048)
    E, bo = fft1024_split(x[0::2])
    O, bo2 = fft1024_split(x[1::2])
    assert np.array_equal(bo, bo2)
    W2 = np.exp(2j * np.pi * np.arange(2048) / 2048)
    X = np.zeros(2048, dtype=complex)
    for b in range(M):
        for e in range(16):
            k = bo[b] + M * e
            t = O[b, e] * W2[k]
            X[k] = E[b, e] + t
            X[k + 1024] = E[b, e] - t
    want = dft(x, 2048)
    err = np.abs(X - want).max() / np.abs(want).max()
    print("2048-point decimation in time: max relative error", err)
    assert err < 1e-12


if __name__ == "__main__":
    main()
