"""CPU suite: permutation_z_polys (backend/hyperplonk/prover.rs:252-345) and BooleanHypercube (util/arithmetic/bh.rs).
The oracle's restatement is pinned by a direct Python-integer computation and by the argument's own identity (for a
satisfied copy constraint the last running product times the last row's factor is one, prover.rs:322-328); the product's
kernels (csrc/poly_kernels.cuh k_perm_*, k_prodscan_*), compiled by g++ for the emulator, are compared with the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bigint_ref as br

R = br.R
MONT = 1 << 256
RINV = pow(MONT, -1, R)
PRIMITIVES = [1, 3, 7, 11, 19, 37, 67, 131, 285, 529, 1033, 2053, 4179, 8219, 16427, 32771, 65581]  # bh.rs:5-22


def to_mont(vals):
    return np.frombuffer(b"".join((int(v) % R * MONT % R).to_bytes(32, "little") for v in vals), dtype=np.uint64).reshape(-1, 4).copy()


def to_ints(arr):
    return [int.from_bytes(row.tobytes(), "little") * RINV % R for row in np.asarray(arr).reshape(-1, 4)]


def bh_iter(k):
    out, b = [0], 1
    for _ in range((1 << k) - 1):
        out.append(b)
        b <<= 1
        b ^= (b >> k) * PRIMITIVES[k]
    return out


def z_polys_python(num_chunks, values, sigmas, beta, gamma, k):
    """prover.rs:252-345 on Python integers."""
    n, count = 1 << k, len(values)
    chunk_size = -(-count // num_chunks)
    products = []
    for c in range(num_chunks):
        idxs = range(c * chunk_size, min((c + 1) * chunk_size, count))
        prod = []
        for b in range(n):
            den = num = 1
            for i in idxs:
                den = den * (beta * sigmas[i][b] + gamma + values[i][b]) % R
                num = num * (beta * ((i << k) + b) + gamma + values[i][b]) % R
            prod.append(num * pow(den, -1, R) % R)
        products.append(prod)
    order = bh_iter(k)
    z = [0] * num_chunks + [1]
    state = 1
    for nth in range(1, n):
        for c in range(num_chunks):
            state = state * products[c][order[nth]] % R
            z.append(state)
    z = z[: num_chunks << k]
    polys = [[0] * n for _ in range(num_chunks)]
    for nth, b in enumerate(order):
        for c in range(num_chunks):
            polys[c][b] = z[c + num_chunks * nth]
    return polys, products, order


def copy_constrained_instance(k, count, rng):
    """Witness columns and permutation polynomials of a satisfied copy constraint: cells are partitioned into cycles, every
    cell of a cycle holds the same value, sigma maps a cell to the next cell of its cycle (ids are (column << k) + row)."""
    n = 1 << k
    # row 0 stays out of every cycle (its cells are fixed points): BooleanHypercube's walk starts after it (bh.rs:118-125),
    # so the grand product only closes over the other rows
    cells = np.array([c for c in rng.permutation(count * n) if c % n != 0], dtype=np.int64)
    values = np.zeros(count * n, dtype=object)
    for c in range(count):
        values[c * n] = int(rng.integers(0, 1 << 62))
    sigma = np.arange(count * n, dtype=object)
    pos = 0
    while pos < len(cells):
        ln = int(rng.integers(1, 5))
        cyc = cells[pos:pos + ln]
        v = int(rng.integers(0, 1 << 62)) * int(rng.integers(1, 1 << 62)) % R
        for j, cell in enumerate(cyc):
            values[cell] = v
            sigma[cell] = int(cyc[(j + 1) % len(cyc)])
        pos += ln
    vals = [[int(values[i * n + b]) for b in range(n)] for i in range(count)]
    sigs = [[int(sigma[i * n + b]) for b in range(n)] for i in range(count)]
    return vals, sigs


def test_boolean_hypercube_order_visits_every_row_once(oracle):
    for k in range(1, 15):
        got = oracle.bh_iter(k)
        assert got.tolist() == bh_iter(k)
        assert sorted(got.tolist()) == list(range(1 << k))


@pytest.mark.parametrize("k,count,num_chunks", [(1, 1, 1), (3, 3, 1), (5, 3, 1), (6, 3, 3), (6, 4, 2), (8, 5, 2)])
def test_oracle_z_polys_match_python_integers_and_close_the_product(oracle, k, count, num_chunks):
    rng = np.random.default_rng(100 * k + count)
    vals, sigs = copy_constrained_instance(k, count, rng)
    beta, gamma = 0x1234567 * 0x89ABCDEF % R, (R - 12345)
    want, products, order = z_polys_python(num_chunks, vals, sigs, beta, gamma, k)
    got = oracle.permutation_z_polys(num_chunks, [to_mont(v) for v in vals], [to_mont(s) for s in sigs], to_mont([beta])[0], to_mont([gamma])[0])
    assert [to_ints(g) for g in got] == want
    # the argument's identity (prover.rs:322-328): the last running product times the last row's factor is one
    last = want[num_chunks - 1][order[-1]]
    assert last * products[num_chunks - 1][order[-1]] % R == 1
    assert all(p[0] == 0 for p in want) and want[0][1] == 1


@pytest.fixture(scope="module")
def emul():
    emul_dir = os.path.join(ROOT, "tests", "emul")
    subprocess.run(["make", "-C", emul_dir], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(emul_dir, "libemul_msm.so"))
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    lib.emul_permutation_z.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp]
    return lib


@pytest.mark.parametrize("k,count,num_chunks", [(1, 1, 1), (2, 2, 1), (5, 3, 1), (6, 4, 2), (9, 3, 1), (9, 3, 3), (10, 5, 2), (13, 3, 1), (14, 3, 3)])
def test_emulated_kernels_match_the_oracle(emul, oracle, k, count, num_chunks):
    n = 1 << k
    values = [oracle.random_scalars(n, 10 * k + i) for i in range(count)]
    sigmas = [oracle.random_scalars(n, 500 + 10 * k + i) for i in range(count)]
    beta, gamma = oracle.random_scalars(2, 77)
    want = oracle.permutation_z_polys(num_chunks, values, sigmas, beta, gamma)
    out = np.zeros((num_chunks * n, 4), dtype=np.uint64)
    vp_ = (ctypes.c_void_p * count)(*[v.ctypes.data for v in values])
    sp_ = (ctypes.c_void_p * count)(*[v.ctypes.data for v in sigmas])
    emul.emul_permutation_z(ctypes.cast(vp_, ctypes.c_void_p), ctypes.cast(sp_, ctypes.c_void_p), count, num_chunks, k, beta.ctypes.data, gamma.ctypes.data,
                            out.ctypes.data)
    for c in range(num_chunks):
        assert out[c * n:(c + 1) * n].tobytes() == want[c].tobytes(), c
