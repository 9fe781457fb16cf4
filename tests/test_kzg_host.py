"""CPU suite: host-side logic of the MultilinearKzg mirror (plonkish_b200/kzg.py) — the
`quotients` bookkeeping of pcs/multilinear.rs:72-107 and the Montgomery conversions.
The MSMs themselves need a GPU (tests/test_msm_gpu.py::test_multilinear_kzg_commit_open_round_trip)."""
import random

from plonkish_b200 import kzg


def _evaluate(evals, x):
    r = kzg.FR_MODULUS
    cur = list(evals)
    for xi in reversed(x):
        h = len(cur) // 2
        cur = [(cur[j] + (cur[j + h] - cur[j]) * xi) % r for j in range(h)]
    return cur[0]


def test_quotients_shapes_and_remainder():
    rnd = random.Random(3)
    r = kzg.FR_MODULUS
    for k in (1, 2, 5, 8):
        evals = [rnd.randrange(r) for _ in range(1 << k)]
        x = [rnd.randrange(r) for _ in range(k)]
        qs, value = kzg.quotients(evals, x)
        assert [len(q) for q in qs] == [1 << i for i in range(k)]  # MSM sizes 1, 2, ..., 2^(k-1)
        assert value == _evaluate(evals, x)
        # f(X) - f(x) = sum_i (X_i - x_i) * q_i(X_0..X_{i-1}) at a random X
        big_x = [rnd.randrange(r) for _ in range(k)]
        total = sum((big_x[i] - x[i]) * _evaluate(qs[i], big_x[:i]) for i in range(k)) % r
        assert total == (_evaluate(evals, big_x) - value) % r


def test_montgomery_round_trip():
    rnd = random.Random(4)
    vals = [0, 1, kzg.FR_MODULUS - 1] + [rnd.randrange(kzg.FR_MODULUS) for _ in range(20)]
    limbs = kzg.fr_to_montgomery(vals)
    assert limbs.shape == (len(vals), 4)
    assert kzg.fr_from_montgomery(limbs) == vals
    # 1 in Montgomery form is R mod r
    assert int.from_bytes(limbs[1].tobytes(), "little") == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
