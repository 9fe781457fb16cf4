"""Generates tests/golden/msm_vectors.json with oracle/bigint_ref.py (pure Python
integers, affine chord-and-tangent arithmetic) — independent of the C oracle and
of the CUDA path.  Run from the repo root:  python tests/golden/make_golden.py

The reference (amit0365/plonkish) holds no known-answer vectors for msm.rs
 and cannot be executed here (Rust, no toolchain), so these
vectors pin the *mathematical* value its variable_base_msm returns after the
callers' to_affine(); the only external anchors are the public BN254 constants
and the EIP-196 value of 2*G recorded in `public_kats`.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bigint_ref as br  # noqa: E402

rnd = random.Random(20261018)


def rand_point():
    return br.scalar_mul(rnd.randrange(1, br.R), br.G)


def case(name, scalars, points):
    result = br.msm(scalars, points)
    assert br.is_on_curve(result)
    return {
        "name": name,
        "scalars_mont_le": [br.scalar_to_bytes(k).hex() for k in scalars],
        "bases_mont_le": [br.point_to_bytes(p).hex() for p in points],
        "result_mont_le": br.point_to_bytes(result).hex(),
        "result_transcript_be": None if result is None else br.transcript_bytes(result).hex(),
    }


cases = []
for n in (1, 2, 3, 17, 33, 64):
    cases.append(case(f"random_n{n}", [rnd.randrange(br.R) for _ in range(n)], [rand_point() for _ in range(n)]))

P1, P2, P3 = rand_point(), rand_point(), rand_point()
edge = br.R - 1
cases += [
    case("all_zero_scalars", [0, 0, 0], [P1, P2, P3]),
    case("scalar_one", [1], [P1]),
    case("scalar_minus_one", [edge], [P1]),
    case("scalar_2pow253", [1 << 253], [P1]),
    case("scalar_half_boundaries", [(br.R - 1) // 2, (br.R + 1) // 2, (br.R - 1) // 2 - 1], [P1, P2, P3]),
    case("window_boundaries", [(1 << 15), (1 << 15) + 1, (1 << 16) - 1, (1 << 16), br.R - (1 << 15), (1 << 127) - 1, (1 << 128)],
         [P1, P2, P3, P1, P2, P3, P1]),
    case("duplicate_bases_equal_scalars", [5, 5, 5, 5], [P1, P1, P1, P1]),
    case("duplicate_bases_random_scalars", [rnd.randrange(br.R) for _ in range(8)], [P2] * 8),
    case("p_and_minus_p_cancel", [7, 7], [P1, br.neg(P1)]),
    case("cancel_to_identity_with_scalars", [9, br.R - 9], [P3, P3]),
    case("identity_bases_mixed", [rnd.randrange(br.R) for _ in range(5)], [P1, None, P2, None, P3]),
    case("all_identity_bases", [3, 4], [None, None]),
    case("protostar_fold_shape", [1, rnd.randrange(br.R)], [None, P2]),  # accumulation/protostar.rs:270
    case("small_integer_scalars", [rnd.randrange(3 * 64) for _ in range(40)], [rand_point() for _ in range(40)]),
    case("selector_like_0_1_minus1", [rnd.choice([0, 1, edge]) for _ in range(40)], [rand_point() for _ in range(40)]),
]

doc = {
    "generator": "tests/golden/make_golden.py (oracle/bigint_ref.py, Python integers)",
    "encoding": "scalars: 32-byte LE Montgomery Fr; bases/result: x||y 32-byte LE Montgomery Fq, zeros = identity; "
                "transcript: x||y 32-byte BE canonical (util/transcript.rs:216-229)",
    "public_kats": {
        "generator": br.point_to_bytes(br.G).hex(),
        "two_g_canonical_be": br.transcript_bytes(br.TWO_G).hex(),
        "two_g_source": "EIP-196 / alt_bn128 test vectors",
    },
    "cases": cases,
}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "msm_vectors.json")
with open(out, "w") as f:
    json.dump(doc, f, indent=1)
print(f"wrote {len(cases)} cases to {out}")
