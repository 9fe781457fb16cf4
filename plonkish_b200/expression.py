"""Host-side mirror of the reference's `Expression` and the compiler that turns one into GPU sum-check tables.

Mirrors /root/reference/plonkish_backend/src/util/expression.rs:
  Rotation, Query                       expression.rs:13-58  (Query orders by (poly, rotation), as #[derive(Ord)] does)
  CommonPolynomial                      expression.rs:60-65  (Identity, Lagrange(i), EqXY(idx))
  Expression + evaluate / degree /
  used_* / distribute_powers            expression.rs:67-244
  operators (Sum, Product, Scaled, ..)  expression.rs:488-560

and, for the GPU, restates what `ProverState` + `EvaluationsProver` do with an expression
(piop/sum_check/classic.rs:40-141, classic/eval.rs): the reference walks the expression tree per hypercube row and keeps
identity / Lagrange / eq_xy polynomials, constants and rotated queries implicit.  The GPU round kernel
(csrc/sumcheck_kernels.cuh) wants sum_t coeff_t * prod_j table[fac_t,j] (times one common table).  `compile_expression`
rewrites the tree into that form WITHOUT expanding linear factors: every maximal sub-expression of degree <= 1, e.g.
w + beta * (offset + id) + gamma (backend/hyperplonk/preprocessor.rs:153-165), is one *atom*, materialised as one
table (`plonkish_cuda_fr_affine_table`).  An atom is a multilinear polynomial and fixing a variable commutes with
sums, so every round polynomial — hence every transcript byte — is the field element the reference computes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R = FR_MODULUS

# util/arithmetic/bh.rs:4-74
BH_PRIMITIVES = [1, 3, 7, 11, 19, 37, 67, 131, 285, 529, 1033, 2053, 4179, 8219, 16427, 32771, 65581, 131081, 262183, 524327, 1048585,
                 2097157, 4194307, 8388641, 16777243, 33554441, 67108935, 134217767, 268435465, 536870917, 1073741907, 2147483657]
BH_X_INVS = [0, 1, 3, 5, 9, 18, 33, 65, 142, 264, 516, 1026, 2089, 4109, 8213, 16385, 32790, 65540, 131091, 262163, 524292, 1048578,
             2097153, 4194320, 8388621, 16777220, 33554467, 67108883, 134217732, 268435458, 536870953, 1073741828]


class BooleanHypercube:
    """util/arithmetic/bh.rs:76-153."""

    def __init__(self, num_vars: int):
        assert num_vars < 32
        self.num_vars = num_vars
        self.primitive = BH_PRIMITIVES[num_vars]
        self.x_inv = BH_X_INVS[num_vars]

    def next(self, b: int) -> int:  # bh.rs:141-146
        b <<= 1
        return b ^ ((b >> self.num_vars) * self.primitive)

    def prev(self, b: int) -> int:  # bh.rs:149-153
        return (b >> 1) ^ ((b & 1) * self.x_inv)

    def rotate(self, b: int, rotation: int) -> int:  # bh.rs:104-121
        for _ in range(rotation):
            b = self.next(b)
        for _ in range(-rotation):
            b = self.prev(b)
        return b

    def iter(self) -> Iterable[int]:  # bh.rs:123-130
        yield 0
        b = 1
        for _ in range((1 << self.num_vars) - 1):
            yield b
            b = self.next(b)

    def nth(self, nth: int) -> int:
        """bh.iter().nth(nth) without walking the cycle: 0, then X^(nth-1) in GF(2)[X] / primitive."""
        assert 0 <= nth < 1 << self.num_vars
        if nth == 0:
            return 0
        k, prim = self.num_vars, self.primitive

        def mul(x: int, y: int) -> int:
            acc = 0
            for i in range(k):
                if (y >> i) & 1:
                    acc ^= x << i
            for i in range(2 * k - 2, k - 1, -1):
                if (acc >> i) & 1:
                    acc ^= prim << (i - k)
            return acc

        result, base, e = 1, self.next(1) if k else 1, nth - 1
        while e:
            if e & 1:
                result = mul(result, base)
            base = mul(base, base)
            e >>= 1
        return result


@dataclass(frozen=True, order=True)
class Query:
    """expression.rs:40-58."""
    poly: int
    rotation: int = 0


# CommonPolynomial (expression.rs:60-65) as tuples: ("identity",), ("lagrange", i), ("eq_xy", idx)
IDENTITY = ("identity",)


def lagrange_poly(i: int):
    return ("lagrange", i)


def eq_xy_poly(idx: int):
    return ("eq_xy", idx)


class Expression:
    """expression.rs:67-78.  kind in {constant, common, poly, challenge, negated, sum, product, scaled, distribute}."""

    __slots__ = ("kind", "a", "b")

    def __init__(self, kind: str, a=None, b=None):
        self.kind, self.a, self.b = kind, a, b

    def __eq__(self, other) -> bool:  # #[derive(PartialEq)]: structural
        if not isinstance(other, Expression) or self.kind != other.kind:
            return False
        if self.kind == "distribute":
            return len(self.a) == len(other.a) and all(x == y for x, y in zip(self.a, other.a)) and self.b == other.b
        return self.a == other.a and self.b == other.b

    __hash__ = None

    # -- constructors (expression.rs:80-106, 317-324)
    @staticmethod
    def constant(v: int) -> "Expression":
        return Expression("constant", v % R)

    @staticmethod
    def zero() -> "Expression":
        return Expression.constant(0)

    @staticmethod
    def one() -> "Expression":
        return Expression.constant(1)

    @staticmethod
    def identity() -> "Expression":
        return Expression("common", IDENTITY)

    @staticmethod
    def lagrange(i: int) -> "Expression":
        return Expression("common", lagrange_poly(i))

    @staticmethod
    def eq_xy(idx: int) -> "Expression":
        return Expression("common", eq_xy_poly(idx))

    @staticmethod
    def polynomial(poly: int, rotation: int = 0) -> "Expression":
        return Expression("poly", Query(poly, rotation))

    @staticmethod
    def challenge(idx: int) -> "Expression":
        return Expression("challenge", idx)

    @staticmethod
    def distribute_powers(exprs: Sequence["Expression"], base: "Expression") -> "Expression":
        exprs = list(exprs)
        assert exprs, "distribute_powers of nothing is unreachable!() in the reference"
        return exprs[0] if len(exprs) == 1 else Expression("distribute", exprs, base)

    # -- operators (expression.rs:488-534): Expression * F is Scaled, a - b is Sum(a, Negated(b))
    def __add__(self, rhs: "Expression") -> "Expression":
        return Expression("sum", self, rhs)

    def __sub__(self, rhs: "Expression") -> "Expression":
        return Expression("sum", self, -rhs)

    def __mul__(self, rhs) -> "Expression":
        if isinstance(rhs, Expression):
            return Expression("product", self, rhs)
        return Expression("scaled", self, int(rhs) % R)

    def __neg__(self) -> "Expression":
        return Expression("negated", self)

    # -- evaluate (expression.rs:108-169)
    def evaluate(self, constant: Callable, common_poly: Callable, poly: Callable, challenge: Callable, negated: Callable, sum: Callable,
                 product: Callable, scaled: Callable):
        ev = lambda e: e.evaluate(constant, common_poly, poly, challenge, negated, sum, product, scaled)  # noqa: E731
        k = self.kind
        if k == "constant":
            return constant(self.a)
        if k == "common":
            return common_poly(self.a)
        if k == "poly":
            return poly(self.a)
        if k == "challenge":
            return challenge(self.a)
        if k == "negated":
            return negated(ev(self.a))
        if k == "sum":
            a = ev(self.a)
            return sum(a, ev(self.b))
        if k == "product":
            a = ev(self.a)
            return product(a, ev(self.b))
        if k == "scaled":
            return scaled(ev(self.a), self.b)
        assert k == "distribute" and self.a
        if len(self.a) == 1:
            return ev(self.a[0])
        base = ev(self.b)
        acc, power = ev(self.a[0]), base                      # exprs[0] + base * exprs[1] + base^2 * exprs[2] + ...  (expression.rs:160-166)
        for e in self.a[1:]:
            acc = sum(acc, product(power, ev(e)))
            power = product(power, base)
        return acc

    def degree(self) -> int:  # expression.rs:171-182
        return self.evaluate(lambda _: 0, lambda _: 1, lambda _: 1, lambda _: 0, lambda a: a, max, lambda a, b: a + b, lambda a, _: a)

    def _used(self, common_poly: Callable, poly: Callable) -> set:  # used_primitive, expression.rs:228-244
        merge = lambda a, b: a | b  # noqa: E731
        return self.evaluate(lambda _: set(), lambda p: common_poly(p), lambda q: poly(q), lambda _: set(), lambda a: a, merge, merge, lambda a, _: a)

    def used_lagrange(self) -> List[int]:
        return sorted(self._used(lambda p: {p[1]} if p[0] == "lagrange" else set(), lambda _: set()))

    def used_query(self) -> List[Query]:
        return sorted(self._used(lambda _: set(), lambda q: {q}))

    def used_poly(self) -> List[int]:
        return sorted(self._used(lambda _: set(), lambda q: {q.poly}))

    def used_rotation(self) -> List[int]:
        return sorted(self._used(lambda _: set(), lambda q: {q.rotation}))

    def max_used_rotation_distance(self) -> int:
        return max([abs(r) for r in self.used_rotation()] or [0])

    def used_challenge(self) -> List[int]:
        merge = lambda a, b: a | b  # noqa: E731
        return sorted(self.evaluate(lambda _: set(), lambda _: set(), lambda _: set(), lambda c: {c}, lambda a: a, merge, merge, lambda a, _: a))

    def evaluate_field(self, common_poly: Callable, poly: Callable, challenges: Sequence[int]) -> int:
        """`evaluate` over canonical integers mod r (what piop/sum_check.rs:59-99 does with closures)."""
        return self.evaluate(lambda c: c % R, common_poly, poly, lambda i: challenges[i] % R, lambda a: -a % R, lambda a, b: (a + b) % R,
                             lambda a, b: a * b % R, lambda a, s: a * s % R)


# ===================================================================================== compiler
Leaf = Tuple  # ("poly", poly, rotation) | ("identity",) | ("lagrange", i) | ("eq_xy", idx)


class Affine:
    """constant + sum_leaf coeff * leaf, coefficients canonical integers; the zero coefficients are dropped."""

    __slots__ = ("const", "terms")

    def __init__(self, const: int = 0, terms: Optional[Dict[Leaf, int]] = None):
        self.const = const % R
        self.terms = {l: c % R for l, c in (terms or {}).items() if c % R}

    def is_const(self) -> bool:
        return not self.terms

    def is_leaf(self) -> bool:
        return self.const == 0 and len(self.terms) == 1 and next(iter(self.terms.values())) == 1

    def leaf(self) -> Leaf:
        return next(iter(self.terms))

    def scaled(self, s: int) -> "Affine":
        return Affine(self.const * s, {l: c * s for l, c in self.terms.items()})

    def plus(self, o: "Affine") -> "Affine":
        t = dict(self.terms)
        for l, c in o.terms.items():
            t[l] = (t.get(l, 0) + c) % R
        return Affine(self.const + o.const, t)

    def key(self) -> Tuple:
        return (self.const, tuple(sorted(self.terms.items())))

    def value(self, leaf_value: Callable[[Leaf], int]) -> int:
        return (self.const + sum(c * leaf_value(l) for l, c in self.terms.items())) % R


# A compiled node is either an Affine or a list of product terms [(coeff, [Affine, ...])] with non-constant factors.
def _as_terms(node) -> List[Tuple[int, List[Affine]]]:
    if isinstance(node, Affine):
        if node.is_const():
            return [(node.const, [])] if node.const else []
        return [(1, [node])]
    return node


def _scale(node, s: int):
    s %= R
    if isinstance(node, Affine):
        return node.scaled(s)
    return [(c * s % R, f) for c, f in node if c * s % R]


def _add(a, b):
    if isinstance(a, Affine) and isinstance(b, Affine):
        return a.plus(b)
    return _as_terms(a) + _as_terms(b)


def _mul(a, b):
    if isinstance(a, Affine) and a.is_const():
        return _scale(b, a.const)
    if isinstance(b, Affine) and b.is_const():
        return _scale(a, b.const)
    return [(ca * cb % R, fa + fb) for ca, fa in _as_terms(a) for cb, fb in _as_terms(b) if ca * cb % R]


@dataclass
class CompiledExpression:
    atoms: List[Affine]                       # table i of the sum check is atoms[i], materialised
    terms: List[Tuple[int, List[int]]]        # (canonical coefficient, atom indices): sum_t coeff_t * prod_j atoms[idx]
    common: int                               # atom multiplying the whole sum (the eq_xy of a zero check), or -1
    degree: int                               # of the round polynomials: max factors per term (+ 1 with a common atom)

    def common_eq_xy(self) -> Optional[int]:
        """idx when the common atom is exactly the leaf eq_xy(idx) — the zero check's factor (preprocessor.rs:49-50),
        which lets the rounds run factored (sumcheck.zero_check_message) — else None."""
        if self.common < 0:
            return None
        a = self.atoms[self.common]
        return a.leaf()[1] if a.is_leaf() and a.leaf()[0] == "eq_xy" else None

    def value(self, leaf_value: Callable[[Leaf], int]) -> int:
        vals = [a.value(leaf_value) for a in self.atoms]
        total = 0
        for c, idx in self.terms:
            p = c
            for i in idx:
                p = p * vals[i] % R
            total = (total + p) % R
        return total * vals[self.common] % R if self.common >= 0 else total


def compile_expression(expression: Expression, challenges: Sequence[int]) -> CompiledExpression:
    """Flatten `expression` (challenges substituted) into product terms over affine atoms; see the module docstring."""
    node = expression.evaluate(
        lambda c: Affine(c),
        lambda p: Affine(0, {tuple(p): 1}),
        lambda q: Affine(0, {("poly", q.poly, q.rotation): 1}),
        lambda i: Affine(challenges[i]),
        lambda a: _scale(a, R - 1),
        _add,
        _mul,
        lambda a, s: _scale(a, s),
    )
    raw = _as_terms(node)
    # merge the terms over the same multiset of atoms
    merged: Dict[Tuple, Tuple[int, List[Affine]]] = {}
    for c, facs in raw:
        facs = sorted(facs, key=lambda a: a.key())
        k = tuple(a.key() for a in facs)
        merged[k] = ((merged[k][0] + c) % R if k in merged else c % R, facs)
    terms = [(c, f) for c, f in merged.values() if c]
    # a leaf present in every term becomes the common factor (the eq_xy(0) of the zero check, preprocessor.rs:49-50)
    common_atom: Optional[Affine] = None
    if terms and all(f for _, f in terms):
        shared = set(a.key() for a in terms[0][1] if a.is_leaf())
        for _, f in terms[1:]:
            shared &= set(a.key() for a in f)
        eq_first = sorted(shared, key=lambda k: (k[1][0][0][0] != "eq_xy", k))
        if eq_first:
            ck = eq_first[0]
            common_atom = next(a for a in terms[0][1] if a.key() == ck)
            stripped = []
            for c, f in terms:
                f = list(f)
                f.pop(next(i for i, a in enumerate(f) if a.key() == ck))
                stripped.append((c, f))
            terms = stripped
    # a coefficient other than one goes into a factor that is materialised anyway (not a plain table)
    absorbed = []
    for c, f in terms:
        if c != 1:
            j = next((i for i, a in enumerate(f) if not a.is_leaf()), None)
            if j is not None:
                f = list(f)
                f[j] = f[j].scaled(c)
                c = 1
        absorbed.append((c, f))
    atoms: List[Affine] = []
    index: Dict[Tuple, int] = {}

    def atom_id(a: Affine) -> int:
        k = a.key()
        if k not in index:
            index[k] = len(atoms)
            atoms.append(a)
        return index[k]

    common = atom_id(common_atom) if common_atom is not None else -1
    out_terms = [(c, [atom_id(a) for a in f]) for c, f in absorbed]
    degree = max([len(f) for _, f in out_terms] or [0]) + (1 if common >= 0 else 0)
    return CompiledExpression(atoms, out_terms, common, degree)
