"""The BASELINE einsums (and the reference's test einsums) built with the
feinsum_b200 front-end.  Shapes follow reference test/test_codegen.py:34-120,
test/test_measure.py:33-52, examples/wave_3d_p4_auto.py:16-63."""

import feinsum_b200 as f


def grad(dtype="float64", ndim=3, ndof=35):
    return f.einsum(
        "xre,rij,ej->xei",
        f.array("J", (ndim, ndim, "E"), dtype),
        f.array("D", (ndim, ndof, ndof), dtype),
        f.array("u", ("E", ndof), dtype),
    )


def grad_batched(b=3, dtype="float64"):
    return f.batched_einsum(
        "xre,rij,ej->xei",
        [[f.array("J", (3, 3, "E"), dtype), f.array("D", (3, 35, 35), dtype),
          f.array(f"u{k}", ("E", 35), dtype)] for k in range(b)],
    )


def div(dtype="float64", ndim=3, ndof=35):
    return f.einsum(
        "xre,rij,xej->ei",
        f.array("J", (ndim, ndim, "E"), dtype),
        f.array("D", (ndim, ndof, ndof), dtype),
        f.array("u", (ndim, "E", ndof), dtype),
    )


def lift_ef(b=4, dtype="float64", nface=4, nvol=35, nfd=15):
    return f.batched_einsum(
        "ef,fij,fej->ei",
        [[f.array("J", ("E", nface), dtype), f.array("R", (nface, nvol, nfd), dtype),
          f.array(f"v{k}", (nface, "E", nfd), dtype)] for k in range(b)],
    )


def lift_fe(b=4, dtype="float64", nface=4, nvol=35, nfd=15):
    return f.batched_einsum(
        "ifj,fe,fej->ei",
        [[f.array("L", (nvol, nface, nfd), dtype), f.array("Jface", (nface, "E"), dtype),
          f.array(f"F_{k}", (nface, "E", nfd), dtype)] for k in range(b)],
    )


def tensor_product(mode=0, n=8, dtype="float64"):
    sub = ["eabc,ia->eibc", "eabc,ib->eaic", "eabc,ic->eabi"][mode]
    return f.einsum(sub, f.array("A", ("E", n, n, n), dtype), f.array("M", (n, n), dtype))


def div_components(dtype="float64"):
    return f.batched_einsum(
        "se, sij, ej -> ei",
        [[f.array(f"J{c}", (3, "E"), dtype), f.array("R", (3, 35, 35), dtype),
          f.array(f"u{c}", ("E", 35), dtype)] for c in "xyz"],
    )


def face_mass_se(dtype="float64"):
    return f.batched_einsum(
        "se, sij, ej -> ei",
        [[f.array("J", (4, "E"), dtype), f.array("R", (4, 15, 15), dtype),
          f.array(f"v{k}", ("E", 15), dtype)] for k in range(4)],
    )


def matvec_f32(long=False):
    A = f.array("A", ("I" if long else 10, 4), "float32")
    return f.batched_einsum(
        "ij, j -> i", [[A, f.array("x", 4, "float32")], [A, f.array("y", 4, "float32")]]
    )


def from_spec(spec):
    """Build from a tests/golden/frontend.json ``spec`` entry."""
    rows = [[f.array(a["name"], tuple(a["shape"]), a["dtype"]) for a in row] for row in spec["args"]]
    return f.batched_einsum(spec["subscripts"], rows)
