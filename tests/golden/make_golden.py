#!/usr/bin/env python
"""
Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU
box; nothing under ``tests/`` reads it at test time):

    python tests/golden/make_golden.py

The reference package cannot be imported as a whole here (loopy, pyopencl,
islpy, pymbolic, immutables, pytools, bidict, pybliss, opt_einsum are not
installed).  Its *front-end* (``einsum.py``, ``make_einsum.py``), its *input
generator* and its *acceptance expression* (``measure.py:63-108,145-159``) only
need numpy plus trivial helpers, so this script installs minimal stand-ins for
the missing modules (an insertion-ordered ``immutables.Map``, a caching
``pytools.memoize_method``, inert ``loopy``/``pyopencl``/``pymbolic``/``islpy``
shells, ``pyopencl.array.to_device`` = identity wrapper) and then imports the
reference's own source files by path.

Outputs
  frontend.json   -- what the reference front-end computes for a list of
                     constructions (subscripts string, shape, index lengths,
                     sum indices, output count) and the exception *type* it
                     raises for invalid ones.
  tccg.json       -- the 48 TCCG contractions as the reference's
                     ``utils.get_tccg_benchmark`` builds them
  numeric_*.npz   -- inputs drawn by the reference's ``generate_input_arrays``
                     (seed 0) and the outputs of the reference's acceptance
                     expression ``np.einsum(get_subscripts(), ..., optimize="optimal")``
                     for the BASELINE einsums at small E.
"""

from __future__ import annotations

import importlib
import json
import os
import sys
import types
from functools import wraps

import numpy as np

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- stubs -----
def _install_stubs() -> None:
    class Map(dict):  # insertion-ordered, hashable stand-in for immutables.Map
        def __hash__(self):  # type: ignore[override]
            return hash(frozenset(self.items()))

        def update(self, other=(), **kw):  # type: ignore[override]
            new = Map(self)
            dict.update(new, other, **kw)
            return new

    immutables = types.ModuleType("immutables")
    immutables.Map = Map
    sys.modules["immutables"] = immutables

    def memoize_method(fn):
        attr = f"_memo_{fn.__name__}"

        @wraps(fn)
        def wrapper(self, *args):
            cache = self.__dict__.setdefault(attr, {}) if hasattr(self, "__dict__") else {}
            if args not in cache:
                cache[args] = fn(self, *args)
            return cache[args]

        return wrapper

    def memoize_on_first_arg(fn):
        return fn

    class UniqueNameGenerator:
        def __init__(self):
            self.names = set()

        def add_names(self, names):
            self.names.update(names)

        def __call__(self, base):
            name, k = base, 0
            while name in self.names:
                name = f"{base}_{k}"
                k += 1
            self.names.add(name)
            return name

    pytools = types.ModuleType("pytools")
    pytools.memoize_method = memoize_method
    pytools.memoize_on_first_arg = memoize_on_first_arg
    pytools.UniqueNameGenerator = UniqueNameGenerator
    sys.modules["pytools"] = pytools

    class _Shell(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {})

    for modname in [
        "loopy", "loopy.symbolic", "loopy.match", "pymbolic", "pymbolic.primitives",
        "pymbolic.mapper", "pymbolic.mapper.evaluator", "islpy", "pyopencl",
        "pyopencl.tools", "opt_einsum", "bidict", "opentuner", "pybliss",
    ]:
        sys.modules[modname] = _Shell(modname)

    class HostArray:
        def __init__(self, a):
            self._a = a

        def get(self):
            return self._a

    cla = _Shell("pyopencl.array")
    cla.to_device = lambda queue, arr: HostArray(arr)
    sys.modules["pyopencl.array"] = cla
    sys.modules["pyopencl"].array = cla

    # bare package shell so that ``feinsum/__init__.py`` (which pulls in loopy
    # based modules) is NOT executed; submodules load from the reference tree
    pkg = types.ModuleType("feinsum")
    pkg.__path__ = [os.path.join(REF_SRC, "feinsum")]
    sys.modules["feinsum"] = pkg


_install_stubs()
ref_make = importlib.import_module("feinsum.make_einsum")
ref_measure = importlib.import_module("feinsum.measure")


# ------------------------------------------------------------- frontend -----
def _enc_dim(d):
    return int(d) if isinstance(d, (int, np.integer)) else {"param": d.name}


def _build(spec):
    rows = [
        [ref_make.array(a["name"], tuple(a["shape"]), a["dtype"]) for a in row]
        for row in spec["args"]
    ]
    return ref_make.batched_einsum(spec["subscripts"], rows)


def _arr(name, shape, dtype="float64"):
    return {"name": name, "shape": list(shape), "dtype": dtype}


VALID = {
    "grad_p4": {
        "subscripts": "xre,rij,ej->xei",
        "args": [[_arr("J", (3, 3, "E")), _arr("D", (3, 35, 35)), _arr("u", ("E", 35))]],
    },
    "div_p4": {
        "subscripts": "xre,rij,xej->ei",
        "args": [[_arr("J", (3, 3, "E")), _arr("D", (3, 35, 35)), _arr("u", (3, "E", 35))]],
    },
    "lift_p4_b4": {
        "subscripts": "ef,fij,fej->ei",
        "args": [
            [_arr("J", ("E", 4)), _arr("R", (4, 35, 15)), _arr(f"v{k}", (4, "E", 15))]
            for k in range(4)
        ],
    },
    "lift_fe_p4_b4": {
        "subscripts": "ifj,fe,fej->ei",
        "args": [
            [_arr("L", (35, 4, 15)), _arr("Jface", (4, "E")), _arr(f"F_{k}", (4, "E", 15))]
            for k in range(4)
        ],
    },
    "tensor_product_p7": {
        "subscripts": "eabc,ia->eibc",
        "args": [[_arr("A", ("E", 8, 8, 8)), _arr("M", (8, 8))]],
    },
    # reference test/test_codegen.py:34-60
    "div_components": {
        "subscripts": "se, sij, ej -> ei",
        "args": [
            [_arr(f"J{c}", (3, "E")), _arr("R", (3, 35, 35)), _arr(f"u{c}", ("E", 35))]
            for c in "xyz"
        ],
    },
    # reference test/test_codegen.py:69-88
    "face_mass_se": {
        "subscripts": "se, sij, ej -> ei",
        "args": [
            [_arr("J", (4, "E")), _arr("R", (4, 15, 15)), _arr(f"v{k}", ("E", 15))]
            for k in range(4)
        ],
    },
    # reference test/test_measure.py:33-52
    "matvec_f32": {
        "subscripts": "ij, j -> i",
        "args": [
            [_arr("A", (10, 4), "float32"), _arr("x", (4,), "float32")],
            [_arr("A", (10, 4), "float32"), _arr("y", (4,), "float32")],
        ],
    },
    "matvec_f32_long": {
        "subscripts": "ij, j -> i",
        "args": [
            [_arr("A", ("I", 4), "float32"), _arr("x", (4,), "float32")],
            [_arr("A", ("I", 4), "float32"), _arr("y", (4,), "float32")],
        ],
    },
    "diag_access": {
        "subscripts": "iij,j->i",
        "args": [[_arr("A", (5, 5, 7)), _arr("x", (7,))]],
    },
    "grad_p4_f32": {
        "subscripts": "xre,rij,ej->xei",
        "args": [[_arr("J", (3, 3, "E"), "float32"), _arr("D", (3, 35, 35), "float32"),
                  _arr("u", ("E", 35), "float32")]],
    },
    # lower-order tets (reference tuning/impls/ifj_fe_fej_to_ei_v3.py:602: ndof 4/10/20, nfacedof 3/6/10)
    "grad_p2": {
        "subscripts": "xre,rij,ej->xei",
        "args": [[_arr("J", (3, 3, "E")), _arr("D", (3, 10, 10)), _arr("u", ("E", 10))]],
    },
    "div_p3": {
        "subscripts": "xre,rij,xej->ei",
        "args": [[_arr("J", (3, 3, "E")), _arr("D", (3, 20, 20)), _arr("u", (3, "E", 20))]],
    },
    "lift_fe_p1_b4": {
        "subscripts": "ifj,fe,fej->ei",
        "args": [
            [_arr("L", (4, 4, 3)), _arr("Jface", (4, "E")), _arr(f"F_{k}", (4, "E", 3))]
            for k in range(4)
        ],
    },
    "lift_p3_b4": {
        "subscripts": "ef,fij,fej->ei",
        "args": [
            [_arr("J", ("E", 4)), _arr("R", (4, 20, 10)), _arr(f"v{k}", (4, "E", 10))]
            for k in range(4)
        ],
    },
    "div_p2_f32": {
        "subscripts": "xre,rij,xej->ei",
        "args": [[_arr("J", (3, 3, "E"), "float32"), _arr("D", (3, 10, 10), "float32"),
                  _arr("u", (3, "E", 10), "float32")]],
    },
}

INVALID = {
    "implicit_mode": {"subscripts": "ij,j", "args": [[_arr("A", (3, 4)), _arr("x", (4,))]]},
    "ellipsis": {"subscripts": "...j,j->...", "args": [[_arr("A", (3, 4)), _arr("x", (4,))]]},
    "repeated_out": {"subscripts": "ij,j->ii", "args": [[_arr("A", (3, 4)), _arr("x", (4,))]]},
    "bad_char": {"subscripts": "i1,j->i", "args": [[_arr("A", (3, 4)), _arr("x", (4,))]]},
    "upper_index": {"subscripts": "Ij,j->I", "args": [[_arr("A", (3, 4)), _arr("x", (4,))]]},
    "out_not_in_inputs": {"subscripts": "ij,j->k", "args": [[_arr("A", (3, 4)), _arr("x", (4,))]]},
    "operand_count": {"subscripts": "ij,j->i", "args": [[_arr("A", (3, 4))]]},
    "rank_mismatch": {"subscripts": "ij,j->i", "args": [[_arr("A", (3, 4, 5)), _arr("x", (4,))]]},
    "extent_mismatch": {"subscripts": "ij,j->i", "args": [[_arr("A", (3, 4)), _arr("x", (5,))]]},
    "dtype_conflict": {
        "subscripts": "ij,j->i",
        "args": [[_arr("A", (3, 4)), _arr("x", (4,))],
                 [_arr("A", (3, 4), "float32"), _arr("y", (4,))]],
    },
    "shape_conflict": {
        "subscripts": "ij,j->i",
        "args": [[_arr("A", (3, 4)), _arr("x", (4,))], [_arr("A", (3, 4)), _arr("x", (4, 1))]],
    },
    "name_clash_arg_index": {"subscripts": "ij,j->i", "args": [[_arr("i", (3, 4)), _arr("x", (4,))]]},
    "name_clash_param_index": {"subscripts": "ej,j->e", "args": [[_arr("A", ("e", 4)), _arr("x", (4,))]]},
}

BAD_SHAPES = {"negative": -1, "float": 2.5, "inf": "np.inf", "none": None}


def frontend_fixture():
    out = {"valid": {}, "invalid": {}, "bad_shape_component": {}}
    for name, spec in VALID.items():
        e = _build(spec)
        out["valid"][name] = {
            "spec": spec,
            "get_subscripts": e.get_subscripts(),
            "b": int(e.b),
            "n": int(e.n),
            "ndim": int(e.ndim),
            "shape": [_enc_dim(d) for d in e.shape],
            "index_to_dim_length": {k: _enc_dim(v) for k, v in e.index_to_dim_length.items()},
            "arg_to_shape": {k: [_enc_dim(d) for d in v] for k, v in e.arg_to_shape.items()},
            "arg_to_dtype": {k: np.dtype(v).name for k, v in e.arg_to_dtype.items()},
            "sum_indices": list(e.sum_indices),
            "all_args": sorted(e.all_args),
            "all_indices": sorted(e.all_indices),
            "all_size_params": sorted(p.name for p in e.all_size_params),
            "access": {
                k: [type(v).__name__, int(getattr(v, "output_index", getattr(v, "index", -1)))]
                for k, v in e.index_to_access_descr.items()
            },
        }
    for name, spec in INVALID.items():
        try:
            _build(spec)
        except Exception as exc:  # noqa: BLE001
            out["invalid"][name] = {"spec": spec, "raises": type(exc).__name__}
        else:
            out["invalid"][name] = {"spec": spec, "raises": None}
    for name, comp in BAD_SHAPES.items():
        val = np.inf if comp == "np.inf" else comp
        try:
            ref_make.array("A", (3, val))
        except Exception as exc:  # noqa: BLE001
            out["bad_shape_component"][name] = type(exc).__name__
        else:
            out["bad_shape_component"][name] = None
    return out


# -------------------------------------------------------------- numeric -----
NUMERIC = {
    "grad_p4": 6, "div_p4": 6, "lift_p4_b4": 6, "lift_fe_p4_b4": 6,
    "tensor_product_p7": 4, "div_components": 5, "face_mass_se": 5,
    "matvec_f32": 1, "matvec_f32_long": 9, "diag_access": 1, "grad_p4_f32": 6,
    "grad_p2": 20, "div_p3": 20, "lift_fe_p1_b4": 20, "lift_p3_b4": 20, "div_p2_f32": 20,
}


def numeric_fixture(name: str, long_dim_length: int):
    e = _build(VALID[name])
    # reference measure.py:80-108 (queue is unused by the stubbed to_device)
    arg_dict = dict(ref_measure.generate_input_arrays(None, e, long_dim_length))
    # reference measure.py:145-159, verbatim expression
    output_names = ["_fe_out", *[f"_fe_out_{i}" for i in range(e.b - 1)]]
    ref_outs = {
        output_name: np.einsum(
            e.get_subscripts(),
            *[arg_dict[arg.name].get() for arg in arg_row],
            optimize="optimal",
        )
        for output_name, arg_row in zip(output_names, e.args, strict=True)
    }
    payload = {f"in__{k}": v.get() for k, v in arg_dict.items()}
    payload.update({f"out__{k}": v for k, v in ref_outs.items()})
    payload["long_dim_length"] = np.int64(long_dim_length)
    return payload


def tccg_fixture():
    """What the reference's ``utils.get_tccg_benchmark(i)`` builds for i = 1..48 (reference
    ``src/feinsum/utils.py:103-233``): subscripts, operand shapes, output shape."""
    ref_utils = importlib.import_module("feinsum.utils")
    out = []
    for i in range(1, 49):
        e = ref_utils.get_tccg_benchmark(i)
        out.append({
            "i": i,
            "subscripts": e.get_subscripts(),
            "arg_shapes": [[int(d) for d in e.arg_to_shape[a.name]] for a in e.args[0]],
            "arg_names": [a.name for a in e.args[0]],
            "shape": [int(d) for d in e.shape],
        })
    return out


def main() -> None:
    with open(os.path.join(HERE, "frontend.json"), "w") as fh:
        json.dump(frontend_fixture(), fh, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "tccg.json"), "w") as fh:
        json.dump(tccg_fixture(), fh, indent=1, sort_keys=True)
    for name, n in NUMERIC.items():
        np.savez_compressed(os.path.join(HERE, f"numeric_{name}.npz"), **numeric_fixture(name, n))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
