"""CPU tier: host-side logic -- the create_semantics dispatch, the formula compiler, the C ABI
surface of the CUDA library (load + symbols + planning, no compute calls), error behaviour."""

from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest

import pystl_b200
from pystl_b200 import capi, create_semantics
from pystl_b200.compiler import affine_of, compile_formula
from pystl_b200.formula import (
    BASELINE_SHAPES, BASELINE_SPECS, Always, And, Eventually, Interval, Not, Or, Predicate, Until, affine_predicate,
    from_spec, gt, lt,
)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------- dispatch
def test_create_semantics_mirrors_the_reference_dispatch():
    # same syntaxes / aliases / kwargs as pystl/semantics/__init__.py:76-125
    for syntax, cls in [
        ("classical", "B200ClassicRobustSemantics"),
        ("smooth", "B200SmoothRobustSemantics"),
        ("cumulative", "B200CumulativeSemantics"),
        ("dgmsr", "B200DgmsrSemantics"),
        ("agm", "B200AgmRobustSemantics"),
        ("arithmetic-geometric-mean", "B200AgmRobustSemantics"),
        (" Smooth ", "B200SmoothRobustSemantics"),
    ]:
        assert type(create_semantics(syntax, backend="b200")).__name__ == cls
    assert create_semantics("smooth", backend="cuda", temperature=0.25).temperature == 0.25
    sem = create_semantics("dgmsr", backend="B200", eps=1e-6, p=2)
    assert (sem.eps, sem.p) == (1e-6, 2)
    for bad in ("not_a_syntax", "classic", "traditional", "ctstl"):  # reference tests/test_api.py:197-205,255-257
        with pytest.raises(KeyError):
            create_semantics(bad, backend="b200")
    with pytest.raises(ValueError):  # torch/semantics.py:13-16
        create_semantics("smooth", backend="b200", temperature=0.0)
    with pytest.raises(TypeError):  # classes without kwargs reject them, like the reference's
        create_semantics("classical", backend="b200", temperature=1.0)


def test_registration_in_the_reference_registry_when_present():
    pystl = pytest.importorskip("pystl", reason="reference package not importable here")
    assert pystl_b200.install()
    import pystl.semantics as ref

    assert "b200" in ref.registry.backends()
    for syntax in ("classical", "smooth", "cumulative", "dgmsr", "agm"):
        assert f"{syntax}/b200" in ref.registry.names()
    # the alias gate of semantics/__init__.py:101-105 now lets the backend through
    sem = pystl.create_semantics("smooth", backend="b200", temperature=0.5)
    assert type(sem).__name__ == "B200SmoothRobustSemantics"
    # and the reference's own backends are untouched
    assert type(pystl.create_semantics("classical")).__name__ == "ClassicRobustSemantics"


# ---------------------------------------------------------------------------- formula / compiler
def test_formula_builders_and_interval_validation():
    with pytest.raises(ValueError):
        Interval(3, 2)
    with pytest.raises(ValueError):
        Interval(-1, 2)
    with pytest.raises(ValueError):
        And()
    p, q = gt(0, 0.0), lt(1, 1.0)
    phi = (p & ~q).always((0, 3)).until(q.eventually(Interval(1, None)), (0, 5), weights_pair=(1.0, 2.0))
    assert isinstance(phi, Until) and isinstance(phi.left, Always) and isinstance(phi.right, Eventually)
    assert phi.right.interval.end is None and phi.weights_pair == (1.0, 2.0)
    assert isinstance((p | q), Or) and isinstance(~p, Not)


def test_compile_is_post_order_and_dedups_predicates():
    p = gt(0, 0.5)
    phi = And(p.always((0, 4)), Not(p), Or(p, lt(1, 2.0)))
    cf = compile_formula(phi, "classical", 2)
    kinds = [cf.nodes[i].kind for i in range(cf.struct.n_nodes)]
    assert kinds[-1] == capi.NODE_AND
    for i in range(cf.struct.n_nodes):  # children precede parents
        n = cf.nodes[i]
        assert all(cf.children[n.child_off + k] < i for k in range(n.n_children))
    assert cf.struct.n_preds == 2  # x>0.5 used three times, y<2 once
    assert cf.channels_used == (0, 1)


def test_affine_probe_recognises_reference_style_lambdas():
    # the reference's own fixtures (tests/test_api.py:71-75)
    p1 = Predicate("speed_ok", fn=lambda s, t: 0.6 - s[t, 0])
    p2 = Predicate("alt_ok", fn=lambda s, t: s[t, 1] - 0.2)
    assert affine_of(p1, 2) == (((0, -1.0),), 0.6)
    terms, bias = affine_of(p2, 2)
    assert terms == ((1, 1.0),) and bias == pytest.approx(-0.2)
    assert affine_of(affine_predicate("p", {3: 2.0, 1: -1.0}, 0.5), 4) == (((1, -1.0), (3, 2.0)), 0.5)
    with pytest.raises(NotImplementedError):
        affine_of(Predicate("sq", fn=lambda s, t: s[t, 0] ** 2), 2)
    with pytest.raises(NotImplementedError):  # looks at another time step
        affine_of(Predicate("lag", fn=lambda s, t: s[t, 0] - s[max(t - 1, 0), 0]), 2)
    with pytest.raises(ValueError):  # classic.py:16-19
        affine_of(Predicate("nofn"), 2)


# ---------------------------------------------------------------------------- C ABI surface
def test_header_and_library_agree():
    """Every function declared in include/pystl_b200.h is exported by the built library."""
    header = open(os.path.join(ROOT, "include", "pystl_b200.h")).read()
    declared = set(re.findall(r"\b(pystl_b200_\w+)\s*\(", header))
    assert declared == set(capi.EXPORTS)
    lib = capi.lib()  # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pystl_b200_abi_version() == capi.ABI_VERSION
    assert C.sizeof(capi.Node) == 48 and C.sizeof(capi.Pred) == 80  # ABI 2: + slack, pair


def test_open_ended_windows_beyond_one_tile_become_leaf_channels():
    """An Always / Eventually without an end (api.py:49-54) keeps the whole trace in one tile; on longer traces the
    node is lowered to a leaf channel the suffix-scan kernels fill (compiler.OpenLeaf), and the predicate's slack
    carries the empty-window rule of api.py:218-221 into the plan."""
    from pystl_b200.compiler import OpenLeaf, start_slack, NEVER

    x = affine_predicate("x", {0: 1.0}, 0.0)
    y = affine_predicate("y", {1: -1.0}, 0.5)
    phi = And(Always(x, Interval(3, None)), Eventually(y, Interval(0, 10)))
    for sem_name in ("classical", "smooth", "agm", "dgmsr"):
        sem = create_semantics(sem_name, backend="b200")
        inside = sem.compile(phi, 2)
        assert inside.has_open and not inside.leaf_preds
        assert not sem._needs_open_leaf(phi, inside, 2048) and sem._needs_open_leaf(phi, inside, 20000)
        # (the gradient kernel keeps more planes: where IT stops fitting one tile, values go the scan path as well)
        cf = sem.compile(phi, 2, open_leaf=True)
        (leaf,) = cf.leaf_preds
        assert isinstance(leaf, OpenLeaf) and leaf.child is x and (leaf.mode, leaf.start) == (0, 3)
        assert cf.signal_channels == (1,) and cf.n_channels == 2 and not cf.has_open and cf.root_leaf is None
        slacks = sorted(cf.preds[i].slack for i in range(cf.struct.n_preds))
        assert slacks == [0, 3]
        short, long = sem.plan(phi, 2048, 2), sem.plan(phi, 20000, 2)
        assert short["n_tiles"] == 1 and long["n_tiles"] > 1
        assert short["start_slack"] == long["start_slack"] == 3 and long["lookahead"] == 10
        # a window over the open-ended one: its whole extent must end before the operand does
        alone = sem.compile(Always(x, Interval(3, None)), 2, open_leaf=True)  # the whole formula is the scanned window
        assert alone.root_leaf == 0 and alone.signal_channels == () and alone.n_channels == 1 and alone.preds[0].slack == 3
        over = Always(Always(x, Interval(3, None)), Interval(0, 5))
        assert sem.plan(over, 2048, 2)["start_slack"] == sem.plan(over, 20000, 2)["start_slack"] == 8
        # an operand that is undefined near the horizon: the open-ended window raises at every t
        never = Eventually(Always(x, Interval(1, 4)), Interval(2, None))
        assert start_slack(never) == NEVER
        assert sem.plan(never, 2048, 2)["raises_always"] == sem.plan(never, 20000, 2)["raises_always"] == 1
        assert sem.compile(never, 2, open_leaf=True).preds[0].slack == -1
    # cumulative values are (pos, neg) pairs: the window is two leaf columns, read back by a PAIR predicate
    cum = create_semantics("cumulative", backend="b200")
    assert cum.plan(phi, 20000, 2)["n_tiles"] > 1 and cum.plan(phi, 20000, 2)["start_slack"] == 3
    ccf = cum.compile(phi, 2, open_leaf=True)
    assert ccf.n_leaf_columns == 2 and ccf.signal_channels == (1,) and ccf.n_channels == 3
    pairs = [(ccf.preds[i].pair, ccf.preds[i].n_terms, ccf.preds[i].slack) for i in range(ccf.struct.n_preds)]
    assert sorted(pairs) == [(0, 1, 0), (1, 2, 3)]
    with pytest.raises(NotImplementedError):  # the cumulative Until without an end stays inside one tile
        cum.plan(Until(x, y, Interval(0, None)), 20000, 2)
    with pytest.raises(ValueError):  # a pair is a cumulative value
        create_semantics("smooth", backend="b200").compile(Predicate("p", metadata={"pair": (0, 1)}), 2)
    with pytest.raises(NotImplementedError):  # window weights are per position of the clipped window
        create_semantics("agm", backend="b200").plan(Always(x, Interval(0, None), weights=[1.0] * 20000), 20000, 2)
    # Until without an end: classical only (a clamp-composition scan); U[a, inf) = min(G[0, a] left, W(t + a))
    from pystl_b200.compiler import OpenUntilLeaf

    cl = create_semantics("classical", backend="b200")
    u = Until(x, y, Interval(4, None))
    cfu = cl.compile(u, 2, open_leaf=True)
    (leaf,) = cfu.leaf_preds
    assert isinstance(leaf, OpenUntilLeaf) and leaf.left is x and leaf.right is y and leaf.start == 4
    kinds = [cfu.nodes[i].kind for i in range(cfu.struct.n_nodes)]
    assert kinds == [capi.NODE_PRED, capi.NODE_PRED, capi.NODE_ALWAYS, capi.NODE_AND]
    assert cl.plan(u, 20000, 2)["start_slack"] == 4 and cl.plan(u, 2048, 2)["start_slack"] == 4
    with pytest.raises(NotImplementedError):
        create_semantics("agm", backend="b200").plan(u, 20000, 2)


@pytest.mark.parametrize("cfg", sorted(BASELINE_SPECS))
def test_plan_reports_tiling_and_validity(cfg):
    B, T, D = BASELINE_SHAPES[cfg]
    phi = from_spec(BASELINE_SPECS[cfg])
    for sem_name in ("classical", "smooth", "cumulative", "agm", "dgmsr"):
        plan = create_semantics(sem_name, backend="b200").plan(phi, T, D)
        assert plan["out_channels"] == (2 if sem_name == "cumulative" else 1)
        assert plan["lookahead"] == {"cfg1": 10, "cfg2": 50, "cfg3": 110, "cfg4": 264, "cfg5": 32}[cfg]
        assert plan["start_slack"] == 0 and plan["raises_always"] == 0
        assert plan["tile_ext"] == 8 * plan["threads"] and plan["threads"] % 32 == 0
        assert plan["n_tiles"] * plan["tile_out"] >= T
        assert plan["smem_bytes"] <= 227 * 1024


def test_plan_validity_rules_follow_the_reference():
    sem = create_semantics("classical", backend="b200")
    x = gt(0, 0.0)
    # api.py:218-221: t + a must not pass the horizon
    assert sem.plan(x.eventually((2, 7)), 30, 1)["start_slack"] == 2
    # a child that raises near the horizon makes every clipped parent window raise: slack = a_child + b
    assert sem.plan(x.eventually((3, 11)).always((1, 33)), 200, 1)["start_slack"] == 36
    # api.py:288-293: Until evaluates its children up to the horizon
    assert sem.plan(x.eventually((1, 2)).until(x, (0, 3)), 80, 1)["raises_always"] == 1
    assert sem.plan(x.eventually((0, 2)).until(x, (0, 3)), 80, 1)["raises_always"] == 0
    # an open interval keeps the trace in one tile; beyond that the window is a scanned leaf channel (next test)
    assert sem.plan(x.always((0, None)), 100000, 1)["n_tiles"] > 1
    assert sem.plan(x.always((0, None)), 4000, 1)["n_tiles"] == 1
    assert sem.plan(x.until(x, (0, None)), 100000, 1)["n_tiles"] > 1  # classical: a scanned leaf channel as well
    with pytest.raises(NotImplementedError):  # a non-classical unbounded Until on a trace longer than any tile
        create_semantics("smooth", backend="b200").plan(x.until(x, (0, None)), 100000, 1)


def test_loud_failure_without_the_library(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.lib()


def test_no_cuda_device_is_an_error_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sem = create_semantics("classical", backend="b200")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sem.robustness_trace(gt(0, 0.0).always((0, 3)), np.zeros((2, 10, 1), np.float32))


def test_non_affine_predicates_become_leaf_channels():
    """Opaque callables that are not affine in signal[t] are compiled to leaf channels behind the
    signal's own (compiler.leaf_values evaluates them with torch on the device)."""
    from pystl_b200.compiler import compile_formula
    from pystl_b200.formula import Always, And, Interval, Predicate

    sq = Predicate("sq", fn=lambda s, t: s[t, 0] ** 2 - 1.0)
    lin = Predicate("lin", fn=lambda s, t: 0.5 - s[t, 1])
    phi = And(Always(sq, Interval(0, 3)), lin, Always(sq, Interval(1, 2)))
    cf = compile_formula(phi, "smooth", 2, temperature=1.0)
    # the kernels read [the signal channels an affine predicate uses] + [one channel per leaf]: channel 0 is only read by fn
    assert len(cf.leaf_preds) == 1 and cf.n_signal_channels == 2 and cf.signal_channels == (1,) and cf.n_channels == 2
    terms = sorted(cf.pred_terms)
    assert (((0, -1.0),), 0.5) in terms and (((1, 1.0),), 0.0) in terms
    both = compile_formula(And(phi, Predicate("lin0", fn=lambda s, t: s[t, 0] + 1.0)), "smooth", 2, temperature=1.0)
    assert both.signal_channels == (0, 1) and both.n_channels == 3


# ---------------------------------------------------------------------------- probing is only a fast reject
def test_probe_rejects_what_it_can_and_real_signals_catch_the_rest():
    """a predicate that depends on t beyond the probe's rows, or is affine only near the
    origin, must not be compiled as an affine predicate."""
    from pystl_b200.compiler import verify_probed

    ref = np.array([0.0] * 4 + [1.0] * 200)
    with pytest.raises(NotImplementedError):  # t-dependence inside the probe's rows
        affine_of(Predicate("a", fn=lambda s, t: 0.5 - abs(ref[t]) + s[t, 0]), 2)
    with pytest.raises(NotImplementedError):  # saturates far from the origin
        affine_of(Predicate("b", fn=lambda s, t: min(s[t, 0], 50.0)), 2)
    # t-dependence beyond the probe's rows: passes the probe, is caught on the real signal
    late = np.array([0.0] * 100 + [1.0] * 100)
    p = Predicate("late", fn=lambda s, t: 0.5 - late[t] + s[t, 0])
    assert affine_of(p, 2) == (((0, 1.0),), 0.5)
    phi = Always(p, Interval(0, 3)) & Predicate("ok", fn=lambda s, t: s[t, 1] - 0.25)
    cf = compile_formula(phi, "classical", 2)
    assert len(cf.probed_preds) == 2 and not cf.leaf_preds
    traces = np.random.default_rng(0).normal(size=(2, 200, 2))
    bad = verify_probed(cf, traces)
    assert bad == (id(p.fn),)
    cf2 = compile_formula(phi, "classical", 2, force_leaf=bad)
    assert len(cf2.leaf_preds) == 1 and cf2.leaf_preds[0] is p and cf2.signal_channels == (1,) and cf2.n_channels == 2
    assert verify_probed(cf2, traces) == ()
    # metadata gives the structured form: never probed, never second-guessed
    assert compile_formula(gt(0, 0.5).always((0, 3)), "classical", 2).probed_preds == ()
    # opting out of the affine path
    forced = Predicate("leaf", fn=lambda s, t: s[t, 0] - 1.0, metadata={"affine": None})
    assert len(compile_formula(forced.always((0, 2)), "smooth", 2).leaf_preds) == 1
    forced2 = Predicate("leaf", fn=lambda s, t: s[t, 0] - 1.0, metadata={"leaf": True})
    assert len(compile_formula(forced2.always((0, 2)), "smooth", 2).leaf_preds) == 1


def test_semantics_parameters_are_part_of_the_compile_cache_key():
    sem = create_semantics("smooth", backend="b200", temperature=1.0)
    phi = gt(0, 0.0).always((0, 5))
    a = sem.compile(phi, 1)
    assert sem.compile(phi, 1) is a
    sem.temperature = 0.25
    b = sem.compile(phi, 1)
    assert b is not a and b.struct.temperature == pytest.approx(0.25)
    sem.invalidate(phi)
    assert sem.compile(phi, 1) is not b


def test_backend_classes_subclass_the_reference_abc_when_it_is_importable():
    try:
        from pystl.semantics.base import Semantics
    except ImportError:
        pytest.skip("the reference package is not importable here")
    import importlib

    import pystl_b200.semantics as mod

    if mod._ReferenceSemantics is object:  # pystl_b200 was imported before the reference became importable
        mod = importlib.reload(mod)
    for cls in (mod.B200ClassicRobustSemantics, mod.B200SmoothRobustSemantics, mod.B200CumulativeSemantics,
                mod.B200AgmRobustSemantics, mod.B200DgmsrSemantics):
        assert issubclass(cls, Semantics)
        cls()  # every abstract method is implemented


def test_custom_ops_are_registered_with_fake_kernels():
    """pystl_b200::robustness / robustness_grad / value_and_grad: shapes propagate through fake
    tensors without a GPU (what torch.compile traces with)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    from pystl_b200 import ops

    sem = create_semantics("cumulative", backend="b200")
    cf = sem.compile(from_spec(BASELINE_SPECS["cfg4"]), 4)
    h = ops.handle_of(cf, sem.out_channels)
    assert ops.handle_of(cf, sem.out_channels) == h
    with FakeTensorMode():
        x = torch.empty(5, 100, 4, device="cuda")
        out = torch.ops.pystl_b200.robustness(x, h)
        assert tuple(out.shape) == (5, 100, 2) and out.device.type == "cuda"
        g = torch.ops.pystl_b200.robustness_grad(x, out, h)
        assert tuple(g.shape) == (5, 100, 4)
        o, acc, gs = torch.ops.pystl_b200.value_and_grad(x, 0.5, h)
        assert tuple(o.shape) == (5, 100, 2) and acc.dtype == torch.float64 and tuple(gs.shape) == (5, 100, 4)
    with pytest.raises(RuntimeError, match="unknown formula handle"):
        ops._lookup(10 ** 9)


def test_operator_interface_validates_like_the_reference():
    """Argument checks of the list-based interface happen before anything touches the device
    (classic.py:28-29,50-61; cumulative.py:85-94)."""
    sem = create_semantics("classical", backend="b200")
    with pytest.raises(ValueError, match="at least one value"):
        sem.boolean_and([])
    with pytest.raises(ValueError, match="non-empty"):
        sem.temporal_until(left_trace=[], right_trace=[], start=0, end=None)
    with pytest.raises(ValueError, match="same length"):
        sem.temporal_until(left_trace=[1.0, 2.0], right_trace=[1.0], start=0, end=None)
    with pytest.raises(ValueError, match="start must be >= 0"):
        sem.temporal_until(left_trace=[1.0], right_trace=[1.0], start=-1, end=None)
    with pytest.raises(ValueError, match="window is empty"):
        sem.temporal_until(left_trace=[1.0, 2.0], right_trace=[1.0, 2.0], start=5, end=7)
    cum = create_semantics("cumulative", backend="b200")
    with pytest.raises(ValueError, match="window is empty"):
        cum.temporal_until(left_trace=[cum._leaf(__import__("torch").zeros(()))], right_trace=[cum._leaf(__import__("torch").zeros(()))],
                           start=3, end=4)


def test_cubin_cache_rejects_damaged_and_foreign_entries(tmp_path, monkeypatch):
    """The on-disk cubin cache (jit.cpp): an entry is trusted only with its descriptor (source length, second hash, cubin
    length); a truncated cubin, a missing descriptor or a descriptor of another source is recompiled, not loaded."""
    import subprocess
    import sys
    import textwrap

    code = textwrap.dedent("""
        import os, sys, glob
        from pystl_b200 import capi
        from pystl_b200.compiler import compile_formula
        from pystl_b200.formula import from_spec, BASELINE_SPECS
        lib = capi.lib()
        cf = compile_formula(from_spec(BASELINE_SPECS["cfg1"]), "classical", 2)
        def build():
            rc = lib.pystl_b200_fused_precompile(cf.ptr(), 1000, 2, 0)
            assert rc == 0, lib.pystl_b200_last_error()
            (cub,) = glob.glob(os.path.join(os.environ["PYSTL_B200_CACHE"], "*.cubin"))
            return cub, cub[:-6] + ".meta"
        cub, meta = build()
        full = os.path.getsize(cub)
        assert full > 1000 and os.path.getsize(meta) == 32
        with open(cub, "r+b") as fh:            # truncated image
            fh.truncate(full // 2)
        cub, meta = build()
        assert os.path.getsize(cub) == full
        os.unlink(meta)                         # image without its descriptor
        cub, meta = build()
        assert os.path.getsize(meta) == 32
        blob = bytearray(open(meta, "rb").read())  # descriptor of another source (a colliding key)
        blob[16] ^= 0xFF
        open(meta, "wb").write(bytes(blob))
        cub, meta = build()
        assert bytes(open(meta, "rb").read()) != bytes(blob) and os.path.getsize(cub) == full
        print("ok")
    """)
    env = dict(os.environ, PYSTL_B200_CACHE=str(tmp_path), PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]
