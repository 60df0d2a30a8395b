"""CPU: the C-ABI library loads and exports every symbol include/ngrtd.h declares (no compute calls without a GPU);
host-side argument validation that does not need a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "libngrtd.so")


def _declared():
    src = open(os.path.join(ROOT, "include", "ngrtd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ngrtd_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(SO)


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libngrtd.so does not export %s" % n
    lib.ngrtd_version.restype = ctypes.c_int
    assert lib.ngrtd_version() == 100


def test_binding_covers_header(lib):
    from noblegas_rtd_mcmc_b200 import _lib
    assert sorted(_lib.EXPORTED) == _declared()


def test_argument_validation_without_device(lib):
    """Invalid arguments are rejected before any CUDA call, with a message."""
    from noblegas_rtd_mcmc_b200 import _lib
    h = ctypes.c_void_p()
    tr = (_lib.Tracer * 1)(_lib.Tracer(0, 0, 0.0, 0, 0))
    x = np.ones((8, 1))
    rc = _lib.lib.ngrtd_plan_create(ctypes.byref(h), 8, 1, _lib.hptr(x), None, 0.0, 1, tr, 99, 0, -1)
    assert rc == -1 and b"mod_type1" in _lib.lib.ngrtd_last_error()
    rc = _lib.lib.ngrtd_plan_create(ctypes.byref(h), 0, 1, _lib.hptr(x), None, 0.0, 1, tr, 2, 0, -1)
    assert rc == -1
    rc = _lib.lib.ngrtd_plan_create(ctypes.byref(h), 8, 1, _lib.hptr(x), None, 0.5, 1, tr, 2, 0, -1)
    assert rc == -1 and b"dtp" in _lib.lib.ngrtd_last_error()
    with pytest.raises(ValueError):
        _lib.Plan(x, [dict(series=0)], "gamma")
    with pytest.raises(ValueError):
        _lib.slot_array(["tau1", "nope"])
    rc = _lib.lib.ngrtd_ce_dev(0, 9, None, None, None, None, None, None, 0.0, 1, None, None)
    assert rc == -1


def test_dropin_signatures_match_reference():
    """Same public names / kwargs as the reference classes."""
    import inspect
    from noblegas_rtd_mcmc_b200 import convolution_integral_utils as conv, noble_gas_utils as ng, run_age_mcmc_utils as ram
    assert list(inspect.signature(conv.tracer_conv_integral.__init__).parameters) == ["self", "C_t", "t_samp"]
    for m in ("update_pars", "thalf_2_lambda", "gen_g_tp", "convolve"):
        assert hasattr(conv.tracer_conv_integral, m)
    assert list(inspect.signature(ng.noble_gas_fun.__init__).parameters) == ["self", "gases", "E", "T", "Ae", "F", "P", "S"]
    for m in ("parse_P", "lapse_rate", "solubility", "vapor_pressure", "equil_conc", "equil_conc_dry", "ce_exc", "update_pars"):
        assert hasattr(ng.noble_gas_fun, m)
    assert list(inspect.signature(ram.ForwardMod.__init__).parameters) == ["self", "conv_kwgs", "par_names", "tracer"]
    assert list(inspect.signature(ram.ForwardMod.perform).parameters) == ["self", "node", "inputs", "outputs"]
    assert ng.J_flux(1, 2700, 1000, 3.7, 10.2, 0.05) == 3.7657277999999995e-11


def test_header_is_plain_c_and_links(lib, tmp_path):
    """include/ngrtd.h compiles as strict C99 and examples/c_abi_demo.c links against libngrtd.so (it runs on a GPU box only)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L" + os.path.dirname(SO), "-lngrtd",
                           "-Wl,-rpath," + os.path.dirname(SO), "-lm"])
    assert os.path.exists(exe)


def test_theano_op_adapters_with_a_stand_in_theano(monkeypatch):
    """The reference's operator boundary is a Theano Op (run_age_mcmc_utils.py:47-52: itypes [dvector], otypes [dscalar];
    noble_gas_mcmc.py:205: @as_op dvector -> dvector).  theano / aesara are not in the image, so a stand-in module with the
    three names the adapters touch checks the wiring: class attributes, perform() delegation, as_op signature."""
    import sys
    import types
    tt = types.ModuleType("theano.tensor")
    tt.dvector, tt.dscalar = "dvector", "dscalar"

    class Op(object):
        def __call__(self, x):                       # what theano does when the node is evaluated
            out = [[None]]
            self.perform(None, [np.asarray(x, dtype=np.float64)], out)
            return out[0][0]
    tt.Op = Op
    th = types.ModuleType("theano")
    th.tensor = tt
    ops = types.ModuleType("theano.compile.ops")
    seen = {}

    def as_op(itypes, otypes):
        seen["sig"] = (itypes, otypes)
        return lambda fn: fn
    ops.as_op = as_op
    comp = types.ModuleType("theano.compile")
    comp.ops = ops
    for name, mod in (("theano", th), ("theano.tensor", tt), ("theano.compile", comp), ("theano.compile.ops", ops)):
        monkeypatch.setitem(sys.modules, name, mod)
    from noblegas_rtd_mcmc_b200 import run_age_mcmc_utils as R
    from noblegas_rtd_mcmc_b200 import noble_gas_mcmc as N

    class Fwd(object):                               # anything with the reference's perform(node, inputs, outputs)
        def perform(self, node, inputs, outputs):
            outputs[0][0] = np.array(float(np.sum(inputs[0])) * 2.0)
    op = R.as_theano_op(Fwd())
    assert isinstance(op, Op) and type(op).itypes == ["dvector"] and type(op).otypes == ["dscalar"]
    assert float(op([1.0, 2.5])) == 7.0
    wrapped = N.as_theano_op(["He", "Ne"])
    assert seen["sig"] == (["dvector"], ["dvector"]) and callable(wrapped)
