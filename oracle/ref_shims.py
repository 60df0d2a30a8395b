"""TEST INFRASTRUCTURE ONLY -- import shims that let the *untouched* reference
(/root/reference, read-only, Python) import in the build container.

Used solely by oracle/gen_golden.py (golden-vector generation) and by the
`needs_reference` tests that validate the oracle restatement against the real
reference when /root/reference is mounted.  Nothing under noblegas_rtd_mcmc_b200/
may import this file.  /root/reference does not exist on the GPU box.

What is shimmed:
  * matplotlib{,.pyplot,.patches,.ticker,.style} are absent -> empty stub modules
    (imported at utils/noble_gas_utils.py:19-21, run_age_mcmc_utils.py:20-23)
  * scipy.integrate.trapz / cumtrapz were removed upstream -> aliases
    (imported at utils/convolution_integral_utils.py:11-12)
  * theano / theano.tensor / pymc3 / arviz are absent -> stubs with an `Op`
    base class (run_age_mcmc_utils.py:31-35,47)
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("NGRTD_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils"))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install():
    """Install the stubs and sys.path entries. Idempotent."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot", rcParams={})
        pat = _stub("matplotlib.patches")
        tick = _stub("matplotlib.ticker", MaxNLocator=object, MultipleLocator=object,
                     AutoMinorLocator=object, LogLocator=object)
        sty = _stub("matplotlib.style")
        mpl.pyplot, mpl.patches, mpl.ticker, mpl.style = plt, pat, tick, sty
    import scipy.integrate as si
    if not hasattr(si, "trapz"):
        si.trapz = si.trapezoid
    if not hasattr(si, "cumtrapz"):
        si.cumtrapz = si.cumulative_trapezoid

    class Op(object):
        pass

    tt = _stub("theano.tensor", Op=Op, dvector=object(), dscalar=object())
    th = _stub("theano", tensor=tt)
    th.tensor = tt
    _stub("pymc3")
    _stub("arviz")
    for sub in ("utils", "age_ens_runs_mcmc"):
        p = os.path.join(REFERENCE_ROOT, sub)
        if p not in sys.path:
            sys.path.insert(0, p)


def load():
    """Return (convolution_integral_utils, noble_gas_utils, run_age_mcmc_utils) of the reference."""
    install()
    import convolution_integral_utils as conv
    import noble_gas_utils as ngu
    cwd = os.getcwd()
    try:
        import run_age_mcmc_utils as ramu
    finally:
        os.chdir(cwd)
    return conv, ngu, ramu
