"""TEST INFRASTRUCTURE ONLY -- loads the real PFLeget/cosmogp from /root/reference.

The reference is Python 2 (three `print` statements: cosmogp/inv_matrix.py:9,
cosmogp/kernel.py:138,153).  This loader reads the sources where they lie under
/root/reference, rewrites those statements IN MEMORY and executes the modules in
the order cosmogp/__init__.py:12-25 imports them.  Nothing is copied into the repo.

/root/reference exists only in the build container, never on the GPU box: the
loader is used by tests/golden/make_golden.py (fixture generation) and by the
`-m "not gpu"` tests that pin oracle/gp_oracle.py against the real reference;
those tests skip when the path is absent.  The product never imports this file.

On the GPU box the same unmodified sources are found under baseline/_ref (the offline
pip install of the reference done by baseline/fetch_ref.py: git-ignored, shipped by
gpurun); bench.py's CPU arm (--impl reference, cpu_baseline) times them from there.
Search order: $COSMOGP_REFERENCE_ROOT, /root/reference, <repo>/baseline/_ref.
"""
import contextlib
import importlib.machinery
import importlib.util
import io
import os
import re
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    cands = [os.environ.get("COSMOGP_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "cosmogp", "__init__.py")):
            return c
    return cands[0] or cands[1]


REFERENCE_ROOT = _find_root()
_PRINT_STMT = re.compile(r"^(\s*)print (.*)$", re.M)
_ORDER = ["inv_matrix", "mean", "Gaussian_process", "kernel", "pull"]


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cosmogp", "__init__.py"))


def load(name="cosmogp"):
    """Return the reference package as a module object (cached in sys.modules)."""
    if name in sys.modules and getattr(sys.modules[name], "__cosmogp_reference__", False):
        return sys.modules[name]
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    pkg_dir = os.path.join(REFERENCE_ROOT, "cosmogp")

    def source(mod):
        fn = os.path.join(pkg_dir, mod + ".py")
        with open(fn) as f:
            return fn, _PRINT_STMT.sub(r"\1print(\2)", f.read())

    pkg = types.ModuleType(name)
    pkg.__path__ = []           # no on-disk submodule lookup: everything is exec'd here
    pkg.__package__ = name
    pkg.__cosmogp_reference__ = True
    saved = sys.modules.get(name)
    sys.modules[name] = pkg
    try:
        fn, src = source("__init__")
        for mod in _ORDER:
            mfn, msrc = source(mod)
            m = types.ModuleType(name + "." + mod)
            m.__file__ = mfn
            m.__package__ = name
            sys.modules[name + "." + mod] = m
            # `from cosmogp import X` inside the modules must hit this package
            code = compile(msrc.replace("from cosmogp import", "from %s import" % name)
                               .replace("import cosmogp\n", "import %s as cosmogp\n" % name),
                           mfn, "exec")
            exec(code, m.__dict__)
            setattr(pkg, mod, m)
            # replay the re-exports __init__ does after importing this module
            for line in src.splitlines():
                mm = re.match(r"from \.%s import (\w+)" % mod, line)
                if mm:
                    setattr(pkg, mm.group(1), getattr(m, mm.group(1)))
    except Exception:
        if saved is not None:
            sys.modules[name] = saved
        else:
            sys.modules.pop(name, None)
        raise
    return pkg


@contextlib.contextmanager
def quiet():
    """HEAD's rbf_kernel_2d prints on every call (kernel.py:138,153): swallow it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
