"""TEST INFRASTRUCTURE ONLY (oracle).

Loads the UNMODIFIED reference Python sources from ``/root/reference/feos_torch`` (never copied
into this repo) as the package ``feos_torch`` with

* ``si_units``            -> ``oracle/refshim/si_units`` (plain floats), and
* ``feos_torch.feos_torch`` (the Rust/PyO3 extension, ``/root/reference/src/lib.rs:10-16``) ->
  a stub whose ``PcSaft`` / ``GcPcSaft`` classes delegate to the oracle's CPU solver
  (``oracle/solver.py``), returning the same ``(rho[n_ok,...], status[N])`` tuples as
  ``/root/reference/src/pcsaft.rs:82-129,150-231`` and ``src/gc_pcsaft.rs:103-171``.

This only works where ``/root/reference`` exists (the build container). It is used by
``oracle/make_golden.py`` to produce ``tests/golden/*.json`` and by CPU tests that are skipped
when the reference is absent.  Nothing in the product (``feos_torch_b200``) imports it.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FEOS_TORCH_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "feos_torch", "pcsaft_pure.py"))


def load_reference():
    """Return the reference ``feos_torch`` package (imported from REFERENCE_ROOT)."""
    if "feos_torch" in sys.modules and getattr(sys.modules["feos_torch"], "_oracle_shim", False):
        return sys.modules["feos_torch"]
    if not reference_available():
        raise ImportError(f"reference sources not found under {REFERENCE_ROOT}")
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)  # makes `import si_units` resolve to the shim

    from oracle import native_stub

    pkg_dir = os.path.join(REFERENCE_ROOT, "feos_torch")
    pkg = types.ModuleType("feos_torch")
    pkg.__path__ = [pkg_dir]
    pkg.__file__ = os.path.join(pkg_dir, "__init__.py")
    pkg._oracle_shim = True
    sys.modules["feos_torch"] = pkg

    stub = types.ModuleType("feos_torch.feos_torch")
    stub.PcSaft = native_stub.PcSaft
    stub.GcPcSaft = native_stub.GcPcSaft
    stub.__version__ = "0.1.0-oracle-stub"
    sys.modules["feos_torch.feos_torch"] = stub
    pkg.feos_torch = stub
    pkg.PcSaft = stub.PcSaft
    pkg.GcPcSaft = stub.GcPcSaft

    for name in ("dual", "dual_torch", "pcsaft_pure", "pcsaft_mix", "gc_pcsaft"):
        spec = importlib.util.spec_from_file_location(
            f"feos_torch.{name}", os.path.join(pkg_dir, f"{name}.py")
        )
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"feos_torch.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    pkg.PcSaftPure = pkg.pcsaft_pure.PcSaftPure
    pkg.PcSaftMix = pkg.pcsaft_mix.PcSaftMix
    pkg.GcPcSaftMix = pkg.gc_pcsaft.GcPcSaftMix
    return pkg
