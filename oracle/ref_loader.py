"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

Imports `diff_model` (dm1), `ddim2/diff_model2` (dm2) and the IDDM sampler from
$ADVSHADOW_REF (default /root/reference) after installing import-time stubs for
packages the reference imports but never uses on the hot path (matplotlib,
fastai, coloredlogs).  Used only by `oracle/make_golden.py` (to mint the
fixtures under tests/golden/) and by tests that run in the dev container.
`/root/reference` does not exist on the GPU box: nothing under tests marked
`gpu`, `smoke()` or `bench.py` may call this module.
"""
import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("ADVSHADOW_REF", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "diff_model.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so sub-imports resolve
    sys.modules[name] = m
    return m


def install_stubs():
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("fastai")
    _stub("fastai.vision")
    _stub("fastai.vision.core", PILImage=object)
    _stub("coloredlogs", install=lambda *a, **k: None)


def _load(modname, relpath):
    install_stubs()
    key = "advshadow_ref_" + modname
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def dm1():
    """reference diff_model.py (UNetModel dm1:157-267, GaussianDiffusion dm1:286-484)."""
    return _load("dm1", "diff_model.py")


def dm2():
    """reference ddim2/diff_model2.py (bigger UNet defaults, apply_shadow dm2:615-654)."""
    return _load("dm2", "ddim2/diff_model2.py")


def iddm():
    """(UNet, DDIMDiffusion) of the vendored IDDM tree: model/networks/unet.py, model/samples/ddim.py."""
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from model.networks.unet import UNet          # noqa: E402
    from model.samples.ddim import DDIMDiffusion  # noqa: E402
    return UNet, DDIMDiffusion
