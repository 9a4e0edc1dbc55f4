"""CPU, this container only: the drop-in classes keep the reference's call surface.

The reference's sources are parsed (not imported: two of its modules need PyWavelets) and every public method of every
mirrored class is compared with the class of the same name under ``offmark_b200``: same method names, same positional
parameters in the same order with the same defaults; the B200 classes may only ADD trailing parameters that have
defaults (``device``, ``batch_frames``).  Skipped where the reference checkout is absent (the GPU box)."""
import ast
import importlib
import inspect
import os

import pytest

REFERENCE = os.path.join(os.sep, "root", "reference", "src", "offmark")
MIRRORED = {                                   # module path under offmark / offmark_b200 -> class
    "embed.dwt_dct_svd_encoder": "DwtDctSvdEncoder", "embed.dct_encoder": "DctEncoder",
    "extract.dwt_dct_svd_decoder": "DwtDctSvdDecoder", "extract.dct_decoder": "DctDecoder",
    "generator.shuffler": "Shuffler", "generator.grayscale": "GrayScale",
    "degenerator.de_shuffler": "DeShuffler", "degenerator.de_grayscale": "DeGrayScale",
    "video.embedder": "Embedder", "video.extractor": "Extractor",
}

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")


def _reference_methods(module, cls_name):
    """{method name: [(parameter, default or inspect.Parameter.empty), ...]} of the class as written in the reference."""
    tree = ast.parse(open(os.path.join(REFERENCE, *module.split(".")) + ".py").read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == cls_name)
    out = {}
    for fn in cls.body:
        if not isinstance(fn, ast.FunctionDef):
            continue
        name = fn.name
        if name.startswith("_" + cls_name + "__") or (name.startswith("__") and name != "__init__"):
            continue                           # name-mangled privates (self.__encode_frame ...)
        if name.startswith("_") and name != "__init__":
            continue
        params = [a.arg for a in fn.args.args][1:]                  # without self
        defaults = [inspect.Parameter.empty] * (len(params) - len(fn.args.defaults)) + \
                   [ast.literal_eval(d) for d in fn.args.defaults]
        out[name] = list(zip(params, defaults))
    return out


@pytest.mark.parametrize("module", sorted(MIRRORED))
def test_class_surface_equals_the_reference(module):
    cls_name = MIRRORED[module]
    want = _reference_methods(module, cls_name)
    assert "__init__" in want or cls_name in ("Embedder", "Extractor")
    ours = getattr(importlib.import_module("offmark_b200." + module), cls_name)
    for name, ref_params in want.items():
        fn = getattr(ours, name, None)
        assert callable(fn), f"{cls_name}.{name} is missing"
        got = [(p.name, p.default) for p in list(inspect.signature(fn).parameters.values())[1:]
               if p.kind in (p.POSITIONAL_OR_KEYWORD, p.POSITIONAL_ONLY)]
        assert got[:len(ref_params)] == ref_params, f"{cls_name}.{name}: {got} vs reference {ref_params}"
        for extra, default in got[len(ref_params):]:
            assert default is not inspect.Parameter.empty, f"{cls_name}.{name}: added parameter {extra} needs a default"


@pytest.mark.parametrize("module", sorted(MIRRORED))
def test_instances_carry_the_attributes_the_reference_sets_in_init(module):
    """``self.key``, ``self.scales``, ``self.frame_reader`` ...: scripts built on the reference read them."""
    cls_name = MIRRORED[module]
    tree = ast.parse(open(os.path.join(REFERENCE, *module.split(".")) + ".py").read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == cls_name)
    init = next((f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name == "__init__"), None)
    if init is None:
        pytest.skip("no __init__ in the reference class")
    attrs = {t.attr for n in ast.walk(init) if isinstance(n, ast.Assign) for t in n.targets
             if isinstance(t, ast.Attribute) and isinstance(t.value, ast.Name) and t.value.id == "self"
             and not t.attr.startswith("_")}
    ours = getattr(importlib.import_module("offmark_b200." + module), cls_name)
    required = [p for p in list(inspect.signature(ours.__init__).parameters.values())[1:] if p.default is p.empty]
    obj = ours(*[None] * len(required))
    missing = sorted(a for a in attrs if not hasattr(obj, a))
    assert not missing, f"{cls_name} lacks attributes the reference sets: {missing}"


def _reference_module(module):
    """One reference module loaded from its file (no package import: offmark/__init__ pulls in ffmpeg bindings)."""
    import importlib.util
    path = os.path.join(REFERENCE, *module.split(".")) + ".py"
    spec = importlib.util.spec_from_file_location("reference_" + module.replace(".", "_"), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_generators_equal_the_reference_on_random_payloads():
    """generator/shuffler.py:15-25, generator/grayscale.py:16-31 run live next to the host-side mirrors."""
    import warnings

    import numpy as np
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.generator.grayscale import GrayScale
    ref_s, ref_g = _reference_module("generator.shuffler").Shuffler, _reference_module("generator.grayscale").GrayScale
    rng = np.random.RandomState(5)
    for key in (None, 0, 7, 123456):
        for length, capacity in ((8, (1, 32400)), (8, (1, 5)), (13, (1, 1200)), (1, (1, 9)), (64, (3, 100)), (40, (1, 40))):
            payload = rng.randint(0, 2, length)
            keep = payload.copy()
            if key is None:                     # unseeded: only shapes, dtypes and the multiset of bits can agree
                a = Shuffler().generate_wm(payload, capacity)
                assert a.shape == tuple(capacity) and a.dtype == ref_s().generate_wm(payload, capacity).dtype
                continue
            a, b = Shuffler(key=key).generate_wm(payload, capacity), ref_s(key=key).generate_wm(payload, capacity)
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
            assert np.array_equal(payload, keep)                 # neither mutates its argument
        for shape, capacity in (((4, 6), (1, 100)), ((9, 9), (1, 50)), ((2, 2), (1, 4))):
            image = rng.randint(0, 256, shape).astype(np.uint8)
            with warnings.catch_warnings(record=True) as w_ours:
                warnings.simplefilter("always")
                a = GrayScale(key=key or 0).generate_wm(image, capacity)
            with warnings.catch_warnings(record=True) as w_ref:
                warnings.simplefilter("always")
                b = ref_g(key=key or 0).generate_wm(image, capacity)
            assert a.dtype == b.dtype and np.array_equal(a, b) and len(w_ours) == len(w_ref)
    assert Shuffler.wm_type() == ref_s.wm_type() and GrayScale.wm_type() == ref_g.wm_type()


def test_every_mirrored_module_is_listed_by_the_package():
    import offmark_b200
    listed = set(getattr(offmark_b200, "MIRRORED_MODULES", ()))
    if listed:
        assert {m for m in MIRRORED} <= {m.replace("offmark_b200.", "").replace("offmark.", "") for m in listed}
