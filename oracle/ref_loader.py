"""TEST / BENCH INFRASTRUCTURE -- loads the UNMODIFIED reference modules of the hot path by file path.

Where the files come from (first hit wins):
  1. /root/reference/fish_diffusion/...        (build container only)
  2. oracle/_ref/fish_diffusion/...            (verbatim copies made by oracle/build_ref.py; git-ignored, travels to
                                                the GPU box with the gpurun snapshot like a built .so)
Nothing under fish_diffusion_b200/ imports this module; only tests/, bench.py's reference / cpu_baseline legs and
tests/golden/make_golden.py do.

The reference package cannot be imported as a package here (its __init__ chain needs mmengine, pytorch_lightning,
librosa, ... -- SURVEY.md section 8c), so each file is executed on its own:
  modules/wavenet.py, modules/vocoders/nsf_hifigan/models.py                       -> torch / numpy only
  archs/diffsinger/diffusions/{uni_pc,noise_predictor,diffusion}.py                -> `.builder` (mmengine + unrelated
        denoisers) is replaced by a stub holding two minimal registries with the reference WaveNet registered
  utils/pitch_adjustable_mel.py                                                    -> `librosa.filters.mel` is served
        by the oracle's Slaney filterbank when librosa is absent (pins everything but the filterbank constants)
"""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ("/root/reference", os.path.join(HERE, "_ref"))

# files of the reference that the hot path consists of (relative to the reference root); build_ref.py copies these
REF_FILES = (
    "fish_diffusion/modules/wavenet.py",
    "fish_diffusion/archs/diffsinger/diffusions/uni_pc.py",
    "fish_diffusion/archs/diffsinger/diffusions/noise_predictor.py",
    "fish_diffusion/archs/diffsinger/diffusions/diffusion.py",
    "fish_diffusion/modules/vocoders/nsf_hifigan/models.py",
    "fish_diffusion/utils/pitch_adjustable_mel.py",
)


def reference_root():
    """Directory that holds fish_diffusion/ (None if neither the reference nor its oracle/_ref copy is present)."""
    for root in CANDIDATES:
        if all(os.path.exists(os.path.join(root, f)) for f in REF_FILES):
            return root
    return None


class _Registry:
    """The subset of mmengine.Registry the reference diffusion.py uses: register_module(name=, module=) and
    build(cfg) = pop `type`, instantiate with the remaining keys."""

    def __init__(self, name):
        self.name, self._m = name, {}

    def register_module(self, name=None, module=None, force=False):
        if module is not None:
            self._m[name or module.__name__] = module
            return module

        def deco(cls):
            self._m[name or cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self._m[cfg.pop("type")](**cfg)


def _load(name, path, package=None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_reference(root=None, with_mel=True):
    """-> namespace(wavenet, uni_pc, noise_predictor, diffusion, nsf, mel, root).  Raises if no reference is present."""
    root = root or reference_root()
    if root is None:
        raise FileNotFoundError("no reference found: neither /root/reference nor oracle/_ref (run oracle/build_ref.py "
                                "in the build container)")
    key = (root, with_mel)
    if key in _cache:
        return _cache[key]
    fd = os.path.join(root, "fish_diffusion")
    ref = types.SimpleNamespace(root=root)
    ref.wavenet = _load("ref_wavenet", f"{fd}/modules/wavenet.py")
    pkg = types.ModuleType("refdiff")
    pkg.__path__ = [f"{fd}/archs/diffsinger/diffusions"]
    sys.modules["refdiff"] = pkg
    builder = types.ModuleType("refdiff.builder")
    builder.DIFFUSIONS = _Registry("diffusions")
    builder.DENOISERS = _Registry("denoisers")
    builder.DENOISERS.register_module(name="WaveNetDenoiser", module=ref.wavenet.WaveNet)
    sys.modules["refdiff.builder"] = builder
    ref.uni_pc = _load("refdiff.uni_pc", f"{fd}/archs/diffsinger/diffusions/uni_pc.py", "refdiff")
    ref.noise_predictor = _load("refdiff.noise_predictor", f"{fd}/archs/diffsinger/diffusions/noise_predictor.py",
                                "refdiff")
    ref.diffusion = _load("refdiff.diffusion", f"{fd}/archs/diffsinger/diffusions/diffusion.py", "refdiff")
    ref.nsf = _load("ref_nsf_models", f"{fd}/modules/vocoders/nsf_hifigan/models.py")
    ref.mel = None
    if with_mel:
        try:
            import librosa  # noqa: F401
            import librosa.filters  # noqa: F401
        except Exception:  # noqa: BLE001 -- librosa absent: serve the filterbank from the oracle restatement
            from . import mel as omel
            lib = types.ModuleType("librosa")
            filt = types.ModuleType("librosa.filters")
            filt.mel = lambda sr, n_fft, n_mels, fmin, fmax: omel.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
            lib.filters = filt
            sys.modules.setdefault("librosa", lib)
            sys.modules.setdefault("librosa.filters", filt)
        try:
            ref.mel = _load("ref_pam", f"{fd}/utils/pitch_adjustable_mel.py")
        except Exception:  # noqa: BLE001 -- loguru missing etc.: the mel front end is then not available
            ref.mel = None
    _cache[key] = ref
    return ref
