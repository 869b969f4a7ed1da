"""Import-time compatibility shim for running the UNMODIFIED reference in this container.

Used ONLY by tools/make_golden.py (golden-vector generation) — never by the product,
the tests, smoke() or bench.py (the reference tree does not exist on the GPU box).

The reference targets torch ~1.2-1.4. Nothing under /root/reference
is edited or copied; we only patch the interpreter around it:
  1. stub modules: soundfile.read, kenlm.LanguageModel, Levenshtein.{distance,editops}
     (data.py:12, model.py:13, util.py:9)
  2. legacy torch.stft real-valued output [.., 257, T, 2] (data.py:205-221)
  3. legacy integer floor division: torch.div(long, long, out=long) (model.py:866) and
     LongTensor / int (model.py:886)
  4. dict.pkl is opened by relative path (data.py:373) -> chdir while constructing AudioBase
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ASR_REFERENCE_DIR", "/root/reference")

_wav_store = {}  # path -> float32 numpy array, served by the soundfile stub


def register_wav(path, pcm):
    _wav_store[path] = np.asarray(pcm, dtype=np.float32)


def _install_stubs():
    sf = types.ModuleType("soundfile")

    def read(path, dtype="float32"):
        return _wav_store[path].copy(), 16000

    sf.read = read
    sys.modules["soundfile"] = sf

    km = types.ModuleType("kenlm")

    class LanguageModel:  # only constructed by main.ASR when lm_path is set
        def __init__(self, path):
            raise RuntimeError("kenlm is not available; inject an lm_model object instead")

    km.LanguageModel = LanguageModel
    km.State = object
    sys.modules["kenlm"] = km

    lv = types.ModuleType("Levenshtein")
    lv.distance = lambda a, b: 0
    lv.editops = lambda a, b: []
    sys.modules["Levenshtein"] = lv


_orig_stft = torch.stft
_orig_div = torch.div
_orig_truediv = torch.Tensor.__truediv__


def _stft(*args, **kw):
    kw["return_complex"] = True
    return torch.view_as_real(_orig_stft(*args, **kw))


def _div(a, b, *args, **kw):
    out = kw.get("out")
    if out is not None and not out.dtype.is_floating_point and "rounding_mode" not in kw:
        kw["rounding_mode"] = "floor"
    return _orig_div(a, b, *args, **kw)


def _truediv(self, other):
    if not self.dtype.is_floating_point and not self.dtype.is_complex and isinstance(other, int):
        return torch.floor_divide(self, other)
    return _orig_truediv(self, other)


@contextlib.contextmanager
def legacy_torch():
    torch.stft = _stft
    torch.div = _div
    torch.Tensor.__truediv__ = _truediv
    try:
        yield
    finally:
        torch.stft = _orig_stft
        torch.div = _orig_div
        torch.Tensor.__truediv__ = _orig_truediv


_mods = None


def load_reference():
    """Returns a namespace with the reference modules (gpd, data, model, util, ...)."""
    global _mods
    if _mods is not None:
        return _mods
    _install_stubs()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import gpd as r_gpd  # noqa
        r_gpd.gpd["use_cuda"] = False
        r_gpd.gpd["verbose"] = False
        r_gpd.gpd["eval_num_workers"] = 0
        r_gpd.gpd["temperature"] = 1
        import util as r_util  # noqa
        import data as r_data  # noqa
        import encoder as r_encoder  # noqa
        import attention as r_attention  # noqa
        import decoder as r_decoder  # noqa
        import model as r_model  # noqa
        audio_base = r_data.AudioBase()
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    _mods = types.SimpleNamespace(gpd=r_gpd.gpd, util=r_util, data=r_data, encoder=r_encoder,
                                  attention=r_attention, decoder=r_decoder, model=r_model,
                                  audio_base=audio_base)
    return _mods
