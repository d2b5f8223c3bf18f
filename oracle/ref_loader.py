"""Oracle: run the VERBATIM reference ``speechpipe.py`` with ``snac_ref`` injected as ``snac``.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Works only where
``/root/reference`` is mounted (the authoring container); the GPU box has no such
tree, so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.
Mirrors the injection trick of the reference's own
``tests/test_speechpipe_snac_path.py:7-33`` (a fake ``snac`` module in
``sys.modules`` + loading the file by path).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types
from typing import Dict, Optional

import torch

from . import snac_ref

REFERENCE_ROOT = os.environ.get("MORPHEUS_REFERENCE_ROOT", "/root/reference")
SPEECHPIPE_PATH = os.path.join(REFERENCE_ROOT, "Morpheus_Client", "tts_engine", "speechpipe.py")


def reference_available() -> bool:
    return os.path.isfile(SPEECHPIPE_PATH)


def load_reference_speechpipe(state_dict: Optional[Dict[str, torch.Tensor]] = None, quiet: bool = True):
    """Import the reference file under a private module name; returns the module.

    ``module.model`` is the ``snac_ref.SNAC`` instance the reference built through
    ``SNAC.from_pretrained("hubertsiuzdak/snac_24khz").eval().to(device)``.
    """
    if not reference_available():
        raise FileNotFoundError(SPEECHPIPE_PATH)
    shim = types.ModuleType("snac")
    shim.SNAC = snac_ref.SNAC
    saved = sys.modules.get("snac")
    saved_env = os.environ.pop("ORPHEUS_SNAC_PATH", None)
    if state_dict is not None:
        snac_ref.PRETRAINED["hubertsiuzdak/snac_24khz"] = state_dict
    sys.modules["snac"] = shim
    try:
        spec = importlib.util.spec_from_file_location("_morpheus_reference_speechpipe", SPEECHPIPE_PATH)
        module = importlib.util.module_from_spec(spec)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext():
            spec.loader.exec_module(module)
    finally:
        if saved is None:
            sys.modules.pop("snac", None)
        else:
            sys.modules["snac"] = saved
        if saved_env is not None:
            os.environ["ORPHEUS_SNAC_PATH"] = saved_env
        snac_ref.PRETRAINED.pop("hubertsiuzdak/snac_24khz", None)
    return module
