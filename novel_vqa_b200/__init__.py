"""Import shim: the package directory is ``novel-vqa_b200/`` (not a valid Python identifier);
``import novel_vqa_b200`` resolves its sub-modules from there."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "novel-vqa_b200"))

from .api import *  # noqa: F401,F403,E402
