"""Import alias for the package directory ``phoneme-vqa_b200/``.

The repository layout names the package ``phoneme-vqa_b200`` (with a hyphen, after the
reference repo's name), which is not a legal Python identifier.  This module makes
``import phoneme_vqa_b200`` (and ``phoneme_vqa_b200.<submodule>``) resolve to that
directory: it sets ``__path__`` so the import system treats it as a package and then
executes the directory's ``__init__.py`` in this module's namespace.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "phoneme-vqa_b200")
__path__ = [_PKG_DIR]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(_PKG_DIR, "__init__.py")
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"))
