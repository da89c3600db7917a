"""Import shim: the package directory is `chrono-mind_b200/` (hyphenated, as the repo layout
requires), which Python cannot name in an `import` statement. `import chronomind_b200` loads it;
`python -m chronomind_b200 query ...` runs its CLI (chrono-mind_b200/cli.py)."""
import importlib.util
import os
import sys

_NAME = "chronomind_b200"
_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chrono-mind_b200")
_spec = importlib.util.spec_from_file_location(
    _NAME, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod
_spec.loader.exec_module(_mod)

if __name__ == "__main__":
    import importlib
    sys.exit(importlib.import_module(_NAME + ".cli").main())
