"""Import shim: the package directory is named `opensearch-jvector_b200` (hyphen, per the repo layout),
which Python cannot import by name.  `import jvpkg; jv = jvpkg.load()` registers it as the module
`opensearch_jvector_b200`."""
import importlib.util
import sys
from pathlib import Path

_NAME = "opensearch_jvector_b200"
_DIR = Path(__file__).resolve().parent / "opensearch-jvector_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, _DIR / "__init__.py", submodule_search_locations=[str(_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
