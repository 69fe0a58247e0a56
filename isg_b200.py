"""Import alias: `import isg_b200` resolves to the package directory
`intrinsic-subgraph-generation-for-vqa_b200/` (a hyphenated name cannot be imported directly)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                          "intrinsic-subgraph-generation-for-vqa_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
