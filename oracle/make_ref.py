"""ORACLE (test infrastructure, NOT product code) -- recipe for the reference arm's own sources.

The reference's path is five pure-Python files with no build step.  This recipe snapshots the three it needs
(`src/models.py`, `src/loss.py`, `src/utils.py`) from /root/reference into the reserved, git-ignored directory
`oracle/_ref/src/`, unmodified, so that `bench.py --impl reference` can time the UNMODIFIED reference classes on the GPU
box's host cores (`cpu_baseline.kind = "reference"`), where /root/reference does not exist.  `oracle/_ref/` never enters
the repository history (.gitignore) but travels with the snapshot to the GPU box like the built `.so`.  Run by
`__graft_entry__.build()` whenever /root/reference is present; nothing in the product package imports it.

    python oracle/make_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("models.py", "loss.py", "utils.py")


def make(ref_root="/root/reference"):
    src = os.path.join(ref_root, "src")
    if not all(os.path.exists(os.path.join(src, f)) for f in FILES):
        return None
    dst = os.path.join(HERE, "_ref", "src")
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    open(os.path.join(dst, "__init__.py"), "w").close()
    return dst


def load():
    """(models, loss) modules of the snapshot, or None when it has not been made."""
    dst = os.path.join(HERE, "_ref")
    if not os.path.exists(os.path.join(dst, "src", "models.py")):
        return None
    import importlib.util
    mods = {}
    for name in ("utils", "models", "loss"):
        spec = importlib.util.spec_from_file_location(f"_hidenn_ref_src.{name}", os.path.join(dst, "src", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        # loss.py does `from src.utils import ...` / `from .utils import ...`: provide both spellings
        sys.modules.setdefault("src", type(sys)("src"))
        sys.modules[f"src.{name}"] = mod
        sys.modules[f"_hidenn_ref_src.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["models"], mods["loss"]


if __name__ == "__main__":
    print(make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
