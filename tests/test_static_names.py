"""Static guard: every name a function of the product / bench / GPU workers reads must be bound somewhere (parameter, local,
module level or builtin).  The multi-GPU host code cannot run in the CPU container, so a typo there would only show up on the
GPU box - this test catches the NameError class of mistakes without executing anything."""
import builtins
import glob
import os
import symtable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = (glob.glob(os.path.join(ROOT, "pcg_mpi_solver_b200", "*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py")) +
         glob.glob(os.path.join(ROOT, "oracle", "*.py")) +
         [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py"), os.path.join(ROOT, "tests", "mgpu_worker.py")])


def _undefined(path):
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = {s.get_name() for s in top.get_symbols()}
    known = module_names | set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    bad = []

    def walk(tab):
        for child in tab.get_children():
            if child.get_type() in ("function", "class"):
                for s in child.get_symbols():
                    if s.is_referenced() and s.is_global() and not s.is_declared_global() and s.get_name() not in known:
                        bad.append((os.path.relpath(path, ROOT), child.get_name(), s.get_name()))
                walk(child)

    walk(top)
    return bad


def test_no_unbound_names_in_host_code():
    bad = [b for f in FILES for b in _undefined(f)]
    assert not bad, bad
