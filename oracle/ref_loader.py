"""Executes the reference's own source over the TF1 shim -- TEST INFRASTRUCTURE ONLY.

Only usable where /root/reference exists (the build container, never the GPU box).
Nothing is copied: whole modules are imported from where they lie; for files whose
module-level imports cannot be satisfied (my_losses.py, Demon_Data_loader.py pull in
DeMoN / slim) only the named function definitions are compiled out of the parsed file.
"""
import ast
import importlib.util
import os
import sys

REF_ROOT = os.environ.get('VSL_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tf1_shim')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'utils_lr.py'))


def _with_shim():
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import tensorflow as tf  # the shim
    assert tf.__file__.startswith(_SHIM), 'a real tensorflow shadows the shim: %s' % tf.__file__
    return tf


def load_module(filename, alias):
    """Import /root/reference/<filename> unmodified under the module name `alias`."""
    _with_shim()
    import pdb
    pdb.set_trace = lambda *a, **k: None  # utils.py:73 / :314 hold live breakpoints
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF_ROOT, filename))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_functions(filename, names, extra_globals=None):
    """Compile only the top-level functions `names` of a reference file; return them in a dict."""
    tf = _with_shim()
    import numpy as np
    path = os.path.join(REF_ROOT, filename)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    assert not missing, 'not found in %s: %s' % (filename, sorted(missing))
    code = compile(ast.Module(body=keep, type_ignores=[]), path, 'exec',
                   flags=__import__('__future__').division.compiler_flag)
    ns = {'tf': tf, 'np': np}
    ns.update(extra_globals or {})
    exec(code, ns)
    return {n: ns[n] for n in names}
