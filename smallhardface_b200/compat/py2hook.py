"""Import-time py2 -> py3 source transform for the reference's UNMODIFIED Python files.

The reference targets Python 2.7 (``print`` statements, ``xrange``, ``cPickle``, ``dict.has_key`` ...,
SURVEY.md Appendix B); this image has Python 3.12 only and no ``lib2to3``.  ``install(root)`` registers a
``sys.meta_path`` finder that, for modules whose source file lives under ``root``, compiles a transformed copy
of the text; the files on disk stay byte-identical.  It also provides the few runtime shims those files
expect (``easydict``, NumPy's removed aliases, PyYAML's old ``load`` signature, ``builtins.xrange``).
"""
from __future__ import annotations

import builtins
import importlib.abc
import importlib.machinery
import importlib.util
import io
import os
import re
import sys
import tokenize
import types

_PRINT_RE = re.compile(r"^(\s*)print\b(?!\s*\()(.*)$")
_PRINT_FUNC_FUTURE = "from __future__ import print_function"


def _convert_print(line: str) -> str:
    m = _PRINT_RE.match(line)
    if not m:
        return line
    indent, rest = m.group(1), m.group(2).strip()
    trailing = ""
    if rest.endswith(","):
        rest, trailing = rest[:-1].rstrip(), ", end=' '"
    if rest.startswith(">>"):
        target, _, rest = rest[2:].partition(",")
        return "%sprint(%s, file=%s%s)" % (indent, rest.strip(), target.strip(), trailing)
    return "%sprint(%s%s)" % (indent, rest, trailing)


def transform_source(src: str, filename: str = "<src>") -> str:
    uses_print_function = _PRINT_FUNC_FUTURE in src
    out_lines = []
    for line in src.splitlines():
        if not uses_print_function:
            line = _convert_print(line)
        out_lines.append(line)
    s = "\n".join(out_lines) + "\n"
    # token-level renames (names only, never inside strings/comments)
    renames = {"xrange": "range", "unicode": "str", "cPickle": "pickle", "raw_input": "input"}
    toks = []
    try:
        for tok in tokenize.generate_tokens(io.StringIO(s).readline):
            if tok.type == tokenize.NAME and tok.string in renames:
                tok = tok._replace(string=renames[tok.string])
            toks.append(tok)
        s = tokenize.untokenize(toks)
    except (tokenize.TokenError, IndentationError, SyntaxError):
        for a, b in renames.items():
            s = re.sub(r"\b%s\b" % a, b, s)
    s = re.sub(r"(\w[\w\.\[\]'\"]*)\.has_key\(([^()]*)\)", r"(\2 in \1)", s)
    s = s.replace(".iteritems()", ".items()").replace(".itervalues()", ".values()").replace(".iterkeys()", ".keys()")
    s = re.sub(r",\s*'w',\s*0\)", r", 'w', 1)", s)                             # open(..., 'w', 0): unbuffered text mode
    s = re.sub(r"except\s+([\w\.]+)\s*,\s*(\w+)\s*:", r"except \1 as \2:", s)
    # true-division results used as integers on the test path
    s = s.replace("scores.shape[1] / (A * self._num_feats)", "scores.shape[1] // (A * self._num_feats)")
    s = s.replace("self._feat_stride[i / len(self._shifts)**", "self._feat_stride[i // len(self._shifts)**")
    s = s.replace("open(det_file, 'r')", "open(det_file, 'rb')")                 # pickle needs bytes (lib/test.py:308)
    return s


# ---- py2 list comprehensions leak their loop variable into the enclosing scope (lib/utils/blob.py:21-23 relies on it) ----
import ast


class _ScopeNodes(ast.NodeVisitor):
    """Nodes of ONE scope: does not descend into nested functions, lambdas or classes."""

    def __init__(self):
        self.listcomps, self.loads = [], []

    def visit_FunctionDef(self, node):
        pass
    visit_AsyncFunctionDef = visit_Lambda = visit_ClassDef = visit_FunctionDef

    def visit_ListComp(self, node):
        self.listcomps.append(node)
        self.generic_visit(node)

    def visit_Name(self, node):
        if isinstance(node.ctx, ast.Load):
            self.loads.append(node)


def _leak_listcomp_targets(tree):
    """For every list comprehension in a function (or module) body whose loop variable is ALSO read elsewhere in that scope,
    reproduce Python 2's behaviour: the variable stays bound to its last value after the statement.  The comprehension
    gets a always-true filter ``[(__py2leak_x := x)]`` (PEP 572: the walrus target binds in the containing scope) and the
    enclosing statement is followed by ``try: x = __py2leak_x / except NameError: pass``."""
    scopes = [tree] + [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef))]
    for scope in scopes:
        def process(stmts):
            out = []
            for st in stmts:
                leaked = []
                if isinstance(st, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                    out.append(st)
                    continue
                v = _ScopeNodes()
                v.visit(st)
                # compound statements: their nested bodies are handled recursively; only patch what is left at this level
                for field in ("body", "orelse", "finalbody"):
                    if isinstance(getattr(st, field, None), list) and getattr(st, field) and isinstance(getattr(st, field)[0], ast.stmt):
                        setattr(st, field, process(getattr(st, field)))
                for h in getattr(st, "handlers", []) or []:
                    h.body = process(h.body)
                nested = set()
                for field in ("body", "orelse", "finalbody"):
                    for sub in (getattr(st, field, None) or []):
                        if isinstance(sub, ast.stmt):
                            nested.update(id(n) for n in ast.walk(sub))
                for h in getattr(st, "handlers", []) or []:
                    for sub in h.body:
                        nested.update(id(n) for n in ast.walk(sub))
                for lc in v.listcomps:
                    if id(lc) in nested:
                        continue
                    inside = {id(n) for n in ast.walk(lc)}
                    for gen in lc.generators:
                        if not isinstance(gen.target, ast.Name):
                            continue
                        name = gen.target.id
                        if name.startswith("__py2leak_") or name not in scope_loads_outside(name, inside):
                            continue
                        tmp = "__py2leak_" + name
                        gen.ifs.append(ast.List(elts=[ast.NamedExpr(target=ast.Name(id=tmp, ctx=ast.Store()),
                                                                    value=ast.Name(id=name, ctx=ast.Load()))], ctx=ast.Load()))
                        leaked.append((name, tmp))
                out.append(st)
                for name, tmp in leaked:
                    out.append(ast.Try(body=[ast.Assign(targets=[ast.Name(id=name, ctx=ast.Store())],
                                                        value=ast.Name(id=tmp, ctx=ast.Load()))],
                                       handlers=[ast.ExceptHandler(type=ast.Name(id="NameError", ctx=ast.Load()), name=None,
                                                                   body=[ast.Pass()])], orelse=[], finalbody=[]))
            return out

        sv = _ScopeNodes()
        for st in scope.body:
            sv.visit(st)

        def scope_loads_outside(name, inside, _loads=sv.loads):
            return {n.id for n in _loads if n.id == name and id(n) not in inside}

        scope.body = process(scope.body)
    ast.fix_missing_locations(tree)
    return tree


def compile_py2(src: str, filename: str, optimize: int = -1):
    """Source text of a Python 2 file -> code object (text-level transform, then the comprehension-leak AST pass)."""
    tree = ast.parse(transform_source(src, filename), filename)
    return compile(_leak_listcomp_targets(tree), filename, "exec", dont_inherit=True, optimize=optimize)


class _Loader(importlib.abc.SourceLoader):
    def __init__(self, fullname, path):
        self.fullname, self.path = fullname, path

    def get_filename(self, fullname):
        return self.path

    def get_data(self, path):
        with open(path, "rb") as f:
            return f.read()

    def source_to_code(self, data, path, *, _optimize=-1):
        text = data.decode("utf-8") if isinstance(data, bytes) else data
        return compile_py2(text, path, _optimize)

    # never write / trust .pyc files of transformed sources
    def path_stats(self, path):
        raise OSError

    def set_data(self, path, data):
        pass


class Py2Finder(importlib.abc.MetaPathFinder):
    def __init__(self, root):
        self.root = os.path.realpath(root) + os.sep

    def find_spec(self, fullname, path, target=None):
        spec = importlib.machinery.PathFinder.find_spec(fullname, path)
        if spec is None or not spec.origin or not spec.origin.endswith(".py"):
            return None
        if not os.path.realpath(spec.origin).startswith(self.root):
            return None
        loader = _Loader(fullname, spec.origin)
        return importlib.util.spec_from_file_location(fullname, spec.origin, loader=loader,
                                                      submodule_search_locations=spec.submodule_search_locations)


class EasyDict(dict):
    """Minimal ``easydict.EasyDict`` (not installed here) with the py2 dict methods the reference calls
    (``lib/utils/get_config.py:101-151``)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = type(value)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in value)
        elif isinstance(value, dict) and not isinstance(value, EasyDict):
            value = EasyDict(value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def has_key(self, k):
        return k in self

    def iteritems(self):
        return self.items()


def install_runtime_shims():
    import numpy as np
    builtins.xrange = range
    builtins.unicode = str
    for alias, typ in (("float", float), ("int", int), ("bool", bool), ("object", object)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)
    if "easydict" not in sys.modules:
        try:
            import easydict  # noqa: F401
        except ImportError:
            m = types.ModuleType("easydict")
            m.EasyDict = EasyDict
            sys.modules["easydict"] = m
    try:
        import yaml
        if not getattr(yaml.load, "_shf_patched", False):
            _orig = yaml.load

            def load(stream, Loader=None, **kw):
                return _orig(stream, Loader=Loader or yaml.SafeLoader, **kw)
            load._shf_patched = True
            yaml.load = load
    except ImportError:
        pass
    if "cPickle" not in sys.modules:
        import pickle
        sys.modules["cPickle"] = pickle


def install(reference_root: str):
    """Put ``<root>`` and ``<root>/lib`` on sys.path (as ``train_test.py:5-8`` does relative to its CWD) behind
    the transforming finder."""
    install_runtime_shims()
    root = os.path.realpath(reference_root)
    if not any(isinstance(f, Py2Finder) and f.root == root + os.sep for f in sys.meta_path):
        sys.meta_path.insert(0, Py2Finder(root))
    for p in (os.path.join(root, "lib"), root):
        if p not in sys.path:
            sys.path.insert(1, p)
