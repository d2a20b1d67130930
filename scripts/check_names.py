"""Poor man's pyflakes (no linter is installed in the image): report names that a function reads as implicit globals
but that no module-level statement, import or builtin defines — the class of bug that only explodes when the line runs.

    python scripts/check_names.py [paths...]      (default: the package, bench.py, scripts/, tests/)
"""
import ast
import builtins
import os
import symtable
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def module_names(tree):
    names = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__package__", "__spec__", "__builtins__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(node.name)
        elif isinstance(node, ast.Import):
            names.update((a.asname or a.name).split(".")[0] for a in node.names)
        elif isinstance(node, ast.ImportFrom):
            names.update(a.asname or a.name for a in node.names)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            names.add(node.id)
        elif isinstance(node, ast.Global):
            names.update(node.names)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            names.add(node.name)
    return names


def check(path):
    src = open(path).read()
    try:
        tree = ast.parse(src)
        table = symtable.symtable(src, path, "exec")
    except SyntaxError as e:
        return [f"{path}: syntax error {e}"]
    if any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree)):
        return []
    known = module_names(tree)
    out = []

    def walk(t):
        for s in t.get_symbols():
            if s.is_referenced() and s.is_global() and not s.is_declared_global() and s.get_name() not in known:
                out.append(f"{path}: '{s.get_name()}' used in {t.get_type()} '{t.get_name()}' (line {t.get_lineno()}) is never defined")
        for c in t.get_children():
            walk(c)
    walk(table)
    return out


def main(argv):
    paths = argv or [os.path.join(ROOT, p) for p in ("nn_distributed_training_b200", "bench.py", "__graft_entry__.py", "scripts", "tests", "experiments")]
    files = []
    for p in paths:
        if os.path.isdir(p):
            for d, _, fs in os.walk(p):
                if "/baseline/" in d or "__pycache__" in d:
                    continue
                files += [os.path.join(d, f) for f in fs if f.endswith(".py")]
        else:
            files.append(p)
    problems = [m for f in sorted(files) for m in check(f)]
    print("\n".join(problems) if problems else f"{len(files)} files: no undefined names")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
