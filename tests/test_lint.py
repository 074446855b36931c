"""CPU: cheap static check for names that are used but never bound (a NameError on a GPU-only code path costs a
GPU round-trip to discover)."""
import ast
import builtins
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _undefined_names(path: Path):
    tree = ast.parse(path.read_text())
    issues = []

    def binds(nodes, into):
        for m in nodes:
            if isinstance(m, ast.Name) and isinstance(m.ctx, (ast.Store, ast.Del)):
                into.add(m.id)
            elif isinstance(m, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
                into.add(m.name)
            elif isinstance(m, (ast.Import, ast.ImportFrom)):
                for al in m.names:
                    into.add((al.asname or al.name).split(".")[0])
            elif isinstance(m, ast.ExceptHandler) and m.name:
                into.add(m.name)
            elif isinstance(m, ast.arg):
                into.add(m.arg)

    module_names = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for n in tree.body:
        binds(ast.walk(n) if not isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)) else [n], module_names)

    def walk_scope(node, enclosing):
        """A name loaded inside `node` must be bound in it, in an enclosing function, or at module level."""
        local = set(enclosing)
        binds(ast.walk(node), local)          # generous: includes names bound in nested defs
        for child in ast.walk(node):
            if isinstance(child, ast.Name) and isinstance(child.ctx, ast.Load) and child.id not in local:
                issues.append((child.lineno, child.id))

    def visit(node, enclosing):
        for child in ast.iter_child_nodes(node):
            if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)):
                local = set(enclosing)
                binds(ast.walk(child), local)
                walk_scope(child, enclosing)
                visit(child, local)
            else:
                visit(child, enclosing)

    visit(tree, module_names)
    return sorted(set(issues))


def test_no_undefined_names_in_python_sources():
    files = [*ROOT.glob("wax_b200/*.py"), ROOT / "bench.py", ROOT / "__graft_entry__.py", *ROOT.glob("scripts/*.py"),
             *ROOT.glob("oracle/*.py")]
    bad = {str(f.relative_to(ROOT)): _undefined_names(f) for f in files}
    bad = {k: v for k, v in bad.items() if v}
    assert not bad, bad
