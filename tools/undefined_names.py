"""Poor man's pyflakes (none is installed): names loaded in a function that no enclosing scope defines."""
import ast
import builtins
import sys


def names_defined(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            out.add(n.name)
            if not isinstance(n, ast.ClassDef):
                a = n.args
                for x in a.args + a.kwonlyargs + a.posonlyargs + ([a.vararg] if a.vararg else []) + (
                        [a.kwarg] if a.kwarg else []):
                    out.add(x.arg)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for al in n.names:
                out.add((al.asname or al.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        elif isinstance(n, ast.Lambda):
            for x in n.args.args:
                out.add(x.arg)
    return out


def module_level(tree):
    """Names bound at module level (not inside function bodies)."""
    out = set()
    for n in tree.body:
        if isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            out.add(n.name)
        else:
            out |= names_defined(n)
    return out


def main(path):
    tree = ast.parse(open(path).read())
    top = module_level(tree) | set(dir(builtins)) | {"__file__", "__name__"}
    bad = 0

    def check(node, known):
        nonlocal bad
        for n in ast.walk(node):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in known:
                print("%s:%d: undefined name %r" % (path, n.lineno, n.id))
                bad += 1
    for n in tree.body:
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef)):
            check(n, top | names_defined(n))
        elif isinstance(n, ast.ClassDef):
            for m in n.body:
                check(m, top | names_defined(m) | names_defined(n))
        else:
            check(n, top)
    return bad


if __name__ == "__main__":
    sys.exit(1 if sum(main(p) for p in sys.argv[1:]) else 0)
