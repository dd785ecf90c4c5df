"""Records the public surface of the reference package (names, methods, positional parameters) as
tests/golden/api_surface.json.  Run in the build container, where /root/reference exists:
    python tests/golden/make_api_surface.py
tests/test_host_api.py::test_api_surface_covers_the_reference checks physicl_b200 against the file."""
import ast
import json
import os
import warnings

REF = "/root/reference/physicl/"
HERE = os.path.dirname(os.path.abspath(__file__))


def params(fn):
    a = fn.args
    return {"args": [x.arg for x in a.args], "ndefaults": len(a.defaults), "vararg": bool(a.vararg), "kwarg": bool(a.kwarg)}


def main():
    out = {}
    for mod in ("__init__", "light", "newton"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            tree = ast.parse(open(REF + mod + ".py").read())
        d = {}
        for node in tree.body:
            if isinstance(node, ast.ClassDef):
                d[node.name] = {"kind": "class", "bases": [ast.unparse(b) for b in node.bases],
                                "methods": {n.name: params(n) for n in node.body if isinstance(n, ast.FunctionDef)}}
            elif isinstance(node, ast.FunctionDef):
                d[node.name] = dict(params(node), kind="function")
            elif isinstance(node, ast.Assign):
                for t in node.targets:
                    if isinstance(t, ast.Name):
                        d[t.id] = {"kind": "value"}
        out[mod] = d
    with open(os.path.join(HERE, "api_surface.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({m: len(d) for m, d in out.items()})


if __name__ == "__main__":
    main()
